// 1-D inverse transforms on register-resident vectors (one CUDA thread = one row or column).
//
// Bit-exact counterparts of dav1d's C 1-D kernels (reference src/itx_1d.c:65-1081):
// inverse DCT 4..64, ADST 4..16 (+ flipped), identity 4..32, WHT4. Every rotation is the
// canonical  ((x*cx + y*cy + rnd) >> sh) + adj  with the reduced multipliers the reference
// uses (c, c-4096, or c/2 with sh=11), evaluated in unsigned so out-of-spec inputs wrap
// exactly like the x86 C build. Vectors are int[N] with compile-time indices only, so
// after inlining everything lives in registers.
#pragma once
#include "common.cuh"

namespace b200 {

B200_DEV int R12(int x, int cx, int y, int cy) {
    return (int)((unsigned)x * (unsigned)cx + (unsigned)y * (unsigned)cy + 2048u) >> 12;
}
B200_DEV int R11(int x, int cx, int y, int cy) {
    return (int)((unsigned)x * (unsigned)cx + (unsigned)y * (unsigned)cy + 1024u) >> 11;
}
B200_DEV int M12(int x, int cx) { return (int)((unsigned)x * (unsigned)cx + 2048u) >> 12; }
// ((a + b) * 181 + 128) >> 8
B200_DEV int H181(int a, int b) { return (int)(((unsigned)a + (unsigned)b) * 181u + 128u) >> 8; }

#define CL(v) iclip((v), lo, hi)

// nodes b..b+3 of a "pair" stage: sums/differences with alternating orientation
template <int FROM, int TO, int N>
B200_DEV void pair_stage(int (&t)[N], const int (&u)[N], int lo, int hi) {
#pragma unroll
    for (int b = FROM; b < TO; b += 4) {
        t[b + 0] = CL(u[b + 0] + u[b + 1]);
        t[b + 1] = CL(u[b + 0] - u[b + 1]);
        t[b + 2] = CL(u[b + 3] - u[b + 2]);
        t[b + 3] = CL(u[b + 3] + u[b + 2]);
    }
}

// out[i] = e[i] + T[N-1-i], out[N-1-i] = e[i] - T[N-1-i]
template <int N>
B200_DEV void merge_halves(int (&c)[N], const int (&e)[N / 2], const int (&T)[N], int lo, int hi) {
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
        c[i]         = CL(e[i] + T[N - 1 - i]);
        c[N - 1 - i] = CL(e[i] - T[N - 1 - i]);
    }
}

template <int N, bool TX64> struct IDct;

template <bool TX64> struct IDct<4, TX64> {
    static B200_DEV void run(int (&c)[4], int lo, int hi) {
        int t0, t1, t2, t3;
        if (TX64) {
            t0 = t1 = H181(c[0], 0);
            t2 = M12(c[1], 1567);
            t3 = M12(c[1], 3784);
        } else {
            t0 = H181(c[0], c[2]);
            t1 = H181(c[0], -c[2]);
            t2 = R12(c[1], 1567, c[3], 4096 - 3784) - c[3];
            t3 = R12(c[1], 3784 - 4096, c[3], 1567) + c[1];
        }
        c[0] = CL(t0 + t3);
        c[1] = CL(t1 + t2);
        c[2] = CL(t1 - t2);
        c[3] = CL(t0 - t3);
    }
};

template <bool TX64> struct IDct<8, TX64> {
    static B200_DEV void run(int (&c)[8], int lo, int hi) {
        int e[4] = { c[0], c[2], c[4], c[6] };
        IDct<4, TX64>::run(e, lo, hi);
        int t[8], u[8];
        if (TX64) {
            u[4] = M12(c[1], 799);
            u[5] = M12(c[3], -2276);
            u[6] = M12(c[3], 3406);
            u[7] = M12(c[1], 4017);
        } else {
            u[4] = R12(c[1], 799, c[7], 4096 - 4017) - c[7];
            u[5] = R11(c[5], 1703, c[3], -1138);
            u[6] = R11(c[5], 1138, c[3], 1703);
            u[7] = R12(c[1], 4017 - 4096, c[7], 799) + c[1];
        }
        pair_stage<4, 8>(t, u, lo, hi);
        int T[8];
        T[4] = t[4];
        T[5] = H181(t[6], -t[5]);
        T[6] = H181(t[6], t[5]);
        T[7] = t[7];
        merge_halves<8>(c, e, T, lo, hi);
    }
};

template <bool TX64> struct IDct<16, TX64> {
    static B200_DEV void run(int (&c)[16], int lo, int hi) {
        int e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) e[i] = c[2 * i];
        IDct<8, TX64>::run(e, lo, hi);
        int t[16], u[16];
        if (TX64) {
            u[8]  = M12(c[1], 401);   u[9]  = M12(c[7], -2598);
            u[10] = M12(c[5], 1931);  u[11] = M12(c[3], -1189);
            u[12] = M12(c[3], 3920);  u[13] = M12(c[5], 3612);
            u[14] = M12(c[7], 3166);  u[15] = M12(c[1], 4076);
        } else {
            u[8]  = R12(c[1], 401, c[15], 4096 - 4076) - c[15];
            u[9]  = R11(c[9], 1583, c[7], -1299);
            u[10] = R12(c[5], 1931, c[11], 4096 - 3612) - c[11];
            u[11] = R12(c[13], 3920 - 4096, c[3], -1189) + c[13];
            u[12] = R12(c[13], 1189, c[3], 3920 - 4096) + c[3];
            u[13] = R12(c[5], 3612 - 4096, c[11], 1931) + c[5];
            u[14] = R11(c[9], 1299, c[7], 1583);
            u[15] = R12(c[1], 4076 - 4096, c[15], 401) + c[1];
        }
        pair_stage<8, 16>(t, u, lo, hi);

        u[9]  = R12(t[14], 1567, t[9], 4096 - 3784) - t[9];
        u[14] = R12(t[14], 3784 - 4096, t[9], 1567) + t[14];
        u[10] = R12(t[13], 4096 - 3784, t[10], -1567) - t[13];
        u[13] = R12(t[13], 1567, t[10], 4096 - 3784) - t[10];

        u[8]  = CL(t[8] + t[11]);
        u[11] = CL(t[8] - t[11]);
        t[9]  = CL(u[9] + u[10]);
        t[10] = CL(u[9] - u[10]);
        u[12] = CL(t[15] - t[12]);
        u[15] = CL(t[15] + t[12]);
        t[13] = CL(u[14] - u[13]);
        t[14] = CL(u[14] + u[13]);

        int T[16];
        T[8]  = u[8];
        T[9]  = t[9];
        T[10] = H181(t[13], -t[10]);
        T[11] = H181(u[12], -u[11]);
        T[12] = H181(u[12], u[11]);
        T[13] = H181(t[13], t[10]);
        T[14] = t[14];
        T[15] = u[15];
        merge_halves<16>(c, e, T, lo, hi);
    }
};

template <bool TX64> struct IDct<32, TX64> {
    static B200_DEV void run(int (&c)[32], int lo, int hi) {
        int e[16];
#pragma unroll
        for (int i = 0; i < 16; i++) e[i] = c[2 * i];
        IDct<16, TX64>::run(e, lo, hi);
        int t[32], u[32];
        if (TX64) {
            u[16] = M12(c[1], 201);    u[17] = M12(c[15], -2751);
            u[18] = M12(c[9], 1751);   u[19] = M12(c[7], -1380);
            u[20] = M12(c[5], 995);    u[21] = M12(c[11], -2106);
            u[22] = M12(c[13], 2440);  u[23] = M12(c[3], -601);
            u[24] = M12(c[3], 4052);   u[25] = M12(c[13], 3290);
            u[26] = M12(c[11], 3513);  u[27] = M12(c[5], 3973);
            u[28] = M12(c[7], 3857);   u[29] = M12(c[9], 3703);
            u[30] = M12(c[15], 3035);  u[31] = M12(c[1], 4091);
        } else {
            u[16] = R12(c[1], 201, c[31], 4096 - 4091) - c[31];
            u[17] = R12(c[17], 3035 - 4096, c[15], -2751) + c[17];
            u[18] = R12(c[9], 1751, c[23], 4096 - 3703) - c[23];
            u[19] = R12(c[25], 3857 - 4096, c[7], -1380) + c[25];
            u[20] = R12(c[5], 995, c[27], 4096 - 3973) - c[27];
            u[21] = R12(c[21], 3513 - 4096, c[11], -2106) + c[21];
            u[22] = R11(c[13], 1220, c[19], -1645);
            u[23] = R12(c[29], 4052 - 4096, c[3], -601) + c[29];
            u[24] = R12(c[29], 601, c[3], 4052 - 4096) + c[3];
            u[25] = R11(c[13], 1645, c[19], 1220);
            u[26] = R12(c[21], 2106, c[11], 3513 - 4096) + c[11];
            u[27] = R12(c[5], 3973 - 4096, c[27], 995) + c[5];
            u[28] = R12(c[25], 1380, c[7], 3857 - 4096) + c[7];
            u[29] = R12(c[9], 3703 - 4096, c[23], 1751) + c[9];
            u[30] = R12(c[17], 2751, c[15], 3035 - 4096) + c[15];
            u[31] = R12(c[1], 4091 - 4096, c[31], 201) + c[1];
        }
        pair_stage<16, 32>(t, u, lo, hi);

        u[17] = R12(t[30], 799, t[17], 4096 - 4017) - t[17];
        u[30] = R12(t[30], 4017 - 4096, t[17], 799) + t[30];
        u[18] = R12(t[29], 4096 - 4017, t[18], -799) - t[29];
        u[29] = R12(t[29], 799, t[18], 4096 - 4017) - t[18];
        u[21] = R11(t[26], 1703, t[21], -1138);
        u[26] = R11(t[26], 1138, t[21], 1703);
        u[22] = R11(t[25], -1138, t[22], -1703);
        u[25] = R11(t[25], 1703, t[22], -1138);

        u[16] = CL(t[16] + t[19]);
        u[19] = CL(t[16] - t[19]);
        t[17] = CL(u[17] + u[18]);
        t[18] = CL(u[17] - u[18]);
        u[20] = CL(t[23] - t[20]);
        u[23] = CL(t[23] + t[20]);
        t[21] = CL(u[22] - u[21]);
        t[22] = CL(u[22] + u[21]);
        u[24] = CL(t[24] + t[27]);
        u[27] = CL(t[24] - t[27]);
        t[25] = CL(u[25] + u[26]);
        t[26] = CL(u[25] - u[26]);
        u[28] = CL(t[31] - t[28]);
        u[31] = CL(t[31] + t[28]);
        t[29] = CL(u[30] - u[29]);
        t[30] = CL(u[30] + u[29]);

        u[18] = R12(t[29], 1567, t[18], 4096 - 3784) - t[18];
        u[29] = R12(t[29], 3784 - 4096, t[18], 1567) + t[29];
        t[19] = R12(u[28], 1567, u[19], 4096 - 3784) - u[19];
        t[28] = R12(u[28], 3784 - 4096, u[19], 1567) + u[28];
        t[20] = R12(u[27], 4096 - 3784, u[20], -1567) - u[27];
        t[27] = R12(u[27], 1567, u[20], 4096 - 3784) - u[20];
        u[21] = R12(t[26], 4096 - 3784, t[21], -1567) - t[26];
        u[26] = R12(t[26], 1567, t[21], 4096 - 3784) - t[21];

        t[16] = CL(u[16] + u[23]);
        t[23] = CL(u[16] - u[23]);
        u[17] = CL(t[17] + t[22]);
        u[22] = CL(t[17] - t[22]);
        t[18] = CL(u[18] + u[21]);
        t[21] = CL(u[18] - u[21]);
        u[19] = CL(t[19] + t[20]);
        u[20] = CL(t[19] - t[20]);
        t[24] = CL(u[31] - u[24]);
        t[31] = CL(u[31] + u[24]);
        u[25] = CL(t[30] - t[25]);
        u[30] = CL(t[30] + t[25]);
        t[26] = CL(u[29] - u[26]);
        t[29] = CL(u[29] + u[26]);
        u[27] = CL(t[28] - t[27]);
        u[28] = CL(t[28] + t[27]);

        int T[32];
        T[16] = t[16];
        T[17] = u[17];
        T[18] = t[18];
        T[19] = u[19];
        T[20] = H181(u[27], -u[20]);
        T[21] = H181(t[26], -t[21]);
        T[22] = H181(u[25], -u[22]);
        T[23] = H181(t[24], -t[23]);
        T[24] = H181(t[24], t[23]);
        T[25] = H181(u[25], u[22]);
        T[26] = H181(t[26], t[21]);
        T[27] = H181(u[27], u[20]);
        T[28] = u[28];
        T[29] = t[29];
        T[30] = u[30];
        T[31] = t[31];
        merge_halves<32>(c, e, T, lo, hi);
    }
};

// 64-point: only the low 32 inputs are ever coded (c[32..63] ignored on input)
B200_DEV void idct64(int (&c)[64], int lo, int hi) {
    int e[32];
#pragma unroll
    for (int i = 0; i < 32; i++) e[i] = c[2 * i];
    IDct<32, true>::run(e, lo, hi);
    int t[64], u[64];
    constexpr int k_in[32]  = {  1, 31, 17, 15,  9, 23, 25,  7,  5, 27, 21, 11, 13, 19, 29,  3,
                                 3, 29, 19, 13, 11, 21, 27,  5,  7, 25, 23,  9, 15, 17, 31,  1 };
    constexpr int k_mul[32] = { 101, -2824, 1660, -1474, 897, -2191, 2359, -700,
                                501, -2520, 2019, -1092, 1285, -1842, 2675, -301,
                                4085, 3102, 3659, 3889, 3948, 3564, 3229, 4065,
                                4036, 3349, 3461, 3996, 3822, 3745, 2967, 4095 };
#pragma unroll
    for (int i = 0; i < 32; i++) u[32 + i] = M12(c[k_in[i]], k_mul[i]);
    pair_stage<32, 64>(t, u, lo, hi);

    u[33] = R12(t[33], 4096 - 4076, t[62], 401) - t[33];
    u[34] = R12(t[34], -401, t[61], 4096 - 4076) - t[61];
    u[37] = R11(t[37], -1299, t[58], 1583);
    u[38] = R11(t[38], -1583, t[57], -1299);
    u[41] = R12(t[41], 4096 - 3612, t[54], 1931) - t[41];
    u[42] = R12(t[42], -1931, t[53], 4096 - 3612) - t[53];
    u[45] = R12(t[45], -1189, t[50], 3920 - 4096) + t[50];
    u[46] = R12(t[46], 4096 - 3920, t[49], -1189) - t[46];
    u[49] = R12(t[46], -1189, t[49], 3920 - 4096) + t[49];
    u[50] = R12(t[45], 3920 - 4096, t[50], 1189) + t[45];
    u[53] = R12(t[42], 4096 - 3612, t[53], 1931) - t[42];
    u[54] = R12(t[41], 1931, t[54], 3612 - 4096) + t[54];
    u[57] = R11(t[38], -1299, t[57], 1583);
    u[58] = R11(t[37], 1583, t[58], 1299);
    u[61] = R12(t[34], 4096 - 4076, t[61], 401) - t[34];
    u[62] = R12(t[33], 401, t[62], 4076 - 4096) + t[62];

    // generation 4: outer pairs come from t, inner pairs from u
    int g[64];
#pragma unroll
    for (int b = 32; b < 64; b += 16) {
        g[b + 0]  = CL(t[b + 0] + t[b + 3]);   g[b + 3]  = CL(t[b + 0] - t[b + 3]);
        g[b + 1]  = CL(u[b + 1] + u[b + 2]);   g[b + 2]  = CL(u[b + 1] - u[b + 2]);
        g[b + 4]  = CL(t[b + 7] - t[b + 4]);   g[b + 7]  = CL(t[b + 7] + t[b + 4]);
        g[b + 5]  = CL(u[b + 6] - u[b + 5]);   g[b + 6]  = CL(u[b + 6] + u[b + 5]);
        g[b + 8]  = CL(t[b + 8] + t[b + 11]);  g[b + 11] = CL(t[b + 8] - t[b + 11]);
        g[b + 9]  = CL(u[b + 9] + u[b + 10]);  g[b + 10] = CL(u[b + 9] - u[b + 10]);
        g[b + 12] = CL(t[b + 15] - t[b + 12]); g[b + 15] = CL(t[b + 15] + t[b + 12]);
        g[b + 13] = CL(u[b + 14] - u[b + 13]); g[b + 14] = CL(u[b + 14] + u[b + 13]);
    }

    // generation 5: rotations by (799,4017) on nodes 34..37/58..61 and (1138,1703)/2 on 42..45/50..53
    int r[64];
#pragma unroll
    for (int i = 32; i < 64; i++) r[i] = g[i];
    r[34] = R12(g[34], 4096 - 4017, g[61], 799) - g[34];
    r[35] = R12(g[35], 4096 - 4017, g[60], 799) - g[35];
    r[36] = R12(g[36], -799, g[59], 4096 - 4017) - g[59];
    r[37] = R12(g[37], -799, g[58], 4096 - 4017) - g[58];
    r[42] = R11(g[42], -1138, g[53], 1703);
    r[43] = R11(g[43], -1138, g[52], 1703);
    r[44] = R11(g[44], -1703, g[51], -1138);
    r[45] = R11(g[45], -1703, g[50], -1138);
    r[50] = R11(g[45], -1138, g[50], 1703);
    r[51] = R11(g[44], -1138, g[51], 1703);
    r[52] = R11(g[43], 1703, g[52], 1138);
    r[53] = R11(g[42], 1703, g[53], 1138);
    r[58] = R12(g[37], 4096 - 4017, g[58], 799) - g[37];
    r[59] = R12(g[36], 4096 - 4017, g[59], 799) - g[36];
    r[60] = R12(g[35], 799, g[60], 4017 - 4096) + g[60];
    r[61] = R12(g[34], 799, g[61], 4017 - 4096) + g[61];

    // generation 6: butterflies over groups of 8
#pragma unroll
    for (int i = 0; i < 4; i++) {
        t[32 + i] = CL(r[32 + i] + r[39 - i]);  t[39 - i] = CL(r[32 + i] - r[39 - i]);
        t[40 + i] = CL(r[47 - i] - r[40 + i]);  t[47 - i] = CL(r[47 - i] + r[40 + i]);
        t[48 + i] = CL(r[48 + i] + r[55 - i]);  t[55 - i] = CL(r[48 + i] - r[55 - i]);
        t[56 + i] = CL(r[63 - i] - r[56 + i]);  t[63 - i] = CL(r[63 - i] + r[56 + i]);
    }

    // generation 7: rotations by (1567,3784) on nodes 36..43 / 52..59
#pragma unroll
    for (int i = 32; i < 64; i++) u[i] = t[i];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int a = 36 + i, b = 59 - i;
        u[a] = R12(t[a], 4096 - 3784, t[b], 1567) - t[a];
        u[b] = R12(t[a], 1567, t[b], 3784 - 4096) + t[b];
        const int p = 40 + i, q = 55 - i;
        u[p] = R12(t[p], -1567, t[q], 4096 - 3784) - t[q];
        u[q] = R12(t[p], 4096 - 3784, t[q], 1567) - t[p];
    }

    // generation 8: butterflies over groups of 16
#pragma unroll
    for (int i = 0; i < 8; i++) {
        t[32 + i] = CL(u[32 + i] + u[47 - i]);
        t[47 - i] = CL(u[32 + i] - u[47 - i]);
        t[48 + i] = CL(u[63 - i] - u[48 + i]);
        t[63 - i] = CL(u[63 - i] + u[48 + i]);
    }

    int T[64];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        T[32 + i] = t[32 + i];
        T[56 + i] = t[56 + i];
        T[40 + i] = H181(t[55 - i], -t[40 + i]);
        T[55 - i] = H181(t[55 - i], t[40 + i]);
    }
    merge_halves<64>(c, e, T, lo, hi);
}

template <int N> B200_DEV void idct(int (&c)[N], int lo, int hi) {
    if constexpr (N == 64) idct64(c, lo, hi);
    else IDct<N, false>::run(c, lo, hi);
}

// ---- ADST (FLIP writes the outputs back to front) ----
template <bool FLIP> B200_DEV void iadst4(int (&c)[4]) {
    const int in0 = c[0], in1 = c[1], in2 = c[2], in3 = c[3];
    const unsigned a0 = in0, a1 = in1, a2 = in2, a3 = in3;
    const int o0 = ((int)(1321u * a0 + (unsigned)(3803 - 4096) * a2 + (unsigned)(2482 - 4096) * a3 +
                          (unsigned)(3344 - 4096) * a1 + 2048u) >> 12) + in2 + in3 + in1;
    const int o1 = ((int)((unsigned)(2482 - 4096) * a0 - 1321u * a2 - (unsigned)(3803 - 4096) * a3 +
                          (unsigned)(3344 - 4096) * a1 + 2048u) >> 12) + in0 - in3 + in1;
    const int o2 = (int)(209u * (a0 - a2 + a3) + 128u) >> 8;
    const int o3 = ((int)((unsigned)(3803 - 4096) * a0 + (unsigned)(2482 - 4096) * a2 - 1321u * a3 -
                          (unsigned)(3344 - 4096) * a1 + 2048u) >> 12) + in0 + in2 - in1;
    c[FLIP ? 3 : 0] = o0; c[FLIP ? 2 : 1] = o1; c[FLIP ? 1 : 2] = o2; c[FLIP ? 0 : 3] = o3;
}

template <bool FLIP> B200_DEV void iadst8(int (&c)[8], int lo, int hi) {
    int u[8], t[8];
    u[0] = R12(c[7], 4076 - 4096, c[0], 401) + c[7];
    u[1] = R12(c[7], 401, c[0], 4096 - 4076) - c[0];
    u[2] = R12(c[5], 3612 - 4096, c[2], 1931) + c[5];
    u[3] = R12(c[5], 1931, c[2], 4096 - 3612) - c[2];
    u[4] = R11(c[3], 1299, c[4], 1583);
    u[5] = R11(c[3], 1583, c[4], -1299);
    u[6] = R12(c[1], 1189, c[6], 3920 - 4096) + c[6];
    u[7] = R12(c[1], 3920 - 4096, c[6], -1189) + c[1];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        t[i]     = CL(u[i] + u[i + 4]);
        t[i + 4] = CL(u[i] - u[i + 4]);
    }
    u[4] = R12(t[4], 3784 - 4096, t[5], 1567) + t[4];
    u[5] = R12(t[4], 1567, t[5], 4096 - 3784) - t[5];
    u[6] = R12(t[7], 3784 - 4096, t[6], -1567) + t[7];
    u[7] = R12(t[7], 1567, t[6], 3784 - 4096) + t[6];

    int o[8];
    o[0] =  CL(t[0] + t[2]);
    o[7] = -CL(t[1] + t[3]);
    const int v2 = CL(t[0] - t[2]), v3 = CL(t[1] - t[3]);
    o[1] = -CL(u[4] + u[6]);
    o[6] =  CL(u[5] + u[7]);
    const int v6 = CL(u[4] - u[6]), v7 = CL(u[5] - u[7]);
    o[3] = -H181(v2, v3);
    o[4] =  H181(v2, -v3);
    o[2] =  H181(v6, v7);
    o[5] = -H181(v6, -v7);
#pragma unroll
    for (int i = 0; i < 8; i++) c[FLIP ? 7 - i : i] = o[i];
}

template <bool FLIP> B200_DEV void iadst16(int (&c)[16], int lo, int hi) {
    int t[16], u[16];
    t[0]  = R12(c[15], 4091 - 4096, c[0], 201) + c[15];
    t[1]  = R12(c[15], 201, c[0], 4096 - 4091) - c[0];
    t[2]  = R12(c[13], 3973 - 4096, c[2], 995) + c[13];
    t[3]  = R12(c[13], 995, c[2], 4096 - 3973) - c[2];
    t[4]  = R12(c[11], 3703 - 4096, c[4], 1751) + c[11];
    t[5]  = R12(c[11], 1751, c[4], 4096 - 3703) - c[4];
    t[6]  = R11(c[9], 1645, c[6], 1220);
    t[7]  = R11(c[9], 1220, c[6], -1645);
    t[8]  = R12(c[7], 2751, c[8], 3035 - 4096) + c[8];
    t[9]  = R12(c[7], 3035 - 4096, c[8], -2751) + c[7];
    t[10] = R12(c[5], 2106, c[10], 3513 - 4096) + c[10];
    t[11] = R12(c[5], 3513 - 4096, c[10], -2106) + c[5];
    t[12] = R12(c[3], 1380, c[12], 3857 - 4096) + c[12];
    t[13] = R12(c[3], 3857 - 4096, c[12], -1380) + c[3];
    t[14] = R12(c[1], 601, c[14], 4052 - 4096) + c[14];
    t[15] = R12(c[1], 4052 - 4096, c[14], -601) + c[1];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u[i]     = CL(t[i] + t[i + 8]);
        u[i + 8] = CL(t[i] - t[i + 8]);
    }
    t[8]  = R12(u[8], 4017 - 4096, u[9], 799) + u[8];
    t[9]  = R12(u[8], 799, u[9], 4096 - 4017) - u[9];
    t[10] = R12(u[10], 2276, u[11], 3406 - 4096) + u[11];
    t[11] = R12(u[10], 3406 - 4096, u[11], -2276) + u[10];
    t[12] = R12(u[13], 4017 - 4096, u[12], -799) + u[13];
    t[13] = R12(u[13], 799, u[12], 4017 - 4096) + u[12];
    t[14] = R12(u[15], 2276, u[14], 4096 - 3406) - u[14];
    t[15] = R12(u[15], 3406 - 4096, u[14], 2276) + u[15];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        t[i]     = CL(u[i] + u[i + 4]);
        t[i + 4] = CL(u[i] - u[i + 4]);
    }
#pragma unroll
    for (int i = 8; i < 12; i++) {
        u[i]     = CL(t[i] + t[i + 4]);
        u[i + 4] = CL(t[i] - t[i + 4]);
    }
    u[4]  = R12(t[4], 3784 - 4096, t[5], 1567) + t[4];
    u[5]  = R12(t[4], 1567, t[5], 4096 - 3784) - t[5];
    u[6]  = R12(t[7], 3784 - 4096, t[6], -1567) + t[7];
    u[7]  = R12(t[7], 1567, t[6], 3784 - 4096) + t[6];
    t[12] = R12(u[12], 3784 - 4096, u[13], 1567) + u[12];
    t[13] = R12(u[12], 1567, u[13], 4096 - 3784) - u[13];
    t[14] = R12(u[15], 3784 - 4096, u[14], -1567) + u[15];
    t[15] = R12(u[15], 1567, u[14], 3784 - 4096) + u[14];

    int o[16];
    o[0]  =  CL(t[0] + t[2]);
    o[15] = -CL(t[1] + t[3]);
    const int a2 = CL(t[0] - t[2]), a3 = CL(t[1] - t[3]);
    o[3]  = -CL(u[4] + u[6]);
    o[12] =  CL(u[5] + u[7]);
    const int a6 = CL(u[4] - u[6]), a7 = CL(u[5] - u[7]);
    o[1]  = -CL(u[8] + u[10]);
    o[14] =  CL(u[9] + u[11]);
    const int a10 = CL(u[8] - u[10]), a11 = CL(u[9] - u[11]);
    o[2]  =  CL(t[12] + t[14]);
    o[13] = -CL(t[13] + t[15]);
    const int a14 = CL(t[12] - t[14]), a15 = CL(t[13] - t[15]);
    o[7]  = -H181(a2, a3);
    o[8]  =  H181(a2, -a3);
    o[4]  =  H181(a6, a7);
    o[11] = -H181(a6, -a7);
    o[6]  =  H181(a10, a11);
    o[9]  = -H181(a10, -a11);
    o[5]  = -H181(a14, a15);
    o[10] =  H181(a14, -a15);
#pragma unroll
    for (int i = 0; i < 16; i++) c[FLIP ? 15 - i : i] = o[i];
}

template <int N, bool FLIP> B200_DEV void iadst(int (&c)[N], int lo, int hi) {
    if constexpr (N == 4) iadst4<FLIP>(c);
    else if constexpr (N == 8) iadst8<FLIP>(c, lo, hi);
    else iadst16<FLIP>(c, lo, hi);
}

// ---- identity ----
template <int N> B200_DEV void iidentity(int (&c)[N]) {
#pragma unroll
    for (int i = 0; i < N; i++) {
        const int v = c[i];
        if (N == 4)       c[i] = v + M12(v, 1697);
        else if (N == 8)  c[i] = (int)((unsigned)v * 2u);
        else if (N == 16) c[i] = (int)(2u * (unsigned)v) + ((int)((unsigned)v * 1697u + 1024u) >> 11);
        else              c[i] = (int)((unsigned)v * 4u);
    }
}

// ---- WHT4 (lossless) ----
B200_DEV void iwht4(int (&c)[4]) {
    const int t0 = c[0] + c[1];
    const int t2 = c[2] - c[3];
    const int t4 = (t0 - t2) >> 1;
    const int t3 = t4 - c[3];
    const int t1 = t4 - c[1];
    c[0] = t0 - t3;
    c[1] = t3;
    c[2] = t1;
    c[3] = t2 + t1;
}

#undef CL

enum Tx1d { TX1D_DCT = 0, TX1D_ADST = 1, TX1D_FLIPADST = 2, TX1D_IDENTITY = 3 };

// run the 1-D transform `type` of length N on c (types that do not exist for N are never requested)
template <int N> B200_DEV void tx1d_apply(int (&c)[N], int type, int lo, int hi) {
    if constexpr (N == 64) {
        idct64(c, lo, hi);
    } else if constexpr (N == 32) {
        if (type == TX1D_DCT) idct<N>(c, lo, hi); else iidentity<N>(c);
    } else {
        switch (type) {
        case TX1D_DCT: idct<N>(c, lo, hi); break;
        case TX1D_ADST: iadst<N, false>(c, lo, hi); break;
        case TX1D_FLIPADST: iadst<N, true>(c, lo, hi); break;
        default: iidentity<N>(c); break;
        }
    }
}

}  // namespace b200
