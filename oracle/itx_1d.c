/*
 * oracle/itx_1d.c — TEST INFRASTRUCTURE (see oracle_common.h, itx_1d.h).
 * Restates reference src/itx_1d.c (file:line ranges in itx_1d.h).
 *
 * Formulation: two scratch arrays t[] / u[] indexed by butterfly-node number hold the
 * alternating generations of each node; the even half of an N-point DCT is the
 * (N/2)-point DCT run in place on the even-indexed inputs (stride 2s), the odd half is
 * computed into T[N/2..N-1] and the two are merged by  out[i] = e[i] +- T[N-1-i].
 */
#include "itx_1d.h"

#define X(i)   c[(ptrdiff_t)(i) * s]
#define CL(v)  o_clip((v), lo, hi)

static inline int R12(int x, int cx, int y, int cy) { return o_mac2(x, cx, y, cy, 2048u, 12); }
static inline int R11(int x, int cx, int y, int cy) { return o_mac2(x, cx, y, cy, 1024u, 11); }
static inline int M12(int x, int cx) { return (int)((unsigned)x * (unsigned)cx + 2048u) >> 12; }
/* ((a + b) * 181 + 128) >> 8 */
static inline int H181(int a, int b) { return (int)(((unsigned)a + (unsigned)b) * 181u + 128u) >> 8; }

/* stage "pairs": t[b]=u[b]+u[b+1], t[b+1]=u[b]-u[b+1], t[b+2]=u[b+3]-u[b+2], t[b+3]=u[b+3]+u[b+2] */
static inline void pair_stage(int *t, const int *u, int from, int to, int lo, int hi) {
    for (int b = from; b < to; b += 4) {
        t[b + 0] = CL(u[b + 0] + u[b + 1]);
        t[b + 1] = CL(u[b + 0] - u[b + 1]);
        t[b + 2] = CL(u[b + 3] - u[b + 2]);
        t[b + 3] = CL(u[b + 3] + u[b + 2]);
    }
}

/* merge even half (already in place at even positions) with odd half T[n/2..n-1] */
static inline void merge(int32_t *c, ptrdiff_t s, const int *T, int n, int lo, int hi) {
    int e[32];
    for (int i = 0; i < n / 2; i++) e[i] = X(2 * i);
    for (int i = 0; i < n / 2; i++) {
        X(i)         = CL(e[i] + T[n - 1 - i]);
        X(n - 1 - i) = CL(e[i] - T[n - 1 - i]);
    }
}

/* ---- DCT4 : reference src/itx_1d.c:65-93 ---- */
static void dct4_i(int32_t *c, ptrdiff_t s, int lo, int hi, int tx64) {
    const int in0 = X(0), in1 = X(1);
    int t0, t1, t2, t3;
    if (tx64) {
        t0 = t1 = H181(in0, 0);
        t2 = M12(in1, 1567);
        t3 = M12(in1, 3784);
    } else {
        const int in2 = X(2), in3 = X(3);
        t0 = H181(in0, in2);
        t1 = H181(in0, -in2);
        t2 = R12(in1, 1567, in3, 4096 - 3784) - in3;
        t3 = R12(in1, 3784 - 4096, in3, 1567) + in1;
    }
    X(0) = CL(t0 + t3);
    X(1) = CL(t1 + t2);
    X(2) = CL(t1 - t2);
    X(3) = CL(t0 - t3);
}

/* ---- DCT8 : reference src/itx_1d.c:101-151 ---- */
static void dct8_i(int32_t *c, ptrdiff_t s, int lo, int hi, int tx64) {
    int t[8], u[8];
    dct4_i(c, s * 2, lo, hi, tx64);
    const int in1 = X(1), in3 = X(3);
    if (tx64) {
        u[4] = M12(in1, 799);
        u[5] = M12(in3, -2276);
        u[6] = M12(in3, 3406);
        u[7] = M12(in1, 4017);
    } else {
        const int in5 = X(5), in7 = X(7);
        u[4] = R12(in1, 799, in7, 4096 - 4017) - in7;
        u[5] = R11(in5, 1703, in3, -1138);
        u[6] = R11(in5, 1138, in3, 1703);
        u[7] = R12(in1, 4017 - 4096, in7, 799) + in1;
    }
    pair_stage(t, u, 4, 8, lo, hi);   /* t4, t5(a), t6(a), t7 */
    u[5] = H181(t[6], -t[5]);
    u[6] = H181(t[6], t[5]);
    int T[8];
    T[4] = t[4]; T[5] = u[5]; T[6] = u[6]; T[7] = t[7];
    merge(c, s, T, 8, lo, hi);
}

/* ---- DCT16 : reference src/itx_1d.c:159-251 ---- */
static void dct16_i(int32_t *c, ptrdiff_t s, int lo, int hi, int tx64) {
    int t[16], u[16];
    dct8_i(c, s * 2, lo, hi, tx64);
    const int in1 = X(1), in3 = X(3), in5 = X(5), in7 = X(7);
    if (tx64) {
        u[8]  = M12(in1, 401);   u[9]  = M12(in7, -2598);
        u[10] = M12(in5, 1931);  u[11] = M12(in3, -1189);
        u[12] = M12(in3, 3920);  u[13] = M12(in5, 3612);
        u[14] = M12(in7, 3166);  u[15] = M12(in1, 4076);
    } else {
        const int in9 = X(9), in11 = X(11), in13 = X(13), in15 = X(15);
        u[8]  = R12(in1, 401, in15, 4096 - 4076) - in15;
        u[9]  = R11(in9, 1583, in7, -1299);
        u[10] = R12(in5, 1931, in11, 4096 - 3612) - in11;
        u[11] = R12(in13, 3920 - 4096, in3, -1189) + in13;
        u[12] = R12(in13, 1189, in3, 3920 - 4096) + in3;
        u[13] = R12(in5, 3612 - 4096, in11, 1931) + in5;
        u[14] = R11(in9, 1299, in7, 1583);
        u[15] = R12(in1, 4076 - 4096, in15, 401) + in1;
    }
    pair_stage(t, u, 8, 16, lo, hi);

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
    merge(c, s, T, 16, lo, hi);
}

/* ---- DCT32 : reference src/itx_1d.c:259-429 ---- */
static void dct32_i(int32_t *c, ptrdiff_t s, int lo, int hi, int tx64) {
    int t[32], u[32];
    dct16_i(c, s * 2, lo, hi, tx64);
    const int in1 = X(1), in3 = X(3), in5 = X(5), in7 = X(7);
    const int in9 = X(9), in11 = X(11), in13 = X(13), in15 = X(15);
    if (tx64) {
        u[16] = M12(in1, 201);    u[17] = M12(in15, -2751);
        u[18] = M12(in9, 1751);   u[19] = M12(in7, -1380);
        u[20] = M12(in5, 995);    u[21] = M12(in11, -2106);
        u[22] = M12(in13, 2440);  u[23] = M12(in3, -601);
        u[24] = M12(in3, 4052);   u[25] = M12(in13, 3290);
        u[26] = M12(in11, 3513);  u[27] = M12(in5, 3973);
        u[28] = M12(in7, 3857);   u[29] = M12(in9, 3703);
        u[30] = M12(in15, 3035);  u[31] = M12(in1, 4091);
    } else {
        const int in17 = X(17), in19 = X(19), in21 = X(21), in23 = X(23);
        const int in25 = X(25), in27 = X(27), in29 = X(29), in31 = X(31);
        u[16] = R12(in1, 201, in31, 4096 - 4091) - in31;
        u[17] = R12(in17, 3035 - 4096, in15, -2751) + in17;
        u[18] = R12(in9, 1751, in23, 4096 - 3703) - in23;
        u[19] = R12(in25, 3857 - 4096, in7, -1380) + in25;
        u[20] = R12(in5, 995, in27, 4096 - 3973) - in27;
        u[21] = R12(in21, 3513 - 4096, in11, -2106) + in21;
        u[22] = R11(in13, 1220, in19, -1645);
        u[23] = R12(in29, 4052 - 4096, in3, -601) + in29;
        u[24] = R12(in29, 601, in3, 4052 - 4096) + in3;
        u[25] = R11(in13, 1645, in19, 1220);
        u[26] = R12(in21, 2106, in11, 3513 - 4096) + in11;
        u[27] = R12(in5, 3973 - 4096, in27, 995) + in5;
        u[28] = R12(in25, 1380, in7, 3857 - 4096) + in7;
        u[29] = R12(in9, 3703 - 4096, in23, 1751) + in9;
        u[30] = R12(in17, 2751, in15, 3035 - 4096) + in15;
        u[31] = R12(in1, 4091 - 4096, in31, 201) + in1;
    }
    pair_stage(t, u, 16, 32, lo, hi);

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
    merge(c, s, T, 32, lo, hi);
}

/* ---- DCT64 : reference src/itx_1d.c:437-781 (only the low 32 inputs are coded) ---- */
static void dct64_i(int32_t *c, ptrdiff_t s, int lo, int hi) {
    int t[64], u[64];
    dct32_i(c, s * 2, lo, hi, 1);
    /* (input index, multiplier) for nodes 32..63 */
    static const int16_t k_in[32]  = {  1, 31, 17, 15,  9, 23, 25,  7,  5, 27, 21, 11, 13, 19, 29,  3,
                                        3, 29, 19, 13, 11, 21, 27,  5,  7, 25, 23,  9, 15, 17, 31,  1 };
    static const int16_t k_mul[32] = { 101, -2824, 1660, -1474, 897, -2191, 2359, -700,
                                       501, -2520, 2019, -1092, 1285, -1842, 2675, -301,
                                       4085, 3102, 3659, 3889, 3948, 3564, 3229, 4065,
                                       4036, 3349, 3461, 3996, 3822, 3745, 2967, 4095 };
    for (int i = 0; i < 32; i++) u[32 + i] = M12(X(k_in[i]), k_mul[i]);
    pair_stage(t, u, 32, 64, lo, hi);

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

    u[32] = CL(t[32] + t[35]);  u[35] = CL(t[32] - t[35]);
    t[33] = CL(u[33] + u[34]);  t[34] = CL(u[33] - u[34]);
    u[36] = CL(t[39] - t[36]);  u[39] = CL(t[39] + t[36]);
    t[37] = CL(u[38] - u[37]);  t[38] = CL(u[38] + u[37]);
    u[40] = CL(t[40] + t[43]);  u[43] = CL(t[40] - t[43]);
    t[41] = CL(u[41] + u[42]);  t[42] = CL(u[41] - u[42]);
    u[44] = CL(t[47] - t[44]);  u[47] = CL(t[47] + t[44]);
    t[45] = CL(u[46] - u[45]);  t[46] = CL(u[46] + u[45]);
    u[48] = CL(t[48] + t[51]);  u[51] = CL(t[48] - t[51]);
    t[49] = CL(u[49] + u[50]);  t[50] = CL(u[49] - u[50]);
    u[52] = CL(t[55] - t[52]);  u[55] = CL(t[55] + t[52]);
    t[53] = CL(u[54] - u[53]);  t[54] = CL(u[54] + u[53]);
    u[56] = CL(t[56] + t[59]);  u[59] = CL(t[56] - t[59]);
    t[57] = CL(u[57] + u[58]);  t[58] = CL(u[57] - u[58]);
    u[60] = CL(t[63] - t[60]);  u[63] = CL(t[63] + t[60]);
    t[61] = CL(u[62] - u[61]);  t[62] = CL(u[62] + u[61]);

    {
        const int a34 = R12(t[34], 4096 - 4017, t[61], 799) - t[34];
        const int b35 = R12(u[35], 4096 - 4017, u[60], 799) - u[35];
        const int b36 = R12(u[36], -799, u[59], 4096 - 4017) - u[59];
        const int a37 = R12(t[37], -799, t[58], 4096 - 4017) - t[58];
        const int a42 = R11(t[42], -1138, t[53], 1703);
        const int b43 = R11(u[43], -1138, u[52], 1703);
        const int b44 = R11(u[44], -1703, u[51], -1138);
        const int a45 = R11(t[45], -1703, t[50], -1138);
        const int a50 = R11(t[45], -1138, t[50], 1703);
        const int b51 = R11(u[44], -1138, u[51], 1703);
        const int b52 = R11(u[43], 1703, u[52], 1138);
        const int a53 = R11(t[42], 1703, t[53], 1138);
        const int a58 = R12(t[37], 4096 - 4017, t[58], 799) - t[37];
        const int b59 = R12(u[36], 4096 - 4017, u[59], 799) - u[36];
        const int b60 = R12(u[35], 799, u[60], 4017 - 4096) + u[60];
        const int a61 = R12(t[34], 799, t[61], 4017 - 4096) + t[61];
        u[34] = a34; t[35] = b35; t[36] = b36; u[37] = a37;
        u[42] = a42; t[43] = b43; t[44] = b44; u[45] = a45;
        u[50] = a50; t[51] = b51; t[52] = b52; u[53] = a53;
        u[58] = a58; t[59] = b59; t[60] = b60; u[61] = a61;
    }

    {
        int n[64];
        n[32] = CL(u[32] + u[39]);  n[39] = CL(u[32] - u[39]);
        n[33] = CL(t[33] + t[38]);  n[38] = CL(t[33] - t[38]);
        n[34] = CL(u[34] + u[37]);  n[37] = CL(u[34] - u[37]);
        n[35] = CL(t[35] + t[36]);  n[36] = CL(t[35] - t[36]);
        n[40] = CL(u[47] - u[40]);  n[47] = CL(u[47] + u[40]);
        n[41] = CL(t[46] - t[41]);  n[46] = CL(t[46] + t[41]);
        n[42] = CL(u[45] - u[42]);  n[45] = CL(u[45] + u[42]);
        n[43] = CL(t[44] - t[43]);  n[44] = CL(t[44] + t[43]);
        n[48] = CL(u[48] + u[55]);  n[55] = CL(u[48] - u[55]);
        n[49] = CL(t[49] + t[54]);  n[54] = CL(t[49] - t[54]);
        n[50] = CL(u[50] + u[53]);  n[53] = CL(u[50] - u[53]);
        n[51] = CL(t[51] + t[52]);  n[52] = CL(t[51] - t[52]);
        n[56] = CL(u[63] - u[56]);  n[63] = CL(u[63] + u[56]);
        n[57] = CL(t[62] - t[57]);  n[62] = CL(t[62] + t[57]);
        n[58] = CL(u[61] - u[58]);  n[61] = CL(u[61] + u[58]);
        n[59] = CL(t[60] - t[59]);  n[60] = CL(t[60] + t[59]);
        for (int i = 32; i < 64; i++) t[i] = n[i];   /* t[] now holds the whole generation */
    }

    /* rotation by (1567, 3784) across the middle 16+... nodes 36..43 / 52..59 */
    for (int i = 0; i < 4; i++) {
        const int a = 36 + i, b = 59 - i;      /* (36,59) (37,58) (38,57) (39,56) */
        u[a] = R12(t[a], 4096 - 3784, t[b], 1567) - t[a];
        u[b] = R12(t[a], 1567, t[b], 3784 - 4096) + t[b];
        const int p = 40 + i, q = 55 - i;      /* (40,55) (41,54) (42,53) (43,52) */
        u[p] = R12(t[p], -1567, t[q], 4096 - 3784) - t[q];
        u[q] = R12(t[p], 4096 - 3784, t[q], 1567) - t[p];
    }
    for (int i = 32; i < 36; i++) u[i] = t[i];
    for (int i = 44; i < 52; i++) u[i] = t[i];
    for (int i = 60; i < 64; i++) u[i] = t[i];

    for (int i = 0; i < 8; i++) {
        t[32 + i] = CL(u[32 + i] + u[47 - i]);
        t[47 - i] = CL(u[32 + i] - u[47 - i]);
        t[48 + i] = CL(u[63 - i] - u[48 + i]);
        t[63 - i] = CL(u[63 - i] + u[48 + i]);
    }

    int T[64];
    for (int i = 32; i < 40; i++) T[i] = t[i];
    for (int i = 56; i < 64; i++) T[i] = t[i];
    for (int i = 0; i < 8; i++) {
        T[40 + i] = H181(t[55 - i], -t[40 + i]);
        T[55 - i] = H181(t[55 - i], t[40 + i]);
    }
    merge(c, s, T, 64, lo, hi);
}

void oracle_dct4 (int32_t *c, ptrdiff_t s, int lo, int hi) { dct4_i (c, s, lo, hi, 0); }
void oracle_dct8 (int32_t *c, ptrdiff_t s, int lo, int hi) { dct8_i (c, s, lo, hi, 0); }
void oracle_dct16(int32_t *c, ptrdiff_t s, int lo, int hi) { dct16_i(c, s, lo, hi, 0); }
void oracle_dct32(int32_t *c, ptrdiff_t s, int lo, int hi) { dct32_i(c, s, lo, hi, 0); }
void oracle_dct64(int32_t *c, ptrdiff_t s, int lo, int hi) { dct64_i(c, s, lo, hi); }

/* ---- ADST4 : reference src/itx_1d.c:783-802 (no intermediate clipping) ---- */
static void adst4_i(const int32_t *c, ptrdiff_t s, int32_t *o, ptrdiff_t os) {
    const int in0 = X(0), in1 = X(1), in2 = X(2), in3 = X(3);
    const unsigned a0 = in0, a1 = in1, a2 = in2, a3 = in3;
    const int o0 = ((int)(1321u * a0 + (unsigned)(3803 - 4096) * a2 + (unsigned)(2482 - 4096) * a3 +
                          (unsigned)(3344 - 4096) * a1 + 2048u) >> 12) + in2 + in3 + in1;
    const int o1 = ((int)((unsigned)(2482 - 4096) * a0 - 1321u * a2 - (unsigned)(3803 - 4096) * a3 +
                          (unsigned)(3344 - 4096) * a1 + 2048u) >> 12) + in0 - in3 + in1;
    const int o2 = (int)(209u * (a0 - a2 + a3) + 128u) >> 8;
    const int o3 = ((int)((unsigned)(3803 - 4096) * a0 + (unsigned)(2482 - 4096) * a2 - 1321u * a3 -
                          (unsigned)(3344 - 4096) * a1 + 2048u) >> 12) + in0 + in2 - in1;
    o[0 * os] = o0; o[1 * os] = o1; o[2 * os] = o2; o[3 * os] = o3;
}

/* ---- ADST8 : reference src/itx_1d.c:804-851 ---- */
static void adst8_i(const int32_t *c, ptrdiff_t s, int lo, int hi, int32_t *o, ptrdiff_t os) {
    const int in0 = X(0), in1 = X(1), in2 = X(2), in3 = X(3);
    const int in4 = X(4), in5 = X(5), in6 = X(6), in7 = X(7);
    int u[8], t[8];
    u[0] = R12(in7, 4076 - 4096, in0, 401) + in7;
    u[1] = R12(in7, 401, in0, 4096 - 4076) - in0;
    u[2] = R12(in5, 3612 - 4096, in2, 1931) + in5;
    u[3] = R12(in5, 1931, in2, 4096 - 3612) - in2;
    u[4] = R11(in3, 1299, in4, 1583);
    u[5] = R11(in3, 1583, in4, -1299);
    u[6] = R12(in1, 1189, in6, 3920 - 4096) + in6;
    u[7] = R12(in1, 3920 - 4096, in6, -1189) + in1;
    for (int i = 0; i < 4; i++) {
        t[i]     = CL(u[i] + u[i + 4]);
        t[i + 4] = CL(u[i] - u[i + 4]);
    }
    u[4] = R12(t[4], 3784 - 4096, t[5], 1567) + t[4];
    u[5] = R12(t[4], 1567, t[5], 4096 - 3784) - t[5];
    u[6] = R12(t[7], 3784 - 4096, t[6], -1567) + t[7];
    u[7] = R12(t[7], 1567, t[6], 3784 - 4096) + t[6];

    const int o0 =  CL(t[0] + t[2]);
    const int o7 = -CL(t[1] + t[3]);
    const int v2 =  CL(t[0] - t[2]);
    const int v3 =  CL(t[1] - t[3]);
    const int o1 = -CL(u[4] + u[6]);
    const int o6 =  CL(u[5] + u[7]);
    const int v6 =  CL(u[4] - u[6]);
    const int v7 =  CL(u[5] - u[7]);
    o[0 * os] = o0; o[7 * os] = o7; o[1 * os] = o1; o[6 * os] = o6;
    o[3 * os] = -H181(v2, v3);
    o[4 * os] =  H181(v2, -v3);
    o[2 * os] =  H181(v6, v7);
    o[5 * os] = -H181(v6, -v7);
}

/* ---- ADST16 : reference src/itx_1d.c:853-952 ---- */
static void adst16_i(const int32_t *c, ptrdiff_t s, int lo, int hi, int32_t *o, ptrdiff_t os) {
    int in[16], t[16], u[16];
    for (int i = 0; i < 16; i++) in[i] = X(i);
    t[0]  = R12(in[15], 4091 - 4096, in[0], 201) + in[15];
    t[1]  = R12(in[15], 201, in[0], 4096 - 4091) - in[0];
    t[2]  = R12(in[13], 3973 - 4096, in[2], 995) + in[13];
    t[3]  = R12(in[13], 995, in[2], 4096 - 3973) - in[2];
    t[4]  = R12(in[11], 3703 - 4096, in[4], 1751) + in[11];
    t[5]  = R12(in[11], 1751, in[4], 4096 - 3703) - in[4];
    t[6]  = R11(in[9], 1645, in[6], 1220);
    t[7]  = R11(in[9], 1220, in[6], -1645);
    t[8]  = R12(in[7], 2751, in[8], 3035 - 4096) + in[8];
    t[9]  = R12(in[7], 3035 - 4096, in[8], -2751) + in[7];
    t[10] = R12(in[5], 2106, in[10], 3513 - 4096) + in[10];
    t[11] = R12(in[5], 3513 - 4096, in[10], -2106) + in[5];
    t[12] = R12(in[3], 1380, in[12], 3857 - 4096) + in[12];
    t[13] = R12(in[3], 3857 - 4096, in[12], -1380) + in[3];
    t[14] = R12(in[1], 601, in[14], 4052 - 4096) + in[14];
    t[15] = R12(in[1], 4052 - 4096, in[14], -601) + in[1];
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
    for (int i = 0; i < 4; i++) {
        t[i]     = CL(u[i] + u[i + 4]);
        t[i + 4] = CL(u[i] - u[i + 4]);
    }
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

    o[ 0 * os] =  CL(t[0] + t[2]);
    o[15 * os] = -CL(t[1] + t[3]);
    const int a2 = CL(t[0] - t[2]), a3 = CL(t[1] - t[3]);
    o[ 3 * os] = -CL(u[4] + u[6]);
    o[12 * os] =  CL(u[5] + u[7]);
    const int a6 = CL(u[4] - u[6]), a7 = CL(u[5] - u[7]);
    o[ 1 * os] = -CL(u[8] + u[10]);
    o[14 * os] =  CL(u[9] + u[11]);
    const int a10 = CL(u[8] - u[10]), a11 = CL(u[9] - u[11]);
    o[ 2 * os] =  CL(t[12] + t[14]);
    o[13 * os] = -CL(t[13] + t[15]);
    const int a14 = CL(t[12] - t[14]), a15 = CL(t[13] - t[15]);

    o[ 7 * os] = -H181(a2, a3);
    o[ 8 * os] =  H181(a2, -a3);
    o[ 4 * os] =  H181(a6, a7);
    o[11 * os] = -H181(a6, -a7);
    o[ 6 * os] =  H181(a10, a11);
    o[ 9 * os] = -H181(a10, -a11);
    o[ 5 * os] = -H181(a14, a15);
    o[10 * os] =  H181(a14, -a15);
}

/* in-place and flipped (output written back to front) entry points: src/itx_1d.c:954-973 */
void oracle_adst4 (int32_t *c, ptrdiff_t s, int lo, int hi) { (void)lo; (void)hi; adst4_i(c, s, c, s); }
void oracle_adst8 (int32_t *c, ptrdiff_t s, int lo, int hi) { adst8_i (c, s, lo, hi, c, s); }
void oracle_adst16(int32_t *c, ptrdiff_t s, int lo, int hi) { adst16_i(c, s, lo, hi, c, s); }
void oracle_flipadst4 (int32_t *c, ptrdiff_t s, int lo, int hi) { (void)lo; (void)hi; adst4_i(c, s, c + 3 * s, -s); }
void oracle_flipadst8 (int32_t *c, ptrdiff_t s, int lo, int hi) { adst8_i (c, s, lo, hi, c + 7 * s, -s); }
void oracle_flipadst16(int32_t *c, ptrdiff_t s, int lo, int hi) { adst16_i(c, s, lo, hi, c + 15 * s, -s); }

/* ---- identity : reference src/itx_1d.c:983-1017 ---- */
void oracle_identity4(int32_t *c, ptrdiff_t s, int lo, int hi) {
    (void)lo; (void)hi;
    for (int i = 0; i < 4; i++) { const int v = X(i); X(i) = v + M12(v, 1697); }
}
void oracle_identity8(int32_t *c, ptrdiff_t s, int lo, int hi) {
    (void)lo; (void)hi;
    for (int i = 0; i < 8; i++) X(i) = (int)((unsigned)X(i) * 2u);
}
void oracle_identity16(int32_t *c, ptrdiff_t s, int lo, int hi) {
    (void)lo; (void)hi;
    for (int i = 0; i < 16; i++) {
        const int v = X(i);
        X(i) = (int)(2u * (unsigned)v) + ((int)((unsigned)v * 1697u + 1024u) >> 11);
    }
}
void oracle_identity32(int32_t *c, ptrdiff_t s, int lo, int hi) {
    (void)lo; (void)hi;
    for (int i = 0; i < 32; i++) X(i) = (int)((unsigned)X(i) * 4u);
}

/* ---- WHT4 : reference src/itx_1d.c:1066-1081 ---- */
void oracle_wht4(int32_t *c, ptrdiff_t s) {
    const int in0 = X(0), in1 = X(1), in2 = X(2), in3 = X(3);
    const int t0 = in0 + in1;
    const int t2 = in2 - in3;
    const int t4 = (t0 - t2) >> 1;
    const int t3 = t4 - in3;
    const int t1 = t4 - in1;
    X(0) = t0 - t3;
    X(1) = t3;
    X(2) = t1;
    X(3) = t2 + t1;
}
