// Device-side bodies of the intra predictors, shared by the batched ipred kernel (ipred.cu) and the fused
// intra reconstruction kernel (intra.cu). A group of threads (a CTA of kIpT threads, or one warp) works on one block; the edge array lives in
// shared memory as ints (tl = IpShared::edge + 128, valid for tl[-(w+h) .. w+h]).
#pragma once
#include "host_util.h"
#ifndef B200_TBL
#define B200_TBL __constant__
#endif
#include "tables_gen.h"

namespace b200 {


#ifndef B200_IPT
#define B200_IPT 128
#endif
constexpr int kIpT = B200_IPT;   // threads per block

// who works on a block: a whole CTA of kIpT threads (batched ipred kernel, round-1 intra kernels) or one warp (the
// warp-per-block intra kernel: independent blocks of a wavefront side by side in one CTA, no CTA barriers)
struct IpCta {
    static B200_DEV int tid() { return threadIdx.x; }
    static constexpr int size = kIpT;
    static B200_DEV void sync() { __syncthreads(); }
};
struct IpWarp {
    static B200_DEV int tid() { return threadIdx.x & 31; }
    static constexpr int size = 32;
    static B200_DEV void sync() { __syncwarp(); }
};

B200_DEV int ip_filter_strength(int wh, int angle, int is_sm) {
    if (is_sm) {
        if (wh <= 8) return angle >= 64 ? 2 : angle >= 40 ? 1 : 0;
        if (wh <= 16) return angle >= 48 ? 2 : angle >= 20 ? 1 : 0;
        if (wh <= 24) return angle >= 4 ? 3 : 0;
        return 3;
    }
    if (wh <= 8) return angle >= 56 ? 1 : 0;
    if (wh <= 16) return angle >= 40 ? 1 : 0;
    if (wh <= 24) return angle >= 32 ? 3 : angle >= 16 ? 2 : angle >= 8 ? 1 : 0;
    if (wh <= 32) return angle >= 32 ? 3 : angle >= 4 ? 2 : 1;
    return 3;
}
B200_DEV int ip_upsample(int wh, int angle, int is_sm) { return angle < 40 && wh <= (16 >> is_sm); }

static __constant__ uint8_t c_edge_kernel[3][5] = { { 0, 4, 8, 4, 0 }, { 0, 5, 6, 5, 0 }, { 2, 4, 4, 4, 2 } };

// out[i], i in [0, sz): in[clamp(i)] or the 5-tap smoothed value inside [lim_from, lim_to)
template <class G = IpCta>
B200_DEV void ip_edge_filter(int *out, int sz, int lim_from, int lim_to, const int *in, int from, int to, int strength) {
    for (int i = G::tid(); i < sz; i += G::size) {
        if (i < imin(sz, lim_from) || i >= imin(lim_to, sz)) { out[i] = in[iclip(i, from, to - 1)]; continue; }
        int s = 0;
#pragma unroll
        for (int j = 0; j < 5; j++) s += in[iclip(i - 2 + j, from, to - 1)] * c_edge_kernel[strength - 1][j];
        out[i] = (s + 8) >> 4;
    }
}
template <class G = IpCta>
B200_DEV void ip_edge_upsample(int *out, int hsz, const int *in, int from, int to, int bdmax) {
    for (int i = G::tid(); i < hsz; i += G::size) {
        out[i * 2] = in[iclip(i, from, to - 1)];
        if (i < hsz - 1) {
            const int s = -in[iclip(i - 1, from, to - 1)] + 9 * in[iclip(i, from, to - 1)] +
                          9 * in[iclip(i + 1, from, to - 1)] - in[iclip(i + 2, from, to - 1)];
            out[i * 2 + 1] = iclip((s + 8) >> 4, 0, bdmax);
        }
    }
}

B200_DEV int ip_dc(const int *tl, int w, int h, int mode, int bdmax, bool hbd) {
    if (mode == B200_DC_128_PRED) return hbd ? (bdmax + 1) >> 1 : 128;
    unsigned dc;
    if (mode == B200_TOP_DC_PRED) { dc = w >> 1; for (int i = 0; i < w; i++) dc += tl[1 + i]; return (int)(dc >> (31 - __clz(w))); }
    if (mode == B200_LEFT_DC_PRED) { dc = h >> 1; for (int i = 0; i < h; i++) dc += tl[-(1 + i)]; return (int)(dc >> (31 - __clz(h))); }
    dc = (w + h) >> 1;
    for (int i = 0; i < w; i++) dc += tl[1 + i];
    for (int i = 0; i < h; i++) dc += tl[-(1 + i)];
    dc >>= __ffs(w + h) - 1;
    if (w != h) {
        dc *= (w > h * 2 || h > w * 2) ? (hbd ? 0x6667u : 0x3334u) : (hbd ? 0xAAABu : 0x5556u);
        dc >>= hbd ? 17 : 16;
    }
    return (int)dc;
}


struct IpShared {
    int edge[2 * 128 + 1];           // tl = edge + 128
    int aux[64 + 64 + 1 + 64 + 8];
    int tile[32 * 32];               // filter-intra working tile / reductions
    int dc;
};

// cfl_ac: (sub-sampled, padded) luma -> zero-mean int16 ac[w * h]   (reference src/ipred_tmpl.c:657-715)
template <bool HBD, class G = IpCta>
B200_DEV void ipred_cfl_ac_body(IpShared &S, const typename Bd<HBD>::pixel *ypx, int ys, int ssh, int ssv, int w, int h,
                                int w_pad, int h_pad, int16_t *ac)
{
    typedef typename Bd<HBD>::pixel pixel;
    const int tid = G::tid();
    int *const s_tile = S.tile;
    int &s_dc = S.dc;
        int part = 0;
        for (int i = tid; i < w * h; i += G::size) {
            const int y = i / w, x = i - y * w;
            const int sy = imin(y, h - 4 * h_pad - 1), sx = imin(x, w - 4 * w_pad - 1);
            const pixel *p = ypx + (ptrdiff_t)(sy << ssv) * ys + (sx << ssh);
            int s = p[0];
            if (ssh) s += p[1];
            if (ssv) { s += p[ys]; if (ssh) s += p[ys + 1]; }
            s <<= 1 + !ssv + !ssh;
            ac[i] = (int16_t)s;
            part += s;
        }
        s_tile[tid] = part;
        G::sync();
        if (tid == 0) {
            const int log2sz = (__ffs(w) - 1) + (__ffs(h) - 1);
            int sum = (1 << log2sz) >> 1;
            for (int i = 0; i < G::size; i++) sum += s_tile[i];
            s_dc = sum >> log2sz;
        }
        G::sync();
        const int dc = s_dc;
        for (int i = tid; i < w * h; i += G::size) ac[i] = (int16_t)(ac[i] - dc);
}

// cfl_pred: dc of the edges + alpha * ac   (reference src/ipred_tmpl.c:71-84, 717-760)
template <bool HBD, class G = IpCta>
B200_DEV void ipred_cfl_pred_body(IpShared &S, typename Bd<HBD>::pixel *dst, int st, int w, int h, int mode, int alpha,
                                  const int16_t *ac, int bdmax)
{
    typedef typename Bd<HBD>::pixel pixel;
    const int tid = G::tid();
    int *const tl = S.edge + 128;
    int &s_dc = S.dc;
        if (tid == 0) s_dc = ip_dc(tl, w, h, mode, bdmax, HBD);
        G::sync();
        const int dc = s_dc;
        for (int i = tid; i < w * h; i += G::size) {
            const int y = i / w, x = i - y * w;
            const int diff = alpha * ac[i];
            const int m = (iabs(diff) + 32) >> 6;
            dst[(ptrdiff_t)y * st + x] = (pixel)iclip(dc + (diff < 0 ? -m : m), 0, bdmax);
        }
}

// the 14 predictors; `angle` carries dav1d's flags (|512 smooth neighbour, |1024 edge filter enabled), FILTER: index
template <bool HBD, class G = IpCta>
B200_DEV void ipred_pred_body(IpShared &S, typename Bd<HBD>::pixel *dst, int st, int w, int h, int mode, int angle_in,
                              int max_w, int max_h, int bdmax)
{
    typedef typename Bd<HBD>::pixel pixel;
    const int tid = G::tid();
    int *const tl = S.edge + 128;
    int *const s_aux = S.aux;
    int *const s_tile = S.tile;
    int &s_dc = S.dc;
    struct { int max_w, max_h; } b = { max_w, max_h };
    int angle = angle_in;
    switch (mode) {
    case B200_DC_PRED: case B200_TOP_DC_PRED: case B200_LEFT_DC_PRED: case B200_DC_128_PRED: {
        if (tid == 0) s_dc = ip_dc(tl, w, h, mode, bdmax, HBD);
        G::sync();
        const int dc = s_dc;
        for (int i = tid; i < w * h; i += G::size) dst[(ptrdiff_t)(i / w) * st + (i % w)] = (pixel)dc;
        break; }
    case B200_VERT_PRED:
        for (int i = tid; i < w * h; i += G::size) dst[(ptrdiff_t)(i / w) * st + (i % w)] = (pixel)tl[1 + (i % w)];
        break;
    case B200_HOR_PRED:
        for (int i = tid; i < w * h; i += G::size) dst[(ptrdiff_t)(i / w) * st + (i % w)] = (pixel)tl[-(1 + i / w)];
        break;
    case B200_PAETH_PRED:
        for (int i = tid; i < w * h; i += G::size) {
            const int y = i / w, x = i - y * w;
            const int l = tl[-(y + 1)], t = tl[1 + x], c = tl[0], base = l + t - c;
            const int ld = iabs(l - base), td = iabs(t - base), cd = iabs(c - base);
            dst[(ptrdiff_t)y * st + x] = (pixel)(ld <= td && ld <= cd ? l : td <= cd ? t : c);
        }
        break;
    case B200_SMOOTH_PRED: case B200_SMOOTH_V_PRED: case B200_SMOOTH_H_PRED: {
        const int right = tl[w], bottom = tl[-h];
        for (int i = tid; i < w * h; i += G::size) {
            const int y = i / w, x = i - y * w;
            const int wv = b200_sm_weights[h + y], wh = b200_sm_weights[w + x];
            int v;
            if (mode == B200_SMOOTH_PRED) v = (wv * tl[1 + x] + (256 - wv) * bottom + wh * tl[-(1 + y)] + (256 - wh) * right + 256) >> 9;
            else if (mode == B200_SMOOTH_V_PRED) v = (wv * tl[1 + x] + (256 - wv) * bottom + 128) >> 8;
            else v = (wh * tl[-(1 + y)] + (256 - wh) * right + 128) >> 8;
            dst[(ptrdiff_t)y * st + x] = (pixel)v;
        }
        break; }
    case B200_Z1_PRED: {
        const int is_sm = (angle >> 9) & 1, eief = angle >> 10; angle &= 511;
        int dx = b200_dr_intra_derivative[angle >> 1];
        const int up = eief ? ip_upsample(w + h, 90 - angle, is_sm) : 0;
        const int fs = (!up && eief) ? ip_filter_strength(w + h, 90 - angle, is_sm) : 0;
        const int *top; int max_base_x;
        if (up) { ip_edge_upsample<G>(s_aux, w + h, &tl[1], -1, w + imin(w, h), bdmax); top = s_aux; max_base_x = 2 * (w + h) - 2; dx <<= 1; }
        else if (fs) { ip_edge_filter<G>(s_aux, w + h, 0, w + h, &tl[1], -1, w + imin(w, h), fs); top = s_aux; max_base_x = w + h - 1; }
        else { top = &tl[1]; max_base_x = w + imin(w, h) - 1; }
        G::sync();
        const int inc = 1 + up;
        for (int i = tid; i < w * h; i += G::size) {
            const int y = i / w, x = i - y * w;
            const int xpos = dx * (y + 1), frac = xpos & 0x3E, base = (xpos >> 6) + x * inc;
            dst[(ptrdiff_t)y * st + x] = (pixel)(base < max_base_x ? (top[base] * (64 - frac) + top[base + 1] * frac + 32) >> 6 : top[max_base_x]);
        }
        break; }
    case B200_Z2_PRED: {
        const int is_sm = (angle >> 9) & 1, eief = angle >> 10; angle &= 511;
        int dy = b200_dr_intra_derivative[(angle - 90) >> 1], dx = b200_dr_intra_derivative[(180 - angle) >> 1];
        const int up_l = eief ? ip_upsample(w + h, 180 - angle, is_sm) : 0;
        const int up_a = eief ? ip_upsample(w + h, angle - 90, is_sm) : 0;
        int *const e = s_aux + 128;               // e[-2h .. 2w]
        if (up_a) { ip_edge_upsample<G>(e, w + 1, tl, 0, w + 1, bdmax); dx <<= 1; }
        else {
            const int fs = eief ? ip_filter_strength(w + h, angle - 90, is_sm) : 0;
            if (fs) ip_edge_filter<G>(&e[1], w, 0, b.max_w, &tl[1], -1, w, fs);
            else for (int i = tid; i < w; i += G::size) e[1 + i] = tl[1 + i];
        }
        if (up_l) { ip_edge_upsample<G>(&e[-h * 2], h + 1, &tl[-h], 0, h + 1, bdmax); dy <<= 1; }
        else {
            const int fs = eief ? ip_filter_strength(w + h, 180 - angle, is_sm) : 0;
            if (fs) ip_edge_filter<G>(&e[-h], h, h - b.max_h, h, &tl[-h], 0, h + 1, fs);
            else for (int i = tid; i < h; i += G::size) e[-h + i] = tl[-h + i];
        }
        G::sync();
        if (tid == 0) e[0] = tl[0];               // after the upsamplers (which also write e[0])
        G::sync();
        const int inc_x = 1 + up_a;
        const int *left = &e[-(1 + up_l)];
        for (int i = tid; i < w * h; i += G::size) {
            const int y = i / w, x = i - y * w;
            const int xpos = ((1 + up_a) << 6) - dx * (y + 1);
            const int base_x = (xpos >> 6) + x * inc_x, frac_x = xpos & 0x3E;
            int v;
            if (base_x >= 0) v = e[base_x] * (64 - frac_x) + e[base_x + 1] * frac_x;
            else {
                const int ypos = (y << (6 + up_l)) - dy * (x + 1);
                const int by = ypos >> 6, fy = ypos & 0x3E;
                v = left[-by] * (64 - fy) + left[-(by + 1)] * fy;
            }
            dst[(ptrdiff_t)y * st + x] = (pixel)((v + 32) >> 6);
        }
        break; }
    case B200_Z3_PRED: {
        const int is_sm = (angle >> 9) & 1, eief = angle >> 10; angle &= 511;
        int dy = b200_dr_intra_derivative[(270 - angle) >> 1];
        const int up = eief ? ip_upsample(w + h, angle - 180, is_sm) : 0;
        const int fs = (!up && eief) ? ip_filter_strength(w + h, angle - 180, is_sm) : 0;
        const int *left; int max_base_y;
        if (up) { ip_edge_upsample<G>(s_aux, w + h, &tl[-(w + h)], imax(w - h, 0), w + h + 1, bdmax); left = &s_aux[2 * (w + h) - 2]; max_base_y = 2 * (w + h) - 2; dy <<= 1; }
        else if (fs) { ip_edge_filter<G>(s_aux, w + h, 0, w + h, &tl[-(w + h)], imax(w - h, 0), w + h + 1, fs); left = &s_aux[w + h - 1]; max_base_y = w + h - 1; }
        else { left = &tl[-1]; max_base_y = h + imin(w, h) - 1; }
        G::sync();
        const int inc = 1 + up;
        for (int i = tid; i < w * h; i += G::size) {
            const int y = i / w, x = i - y * w;
            const int ypos = dy * (x + 1), frac = ypos & 0x3E, base = (ypos >> 6) + y * inc;
            dst[(ptrdiff_t)y * st + x] = (pixel)(base < max_base_y ? (left[-base] * (64 - frac) + left[-(base + 1)] * frac + 32) >> 6 : left[-max_base_y]);
        }
        break; }
    case B200_FILTER_PRED: {
        const int8_t *flt = b200_filter_intra_taps[angle & 511];
        const int uw = w >> 2, uh = h >> 1;
        for (int d = 0; d < uw + uh - 1; d++) {
            for (int ux = tid; ux < uw; ux += G::size) {
                const int uy = d - ux;
                if (uy < 0 || uy >= uh) continue;
                const int x = ux * 4, y = uy * 2;
                int p[7];
                p[0] = y ? (x ? s_tile[(y - 1) * 32 + x - 1] : tl[-y]) : tl[x];
#pragma unroll
                for (int i = 0; i < 4; i++) p[1 + i] = y ? s_tile[(y - 1) * 32 + x + i] : tl[1 + x + i];
#pragma unroll
                for (int i = 0; i < 2; i++) p[5 + i] = x ? s_tile[(y + i) * 32 + x - 1] : tl[-(1 + y + i)];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int8_t *t = flt + k * 2;    // ARCH_X86 tap layout (reference src/ipred_tmpl.c:537-545)
                    const int acc = t[0] * p[0] + t[1] * p[1] + t[16] * p[2] + t[17] * p[3] + t[32] * p[4] + t[33] * p[5] + t[48] * p[6];
                    const int v = iclip((acc + 8) >> 4, 0, bdmax);
                    s_tile[(y + (k >> 2)) * 32 + x + (k & 3)] = v;
                    dst[(ptrdiff_t)(y + (k >> 2)) * st + x + (k & 3)] = (pixel)v;
                }
            }
            G::sync();
        }
        break; }
    }
}

}  // namespace b200
