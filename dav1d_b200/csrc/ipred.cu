// Intra prediction (dav1d Dav1dIntraPredDSPContext; reference src/ipred_tmpl.c:39-675).
// One CTA per block: the edge array is copied into shared memory as ints, directional modes first
// build their filtered / upsampled edge there, then every thread produces pixels of the block with
// consecutive threads on consecutive columns. Filter-intra walks its 4x2 units along anti-diagonals
// (each unit depends on its left / top / top-left neighbours only). Integer, bit-exact.
#include "host_util.h"
#define B200_TBL __constant__
#include "tables_gen.h"

namespace b200 {

constexpr int kIpT = 128;   // threads per block

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

__constant__ uint8_t c_edge_kernel[3][5] = { { 0, 4, 8, 4, 0 }, { 0, 5, 6, 5, 0 }, { 2, 4, 4, 4, 2 } };

// out[i], i in [0, sz): in[clamp(i)] or the 5-tap smoothed value inside [lim_from, lim_to)
B200_DEV void ip_edge_filter(int *out, int sz, int lim_from, int lim_to, const int *in, int from, int to, int strength) {
    for (int i = threadIdx.x; i < sz; i += kIpT) {
        if (i < imin(sz, lim_from) || i >= imin(lim_to, sz)) { out[i] = in[iclip(i, from, to - 1)]; continue; }
        int s = 0;
#pragma unroll
        for (int j = 0; j < 5; j++) s += in[iclip(i - 2 + j, from, to - 1)] * c_edge_kernel[strength - 1][j];
        out[i] = (s + 8) >> 4;
    }
}
B200_DEV void ip_edge_upsample(int *out, int hsz, const int *in, int from, int to, int bdmax) {
    for (int i = threadIdx.x; i < hsz; i += kIpT) {
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

template <bool HBD>
__global__ void __launch_bounds__(kIpT) ipred_kernel(const B200IpredBlock *__restrict__ blocks, int n, B200IpredFrame f, int bdmax)
{
    typedef typename Bd<HBD>::pixel pixel;
    __shared__ int s_edge[2 * 128 + 1];      // tl = s_edge + 128
    __shared__ int s_aux[64 + 64 + 1 + 64 + 8];
    __shared__ int s_tile[32 * 32];          // filter-intra working tile / reductions
    __shared__ int s_dc;
    const B200IpredBlock b = blocks[blockIdx.x];
    const int w = b.w, h = b.h, tid = threadIdx.x;
    pixel *const dst = (pixel *)f.dst + b.dst_off;
    const int st = f.dst_stride[b.plane];
    int *const tl = s_edge + 128;

    if (b.op == B200_IPRED_OP_PAL_PRED) {
        const pixel *pal = (const pixel *)f.edge + b.edge_off;
        const uint8_t *idx = f.pal_idx + b.ac_off;
        for (int i = tid; i < (w * h) >> 1; i += kIpT) {
            const int y = i / (w >> 1), x = (i - y * (w >> 1)) * 2, v = idx[i];
            dst[(ptrdiff_t)y * st + x] = pal[v & 7];
            dst[(ptrdiff_t)y * st + x + 1] = pal[v >> 4];
        }
        return;
    }
    if (b.op == B200_IPRED_OP_CFL_AC) {
        // dst_off addresses the luma block inside the picture (plane 0)
        const pixel *ypx = (const pixel *)f.dst + b.dst_off;
        const int ys = f.dst_stride[0], ssh = f.ss_hor, ssv = f.ss_ver;
        const int w_pad = b.angle & 0xff, h_pad = (b.angle >> 8) & 0xff;
        int16_t *ac = f.ac + b.ac_off;
        int part = 0;
        for (int i = tid; i < w * h; i += kIpT) {
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
        __syncthreads();
        if (tid == 0) {
            const int log2sz = (__ffs(w) - 1) + (__ffs(h) - 1);
            int sum = (1 << log2sz) >> 1;
            for (int i = 0; i < kIpT; i++) sum += s_tile[i];
            s_dc = sum >> log2sz;
        }
        __syncthreads();
        const int dc = s_dc;
        for (int i = tid; i < w * h; i += kIpT) ac[i] = (int16_t)(ac[i] - dc);
        return;
    }

    // ---- edge array to shared memory: tl[-(w+h) .. w+h] ----
    {
        const pixel *e = (const pixel *)f.edge + b.edge_off;
        for (int i = tid - (w + h); i <= w + h; i += kIpT) tl[i] = e[i];
    }
    __syncthreads();
    const int mode = b.mode;

    if (b.op == B200_IPRED_OP_CFL_PRED) {
        if (tid == 0) s_dc = ip_dc(tl, w, h, mode, bdmax, HBD);
        __syncthreads();
        const int dc = s_dc, alpha = b.alpha;
        const int16_t *ac = f.ac + b.ac_off;
        for (int i = tid; i < w * h; i += kIpT) {
            const int y = i / w, x = i - y * w;
            const int diff = alpha * ac[i];
            const int m = (iabs(diff) + 32) >> 6;
            dst[(ptrdiff_t)y * st + x] = (pixel)iclip(dc + (diff < 0 ? -m : m), 0, bdmax);
        }
        return;
    }

    int angle = b.angle;
    switch (mode) {
    case B200_DC_PRED: case B200_TOP_DC_PRED: case B200_LEFT_DC_PRED: case B200_DC_128_PRED: {
        if (tid == 0) s_dc = ip_dc(tl, w, h, mode, bdmax, HBD);
        __syncthreads();
        const int dc = s_dc;
        for (int i = tid; i < w * h; i += kIpT) dst[(ptrdiff_t)(i / w) * st + (i % w)] = (pixel)dc;
        break; }
    case B200_VERT_PRED:
        for (int i = tid; i < w * h; i += kIpT) dst[(ptrdiff_t)(i / w) * st + (i % w)] = (pixel)tl[1 + (i % w)];
        break;
    case B200_HOR_PRED:
        for (int i = tid; i < w * h; i += kIpT) dst[(ptrdiff_t)(i / w) * st + (i % w)] = (pixel)tl[-(1 + i / w)];
        break;
    case B200_PAETH_PRED:
        for (int i = tid; i < w * h; i += kIpT) {
            const int y = i / w, x = i - y * w;
            const int l = tl[-(y + 1)], t = tl[1 + x], c = tl[0], base = l + t - c;
            const int ld = iabs(l - base), td = iabs(t - base), cd = iabs(c - base);
            dst[(ptrdiff_t)y * st + x] = (pixel)(ld <= td && ld <= cd ? l : td <= cd ? t : c);
        }
        break;
    case B200_SMOOTH_PRED: case B200_SMOOTH_V_PRED: case B200_SMOOTH_H_PRED: {
        const int right = tl[w], bottom = tl[-h];
        for (int i = tid; i < w * h; i += kIpT) {
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
        if (up) { ip_edge_upsample(s_aux, w + h, &tl[1], -1, w + imin(w, h), bdmax); top = s_aux; max_base_x = 2 * (w + h) - 2; dx <<= 1; }
        else if (fs) { ip_edge_filter(s_aux, w + h, 0, w + h, &tl[1], -1, w + imin(w, h), fs); top = s_aux; max_base_x = w + h - 1; }
        else { top = &tl[1]; max_base_x = w + imin(w, h) - 1; }
        __syncthreads();
        const int inc = 1 + up;
        for (int i = tid; i < w * h; i += kIpT) {
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
        if (up_a) { ip_edge_upsample(e, w + 1, tl, 0, w + 1, bdmax); dx <<= 1; }
        else {
            const int fs = eief ? ip_filter_strength(w + h, angle - 90, is_sm) : 0;
            if (fs) ip_edge_filter(&e[1], w, 0, b.max_w, &tl[1], -1, w, fs);
            else for (int i = tid; i < w; i += kIpT) e[1 + i] = tl[1 + i];
        }
        if (up_l) { ip_edge_upsample(&e[-h * 2], h + 1, &tl[-h], 0, h + 1, bdmax); dy <<= 1; }
        else {
            const int fs = eief ? ip_filter_strength(w + h, 180 - angle, is_sm) : 0;
            if (fs) ip_edge_filter(&e[-h], h, h - b.max_h, h, &tl[-h], 0, h + 1, fs);
            else for (int i = tid; i < h; i += kIpT) e[-h + i] = tl[-h + i];
        }
        __syncthreads();
        if (tid == 0) e[0] = tl[0];               // after the upsamplers (which also write e[0])
        __syncthreads();
        const int inc_x = 1 + up_a;
        const int *left = &e[-(1 + up_l)];
        for (int i = tid; i < w * h; i += kIpT) {
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
        if (up) { ip_edge_upsample(s_aux, w + h, &tl[-(w + h)], imax(w - h, 0), w + h + 1, bdmax); left = &s_aux[2 * (w + h) - 2]; max_base_y = 2 * (w + h) - 2; dy <<= 1; }
        else if (fs) { ip_edge_filter(s_aux, w + h, 0, w + h, &tl[-(w + h)], imax(w - h, 0), w + h + 1, fs); left = &s_aux[w + h - 1]; max_base_y = w + h - 1; }
        else { left = &tl[-1]; max_base_y = h + imin(w, h) - 1; }
        __syncthreads();
        const int inc = 1 + up;
        for (int i = tid; i < w * h; i += kIpT) {
            const int y = i / w, x = i - y * w;
            const int ypos = dy * (x + 1), frac = ypos & 0x3E, base = (ypos >> 6) + y * inc;
            dst[(ptrdiff_t)y * st + x] = (pixel)(base < max_base_y ? (left[-base] * (64 - frac) + left[-(base + 1)] * frac + 32) >> 6 : left[-max_base_y]);
        }
        break; }
    case B200_FILTER_PRED: {
        const int8_t *flt = b200_filter_intra_taps[angle & 511];
        const int uw = w >> 2, uh = h >> 1;
        for (int d = 0; d < uw + uh - 1; d++) {
            for (int ux = tid; ux < uw; ux += kIpT) {
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
            __syncthreads();
        }
        break; }
    }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_ipred_batch(int bdmax, const B200IpredFrame *f, const B200IpredBlock *d_blocks, int n, void *stream)
{
    if (bdmax != 255 && bdmax != 1023 && bdmax != 4095) { b200_set_error("b200_ipred_batch: bad bitdepth_max"); return -2; }
    if (n <= 0) return 0;
    if (bdmax > 255) { auto k = ipred_kernel<true>; B200_LAUNCH(k, dim3(n), dim3(kIpT), 0, (cudaStream_t)stream, d_blocks, n, *f, bdmax); }
    else { auto k = ipred_kernel<false>; B200_LAUNCH(k, dim3(n), dim3(kIpT), 0, (cudaStream_t)stream, d_blocks, n, *f, bdmax); }
    b200_count_launch();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}

// ---- Level 1 -------------------------------------------------------------------------------
namespace {
Scratch s_dst, s_edge, s_ac, s_idx, s_desc;
uint8_t h_px[64 * 64 * 2];

int ipred_l1(int op, int mode, void *dst, ptrdiff_t stride, const void *topleft, int w, int h, int angle, int max_w,
             int max_h, const int16_t *ac, int alpha, int bdmax)
{
    std::lock_guard<std::mutex> lk(host_lock());
    const size_t px = bdmax > 255 ? 2 : 1;
    const int n_edge = 2 * (w + h) + 1;
    // the edge window [-(w+h), w+h] around topleft
    if (s_edge.upload((const uint8_t *)topleft - (ptrdiff_t)(w + h) * (ptrdiff_t)px, (size_t)n_edge * px)) return -1;
    if (s_dst.reserve((size_t)w * h * px) || s_desc.reserve(sizeof(B200IpredBlock))) return -1;
    if (op == B200_IPRED_OP_CFL_PRED && s_ac.upload(ac, (size_t)w * h * 2)) return -1;
    B200IpredFrame f;
    memset(&f, 0, sizeof(f));
    f.dst = s_dst.p; f.dst_stride[0] = w; f.edge = s_edge.p; f.ac = (int16_t *)s_ac.p;
    B200IpredBlock b;
    memset(&b, 0, sizeof(b));
    b.edge_off = (uint32_t)(w + h); b.w = (uint8_t)w; b.h = (uint8_t)h; b.mode = (uint8_t)mode; b.op = (uint8_t)op;
    b.angle = (int16_t)angle; b.max_w = max_w; b.max_h = max_h; b.alpha = (int8_t)alpha;
    if (s_desc.upload(&b, sizeof(b))) return -1;
    int r = b200_ipred_batch(bdmax, &f, (const B200IpredBlock *)s_desc.p, 1, 0);
    if (r) return r;
    if (s_dst.download(h_px, (size_t)w * h * px)) return -1;
    B200_CUDA_OK(cudaStreamSynchronize(0));
    unpack_rect(dst, stride, h_px, w, h, px);
    return 0;
}
}  // namespace

extern "C" {

int b200_ipred(int mode, void *dst, ptrdiff_t stride, const void *topleft, int w, int h, int angle, int max_w, int max_h, int bdmax)
{
    if (mode < 0 || mode > 13 || w < 4 || w > 64 || h < 4 || h > 64 || (mode == 13 && (w > 32 || h > 32))) { b200_set_error("b200_ipred: bad arguments"); return -2; }
    return ipred_l1(B200_IPRED_OP_PRED, mode, dst, stride, topleft, w, h, angle, max_w, max_h, nullptr, 0, bdmax);
}
int b200_cfl_pred(int mode, void *dst, ptrdiff_t stride, const void *topleft, int w, int h, const int16_t *ac, int alpha, int bdmax)
{
    if (!(mode == 0 || mode == 3 || mode == 4 || mode == 5) || w < 4 || w > 32 || h < 4 || h > 32) { b200_set_error("b200_cfl_pred: bad arguments"); return -2; }
    return ipred_l1(B200_IPRED_OP_CFL_PRED, mode, dst, stride, topleft, w, h, 0, 0, 0, ac, alpha, bdmax);
}
int b200_cfl_ac(int16_t *ac, const void *ypx, ptrdiff_t stride, int w_pad, int h_pad, int cw, int ch, int ss_hor, int ss_ver, int bdmax)
{
    if (cw < 4 || cw > 32 || ch < 4 || ch > 32) { b200_set_error("b200_cfl_ac: bad arguments"); return -2; }
    std::lock_guard<std::mutex> lk(host_lock());
    const size_t px = bdmax > 255 ? 2 : 1;
    const int lw = (cw - 4 * w_pad) << ss_hor, lh = (ch - 4 * h_pad) << ss_ver;   // luma samples actually read
    pack_rect(h_px, ypx, stride, lw, lh, px);
    if (s_dst.upload(h_px, (size_t)lw * lh * px) || s_ac.reserve((size_t)cw * ch * 2) || s_desc.reserve(sizeof(B200IpredBlock))) return -1;
    B200IpredFrame f;
    memset(&f, 0, sizeof(f));
    f.dst = s_dst.p; f.dst_stride[0] = lw; f.ss_hor = ss_hor; f.ss_ver = ss_ver; f.ac = (int16_t *)s_ac.p;
    B200IpredBlock b;
    memset(&b, 0, sizeof(b));
    b.w = (uint8_t)cw; b.h = (uint8_t)ch; b.op = B200_IPRED_OP_CFL_AC; b.angle = (int16_t)(w_pad | (h_pad << 8));
    if (s_desc.upload(&b, sizeof(b))) return -1;
    int r = b200_ipred_batch(bdmax, &f, (const B200IpredBlock *)s_desc.p, 1, 0);
    if (r) return r;
    if (s_ac.download(ac, (size_t)cw * ch * 2)) return -1;
    B200_CUDA_OK(cudaStreamSynchronize(0));
    return 0;
}
int b200_pal_pred(void *dst, ptrdiff_t stride, const void *pal, const uint8_t *idx, int w, int h, int bdmax)
{
    if (w < 4 || w > 64 || h < 4 || h > 64) { b200_set_error("b200_pal_pred: bad arguments"); return -2; }
    std::lock_guard<std::mutex> lk(host_lock());
    const size_t px = bdmax > 255 ? 2 : 1;
    if (s_edge.upload(pal, 8 * px) || s_idx.upload(idx, (size_t)w * h / 2) || s_dst.reserve((size_t)w * h * px) || s_desc.reserve(sizeof(B200IpredBlock))) return -1;
    B200IpredFrame f;
    memset(&f, 0, sizeof(f));
    f.dst = s_dst.p; f.dst_stride[0] = w; f.edge = s_edge.p; f.pal_idx = (const uint8_t *)s_idx.p;
    B200IpredBlock b;
    memset(&b, 0, sizeof(b));
    b.w = (uint8_t)w; b.h = (uint8_t)h; b.op = B200_IPRED_OP_PAL_PRED;
    if (s_desc.upload(&b, sizeof(b))) return -1;
    int r = b200_ipred_batch(bdmax, &f, (const B200IpredBlock *)s_desc.p, 1, 0);
    if (r) return r;
    if (s_dst.download(h_px, (size_t)w * h * px)) return -1;
    B200_CUDA_OK(cudaStreamSynchronize(0));
    unpack_rect(dst, stride, h_px, w, h, px);
    return 0;
}

}  // extern "C"

namespace {
template <int M> void ip8(uint8_t *d, ptrdiff_t s, const uint8_t *tl, int w, int h, int a, int mw, int mh) { if (b200_ipred(M, d, s, tl, w, h, a, mw, mh, 255)) die("intra_pred"); }
template <int M> void ip16(uint16_t *d, ptrdiff_t s, const uint16_t *tl, int w, int h, int a, int mw, int mh, int bd) { if (b200_ipred(M, d, s, tl, w, h, a, mw, mh, bd)) die("intra_pred"); }
template <int M> void cp8(uint8_t *d, ptrdiff_t s, const uint8_t *tl, int w, int h, const int16_t *ac, int al) { if (b200_cfl_pred(M, d, s, tl, w, h, ac, al, 255)) die("cfl_pred"); }
template <int M> void cp16(uint16_t *d, ptrdiff_t s, const uint16_t *tl, int w, int h, const int16_t *ac, int al, int bd) { if (b200_cfl_pred(M, d, s, tl, w, h, ac, al, bd)) die("cfl_pred"); }
template <int BD, int SH, int SV> void ca(int16_t *ac, const void *y, ptrdiff_t s, int wp, int hp, int cw, int ch) { if (b200_cfl_ac(ac, y, s, wp, hp, cw, ch, SH, SV, BD)) die("cfl_ac"); }
template <int BD> void pp(void *d, ptrdiff_t s, const void *pal, const uint8_t *idx, int w, int h) { if (b200_pal_pred(d, s, pal, idx, w, h, BD)) die("pal_pred"); }
template <int... M> void fill_ip8(B200IntraPredDSPContext *c, std::integer_sequence<int, M...>) { ((c->intra_pred[M] = (void *)ip8<M>), ...); }
template <int... M> void fill_ip16(B200IntraPredDSPContext *c, std::integer_sequence<int, M...>) { ((c->intra_pred[M] = (void *)ip16<M>), ...); }
}
extern "C" {
void b200_intra_pred_dsp_init_8bpc(B200IntraPredDSPContext *c) {
    memset(c, 0, sizeof(*c));
    fill_ip8(c, std::make_integer_sequence<int, 14>{});
    c->cfl_ac[0] = (void *)ca<255, 1, 1>; c->cfl_ac[1] = (void *)ca<255, 1, 0>; c->cfl_ac[2] = (void *)ca<255, 0, 0>;
    c->cfl_pred[0] = (void *)cp8<0>; c->cfl_pred[3] = (void *)cp8<3>; c->cfl_pred[4] = (void *)cp8<4>; c->cfl_pred[5] = (void *)cp8<5>;
    c->pal_pred = (void *)pp<255>;
}
void b200_intra_pred_dsp_init_16bpc(B200IntraPredDSPContext *c) {
    memset(c, 0, sizeof(*c));
    fill_ip16(c, std::make_integer_sequence<int, 14>{});
    // cfl_ac / pal_pred carry no bit-depth argument in dav1d; only the pixel width matters
    c->cfl_ac[0] = (void *)ca<1023, 1, 1>; c->cfl_ac[1] = (void *)ca<1023, 1, 0>; c->cfl_ac[2] = (void *)ca<1023, 0, 0>;
    c->cfl_pred[0] = (void *)cp16<0>; c->cfl_pred[3] = (void *)cp16<3>; c->cfl_pred[4] = (void *)cp16<4>; c->cfl_pred[5] = (void *)cp16<5>;
    c->pal_pred = (void *)pp<1023>;
}
}
