/*
 * oracle/ipred.c — TEST INFRASTRUCTURE (see oracle_common.h).
 *
 * Plain-C restatement of dav1d's intra predictors (reference src/ipred_tmpl.c):
 *   dc / dc_top / dc_left / dc_128 (:39-69, 86-203), v / h (:205-229), paeth (:231-252),
 *   smooth / smooth_v / smooth_h (:254-312), directional z1 / z2 / z3 with edge filter and
 *   upsampling (:314-535), recursive filter intra (:537-600), cfl_ac (:602-660) + cfl_pred
 *   (:71-84), pal_pred (:662-675).
 * Edge convention: `tl` points at the top-left sample; tl[1 + x] is the row above, tl[-(1 + y)]
 * the column to the left. Modes are the DSP-table indices (reference src/levels.h:112-136).
 */
#include "oracle_common.h"
#include "tables_gen.h"

enum { M_DC = 0, M_VERT = 1, M_HOR = 2, M_LEFT_DC = 3, M_TOP_DC = 4, M_DC_128 = 5, M_Z1 = 6, M_Z2 = 7, M_Z3 = 8,
       M_SMOOTH = 9, M_SMOOTH_V = 10, M_SMOOTH_H = 11, M_PAETH = 12, M_FILTER = 13 };

static inline int PX(const void *p, int hbd, ptrdiff_t i) {
    return hbd ? ((const uint16_t *)p)[i] : ((const uint8_t *)p)[i];
}
static inline void SPX(void *p, int hbd, ptrdiff_t i, int v) {
    if (hbd) ((uint16_t *)p)[i] = (uint16_t)v; else ((uint8_t *)p)[i] = (uint8_t)v;
}
static inline int ctz_(unsigned v) { return __builtin_ctz(v); }

static unsigned dc_value(const int *tl, int w, int h, int mode, int bdmax) {
    const int hbd = bdmax > 255;
    if (mode == M_DC_128) return hbd ? (unsigned)(bdmax + 1) >> 1 : 128;
    unsigned dc = 0;
    if (mode == M_TOP_DC) { dc = w >> 1; for (int i = 0; i < w; i++) dc += tl[1 + i]; return dc >> ctz_(w); }
    if (mode == M_LEFT_DC) { dc = h >> 1; for (int i = 0; i < h; i++) dc += tl[-(1 + i)]; return dc >> ctz_(h); }
    dc = (w + h) >> 1;
    for (int i = 0; i < w; i++) dc += tl[1 + i];
    for (int i = 0; i < h; i++) dc += tl[-(1 + i)];
    dc >>= ctz_(w + h);
    if (w != h) {
        const unsigned m12 = hbd ? 0xAAAB : 0x5556, m14 = hbd ? 0x6667 : 0x3334;
        dc *= (w > h * 2 || h > w * 2) ? m14 : m12;
        dc >>= hbd ? 17 : 16;
    }
    return dc;
}

static int filter_strength(int wh, int angle, int is_sm) {
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
static int do_upsample(int wh, int angle, int is_sm) { return angle < 40 && wh <= (16 >> is_sm); }

/* out[0..sz) from in[] with index clamping to [from, to) ; filtered inside [lim_from, lim_to) */
static void edge_filter(int *out, int sz, int lim_from, int lim_to, const int *in, int from, int to, int strength) {
    static const uint8_t k[3][5] = { { 0, 4, 8, 4, 0 }, { 0, 5, 6, 5, 0 }, { 2, 4, 4, 4, 2 } };
    for (int i = 0; i < sz; i++) {
        if (i < o_min(sz, lim_from) || i >= o_min(lim_to, sz)) { out[i] = in[o_clip(i, from, to - 1)]; continue; }
        int s = 0;
        for (int j = 0; j < 5; j++) s += in[o_clip(i - 2 + j, from, to - 1)] * k[strength - 1][j];
        out[i] = (s + 8) >> 4;
    }
}
static void edge_upsample(int *out, int hsz, const int *in, int from, int to, int bdmax) {
    static const int8_t k[4] = { -1, 9, 9, -1 };
    int i;
    for (i = 0; i < hsz - 1; i++) {
        out[i * 2] = in[o_clip(i, from, to - 1)];
        int s = 0;
        for (int j = 0; j < 4; j++) s += in[o_clip(i + j - 1, from, to - 1)] * k[j];
        out[i * 2 + 1] = o_clip((s + 8) >> 4, 0, bdmax);
    }
    out[i * 2] = in[o_clip(i, from, to - 1)];
}

ORACLE_API void oracle_ipred(int mode, void *dst, ptrdiff_t stride_bytes, const void *topleft, int w, int h,
                             int angle, int max_w, int max_h, int bdmax)
{
    const int hbd = bdmax > 255;
    const ptrdiff_t ps = hbd ? stride_bytes / 2 : stride_bytes;
    /* int copy of the edge: e[256 + i] = topleft[i] for the range the predictors may touch */
    int ebuf[513], *tl = ebuf + 256;
    for (int i = -(2 * h + 0); i <= 2 * w; i++) tl[i] = PX(topleft, hbd, i);
#define OUT(x, y, v) SPX(dst, hbd, (ptrdiff_t)(y) * ps + (x), (v))
    switch (mode) {
    case M_DC: case M_TOP_DC: case M_LEFT_DC: case M_DC_128: {
        const int dc = (int)dc_value(tl, w, h, mode, bdmax);
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) OUT(x, y, dc);
        break; }
    case M_VERT: for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) OUT(x, y, tl[1 + x]); break;
    case M_HOR:  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) OUT(x, y, tl[-(1 + y)]); break;
    case M_PAETH:
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
            const int l = tl[-(y + 1)], t = tl[1 + x], c = tl[0], base = l + t - c;
            const int ld = o_abs(l - base), td = o_abs(t - base), cd = o_abs(c - base);
            OUT(x, y, ld <= td && ld <= cd ? l : td <= cd ? t : c);
        }
        break;
    case M_SMOOTH: case M_SMOOTH_V: case M_SMOOTH_H: {
        const uint8_t *wh_ = &b200_sm_weights[w], *wv = &b200_sm_weights[h];
        const int right = tl[w], bottom = tl[-h];
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
            if (mode == M_SMOOTH)
                OUT(x, y, (wv[y] * tl[1 + x] + (256 - wv[y]) * bottom + wh_[x] * tl[-(1 + y)] + (256 - wh_[x]) * right + 256) >> 9);
            else if (mode == M_SMOOTH_V) OUT(x, y, (wv[y] * tl[1 + x] + (256 - wv[y]) * bottom + 128) >> 8);
            else OUT(x, y, (wh_[x] * tl[-(y + 1)] + (256 - wh_[x]) * right + 128) >> 8);
        }
        break; }
    case M_Z1: {
        const int is_sm = (angle >> 9) & 1, eief = angle >> 10; angle &= 511;
        int dx = b200_dr_intra_derivative[angle >> 1];
        int top_out[128]; const int *top; int max_base_x;
        const int up = eief ? do_upsample(w + h, 90 - angle, is_sm) : 0;
        if (up) {
            edge_upsample(top_out, w + h, &tl[1], -1, w + o_min(w, h), bdmax);
            top = top_out; max_base_x = 2 * (w + h) - 2; dx <<= 1;
        } else {
            const int fs = eief ? filter_strength(w + h, 90 - angle, is_sm) : 0;
            if (fs) { edge_filter(top_out, w + h, 0, w + h, &tl[1], -1, w + o_min(w, h), fs); top = top_out; max_base_x = w + h - 1; }
            else { top = &tl[1]; max_base_x = w + o_min(w, h) - 1; }
        }
        const int inc = 1 + up;
        for (int y = 0, xpos = dx; y < h; y++, xpos += dx) {
            const int frac = xpos & 0x3E;
            for (int x = 0, base = xpos >> 6; x < w; x++, base += inc)
                OUT(x, y, base < max_base_x ? (top[base] * (64 - frac) + top[base + 1] * frac + 32) >> 6 : top[max_base_x]);
        }
        break; }
    case M_Z2: {
        const int is_sm = (angle >> 9) & 1, eief = angle >> 10; angle &= 511;
        int dy = b200_dr_intra_derivative[(angle - 90) >> 1], dx = b200_dr_intra_derivative[(180 - angle) >> 1];
        const int up_l = eief ? do_upsample(w + h, 180 - angle, is_sm) : 0;
        const int up_a = eief ? do_upsample(w + h, angle - 90, is_sm) : 0;
        int edge[64 + 64 + 1 + 64], *e = &edge[128];
        if (up_a) { edge_upsample(e, w + 1, tl, 0, w + 1, bdmax); dx <<= 1; }
        else {
            const int fs = eief ? filter_strength(w + h, angle - 90, is_sm) : 0;
            if (fs) edge_filter(&e[1], w, 0, max_w, &tl[1], -1, w, fs);
            else for (int i = 0; i < w; i++) e[1 + i] = tl[1 + i];
        }
        if (up_l) { edge_upsample(&e[-h * 2], h + 1, &tl[-h], 0, h + 1, bdmax); dy <<= 1; }
        else {
            const int fs = eief ? filter_strength(w + h, 180 - angle, is_sm) : 0;
            if (fs) edge_filter(&e[-h], h, h - max_h, h, &tl[-h], 0, h + 1, fs);
            else for (int i = 0; i < h; i++) e[-h + i] = tl[-h + i];
        }
        e[0] = tl[0];
        const int inc_x = 1 + up_a;
        const int *left = &e[-(1 + up_l)];
        for (int y = 0, xpos = ((1 + up_a) << 6) - dx; y < h; y++, xpos -= dx) {
            int base_x = xpos >> 6; const int frac_x = xpos & 0x3E;
            for (int x = 0, ypos = (y << (6 + up_l)) - dy; x < w; x++, base_x += inc_x, ypos -= dy) {
                int v;
                if (base_x >= 0) v = e[base_x] * (64 - frac_x) + e[base_x + 1] * frac_x;
                else { const int by = ypos >> 6, fy = ypos & 0x3E; v = left[-by] * (64 - fy) + left[-(by + 1)] * fy; }
                OUT(x, y, (v + 32) >> 6);
            }
        }
        break; }
    case M_Z3: {
        const int is_sm = (angle >> 9) & 1, eief = angle >> 10; angle &= 511;
        int dy = b200_dr_intra_derivative[(270 - angle) >> 1];
        int left_out[128]; const int *left; int max_base_y;
        const int up = eief ? do_upsample(w + h, angle - 180, is_sm) : 0;
        if (up) {
            edge_upsample(left_out, w + h, &tl[-(w + h)], o_max(w - h, 0), w + h + 1, bdmax);
            left = &left_out[2 * (w + h) - 2]; max_base_y = 2 * (w + h) - 2; dy <<= 1;
        } else {
            const int fs = eief ? filter_strength(w + h, angle - 180, is_sm) : 0;
            if (fs) { edge_filter(left_out, w + h, 0, w + h, &tl[-(w + h)], o_max(w - h, 0), w + h + 1, fs); left = &left_out[w + h - 1]; max_base_y = w + h - 1; }
            else { left = &tl[-1]; max_base_y = h + o_min(w, h) - 1; }
        }
        const int inc = 1 + up;
        for (int x = 0, ypos = dy; x < w; x++, ypos += dy) {
            const int frac = ypos & 0x3E;
            for (int y = 0, base = ypos >> 6; y < h; y++, base += inc)
                OUT(x, y, base < max_base_y ? (left[-base] * (64 - frac) + left[-(base + 1)] * frac + 32) >> 6 : left[-max_base_y]);
        }
        break; }
    case M_FILTER: {
        /* 4x2 units in raster order; each reads its 7 neighbours (above row, top-left, 2 left) from
         * the edge or from already predicted output. Tap layout is the ARCH_X86 one (:537-545). */
        const int8_t *f = b200_filter_intra_taps[angle & 511];
        for (int y = 0; y < h; y += 2)
            for (int x = 0; x < w; x += 4) {
                int p[7];
                p[0] = y ? (x ? PX(dst, hbd, (y - 1) * ps + x - 1) : tl[-y]) : tl[x];
                for (int i = 0; i < 4; i++) p[1 + i] = y ? PX(dst, hbd, (y - 1) * ps + x + i) : tl[1 + x + i];
                for (int i = 0; i < 2; i++) p[5 + i] = x ? PX(dst, hbd, (y + i) * ps + x - 1) : tl[-(1 + y + i)];
                for (int yy = 0; yy < 2; yy++)
                    for (int xx = 0; xx < 4; xx++) {
                        const int8_t *t = f + (yy * 4 + xx) * 2;
                        const int acc = t[0] * p[0] + t[1] * p[1] + t[16] * p[2] + t[17] * p[3] + t[32] * p[4] + t[33] * p[5] + t[48] * p[6];
                        OUT(x + xx, y + yy, o_clip((acc + 8) >> 4, 0, bdmax));
                    }
            }
        break; }
    }
#undef OUT
}

ORACLE_API void oracle_cfl_ac(int16_t *ac, const void *ypx, ptrdiff_t stride_bytes, int w_pad, int h_pad, int w, int h,
                              int ss_hor, int ss_ver, int bdmax)
{
    const int hbd = bdmax > 255;
    const ptrdiff_t ps = hbd ? stride_bytes / 2 : stride_bytes;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int sy = o_min(y, h - 4 * h_pad - 1), sx = o_min(x, w - 4 * w_pad - 1);   /* padding repeats the last real sample */
            const ptrdiff_t o = (ptrdiff_t)(sy << ss_ver) * ps + (sx << ss_hor);
            int s = PX(ypx, hbd, o);
            if (ss_hor) s += PX(ypx, hbd, o + 1);
            if (ss_ver) { s += PX(ypx, hbd, o + ps); if (ss_hor) s += PX(ypx, hbd, o + ps + 1); }
            ac[y * w + x] = (int16_t)(s << (1 + !ss_ver + !ss_hor));
        }
    const int log2sz = ctz_(w) + ctz_(h);
    int sum = (1 << log2sz) >> 1;
    for (int i = 0; i < w * h; i++) sum += ac[i];
    sum >>= log2sz;
    for (int i = 0; i < w * h; i++) ac[i] = (int16_t)(ac[i] - sum);
}

/* mode: DC_PRED 0, LEFT_DC 3, TOP_DC 4, DC_128 5 (index of c->cfl_pred[]) */
ORACLE_API void oracle_cfl_pred(int mode, void *dst, ptrdiff_t stride_bytes, const void *topleft, int w, int h,
                                const int16_t *ac, int alpha, int bdmax)
{
    const int hbd = bdmax > 255;
    const ptrdiff_t ps = hbd ? stride_bytes / 2 : stride_bytes;
    int ebuf[513], *tl = ebuf + 256;
    for (int i = -h; i <= w; i++) tl[i] = PX(topleft, hbd, i);
    const int dc = (int)dc_value(tl, w, h, mode, bdmax);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int diff = alpha * ac[y * w + x];
            const int m = (o_abs(diff) + 32) >> 6;
            SPX(dst, hbd, y * ps + x, o_clip(dc + (diff < 0 ? -m : m), 0, bdmax));
        }
}

ORACLE_API void oracle_pal_pred(void *dst, ptrdiff_t stride_bytes, const void *pal, const uint8_t *idx, int w, int h, int bdmax)
{
    const int hbd = bdmax > 255;
    const ptrdiff_t ps = hbd ? stride_bytes / 2 : stride_bytes;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x += 2) {
            const int i = *idx++;
            SPX(dst, hbd, y * ps + x, PX(pal, hbd, i & 7));
            SPX(dst, hbd, y * ps + x + 1, PX(pal, hbd, i >> 4));
        }
}
