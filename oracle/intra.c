/*
 * oracle/intra.c — TEST INFRASTRUCTURE (see oracle_common.h).
 *
 * Plain-C restatement of intra reconstruction at transform-block granularity:
 *   edge preparation            reference src/ipred_prepare_tmpl.c:75-204 (dav1d_prepare_intra_edges)
 *   per-tx-block glue           reference src/recon_tmpl.c:1235-1330 (luma), 1342-1398 (CFL), 1418-1540 (chroma)
 *   palette / inter-intra glue  reference src/recon_tmpl.c:1201-1223, 1400-1419 (pal_pred over the whole block),
 *                               1601-1626, 1737-1777 (intra predictor over the block, blended into the inter prediction)
 * Records (B200IntraTx, include/b200av1.h) are processed sequentially in the order given. The top edge is
 * read from the picture itself: with whole-frame reconstruction the row above is still unfiltered, which is
 * what f->ipred_edge preserves in the reference's superblock-row pipeline.
 */
#include "oracle_common.h"
#include "../include/b200av1.h"

void oracle_ipred(int mode, void *dst, ptrdiff_t stride_bytes, const void *topleft, int w, int h, int angle, int max_w, int max_h, int bdmax);
void oracle_cfl_ac(int16_t *ac, const void *ypx, ptrdiff_t stride_bytes, int w_pad, int h_pad, int w, int h, int ss_hor, int ss_ver, int bdmax);
void oracle_cfl_pred(int mode, void *dst, ptrdiff_t stride_bytes, const void *topleft, int w, int h, const int16_t *ac, int alpha, int bdmax);
void oracle_pal_pred(void *dst, ptrdiff_t stride_bytes, const void *pal, const uint8_t *idx, int w, int h, int bdmax);
void oracle_blend(void *dst, ptrdiff_t dst_stride, const void *tmp, int w, int h, const uint8_t *mask, int bdmax);
int oracle_inv_txfm_add(void *dst, ptrdiff_t stride_bytes, void *coeff, int eob, int tx, int txtp, int bdmax);
void oracle_emu_edge(intptr_t bw, intptr_t bh, intptr_t iw, intptr_t ih, intptr_t x, intptr_t y, void *dst, ptrdiff_t dst_stride,
                     const void *ref, ptrdiff_t ref_stride, int bdmax);
void oracle_mc_put(void *dst, ptrdiff_t dst_stride, const void *src, ptrdiff_t src_stride, int w, int h, int mx, int my, int filter2d, int bdmax);

static const uint8_t k_w4[19] = { 1, 2, 4, 8, 16, 1, 2, 2, 4, 4, 8, 8, 16, 1, 4, 2, 8, 4, 16 };
static const uint8_t k_h4[19] = { 1, 2, 4, 8, 16, 2, 1, 4, 2, 8, 4, 16, 8, 4, 1, 8, 2, 16, 4 };
enum { M_DC = 0, M_VERT = 1, M_HOR = 2, M_LEFT_DC = 3, M_TOP_DC = 4, M_DC128 = 5, M_Z1 = 6, M_Z2 = 7, M_Z3 = 8, M_PAETH = 12 };

static int rd(const void *p, int hbd, ptrdiff_t i) { return hbd ? ((const uint16_t *)p)[i] : ((const uint8_t *)p)[i]; }
static void wr(void *p, int hbd, ptrdiff_t i, int v) { if (hbd) ((uint16_t *)p)[i] = (uint16_t)v; else ((uint8_t *)p)[i] = (uint8_t)v; }

/* Fills tl[-2*h4*4 .. 2*w4*4] (pixel units, `tl` indexes pixels of the picture's type) and returns the DSP mode.
 * x, y, xend, yend in 4-sample units; dst = top-left sample of the block; stride in samples. */
ORACLE_API int oracle_prepare_intra_edges(int x, int have_left, int y, int have_top, int xend, int yend, int flags,
                                          const void *pic, ptrdiff_t dst_idx, ptrdiff_t stride, int mode, int *angle,
                                          int tw, int th, int filter_edge, void *tl_base, int bdmax)
{
    const int hbd = bdmax > 255, bitdepth = o_ulog2((unsigned)bdmax) + 1, half = (1 << bitdepth) >> 1;
    static const uint8_t base_angle[8] = { 90, 180, 45, 135, 113, 157, 203, 67 };
    if (mode >= 1 && mode <= 8) {
        *angle = base_angle[mode - 1] + 3 * *angle;
        if (*angle <= 90) mode = (*angle < 90 && have_top) ? M_Z1 : M_VERT;
        else if (*angle < 180) mode = M_Z2;
        else mode = (*angle > 180 && have_left) ? M_Z3 : M_HOR;
    } else if (mode == M_DC) {
        mode = have_left ? (have_top ? M_DC : M_LEFT_DC) : (have_top ? M_TOP_DC : M_DC128);
    } else if (mode == M_PAETH) {
        mode = have_left ? (have_top ? M_PAETH : M_HOR) : (have_top ? M_VERT : M_DC128);
    }
    const int w = tw * 4, h = th * 4;
    const ptrdiff_t top = dst_idx - stride;
    /* left column, bottom first in memory: tl[-(1+i)] is row i */
    {
        const int have = o_min(h, (yend - y) * 4);
        const int fill = have_top ? rd(pic, hbd, top) : half + 1;
        for (int i = 0; i < h; i++)
            wr(tl_base, hbd, 128 - (1 + i), have_left ? rd(pic, hbd, dst_idx + (ptrdiff_t)o_min(i, have - 1) * stride - 1) : fill);
        const int have_bl = have_left && y + th < yend && (flags & B200_INTRA_LEFT_HAS_BOTTOM);
        const int have2 = o_min(h, (yend - y - th) * 4);
        const int last = rd(tl_base, hbd, 128 - h);
        for (int i = 0; i < h; i++)
            wr(tl_base, hbd, 128 - (1 + h + i), have_bl ? rd(pic, hbd, dst_idx + (ptrdiff_t)(h + o_min(i, have2 - 1)) * stride - 1) : last);
    }
    {
        const int have = o_min(w, (xend - x) * 4);
        const int fill = have_left ? rd(pic, hbd, dst_idx - 1) : half - 1;
        for (int i = 0; i < w; i++)
            wr(tl_base, hbd, 128 + 1 + i, have_top ? rd(pic, hbd, top + o_min(i, have - 1)) : fill);
        const int have_tr = have_top && x + tw < xend && (flags & B200_INTRA_TOP_HAS_RIGHT);
        const int have2 = o_min(w, (xend - x - tw) * 4);
        const int last = rd(tl_base, hbd, 128 + w);
        for (int i = 0; i < w; i++)
            wr(tl_base, hbd, 128 + 1 + w + i, have_tr ? rd(pic, hbd, top + w + o_min(i, have2 - 1)) : last);
    }
    int c = have_left ? (have_top ? rd(pic, hbd, top - 1) : rd(pic, hbd, dst_idx - 1)) : (have_top ? rd(pic, hbd, top) : half);
    if (mode == M_Z2 && tw + th >= 6 && filter_edge)
        c = ((rd(tl_base, hbd, 127) + rd(tl_base, hbd, 129)) * 5 + c * 6 + 8) >> 4;
    wr(tl_base, hbd, 128, c);
    return mode;
}

ORACLE_API void oracle_intra_frame(int bdmax, const B200IntraFrame *f, const B200IntraTx *tx, int n)
{
    const int hbd = bdmax > 255;
    const size_t px = hbd ? 2 : 1, cs = hbd ? 4 : 2;
    uint16_t edge16[257];
    void *const edge = edge16;                                      /* 257 pixels of either width */
    int16_t ac[32 * 32];
    for (int i = 0; i < n; i++) {
        const B200IntraTx *r = &tx[i];
        const int tw = k_w4[r->tx], th = k_h4[r->tx], w = tw * 4, h = th * 4;
        const ptrdiff_t st = f->stride[r->plane];
        void *dst = (uint8_t *)f->pic + (size_t)r->dst_off * px;
        const int hl = !!(r->flags & B200_INTRA_HAVE_LEFT), ht = !!(r->flags & B200_INTRA_HAVE_TOP);
        int angle = r->angle;
        const void *tl = (const uint8_t *)edge + 128 * px;
        if (r->mode == B200_INTRA_MODE_RESID) {
            /* nothing to predict: the block was predicted by a PAL / II record, only the residual below is added */
        } else if (r->mode == B200_INTRA_MODE_PAL) {
            const uint8_t *pd = f->pal + r->luma_off;
            oracle_pal_pred(dst, st * (ptrdiff_t)px, pd, pd + 8 * px, w, h, bdmax);
        } else if (r->mode == B200_INTRA_MODE_IBC) {
            /* intra block copy: mc() with the picture itself as reference and the bilinear filter (reference
             * src/recon_tmpl.c:1583-1596, 938-988): emu_edge window against the plane area w4*4 x h4*4, then put_bilin */
            static __thread uint8_t win[72 * 72 * 2];
            const int pl = r->plane, sx = (int)(r->luma_off & 0xffff), sy = (int)(r->luma_off >> 16);
            const uint8_t *plane = (const uint8_t *)f->pic + ((size_t)r->dst_off - ((size_t)r->y4 * 4 * st + (size_t)r->x4 * 4)) * px;
            oracle_emu_edge(w + 7, h + 7, f->w4[pl] * 4, f->h4[pl] * 4, sx - 3, sy - 3, win, 72 * (ptrdiff_t)px, plane, st * (ptrdiff_t)px, bdmax);
            oracle_mc_put(dst, st * (ptrdiff_t)px, win + (3 * 72 + 3) * px, 72 * (ptrdiff_t)px, w, h, r->cfl_w_pad, r->cfl_h_pad, 9, bdmax);
        } else if (r->mode == B200_INTRA_MODE_II) {
            /* the predictor named by `angle` over the whole block into a scratch, then dsp->mc.blend with the mask */
            uint16_t tmp16[64 * 64];
            int a0 = 0;
            const int m = oracle_prepare_intra_edges(r->x4, hl, r->y4, ht, r->xend4, r->yend4, 0, f->pic, r->dst_off, st, r->angle,
                                                     &a0, tw, th, 0, edge, bdmax);
            oracle_ipred(m, tmp16, w * (ptrdiff_t)px, tl, w, h, 0, 0, 0, bdmax);
            oracle_blend(dst, st * (ptrdiff_t)px, tmp16, w, h, f->mask + r->luma_off, bdmax);
        } else if (r->mode == B200_INTRA_MODE_CFL && r->cfl_alpha) {
            angle = 0;
            oracle_cfl_ac(ac, (const uint8_t *)f->pic + (size_t)r->luma_off * px, f->stride[0] * (ptrdiff_t)px, r->cfl_w_pad,
                          r->cfl_h_pad, w, h, f->ss_hor, f->ss_ver, bdmax);
            const int m = oracle_prepare_intra_edges(r->x4, hl, r->y4, ht, r->xend4, r->yend4, 0, f->pic, r->dst_off, st, M_DC,
                                                     &angle, tw, th, 0, edge, bdmax);
            oracle_cfl_pred(m, dst, st * (ptrdiff_t)px, tl, w, h, ac, r->cfl_alpha, bdmax);
        } else {
            const int in = r->mode == B200_INTRA_MODE_CFL ? M_DC : r->mode;
            const int m = oracle_prepare_intra_edges(r->x4, hl, r->y4, ht, r->xend4, r->yend4, r->flags, f->pic, r->dst_off, st, in,
                                                     &angle, tw, th, (r->angle_flags >> 10) & 1, edge, bdmax);
            oracle_ipred(m, dst, st * (ptrdiff_t)px, tl, w, h, angle | r->angle_flags, r->max_w, r->max_h, bdmax);
        }
        if (r->eob >= 0) {
            const int ncf = o_min(w, 32) * o_min(h, 32);
            uint8_t save[1024 * 4];
            uint8_t *cf = (uint8_t *)f->d_coef + (size_t)r->coef_off * cs;
            memcpy(save, cf, ncf * cs);
            oracle_inv_txfm_add(dst, st * (ptrdiff_t)px, cf, r->eob, r->tx, r->txtp, bdmax);
            if (!f->zero_coefs) memcpy(cf, save, ncf * cs);
        }
    }
}
