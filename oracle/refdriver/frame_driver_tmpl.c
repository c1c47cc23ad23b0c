/*
 * oracle/refdriver/frame_driver_tmpl.c — TEST INFRASTRUCTURE, compiled at BITDEPTH 8 and 16 and
 * linked into oracle/_ref/libdav1d_ref.so.
 *
 * Drives the reference's OWN frame-level drivers (src/lf_apply_tmpl.c, ...) from a hand-built
 * Dav1dFrameContext holding just the fields those drivers read, so that the frame-wide CUDA
 * sweeps can be checked against dav1d's real per-superblock-row code, not only against our
 * restatement. The caller's frame description restates B200LfFrame (include/b200av1.h).
 */
#include "config.h"
#include <stdlib.h>
#include <string.h>
#include "common/bitdepth.h"
#include "src/internal.h"
#include "src/lf_apply.h"
#include "src/loopfilter.h"

#define API __attribute__((visibility("default")))

typedef struct {
    void *pic;
    uint32_t plane_off[3];
    int32_t stride[3];
    int32_t w4, h4, sb128w, b4_stride, ss_hor, ss_ver, sb128, filter_y, filter_uv;
    Av1Filter *mask;
    uint8_t (*level)[4];
    struct { uint8_t e[64], i[64]; uint64_t sharp[2]; } lut;   /* unaligned, as the caller packs it */
} RefLfFrame;

API void bitfn(refdrv_lf_frame)(const int bitdepth_max, const RefLfFrame *const fr)
{
    Dav1dFrameContext *const f = calloc(1, sizeof(*f));
    Dav1dSequenceHeader *const seq = calloc(1, sizeof(*seq));
    Dav1dFrameHeader *const hdr = calloc(1, sizeof(*hdr));
    Dav1dDSPContext *const dsp = calloc(1, sizeof(*dsp));
    bitfn(dav1d_loop_filter_dsp_init)(&dsp->lf);
    f->dsp = dsp;
    f->seq_hdr = seq;
    f->frame_hdr = hdr;
    seq->sb128 = fr->sb128;
    f->w4 = fr->w4; f->h4 = fr->h4;
    f->bw = fr->w4; f->bh = fr->h4;            /* no super-res, frame size == picture size here */
    f->sb128w = fr->sb128w;
    f->b4_stride = fr->b4_stride;
    f->sb_step = 32 >> !fr->sb128;
    f->sbh = (fr->h4 + f->sb_step - 1) / f->sb_step;
    f->cur.p.layout = !fr->ss_hor ? DAV1D_PIXEL_LAYOUT_I444 : fr->ss_ver ? DAV1D_PIXEL_LAYOUT_I420 : DAV1D_PIXEL_LAYOUT_I422;
    f->cur.stride[0] = fr->stride[0] * (ptrdiff_t)sizeof(pixel);
    f->cur.stride[1] = fr->stride[1] * (ptrdiff_t)sizeof(pixel);
#if BITDEPTH == 16
    f->bitdepth_max = bitdepth_max;
#endif
    f->lf.level = fr->level;
    memcpy(f->lf.lim_lut.e, fr->lut.e, 64);
    memcpy(f->lf.lim_lut.i, fr->lut.i, 64);
    f->lf.lim_lut.sharp[0] = fr->lut.sharp[0]; f->lf.lim_lut.sharp[1] = fr->lut.sharp[1];
    f->lf.mask = fr->mask;
    hdr->loopfilter.level_u = hdr->loopfilter.level_v = fr->filter_uv;
    hdr->tiling.cols = 1;
    hdr->tiling.col_start_sb[0] = 0;
    hdr->tiling.col_start_sb[1] = 1 << 14;    /* single tile column: loop at lf_apply_tmpl.c:334 exits at once */
    static uint8_t lpf_edge[2][8192];
    f->lf.tx_lpf_right_edge[0] = lpf_edge[0];
    f->lf.tx_lpf_right_edge[1] = lpf_edge[1];
    pixel *const base = fr->pic;
    if (fr->filter_y) {
        for (int sby = 0; sby < f->sbh; sby++) {
            const int y = sby * f->sb_step * 4;
            pixel *const p[3] = {
                base + fr->plane_off[0] + (ptrdiff_t)y * fr->stride[0],
                base + fr->plane_off[1] + (ptrdiff_t)(y >> fr->ss_ver) * fr->stride[1],
                base + fr->plane_off[2] + (ptrdiff_t)(y >> fr->ss_ver) * fr->stride[2],
            };
            Av1Filter *const mask = f->lf.mask + (sby >> !seq->sb128) * f->sb128w;
            bytefn(dav1d_loopfilter_sbrow_cols)(f, p, mask, sby, 0);
            bytefn(dav1d_loopfilter_sbrow_rows)(f, p, mask, sby);
        }
    }
    free(dsp); free(hdr); free(seq); free(f);
}


/* ---- CDEF: the reference's own dav1d_cdef_brow over a fully deblocked picture, in place ---- */
#include "src/cdef_apply.h"
#include "src/cdef.h"
typedef struct {
    void *src, *dst;               /* dst unused: the reference filters `src` in place */
    uint32_t plane_off[3];
    int32_t stride[3];
    int32_t bw, bh, sb128w, ss_hor, ss_ver, damping;
    int32_t y_strength[8], uv_strength[8];
    Av1Filter *mask;
} RefCdefFrame;

API void bitfn(refdrv_cdef_frame)(const int bitdepth_max, const RefCdefFrame *const fr)
{
    Dav1dFrameContext *const f = calloc(1, sizeof(*f));
    Dav1dContext *const c = calloc(1, sizeof(*c));
    Dav1dSequenceHeader *const seq = calloc(1, sizeof(*seq));
    Dav1dFrameHeader *const hdr = calloc(1, sizeof(*hdr));
    Dav1dDSPContext *const dsp = calloc(1, sizeof(*dsp));
    Dav1dTaskContext *const tc = calloc(1, sizeof(*tc));
    bitfn(dav1d_cdef_dsp_init)(&dsp->cdef);
    f->dsp = dsp; f->seq_hdr = seq; f->frame_hdr = hdr; f->c = c;
    c->n_tc = 1;
    tc->f = f;
    f->bw = fr->bw; f->bh = fr->bh; f->sb128w = fr->sb128w;
    f->cur.p.layout = !fr->ss_hor ? DAV1D_PIXEL_LAYOUT_I444 : fr->ss_ver ? DAV1D_PIXEL_LAYOUT_I420 : DAV1D_PIXEL_LAYOUT_I422;
    f->cur.p.bpc = 32 - clz(bitdepth_max);
    f->cur.stride[0] = fr->stride[0] * (ptrdiff_t)sizeof(pixel);
    f->cur.stride[1] = fr->stride[1] * (ptrdiff_t)sizeof(pixel);
#if BITDEPTH == 16
    f->bitdepth_max = bitdepth_max;
#endif
    hdr->cdef.damping = fr->damping;
    for (int i = 0; i < 8; i++) { hdr->cdef.y_strength[i] = fr->y_strength[i]; hdr->cdef.uv_strength[i] = fr->uv_strength[i]; }
    hdr->width[0] = hdr->width[1] = fr->bw * 4;
    /* two-line pre-filter backups, toggled per 8-row band (src/decode.c:2916-2936, have_tt == 0 layout) */
    pixel *const lines = calloc((size_t)(fr->stride[0] * 4 + fr->stride[1] * 8 + 64), sizeof(pixel));
    f->lf.cdef_line[0][0] = lines;
    f->lf.cdef_line[1][0] = lines + fr->stride[0] * 2;
    pixel *const uvl = lines + fr->stride[0] * 4;
    f->lf.cdef_line[0][1] = uvl; f->lf.cdef_line[0][2] = uvl + fr->stride[1] * 2;
    f->lf.cdef_line[1][1] = uvl + fr->stride[1] * 4; f->lf.cdef_line[1][2] = uvl + fr->stride[1] * 6;
    pixel *const base = fr->src;
    for (int r = 0; r * 32 < fr->bh; r++) {
        const int y = r * 128;
        pixel *const p[3] = {
            base + fr->plane_off[0] + (ptrdiff_t)y * fr->stride[0],
            base + fr->plane_off[1] + (ptrdiff_t)(y >> fr->ss_ver) * fr->stride[1],
            base + fr->plane_off[2] + (ptrdiff_t)(y >> fr->ss_ver) * fr->stride[2],
        };
        const int end = r * 32 + 32 < fr->bh ? r * 32 + 32 : fr->bh;
        bytefn(dav1d_cdef_brow)(tc, p, fr->mask + r * fr->sb128w, r * 32, end, 0, 0);
    }
    free(lines); free(tc); free(dsp); free(hdr); free(seq); free(c); free(f);
}

/* ---- loop restoration: the reference's own dav1d_copy_lpf + dav1d_lr_sbrow, in place on `cdef` ---- */
#include "src/lr_apply.h"
#include "src/looprestoration.h"
typedef struct {
    void *cdef; const void *dbl; void *dst;   /* dst unused: the reference restores `cdef` in place */
    uint32_t plane_off[3]; int32_t stride[3];
    int32_t w, h, ss_hor, ss_ver, sb128, sr_sb128w;
    int32_t unit_size_log2[2];
    int32_t restore_planes;
    Av1Restoration *lr_mask;
} RefLrFrame;

API void bitfn(refdrv_lr_frame)(const int bitdepth_max, const RefLrFrame *const fr)
{
    Dav1dFrameContext *const f = calloc(1, sizeof(*f));
    Dav1dContext *const c = calloc(1, sizeof(*c));
    Dav1dSequenceHeader *const seq = calloc(1, sizeof(*seq));
    Dav1dFrameHeader *const hdr = calloc(1, sizeof(*hdr));
    Dav1dDSPContext *const dsp = calloc(1, sizeof(*dsp));
    bitfn(dav1d_loop_restoration_dsp_init)(&dsp->lr, 32 - clz(bitdepth_max));
    f->dsp = dsp; f->seq_hdr = seq; f->frame_hdr = hdr; f->c = c;
    c->n_tc = 1;
    seq->sb128 = fr->sb128; seq->cdef = 1;
    const enum Dav1dPixelLayout layout = !fr->ss_hor ? DAV1D_PIXEL_LAYOUT_I444 : fr->ss_ver ? DAV1D_PIXEL_LAYOUT_I420 : DAV1D_PIXEL_LAYOUT_I422;
    f->cur.p.w = f->sr_cur.p.p.w = fr->w; f->cur.p.h = f->sr_cur.p.p.h = fr->h;
    f->cur.p.layout = f->sr_cur.p.p.layout = layout;
    f->cur.p.bpc = f->sr_cur.p.p.bpc = 32 - clz(bitdepth_max);
    f->cur.stride[0] = f->sr_cur.p.stride[0] = fr->stride[0] * (ptrdiff_t)sizeof(pixel);
    f->cur.stride[1] = f->sr_cur.p.stride[1] = fr->stride[1] * (ptrdiff_t)sizeof(pixel);
#if BITDEPTH == 16
    f->bitdepth_max = bitdepth_max;
#endif
    f->bw = (fr->w + 3) >> 2; f->bh = (fr->h + 3) >> 2;
    f->sb_step = 16 << fr->sb128;
    f->sbh = (f->bh + f->sb_step - 1) / f->sb_step;
    f->sr_sb128w = fr->sr_sb128w;
    f->lf.lr_mask = fr->lr_mask;
    f->lf.restore_planes = fr->restore_planes;
    hdr->width[0] = hdr->width[1] = fr->w;
    hdr->restoration.unit_size[0] = fr->unit_size_log2[0];
    hdr->restoration.unit_size[1] = fr->unit_size_log2[1];
    pixel *lines[3];
    for (int p = 0; p < 3; p++) {
        lines[p] = calloc((size_t)fr->stride[p] * 16 + 64, sizeof(pixel));
        f->lf.lr_lpf_line[p] = lines[p];
    }
    pixel *const C = fr->cdef;
    pixel *const D = (pixel *)fr->dbl;
    for (int sby = 0; sby < f->sbh; sby++) {
        const int y = sby * f->sb_step * 4;
        pixel *const d[3] = { D + fr->plane_off[0] + (ptrdiff_t)y * fr->stride[0],
                              D + fr->plane_off[1] + (ptrdiff_t)(y >> fr->ss_ver) * fr->stride[1],
                              D + fr->plane_off[2] + (ptrdiff_t)(y >> fr->ss_ver) * fr->stride[2] };
        pixel *const s[3] = { C + fr->plane_off[0] + (ptrdiff_t)y * fr->stride[0],
                              C + fr->plane_off[1] + (ptrdiff_t)(y >> fr->ss_ver) * fr->stride[1],
                              C + fr->plane_off[2] + (ptrdiff_t)(y >> fr->ss_ver) * fr->stride[2] };
        bytefn(dav1d_copy_lpf)(f, d, sby);
        bytefn(dav1d_lr_sbrow)(f, s, sby);
    }
    for (int p = 0; p < 3; p++) free(lines[p]);
    free(dsp); free(hdr); free(seq); free(c); free(f);
}

/* ---- whole frame on the CPU with the reference's own functions (bench.py --impl reference,
 *      cpu_baseline): same B200FrameJob records, host pointers, single thread per call ---- */
#include "../../include/b200av1.h"
#include "src/mc.h"
#include "src/itx.h"

static void bitfn(frame_run_fg)(int bitdepth_max, const B200FrameJob *j);
static void bitfn(intra_records)(int bitdepth_max, const B200IntraFrame *fr, const B200IntraTx *tx, int n);
API void bitfn(refdrv_frame_run)(const B200FrameJob *const j)
{
    static __thread Dav1dMCDSPContext mc;
    static __thread Dav1dInvTxfmDSPContext itx;
    static __thread int inited;
    if (!inited) {
        bitfn(dav1d_mc_dsp_init)(&mc);
        bitfn(dav1d_itx_dsp_init)(&itx, BITDEPTH == 8 ? 8 : 32 - clz(j->bitdepth_max));
        inited = 1;
    }
    const int bitdepth_max = j->bitdepth_max;
    (void)bitdepth_max;
    pixel *const dst = j->mc.dst;
    ALIGN_STK_64(pixel, emu, 192 * 160,);
    /* prediction: emu_edge when the footprint leaves the picture, like mc() in src/recon_tmpl.c:956-988 */
    for (int i = 0; i < j->n_pred; i++) {
        const B200McBlock *b = &j->d_pred[i];
        const int pl = b->plane, w = b->w, h = b->h, mx = b->mx, my = b->my;
        const pixel *ref = (const pixel *)j->mc.ref[b->ref] + j->mc.ref_plane_off[pl];
        ptrdiff_t rs = j->mc.ref_stride[pl] * (ptrdiff_t)sizeof(pixel);
        const int dx = b->src_x, dy = b->src_y, iw = j->mc.ref_w[pl], ih = j->mc.ref_h[pl];
        const pixel *src;
        if (dx < !!mx * 3 || dy < !!my * 3 || dx + w + !!mx * 4 > iw || dy + h + !!my * 4 > ih) {
            mc.emu_edge(w + !!mx * 7, h + !!my * 7, iw, ih, dx - !!mx * 3, dy - !!my * 3, emu, 192 * sizeof(pixel), ref, rs);
            src = &emu[192 * !!my * 3 + !!mx * 3];
            rs = 192 * sizeof(pixel);
        } else {
            src = ref + (ptrdiff_t)dy * j->mc.ref_stride[pl] + dx;
        }
        if (b->op == 2)       /* obmc()'s `lap` prediction: a put into the pixel scratch, pitch w (src/recon_tmpl.c:1052-1113) */
            mc.mc[b->filter2d]((pixel *)j->mc.px_tmp + b->dst_off, w * (ptrdiff_t)sizeof(pixel), src, rs, w, h, mx, my HIGHBD_TAIL_SUFFIX);
        else if (b->op) mc.mct[b->filter2d](j->mc.tmp + b->dst_off, src, rs, w, h, mx, my HIGHBD_TAIL_SUFFIX);
        else mc.mc[b->filter2d](dst + b->dst_off, j->mc.dst_stride[pl] * (ptrdiff_t)sizeof(pixel), src, rs, w, h, mx, my HIGHBD_TAIL_SUFFIX);
    }
    /* warped blocks: emu_edge of the 15x15 window when it leaves the picture, then warp8x8 / warp8x8t (warp_affine,
     * src/recon_tmpl.c:1115-1165) */
    for (int i = 0; i < j->n_warp; i++) {
        const B200WarpBlock *b = &j->d_warp[i];
        const int pl = b->plane, iw = j->mc.ref_w[pl], ih = j->mc.ref_h[pl], dx = b->src_x, dy = b->src_y;
        const pixel *ref = (const pixel *)j->mc.ref[b->ref] + j->mc.ref_plane_off[pl];
        ptrdiff_t rs = j->mc.ref_stride[pl] * (ptrdiff_t)sizeof(pixel);
        const pixel *src;
        if (dx < 3 || dx + 8 + 4 > iw || dy < 3 || dy + 8 + 4 > ih) {
            mc.emu_edge(15, 15, iw, ih, dx - 3, dy - 3, emu, 32 * sizeof(pixel), ref, rs);
            src = &emu[32 * 3 + 3];
            rs = 32 * sizeof(pixel);
        } else {
            src = ref + (ptrdiff_t)dy * j->mc.ref_stride[pl] + dx;
        }
        if (b->op) mc.warp8x8t(j->mc.tmp + b->dst_off, b->tmp_stride, src, rs, b->abcd, b->mx, b->my HIGHBD_TAIL_SUFFIX);
        else mc.warp8x8(dst + b->dst_off, j->mc.dst_stride[pl] * (ptrdiff_t)sizeof(pixel), src, rs, b->abcd, b->mx, b->my HIGHBD_TAIL_SUFFIX);
    }
    for (int pass = 0; pass < 2; pass++) {
        const B200CompBlock *cb = pass ? j->d_comp2 : j->d_comp;
        const int n = pass ? j->n_comp2 : j->n_comp;
        for (int i = 0; i < n; i++, cb++) {
            pixel *d = dst + cb->dst_off;
            const ptrdiff_t ds = j->mc.dst_stride[cb->plane] * (ptrdiff_t)sizeof(pixel);
            const int16_t *t1 = j->mc.tmp + cb->tmp1_off, *t2 = j->mc.tmp + cb->tmp2_off;
            uint8_t *m = j->mc.mask + cb->mask_off;
            switch (cb->op) {
            case B200_COMP_AVG:   mc.avg(d, ds, t1, t2, cb->w, cb->h HIGHBD_TAIL_SUFFIX); break;
            case B200_COMP_W_AVG: mc.w_avg(d, ds, t1, t2, cb->w, cb->h, cb->param HIGHBD_TAIL_SUFFIX); break;
            case B200_COMP_MASK:  mc.mask(d, ds, t1, t2, cb->w, cb->h, m HIGHBD_TAIL_SUFFIX); break;
            default: mc.w_mask[cb->op - B200_COMP_W_MASK_444](d, ds, t1, t2, cb->w, cb->h, m, cb->param HIGHBD_TAIL_SUFFIX); break;
            }
        }
    }
    /* blends: every blend_h (predictions of the blocks above), then every blend_v (blocks to the left), as obmc() orders them
     * inside a block; blocks do not overlap, so the order across blocks is free */
    for (int pass = 0; pass < 2; pass++) {
        const B200BlendBlock *bb = pass ? j->d_blend2 : j->d_blend;
        const int n = pass ? j->n_blend2 : j->n_blend;
        for (int i = 0; i < n; i++, bb++) {
            pixel *d = dst + bb->dst_off;
            const ptrdiff_t ds = j->mc.dst_stride[bb->plane] * (ptrdiff_t)sizeof(pixel);
            const pixel *lap = (const pixel *)j->mc.px_tmp + bb->tmp_off;
            if (bb->op == B200_BLEND_H) mc.blend_h(d, ds, lap, bb->w, bb->h);
            else if (bb->op == B200_BLEND_V) mc.blend_v(d, ds, lap, bb->w, bb->h);
            else mc.blend(d, ds, lap, bb->w, bb->h, j->mc.mask + bb->mask_off);
        }
    }
    for (int tx = 0; tx < N_RECT_TX_SIZES; tx++) {
        const TxfmInfo *t = &dav1d_txfm_dimensions[tx];
        const int n_cf = imin(t->w * 4, 32) * imin(t->h * 4, 32);
        for (int i = 0; i < j->n_itx[tx]; i++) {
            const B200ItxBlock *b = &j->d_itx[tx][i];
            coef *cf = (coef *)j->d_coef + b->coef_off, save[1024];
            memcpy(save, cf, sizeof(coef) * n_cf);
            itx.itxfm_add[tx][b->txtp](dst + b->dst_off, j->itx_stride[b->plane] * (ptrdiff_t)sizeof(pixel), cf, b->eob HIGHBD_TAIL_SUFFIX);
            memcpy(cf, save, sizeof(coef) * n_cf);
        }
    }
    if (j->n_intra > 0) bitfn(intra_records)(bitdepth_max, &j->intra, j->d_intra, j->n_intra);
    /* the three frame drivers take frame structs with the same leading layout as the B200 ones */
    if (j->run_lf) {
        RefLfFrame lf;
        memset(&lf, 0, sizeof(lf));
        lf.pic = j->lf.pic;
        for (int p = 0; p < 3; p++) { lf.plane_off[p] = j->lf.plane_off[p]; lf.stride[p] = j->lf.stride[p]; }
        lf.w4 = j->lf.w4; lf.h4 = j->lf.h4; lf.sb128w = j->lf.sb128w; lf.b4_stride = j->lf.b4_stride;
        lf.ss_hor = j->lf.ss_hor; lf.ss_ver = j->lf.ss_ver; lf.sb128 = j->lf.sb128; lf.filter_y = j->lf.filter_y; lf.filter_uv = j->lf.filter_uv;
        lf.mask = (Av1Filter *)j->lf.mask; lf.level = (uint8_t (*)[4])j->lf.level;
        memcpy(&lf.lut, &j->lf.lut, sizeof(lf.lut));
        bitfn(refdrv_lf_frame)(bitdepth_max, &lf);
    }
    size_t pic_bytes = 0;
    {   /* extent of the picture allocation: last plane offset + its rows */
        const int ssv = j->lf.ss_ver;
        const int rows = (((j->lf.h4 * 4 + 127) & ~127) >> ssv);
        pic_bytes = ((size_t)j->lf.plane_off[2] + (size_t)rows * j->lf.stride[2]) * sizeof(pixel);
    }
    if (j->run_cdef) {
        memcpy(j->cdef.dst, j->cdef.src, pic_bytes);
        RefCdefFrame cd;
        memset(&cd, 0, sizeof(cd));
        cd.src = j->cdef.dst;
        for (int p = 0; p < 3; p++) { cd.plane_off[p] = j->cdef.plane_off[p]; cd.stride[p] = j->cdef.stride[p]; }
        cd.bw = j->cdef.bw; cd.bh = j->cdef.bh; cd.sb128w = j->cdef.sb128w; cd.ss_hor = j->cdef.ss_hor; cd.ss_ver = j->cdef.ss_ver;
        cd.damping = j->cdef.damping;
        for (int i = 0; i < 8; i++) { cd.y_strength[i] = j->cdef.y_strength[i]; cd.uv_strength[i] = j->cdef.uv_strength[i]; }
        cd.mask = (Av1Filter *)j->cdef.mask;
        bitfn(refdrv_cdef_frame)(bitdepth_max, &cd);
    }
    if (j->run_lr) {
        memcpy(j->lr.dst, j->lr.cdef, pic_bytes);
        RefLrFrame lr;
        memset(&lr, 0, sizeof(lr));
        lr.cdef = j->lr.dst; lr.dbl = j->lr.dbl;
        for (int p = 0; p < 3; p++) { lr.plane_off[p] = j->lr.plane_off[p]; lr.stride[p] = j->lr.stride[p]; }
        lr.w = j->lr.w; lr.h = j->lr.h; lr.ss_hor = j->lr.ss_hor; lr.ss_ver = j->lr.ss_ver; lr.sb128 = j->lr.sb128;
        lr.sr_sb128w = j->lr.sr_sb128w; lr.unit_size_log2[0] = j->lr.unit_size_log2[0]; lr.unit_size_log2[1] = j->lr.unit_size_log2[1];
        lr.restore_planes = j->lr.restore_planes; lr.lr_mask = (Av1Restoration *)j->lr.lr_mask;
        bitfn(refdrv_lr_frame)(bitdepth_max, &lr);
    }
    if (j->run_fg) bitfn(frame_run_fg)(bitdepth_max, j);
}

/* ---- film grain: the reference's own dav1d_apply_grain (prep + every 32-row strip) ---- */
#include "src/fg_apply.h"
#include "src/filmgrain.h"
typedef struct {
    const void *in; void *out;
    uint32_t plane_off[3]; int32_t stride[3];
    int32_t w, h, ss_hor, ss_ver, is_id;
    Dav1dFilmGrainData data;
} RefFgFrame;

API void bitfn(refdrv_fg_frame)(const int bitdepth_max, const RefFgFrame *const fr)
{
    Dav1dFilmGrainDSPContext dsp;
    bitfn(dav1d_film_grain_dsp_init)(&dsp);
    Dav1dPicture in, out;
    Dav1dFrameHeader *const hdr = calloc(1, sizeof(*hdr));
    Dav1dSequenceHeader *const seq = calloc(1, sizeof(*seq));
    memset(&in, 0, sizeof(in)); memset(&out, 0, sizeof(out));
    hdr->film_grain.data = fr->data;
    seq->mtrx = fr->is_id ? DAV1D_MC_IDENTITY : DAV1D_MC_BT709;
    const enum Dav1dPixelLayout layout = !fr->ss_hor ? DAV1D_PIXEL_LAYOUT_I444 : fr->ss_ver ? DAV1D_PIXEL_LAYOUT_I420 : DAV1D_PIXEL_LAYOUT_I422;
    Dav1dPicture *pics[2] = { &in, &out };
    for (int i = 0; i < 2; i++) {
        Dav1dPicture *p = pics[i];
        pixel *base = (pixel *)(i ? fr->out : (void *)fr->in);
        for (int pl = 0; pl < 3; pl++) p->data[pl] = base + fr->plane_off[pl];
        p->stride[0] = fr->stride[0] * (ptrdiff_t)sizeof(pixel);
        p->stride[1] = fr->stride[1] * (ptrdiff_t)sizeof(pixel);
        p->p.w = fr->w; p->p.h = fr->h; p->p.layout = layout; p->p.bpc = 32 - clz(bitdepth_max);
        p->frame_hdr = hdr; p->seq_hdr = seq;
    }
    bitfn(dav1d_apply_grain)(&dsp, &out, &in);
    free(hdr); free(seq);
}

static void bitfn(frame_run_fg)(const int bitdepth_max, const B200FrameJob *const j)
{
    RefFgFrame fr;
    memset(&fr, 0, sizeof(fr));
    fr.in = j->fg.in; fr.out = j->fg.out;
    for (int p = 0; p < 3; p++) { fr.plane_off[p] = j->fg.plane_off[p]; fr.stride[p] = j->fg.stride[p]; }
    fr.w = j->fg.w; fr.h = j->fg.h; fr.ss_hor = j->fg.ss_hor; fr.ss_ver = j->fg.ss_ver; fr.is_id = j->fg.is_id;
    memcpy(&fr.data, &j->fg.data, sizeof(fr.data));
    bitfn(refdrv_fg_frame)(bitdepth_max, &fr);
}

/* ---- intra reconstruction: the reference's own dav1d_prepare_intra_edges + intra_pred / cfl / itxfm_add per
 * transform-block record, in record order (what dav1d_recon_b_intra does per tx block, src/recon_tmpl.c:1235-1330,
 * 1342-1398, 1418-1540). The top edge is read from the picture: with whole-frame reconstruction the row above is
 * still unfiltered, which is what f->ipred_edge preserves in the sbrow pipeline. */
#include "src/ipred.h"
#include "src/ipred_prepare.h"
static void bitfn(intra_records)(const int bitdepth_max, const B200IntraFrame *const fr, const B200IntraTx *const tx, const int n)
{
    Dav1dIntraPredDSPContext ip;
    Dav1dInvTxfmDSPContext itx;
    Dav1dMCDSPContext mcd;
    bitfn(dav1d_mc_dsp_init)(&mcd);
    bitfn(dav1d_intra_pred_dsp_init)(&ip);
    bitfn(dav1d_itx_dsp_init)(&itx, 32 - clz(bitdepth_max));
    pixel edge_buf[257];
    pixel *const edge = edge_buf + 128;
    int16_t ac[32 * 32];
    const int layout_idx = !fr->ss_hor ? 2 : fr->ss_ver ? 0 : 1;
    for (int i = 0; i < n; i++) {
        const B200IntraTx *const r = &tx[i];
        const TxfmInfo *const t = &dav1d_txfm_dimensions[r->tx];
        pixel *const dst = (pixel *)fr->pic + r->dst_off;
        const ptrdiff_t stride = fr->stride[r->plane] * (ptrdiff_t)sizeof(pixel);
        const int have_left = !!(r->flags & B200_INTRA_HAVE_LEFT), have_top = !!(r->flags & B200_INTRA_HAVE_TOP);
        const enum EdgeFlags ef = ((r->flags & B200_INTRA_TOP_HAS_RIGHT) ? EDGE_I444_TOP_HAS_RIGHT : 0) |
                                  ((r->flags & B200_INTRA_LEFT_HAS_BOTTOM) ? EDGE_I444_LEFT_HAS_BOTTOM : 0);
        int angle = r->angle;
        if (r->mode == B200_INTRA_MODE_RESID) {
            /* residual only */
        } else if (r->mode == B200_INTRA_MODE_PAL) {
            /* the reference's own pal_pred over the whole block (src/recon_tmpl.c:1220) */
            const pixel *const pal = (const pixel *)(fr->pal + r->luma_off);
            ip.pal_pred(dst, stride, pal, (const uint8_t *)(pal + 8), t->w * 4, t->h * 4);
        } else if (r->mode == B200_INTRA_MODE_IBC) {
            /* the reference's own mc() route for intra block copy (src/recon_tmpl.c:1583-1596, 956-988): emu_edge when the
             * footprint leaves the plane area, then mc[FILTER_2D_BILINEAR] */
            ALIGN_STK_64(pixel, emu, 192 * 72,);
            const int pl = r->plane, w = t->w * 4, h = t->h * 4, mx = r->cfl_w_pad, my = r->cfl_h_pad;
            const int dx = (int)(r->luma_off & 0xffff), dy = (int)(r->luma_off >> 16), iw = fr->w4[pl] * 4, ih = fr->h4[pl] * 4;
            const pixel *const plane = (const pixel *)fr->pic + r->dst_off - ((ptrdiff_t)r->y4 * 4 * fr->stride[pl] + r->x4 * 4);
            const pixel *src;
            ptrdiff_t rs = stride;
            if (dx < !!mx * 3 || dy < !!my * 3 || dx + w + !!mx * 4 > iw || dy + h + !!my * 4 > ih) {
                mcd.emu_edge(w + !!mx * 7, h + !!my * 7, iw, ih, dx - !!mx * 3, dy - !!my * 3, emu, 192 * sizeof(pixel), plane, stride);
                src = &emu[192 * !!my * 3 + !!mx * 3];
                rs = 192 * sizeof(pixel);
            } else {
                src = plane + (ptrdiff_t)dy * fr->stride[pl] + dx;
            }
            mcd.mc[FILTER_2D_BILINEAR](dst, stride, src, rs, w, h, mx, my HIGHBD_TAIL_SUFFIX);
        } else if (r->mode == B200_INTRA_MODE_II) {
            /* the reference's own edge preparation, predictor and blend (src/recon_tmpl.c:1601-1626) */
            pixel tmp[64 * 64];
            int a0 = 0;
            const enum IntraPredMode m = bytefn(dav1d_prepare_intra_edges)(r->x4, have_left, r->y4, have_top, r->xend4, r->yend4, 0, dst, stride,
                                                                          NULL, (enum IntraPredMode)r->angle, &a0, t->w, t->h, 0,
                                                                          edge HIGHBD_TAIL_SUFFIX);
            ip.intra_pred[m](tmp, t->w * 4 * sizeof(pixel), edge, t->w * 4, t->h * 4, 0, 0, 0 HIGHBD_TAIL_SUFFIX);
            mcd.blend(dst, stride, tmp, t->w * 4, t->h * 4, fr->mask + r->luma_off);
        } else if (r->mode == B200_INTRA_MODE_CFL && r->cfl_alpha) {
            angle = 0;
            ip.cfl_ac[layout_idx](ac, (const pixel *)fr->pic + r->luma_off, fr->stride[0] * (ptrdiff_t)sizeof(pixel),
                                  r->cfl_w_pad, r->cfl_h_pad, t->w * 4, t->h * 4);
            const enum IntraPredMode m = bytefn(dav1d_prepare_intra_edges)(r->x4, have_left, r->y4, have_top, r->xend4, r->yend4,
                                                                          0, dst, stride, NULL, DC_PRED, &angle, t->w, t->h, 0,
                                                                          edge HIGHBD_TAIL_SUFFIX);
            ip.cfl_pred[m](dst, stride, edge, t->w * 4, t->h * 4, ac, r->cfl_alpha HIGHBD_TAIL_SUFFIX);
        } else {
            const enum IntraPredMode in = r->mode == B200_INTRA_MODE_CFL ? DC_PRED : (enum IntraPredMode)r->mode;
            const enum IntraPredMode m = bytefn(dav1d_prepare_intra_edges)(r->x4, have_left, r->y4, have_top, r->xend4, r->yend4,
                                                                          ef, dst, stride, NULL, in, &angle, t->w, t->h,
                                                                          (r->angle_flags >> 10) & 1, edge HIGHBD_TAIL_SUFFIX);
            ip.intra_pred[m](dst, stride, edge, t->w * 4, t->h * 4, angle | r->angle_flags, r->max_w, r->max_h HIGHBD_TAIL_SUFFIX);
        }
        if (r->eob >= 0) {
            const int n_cf = imin(t->w * 4, 32) * imin(t->h * 4, 32);
            coef *cf = (coef *)fr->d_coef + r->coef_off, save[1024];
            memcpy(save, cf, sizeof(coef) * n_cf);
            itx.itxfm_add[r->tx][r->txtp](dst, stride, cf, r->eob HIGHBD_TAIL_SUFFIX);
            if (!fr->zero_coefs) memcpy(cf, save, sizeof(coef) * n_cf);
        }
    }
}
API void bitfn(refdrv_intra_frame)(const int bitdepth_max, const B200IntraFrame *const fr, const B200IntraTx *const tx, const int n)
{
    bitfn(intra_records)(bitdepth_max, fr, tx, n);
}
