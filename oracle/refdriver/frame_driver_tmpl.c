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
