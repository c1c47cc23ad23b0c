/*
 * integration/dav1d/b200_hooks_tmpl.c — the dav1d `f->bd_fn` hooks (reference src/internal.h:247-262, typedefs
 * src/recon.h:39-70) as B200 record emitters. Compiled at BITDEPTH 8 and 16 against dav1d's internal headers.
 *
 * dav1d runs in its two-pass (frame-threaded) mode: pass 1 entropy-decodes every block into
 * f->frame_thread.{b,cbi,cf}; pass 2 calls the hooks below, which translate each block into the records of
 * include/b200av1.h instead of reconstructing it on the CPU:
 *   recon_b_intra   -> one B200IntraTx per transform block (mode, angle, edge availability, CFL parameters,
 *                      transform type / eob) + its dequantised coefficients copied into a pinned staging buffer
 *                      (what dav1d_recon_b_intra does per tx block, reference src/recon_tmpl.c:1176-1555)
 *   backup_ipred_edge -> "tile superblock row complete" marker; when the last one of a frame arrives the frame's
 *                      records, dav1d's own Av1Filter / level / Av1Restoration arrays and the frame-header
 *                      parameters are shipped to HBM, b200_frame_run_host reconstructs and filters the whole frame,
 *                      and the finished picture is copied into f->cur (what the output / reference logic reads)
 *   filter_sbrow_*  -> nothing left to do on the CPU (the device job already ran the whole post-filter sweep)
 *   recon_b_inter   -> not translated yet: the tile fails loudly (no CPU fallback)
 * Palette blocks are not translated yet either (streams with allow_screen_content_tools fail loudly).
 */
#include "config.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "common/attributes.h"
#include "common/bitdepth.h"
#include "common/intops.h"
#include "src/internal.h"
#include "src/ipred_prepare.h"
#include "src/recon.h"
#include "src/tables.h"
#include "b200_hooks.h"

static inline double bitfn(now_ms)(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* device picture geometry derived from the host picture: same strides, planes back to back */
typedef struct PicGeom { int stride[3]; uint32_t off[3]; int rows[3]; size_t bytes; } PicGeom;
static void bitfn(pic_geom)(const Dav1dFrameContext *const f, PicGeom *const g)
{
    const int ss_ver = f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420;
    const int rows = (f->cur.p.h + 127) & ~127;
    g->stride[0] = (int)PXSTRIDE(f->cur.stride[0]);
    g->stride[1] = g->stride[2] = (int)PXSTRIDE(f->cur.stride[1]);
    g->rows[0] = rows; g->rows[1] = g->rows[2] = rows >> ss_ver;
    g->off[0] = 0;
    g->off[1] = (uint32_t)g->stride[0] * rows;
    g->off[2] = g->off[1] + (uint32_t)g->stride[1] * g->rows[1];
    g->bytes = ((size_t)g->off[2] + (size_t)g->stride[2] * g->rows[2]) * sizeof(pixel);
}

/* ---- one transform block -> one record ------------------------------------------------------------------ */
typedef struct TxCtx {
    HookFrame *hf;
    const Dav1dTaskContext *t;
    const Av1Block *b;
    PicGeom g;
} TxCtx;

/* copies the block's coefficients out of dav1d's pass-1 buffer (and clears them there, as the reference's
 * inverse transform would have: the buffer must be all zero for the next frame's pass 1) */
static int bitfn(stage_coefs)(HookFrame *const hf, coef *const cf, const int n, uint32_t *const off)
{
    if (b200hook_buf_reserve(&hf->coef, (hf->n_coef + n) * sizeof(coef), 1, 1)) return -1;
    memcpy((coef *)hf->coef.host + hf->n_coef, cf, n * sizeof(coef));
    memset(cf, 0, n * sizeof(coef));
    *off = (uint32_t)hf->n_coef;
    hf->n_coef += n;
    return 0;
}

static B200IntraTx *bitfn(new_record)(HookFrame *const hf)
{
    if (b200hook_buf_reserve(&hf->tx, (size_t)(hf->n_tx + 1) * sizeof(B200IntraTx), 1, 1)) return NULL;
    B200IntraTx *const r = (B200IntraTx *)hf->tx.host + hf->n_tx++;
    memset(r, 0, sizeof(*r));
    return r;
}

/* the residual of the transform block the tile's cbi / cf cursors point at (pass 2: reference
 * src/recon_tmpl.c:1296-1302, 1508-1514); advances the cursors exactly like the reference */
static int bitfn(take_residual)(TxCtx *const c, B200IntraTx *const r, const TxfmInfo *const td, const int chroma)
{
    Dav1dTileState *const ts = c->t->ts;
    const int p = c->t->frame_thread.pass & 1;
    r->eob = -1;
    if (c->b->skip) return 0;
    const int cbi = *ts->frame_thread[p].cbi++;
    coef *const cf = ts->frame_thread[p].cf;
    const int n = chroma ? td->w * td->h * 16 : imin(td->w, 8) * imin(td->h, 8) * 16;
    ts->frame_thread[p].cf = cf + n;
    r->eob = (int16_t)(cbi >> 5);
    r->txtp = (uint8_t)(cbi & 0x1f);
    if (r->eob >= 0) return bitfn(stage_coefs)(c->hf, cf, n, &r->coef_off);
    return 0;
}

void bitfn(b200hook_recon_b_intra)(Dav1dTaskContext *const t, const enum BlockSize bs,
                                   const enum EdgeFlags intra_edge_flags, const Av1Block *const b)
{
    const Dav1dFrameContext *const f = t->f;
    Dav1dTileState *const ts = t->ts;
    HookFrame *const hf = b200hook_frame(f);
    if (!hf) return;
    if (t->frame_thread.pass != 2) {
        /* single-pass decoding interleaves entropy decoding with reconstruction inside this hook; the B200 back end
         * needs dav1d's two-pass mode (n_threads >= 2 with max_frame_delay >= 2, reference src/lib.c get_num_threads) */
        pthread_mutex_lock(&hf->lock); hf->unsupported |= 4; pthread_mutex_unlock(&hf->lock);
        return;
    }
    TxCtx c = { hf, t, b };
    bitfn(pic_geom)(f, &c.g);
    const int ss_ver = f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420;
    const int ss_hor = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444;
    const int bx = t->bx, by = t->by, bx4 = bx & 31, by4 = by & 31;
    const uint8_t *const dim = dav1d_block_dimensions[bs];
    const int bw4 = dim[0], bh4 = dim[1];
    const int w4 = imin(bw4, f->bw - bx), h4 = imin(bh4, f->bh - by);
    const int cw4 = (w4 + ss_hor) >> ss_hor, ch4 = (h4 + ss_ver) >> ss_ver;
    const int cbw4 = (bw4 + ss_hor) >> ss_hor, cbh4 = (bh4 + ss_ver) >> ss_ver;
    const int has_chroma = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I400 && (bw4 > ss_hor || bx & 1) && (bh4 > ss_ver || by & 1);
    const TxfmInfo *const yt = &dav1d_txfm_dimensions[b->tx], *const ct = &dav1d_txfm_dimensions[b->uvtx];
    const int edge_filter_bit = f->seq_hdr->intra_edge_filter << 10;
    const int layout_shift = f->cur.p.layout - 1;      /* EDGE_I420_* >> (layout - 1) selects this layout's chroma flags */

    pthread_mutex_lock(&hf->lock);
    if (b->pal_sz[0] || (has_chroma && b->pal_sz[1])) hf->unsupported |= 1;
    /* the reference walks a block in 64x64-luma chunks: luma transform blocks of the chunk, then its chroma */
    for (int iy = 0; iy < h4; iy += 16) {
        const int y_end = imin(h4, iy + 16), cy_end = imin(ch4, (iy + 16) >> ss_ver);
        for (int ix = 0; ix < w4; ix += 16) {
            const int x_end = imin(w4, ix + 16), cx_end = imin(cw4, (ix + 16) >> ss_hor);
            /* ---- luma ---- */
            const int y_flags = sm_flag(t->a, bx4) | sm_flag(&t->l, by4) | edge_filter_bit;
            const int chunk_tr = ix + 16 < w4 ? 1 : iy ? 0 : !!(intra_edge_flags & EDGE_I444_TOP_HAS_RIGHT);
            const int chunk_bl = ix ? 0 : iy + 16 < h4 ? 1 : !!(intra_edge_flags & EDGE_I444_LEFT_HAS_BOTTOM);
            for (int y = iy; y < y_end; y += yt->h)
                for (int x = ix; x < x_end; x += yt->w) {
                    B200IntraTx *const r = bitfn(new_record)(hf);
                    if (!r) { hf->unsupported |= 8; goto out; }
                    const int px = bx + x, py = by + y;
                    r->plane = 0; r->tx = b->tx;
                    r->x4 = px; r->y4 = py; r->xend4 = ts->tiling.col_end; r->yend4 = ts->tiling.row_end;
                    r->dst_off = c.g.off[0] + (uint32_t)(4 * py) * c.g.stride[0] + 4 * px;
                    r->mode = b->y_mode; r->angle = b->y_angle;
                    r->angle_flags = y_flags;
                    r->max_w = 4 * f->bw - 4 * px; r->max_h = 4 * f->bh - 4 * py;
                    const int last_col = x + yt->w >= x_end, last_row = y + yt->h >= y_end;
                    r->flags = (px > ts->tiling.col_start ? B200_INTRA_HAVE_LEFT : 0) |
                               (py > ts->tiling.row_start ? B200_INTRA_HAVE_TOP : 0) |
                               (((y > iy || !chunk_tr) && last_col) ? 0 : B200_INTRA_TOP_HAS_RIGHT) |
                               ((x > ix || (!chunk_bl && last_row)) ? 0 : B200_INTRA_LEFT_HAS_BOTTOM);
                    if (bitfn(take_residual)(&c, r, yt, 0)) { hf->unsupported |= 8; goto out; }
                }
            if (!has_chroma) continue;
            /* ---- chroma ---- */
            const int is_cfl = b->uv_mode == CFL_PRED;
            const int uv_flags = sm_uv_flag(t->a, bx4 >> ss_hor) | sm_uv_flag(&t->l, by4 >> ss_ver) | edge_filter_bit;
            const int uv_tr = ((ix + 16) >> ss_hor) < cw4 ? 1 : iy ? 0 :
                              !!(intra_edge_flags & (EDGE_I420_TOP_HAS_RIGHT >> layout_shift));
            const int uv_bl = ix ? 0 : ((iy + 16) >> ss_ver) < ch4 ? 1 :
                              !!(intra_edge_flags & (EDGE_I420_LEFT_HAS_BOTTOM >> layout_shift));
            /* CFL: the ac block is derived from the whole co-located luma block, padded past the frame edge
             * (reference :1342-1362; a CFL block is a single chunk and a single chroma transform block) */
            int cfl_wpad = 0, cfl_hpad = 0;
            if (is_cfl) {
                const int far_r = ((cw4 << ss_hor) + yt->w - 1) & ~(yt->w - 1);
                const int far_b = ((ch4 << ss_ver) + yt->h - 1) & ~(yt->h - 1);
                cfl_wpad = cbw4 - (far_r >> ss_hor); cfl_hpad = cbh4 - (far_b >> ss_ver);
            }
            for (int pl = 1; pl <= 2; pl++)
                for (int y = iy >> ss_ver; y < cy_end; y += ct->h)
                    for (int x = ix >> ss_hor; x < cx_end; x += ct->w) {
                        B200IntraTx *const r = bitfn(new_record)(hf);
                        if (!r) { hf->unsupported |= 8; goto out; }
                        /* luma-unit position the reference's t->bx / t->by would hold here */
                        const int lx = bx + (x << ss_hor), ly = by + (y << ss_ver);
                        const int px = lx >> ss_hor, py = ly >> ss_ver;
                        r->plane = pl; r->tx = b->uvtx;
                        r->x4 = px; r->y4 = py;
                        r->xend4 = ts->tiling.col_end >> ss_hor; r->yend4 = ts->tiling.row_end >> ss_ver;
                        r->dst_off = c.g.off[pl] + (uint32_t)(4 * py) * c.g.stride[pl] + 4 * px;
                        r->max_w = (4 * f->bw + ss_hor - 4 * (lx & ~ss_hor)) >> ss_hor;
                        r->max_h = (4 * f->bh + ss_ver - 4 * (ly & ~ss_ver)) >> ss_ver;
                        r->angle_flags = uv_flags;
                        const int last_col = x + ct->w >= cx_end, last_row = y + ct->h >= cy_end;
                        r->flags = (px > (ts->tiling.col_start >> ss_hor) ? B200_INTRA_HAVE_LEFT : 0) |
                                   (py > (ts->tiling.row_start >> ss_ver) ? B200_INTRA_HAVE_TOP : 0);
                        if (is_cfl) {
                            r->mode = B200_INTRA_MODE_CFL;
                            r->cfl_alpha = b->cfl_alpha[pl - 1];
                            r->cfl_w_pad = cfl_wpad; r->cfl_h_pad = cfl_hpad;
                            r->luma_off = c.g.off[0] + (uint32_t)(4 * (by & ~ss_ver)) * c.g.stride[0] + 4 * (bx & ~ss_hor);
                        } else {
                            r->mode = b->uv_mode; r->angle = b->uv_angle;
                        }
                        if (!is_cfl || !r->cfl_alpha)      /* alpha == 0 is a plain DC_PRED with the usual edge rules */
                            r->flags |= (((y > (iy >> ss_ver) || !uv_tr) && last_col) ? 0 : B200_INTRA_TOP_HAS_RIGHT) |
                                        ((x > (ix >> ss_hor) || (!uv_bl && last_row)) ? 0 : B200_INTRA_LEFT_HAS_BOTTOM);
                        if (bitfn(take_residual)(&c, r, ct, 1)) { hf->unsupported |= 8; goto out; }
                    }
        }
    }
out:
    pthread_mutex_unlock(&hf->lock);
}

int bitfn(b200hook_recon_b_inter)(Dav1dTaskContext *const t, const enum BlockSize bs, const Av1Block *const b)
{
    (void)bs; (void)b;
    HookFrame *const hf = b200hook_frame(t->f);
    if (hf) { pthread_mutex_lock(&hf->lock); hf->unsupported |= 2; pthread_mutex_unlock(&hf->lock); }
    return -1;      /* aborts the tile (reference src/decode.c:771): inter blocks are not translated yet */
}

/* ---- frame completion ------------------------------------------------------------------------------------- */

/* moves the edge at bit `bit` of a (classes x 2 halves) mask row into class min(current class, cap) */
static inline void bitfn(cap_edge_class)(uint16_t (*const m)[2], const int n_cls, const int bit, const int half_bits, const int cap)
{
    const int half = bit >= half_bits;
    const unsigned sel = 1u << (bit - half * half_bits);
    int cls = 0;
    for (int k = n_cls - 1; k > 0; k--)
        if (m[k][half] & sel) { cls = k; break; }
    for (int k = 0; k < n_cls; k++) m[k][half] &= ~sel;
    m[imin(cls, cap)][half] |= sel;
}

/* The deblocking masks pass 1 built describe each tile on its own; at tile boundaries the edge class is limited by
 * the transform size on the other side, which the reference patches in when it filters a superblock row
 * (reference src/lf_apply_tmpl.c:331-401). Same patch, applied to our staged copy for every superblock row. */
static void bitfn(fix_tile_edges)(const Dav1dFrameContext *const f, Av1Filter *const masks)
{
    const int is_sb64 = !f->seq_hdr->sb128, sbsz = 32 >> is_sb64, sbl2 = 5 - is_sb64;
    const int ss_ver = f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420, ss_hor = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444;
    const int has_chroma = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I400;
    const int halign = (f->bh + 31) & ~31;
    for (int sby = 0; sby < f->sbh; sby++) {
        Av1Filter *const row = masks + (sby >> is_sb64) * f->sb128w;
        const int y0 = (sby & is_sb64) << 4, y1 = y0 + imin(f->h4 - sby * sbsz, sbsz);
        const int cy0 = y0 >> ss_ver, cy1 = (y1 + ss_ver) >> ss_ver;
        for (int tc = 1; ; tc++) {
            const int sbx = f->frame_hdr->tiling.col_start_sb[tc];
            if ((sbx << sbl2) >= f->bw) break;
            const int col4 = (sbx & is_sb64) ? 16 : 0;
            Av1Filter *const m = &row[sbx >> is_sb64];
            const uint8_t *const cap_y = &f->lf.tx_lpf_right_edge[0][(sby << sbl2) + (size_t)halign * (tc - 1)];
            for (int y = y0; y < y1; y++)
                bitfn(cap_edge_class)(m->filter_y[0][col4], 3, y, 16, cap_y[y - y0]);
            if (has_chroma) {
                const uint8_t *const cap_uv = &f->lf.tx_lpf_right_edge[1][(sby << (sbl2 - ss_ver)) + (size_t)(halign >> ss_ver) * (tc - 1)];
                for (int y = cy0; y < cy1; y++)
                    bitfn(cap_edge_class)(m->filter_uv[0][col4 >> ss_hor], 2, y, 16 >> ss_ver, cap_uv[y - cy0]);
            }
        }
        const int tile_row = f->lf.start_of_tile_row[sby];
        if (!tile_row) continue;
        const BlockContext *a = &f->a[f->sb128w * (tile_row - 1)];
        for (int x = 0; x < f->sb128w; x++, a++) {
            const int w = imin(32, f->w4 - (x << 5));
            for (int i = 0; i < w; i++)
                bitfn(cap_edge_class)(row[x].filter_y[1][y0], 3, i, 16, a->tx_lpf_y[i]);
            if (has_chroma) {
                const int cw = (w + ss_hor) >> ss_hor;
                for (int i = 0; i < cw; i++)
                    bitfn(cap_edge_class)(row[x].filter_uv[1][cy0], 2, i, 16 >> ss_hor, a->tx_lpf_uv[i]);
            }
        }
    }
}

static int bitfn(run_frame)(HookFrame *const hf, const Dav1dFrameContext *const f)
{
    const B200Backend *const be = b200hook_backend();
    if (!be) return -1;
    if (hf->unsupported) {
        fprintf(stderr, "b200hook: frame uses tools the emitters do not translate yet (%s%s%s%s)\n",
                hf->unsupported & 1 ? " palette" : "", hf->unsupported & 2 ? " inter" : "",
                hf->unsupported & 4 ? " single-pass-decoding" : "", hf->unsupported & 8 ? " out-of-memory" : "");
        return -1;
    }
    PicGeom g;
    bitfn(pic_geom)(f, &g);
    const Dav1dFrameHeader *const hdr = f->frame_hdr;
    const int ss_ver = f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420, ss_hor = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444;
    const int n_sb128 = f->sb128w * f->sb128h;
    const size_t mask_bytes = (size_t)n_sb128 * sizeof(Av1Filter), level_bytes = (size_t)n_sb128 * 32 * 32 * 4;
    const size_t lr_bytes = (size_t)f->sr_sb128w * f->sb128h * sizeof(Av1Restoration);
    if (!hf->stream && !(hf->stream = be->stream_create())) { fprintf(stderr, "b200hook: %s\n", be->last_error()); return -1; }
    for (int k = 0; k < 3; k++)
        if (b200hook_buf_reserve(&hf->pic[k], g.bytes, 0, 0)) return -1;
    if (b200hook_buf_reserve(&hf->mask, mask_bytes, 1, 0) || b200hook_buf_reserve(&hf->level, level_bytes, 1, 0) ||
        b200hook_buf_reserve(&hf->lr_mask, lr_bytes, 1, 0) ||
        b200hook_buf_reserve(&hf->tx, (size_t)imax(hf->n_tx, 1) * sizeof(B200IntraTx), 1, 1) ||
        b200hook_buf_reserve(&hf->coef, (hf->n_coef + 1) * sizeof(coef), 1, 1))
        return -1;
    memcpy(hf->mask.host, f->lf.mask, mask_bytes);
    bitfn(fix_tile_edges)(f, (Av1Filter *)hf->mask.host);
    memcpy(hf->level.host, f->lf.level, level_bytes);
    memcpy(hf->lr_mask.host, f->lf.lr_mask, lr_bytes);

    B200FrameJob j;
    memset(&j, 0, sizeof(j));
#if BITDEPTH == 8
    j.bitdepth_max = 255;
#else
    j.bitdepth_max = f->bitdepth_max;
#endif
    void *const p0 = hf->pic[0].dev, *const p1 = hf->pic[1].dev, *const p2 = hf->pic[2].dev;
    j.mc.dst = p0;
    j.d_coef = hf->coef.dev;
    for (int p = 0; p < 3; p++) { j.itx_stride[p] = g.stride[p]; j.mc.dst_stride[p] = g.stride[p]; }
    /* intra reconstruction */
    j.d_intra = (const B200IntraTx *)hf->tx.dev; j.n_intra = hf->n_tx;
    j.intra.pic = p0; j.intra.d_coef = hf->coef.dev;
    j.intra.ss_hor = ss_hor; j.intra.ss_ver = ss_ver;
    for (int p = 0; p < 3; p++) {
        j.intra.stride[p] = g.stride[p]; j.intra.plane_off[p] = g.off[p];
        j.intra.w4[p] = p ? (f->bw + ss_hor) >> ss_hor : f->bw;
        j.intra.h4[p] = p ? (f->bh + ss_ver) >> ss_ver : f->bh;
    }
    if (b200hook_buf_reserve(&hf->scratch, be->intra_scratch_bytes(&j.intra), 0, 0)) return -1;
    j.intra.scratch = hf->scratch.dev;
    /* decode order -> wavefront order (B200HOOK_WAVE_SORT=0 keeps decode order, which is also valid) */
    const HookBuf *txb = &hf->tx;
    static int wave_sort = -1;
    if (wave_sort < 0) { const char *e = getenv("B200HOOK_WAVE_SORT"); wave_sort = !e || atoi(e) != 0; }
    if (wave_sort && hf->n_tx > 0) {
        if (b200hook_buf_reserve(&hf->tx_sorted, (size_t)hf->n_tx * sizeof(B200IntraTx), 1, 0)) return -1;
        if (b200hook_wave_sort((const B200IntraTx *)hf->tx.host, (B200IntraTx *)hf->tx_sorted.host, hf->n_tx,
                               j.intra.w4, j.intra.h4, ss_hor, ss_ver) < 0) return -1;
        txb = &hf->tx_sorted;
        j.d_intra = (const B200IntraTx *)txb->dev;
    }
    /* deblock (reference src/recon_tmpl.c:1987-2027), in place on p0 */
    const int do_lf = (f->c->inloop_filters & DAV1D_INLOOPFILTER_DEBLOCK) && (hdr->loopfilter.level_y[0] || hdr->loopfilter.level_y[1]);
    j.run_lf = do_lf;
    j.lf.pic = p0;
    for (int p = 0; p < 3; p++) { j.lf.plane_off[p] = g.off[p]; j.lf.stride[p] = g.stride[p]; }
    j.lf.w4 = f->w4; j.lf.h4 = f->h4; j.lf.sb128w = f->sb128w; j.lf.b4_stride = (int)f->b4_stride;
    j.lf.ss_hor = ss_hor; j.lf.ss_ver = ss_ver; j.lf.sb128 = f->seq_hdr->sb128;
    j.lf.filter_y = do_lf; j.lf.filter_uv = hdr->loopfilter.level_u || hdr->loopfilter.level_v;
    j.lf.mask = (const B200Av1Filter *)hf->mask.dev;
    j.lf.level = (const uint8_t (*)[4])hf->level.dev;
    memcpy(j.lf.lut.e, f->lf.lim_lut.e, 64); memcpy(j.lf.lut.i, f->lf.lim_lut.i, 64);
    j.lf.lut.sharp[0] = f->lf.lim_lut.sharp[0]; j.lf.lut.sharp[1] = f->lf.lim_lut.sharp[1];
    /* CDEF (:2029-2058), p0 -> p1 */
    const int do_cdef = f->seq_hdr->cdef && (f->c->inloop_filters & DAV1D_INLOOPFILTER_CDEF);
    j.run_cdef = do_cdef;
    j.cdef.src = p0; j.cdef.dst = p1;
    for (int p = 0; p < 3; p++) { j.cdef.plane_off[p] = g.off[p]; j.cdef.stride[p] = g.stride[p]; }
    j.cdef.bw = f->bw; j.cdef.bh = f->bh; j.cdef.sb128w = f->sb128w; j.cdef.ss_hor = ss_hor; j.cdef.ss_ver = ss_ver;
    j.cdef.damping = hdr->cdef.damping;
    for (int i = 0; i < 8; i++) { j.cdef.y_strength[i] = hdr->cdef.y_strength[i]; j.cdef.uv_strength[i] = hdr->cdef.uv_strength[i]; }
    j.cdef.mask = (const B200Av1Filter *)hf->mask.dev;
    /* loop restoration (:2100-2109), -> p2 */
    const int do_lr = f->lf.restore_planes && (f->c->inloop_filters & DAV1D_INLOOPFILTER_RESTORATION);
    j.run_lr = do_lr;
    j.lr.cdef = do_cdef ? p1 : p0; j.lr.dbl = p0; j.lr.dst = p2;
    for (int p = 0; p < 3; p++) { j.lr.plane_off[p] = g.off[p]; j.lr.stride[p] = g.stride[p]; }
    j.lr.w = f->sr_cur.p.p.w; j.lr.h = f->sr_cur.p.p.h; j.lr.ss_hor = ss_hor; j.lr.ss_ver = ss_ver;
    j.lr.sb128 = f->seq_hdr->sb128; j.lr.sr_sb128w = f->sr_sb128w;
    j.lr.unit_size_log2[0] = hdr->restoration.unit_size[0]; j.lr.unit_size_log2[1] = hdr->restoration.unit_size[1];
    j.lr.restore_planes = f->lf.restore_planes;
    j.lr.lr_mask = (const B200Av1Restoration *)hf->lr_mask.dev;

    const B200Xfer up[5] = {
        { txb->host, txb->dev, (uint64_t)hf->n_tx * sizeof(B200IntraTx) },
        { hf->coef.host, hf->coef.dev, (uint64_t)hf->n_coef * sizeof(coef) },
        { hf->mask.host, hf->mask.dev, mask_bytes },
        { hf->level.host, hf->level.dev, level_bytes },
        { hf->lr_mask.host, hf->lr_mask.dev, lr_bytes },
    };
    uint8_t *const out = do_lr ? p2 : do_cdef ? p1 : p0;
    B200Xfer down[3];
    uint64_t d2h = 0, h2d = 0;
    for (int p = 0; p < 3; p++) {
        const int rows = p ? (f->cur.p.h + ss_ver) >> ss_ver : f->cur.p.h;
        down[p].host = f->cur.data[p];
        down[p].dev = out + (size_t)g.off[p] * sizeof(pixel);
        down[p].bytes = (uint64_t)rows * g.stride[p] * sizeof(pixel);
        d2h += down[p].bytes;
    }
    for (int i = 0; i < 5; i++) h2d += up[i].bytes;
    const double t0 = bitfn(now_ms)();
    b200hook_job_enter();
    const int r = be->frame_run_host(&j, up, 5, down, 3, hf->stream);
    b200hook_job_leave();
    if (r) { fprintf(stderr, "b200hook: b200_frame_run_host failed (%d): %s\n", r, be->last_error()); return -1; }
    b200hook_account((uint64_t)hf->n_tx, hf->n_coef, h2d, d2h, bitfn(now_ms)() - t0);
    return 0;
}

/* "tile superblock row reconstructed" (pass 2 calls this after every tile superblock row, reference
 * src/decode.c:2620-2635): the frame is complete when every tile has delivered all of its rows */
void bitfn(b200hook_backup_ipred_edge)(Dav1dTaskContext *const t)
{
    Dav1dFrameContext *const f = (Dav1dFrameContext *)t->f;
    HookFrame *const hf = b200hook_frame(f);
    if (!hf) return;
    pthread_mutex_lock(&hf->lock);
    if (t->frame_thread.pass != 2) hf->unsupported |= 4;
    const int total = f->sbh * f->frame_hdr->tiling.cols;
    if (++hf->tile_sbrows_done >= total) {
        if (bitfn(run_frame)(hf, f))
            atomic_fetch_or(&f->task_thread.error, 1);      /* the frame is reported as a decoding error */
        hf->tile_sbrows_done = 0; hf->n_tx = 0; hf->n_coef = 0; hf->unsupported = 0;
    }
    pthread_mutex_unlock(&hf->lock);
}

/* the post-filter sweep already ran on the device as part of the frame job */
void bitfn(b200hook_filter_sbrow)(Dav1dFrameContext *const f, const int sby) { (void)f; (void)sby; }
void bitfn(b200hook_filter_sbrow_deblock_cols)(Dav1dFrameContext *const f, const int sby) { (void)f; (void)sby; }
void bitfn(b200hook_filter_sbrow_deblock_rows)(Dav1dFrameContext *const f, const int sby) { (void)f; (void)sby; }
void bitfn(b200hook_filter_sbrow_cdef)(Dav1dTaskContext *const tc, const int sby) { (void)tc; (void)sby; }
void bitfn(b200hook_filter_sbrow_resize)(Dav1dFrameContext *const f, const int sby) { (void)f; (void)sby; }
void bitfn(b200hook_filter_sbrow_lr)(Dav1dFrameContext *const f, const int sby) { (void)f; (void)sby; }
