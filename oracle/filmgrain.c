/*
 * oracle/filmgrain.c — TEST INFRASTRUCTURE (see oracle_common.h).
 *
 * Plain-C restatement of dav1d's film grain synthesis:
 *   generate_grain_y / generate_grain_uv      reference src/filmgrain_tmpl.c:50-145
 *   fgy_32x32xn / fguv_32x32xn                reference src/filmgrain_tmpl.c:169-402
 *   generate_scaling, prep_grain, apply_grain reference src/fg_apply_tmpl.c:41-240
 * The apply functions are written per pixel: the grain of a pixel is the LUT sample of its own
 * 32x32 block, blended with the left / top / top-left blocks' samples inside the 2-sample overlap.
 * Block offsets are the k-th draw of the row's 16-bit LFSR (reference :190-214).
 * Grain LUT entries are int8 (8 bpc) or int16 (10/12 bpc) with a row pitch of 82.
 */
#include "oracle_common.h"
#include "tables_gen.h"

typedef struct {             /* layout of Dav1dFilmGrainData, reference include/dav1d/headers.h:315-333 */
    unsigned seed; int num_y_points; uint8_t y_points[14][2]; int chroma_scaling_from_luma; int num_uv_points[2];
    uint8_t uv_points[2][10][2]; int scaling_shift; int ar_coeff_lag; int8_t ar_coeffs_y[24]; int8_t ar_coeffs_uv[2][28];
    uint64_t ar_coeff_shift; int grain_scale_shift; int uv_mult[2]; int uv_luma_mult[2]; int uv_offset[2];
    int overlap_flag; int clip_to_restricted_range;
} FgData;

#define GW 82
#define GH 73
static inline int rnd_next(int bits, unsigned *state) {
    const int r = (int)*state;
    const unsigned bit = ((r >> 0) ^ (r >> 1) ^ (r >> 3) ^ (r >> 12)) & 1;
    *state = (r >> 1) | (bit << 15);
    return (*state >> (16 - bits)) & ((1 << bits) - 1);
}
static inline int round2(int x, int sh) { return (x + ((1 << sh) >> 1)) >> sh; }
static inline int LUT(const void *l, int hbd, int y, int x) { return hbd ? ((const int16_t *)l)[y * GW + x] : ((const int8_t *)l)[y * GW + x]; }
static inline void SLUT(void *l, int hbd, int y, int x, int v) { if (hbd) ((int16_t *)l)[y * GW + x] = (int16_t)v; else ((int8_t *)l)[y * GW + x] = (int8_t)v; }
static inline int PX(const void *p, int hbd, ptrdiff_t i) { return hbd ? ((const uint16_t *)p)[i] : ((const uint8_t *)p)[i]; }
static inline void SPX(void *p, int hbd, ptrdiff_t i, int v) { if (hbd) ((uint16_t *)p)[i] = (uint16_t)v; else ((uint8_t *)p)[i] = (uint8_t)v; }

/* uv < 0: luma LUT; else chroma plane uv with subsampling subx/suby and the finished luma LUT buf_y */
ORACLE_API void oracle_fg_generate_grain(void *buf, const void *buf_y, const FgData *d, int uv, int subx, int suby, int bdmax)
{
    const int hbd = bdmax > 255, b8 = (o_ulog2((unsigned)bdmax) + 1) - 8;
    unsigned seed = d->seed ^ (uv < 0 ? 0 : uv ? 0x49d8 : 0xb524);
    const int shift = 4 - b8 + d->grain_scale_shift;
    const int gmin = -(128 << b8), gmax = (128 << b8) - 1;
    const int cw = uv >= 0 && subx ? 44 : GW, ch = uv >= 0 && suby ? 38 : GH;
    for (int y = 0; y < ch; y++)
        for (int x = 0; x < cw; x++)
            SLUT(buf, hbd, y, x, round2(b200_gaussian_sequence[rnd_next(11, &seed)], shift));
    const int lag = d->ar_coeff_lag;
    for (int y = 3; y < ch; y++)
        for (int x = 3; x < cw - 3; x++) {
            const int8_t *coeff = uv < 0 ? d->ar_coeffs_y : d->ar_coeffs_uv[uv];
            int sum = 0;
            for (int dy = -lag; dy <= 0; dy++)
                for (int dx = -lag; dx <= lag; dx++) {
                    if (!dx && !dy) {
                        if (uv >= 0 && d->num_y_points) {       /* luma grain contribution (:112-127) */
                            int luma = 0;
                            const int lx = ((x - 3) << subx) + 3, ly = ((y - 3) << suby) + 3;
                            for (int i = 0; i <= suby; i++)
                                for (int j = 0; j <= subx; j++) luma += LUT(buf_y, hbd, ly + i, lx + j);
                            sum += round2(luma, subx + suby) * *coeff;
                        }
                        break;
                    }
                    sum += *(coeff++) * LUT(buf, hbd, y + dy, x + dx);
                }
            SLUT(buf, hbd, y, x, o_clip(LUT(buf, hbd, y, x) + round2(sum, (int)d->ar_coeff_shift), gmin, gmax));
        }
}

ORACLE_API void oracle_fg_scaling(int bitdepth, const uint8_t points[][2], int num, uint8_t *scaling)
{
    const int shift_x = bitdepth - 8, size = 1 << bitdepth;
    if (!num) { memset(scaling, 0, size); return; }
    memset(scaling, points[0][1], points[0][0] << shift_x);
    for (int i = 0; i < num - 1; i++) {
        const int bx = points[i][0], by = points[i][1], dx = points[i + 1][0] - bx, dy = points[i + 1][1] - by;
        const int delta = dy * ((0x10000 + (dx >> 1)) / dx);
        for (int x = 0; x < dx; x++) scaling[(bx + x) << shift_x] = (uint8_t)(by + ((0x8000 + x * delta) >> 16));
    }
    const int n = points[num - 1][0] << shift_x;
    memset(&scaling[n], points[num - 1][1], size - n);
    if (shift_x) {
        const int pad = 1 << shift_x, rnd = pad >> 1;
        for (int i = 0; i < num - 1; i++) {
            const int bx = points[i][0] << shift_x, dx = (points[i + 1][0] << shift_x) - bx;
            for (int x = 0; x < dx; x += pad) {
                const int range = scaling[bx + x + pad] - scaling[bx + x];
                for (int k = 1; k < pad; k++) scaling[bx + x + k] = (uint8_t)(scaling[bx + x] + ((rnd + k * range) >> shift_x));
            }
        }
    }
}

static int row_offset(const FgData *d, int row, int k) {   /* k-th block offset of 32-row strip `row` */
    unsigned s = d->seed;
    s ^= (((row * 37 + 178) & 0xFF) << 8);
    s ^= ((row * 173 + 105) & 0xFF);
    int v = 0;
    for (int i = 0; i <= k; i++) v = rnd_next(8, &s);
    return v;
}
static int lut_sample(const void *lut, int hbd, int randval, int subx, int suby, int x, int y) {
    const int offx = 3 + (2 >> subx) * (3 + (randval >> 4)), offy = 3 + (2 >> suby) * (3 + (randval & 0xF));
    return LUT(lut, hbd, offy + y, offx + x);
}

/* grain value of strip pixel (x, y); pw = strip width, bh = strip height (plane units) */
static int pixel_grain(const FgData *d, const void *lut, int hbd, int b8, int row, int x, int y, int pw, int bh, int sx, int sy)
{
    static const int W[2][2][2] = { { { 27, 17 }, { 17, 27 } }, { { 23, 22 }, { 0, 0 } } };
    const int gmin = -(128 << b8), gmax = (128 << b8) - 1;
    const int bs = 32 >> sx, bsy = 32 >> sy, bi = x / bs, xin = x - bi * bs;
    const int bw = o_min(bs, pw - bi * bs);
    const int xov = d->overlap_flag && bi && xin < o_min(2 >> sx, bw);
    const int yov = d->overlap_flag && row > 0 && y < o_min(2 >> sy, bh);
    int g = lut_sample(lut, hbd, row_offset(d, row, bi), sx, sy, xin, y);
    if (xov) {
        const int old = lut_sample(lut, hbd, row_offset(d, row, bi - 1), sx, sy, xin + bs, y);
        g = o_clip(round2(old * W[sx][xin][0] + g * W[sx][xin][1], 5), gmin, gmax);
    }
    if (yov) {
        int top = lut_sample(lut, hbd, row_offset(d, row - 1, bi), sx, sy, xin, y + bsy);
        if (xov) {
            const int old = lut_sample(lut, hbd, row_offset(d, row - 1, bi - 1), sx, sy, xin + bs, y + bsy);
            top = o_clip(round2(old * W[sx][xin][0] + top * W[sx][xin][1], 5), gmin, gmax);
        }
        g = o_clip(round2(top * W[sy][y][0] + g * W[sy][y][1], 5), gmin, gmax);
    }
    return g;
}

ORACLE_API void oracle_fgy_32x32xn(void *dst_row, const void *src_row, ptrdiff_t stride_bytes, const FgData *d, size_t pw,
                                   const uint8_t *scaling, const void *lut, int bh, int row_num, int bdmax)
{
    const int hbd = bdmax > 255, b8 = (o_ulog2((unsigned)bdmax) + 1) - 8;
    const ptrdiff_t ps = hbd ? stride_bytes / 2 : stride_bytes;
    const int mn = d->clip_to_restricted_range ? 16 << b8 : 0, mx = d->clip_to_restricted_range ? 235 << b8 : bdmax;
    for (int y = 0; y < bh; y++)
        for (int x = 0; x < (int)pw; x++) {
            const int g = pixel_grain(d, lut, hbd, b8, row_num, x, y, (int)pw, bh, 0, 0);
            const int s = PX(src_row, hbd, y * ps + x);
            SPX(dst_row, hbd, y * ps + x, o_clip(s + round2(scaling[s] * g, d->scaling_shift), mn, mx));
        }
}

ORACLE_API void oracle_fguv_32x32xn(void *dst_row, const void *src_row, ptrdiff_t stride_bytes, const FgData *d, size_t pw,
                                    const uint8_t *scaling, const void *lut, int bh, int row_num, const void *luma_row,
                                    ptrdiff_t luma_stride_bytes, int uv, int is_id, int sx, int sy, int bdmax)
{
    const int hbd = bdmax > 255, b8 = (o_ulog2((unsigned)bdmax) + 1) - 8;
    const ptrdiff_t ps = hbd ? stride_bytes / 2 : stride_bytes, ls = hbd ? luma_stride_bytes / 2 : luma_stride_bytes;
    const int mn = d->clip_to_restricted_range ? 16 << b8 : 0;
    const int mx = d->clip_to_restricted_range ? (is_id ? 235 : 240) << b8 : bdmax;
    for (int y = 0; y < bh; y++)
        for (int x = 0; x < (int)pw; x++) {
            const int g = pixel_grain(d, lut, hbd, b8, row_num, x, y, (int)pw, bh, sx, sy);
            const ptrdiff_t lo = (ptrdiff_t)(y << sy) * ls + (x << sx);
            int avg = PX(luma_row, hbd, lo);
            if (sx) avg = (avg + PX(luma_row, hbd, lo + 1) + 1) >> 1;
            const int s = PX(src_row, hbd, y * ps + x);
            int val = avg;
            if (!d->chroma_scaling_from_luma) {
                const int combined = avg * d->uv_luma_mult[uv] + s * d->uv_mult[uv];
                val = o_clip((combined >> 6) + d->uv_offset[uv] * (1 << b8), 0, bdmax);
            }
            SPX(dst_row, hbd, y * ps + x, o_clip(s + round2(scaling[val] * g, d->scaling_shift), mn, mx));
        }
}

/* ---- whole picture (dav1d_apply_grain, reference src/fg_apply_tmpl.c:225-240) ---- */
typedef struct {           /* restates B200FgFrame (include/b200av1.h) */
    const void *in; void *out;
    uint32_t plane_off[3]; int32_t stride[3];
    int32_t w, h, ss_hor, ss_ver, is_id;
    FgData data;
} OracleFgFrame;

ORACLE_API void oracle_fg_apply_frame(int bdmax, const OracleFgFrame *f)
{
    const int hbd = bdmax > 255, bitdepth = o_ulog2((unsigned)bdmax) + 1; const size_t px = hbd ? 2 : 1;
    const FgData *d = &f->data;
    static __thread int16_t lut[3][(GH + 1) * GW];
    static __thread uint8_t scaling[3][4096];
    const int sx = f->ss_hor, sy = f->ss_ver;
    oracle_fg_generate_grain(lut[0], NULL, d, -1, 0, 0, bdmax);
    for (int uv = 0; uv < 2; uv++)
        if (d->num_uv_points[uv] || d->chroma_scaling_from_luma) oracle_fg_generate_grain(lut[1 + uv], lut[0], d, uv, sx, sy, bdmax);
    if (d->num_y_points || d->chroma_scaling_from_luma) oracle_fg_scaling(bitdepth, d->y_points, d->num_y_points, scaling[0]);
    for (int uv = 0; uv < 2; uv++)
        if (d->num_uv_points[uv]) oracle_fg_scaling(bitdepth, d->uv_points[uv], d->num_uv_points[uv], scaling[1 + uv]);
    const int cw = (f->w + sx) >> sx, chh = (f->h + sy) >> sy;
    /* planes without grain are copied through */
    for (int pl = 0; pl < 3; pl++) {
        const int pw = pl ? cw : f->w, ph = pl ? chh : f->h;
        const int grained = pl ? (d->chroma_scaling_from_luma || d->num_uv_points[pl - 1]) : d->num_y_points;
        if (grained) continue;
        for (int y = 0; y < ph; y++)
            memcpy((uint8_t *)f->out + ((size_t)f->plane_off[pl] + (size_t)y * f->stride[pl]) * px,
                   (const uint8_t *)f->in + ((size_t)f->plane_off[pl] + (size_t)y * f->stride[pl]) * px, (size_t)pw * px);
    }
    const int rows = (f->h + 31) / 32;
    for (int row = 0; row < rows; row++) {
        const int bh = o_min(f->h - row * 32, 32);
        const uint8_t *luma_src = (const uint8_t *)f->in + ((size_t)f->plane_off[0] + (size_t)row * 32 * f->stride[0]) * px;
        if (d->num_y_points)
            oracle_fgy_32x32xn((uint8_t *)f->out + ((size_t)f->plane_off[0] + (size_t)row * 32 * f->stride[0]) * px, luma_src,
                               f->stride[0] * (ptrdiff_t)px, d, f->w, scaling[0], lut[0], bh, row, bdmax);
        const int cbh = (bh + sy) >> sy;
        for (int uv = 0; uv < 2; uv++) {
            if (!(d->chroma_scaling_from_luma || d->num_uv_points[uv])) continue;
            const size_t off = ((size_t)f->plane_off[1 + uv] + (size_t)(row * 32 >> sy) * f->stride[1 + uv]) * px;
            /* odd widths: the reference extends the luma row by one sample; clamp instead (same value) */
            if (f->w & sx) {
                /* emulate ptr[w] = ptr[w-1] on a private copy of the luma strip */
                static __thread uint8_t strip[32 * 8200 * 2];
                const size_t rowb = (size_t)f->stride[0] * px;
                for (int y = 0; y < bh; y++) {
                    memcpy(strip + y * rowb, luma_src + y * rowb, (size_t)f->w * px);
                    memcpy(strip + y * rowb + (size_t)f->w * px, strip + y * rowb + (size_t)(f->w - 1) * px, px);
                }
                oracle_fguv_32x32xn((uint8_t *)f->out + off, (const uint8_t *)f->in + off, f->stride[1 + uv] * (ptrdiff_t)px, d, cw,
                                    scaling[d->chroma_scaling_from_luma ? 0 : 1 + uv], lut[1 + uv], cbh, row, strip,
                                    f->stride[0] * (ptrdiff_t)px, uv, f->is_id, sx, sy, bdmax);
            } else {
                oracle_fguv_32x32xn((uint8_t *)f->out + off, (const uint8_t *)f->in + off, f->stride[1 + uv] * (ptrdiff_t)px, d, cw,
                                    scaling[d->chroma_scaling_from_luma ? 0 : 1 + uv], lut[1 + uv], cbh, row, luma_src,
                                    f->stride[0] * (ptrdiff_t)px, uv, f->is_id, sx, sy, bdmax);
            }
        }
    }
}
