/*
 * oracle/itx.c — TEST INFRASTRUCTURE (see oracle_common.h).
 *
 * Plain-C restatement of the 2-D inverse transform + add:
 *   inv_txfm_add_c           reference src/itx_tmpl.c:43-119
 *   inv_txfm_add_wht_wht_4x4 reference src/itx_tmpl.c:184-203
 *   per-size shift table      reference src/itx_tmpl.c:160-178
 *   1-D type pairs            reference src/itx_1d.c:1043-1060
 * One entry point serves 8 bpc (uint8 pixels, int16 coefs) and 10/12 bpc (uint16 pixels,
 * int32 coefs); `bitdepth_max` = 255 / 1023 / 4095 selects it.
 */
#include "itx_1d.h"
#include "tables_gen.h"

/* RectTxfmSize order (reference src/levels.h:38-79) */
static const uint8_t tx_w[19] = { 4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64 };
static const uint8_t tx_h[19] = { 4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16 };
static const uint8_t tx_shift[19] = { 0, 1, 2, 2, 2, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2 };

enum { T_DCT, T_ADST, T_FLIPADST, T_IDENTITY };
/* TxfmType slot X_Y of the itxfm_add table means X vertical, Y horizontal
 * (reference src/levels.h:81-83); the wrapper installed in that slot passes the MIRRORED
 * enum into inv_txfm_add_c (src/itx_tmpl.c:142-143,232-262), so per slot:
 *   { first = row pass (horizontal, length w), second = column pass (vertical, length h) } */
static const uint8_t tx_1d[16][2] = {
    /* DCT_DCT           */ { T_DCT, T_DCT },
    /* ADST_DCT          */ { T_DCT, T_ADST },
    /* DCT_ADST          */ { T_ADST, T_DCT },
    /* ADST_ADST         */ { T_ADST, T_ADST },
    /* FLIPADST_DCT      */ { T_DCT, T_FLIPADST },
    /* DCT_FLIPADST      */ { T_FLIPADST, T_DCT },
    /* FLIPADST_FLIPADST */ { T_FLIPADST, T_FLIPADST },
    /* ADST_FLIPADST     */ { T_FLIPADST, T_ADST },
    /* FLIPADST_ADST     */ { T_ADST, T_FLIPADST },
    /* IDTX              */ { T_IDENTITY, T_IDENTITY },
    /* V_DCT             */ { T_IDENTITY, T_DCT },
    /* H_DCT             */ { T_DCT, T_IDENTITY },
    /* V_ADST            */ { T_IDENTITY, T_ADST },
    /* H_ADST            */ { T_ADST, T_IDENTITY },
    /* V_FLIPADST        */ { T_IDENTITY, T_FLIPADST },
    /* H_FLIPADST        */ { T_FLIPADST, T_IDENTITY },
};

static oracle_tx1d_fn pick_1d(int len, int type) {
    switch (type) {
    case T_DCT:
        return len == 4 ? oracle_dct4 : len == 8 ? oracle_dct8 : len == 16 ? oracle_dct16 :
               len == 32 ? oracle_dct32 : oracle_dct64;
    case T_ADST:
        return len == 4 ? oracle_adst4 : len == 8 ? oracle_adst8 : len == 16 ? oracle_adst16 : NULL;
    case T_FLIPADST:
        return len == 4 ? oracle_flipadst4 : len == 8 ? oracle_flipadst8 :
               len == 16 ? oracle_flipadst16 : NULL;
    default:
        return len == 4 ? oracle_identity4 : len == 8 ? oracle_identity8 :
               len == 16 ? oracle_identity16 : len == 32 ? oracle_identity32 : NULL;
    }
}

static inline int ld_coef(const void *cf, int hbd, int i) {
    return hbd ? ((const int32_t *)cf)[i] : ((const int16_t *)cf)[i];
}
static inline int ld_px(const void *p, int hbd, ptrdiff_t i) {
    return hbd ? ((const uint16_t *)p)[i] : ((const uint8_t *)p)[i];
}
static inline void st_px(void *p, int hbd, ptrdiff_t i, int v) {
    if (hbd) ((uint16_t *)p)[i] = (uint16_t)v; else ((uint8_t *)p)[i] = (uint8_t)v;
}

/* returns 0 on success, -1 if (tx, txtp) is not a defined transform */
ORACLE_API int oracle_inv_txfm_add(void *dst, ptrdiff_t stride_bytes, void *coeff, int eob,
                                   int tx, int txtp, int bitdepth_max)
{
    const int hbd = bitdepth_max > 255;
    const ptrdiff_t ps = hbd ? stride_bytes / 2 : stride_bytes;   /* PXSTRIDE */
    if (tx < 0 || tx >= 19) return -1;

    if (txtp == 16) { /* WHT_WHT, lossless 4x4 only */
        if (tx != 0) return -1;
        int32_t tmp[16];
        for (int y = 0; y < 4; y++) {
            for (int x = 0; x < 4; x++) tmp[y * 4 + x] = ld_coef(coeff, hbd, y + x * 4) >> 2;
            oracle_wht4(&tmp[y * 4], 1);
        }
        memset(coeff, 0, (hbd ? 4 : 2) * 16);
        for (int x = 0; x < 4; x++) oracle_wht4(&tmp[x], 4);
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++)
                st_px(dst, hbd, y * ps + x, o_clip(ld_px(dst, hbd, y * ps + x) + tmp[y * 4 + x], 0, bitdepth_max));
        return 0;
    }
    if (txtp < 0 || txtp > 16) return -1;

    const int w = tx_w[tx], h = tx_h[tx], shift = tx_shift[tx];
    const int is_rect2 = w * 2 == h || h * 2 == w;
    const int rnd = (1 << shift) >> 1;

    if (eob < (txtp == 0)) { /* DCT_DCT dc-only shortcut, src/itx_tmpl.c:58-70 */
        int dc = ld_coef(coeff, hbd, 0);
        if (hbd) ((int32_t *)coeff)[0] = 0; else ((int16_t *)coeff)[0] = 0;
        if (is_rect2) dc = (dc * 181 + 128) >> 8;
        dc = (dc * 181 + 128) >> 8;
        dc = (dc + rnd) >> shift;
        dc = (dc * 181 + 128 + 2048) >> 12;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                st_px(dst, hbd, y * ps + x, o_clip(ld_px(dst, hbd, y * ps + x) + dc, 0, bitdepth_max));
        return 0;
    }

    const int t0 = tx_1d[txtp][0], t1 = tx_1d[txtp][1];
    const oracle_tx1d_fn f0 = pick_1d(w, t0), f1 = pick_1d(h, t1);
    if (!f0 || !f1) return -1;
    const int sh = o_min(h, 32), sw = o_min(w, 32);
    const int row_lo = hbd ? (int)((unsigned)~bitdepth_max << 7) : INT16_MIN;
    const int col_lo = hbd ? (int)((unsigned)~bitdepth_max << 5) : INT16_MIN;
    const int row_hi = ~row_lo, col_hi = ~col_lo;

    /* number of coefficient "rows" that can be non-zero, src/itx_tmpl.c:88-95 */
    int last;
    const int lw = w == 4 ? 0 : w == 8 ? 1 : w == 16 ? 2 : w == 32 ? 3 : 4;
    if (t1 == T_IDENTITY && t0 != T_IDENTITY)      last = o_min(sh - 1, eob);
    else if (t0 == T_IDENTITY && t1 != T_IDENTITY) last = eob >> (lw + 2);
    else                                           last = b200_lnz_col[b200_lnz_col_off[tx] + eob];

    static __thread int32_t tmp[64 * 64];
    int32_t *c = tmp;
    for (int y = 0; y <= last; y++, c += w) {
        for (int x = 0; x < sw; x++) {
            int v = ld_coef(coeff, hbd, y + x * sh);
            c[x] = is_rect2 ? (int)((unsigned)v * 181u + 128u) >> 8 : v;
        }
        for (int x = sw; x < w; x++) c[x] = 0;   /* tx64: upper half never read by dct64 */
        f0(c, 1, row_lo, row_hi);
    }
    if (last + 1 < sh) memset(c, 0, sizeof(*c) * (size_t)(sh - last - 1) * w);
    memset(coeff, 0, (size_t)(hbd ? 4 : 2) * sw * sh);
    for (int i = 0; i < w * sh; i++)
        tmp[i] = o_clip((tmp[i] + rnd) >> shift, col_lo, col_hi);
    for (int x = 0; x < w; x++) f1(&tmp[x], w, col_lo, col_hi);
    c = tmp;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++, c++)
            st_px(dst, hbd, y * ps + x, o_clip(ld_px(dst, hbd, y * ps + x) + ((*c + 8) >> 4), 0, bitdepth_max));
    return 0;
}

/* is (tx, txtp) a defined slot of the itxfm_add table? (reference src/itx_tmpl.c:220-288) */
ORACLE_API int oracle_itx_defined(int tx, int txtp) {
    if (tx < 0 || tx >= 19 || txtp < 0 || txtp > 16) return 0;
    if (txtp == 16) return tx == 0;
    const int w = tx_w[tx], h = tx_h[tx], mx = o_max(w, h);
    if (mx == 64) return txtp == 0;
    if (mx == 32) return txtp == 0 || txtp == 9;
    if (mx == 16 && o_min(w, h) == 16) return txtp <= 11;   /* 16x16: no 1-D ADST+identity */
    return 1;
}

/* Batched form over the same block records the CUDA kernels consume (layout of
 * B200ItxBlock in include/b200av1.h, restated here so oracle/ stays self-contained). */
typedef struct OracleItxBlock {
    uint32_t dst_off, coef_off;
    int16_t eob;
    uint8_t txtp, plane;
} OracleItxBlock;

ORACLE_API int oracle_itx_add_batch(int bitdepth_max, int tx, const OracleItxBlock *blocks, int n,
                                    void *coef, void *pic, const int32_t stride_px[3], int zero_coefs)
{
    const int hbd = bitdepth_max > 255;
    const int sw = o_min(tx_w[tx], 32), sh = o_min(tx_h[tx], 32);
    const size_t cs = hbd ? 4 : 2, ps = hbd ? 2 : 1;
    for (int i = 0; i < n; i++) {
        const OracleItxBlock *b = &blocks[i];
        uint8_t *cf = (uint8_t *)coef + (size_t)b->coef_off * cs;
        uint8_t save[32 * 32 * 4];
        if (!zero_coefs) memcpy(save, cf, cs * sw * sh);
        int r = oracle_inv_txfm_add((uint8_t *)pic + (size_t)b->dst_off * ps,
                                    (ptrdiff_t)stride_px[b->plane] * (ptrdiff_t)ps, cf, b->eob, tx,
                                    b->txtp, bitdepth_max);
        if (!zero_coefs) memcpy(cf, save, cs * sw * sh);
        if (r) return r;
    }
    return 0;
}
