/*
 * b200av1.h — C ABI of the B200-native AV1 reconstruction / post-filter back end.
 *
 * Drop-in boundary for videolan/dav1d's block-reconstruction path:
 *
 *   Level 1  b200_*_dsp_init_{8,16}bpc() fill tables of function pointers that have exactly
 *            the signatures of dav1d's Dav1dDSPContext members (reference src/internal.h:62-70;
 *            itx: src/itx.h:37-40,70-72). Each call ships its operands to HBM, launches the
 *            CUDA kernel and waits — correct but one block per launch; it is the semantic
 *            definition of the batched kernels and what the parity tests drive.
 *   Level 2  b200_*_batch() take arrays of block records already resident in HBM (device
 *            pointers) and process a whole frame's worth of work per launch; this is what
 *            a dav1d `f->bd_fn` record emitter (reference src/internal.h:247-262) feeds.
 *
 * Plain C: pointers, sizes, no C++/torch types. All functions return 0 on success and a
 * negative value on error (b200_last_error() gives the message) unless they mirror a
 * `void` dav1d signature. There is NO CPU fallback: without a CUDA device every entry
 * point fails.
 */
#ifndef B200AV1_H
#define B200AV1_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_API __attribute__((visibility("default")))

/* ---- library / context ------------------------------------------------------------- */
B200_API int b200_version(void);
B200_API const char *b200_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
B200_API uint64_t b200_launch_count(void);

/* enum RectTxfmSize / enum TxfmType numbering is dav1d's (reference src/levels.h:38-110) */
#define B200_N_RECT_TX_SIZES 19
#define B200_N_TX_TYPES_PLUS_LL 17
#define B200_WHT_WHT 16

/* ---- itx: Level 1 (replaces dav1d_itx_dsp_init_{8,16}bpc, reference src/itx_tmpl.c:220-311) */
/* typedefs mirror decl_itx_fn (reference src/itx.h:37-40) */
typedef void (*b200_itxfm_fn_8bpc)(uint8_t *dst, ptrdiff_t dst_stride, int16_t *coeff, int eob);
typedef void (*b200_itxfm_fn_16bpc)(uint16_t *dst, ptrdiff_t dst_stride, int32_t *coeff, int eob,
                                    int bitdepth_max);
/* same layout as Dav1dInvTxfmDSPContext (reference src/itx.h:70-72) */
typedef struct B200InvTxfmDSPContext8 {
    b200_itxfm_fn_8bpc itxfm_add[B200_N_RECT_TX_SIZES][B200_N_TX_TYPES_PLUS_LL];
} B200InvTxfmDSPContext8;
typedef struct B200InvTxfmDSPContext16 {
    b200_itxfm_fn_16bpc itxfm_add[B200_N_RECT_TX_SIZES][B200_N_TX_TYPES_PLUS_LL];
} B200InvTxfmDSPContext16;
B200_API void b200_itx_dsp_init_8bpc(B200InvTxfmDSPContext8 *c, int bpc);
B200_API void b200_itx_dsp_init_16bpc(B200InvTxfmDSPContext16 *c, int bpc);
/* non-table form of the same call (host pointers); bitdepth_max 255 selects 8 bpc */
B200_API int b200_inv_txfm_add(void *dst, ptrdiff_t dst_stride, void *coeff, int eob, int tx,
                               int txtp, int bitdepth_max);

/* ---- itx: Level 2 (batched, device-resident) ---------------------------------------- */
/* One transform block of a frame. All blocks of one b200_itx_add_batch call share `tx`.
 * dst_off: offset of the block's top-left pixel, in PIXELS, from the picture base pointer;
 * coef_off: offset in COEFFICIENTS into the coefficient stream. The block's coefficients are
 * min(w,32)*min(h,32) entries laid out as dav1d's decode_coefs writes them (x-frequency major:
 * coeff[y + x*min(h,32)], reference src/itx_tmpl.c:96-102), dequantised. */
typedef struct B200ItxBlock {
    uint32_t dst_off;
    uint32_t coef_off;
    int16_t eob;      /* as passed to itxfm_add (>= 0) */
    uint8_t txtp;     /* enum TxfmType, 16 = WHT_WHT */
    uint8_t plane;    /* index into stride_px[] */
} B200ItxBlock;

/* d_blocks/d_coef/d_pic are DEVICE pointers; stride_px[3] per-plane picture strides in pixels
 * (may be negative); stream is a cudaStream_t (NULL = default stream). The call is
 * asynchronous with respect to the host. If zero_coefs != 0 the consumed coefficients are
 * zeroed like dav1d's callee contract (reference src/itx_tmpl.c:108). */
B200_API int b200_itx_add_batch(int bitdepth_max, int tx, const B200ItxBlock *d_blocks, int n_blocks,
                                void *d_coef, void *d_pic, const int32_t stride_px[3],
                                int zero_coefs, void *stream);

/* Same work through HOST buffers (the end-to-end leg of bench.py): copies blocks, coefficients
 * and the picture to HBM, runs b200_itx_add_batch, copies the picture back, synchronises. */
B200_API int b200_itx_add_batch_host(int bitdepth_max, int tx, const B200ItxBlock *blocks, int n_blocks,
                                     void *coef, size_t coef_bytes, void *pic, size_t pic_bytes,
                                     const int32_t stride_px[3], int zero_coefs);

#ifdef __cplusplus
}
#endif
#endif /* B200AV1_H */
