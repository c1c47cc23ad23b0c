/*
 * oracle/refdriver/ref_driver.c — TEST INFRASTRUCTURE.
 *
 * Our own thin C layer linked INTO oracle/_ref/libdav1d_ref.so next to the unmodified
 * reference objects, so that it can reach the reference's hidden-visibility tables and
 * drive its DSP function pointers in bulk (parity sweeps, CPU baseline timing).
 * It includes reference headers at build time only; no reference source is copied.
 */
#include "config.h"
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <pthread.h>
#include <stdlib.h>
#include <time.h>

#include "src/tables.h"
#include "src/scan.h"
#include "src/levels.h"

#define API __attribute__((visibility("default")))

#define T(sym) if (!strcmp(name, #sym)) { *bytes = sizeof(sym); return sym; }
API const void *refdrv_table(const char *name, size_t *bytes) {
    T(dav1d_cdef_directions) T(dav1d_sgr_params) T(dav1d_sgr_x_by_x)
    T(dav1d_mc_subpel_filters) T(dav1d_mc_warp_filter) T(dav1d_resize_filter)
    T(dav1d_sm_weights) T(dav1d_dr_intra_derivative) T(dav1d_filter_intra_taps)
    T(dav1d_obmc_masks) T(dav1d_gaussian_sequence)
    *bytes = 0;
    return NULL;
}

API const uint8_t *refdrv_last_nonzero_col(int tx) {
    dav1d_init_last_nonzero_col_from_eob_tables();
    return dav1d_last_nonzero_col_from_eob[tx];
}

/* scan order + transform class, for the checkasm-style coefficient generator in tests/ */
API const uint16_t *refdrv_scan(int tx) { return dav1d_scans[tx]; }
API int refdrv_tx_type_class(int txtp) { return dav1d_tx_type_class[txtp]; }

/* ---- batched itx over B200ItxBlock-shaped records, optionally multi-threaded ------------- */
#include "src/itx.h"
typedef struct RefItxBlock { uint32_t dst_off, coef_off; int16_t eob; uint8_t txtp, plane; } RefItxBlock;
typedef void (*itx8_fn)(uint8_t *, ptrdiff_t, int16_t *, int);
typedef void (*itx16_fn)(uint16_t *, ptrdiff_t, int32_t *, int, int);

typedef struct {
    int bdmax, tx, n, zero; const RefItxBlock *blocks; void *coef, *pic; const int32_t *st;
    void *tbl; int lo, hi;
} ItxJob;

static void *itx_worker(void *arg) {
    ItxJob *j = arg;
    const int hbd = j->bdmax > 255;
    const TxfmInfo *t = &dav1d_txfm_dimensions[j->tx];
    const int sw = t->w * 4 > 32 ? 32 : t->w * 4, sh = t->h * 4 > 32 ? 32 : t->h * 4;
    void *(*tbl)[N_TX_TYPES_PLUS_LL] = j->tbl;
    for (int i = j->lo; i < j->hi; i++) {
        const RefItxBlock *b = &j->blocks[i];
        if (hbd) {
            int32_t *cf = (int32_t *)j->coef + b->coef_off, save[1024];
            if (!j->zero) memcpy(save, cf, 4 * sw * sh);
            ((itx16_fn)tbl[j->tx][b->txtp])((uint16_t *)j->pic + b->dst_off, (ptrdiff_t)j->st[b->plane] * 2, cf, b->eob, j->bdmax);
            if (!j->zero) memcpy(cf, save, 4 * sw * sh);
        } else {
            int16_t *cf = (int16_t *)j->coef + b->coef_off, save[1024];
            if (!j->zero) memcpy(save, cf, 2 * sw * sh);
            ((itx8_fn)tbl[j->tx][b->txtp])((uint8_t *)j->pic + b->dst_off, (ptrdiff_t)j->st[b->plane], cf, b->eob);
            if (!j->zero) memcpy(cf, save, 2 * sw * sh);
        }
    }
    return NULL;
}

/* Runs the reference's own itxfm_add[tx][txtp] C functions over the records. Blocks must not
 * overlap in the picture when nthreads > 1. Returns elapsed seconds. */
API double refdrv_itx_add_batch(int bitdepth_max, int tx, const RefItxBlock *blocks, int n, void *coef,
                                void *pic, const int32_t stride_px[3], int zero_coefs, int nthreads)
{
    static void *tbl8[N_RECT_TX_SIZES][N_TX_TYPES_PLUS_LL], *tbl16[N_RECT_TX_SIZES][N_TX_TYPES_PLUS_LL];
    static int init8, init16;
    const int hbd = bitdepth_max > 255;
    if (hbd && !init16) { dav1d_itx_dsp_init_16bpc((void *)tbl16, 10); init16 = 1; }
    if (!hbd && !init8) { dav1d_itx_dsp_init_8bpc((void *)tbl8, 8); init8 = 1; }
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    ItxJob jobs[256]; pthread_t th[256];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < nthreads; k++) {
        jobs[k] = (ItxJob){ bitdepth_max, tx, n, zero_coefs, blocks, coef, pic, stride_px,
                            hbd ? (void *)tbl16 : (void *)tbl8,
                            (int)((long long)n * k / nthreads), (int)((long long)n * (k + 1) / nthreads) };
        if (nthreads > 1) pthread_create(&th[k], NULL, itx_worker, &jobs[k]);
    }
    if (nthreads == 1) itx_worker(&jobs[0]);
    else for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
