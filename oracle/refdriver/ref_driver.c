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
