/*
 * oracle/itx_1d.h — TEST INFRASTRUCTURE (see oracle_common.h).
 *
 * Plain-C restatement of the 1-D inverse transforms of dav1d:
 *   inverse DCT 4/8/16/32/64   reference src/itx_1d.c:65-781
 *   inverse ADST 4/8/16 (+flip) reference src/itx_1d.c:783-979
 *   identity 4/8/16/32          reference src/itx_1d.c:983-1017
 *   WHT4                        reference src/itx_1d.c:1066-1081
 * Every rotation is written in the canonical form
 *      ((x*cx + y*cy + rnd) >> sh) + adj
 * with cx/cy the *reduced* multipliers the reference uses (c or c-4096, or c/2 with
 * sh=11) so that results agree with it even on out-of-spec inputs that wrap.
 * All functions work in place on c[0], c[s], c[2s], ... ; lo/hi are the clip bounds.
 */
#ifndef ORACLE_ITX_1D_H
#define ORACLE_ITX_1D_H
#include "oracle_common.h"

typedef void (*oracle_tx1d_fn)(int32_t *c, ptrdiff_t s, int lo, int hi);

void oracle_dct4 (int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_dct8 (int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_dct16(int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_dct32(int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_dct64(int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_adst4 (int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_adst8 (int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_adst16(int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_flipadst4 (int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_flipadst8 (int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_flipadst16(int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_identity4 (int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_identity8 (int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_identity16(int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_identity32(int32_t *c, ptrdiff_t s, int lo, int hi);
void oracle_wht4(int32_t *c, ptrdiff_t s);

#endif
