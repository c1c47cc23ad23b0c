/*
 * oracle/ — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of dav1d's reconstruction / post-filter DSP arithmetic, used as
 * the checker for the CUDA kernels in dav1d_b200/csrc. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library. The product
 * never links, imports or falls back to it.
 *
 * Pinning: every function here is compared bit-for-bit against the UNMODIFIED reference
 * C path (oracle/_ref/libdav1d_ref.so, built by oracle/Makefile from /root/reference) by
 * tests/test_oracle_vs_ref.py on checkasm-style inputs (the reference tree stores no
 * golden vectors for this path), and committed fixtures generated from the
 * reference live under tests/golden/.
 */
#ifndef ORACLE_COMMON_H
#define ORACLE_COMMON_H
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static inline int o_min(int a, int b) { return a < b ? a : b; }
static inline int o_max(int a, int b) { return a > b ? a : b; }
static inline int o_clip(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static inline int o_abs(int v) { return v < 0 ? -v : v; }
static inline int o_sign(int v) { return v < 0 ? -1 : 1; }
static inline int o_ulog2(unsigned v) { return 31 - __builtin_clz(v); }

/* wrap-around (two's complement) multiply-accumulate helpers: the reference relies on
 * x86 wrap-around for out-of-spec inputs (src/itx_1d.c:39-63); doing the sums in
 * unsigned keeps that behaviour without signed-overflow UB. */
static inline int o_mac2(int x, int cx, int y, int cy, unsigned rnd, int sh) {
    unsigned s = (unsigned)x * (unsigned)cx + (unsigned)y * (unsigned)cy + rnd;
    return (int)s >> sh;
}

#define ORACLE_API __attribute__((visibility("default")))

#endif
