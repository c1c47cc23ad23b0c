/*
 * tests/emu/cuda_emu.h — TEST-ONLY host execution harness for the .cu sources.
 *
 * There is no GPU in the build container, and every gpurun round trip costs minutes of
 * a small budget. This header lets g++ compile dav1d_b200/csrc/*.cu unchanged
 * (-DB200_EMU -include cuda_emu.h) and run each kernel on the CPU with one *fiber* per
 * CUDA thread, so indexing, barriers, shuffles and the bit-exact arithmetic can be
 * debugged against the oracle before any GPU time is spent.
 *
 * It is NOT a product path: the emulated library is built into tests/emu/_build/ by
 * tests/emu/build_emu.py, is only ever loaded by tests (B200AV1_EMU_LIB), and the
 * dav1d_b200 package refuses to run without the real CUDA library.
 *
 * Model: blocks of a launch run in groups of `emu_resident_blocks` (default 1); inside a
 * group all threads are fibers scheduled round-robin; __syncthreads / __syncwarp /
 * shuffles yield until their barrier generation advances. __shared__ variables are
 * plain statics (valid because a statically-shared kernel runs one block at a time);
 * kernels that need several co-resident blocks must use dynamic shared memory.
 */
#ifndef B200_CUDA_EMU_H
#define B200_CUDA_EMU_H
#ifndef B200_EMU
#define B200_EMU 1
#endif

#include <cstdint>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <algorithm>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __grid_constant__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __constant__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#ifndef __restrict__
#define __restrict__ __restrict
#endif

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct short2 { short x, y; };
struct short4 { short x, y, z, w; };
struct ushort2 { unsigned short x, y; };
struct ushort4 { unsigned short x, y, z, w; };
struct uchar2 { unsigned char x, y; };
struct uchar4 { unsigned char x, y, z, w; };
struct char4 { signed char x, y, z, w; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline short2 make_short2(short x, short y) { return short2{x, y}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }

/* ---------------------------------------------------------------- fibers */
struct EmuBlock;
struct EmuFiber {
    void *sp;
    void *stack;
    uint3 tid, bid;
    EmuBlock *blk;
    int lane, warp, linear;
    bool done;
};
struct EmuWarp {
    unsigned arrived; unsigned gen;
    uint64_t slot[32];
};
struct EmuBlock {
    dim3 bdim, gdim;
    int nthreads, alive;
    unsigned bar_arrived, bar_gen;
    std::vector<EmuWarp> warps;
    void *dyn_smem;
};
extern EmuFiber *emu_cur;
extern int emu_resident_blocks;
void emu_yield();
void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);

#define threadIdx (emu_cur->tid)
#define blockIdx (emu_cur->bid)
#define blockDim (emu_cur->blk->bdim)
#define gridDim (emu_cur->blk->gdim)
#define warpSize 32
#define B200_EMU_DYN_SMEM (emu_cur->blk->dyn_smem)

static inline void __syncthreads() {
    EmuBlock *b = emu_cur->blk;
    const unsigned g = b->bar_gen;
    if (++b->bar_arrived >= (unsigned)b->alive) { b->bar_arrived = 0; b->bar_gen++; return; }
    while (b->bar_gen == g) emu_yield();
}
static inline void emu_warp_barrier() {
    EmuBlock *b = emu_cur->blk;
    EmuWarp &w = b->warps[emu_cur->warp];
    const int nw = std::min(32, b->nthreads - emu_cur->warp * 32);
    const unsigned g = w.gen;
    if (++w.arrived >= (unsigned)nw) { w.arrived = 0; w.gen++; return; }
    while (w.gen == g) emu_yield();
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { (void)mask; emu_warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <class T> static inline T emu_shfl_impl(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    EmuWarp &w = emu_cur->blk->warps[emu_cur->warp];
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    w.slot[emu_cur->lane] = raw;
    emu_warp_barrier();
    uint64_t got = w.slot[src_lane & 31];
    emu_warp_barrier();
    T r; memcpy(&r, &got, sizeof(T));
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int lane = emu_cur->lane, base = lane & ~(width - 1);
    return emu_shfl_impl(v, base + (src & (width - 1)));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    const int lane = emu_cur->lane, base = lane & ~(width - 1);
    int s = (lane ^ m); if (s >= base + width) s = lane;
    return emu_shfl_impl(v, s);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    const int lane = emu_cur->lane, base = lane & ~(width - 1);
    int s = lane - (int)d; if (s < base) s = lane;
    return emu_shfl_impl(v, s);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    const int lane = emu_cur->lane, base = lane & ~(width - 1);
    int s = lane + (int)d; if (s >= base + width) s = lane;
    return emu_shfl_impl(v, s);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) { int p = emu_shfl_impl(pred ? 1 : 0, i); r |= (unsigned)(p != 0) << i; }
    const int nw = std::min(32, emu_cur->blk->nthreads - emu_cur->warp * 32);
    if (nw < 32) r &= (1u << nw) - 1;
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) {
    const int nw = std::min(32, emu_cur->blk->nthreads - emu_cur->warp * 32);
    return __ballot_sync(m, p) == (nw == 32 ? 0xffffffffu : (1u << nw) - 1);
}

/* ---------------------------------------------------------------- intrinsics */
template <class T> static inline T __ldg(const T *p) { return *p; }
using std::min;
using std::max;
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline int __mul24(int a, int b) { return a * b; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    uint64_t v = ((uint64_t)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; i++) { unsigned sel = (s >> (4 * i)) & 7; r |= (unsigned)((v >> (8 * sel)) & 0xff) << (8 * i); }
    return r;
}
static inline int __dp4a(int a, int b, int c) {
    for (int i = 0; i < 4; i++) c += (int)(int8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i));
    return c;
}
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
    return c;
}
static inline int __dp4a_us(unsigned a, int b, int c) { /* helper: u8 x s8 */
    for (int i = 0; i < 4; i++) c += (int)((a >> (8 * i)) & 0xff) * (int)(int8_t)(b >> (8 * i));
    return c;
}
// 16x2 SIMD integer intrinsics (semantics per the CUDA math API: per-halfword, wrapping unless noted)
#define EMU_H2(expr_lo, expr_hi) ((unsigned)((expr_lo) & 0xffff) | ((unsigned)((expr_hi) & 0xffff) << 16))
static inline int emu_s16(unsigned v) { return (int)(int16_t)(v & 0xffff); }
static inline unsigned __vadd2(unsigned a, unsigned b) { return EMU_H2((a & 0xffff) + (b & 0xffff), (a >> 16) + (b >> 16)); }
static inline unsigned __vsub2(unsigned a, unsigned b) { return EMU_H2((a & 0xffff) - (b & 0xffff), (a >> 16) - (b >> 16)); }
static inline unsigned __vmaxs2(unsigned a, unsigned b) { return EMU_H2(std::max(emu_s16(a), emu_s16(b)), std::max(emu_s16(a >> 16), emu_s16(b >> 16))); }
static inline unsigned __vmins2(unsigned a, unsigned b) { return EMU_H2(std::min(emu_s16(a), emu_s16(b)), std::min(emu_s16(a >> 16), emu_s16(b >> 16))); }
static inline unsigned __vmaxu2(unsigned a, unsigned b) { return EMU_H2(std::max(a & 0xffff, b & 0xffff), std::max(a >> 16, b >> 16)); }
static inline unsigned __vminu2(unsigned a, unsigned b) { return EMU_H2(std::min(a & 0xffff, b & 0xffff), std::min(a >> 16, b >> 16)); }
static inline unsigned __vimin_s16x2_relu(unsigned a, unsigned b) { return EMU_H2(std::max(0, std::min(emu_s16(a), emu_s16(b))), std::max(0, std::min(emu_s16(a >> 16), emu_s16(b >> 16)))); }
static inline unsigned __vimax_s16x2_relu(unsigned a, unsigned b) { return EMU_H2(std::max(0, std::max(emu_s16(a), emu_s16(b))), std::max(0, std::max(emu_s16(a >> 16), emu_s16(b >> 16)))); }
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vadd2(a, b), c); }
static inline unsigned __viaddmin_s16x2(unsigned a, unsigned b, unsigned c) { return __vmins2(__vadd2(a, b), c); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { sh &= 31; return sh ? (hi << sh) | (lo >> (32 - sh)) : hi; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline void __nanosleep(unsigned) { emu_yield(); }
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicSub(T *p, T v) { T o = *p; *p = o - v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; *p = std::max(o, v); return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; *p = std::min(o, v); return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

/* ---------------------------------------------------------------- runtime API stubs */
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
// "device" memory is poisoned on allocation (like real device memory it is NOT zero): a kernel that reads what nothing wrote
// shows up as a mismatch in the emulator tests too (B200_EMU_POISON=0 switches the fill off)
static inline cudaError_t cudaMalloc(void **p, size_t n)
{
    const size_t sz = (n + 255) & ~(size_t)255;
    *p = n ? aligned_alloc(256, sz) : nullptr;
    static int poison = -1;
    if (poison < 0) { const char *e = getenv("B200_EMU_POISON"); poison = !e || atoi(e) != 0; }
    if (*p && poison) memset(*p, 0xA5, sz);
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { if (n) memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = 0) {
    for (size_t y = 0; y < h; y++) memcpy((char *)d + y * dp, (const char *)s + y * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
#define cudaMemcpyToSymbol(sym, src, n) (memcpy((void *)&(sym), (src), (n)), cudaSuccess)
#define cudaMemcpyToSymbolAsync(sym, src, n, off, kind, st) (memcpy((char *)&(sym) + (off), (src), (n)), cudaSuccess)

/* launch: body is run once per thread */
#define B200_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu_launch((grid), (block), (smem), [=]() { kern(__VA_ARGS__); })

#endif
