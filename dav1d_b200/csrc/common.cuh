// Shared device/host helpers for the sm_100a kernels of the AV1 reconstruction back end.
#pragma once
#ifndef B200_EMU
#include <cuda_runtime.h>
#define B200_LAUNCH(kern, grid, block, smem, stream, ...) \
    kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
#include <stdint.h>
#include <stddef.h>

#define B200_DEV __device__ __forceinline__
#define B200_HD __host__ __device__ __forceinline__

namespace b200 {

B200_HD int imin(int a, int b) { return a < b ? a : b; }
B200_HD int imax(int a, int b) { return a > b ? a : b; }
B200_HD int iclip(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
B200_HD int iabs(int v) { return v < 0 ? -v : v; }
B200_HD int ulog2(unsigned v) {
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

// pixel / coefficient types per bit-depth class (reference include/common/bitdepth.h:42-86)
template <bool HBD> struct Bd;
template <> struct Bd<false> { typedef uint8_t pixel; typedef int16_t coef; };
template <> struct Bd<true> { typedef uint16_t pixel; typedef int32_t coef; };

}  // namespace b200

// error plumbing for the C ABI (capi.cu)
void b200_set_error(const char *fmt, ...);
#define B200_CUDA_OK(expr)                                                                   \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            b200_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)
