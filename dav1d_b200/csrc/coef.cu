// Coefficient stream expansion. The host emitter ships, per coded transform block, only the eob + 1 coefficients
// that can be non-zero, in scan order (what dav1d's decode_coefs walks, reference src/recon_tmpl.c:318-730, before
// it scatters them into the dense frame_thread.cf plane, src/decode.c:2852-2863). This kernel rebuilds the dense
// min(w,32) x min(h,32) blocks the transform kernels read: dense[scan[k]] = compact[k]. The dense buffer is zeroed
// first by the caller (b200_frame_run). One warp per block. Cuts the host->device traffic of a frame ~3x.
#include "host_util.h"
#define B200_SCAN_TBL __device__
#include "scan_gen.h"
#include "launch_count.h"

namespace b200 {

template <class coef>
__global__ void __launch_bounds__(128) coef_expand_kernel(const B200CoefBlock *__restrict__ recs, int n,
                                                          const coef *__restrict__ compact, coef *__restrict__ dense)
{
    B200_PDL_ENTRY();
    const int wi = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (wi >= n) return;
    const B200CoefBlock r = recs[wi];
    const uint16_t *scan = b200_scan + b200_scan_off[r.tx];
    const coef *src = compact + r.compact_off;
    coef *dst = dense + r.dense_off;
    for (int k = lane; k <= r.eob; k += 32) dst[scan[k]] = src[k];
}

}  // namespace b200

extern "C" {

int b200_coef_expand(int bitdepth_max, const B200CoefBlock *d_blocks, int n_blocks, const void *d_compact, void *d_dense,
                     void *stream)
{
    if (bitdepth_max != 255 && bitdepth_max != 1023 && bitdepth_max != 4095) { b200_set_error("b200_coef_expand: bad bitdepth_max"); return -2; }
    if (n_blocks <= 0) return 0;
    using namespace b200;
    const dim3 grid((n_blocks + 3) / 4);
    if (bitdepth_max > 255) { auto k = coef_expand_kernel<int32_t>; B200_LAUNCH_PDL(k, grid, dim3(128), 0, (cudaStream_t)stream, d_blocks, n_blocks, (const int32_t *)d_compact, (int32_t *)d_dense); }
    else { auto k = coef_expand_kernel<int16_t>; B200_LAUNCH_PDL(k, grid, dim3(128), 0, (cudaStream_t)stream, d_blocks, n_blocks, (const int16_t *)d_compact, (int16_t *)d_dense); }
    b200_count_launch();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}
