// Whole-frame job: sequences the batched kernels of one frame on a stream (include/b200av1.h,
// B200FrameJob). The device-side counterpart of dav1d's per-frame task graph
// (TILE_RECONSTRUCTION -> DEBLOCK_COLS -> DEBLOCK_ROWS -> CDEF -> LOOP_RESTORATION,
// reference src/thread_task.c:699-854) with whole-frame stages instead of superblock rows.
#include "host_util.h"

extern "C" {

// the stages before intra reconstruction; *fg_side is set when the film grain preparation was forked to a side stream
static int frame_phase_recon(const B200FrameJob *j, void *stream, void **fg_side)
{
    int r;
    const int bd = j->bitdepth_max;
    *fg_side = nullptr;
#ifndef B200_EMU
    // film grain LUT preparation: one CTA, latency bound, depends only on the frame header -> side stream. The
    // scratch is reused frame after frame on this stream, hence the fork (after everything enqueued so far).
    SideStream *fs = nullptr;
    if (j->run_fg && (fs = side_stream_for((cudaStream_t)stream, 0)) && fs->fork((cudaStream_t)stream)) {
        if ((r = b200_fg_prep(bd, &j->fg, fs->side))) return r;
        *fg_side = fs;
    }
#endif
    if (j->n_expand > 0) {
        B200_CUDA_OK(cudaMemsetAsync(j->d_coef, 0, j->coef_bytes, (cudaStream_t)stream));
        if ((r = b200_coef_expand(bd, j->d_expand, j->n_expand, j->d_ccoef, j->d_coef, stream))) return r;
    }
    if ((r = b200_mc_batch(bd, &j->mc, j->d_pred, j->n_pred, stream))) return r;
    if ((r = b200_mc_scaled_batch(bd, &j->mc, j->d_scaled, j->n_scaled, stream))) return r;
    if ((r = b200_mc_warp_batch(bd, &j->mc, j->d_warp, j->n_warp, stream))) return r;
    if ((r = b200_mc_comp_fused_batch(bd, &j->mc, j->d_cfused, j->n_cfused, stream))) return r;
    if ((r = b200_mc_comp_fused_batch(bd, &j->mc, j->d_cfused2, j->n_cfused2, stream))) return r;
    if ((r = b200_mc_comp_batch(bd, &j->mc, j->d_comp, j->n_comp, stream))) return r;
    if ((r = b200_mc_comp_batch(bd, &j->mc, j->d_comp2, j->n_comp2, stream))) return r;
    if ((r = b200_mc_blend_batch(bd, &j->mc, j->d_blend, j->n_blend, stream))) return r;
    if ((r = b200_mc_blend_batch(bd, &j->mc, j->d_blend2, j->n_blend2, stream))) return r;
    if ((r = b200_itx_add_frame(bd, (const void *const *)j->d_itx, j->n_itx, j->d_coef, j->mc.dst, j->itx_stride,
                                j->zero_coefs, stream)))
        return r;
    return 0;
}

static int frame_phase_post(const B200FrameJob *j, void *stream, void *fg_side)
{
    int r;
    const int bd = j->bitdepth_max;
    if (j->run_lf && (r = b200_lf_frame(bd, &j->lf, stream))) return r;
    if (j->run_cdef && (r = b200_cdef_frame(bd, &j->cdef, stream))) return r;
    if (j->run_lr && (r = b200_lr_frame(bd, &j->lr, stream))) return r;
    if (j->run_fg) {
#ifndef B200_EMU
        if (fg_side && !((SideStream *)fg_side)->join((cudaStream_t)stream)) { b200_set_error("b200_frame_run: stream join failed"); return -1; }
#endif
        if (!fg_side && (r = b200_fg_prep(bd, &j->fg, stream))) return r;
        if ((r = b200_fg_apply(bd, &j->fg, stream))) return r;
    }
    return 0;
}

int b200_frame_run(const B200FrameJob *j, void *stream)
{
    int r;
    void *fg_side;
    if ((r = frame_phase_recon(j, stream, &fg_side))) return r;
    if (j->n_intra > 0 && (r = b200_intra_frame(j->bitdepth_max, &j->intra, j->d_intra, j->n_intra, stream))) return r;
    return frame_phase_post(j, stream, fg_side);
}

int b200_frame_run_batch(const B200FrameJob *const *jobs, int n, void *stream)
{
    if (n <= 0) return 0;
    if (n > 256) { b200_set_error("b200_frame_run_batch: too many jobs"); return -2; }
    int r;
    void *fg_side[256];
    B200IntraFrame frames[256];
    const B200IntraTx *tx[256];
    int32_t ntx[256];
    for (int i = 0; i < n; i++) {
        if (jobs[i]->bitdepth_max != jobs[0]->bitdepth_max) { b200_set_error("b200_frame_run_batch: mixed bit depths"); return -2; }
        if (jobs[i]->run_fg) { b200_set_error("b200_frame_run_batch: film grain jobs must be run one by one"); return -2; }
        if ((r = frame_phase_recon(jobs[i], stream, &fg_side[i]))) return r;
        frames[i] = jobs[i]->intra; tx[i] = jobs[i]->d_intra; ntx[i] = jobs[i]->n_intra;
    }
    if ((r = b200_intra_frames(jobs[0]->bitdepth_max, frames, tx, ntx, n, stream))) return r;
    for (int i = 0; i < n; i++)
        if ((r = frame_phase_post(jobs[i], stream, fg_side[i]))) return r;
    return 0;
}

int b200_struct_size(int which)
{
    switch (which) {
    case 0: return sizeof(B200McFrame); case 1: return sizeof(B200McBlock); case 2: return sizeof(B200CompBlock);
    case 3: return sizeof(B200BlendBlock); case 4: return sizeof(B200WarpBlock); case 5: return sizeof(B200ItxBlock);
    case 6: return sizeof(B200LfFrame); case 7: return sizeof(B200CdefFrame); case 8: return sizeof(B200LrFrame);
    case 9: return sizeof(B200FrameJob); case 10: return sizeof(B200Av1Filter); case 11: return sizeof(B200Av1Restoration);
    case 12: return sizeof(B200FgFrame); case 13: return sizeof(B200FilmGrainData);
    case 14: return sizeof(B200IntraTx); case 15: return sizeof(B200IntraFrame); case 16: return sizeof(B200McScaledBlock); case 17: return sizeof(B200CoefBlock); case 18: return sizeof(B200IntraSb); case 19: return sizeof(B200CompFusedBlock);
    }
    return -1;
}

int b200_frame_submit_host(const B200FrameJob *job, const B200Xfer *up, int n_up, const B200Xfer *down, int n_down,
                           void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < n_up; i++)
        if (up[i].bytes) B200_CUDA_OK(cudaMemcpyAsync(up[i].dev, up[i].host, up[i].bytes, cudaMemcpyHostToDevice, st));
    int r = b200_frame_run(job, stream);
    if (r) return r;
    for (int i = 0; i < n_down; i++)
        if (down[i].bytes) B200_CUDA_OK(cudaMemcpyAsync(down[i].host, down[i].dev, down[i].bytes, cudaMemcpyDeviceToHost, st));
    return 0;
}

int b200_frame_submit_host_batch(const B200FrameJob *const *jobs, int n_jobs, const B200Xfer *up, int n_up,
                                 const B200Xfer *down, int n_down, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < n_up; i++)
        if (up[i].bytes) B200_CUDA_OK(cudaMemcpyAsync(up[i].dev, up[i].host, up[i].bytes, cudaMemcpyHostToDevice, st));
    int r = b200_frame_run_batch(jobs, n_jobs, stream);
    if (r) return r;
    for (int i = 0; i < n_down; i++)
        if (down[i].bytes) B200_CUDA_OK(cudaMemcpyAsync(down[i].host, down[i].dev, down[i].bytes, cudaMemcpyDeviceToHost, st));
    return 0;
}

int b200_frame_wait(void *stream)
{
    B200_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

int b200_frame_run_host(const B200FrameJob *job, const B200Xfer *up, int n_up, const B200Xfer *down, int n_down,
                        void *stream)
{
    int r = b200_frame_submit_host(job, up, n_up, down, n_down, stream);
    return r ? r : b200_frame_wait(stream);
}

}  // extern "C"
