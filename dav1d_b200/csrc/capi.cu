// C ABI of the back end (include/b200av1.h): error plumbing, the Level-1 drop-in function
// tables (record -> launch -> sync shims with dav1d's exact signatures) and the Level-2
// batched entry points. No CPU fallback anywhere: every path ends in a kernel launch.
#include "common.cuh"
#include "../../include/b200av1.h"
#include "host_util.h"
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <utility>

namespace b200 {
int launch_itx_grouped(bool hbd, const void *const *blocks, const int32_t *n, void *coefs, void *pic, const int32_t *st,
                       int bdmax, int zero, cudaStream_t stream);
int launch_itx(int tx, bool hbd, const B200ItxBlock *blocks, int n, void *coefs, void *pic,
               const int32_t *st, int bdmax, int zero, cudaStream_t stream);
}

static thread_local char g_err[512];
namespace b200 { std::mutex &host_lock() { static std::mutex m; return m; } }
static std::atomic<uint64_t> g_launches{0};

void b200_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void b200_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static std::atomic<int> g_pdl{-1};
bool b200_pdl_enabled() {
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0) { v = getenv("B200_NO_PDL") ? 0 : 1; g_pdl.store(v, std::memory_order_relaxed); }
    return v != 0;
}

extern "C" {

int b200_version(void) { return 100; }
const char *b200_last_error(void) { return g_err; }
uint64_t b200_launch_count(void) { return g_launches.load(); }
void b200_set_pdl(int on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }

void *b200_dev_alloc(size_t bytes) {
    void *p = nullptr;
    const cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { b200_set_error("b200_dev_alloc(%zu): %s", bytes, cudaGetErrorString(e)); return nullptr; }
    return p;
}
void b200_dev_free(void *p) { if (p) cudaFree(p); }
void *b200_host_alloc(size_t bytes) {
    void *p = nullptr;
    const cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { b200_set_error("b200_host_alloc(%zu): %s", bytes, cudaGetErrorString(e)); return nullptr; }
    return p;
}
void b200_host_free(void *p) { if (p) cudaFreeHost(p); }
void *b200_stream_create(void) {
    cudaStream_t s = nullptr;
    const cudaError_t e = cudaStreamCreate(&s);
    if (e != cudaSuccess) { b200_set_error("b200_stream_create: %s", cudaGetErrorString(e)); return nullptr; }
#ifdef B200_EMU
    if (!s) return (void *)(uintptr_t)1;      // the host emulator has no stream objects; NULL means failure to callers
#endif
    return (void *)s;
}
void b200_stream_destroy(void *stream) { if (stream) cudaStreamDestroy((cudaStream_t)stream); }
int b200_dev_memset(void *p, int value, size_t bytes, void *stream) {
    B200_CUDA_OK(cudaMemsetAsync(p, value, bytes, (cudaStream_t)stream));
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
namespace {
using b200::Scratch;
Scratch g_s_blocks, g_s_coef, g_s_pic;
#define g_mu (b200::host_lock())

const uint8_t k_tx_w[19] = { 4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64 };
const uint8_t k_tx_h[19] = { 4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16 };

// is (tx, txtp) a slot dav1d defines? (reference src/itx_tmpl.c:220-288)
bool itx_defined(int tx, int txtp) {
    if (tx < 0 || tx >= 19 || txtp < 0 || txtp > 16) return false;
    if (txtp == 16) return tx == 0;
    const int w = k_tx_w[tx], h = k_tx_h[tx], mx = w > h ? w : h, mn = w < h ? w : h;
    if (mx == 64) return txtp == 0;
    if (mx == 32) return txtp == 0 || txtp == 9;
    if (mx == 16 && mn == 16) return txtp <= 11;
    return true;
}

using b200::die;
}  // namespace

extern "C" {

int b200_itx_add_batch(int bitdepth_max, int tx, const B200ItxBlock *d_blocks, int n_blocks,
                       void *d_coef, void *d_pic, const int32_t stride_px[3], int zero_coefs,
                       void *stream)
{
    if (tx < 0 || tx >= 19) { b200_set_error("b200_itx_add_batch: bad tx %d", tx); return -2; }
    if (bitdepth_max != 255 && bitdepth_max != 1023 && bitdepth_max != 4095) {
        b200_set_error("b200_itx_add_batch: bad bitdepth_max %d", bitdepth_max);
        return -2;
    }
    if (n_blocks <= 0) return 0;
    if (b200::launch_itx(tx, bitdepth_max > 255, d_blocks, n_blocks, d_coef, d_pic, stride_px,
                         bitdepth_max, zero_coefs, (cudaStream_t)stream))
        { b200_set_error("b200_itx_add_batch: launch failed"); return -1; }
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

int b200_itx_add_frame(int bitdepth_max, const void *const d_blocks[19], const int32_t n_blocks[19], void *d_coef,
                       void *d_pic, const int32_t stride_px[3], int zero_coefs, void *stream)
{
    if (bitdepth_max != 255 && bitdepth_max != 1023 && bitdepth_max != 4095) {
        b200_set_error("b200_itx_add_frame: bad bitdepth_max %d", bitdepth_max);
        return -2;
    }
    if (b200::launch_itx_grouped(bitdepth_max > 255, d_blocks, n_blocks, d_coef, d_pic, stride_px, bitdepth_max,
                                 zero_coefs, (cudaStream_t)stream))
        { b200_set_error("b200_itx_add_frame: launch failed"); return -1; }
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

int b200_itx_add_batch_host(int bitdepth_max, int tx, const B200ItxBlock *blocks, int n_blocks,
                            void *coef, size_t coef_bytes, void *pic, size_t pic_bytes,
                            const int32_t stride_px[3], int zero_coefs)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (n_blocks <= 0) return 0;
    const size_t bb = (size_t)n_blocks * sizeof(B200ItxBlock);
    if (g_s_blocks.reserve(bb) || g_s_coef.reserve(coef_bytes) || g_s_pic.reserve(pic_bytes)) return -1;
    cudaStream_t st = 0;
    B200_CUDA_OK(cudaMemcpyAsync(g_s_blocks.p, blocks, bb, cudaMemcpyHostToDevice, st));
    B200_CUDA_OK(cudaMemcpyAsync(g_s_coef.p, coef, coef_bytes, cudaMemcpyHostToDevice, st));
    B200_CUDA_OK(cudaMemcpyAsync(g_s_pic.p, pic, pic_bytes, cudaMemcpyHostToDevice, st));
    int r = b200_itx_add_batch(bitdepth_max, tx, (const B200ItxBlock *)g_s_blocks.p, n_blocks,
                               g_s_coef.p, g_s_pic.p, stride_px, zero_coefs, st);
    if (r) return r;
    B200_CUDA_OK(cudaMemcpyAsync(pic, g_s_pic.p, pic_bytes, cudaMemcpyDeviceToHost, st));
    if (zero_coefs)
        B200_CUDA_OK(cudaMemcpyAsync(coef, g_s_coef.p, coef_bytes, cudaMemcpyDeviceToHost, st));
    B200_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

// Level-1 single call, host pointers, arbitrary (possibly negative) byte stride.
int b200_inv_txfm_add(void *dst, ptrdiff_t dst_stride, void *coeff, int eob, int tx, int txtp,
                      int bitdepth_max)
{
    if (!itx_defined(tx, txtp)) { b200_set_error("b200_inv_txfm_add: undefined (tx=%d, txtp=%d)", tx, txtp); return -2; }
    if (eob < 0) { b200_set_error("b200_inv_txfm_add: eob < 0"); return -2; }
    const bool hbd = bitdepth_max > 255;
    const int w = k_tx_w[tx], h = k_tx_h[tx];
    const int sw = w < 32 ? w : 32, sh = h < 32 ? h : 32;
    const size_t px = hbd ? 2 : 1, cs = hbd ? 4 : 2;
    // pack the w x h destination rectangle densely (handles negative strides)
    uint8_t rect[64 * 64 * 2];
    for (int y = 0; y < h; y++)
        memcpy(rect + (size_t)y * w * px, (const uint8_t *)dst + (ptrdiff_t)y * dst_stride, (size_t)w * px);
    B200ItxBlock b;
    b.dst_off = 0; b.coef_off = 0; b.eob = (int16_t)eob; b.txtp = (uint8_t)txtp; b.plane = 0;
    const int32_t st[3] = { w, w, w };
    int r = b200_itx_add_batch_host(bitdepth_max, tx, &b, 1, coeff, (size_t)sw * sh * cs, rect,
                                    (size_t)w * h * px, st, 1);
    if (r) return r;
    for (int y = 0; y < h; y++)
        memcpy((uint8_t *)dst + (ptrdiff_t)y * dst_stride, rect + (size_t)y * w * px, (size_t)w * px);
    return 0;
}

}  // extern "C"

// ---- Level-1 function tables: one thunk per (tx, txtp) slot, dav1d signatures -------------
namespace {
template <int TX, int TXTP>
void itx_thunk8(uint8_t *dst, ptrdiff_t stride, int16_t *coeff, int eob) {
    if (b200_inv_txfm_add(dst, stride, coeff, eob, TX, TXTP, 255)) die("itxfm_add (8 bpc)");
}
template <int TX, int TXTP>
void itx_thunk16(uint16_t *dst, ptrdiff_t stride, int32_t *coeff, int eob, int bitdepth_max) {
    if (b200_inv_txfm_add(dst, stride, coeff, eob, TX, TXTP, bitdepth_max)) die("itxfm_add (16 bpc)");
}
template <int TX, int... TP>
void fill_row(B200InvTxfmDSPContext8 *c8, B200InvTxfmDSPContext16 *c16, std::integer_sequence<int, TP...>) {
    if (c8)  { ((c8->itxfm_add[TX][TP]  = itx_defined(TX, TP) ? itx_thunk8<TX, TP>  : nullptr), ...); }
    if (c16) { ((c16->itxfm_add[TX][TP] = itx_defined(TX, TP) ? itx_thunk16<TX, TP> : nullptr), ...); }
}
template <int... TX>
void fill_all(B200InvTxfmDSPContext8 *c8, B200InvTxfmDSPContext16 *c16, std::integer_sequence<int, TX...>) {
    (fill_row<TX>(c8, c16, std::make_integer_sequence<int, B200_N_TX_TYPES_PLUS_LL>{}), ...);
}
}  // namespace

extern "C" {
void b200_itx_dsp_init_8bpc(B200InvTxfmDSPContext8 *c, int bpc) {
    (void)bpc;
    fill_all(c, nullptr, std::make_integer_sequence<int, B200_N_RECT_TX_SIZES>{});
}
void b200_itx_dsp_init_16bpc(B200InvTxfmDSPContext16 *c, int bpc) {
    (void)bpc;
    fill_all(nullptr, c, std::make_integer_sequence<int, B200_N_RECT_TX_SIZES>{});
}
}
