// Batched inverse transform + add (dav1d Dav1dInvTxfmDSPContext, reference src/itx_tmpl.c:43-203).
//
// Work decomposition (all sizes 4x4 .. 64x64, all 16 transform types + WHT):
//   * a transform block of w x h is owned by a group of L = max(min(h,32), min(w,32)) lanes
//     (32/L blocks per warp, 4 warps per CTA);
//   * pass 1: lane y holds coefficient row y (length w) in registers, runs the horizontal
//     1-D transform, applies the inter-pass rounding/clip and parks the row in a padded
//     shared-memory tile (pitch w+1 words: conflict-free both ways);
//   * pass 2: lane x pulls column x (length h) back into registers, runs the vertical 1-D
//     transform and does the read-modify-write of the picture column, so that consecutive
//     lanes touch consecutive pixels of a row (coalesced).
// Integer only, bit-exact with the reference C path for every input (including coefficient
// garbage: rows beyond the eob-derived bound are treated as zero exactly like the reference).
#include "itx_body.cuh"
#include "host_util.h"
#include "launch_count.h"

namespace b200 {

template <int W, int H, int TX, int SHIFT, bool HBD>
__global__ void __launch_bounds__(kItxWarps * 32)
itx_add_kernel(const B200ItxBlock *__restrict__ blocks, int n_blocks,
               typename Bd<HBD>::coef *__restrict__ coefs, typename Bd<HBD>::pixel *__restrict__ pic,
               int stride0, int stride1, int stride2, int bitdepth_max, int zero_coefs)
{
    B200_PDL_ENTRY();
    typedef ItxGeom<W, H> G;
    __shared__ int tile[ItxGeom<W, H>::BPC * G::SLOT];
    itx_add_body<W, H, TX, SHIFT, HBD>(blockIdx.x, tile, blocks, n_blocks, coefs, pic, stride0, stride1, stride2,
                                       bitdepth_max, zero_coefs);
}

// ---- all transform sizes of a frame in two launches (one per register class), CTAs dealt largest size first ---
struct ItxGroups {
    const B200ItxBlock *blocks[B200_N_RECT_TX_SIZES];
    int n[B200_N_RECT_TX_SIZES];
    int cta_begin[B200_N_RECT_TX_SIZES], cta_end[B200_N_RECT_TX_SIZES];
};

// Two register classes, one launch each: sizes with a 64-point dimension (long butterflies, up to ~170 live
// registers, 33 KB tile) and everything else (<= 64 registers, 17 KB tile, 8 CTAs per SM).
#ifndef B200_ITX_SMALL_MINB
#define B200_ITX_SMALL_MINB 7
#endif
template <bool BIG> struct ItxClass {
    static constexpr int kMinCtas = BIG ? 3 : B200_ITX_SMALL_MINB;
    static constexpr int kTileWords = kItxWarps * (BIG ? ItxGeom<64, 64>::NB * ItxGeom<64, 64>::SLOT
                                                       : ItxGeom<32, 32>::NB * ItxGeom<32, 32>::SLOT);
};

template <bool HBD, bool BIG>
__global__ void __launch_bounds__(kItxWarps * 32, ItxClass<BIG>::kMinCtas)
itx_add_grouped_kernel(const __grid_constant__ ItxGroups g, typename Bd<HBD>::coef *__restrict__ coefs, typename Bd<HBD>::pixel *__restrict__ pic,
                       int stride0, int stride1, int stride2, int bitdepth_max, int zero_coefs)
{
    B200_PDL_ENTRY();
    __shared__ int tile[ItxClass<BIG>::kTileWords];
    const int c = blockIdx.x;
#define X(TX, W, H, SH) \
    if constexpr ((W == 64 || H == 64) == BIG) { \
        if (c < g.cta_end[TX]) { \
            itx_add_body<W, H, TX, SH, HBD>(c - g.cta_begin[TX], tile, g.blocks[TX], g.n[TX], coefs, pic, stride0, stride1, \
                                            stride2, bitdepth_max, zero_coefs); \
            return; \
        } \
    }
    B200_ITX_SIZES(X)
#undef X
}

int launch_itx_grouped(bool hbd, const void *const *blocks, const int32_t *n, void *coefs, void *pic, const int32_t *st,
                       int bdmax, int zero, cudaStream_t stream)
{
    for (int big = 1; big >= 0; big--) {
        ItxGroups g;
        int total = 0;
#define X(TX, W, H, SH) { \
            const int per_cta = ItxGeom<W, H>::BPC; \
            const int mine = ((W == 64 || H == 64) ? 1 : 0) == big; \
            const int ctas = (mine && n[TX] > 0) ? (n[TX] + per_cta - 1) / per_cta : 0; \
            g.blocks[TX] = (const B200ItxBlock *)blocks[TX]; g.n[TX] = n[TX] > 0 ? n[TX] : 0; \
            g.cta_begin[TX] = total; total += ctas; g.cta_end[TX] = total; }
        B200_ITX_SIZES(X)
#undef X
        if (!total) continue;
        if (hbd) {
            if (big) { auto k = itx_add_grouped_kernel<true, true>; B200_LAUNCH_PDL(k, dim3(total), dim3(kItxWarps * 32), 0, stream, g, (int32_t *)coefs, (uint16_t *)pic, st[0], st[1], st[2], bdmax, zero); }
            else { auto k = itx_add_grouped_kernel<true, false>; B200_LAUNCH_PDL(k, dim3(total), dim3(kItxWarps * 32), 0, stream, g, (int32_t *)coefs, (uint16_t *)pic, st[0], st[1], st[2], bdmax, zero); }
        } else {
            if (big) { auto k = itx_add_grouped_kernel<false, true>; B200_LAUNCH_PDL(k, dim3(total), dim3(kItxWarps * 32), 0, stream, g, (int16_t *)coefs, (uint8_t *)pic, st[0], st[1], st[2], bdmax, zero); }
            else { auto k = itx_add_grouped_kernel<false, false>; B200_LAUNCH_PDL(k, dim3(total), dim3(kItxWarps * 32), 0, stream, g, (int16_t *)coefs, (uint8_t *)pic, st[0], st[1], st[2], bdmax, zero); }
        }
        b200_count_launch();
    }
    return 0;
}

template <int W, int H, int TX, int SHIFT>
static int launch_itx_wh(bool hbd, const B200ItxBlock *blocks, int n, void *coefs, void *pic,
                         const int32_t *st, int bdmax, int zero, cudaStream_t stream)
{
    typedef ItxGeom<W, H> G;
    const int per_cta = ItxGeom<W, H>::BPC;
    const int grid = (n + per_cta - 1) / per_cta;
    if (grid <= 0) return 0;
    if (hbd) {
        auto k = itx_add_kernel<W, H, TX, SHIFT, true>;
        B200_LAUNCH(k, dim3(grid), dim3(kItxWarps * 32), 0, stream, blocks, n, (int32_t *)coefs,
                    (uint16_t *)pic, st[0], st[1], st[2], bdmax, zero);
    } else {
        auto k = itx_add_kernel<W, H, TX, SHIFT, false>;
        B200_LAUNCH(k, dim3(grid), dim3(kItxWarps * 32), 0, stream, blocks, n, (int16_t *)coefs,
                    (uint8_t *)pic, st[0], st[1], st[2], bdmax, zero);
    }
    b200_count_launch();
    return 0;
}

// tx -> (w, h, inter-pass shift): reference src/itx_tmpl.c:160-178
int launch_itx(int tx, bool hbd, const B200ItxBlock *blocks, int n, void *coefs, void *pic,
               const int32_t *st, int bdmax, int zero, cudaStream_t stream)
{
#define CASE(TX, W, H, SH) case TX: return launch_itx_wh<W, H, TX, SH>(hbd, blocks, n, coefs, pic, st, bdmax, zero, stream)
    switch (tx) {
    CASE(0, 4, 4, 0);
    CASE(1, 8, 8, 1);
    CASE(2, 16, 16, 2);
    CASE(3, 32, 32, 2);
    CASE(4, 64, 64, 2);
    CASE(5, 4, 8, 0);
    CASE(6, 8, 4, 0);
    CASE(7, 8, 16, 1);
    CASE(8, 16, 8, 1);
    CASE(9, 16, 32, 1);
    CASE(10, 32, 16, 1);
    CASE(11, 32, 64, 1);
    CASE(12, 64, 32, 1);
    CASE(13, 4, 16, 1);
    CASE(14, 16, 4, 1);
    CASE(15, 8, 32, 2);
    CASE(16, 32, 8, 2);
    CASE(17, 16, 64, 2);
    CASE(18, 64, 16, 2);
    }
#undef CASE
    return -1;
}

}  // namespace b200
