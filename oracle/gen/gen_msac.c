/*
 * oracle/gen/gen_msac.c — TEST INFRASTRUCTURE: a stream GENERATOR made of the reference decoder itself.
 *
 * Synthetic AV1 streams with random tile payloads (dav1d_b200/obu.py) follow the statistics of the default CDFs: about
 * half of the blocks carry residual, every transform block is dense. To make streams with CHOSEN statistics (mostly
 * skipped blocks, sparse coefficients, a given share of intra blocks) without writing an AV1 encoder, this file replaces
 * the symbol DEcoder of a dav1d build (src/msac.c is not compiled into oracle/_ref/libdav1d_gen.so, everything else is the
 * unmodified reference) with functions that
 *   1. CHOOSE the value of every symbol the decoder asks for — drawn from the context's own CDF (exactly what a random
 *      payload yields), except for a few syntax elements recognised by where their CDF lives inside the tile's CdfContext
 *      (block skip flag, intra / inter flag, all-zero flag of a transform block, end-of-block position), which follow a policy;
 *   2. adapt the CDF exactly as the decoder does (the arithmetic of src/msac.c:132-216 is normative) and keep the range
 *      `rng` the decoder would have;
 *   3. ENCODE the chosen value with the range encoder that is the inverse of that decoder (AV1's od_ec encoder: same
 *      partition of the range, a 16-bit pre-carry buffer, carries propagated when the tile is finished).
 * The decoder then parses a whole stream "from" a placeholder payload it never reads, reconstructs the pictures that
 * belong to its own choices, and leaves behind, tile by tile, the bytes that make any AV1 decoder take the same choices.
 * tests/streamgen.py puts those payloads into the headers obu.py wrote; the self-check is that stock dav1d decodes
 * the result to the very pictures the generator run produced.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "config.h"
#include "common/intops.h"
#include "src/internal.h"
#include "src/msac.h"

#define API __attribute__((visibility("default")))
#define EC_PROB_SHIFT 6
#define EC_MIN_PROB 4

/* ---- the range encoder -------------------------------------------------------------------------------------------- */
typedef struct Enc { uint64_t low; unsigned rng; int cnt; uint16_t *pre; size_t n, cap; } Enc;
static void enc_init(Enc *e) { memset(e, 0, sizeof(*e)); e->rng = 0x8000; e->cnt = -9; }
static void enc_push(Enc *e, unsigned v)
{
    if (e->n == e->cap) { e->cap = e->cap ? 2 * e->cap : 4096; e->pre = realloc(e->pre, e->cap * sizeof(*e->pre)); if (!e->pre) abort(); }
    e->pre[e->n++] = (uint16_t)v;
}
/* renormalise so that 32768 <= rng < 65536, moving finished bytes (with a possible carry bit on top) to the pre-carry buffer */
static void enc_norm(Enc *e, uint64_t low, unsigned rng)
{
    int c = e->cnt;
    const int d = 15 ^ (31 ^ clz(rng));
    int s = c + d;
    if (s >= 0) {
        c += 16;
        uint64_t m = ((uint64_t)1 << c) - 1;
        if (s >= 8) { enc_push(e, (unsigned)(low >> c)); low &= m; c -= 8; m >>= 8; }
        enc_push(e, (unsigned)(low >> c));
        s = c + d - 24;
        low &= m;
    }
    e->low = low << d; e->rng = rng << d; e->cnt = s;
}
static size_t enc_done(Enc *e, uint8_t **out)
{
    uint64_t l = e->low, m = 0x3fff, x = ((l + m) & ~m) | (m + 1);
    int c = e->cnt, s = 10 + c;
    if (s > 0) {
        uint64_t n = ((uint64_t)1 << (c + 16)) - 1;
        do { enc_push(e, (unsigned)(x >> (c + 16))); x &= n; s -= 8; c -= 8; n >>= 8; } while (s > 0);
    }
    uint8_t *b = malloc(e->n + 4);
    if (!b) abort();
    unsigned carry = 0;
    for (size_t i = e->n; i-- > 0; ) { carry += e->pre[i]; b[i] = (uint8_t)carry; carry >>= 8; }
    /* bytes beyond the end read as zero, which is what the decoder assumes past the buffer (src/msac.c:47-50): harmless slack
     * against its overread check (src/decode.c:2743) */
    memset(b + e->n, 0, 4);
    *out = b;
    const size_t n = e->n + 2;
    free(e->pre); e->pre = NULL; e->n = e->cap = 0;
    return n;
}

/* ---- tiles in the order the decoder initialises them ----------------------------------------------------------------- */
typedef struct Tile { MsacContext *s; Enc e; int open; uint8_t *bytes; size_t n_bytes; } Tile;
static Tile *g_tiles;
static int g_n_tiles, g_cap_tiles;
static struct { uint64_t rs; double p_skip, p_intra, p_txskip; int eob_draws, is422; } g_pol = { 88172645463325252ull, -1, -1, -1, 1, 0 };

static uint64_t rnd64(void) { uint64_t x = g_pol.rs; x ^= x << 13; x ^= x >> 7; x ^= x << 17; return g_pol.rs = x; }
static double rnd01(void) { return (double)(rnd64() >> 11) * (1.0 / 9007199254740992.0); }

static void tile_close(Tile *t) { if (t->open) { t->n_bytes = enc_done(&t->e, &t->bytes); t->open = 0; } }
static Tile *tile_of(MsacContext *s)
{
    for (int i = g_n_tiles - 1; i >= 0; i--)
        if (g_tiles[i].s == s && g_tiles[i].open) return &g_tiles[i];
    abort();
}

API void gen_reset(uint64_t seed, double p_skip, double p_intra, double p_txskip, int eob_draws)
{
    for (int i = 0; i < g_n_tiles; i++) { free(g_tiles[i].bytes); free(g_tiles[i].e.pre); }
    g_n_tiles = 0;
    g_pol.rs = seed * 0x9e3779b97f4a7c15ull + 0x1234567ull; if (!g_pol.rs) g_pol.rs = 1;
    g_pol.p_skip = p_skip; g_pol.p_intra = p_intra; g_pol.p_txskip = p_txskip; g_pol.eob_draws = eob_draws < 1 ? 1 : eob_draws;
}
/* 4:2:2 streams: partitions that would make 2-sample-wide chroma blocks are forbidden and the decoder rejects the tile
 * (reference src/decode.c:2151-2156, 2356): the generator then never chooses them */
API void gen_set_422(int on) { g_pol.is422 = on; }
API int gen_finish(void) { for (int i = 0; i < g_n_tiles; i++) tile_close(&g_tiles[i]); return g_n_tiles; }
API uint64_t gen_tile(int i, uint8_t *out, uint64_t cap)
{
    if (i < 0 || i >= g_n_tiles || g_tiles[i].open) return 0;
    if (out && cap >= g_tiles[i].n_bytes) memcpy(out, g_tiles[i].bytes, g_tiles[i].n_bytes);
    return g_tiles[i].n_bytes;
}

/* ---- the decoder's side: src/msac.c replaced -------------------------------------------------------------------------- */
void dav1d_msac_init(MsacContext *const s, const uint8_t *const data, const size_t sz, const int disable_cdf_update_flag)
{
    for (int i = 0; i < g_n_tiles; i++)
        if (g_tiles[i].s == s) tile_close(&g_tiles[i]);            /* the tile state is reused by the next frame */
    if (g_n_tiles == g_cap_tiles) { g_cap_tiles = g_cap_tiles ? 2 * g_cap_tiles : 64; g_tiles = realloc(g_tiles, g_cap_tiles * sizeof(*g_tiles)); if (!g_tiles) abort(); }
    Tile *const t = &g_tiles[g_n_tiles++];
    memset(t, 0, sizeof(*t));
    t->s = s; t->open = 1; enc_init(&t->e);
    s->buf_pos = data; s->buf_end = data + sz;                       /* never read */
    s->dif = 0; s->rng = 0x8000; s->cnt = 40;                        /* cnt: far from the overread threshold */
    s->allow_update_cdf = !disable_cdf_update_flag;
}

static inline void dec_norm(MsacContext *const s, const unsigned rng) { s->rng = rng << (15 ^ (31 ^ clz(rng))); }

/* which syntax element a CDF belongs to: by its address inside the tile's CdfContext */
enum { EL_OTHER, EL_SKIP, EL_INTRA, EL_TXSKIP, EL_EOB, EL_PARTITION };
static int element_of(const MsacContext *const s, const uint16_t *const cdf)
{
    const Dav1dTileState *const ts = (const Dav1dTileState *)((const char *)s - offsetof(Dav1dTileState, msac));
    const CdfContext *const C = &ts->cdf;
    const char *const p = (const char *)cdf;
#define IN(m) (p >= (const char *)&C->m && p < (const char *)&C->m + sizeof(C->m))
    if (IN(m.skip)) return EL_SKIP;
    if (IN(m.intra)) return EL_INTRA;
    if (IN(coef.skip)) return EL_TXSKIP;
    if (IN(m.partition)) return EL_PARTITION;
    if (p >= (const char *)&C->coef.eob_bin_16 && p < (const char *)&C->coef.eob_base_tok) return EL_EOB;
#undef IN
    return EL_OTHER;
}

/* NB: the decoder's dif counts DOWN from the top of the range: "dif < v" (bit = 1) is the interval of width v at the top,
 * which the encoder reaches by adding rng - v to low (od_ec_encode_bool_q15: if (val) l += r - v; r = val ? v : r - v). */

unsigned dav1d_msac_decode_bool_equi_c(MsacContext *const s)
{
    const unsigned r = s->rng, v = ((r >> 8) << 7) + EC_MIN_PROB;
    const unsigned bit = (unsigned)(rnd64() % r) < v;
    Tile *const t = tile_of(s);
    if (bit) { enc_norm(&t->e, t->e.low + (r - v), v); dec_norm(s, v); }
    else { enc_norm(&t->e, t->e.low, r - v); dec_norm(s, r - v); }
    return bit;
}

/* only caller: the split flag of a block that reaches over the frame edge (decode_sb) */
unsigned dav1d_msac_decode_bool_c(MsacContext *const s, const unsigned f)
{
    const unsigned r = s->rng;
    const unsigned v = ((r >> 8) * (f >> EC_PROB_SHIFT) >> (7 - EC_PROB_SHIFT)) + EC_MIN_PROB;
    const unsigned bit = g_pol.is422 ? 1 : (unsigned)(rnd64() % r) < v;
    Tile *const t = tile_of(s);
    if (bit) { enc_norm(&t->e, t->e.low + (r - v), v); dec_norm(s, v); }
    else { enc_norm(&t->e, t->e.low, r - v); dec_norm(s, r - v); }
    return bit;
}

unsigned dav1d_msac_decode_bool_adapt_c(MsacContext *const s, uint16_t *const cdf)
{
    const unsigned r = s->rng, f = *cdf;
    const unsigned v = ((r >> 8) * (f >> EC_PROB_SHIFT) >> (7 - EC_PROB_SHIFT)) + EC_MIN_PROB;
    unsigned bit = (unsigned)(rnd64() % r) < v;
    const int el = element_of(s, cdf);
    if (el == EL_SKIP && g_pol.p_skip >= 0) bit = rnd01() < g_pol.p_skip;
    else if (el == EL_INTRA && g_pol.p_intra >= 0) bit = !(rnd01() < g_pol.p_intra);       /* the flag says "is inter" */
    else if (el == EL_TXSKIP && g_pol.p_txskip >= 0) bit = rnd01() < g_pol.p_txskip;
    Tile *const t = tile_of(s);
    if (bit) { enc_norm(&t->e, t->e.low + (r - v), v); dec_norm(s, v); }
    else { enc_norm(&t->e, t->e.low, r - v); dec_norm(s, r - v); }
    if (s->allow_update_cdf) {                    /* update_cdf() for boolean CDFs, as the decoder does it */
        const unsigned count = cdf[1];
        const int rate = 4 + (count >> 4);
        if (bit) cdf[0] += (32768 - cdf[0]) >> rate;
        else cdf[0] -= cdf[0] >> rate;
        cdf[1] = count + (count < 32);
    }
    return bit;
}

unsigned dav1d_msac_decode_symbol_adapt_c(MsacContext *const s, uint16_t *const cdf, const size_t n_symbols)
{
    const unsigned r = s->rng >> 8;
    unsigned val = (unsigned)n_symbols + 1;
    /* draw from the CDF (several draws, smallest value, for the end-of-block position when the policy asks for short blocks) */
    const int el = element_of(s, cdf);
    const int draws = el == EL_EOB ? g_pol.eob_draws : 1;
    for (int k = 0; k < draws; k++) {
        unsigned cand;
        do {
            const unsigned c = (unsigned)(rnd64() % s->rng);
            unsigned v;
            cand = (unsigned)-1;
            do {
                cand++;
                v = (r * (cdf[cand] >> EC_PROB_SHIFT) >> (7 - EC_PROB_SHIFT)) + EC_MIN_PROB * ((unsigned)n_symbols - cand);
            } while (c < v);
        } while (el == EL_PARTITION && g_pol.is422 && (cand == PARTITION_V || cand == PARTITION_V4 || cand == PARTITION_T_LEFT_SPLIT || cand == PARTITION_T_RIGHT_SPLIT));
        if (cand < val) val = cand;
    }
    const unsigned u = val ? (r * (cdf[val - 1] >> EC_PROB_SHIFT) >> (7 - EC_PROB_SHIFT)) + EC_MIN_PROB * ((unsigned)n_symbols - (val - 1)) : s->rng;
    const unsigned v = (r * (cdf[val] >> EC_PROB_SHIFT) >> (7 - EC_PROB_SHIFT)) + EC_MIN_PROB * ((unsigned)n_symbols - val);
    Tile *const t = tile_of(s);
    enc_norm(&t->e, t->e.low + (s->rng - u), u - v);
    dec_norm(s, u - v);
    if (s->allow_update_cdf) {
        const unsigned count = cdf[n_symbols];
        const unsigned rate = 4 + (count >> 4) + (n_symbols > 2);
        unsigned i;
        for (i = 0; i < val; i++) cdf[i] += (32768 - cdf[i]) >> rate;
        for (; i < n_symbols; i++) cdf[i] -= cdf[i] >> rate;
        cdf[n_symbols] = count + (count < 32);
    }
    return val;
}

unsigned dav1d_msac_decode_hi_tok_c(MsacContext *const s, uint16_t *const cdf)
{
    unsigned tok_br = dav1d_msac_decode_symbol_adapt_c(s, cdf, 3), tok = 3 + tok_br;
    if (tok_br == 3) {
        tok_br = dav1d_msac_decode_symbol_adapt_c(s, cdf, 3); tok = 6 + tok_br;
        if (tok_br == 3) {
            tok_br = dav1d_msac_decode_symbol_adapt_c(s, cdf, 3); tok = 9 + tok_br;
            if (tok_br == 3) tok = 12 + dav1d_msac_decode_symbol_adapt_c(s, cdf, 3);
        }
    }
    return tok;
}

int dav1d_msac_decode_subexp(MsacContext *const s, const int ref, const int n, unsigned k)
{
    unsigned a = 0;
    if (dav1d_msac_decode_bool_equi(s)) {
        if (dav1d_msac_decode_bool_equi(s)) k += dav1d_msac_decode_bool_equi(s) + 1;
        a = 1 << k;
    }
    const unsigned v = dav1d_msac_decode_bools(s, k) + a;
    return ref * 2 <= n ? inv_recenter(ref, v) : n - 1 - inv_recenter(n - 1 - ref, v);
}
