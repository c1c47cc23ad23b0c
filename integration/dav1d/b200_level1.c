/*
 * integration/dav1d/b200_level1.c — the Level-1 drop-in inside a real dav1d: the arch hook INTEGRATION.md describes.
 *
 * dav1d fills c->dsp[] the first time a bit depth is seen (reference src/decode.c:3387-3416): each
 * dav1d_<family>_dsp_init_{8,16}bpc installs the C functions and then lets an architecture hook override them
 * (src/itx_tmpl.c:290-307 and the same pattern in the other six families). In libdav1d_b200_l1.so decode.c is compiled
 * with those seven calls renamed to the functions below, which run dav1d's own init (so every slot holds the C function)
 * and then let libb200av1's b200_<family>_dsp_init_{8,16}bpc override the slots it implements — the table structs are
 * layout-identical (include/b200av1.h). dav1d's reconstruction code (recon_tmpl.c, lf_apply, cdef_apply, lr_apply,
 * fg_apply) then runs UNCHANGED and every DSP call lands in a CUDA kernel: record -> launch -> sync per call. That is the
 * semantic definition of each kernel and a parity harness, not a throughput path (Level 2 is: integration/dav1d/b200_hooks*).
 * No CPU fallback: without a back end the tables are not usable and the first frame aborts.
 */
#include "config.h"
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "src/internal.h"

#define API __attribute__((visibility("default")))

static void *g_handle;
static unsigned g_families = 0x7f;      /* bit per family: 1 itx, 2 mc, 4 ipred, 8 loopfilter, 16 cdef, 32 looprestoration, 64 filmgrain */

API int b200l1_set_backend(const char *path, unsigned families)
{
    g_handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!g_handle) { fprintf(stderr, "b200l1: cannot load back end %s: %s\n", path, dlerror()); return -1; }
    g_families = families;
    return 0;
}

static void *sym(const char *name)
{
    if (!g_handle) {
        const char *env = getenv("B200AV1_LIB");
        if (!env || b200l1_set_backend(env, g_families)) {
            fprintf(stderr, "b200l1: no back end loaded (b200l1_set_backend / B200AV1_LIB)\n");
            abort();
        }
    }
    void *const p = dlsym(g_handle, name);
    if (!p) { fprintf(stderr, "b200l1: back end lacks %s\n", name); abort(); }
    return p;
}

/* Every member of the seven DSP contexts is a function pointer, so a context is an array of slots. After dav1d's own init
 * a slot is either NULL (a combination AV1 does not define) or one of dav1d's C functions; after the back end's init every
 * non-NULL slot must point somewhere else. A slot the back end forgot would silently keep running on the CPU and the stream
 * tests would still pass: count them (b200l1_c_slots_left) and name them on stderr. */
static int g_c_slots_left, g_slots_replaced;
API int b200l1_c_slots_left(void) { return g_c_slots_left; }
API int b200l1_slots_replaced(void) { return g_slots_replaced; }

static void check_replaced(const char *family, int bd, void *const *before, void *const *after, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        if (!before[i]) continue;
        if (after[i] == before[i]) {
            g_c_slots_left++;
            fprintf(stderr, "b200l1: %s_%dbpc slot %zu of %zu is still dav1d's C function\n", family, bd, i, n);
        } else g_slots_replaced++;
    }
}

#define HOOK_BODY(bit, family, Type, bd, CALL_REF, CALL_B200) \
        CALL_REF; \
        if (g_families & (bit)) { \
            void *before[sizeof(Type) / sizeof(void *)]; \
            _Static_assert(sizeof(Type) % sizeof(void *) == 0, "context is an array of function pointers"); \
            memcpy(before, c, sizeof(Type)); \
            CALL_B200; \
            check_replaced(#family, bd, before, (void *const *)c, sizeof(Type) / sizeof(void *)); \
        }
#define HOOK0(bit, family, Type, bd) \
    void b200l1_##family##_dsp_init_##bd##bpc(Type *const c) { \
        HOOK_BODY(bit, family, Type, bd, dav1d_##family##_dsp_init_##bd##bpc(c), \
                  ((void (*)(void *))sym("b200_" #family "_dsp_init_" #bd "bpc"))(c)) \
    }
#define HOOK1(bit, family, Type, bd) \
    void b200l1_##family##_dsp_init_##bd##bpc(Type *const c, const int bpc) { \
        HOOK_BODY(bit, family, Type, bd, dav1d_##family##_dsp_init_##bd##bpc(c, bpc), \
                  ((void (*)(void *, int))sym("b200_" #family "_dsp_init_" #bd "bpc"))(c, bpc)) \
    }

HOOK1(1, itx, Dav1dInvTxfmDSPContext, 8)
HOOK1(1, itx, Dav1dInvTxfmDSPContext, 16)
HOOK0(2, mc, Dav1dMCDSPContext, 8)
HOOK0(2, mc, Dav1dMCDSPContext, 16)
HOOK0(4, intra_pred, Dav1dIntraPredDSPContext, 8)
HOOK0(4, intra_pred, Dav1dIntraPredDSPContext, 16)
HOOK0(8, loop_filter, Dav1dLoopFilterDSPContext, 8)
HOOK0(8, loop_filter, Dav1dLoopFilterDSPContext, 16)
HOOK0(16, cdef, Dav1dCdefDSPContext, 8)
HOOK0(16, cdef, Dav1dCdefDSPContext, 16)
HOOK1(32, loop_restoration, Dav1dLoopRestorationDSPContext, 8)
HOOK1(32, loop_restoration, Dav1dLoopRestorationDSPContext, 16)
HOOK0(64, film_grain, Dav1dFilmGrainDSPContext, 8)
HOOK0(64, film_grain, Dav1dFilmGrainDSPContext, 16)
