/* tests/emu/cuda_emu.cpp — TEST-ONLY fiber scheduler behind cuda_emu.h (see its header). */
#include "cuda_emu.h"
#include <sys/mman.h>

EmuFiber *emu_cur = nullptr;
int emu_resident_blocks = 1;

extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

static void *sched_sp;
static const std::function<void()> *cur_body;
static const size_t kStack = 256 * 1024;

void emu_yield() {
    EmuFiber *f = emu_cur;
    emu_switch(&f->sp, sched_sp);
}

static void fiber_entry() {
    EmuFiber *f = emu_cur;
    (*cur_body)();
    f->done = true;
    f->blk->alive--;
    /* a thread that exits releases barriers the remaining threads are waiting on */
    EmuBlock *b = f->blk;
    if (b->alive > 0 && b->bar_arrived >= (unsigned)b->alive) { b->bar_arrived = 0; b->bar_gen++; }
    emu_switch(&f->sp, sched_sp);
    abort();
}

static std::vector<void *> stack_pool;
static void *get_stack() {
    if (!stack_pool.empty()) { void *s = stack_pool.back(); stack_pool.pop_back(); return s; }
    void *s = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (s == MAP_FAILED) { perror("emu stack mmap"); abort(); }
    return s;
}

void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    const long nblocks = (long)grid.x * grid.y * grid.z;
    const int resident = std::max(1, emu_resident_blocks);
    cur_body = &body;
    for (long b0 = 0; b0 < nblocks; b0 += resident) {
        const int nb = (int)std::min<long>(resident, nblocks - b0);
        std::vector<EmuBlock> blocks(nb);
        std::vector<EmuFiber> fibers((size_t)nb * nthreads);
        for (int bi = 0; bi < nb; bi++) {
            EmuBlock &B = blocks[bi];
            B.bdim = block; B.gdim = grid; B.nthreads = nthreads; B.alive = nthreads;
            B.bar_arrived = 0; B.bar_gen = 0;
            B.warps.assign((nthreads + 31) / 32, EmuWarp{});
            B.dyn_smem = smem ? aligned_alloc(1024, (smem + 1023) & ~(size_t)1023) : nullptr;
            const long lb = b0 + bi;
            uint3 bid = { (unsigned)(lb % grid.x), (unsigned)((lb / grid.x) % grid.y), (unsigned)(lb / ((long)grid.x * grid.y)) };
            for (int t = 0; t < nthreads; t++) {
                EmuFiber &f = fibers[(size_t)bi * nthreads + t];
                f.blk = &B; f.bid = bid; f.done = false; f.linear = t;
                f.tid = { (unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y)) };
                f.lane = t & 31; f.warp = t >> 5;
                f.stack = get_stack();
                uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
                uint64_t *sp = (uint64_t *)(top - 64);
                for (int i = 0; i < 6; i++) sp[i] = 0;
                sp[6] = (uint64_t)(uintptr_t)&fiber_entry;
                sp[7] = 0;
                f.sp = sp;
            }
        }
        size_t remaining = fibers.size();
        while (remaining) {
            size_t progressed = 0;
            for (EmuFiber &f : fibers) {
                if (f.done) continue;
                emu_cur = &f;
                emu_switch(&sched_sp, f.sp);
                if (f.done) { remaining--; stack_pool.push_back(f.stack); f.stack = nullptr; }
                progressed++;
            }
            if (!progressed) break;
        }
        emu_cur = nullptr;
        for (EmuBlock &B : blocks) free(B.dyn_smem);
    }
}

extern "C" void b200_emu_set_resident_blocks(int n) { emu_resident_blocks = n; }
