// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY. Fiber scheduler + guarded allocator
// behind cuda_emu.h (see the header for the model).
#include "cuda_emu.h"
#include <map>

namespace emu {

static State g_state;
State &S() { return g_state; }

static const size_t kStack = 256 * 1024;

#if RV_EMU_ASM_SWITCH
// emu_switch(from, to): saves the callee-saved registers of the running fiber on its stack and its stack pointer in *from,
// then resumes the fiber whose stack pointer is in *to (System V x86-64; no signal mask, no floating-point environment).
extern "C" void emu_switch(FiberCtx *from, FiberCtx *to);
asm(R"(
.text
.globl emu_switch
.hidden emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq (%rsi), %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");
static inline void switch_to(FiberCtx *from, FiberCtx *to) { emu_switch(from, to); }
// a fresh fiber: its first resumption "returns" into entry() with the stack aligned as after a call
static void prepare(FiberCtx &ctx, char *stack, size_t bytes, void (*entry)()) {
    uintptr_t top = ((uintptr_t)stack + bytes) & ~(uintptr_t)15;
    void **sp = (void **)top;
    *--sp = nullptr;          // the return address entry() would see (it never returns)
    *--sp = (void *)entry;    // popped by emu_switch's ret
    for (int i = 0; i < 6; i++) *--sp = nullptr;  // rbp rbx r12 r13 r14 r15
    ctx.sp = (void *)sp;
}
#else
static inline void switch_to(FiberCtx *from, FiberCtx *to) { swapcontext(from, to); }
static void prepare(FiberCtx &ctx, char *stack, size_t bytes, void (*entry)()) {
    getcontext(&ctx);
    ctx.uc_stack.ss_sp = stack;
    ctx.uc_stack.ss_size = bytes;
    ctx.uc_link = nullptr;
    makecontext(&ctx, entry, 0);
}
#endif

void die(const char *msg) {
    State &s = S();
    fprintf(stderr, "[cuda_emu] FATAL in kernel %s block %u thread %u: %s\n", s.kname, s.b_idx.x, s.t_idx.x, msg);
    abort();
}

void yield() {
    State &s = S();
    Fiber &f = s.fibers[s.cur];
    switch_to(&f.ctx, &s.sched);
}

static void fiber_entry() {
    State &s = S();
    (*s.body)();
    Fiber &f = s.fibers[s.cur];
    f.done = true;
    s.alive--;
    s.progress++;
    // a thread that exits releases a block barrier the rest is waiting on
    if (s.alive > 0 && s.bar_arrived >= s.alive) {
        s.bar_arrived = 0;
        s.bar_gen++;
    }
    switch_to(&f.ctx, &s.sched);
    abort();  // a finished fiber is never resumed
}

void launch(const char *name, dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    State &s = S();
    if (s.cur >= 0) die("nested launch");
    int nt = (int)(block.x * block.y * block.z);
    if (nt <= 0 || nt > 1024) { fprintf(stderr, "[cuda_emu] bad block size %d for %s\n", nt, name); abort(); }
    if (block.y != 1 || block.z != 1) { fprintf(stderr, "[cuda_emu] only 1-D blocks supported (%s)\n", name); abort(); }
    s.kname = name;
    s.body = &body;
    s.b_dim = block;
    s.g_dim = grid;
    s.nthreads = nt;
    if ((int)s.fibers.size() < nt) {
        size_t old = s.fibers.size();
        s.fibers.resize(nt);
        for (size_t i = old; i < (size_t)nt; i++) s.fibers[i].stack = (char *)malloc(kStack);
    }
    s.warps.assign((nt + 31) / 32, WarpState());
    if (smem > s.dyn_smem_cap) {
        free(s.dyn_smem);
        s.dyn_smem = (unsigned char *)aligned_alloc(128, (smem + 127) / 128 * 128);
        s.dyn_smem_cap = smem;
    }
    for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++) {
        s.b_idx = uint3{bx, by, bz};
        s.alive = nt;
        s.bar_arrived = 0;
        for (auto &w : s.warps) { w.arrived = 0; }
        for (int t = 0; t < nt; t++) {
            Fiber &f = s.fibers[t];
            prepare(f.ctx, f.stack, kStack, fiber_entry);
            f.done = false;
            f.tid = uint3{(unsigned)t, 0, 0};
        }
        unsigned long long last_progress = s.progress;
        int idle_rounds = 0;
        while (s.alive > 0) {
            for (int t = 0; t < nt; t++) {
                Fiber &f = s.fibers[t];
                if (f.done) continue;
                s.cur = t;
                s.t_idx = f.tid;
                switch_to(&s.sched, &f.ctx);
            }
            if (s.progress == last_progress) {
                if (++idle_rounds > 4) {
                    s.cur = 0;
                    die("deadlock: no thread of the block can make progress (barrier/collective mismatch or look-back on an unfinished block)");
                }
            } else {
                idle_rounds = 0;
                last_progress = s.progress;
            }
        }
        s.cur = -1;
    }
    s.body = nullptr;
    emu_check_guards();
}

}  // namespace emu

// ---- guarded device allocator ------------------------------------------------
static const size_t kGuard = 256;
static std::map<void *, size_t> &allocs() { static std::map<void *, size_t> m; return m; }

cudaError_t cudaMalloc(void **p, size_t bytes) {
    size_t padded = (bytes + 255) / 256 * 256;
    unsigned char *raw = (unsigned char *)aligned_alloc(256, padded + 2 * kGuard);
    if (!raw) return cudaErrorMemoryAllocation;
    memset(raw, 0xA5, kGuard);
    memset(raw + kGuard, 0xCD, padded);             // poison: uninitialised reads show up as garbage
    memset(raw + kGuard + bytes, 0xA5, padded - bytes + kGuard);
    *p = raw + kGuard;
    allocs()[*p] = bytes;
    return cudaSuccess;
}

static void check_one(void *p, size_t bytes) {
    unsigned char *u = (unsigned char *)p;
    size_t padded = (bytes + 255) / 256 * 256;
    for (size_t i = 0; i < kGuard; i++)
        if (u[-(long)i - 1] != 0xA5) { fprintf(stderr, "[cuda_emu] buffer UNDERRUN: alloc %p (%zu bytes) byte -%zu\n", p, bytes, i + 1); abort(); }
    for (size_t i = bytes; i < padded + kGuard; i++)
        if (u[i] != 0xA5) { fprintf(stderr, "[cuda_emu] buffer OVERRUN: alloc %p (%zu bytes) at offset %zu (kernel %s)\n", p, bytes, i, emu::S().kname); abort(); }
}

cudaError_t emu_check_guards() {
    for (auto &kv : allocs()) check_one(kv.first, kv.second);
    return cudaSuccess;
}

cudaError_t cudaFree(void *p) {
    if (!p) return cudaSuccess;
    auto it = allocs().find(p);
    if (it == allocs().end()) { fprintf(stderr, "[cuda_emu] cudaFree of unknown pointer %p\n", p); abort(); }
    check_one(p, it->second);
    free((unsigned char *)p - kGuard);
    allocs().erase(it);
    return cudaSuccess;
}
