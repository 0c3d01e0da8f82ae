// The scheduler of the CUDA-on-CPU emulator (tests/host/cuda_emu.h): the threads of a CTA as fibers of one host thread.  Test infrastructure only.
#include "../cuda_emu.h"

#include <cstdio>
#include <cstdlib>

#if !defined(__x86_64__)
#error "the emulator's context switch is written for x86-64"
#endif

// void fz_emu_switch(void** save_sp, void* load_sp): save the callee-saved registers and the stack pointer of the running context, continue the other one
extern "C" void fz_emu_switch(void** save_sp, void* load_sp);
asm(R"(
    .text
    .globl fz_emu_switch
    .type fz_emu_switch, @function
fz_emu_switch:
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
    .size fz_emu_switch, .-fz_emu_switch
)");

namespace fz_emu {

Cta g_cta;
Fiber* g_cur = nullptr;
Idx g_blockIdx, g_blockDim, g_gridDim;

namespace {
constexpr size_t STACK_BYTES = 256 * 1024;
constexpr int MAX_THREADS = 1024;
Fiber g_fibers[MAX_THREADS];
char* g_stacks = nullptr;
void* g_sched_sp = nullptr;
const std::function<void()>* g_body = nullptr;

void yield() { fz_emu_switch(&g_cur->sp, g_sched_sp); }

void drop(Bar& b) {      // a thread that has left the kernel: whoever is waiting must not wait for it
    --b.expected;
    if (b.expected > 0 && b.arrived == b.expected) { b.arrived = 0; ++b.gen; }
}

void fiber_main() {
    (*g_body)();
    Fiber* me = g_cur;
    me->done = true;
    drop(g_cta.warp_bar[me->tid.x >> 5]);      // warps are formed along x (the product's kernels are one-dimensional CTAs)
    drop(g_cta.block_bar);
    fz_emu_switch(&me->sp, g_sched_sp);
    std::abort();                               // a finished fiber is never resumed
}
}  // namespace

void bar_wait(Bar& b) {
    const unsigned g = b.gen;
    if (++b.arrived == b.expected) { b.arrived = 0; ++b.gen; return; }
    while (b.gen == g) yield();
}

void run_cta(int n, const std::function<void()>& body) {
    if (n > MAX_THREADS) { std::fprintf(stderr, "cuda_emu: CTA of %d threads\n", n); std::abort(); }
    if (!g_stacks) g_stacks = static_cast<char*>(std::malloc(STACK_BYTES * MAX_THREADS));
    g_body = &body;
    g_cta.n_threads = n;
    g_cta.block_bar = Bar();
    g_cta.block_bar.expected = n;
    for (int w = 0; w < 32; ++w) { g_cta.warp_bar[w] = Bar(); g_cta.warp_bar[w].expected = std::max(0, std::min(32, n - 32 * w)); }
    for (int t = 0; t < n; ++t) {
        Fiber& f = g_fibers[t];
        f.done = false;
        f.tid.x = (unsigned)t % g_blockDim.x;
        f.tid.y = ((unsigned)t / g_blockDim.x) % g_blockDim.y;
        f.tid.z = (unsigned)t / (g_blockDim.x * g_blockDim.y);
        // initial frame: what fz_emu_switch pops (six registers, then the entry point as its return address); the stack pointer is 8 modulo 16 when
        // fiber_main starts, as after a call
        uintptr_t top = reinterpret_cast<uintptr_t>(g_stacks + STACK_BYTES * (size_t)(t + 1)) & ~(uintptr_t)15;
        void** sp = reinterpret_cast<void**>(top) - 8;
        for (int k = 0; k < 6; ++k) sp[k] = nullptr;
        sp[6] = reinterpret_cast<void*>(&fiber_main);
        sp[7] = nullptr;
        f.sp = sp;
    }
    int live = n;
    while (live > 0) {
        const unsigned block_gen = g_cta.block_bar.gen;
        unsigned warp_gens = 0;
        for (int w = 0; w < 32; ++w) warp_gens += g_cta.warp_bar[w].gen;
        const int before = live;
        for (int t = 0; t < n; ++t) {
            Fiber& f = g_fibers[t];
            if (f.done) continue;
            g_cur = &f;
            fz_emu_switch(&g_sched_sp, f.sp);
            if (f.done) --live;
        }
        unsigned warp_gens2 = 0;
        for (int w = 0; w < 32; ++w) warp_gens2 += g_cta.warp_bar[w].gen;
        if (live == before && live > 0 && block_gen == g_cta.block_bar.gen && warp_gens == warp_gens2) {
            // nobody finished and no barrier opened in a whole round: every live thread waits for threads that will never arrive
            bool any_arrival_pending = false;
            for (int w = 0; w < 32; ++w) any_arrival_pending |= g_cta.warp_bar[w].arrived > 0;
            any_arrival_pending |= g_cta.block_bar.arrived > 0;
            if (any_arrival_pending) { std::fprintf(stderr, "cuda_emu: deadlock at a barrier (divergent __syncthreads / __syncwarp?)\n"); std::abort(); }
        }
    }
    g_cur = nullptr;
}

}  // namespace fz_emu
