// cuda_emu.cpp — fiber scheduler for the TEST-ONLY CUDA emulation (see cuda_emu.h).
#include "cuda_emu.h"

namespace emu {

Block* g_blk = nullptr;
uint3 g_threadIdx = {0, 0, 0}, g_blockIdx = {0, 0, 0};
dim3 g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;

static const size_t kStack = 512 * 1024;

static void trampoline() {
    Block* b = g_blk;
    int me = b->cur;
    set_tid(me);
    b->entry(b->entry_arg);
    // thread exit: leave barriers consistent for the survivors
    b->fib[me].done = true;
    b->alive--;
    int w = me / 32;
    b->walive[w]--;
    if (b->alive > 0 && b->bar_count >= b->alive) { b->bar_count = 0; b->bar_gen++; }
    if (b->walive[w] > 0 && b->wbar_count[w] >= b->walive[w]) { b->wbar_count[w] = 0; b->wbar_gen[w]++; }
    swapcontext(&b->fib[me].ctx, &b->sched);
}

void run_block(void (*entry)(void*), void* arg, int nthreads) {
    static Block blk;  // stacks are reused across blocks
    Block* b = &blk;
    g_blk = b;
    b->entry = entry;
    b->entry_arg = arg;
    b->nthreads = nthreads;
    b->alive = nthreads;
    b->bar_count = 0;
    int nw = (nthreads + 31) / 32;
    b->wbar_count.assign(nw, 0);
    b->wbar_gen.assign(nw, 0);
    b->walive.assign(nw, 0);
    for (int t = 0; t < nthreads; t++) b->walive[t / 32]++;
    b->xchg.assign((size_t)nw * 32, 0);
    b->wscr.assign((size_t)nw * 32 * 8, 0);
    if ((int)b->fib.size() < nthreads) b->fib.resize(nthreads);
    for (int t = 0; t < nthreads; t++) {
        Fiber& f = b->fib[t];
        if (f.stack.size() != kStack) f.stack.resize(kStack);
        f.done = false;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.data();
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    while (b->alive > 0) {
        bool progressed = false;
        for (int t = 0; t < nthreads; t++) {
            if (b->fib[t].done) continue;
            b->cur = t;
            set_tid(t);
            swapcontext(&b->sched, &b->fib[t].ctx);
            progressed = true;
        }
        if (!progressed) break;
    }
    g_blk = nullptr;
}

}  // namespace emu
