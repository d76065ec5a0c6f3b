// A very small CUDA execution model on CPU threads, just enough to run the protocol kernels of src/core/kernels.cuh
// (triggers, movers, waits, cooperative agents, the wildcard resolver, pack/unpack) with REAL concurrency and no GPU:
//   * one OS thread per CTA, one ucontext fiber per CUDA thread; __syncthreads() is a fiber barrier, spin loops yield;
//   * "ranks" are host threads that launch kernels one after the other (= stream order); peer-mapped pointers are
//     ordinary pointers into the other rank's buffers;
//   * the .sys-scope PTX primitives are C++ atomics (include/acx_device.cuh, ACX_EMU section).
// x86 is more strongly ordered than the GPU memory model, so this checks protocol LOGIC (matching, generations, credits,
// chunk claims, completion accounting, deadlock freedom), not fences.
#pragma once
#define ACX_EMU 1
#include <cuda_runtime.h>   // host-side vector types (int4, uint4, dim3)
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <ucontext.h>
#include <functional>
#include <thread>
#include <vector>

#undef __shared__
#define __shared__ static thread_local      // one OS thread per CTA => per-CTA storage
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __global__
#define __global__
#undef __device__
#define __device__
#undef __host__
#define __host__
#undef __forceinline__
#define __forceinline__ inline

namespace emu {
struct Idx { unsigned x = 0, y = 0, z = 0; };
struct Fiber {
    ucontext_t ctx;
    Idx tid;
    int state = 0;   // 0 ready, 1 at barrier, 2 done
    std::vector<char> stack;
};
struct Cta {
    Idx bid, gdim, bdim;
    ucontext_t sched;
    Fiber *cur = nullptr;
    std::function<void()> body;
};
inline thread_local Cta *cta = nullptr;
inline uint64_t now_ns() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec; }
inline void yield() { Cta *c = cta; swapcontext(&c->cur->ctx, &c->sched); }
inline void barrier() { Cta *c = cta; c->cur->state = 1; swapcontext(&c->cur->ctx, &c->sched); }
inline void fiber_main() { Cta *c = cta; c->body(); c->cur->state = 2; swapcontext(&c->cur->ctx, &c->sched); }

inline void run_cta(unsigned bx, unsigned gx, unsigned threads, const std::function<void()> &body) {
    Cta c;
    c.bid.x = bx; c.gdim.x = gx; c.gdim.y = c.gdim.z = 1; c.bdim.x = threads; c.bdim.y = c.bdim.z = 1;
    c.body = body;
    cta = &c;
    std::vector<Fiber> f(threads);
    for (unsigned i = 0; i < threads; i++) {
        f[i].tid.x = i;
        f[i].stack.resize(256 << 10);
        getcontext(&f[i].ctx);
        f[i].ctx.uc_stack.ss_sp = f[i].stack.data();
        f[i].ctx.uc_stack.ss_size = f[i].stack.size();
        f[i].ctx.uc_link = &c.sched;
        makecontext(&f[i].ctx, (void (*)())fiber_main, 0);
    }
    for (;;) {
        unsigned done = 0, waiting = 0;
        for (unsigned i = 0; i < threads; i++) {
            if (f[i].state == 0) { c.cur = &f[i]; swapcontext(&c.sched, &f[i].ctx); }
            if (f[i].state == 2) done++; else if (f[i].state == 1) waiting++;
        }
        if (done == threads) break;
        if (waiting && waiting + done == threads) for (auto &x : f) if (x.state == 1) x.state = 0;   // barrier released (exited threads do not take part)
        else sched_yield();                                                                          // somebody is spinning: let the other CTAs / ranks run
    }
    cta = nullptr;
}

// <<<grid, block>>> in stream order: returns when the whole grid has finished (the caller is the stream).
template <typename F>
inline void launch(unsigned grid, unsigned block, F &&body) {
    std::function<void()> fn = body;
    if (grid == 1) { run_cta(0, 1, block, fn); return; }
    std::vector<std::thread> ts;
    for (unsigned b = 0; b < grid; b++) ts.emplace_back([=] { run_cta(b, grid, block, fn); });
    for (auto &t : ts) t.join();
}
}  // namespace emu

#define threadIdx (emu::cta->cur->tid)
#define blockIdx (emu::cta->bid)
#define blockDim (emu::cta->bdim)
#define gridDim (emu::cta->gdim)
inline void __syncthreads() { emu::barrier(); }
inline void __syncwarp() {}
inline void __nanosleep(unsigned) { emu::yield(); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
[[noreturn]] inline void __trap() { fprintf(stderr, "emu: __trap()\n"); abort(); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long val) {
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
    return cmp;
}
