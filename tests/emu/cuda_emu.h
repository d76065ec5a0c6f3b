// A small CUDA execution model on CPU threads: enough to run the library's device code (src/core/kernels.cuh,
// include/acx_device.cuh) and user kernels of the conformance programs with REAL concurrency and no GPU.
//   * one OS thread per CTA, one ucontext fiber per CUDA thread; __syncthreads() is a fiber barrier, spin loops yield;
//   * the .sys-scope PTX primitives are C++ atomics (include/acx_device.cuh, ACX_EMU section);
//   * __trap() unwinds every fiber of the process' "device" and leaves a sticky launch failure, like a real trap.
// Two users:
//   tests/emu/protocol_emu.cpp  calls kernels directly (ranks = threads of one process; runs under ThreadSanitizer);
//   tests/emu/cudart_emu.cpp    a software CUDA runtime (streams, events, graphs, capture, IPC memory) under which the
//                               whole of src/core/acx.cu runs unmodified, one process per rank (make emu-lib).
// x86 is more strongly ordered than the GPU memory model: this checks protocol and runtime LOGIC (matching, generations,
// credits, chunk claims, completion accounting, deadlock freedom, graph/capture bookkeeping). ThreadSanitizer on top
// checks the release/acquire pairing against the C++ memory model.
#pragma once
#define ACX_EMU 1
#include <cuda_runtime.h>   // host-side API declarations and vector types (int4, uint4, dim3)
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <ucontext.h>
#include <atomic>
#include <functional>
#include <thread>
#include <tuple>
#include <type_traits>
#include <utility>
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

// Sanitizer awareness: fibers switch stacks under the sanitizers' feet unless told.
#if defined(__SANITIZE_ADDRESS__)
#define EMU_ASAN 1
extern "C" void __sanitizer_start_switch_fiber(void **fake_stack_save, const void *bottom, size_t size);
extern "C" void __sanitizer_finish_switch_fiber(void *fake_stack_save, const void **bottom_old, size_t *size_old);
#endif
#if defined(__SANITIZE_THREAD__)
#define EMU_TSAN 1
extern "C" void *__tsan_get_current_fiber(void);
extern "C" void *__tsan_create_fiber(unsigned flags);
extern "C" void __tsan_destroy_fiber(void *fiber);
extern "C" void __tsan_switch_to_fiber(void *fiber, unsigned flags);
#endif

namespace emu {
struct Idx { unsigned x = 0, y = 0, z = 0; };
struct Trap {};   // thrown through a fiber by __trap() and by yields once the device is dead
struct Fiber {
    ucontext_t ctx;
    Idx tid;
    int state = 0;   // 0 ready, 1 at __syncthreads, 2 done, 3 at __syncwarp
    char *stack = nullptr;
    void *asan_fake = nullptr, *tsan = nullptr;
};
struct Cta {
    Idx bid, gdim, bdim;
    ucontext_t sched;
    Fiber *cur = nullptr;
    const std::function<void()> *body = nullptr;
    void *asan_fake = nullptr, *tsan = nullptr;        // the scheduler's (= the OS thread's) own context
    const void *asan_bottom = nullptr; size_t asan_size = 0;
};
inline thread_local Cta *cta = nullptr;
inline std::atomic<int> g_trapped{0};   // sticky: the "device" is dead (reported as cudaErrorLaunchFailure by the software runtime)
constexpr size_t kFiberStack = 128u << 10;

inline uint64_t now_ns() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec; }
inline void check_trap() { if (g_trapped.load(std::memory_order_relaxed)) throw Trap(); }
// fiber -> scheduler. `dying`: the fiber never runs again (its fake stack may be released).
inline void to_sched(Cta *c, bool dying) {
    Fiber *f = c->cur;
    (void)dying;
#ifdef EMU_ASAN
    __sanitizer_start_switch_fiber(dying ? nullptr : &f->asan_fake, c->asan_bottom, c->asan_size);
#endif
#ifdef EMU_TSAN
    __tsan_switch_to_fiber(c->tsan, 0);
#endif
    swapcontext(&f->ctx, &c->sched);
#ifdef EMU_ASAN
    __sanitizer_finish_switch_fiber(f->asan_fake, &c->asan_bottom, &c->asan_size);   // resumed
#endif
}
// scheduler -> fiber
inline void to_fiber(Cta *c, Fiber *f) {
    c->cur = f;
#ifdef EMU_ASAN
    __sanitizer_start_switch_fiber(&c->asan_fake, f->stack, kFiberStack);
#endif
#ifdef EMU_TSAN
    __tsan_switch_to_fiber(f->tsan, 0);
#endif
    swapcontext(&c->sched, &f->ctx);
#ifdef EMU_ASAN
    __sanitizer_finish_switch_fiber(c->asan_fake, nullptr, nullptr);
#endif
}
inline void yield() { Cta *c = cta; to_sched(c, false); check_trap(); }
inline void barrier() { Cta *c = cta; c->cur->state = 1; to_sched(c, false); check_trap(); }
inline void warp_barrier() { Cta *c = cta; if (c->bdim.x <= 1) return; c->cur->state = 3; to_sched(c, false); check_trap(); }
inline void fiber_main() {
    Cta *c = cta;
#ifdef EMU_ASAN
    __sanitizer_finish_switch_fiber(nullptr, &c->asan_bottom, &c->asan_size);   // first entry
#endif
    try { (*c->body)(); } catch (const Trap &) {}
    c = cta;
    c->cur->state = 2;
    to_sched(c, true);
}

// Fiber stacks are recycled per OS thread (mmap'ed, never zeroed by us).
struct StackPool {
    std::vector<char *> free_;
    ~StackPool() { for (char *s : free_) munmap(s, kFiberStack); }
    char *get() {
        if (!free_.empty()) { char *s = free_.back(); free_.pop_back(); return s; }
        void *p = mmap(nullptr, kFiberStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { perror("emu: mmap(fiber stack)"); abort(); }
        return (char *)p;
    }
    void put(char *s) { free_.push_back(s); }
};
inline thread_local StackPool stack_pool;

inline void run_cta(unsigned bx, unsigned gx, unsigned threads, const std::function<void()> &body) {
    Cta c;
    c.bid.x = bx; c.gdim.x = gx; c.gdim.y = c.gdim.z = 1; c.bdim.x = threads; c.bdim.y = c.bdim.z = 1;
    c.body = &body;
    Cta *outer = cta;
    cta = &c;
    std::vector<Fiber> f(threads);
    for (unsigned i = 0; i < threads; i++) {
        f[i].tid.x = i;
        f[i].stack = stack_pool.get();
        getcontext(&f[i].ctx);
        f[i].ctx.uc_stack.ss_sp = f[i].stack;
        f[i].ctx.uc_stack.ss_size = kFiberStack;
        f[i].ctx.uc_link = &c.sched;
        makecontext(&f[i].ctx, (void (*)())fiber_main, 0);
#ifdef EMU_TSAN
        f[i].tsan = __tsan_create_fiber(0);
#endif
    }
#ifdef EMU_TSAN
    c.tsan = __tsan_get_current_fiber();
#endif
    unsigned idle = 0, last_done = 0;
    for (;;) {
        unsigned done = 0, waiting = 0;
        for (unsigned i = 0; i < threads; i++) {
            if (f[i].state == 0) to_fiber(&c, &f[i]);
            if (f[i].state == 2) done++; else if (f[i].state == 1) waiting++;
        }
        if (done == threads) break;
        if (waiting && waiting + done == threads) { for (auto &x : f) if (x.state == 1) x.state = 0; idle = 0; continue; }   // barrier released (exited threads do not take part)
        {   // __syncwarp: a warp's lanes go on once every live lane of that warp has arrived
            bool released = false;
            for (unsigned w0 = 0; w0 < threads; w0 += 32) {
                const unsigned w1 = w0 + 32 < threads ? w0 + 32 : threads;
                unsigned at = 0, live = 0;
                for (unsigned i = w0; i < w1; i++) { if (f[i].state == 3) at++; if (f[i].state != 2) live++; }
                if (at && at == live) { for (unsigned i = w0; i < w1; i++) if (f[i].state == 3) f[i].state = 0; released = true; }
            }
            if (released) { idle = 0; continue; }
        }
        // somebody is spinning on another CTA / rank / stream: give the CPU away — politely at first, then for real
        // (a box with fewer cores than emulated threads must not burn them all in sched_yield)
        if (done != last_done) { last_done = done; idle = 0; }
        if (++idle < 64) sched_yield();
        else { timespec ts{0, idle < 1024 ? 2000 : 50000}; nanosleep(&ts, nullptr); }
    }
    for (auto &x : f) {
        stack_pool.put(x.stack);
#ifdef EMU_TSAN
        __tsan_destroy_fiber(x.tsan);
#endif
    }
    cta = outer;
}

// <<<grid, block>>> in stream order: returns when the whole grid has finished (the caller is the stream).
template <typename F>
inline void launch(unsigned grid, unsigned block, F &&body) {
    const std::function<void()> fn = body;
    if (grid == 1) { run_cta(0, 1, block, fn); return; }
    std::vector<std::thread> ts;
    for (unsigned b = 0; b < grid; b++) ts.emplace_back([&fn, b, grid, block] { run_cta(b, grid, block, fn); });
    for (auto &t : ts) t.join();
}

// ---- kernel registry: what cudaLaunchKernel / cudaGraphAddKernelNode of the software runtime need to know about a
// kernel (a plain function here): how big its arguments are and how to call it from an array of argument pointers ----
struct KernelInfo {
    const char *name = "?";
    std::vector<size_t> arg_sizes;
    std::function<void(void **)> invoke;
    bool scalable = false;   // written for any blockDim: the runtime may shrink the block to keep fiber counts low
};
void register_kernel(const void *fn, const KernelInfo &ki);   // cudart_emu.cpp
template <typename... A, size_t... I>
inline void invoke_kernel(void (*k)(A...), void **args, std::index_sequence<I...>) { k(*(std::remove_reference_t<A> *)args[I]...); }
template <typename... A>
inline void register_kernel_fn(void (*k)(A...), const char *name, bool scalable) {
    KernelInfo ki;
    ki.name = name; ki.scalable = scalable;
    ki.arg_sizes = {sizeof(std::remove_reference_t<A>)...};
    ki.invoke = [k](void **args) { invoke_kernel(k, args, std::index_sequence_for<A...>{}); };
    register_kernel((const void *)k, ki);
}
struct KernelRegistrar { template <typename... A> KernelRegistrar(void (*k)(A...), const char *name, bool scalable) { register_kernel_fn(k, name, scalable); } };
// kern<<<grid, block, 0, stream>>>(args...) for user kernels of test programs
template <typename... A, typename... B>
inline cudaError_t launch_user(void (*k)(A...), dim3 grid, dim3 block, cudaStream_t stream, B... args) {   // dim3 converts from unsigned
    register_kernel_fn(k, "user kernel", false);   // idempotent
    std::tuple<std::remove_reference_t<A>...> vals(args...);
    void *ptrs[sizeof...(A) ? sizeof...(A) : 1];
    size_t i = 0;
    std::apply([&](auto &...v) { ((ptrs[i++] = (void *)&v), ...); }, vals);
    return cudaLaunchKernel((const void *)k, grid, block, ptrs, 0, stream);
}
}  // namespace emu
#define ACX_EMU_KERNEL(k, scalable) static const emu::KernelRegistrar emu_registrar_##k(k, #k, scalable);

// threadIdx / blockIdx / blockDim / gridDim: objects (not macros: CUDA's host structs have members called gridDim and
// blockDim) whose .x/.y/.z read the running fiber's coordinates.
namespace emu {
#define EMU_COORD(NAME, EXPR) struct NAME { operator unsigned() const { return (EXPR); } }
EMU_COORD(TidX, cta->cur->tid.x); EMU_COORD(TidY, cta->cur->tid.y); EMU_COORD(TidZ, cta->cur->tid.z);
EMU_COORD(BidX, cta->bid.x);      EMU_COORD(BidY, cta->bid.y);      EMU_COORD(BidZ, cta->bid.z);
EMU_COORD(BdimX, cta->bdim.x);    EMU_COORD(BdimY, cta->bdim.y);    EMU_COORD(BdimZ, cta->bdim.z);
EMU_COORD(GdimX, cta->gdim.x);    EMU_COORD(GdimY, cta->gdim.y);    EMU_COORD(GdimZ, cta->gdim.z);
#undef EMU_COORD
struct Tid { TidX x; TidY y; TidZ z; };
struct Bid { BidX x; BidY y; BidZ z; };
struct Bdim { BdimX x; BdimY y; BdimZ z; };
struct Gdim { GdimX x; GdimY y; GdimZ z; };
}  // namespace emu
static const emu::Tid threadIdx{};
static const emu::Bid blockIdx{};
static const emu::Bdim blockDim{};
static const emu::Gdim gridDim{};
inline void __syncthreads() { emu::barrier(); }
inline void __syncwarp() { emu::warp_barrier(); }
inline void __nanosleep(unsigned) { emu::yield(); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
[[noreturn]] inline void __trap() {
    if (!emu::g_trapped.exchange(1)) fprintf(stderr, "emu: __trap() in block %u thread %u\n", (unsigned)blockIdx.x, (unsigned)threadIdx.x);
    throw emu::Trap();
}
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long val) {
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
    return cmp;
}
inline int atomicCAS(int *p, int cmp, int val) {
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
    return cmp;
}
