// Software CUDA runtime for CI without a GPU: the subset of the CUDA runtime API that src/core/acx.cu, the bundled
// mini-MPI and the conformance programs use, implemented on CPU threads. Linked (instead of libcudart) into
// build/emu/libacx_emu.so together with the UNMODIFIED runtime (src/core/acx.cu compiled as C++ with tests/emu/cuda_emu.h),
// so the whole device transport — windows, CUDA-IPC bootstrap, streams, events, stream capture, explicit graphs, user
// objects, kernels — runs one process per rank under acxrun exactly as it does on a B200, minus the hardware.
//
//   device memory   cudaMalloc = a memfd mapping; cudaIpcGetMemHandle names it {pid, fd}; peers map /proc/<pid>/fd/<fd>
//   streams         one worker thread + FIFO each; legacy default-stream synchronization with blocking streams
//   events          generation counters; cudaStreamWaitEvent blocks the waiting stream's worker
//   kernels         registered plain functions (ACX_EMU_KERNEL / emu::launch_user), one OS thread per CTA, one fiber
//                   per CUDA thread; library kernels are blockDim-agnostic and run with small blocks (ACX_EMU_MAX_BLOCK)
//   graphs          explicit construction, stream capture incl. cross-stream fork/join through events (with the
//                   "unjoined work" check CUDA applies at cudaStreamEndCapture), child graphs, user objects with
//                   reference counts, instantiate = clone, launch = dataflow execution (independent branches run concurrently)
//   errors          a device __trap() is sticky: every later synchronizing call returns cudaErrorLaunchFailure
#include "cuda_emu.h"

#include <cuda.h>   // driver types of the stream memOps (CUstreamBatchMemOpParams)
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <sys/syscall.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <unordered_map>

namespace {

thread_local cudaError_t t_last = cudaSuccess;
thread_local int t_dev = 0;   // CUDA's current device is per thread; a new thread starts on device 0
cudaError_t fail(cudaError_t e) { t_last = e; return e; }
cudaError_t dead_or(cudaError_t ok) { return emu::g_trapped.load() ? fail(cudaErrorLaunchFailure) : ok; }

long env_long(const char *n, long d) { const char *v = getenv(n); return v && *v ? strtol(v, nullptr, 0) : d; }

// ------------------------------------------------------------------------------------------------------------------
// memory
// ------------------------------------------------------------------------------------------------------------------
enum { kDevOwn = 0, kDevIpc = 1, kPinned = 2, kRegistered = 3 };
struct Alloc { char *base; size_t size; uint64_t id; int kind; int fd; int dev = 0; };
std::mutex g_mem_mu;
std::map<uintptr_t, Alloc> g_allocs;   // by base address
uint64_t g_next_id = 1;

const Alloc *find_alloc(const void *p) {   // g_mem_mu held
    auto it = g_allocs.upper_bound((uintptr_t)p);
    if (it == g_allocs.begin()) return nullptr;
    --it;
    const Alloc &a = it->second;
    return ((uintptr_t)p < (uintptr_t)a.base + a.size) ? &a : nullptr;
}

struct IpcHandle { uint32_t magic; int32_t pid, fd; uint32_t dev; uint64_t size, id; };
static_assert(sizeof(IpcHandle) <= sizeof(cudaIpcMemHandle_t), "handle must fit");
constexpr uint32_t kIpcMagic = 0xACE0E001u;

// ------------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------------
std::mutex &reg_mu() { static std::mutex m; return m; }
std::unordered_map<const void *, emu::KernelInfo> &registry() { static std::unordered_map<const void *, emu::KernelInfo> r; return r; }

struct KernelLaunch {
    const emu::KernelInfo *ki = nullptr;
    unsigned grid = 1, block = 1;
    std::vector<std::vector<char>> args;
    void run() const {
        std::vector<void *> ptrs(args.size());
        for (size_t i = 0; i < args.size(); i++) ptrs[i] = (void *)args[i].data();
        void **pp = ptrs.data();
        const emu::KernelInfo *k = ki;
        emu::launch(grid, block, [k, pp] { k->invoke(pp); });
    }
};
cudaError_t make_launch(const void *func, dim3 grid, dim3 block, void **args, KernelLaunch *out) {
    const emu::KernelInfo *ki = nullptr;
    {
        std::lock_guard<std::mutex> lk(reg_mu());
        auto it = registry().find(func);
        if (it != registry().end()) ki = &it->second;   // node addresses of an unordered_map are stable
    }
    if (!ki) { fprintf(stderr, "emu: launch of an unregistered kernel %p\n", func); return cudaErrorInvalidDeviceFunction; }
    if (grid.y != 1 || grid.z != 1 || block.y != 1 || block.z != 1) { fprintf(stderr, "emu: only 1-D launches are emulated (%s)\n", ki->name); return cudaErrorInvalidConfiguration; }
    if (grid.x == 0 || block.x == 0 || block.x > 1024) return cudaErrorInvalidConfiguration;
    static const unsigned max_block = (unsigned)std::max(1L, env_long("ACX_EMU_MAX_BLOCK", 8));
    out->ki = ki; out->grid = grid.x; out->block = ki->scalable ? std::min(block.x, max_block) : block.x;
    out->args.resize(ki->arg_sizes.size());
    for (size_t i = 0; i < ki->arg_sizes.size(); i++) out->args[i].assign((const char *)args[i], (const char *)args[i] + ki->arg_sizes[i]);
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------------------------
// graphs
// ------------------------------------------------------------------------------------------------------------------
struct UserObj { void *ptr; cudaHostFn_t destroy; std::atomic<long> refs; };
// The real runtime may run a user object's destructor on an internal thread some time AFTER the call that dropped the last
// reference has returned (cudaUserObjectNoDestructorSync; observed on hardware in round 1: slots of a destroyed graph were
// still held when cudaGraphDestroy returned). Modelled here: a detached thread, ACX_EMU_UOBJ_DELAY_MS (default 15) later.
static std::atomic<int> g_uobj_async{0};
struct UobjDrain { ~UobjDrain() { for (int i = 0; i < 3000 && g_uobj_async.load() > 0; i++) usleep(1000); } };
static UobjDrain g_uobj_drain;
void uobj_release(UserObj *o, long n) {
    if (o->refs.fetch_sub(n) - n > 0) return;
    static const long delay_ms = getenv("ACX_EMU_UOBJ_DELAY_MS") ? atol(getenv("ACX_EMU_UOBJ_DELAY_MS")) : 15;
    if (delay_ms <= 0) { o->destroy(o->ptr); delete o; return; }
    g_uobj_async.fetch_add(1);
    std::thread([o] { usleep((useconds_t)delay_ms * 1000); o->destroy(o->ptr); delete o; g_uobj_async.fetch_sub(1); }).detach();
}
struct Graph;
enum NodeType { kNodeKernel, kNodeEmpty, kNodeChild, kNodeMemcpy, kNodeMemset, kNodeHost };
struct Node {
    NodeType type = kNodeEmpty;
    Graph *owner = nullptr;
    std::vector<Node *> deps;
    KernelLaunch kl;
    Graph *child = nullptr;
    void *dst = nullptr; const void *src = nullptr; size_t bytes = 0; int value = 0;
    cudaHostFn_t fn = nullptr; void *fn_arg = nullptr;
};
struct Graph {
    std::vector<Node *> nodes;                       // insertion order is a topological order (deps always exist first)
    std::vector<std::pair<UserObj *, long>> uobjs;
    bool is_exec = false;
    int inflight = 0;      // launches queued or running (g_graph_mu)
    bool zombie = false;   // cudaGraphExecDestroy arrived while launches were in flight: the last one deletes
    ~Graph();
};
Graph::~Graph() {
    for (Node *n : nodes) { delete n->child; delete n; }
    for (auto &u : uobjs) uobj_release(u.first, u.second);
}
Graph *clone_graph(const Graph *g) {
    Graph *c = new Graph();
    std::unordered_map<const Node *, Node *> m;
    for (const Node *n : g->nodes) {
        Node *x = new Node(*n);
        x->owner = c;
        x->child = n->child ? clone_graph(n->child) : nullptr;
        for (auto &d : x->deps) d = m.at(d);
        m[n] = x;
        c->nodes.push_back(x);
    }
    for (auto &u : g->uobjs) { u.first->refs.fetch_add(u.second); c->uobjs.push_back(u); }
    return c;
}
void jitter();
void run_graph(const Graph *g);
void run_node(const Node *n) {
    if (emu::g_trapped.load()) return;
    switch (n->type) {
        case kNodeKernel: n->kl.run(); break;
        case kNodeChild: run_graph(n->child); break;
        case kNodeMemcpy: memmove(n->dst, n->src, n->bytes); break;
        case kNodeMemset: memset(n->dst, n->value, n->bytes); break;
        case kNodeHost: n->fn(n->fn_arg); break;
        case kNodeEmpty: break;
    }
}
// Dataflow execution, like the hardware: a node starts as soon as its own dependencies have finished, so independent
// branches of a graph (a captured side-stream mover next to the main chain, the send / receive children of an explicit
// graph) really run concurrently — and ThreadSanitizer checks that every ordering the library relies on is an edge.
// A linear chain (the common case) runs inline; ACX_EMU_GRAPH_SERIAL=1 forces insertion order for debugging.
void run_graph(const Graph *g) {
    const size_t n = g->nodes.size();
    bool chain = true;
    for (size_t i = 0; i < n && chain; i++) {
        const Node *x = g->nodes[i];
        if (i == 0 ? !x->deps.empty() : !(x->deps.size() == 1 && x->deps[0] == g->nodes[i - 1])) chain = false;
    }
    static const bool serial = env_long("ACX_EMU_GRAPH_SERIAL", 0) != 0;
    if (chain || serial || n < 2) { for (const Node *x : g->nodes) run_node(x); return; }
    std::unordered_map<const Node *, size_t> index;
    for (size_t i = 0; i < n; i++) index[g->nodes[i]] = i;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<char> done(n, 0);
    std::vector<std::thread> workers;
    workers.reserve(n);
    for (size_t i = 0; i < n; i++) {
        workers.emplace_back([&, i] {
            const Node *x = g->nodes[i];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { for (const Node *d : x->deps) if (!done[index.at(d)]) return false; return true; });
            }
            jitter();
            run_node(x);
            { std::lock_guard<std::mutex> lk(mu); done[i] = 1; }
            cv.notify_all();
        });
    }
    for (auto &w : workers) w.join();
}
std::mutex g_graph_mu;                 // graph construction is host-side and rare: one lock
std::set<Graph *> g_graphs;            // live handles (validation)

// ------------------------------------------------------------------------------------------------------------------
// streams, events, capture
// ------------------------------------------------------------------------------------------------------------------
struct Stream;
struct Capture {
    Graph *graph;
    Stream *origin;
    std::set<Stream *> members;
    unsigned long long id;
    enum cudaStreamCaptureMode mode = cudaStreamCaptureModeGlobal;
    bool invalid = false;   // an unsafe call was made while it was in progress: cudaStreamEndCapture reports it
};
// ACX_EMU_JITTER_US=n: every stream operation and graph node starts after a pseudo-random 0..n us — the hardware gives no
// promise about when a runnable kernel starts, and late starters are what slot- and buffer-reuse races need.
void jitter() {
    static const long max_us = env_long("ACX_EMU_JITTER_US", 0);
    if (max_us <= 0) return;
    static std::atomic<unsigned> seed{(unsigned)getpid() * 2654435761u};
    unsigned x = seed.fetch_add(0x9e3779b9u);
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13;
    if (x & 3) return;   // three out of four start at once: bursts of order changes, not uniform slowness
    timespec ts{0, (long)(x % (unsigned)max_us) * 1000};
    nanosleep(&ts, nullptr);
}
// ACX_EMU_HW_QUEUES=n: the hardware's work queues. A GPU takes work from a few queues (CUDA_DEVICE_MAX_CONNECTIONS, 8 by
// default) onto which the driver maps all streams of a process, and a queue is strictly in order: an entry is handed to the
// device only when the entries in front of it have been, and an entry whose own stream predecessor has not finished yet (a
// kernel behind a spinning kernel of the same stream, an event or memOp wait that is not satisfied) holds up everything queued
// behind it — operations of OTHER streams included. That false dependency turns "two threads, two streams, each waiting for a
// message" into a deadlock across ranks when the streams share a queue (tests/c/threads.c on hardware).
// Streams are assigned round-robin in creation order; 0 (default) = every stream has its own queue.
struct HwQueue { std::mutex mu; std::condition_variable cv; uint64_t tail = 0, started = 0; };
std::vector<HwQueue *> &hw_queues() {   // never destroyed: stream workers may outlive static destructors, and the queues stay reachable for LeakSanitizer
    static std::vector<HwQueue *> *qs = [] { auto *v = new std::vector<HwQueue *>(); const long n = env_long("ACX_EMU_HW_QUEUES", 0); for (long i = 0; i < n; i++) v->push_back(new HwQueue()); return v; }();
    return *qs;
}
bool holds_hw_queue(const char *what) { return !strcmp(what, "event-wait") || !strcmp(what, "batch-memop"); }   // semaphore acquires: the queue stalls AT them until they are satisfied

struct Stream {
    static constexpr size_t kMaxQueued = 192;
    HwQueue *hwq = nullptr;
    std::deque<uint64_t> hw_seq;   // queue position of each pending operation (parallel to q)
    unsigned flags = 0;
    bool is_null = false;
    int dev = 0;            // the device that was current when the stream was created
    std::thread th;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::deque<std::pair<const char *, std::function<void()>>> q;
    const char *running = nullptr;   // label of the task being executed (debug dump)
    uint64_t submitted = 0, completed = 0;
    bool stop = false;
    // capture state (guarded by g_graph_mu)
    Capture *cap = nullptr;
    std::vector<Node *> cap_deps;

    void worker() {
        for (;;) {
            std::function<void()> t;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || !q.empty(); });
                if (q.empty()) return;
                running = q.front().first;
                t = std::move(q.front().second);
                q.pop_front();
            }
            uint64_t hw_pos = 0;
            bool hw_hold = false;
            if (hwq) {
                { std::lock_guard<std::mutex> lk(mu); hw_pos = hw_seq.front(); hw_seq.pop_front(); }
                hw_hold = holds_hw_queue(running);
                std::unique_lock<std::mutex> lk(hwq->mu);
                hwq->cv.wait(lk, [&] { return hwq->started + 1 == hw_pos; });   // everything in front of it in the queue has been handed to the device
                if (!hw_hold) { hwq->started = hw_pos; hwq->cv.notify_all(); }
            }
            jitter();
            t();
            if (hwq && hw_hold) { std::lock_guard<std::mutex> lk(hwq->mu); hwq->started = hw_pos; hwq->cv.notify_all(); }
            {
                std::lock_guard<std::mutex> lk(mu);
                completed++;
                running = nullptr;
            }
            cv_done.notify_all();
        }
    }
    uint64_t push(const char *what, std::function<void()> t) {
        uint64_t n;
        {
            // like the hardware's launch queue, a stream takes a bounded number of pending operations; the host blocks beyond
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return q.size() < kMaxQueued; });
            if (hwq) { std::lock_guard<std::mutex> hl(hwq->mu); hw_seq.push_back(++hwq->tail); }   // position in the hardware queue = host enqueue order
            q.emplace_back(what, std::move(t));
            n = ++submitted;
        }
        cv.notify_one();
        return n;
    }
    void wait_for(uint64_t n) {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return completed >= n; });
    }
    uint64_t tail() { std::lock_guard<std::mutex> lk(mu); return submitted; }
    bool idle() { std::lock_guard<std::mutex> lk(mu); return completed >= submitted; }
};
std::mutex g_streams_mu;
std::set<Stream *> g_streams;
// Capture-mode rules: while a capture begun in Global mode is in progress (ThreadLocal: by this thread), "potentially unsafe"
// calls — allocation, registration, IPC mapping, synchronous copies, device synchronization — fail with
// cudaErrorStreamCaptureUnsupported and invalidate the capture, unless the calling thread has switched itself to Relaxed
// with cudaThreadExchangeStreamCaptureMode (what a library does around its own such calls).
std::atomic<int> g_global_captures{0};
thread_local int t_local_captures = 0;
thread_local enum cudaStreamCaptureMode t_capture_mode = cudaStreamCaptureModeGlobal;
struct Capture;
void invalidate_all_captures();
Stream *g_null = nullptr;

Stream *new_stream(unsigned flags, bool is_null) {
    Stream *s = new Stream();
    s->flags = flags; s->is_null = is_null; s->dev = t_dev;
    { static std::atomic<unsigned> created{0}; auto &qs = hw_queues(); if (!qs.empty()) s->hwq = qs[created.fetch_add(1) % qs.size()]; }
    s->th = std::thread([s] { s->worker(); });
    std::lock_guard<std::mutex> lk(g_streams_mu);
    g_streams.insert(s);
    return s;
}
Stream *null_stream() {   // one legacy default stream per device
    static std::mutex mu;
    static std::map<int, Stream *> per_dev;
    std::lock_guard<std::mutex> lk(mu);
    Stream *&s = per_dev[t_dev];
    if (!s) s = new_stream(0, true);
    g_null = s;
    return s;
}
Stream *S(cudaStream_t s) {
    if (s == nullptr || s == cudaStreamLegacy || s == cudaStreamPerThread) return null_stream();
    return (Stream *)s;
}
bool unsafe_call_during_capture(const char *what) {
    if (t_capture_mode == cudaStreamCaptureModeRelaxed) return false;
    if (g_global_captures.load() == 0 && t_local_captures == 0) return false;
    fprintf(stderr, "emu: %s during a stream capture (capture invalidated; a library wraps such calls in cudaThreadExchangeStreamCaptureMode(Relaxed))\n", what);
    invalidate_all_captures();
    return true;
}
#define EMU_UNSAFE(what) do { if (unsafe_call_during_capture(what)) return fail(cudaErrorStreamCaptureUnsupported); } while (0)

// `s` will not run anything submitted from now on before `a`'s current tail has completed
void order_after(Stream *s, Stream *a) {
    if (s == a) return;
    const uint64_t n = a->tail();
    if (n == 0 || a->idle()) return;
    s->push("order-after", [a, n] { a->wait_for(n); });
}
// Submit work to a stream outside capture, with the legacy default stream's implicit synchronization.
void submit(Stream *s, const char *what, std::function<void()> t) {
    if (s->is_null) {
        std::vector<Stream *> blocking;
        { std::lock_guard<std::mutex> lk(g_streams_mu); for (Stream *b : g_streams) if (!b->is_null && !(b->flags & cudaStreamNonBlocking)) blocking.push_back(b); }
        for (Stream *b : blocking) order_after(s, b);
    } else if (!(s->flags & cudaStreamNonBlocking)) {
        order_after(s, null_stream());
    }
    s->push(what, std::move(t));
}
// Inside capture: append a node that depends on the stream's current frontier. g_graph_mu held.
Node *capture_node(Stream *s, Node *n) {
    n->owner = s->cap->graph;
    n->deps = s->cap_deps;
    s->cap->graph->nodes.push_back(n);
    s->cap_deps.assign(1, n);
    return n;
}

bool wrong_device(Stream *s, const char *what) {
    if (s->dev == t_dev) return false;
    fprintf(stderr, "emu: %s on a stream of device %d while device %d is current in this thread (missing cudaSetDevice?)\n", what, s->dev, t_dev);
    return true;
}

struct Event {
    std::mutex mu;
    std::condition_variable cv;
    uint64_t recorded = 0, done = 0;
    uint64_t t_ns = 0;
    // capture (guarded by g_graph_mu)
    Capture *cap = nullptr;
    std::vector<Node *> cap_nodes;
};
unsigned long long g_capture_ids = 1;

void invalidate_all_captures() {
    std::lock_guard<std::mutex> lk(g_graph_mu);
    std::lock_guard<std::mutex> l2(g_streams_mu);
    for (Stream *st : g_streams) if (st->cap) st->cap->invalid = true;
}
bool reachable(Node *from, Node *target, std::set<Node *> &seen) {   // is `target` an ancestor-or-self of `from`?
    if (from == target) return true;
    if (!seen.insert(from).second) return false;
    for (Node *d : from->deps) if (reachable(d, target, seen)) return true;
    return false;
}

void sync_all() {
    std::vector<Stream *> all;
    { std::lock_guard<std::mutex> lk(g_streams_mu); all.assign(g_streams.begin(), g_streams.end()); }
    for (Stream *s : all) s->wait_for(s->tail());
}

int fake_cuMemGetAddressRange(unsigned long long *base, size_t *size, unsigned long long ptr) {
    std::lock_guard<std::mutex> lk(g_mem_mu);
    const Alloc *a = find_alloc((const void *)ptr);
    if (!a) return 1;
    if (base) *base = (unsigned long long)a->base;
    if (size) *size = a->size;
    return 0;
}
int fake_cuPointerGetAttribute(void *data, int attr, unsigned long long ptr) {
    std::lock_guard<std::mutex> lk(g_mem_mu);
    const Alloc *a = find_alloc((const void *)ptr);
    if (!a || attr != 7 /*CU_POINTER_ATTRIBUTE_BUFFER_ID*/) return 1;
    *(unsigned long long *)data = a->id;
    return 0;
}

// cuStreamBatchMemOp: value writes (release) and value waits (GEQ / EQ, acquire) executed in array order by the stream.
void submit(Stream *s, const char *what, std::function<void()> t);
int fake_cuStreamBatchMemOp(CUstream cs, unsigned int count, CUstreamBatchMemOpParams *ops, unsigned int) {
    Stream *s = S((cudaStream_t)cs);
    if (wrong_device(s, "cuStreamBatchMemOp")) return 400 /*CUDA_ERROR_INVALID_HANDLE*/;
    { std::lock_guard<std::mutex> lk(g_graph_mu); if (s->cap) return 900 /*CUDA_ERROR_STREAM_CAPTURE_UNSUPPORTED*/; }
    auto v = std::make_shared<std::vector<CUstreamBatchMemOpParams>>(ops, ops + count);
    for (const auto &o : *v)
        if (o.operation != CU_STREAM_MEM_OP_WAIT_VALUE_64 && o.operation != CU_STREAM_MEM_OP_WRITE_VALUE_64 && o.operation != CU_STREAM_MEM_OP_WAIT_VALUE_32 &&
            o.operation != CU_STREAM_MEM_OP_WRITE_VALUE_32) return 1 /*CUDA_ERROR_INVALID_VALUE*/;
    submit(s, "batch-memop", [v] {
        for (const auto &o : *v) {
            if (o.operation == CU_STREAM_MEM_OP_WRITE_VALUE_64) __atomic_store_n((uint64_t *)o.writeValue.address, (uint64_t)o.writeValue.value64, __ATOMIC_RELEASE);
            else if (o.operation == CU_STREAM_MEM_OP_WRITE_VALUE_32) __atomic_store_n((uint32_t *)o.writeValue.address, (uint32_t)o.writeValue.value, __ATOMIC_RELEASE);
            else {
                const bool w64 = o.operation == CU_STREAM_MEM_OP_WAIT_VALUE_64;
                unsigned idle = 0;
                for (;;) {
                    if (emu::g_trapped.load()) return;
                    bool ok;
                    if (w64) { const uint64_t cur = __atomic_load_n((const uint64_t *)o.waitValue.address, __ATOMIC_ACQUIRE); ok = o.waitValue.flags == CU_STREAM_WAIT_VALUE_EQ ? cur == o.waitValue.value64 : (int64_t)(cur - o.waitValue.value64) >= 0; }
                    else { const uint32_t cur = __atomic_load_n((const uint32_t *)o.waitValue.address, __ATOMIC_ACQUIRE); ok = o.waitValue.flags == CU_STREAM_WAIT_VALUE_EQ ? cur == o.waitValue.value : (int32_t)(cur - o.waitValue.value) >= 0; }
                    if (ok) break;
                    if (++idle < 64) sched_yield(); else { timespec ts{0, 5000}; nanosleep(&ts, nullptr); }
                }
            }
        }
    });
    return 0;
}

}  // namespace

namespace emu {
void register_kernel(const void *fn, const KernelInfo &ki) {
    std::lock_guard<std::mutex> lk(reg_mu());
    registry().emplace(fn, ki);
}
}  // namespace emu

extern "C" {

// Debug aid: one line per stream (ACX runtime calls it through dlsym when ACX_DBG_PENDING is set).
void acx_emu_dump_streams(void) {
    std::lock_guard<std::mutex> lk(g_streams_mu);
    for (Stream *s : g_streams) {
        std::lock_guard<std::mutex> l2(s->mu);
        fprintf(stderr, "  stream %p%s flags %u: submitted %llu completed %llu queued %zu running %s\n", (void *)s, s->is_null ? " (null)" : "", s->flags,
                (unsigned long long)s->submitted, (unsigned long long)s->completed, s->q.size(), s->running ? s->running : "-");
    }
}

// ---- device management --------------------------------------------------------------------------------------------
cudaError_t cudaGetDeviceCount(int *n) { *n = (int)env_long("ACX_EMU_DEVICES", 1); return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = t_dev; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= env_long("ACX_EMU_DEVICES", 1)) return fail(cudaErrorInvalidDevice); t_dev = d; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof *p);
    snprintf(p->name, sizeof p->name, "acx-emu software device");
    p->multiProcessorCount = (int)env_long("ACX_EMU_SMS", 4);
    p->major = 10; p->minor = 0; p->warpSize = 32; p->maxThreadsPerBlock = 1024;
    p->totalGlobalMem = 1ull << 34; p->sharedMemPerBlock = 48 << 10; p->unifiedAddressing = 1; p->canMapHostMemory = 1;
    return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int *v, enum cudaDeviceAttr a, int) {
    switch (a) {
        case cudaDevAttrMultiProcessorCount: *v = (int)env_long("ACX_EMU_SMS", 4); break;
        case cudaDevAttrCanUseHostPointerForRegisteredMem: *v = env_long("ACX_EMU_NO_HOSTREG_UVA", 0) ? 0 : 1; break;   // 0: exercise the runtime's path for devices whose registered host memory has a separate device alias
        case cudaDevAttrPageableMemoryAccess: *v = 0; break;
        case cudaDevAttrComputeCapabilityMajor: *v = 10; break;
        case cudaDevAttrComputeCapabilityMinor: *v = 0; break;
        case cudaDevAttrMaxThreadsPerBlock: *v = 1024; break;
        case cudaDevAttrWarpSize: *v = 32; break;
        default: *v = 0; break;
    }
    return cudaSuccess;
}
cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) {
    EMU_UNSAFE("cudaDeviceSynchronize");
    sync_all(); return dead_or(cudaSuccess); }
cudaError_t cudaGetLastError(void) { cudaError_t e = t_last; t_last = cudaSuccess; return e; }
cudaError_t cudaPeekAtLastError(void) { return t_last; }
const char *cudaGetErrorString(cudaError_t e) {
    switch (e) {
        case cudaSuccess: return "no error";
        case cudaErrorInvalidValue: return "invalid argument";
        case cudaErrorMemoryAllocation: return "out of memory";
        case cudaErrorLaunchFailure: return "unspecified launch failure (device trap)";
        case cudaErrorInvalidDeviceFunction: return "invalid device function (kernel not registered with the emulation)";
        case cudaErrorInvalidConfiguration: return "invalid configuration argument";
        case cudaErrorInvalidResourceHandle: return "invalid resource handle";
        case cudaErrorNotReady: return "device not ready";
        case cudaErrorStreamCaptureUnjoined: return "capture was not joined to the origin stream (unjoined work)";
        case cudaErrorStreamCaptureUnsupported: return "operation not permitted when stream is capturing";
        case cudaErrorStreamCaptureInvalidated: return "operation failed due to a previous error during capture";
        case cudaErrorStreamCaptureIsolation: return "dependency created on uncaptured work in another stream";
        case cudaErrorStreamCaptureUnmatched: return "capture was not initiated in this stream";
        case cudaErrorIllegalState: return "illegal state";
        case cudaErrorInvalidDevice: return "invalid device ordinal";
        case cudaErrorNotSupported: return "operation not supported by the emulation";
        default: return "CUDA error (emulated runtime)";
    }
}
const char *cudaGetErrorName(cudaError_t e) { return cudaGetErrorString(e); }
cudaError_t cudaDriverGetVersion(int *v) { *v = 12090; return cudaSuccess; }
cudaError_t cudaRuntimeGetVersion(int *v) { *v = 12090; return cudaSuccess; }
cudaError_t cudaGetDriverEntryPoint(const char *symbol, void **fn, unsigned long long, enum cudaDriverEntryPointQueryResult *st) {
    *fn = nullptr;
    if (!strcmp(symbol, "cuMemGetAddressRange")) *fn = (void *)fake_cuMemGetAddressRange;
    else if (!strcmp(symbol, "cuPointerGetAttribute")) *fn = (void *)fake_cuPointerGetAttribute;
    else if (!strcmp(symbol, "cuStreamBatchMemOp")) *fn = (void *)fake_cuStreamBatchMemOp;
    if (st) *st = *fn ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
    return *fn ? cudaSuccess : fail(cudaErrorNotSupported);
}
cudaError_t cudaFuncSetAttribute(const void *, enum cudaFuncAttribute, int) { return cudaSuccess; }

// ---- memory -------------------------------------------------------------------------------------------------------
cudaError_t cudaMalloc(void **p, size_t size) {
    EMU_UNSAFE("cudaMalloc");
    if (!p) return fail(cudaErrorInvalidValue);
    if (size == 0) { *p = nullptr; return cudaSuccess; }
    const size_t map = (size + 4095) & ~(size_t)4095;
    int fd = (int)syscall(SYS_memfd_create, "acxemu", 0u);
    if (fd < 0 || ftruncate(fd, (off_t)map) != 0) { if (fd >= 0) close(fd); return fail(cudaErrorMemoryAllocation); }
    void *m = mmap(nullptr, map, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) { close(fd); return fail(cudaErrorMemoryAllocation); }
    std::lock_guard<std::mutex> lk(g_mem_mu);
    g_allocs[(uintptr_t)m] = Alloc{(char *)m, map, g_next_id++, kDevOwn, fd, t_dev};
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
    if (!p) return cudaSuccess;
    EMU_UNSAFE("cudaFree");
    sync_all();   // cudaFree synchronizes the device: keeps the hazard the design documents (spinning gates vs. cudaFree) observable
    std::lock_guard<std::mutex> lk(g_mem_mu);
    auto it = g_allocs.find((uintptr_t)p);
    if (it == g_allocs.end() || it->second.kind != kDevOwn) return fail(cudaErrorInvalidValue);
    munmap(it->second.base, it->second.size);
    close(it->second.fd);
    g_allocs.erase(it);
    return cudaSuccess;
}
cudaError_t cudaMallocManaged(void **p, size_t size, unsigned) { return cudaMalloc(p, size); }
cudaError_t cudaHostAlloc(void **p, size_t size, unsigned) {
    EMU_UNSAFE("cudaHostAlloc");
    void *m = nullptr;
    if (posix_memalign(&m, 4096, size ? size : 1)) return fail(cudaErrorMemoryAllocation);
    std::lock_guard<std::mutex> lk(g_mem_mu);
    g_allocs[(uintptr_t)m] = Alloc{(char *)m, size ? size : 1, g_next_id++, kPinned, -1};
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void **p, size_t size) { return cudaHostAlloc(p, size, 0); }
cudaError_t cudaFreeHost(void *p) {
    if (!p) return cudaSuccess;
    std::lock_guard<std::mutex> lk(g_mem_mu);
    auto it = g_allocs.find((uintptr_t)p);
    if (it == g_allocs.end() || it->second.kind != kPinned) return fail(cudaErrorInvalidValue);
    g_allocs.erase(it);
    free(p);
    return cudaSuccess;
}
cudaError_t cudaHostRegister(void *p, size_t size, unsigned) {
    EMU_UNSAFE("cudaHostRegister");
    std::lock_guard<std::mutex> lk(g_mem_mu);
    // like CUDA: a range overlapping an existing registration is rejected
    auto it = g_allocs.lower_bound((uintptr_t)p);
    if (it != g_allocs.end() && it->first < (uintptr_t)p + size) return fail(cudaErrorHostMemoryAlreadyRegistered);
    if (it != g_allocs.begin()) { --it; if ((uintptr_t)it->second.base + it->second.size > (uintptr_t)p) return fail(cudaErrorHostMemoryAlreadyRegistered); }
    g_allocs[(uintptr_t)p] = Alloc{(char *)p, size, g_next_id++, kRegistered, -1};
    return cudaSuccess;
}
cudaError_t cudaHostUnregister(void *p) {
    EMU_UNSAFE("cudaHostUnregister");
    std::lock_guard<std::mutex> lk(g_mem_mu);
    auto it = g_allocs.find((uintptr_t)p);
    if (it == g_allocs.end() || it->second.kind != kRegistered) return fail(cudaErrorHostMemoryNotRegistered);
    g_allocs.erase(it);
    return cudaSuccess;
}
cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned) { *d = h; return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) {
    memset(a, 0, sizeof *a);
    std::lock_guard<std::mutex> lk(g_mem_mu);
    const Alloc *x = find_alloc(p);
    if (!x) { a->type = cudaMemoryTypeUnregistered; a->device = -1; return cudaSuccess; }
    if (x->kind == kDevOwn || x->kind == kDevIpc) { a->type = cudaMemoryTypeDevice; a->device = x->dev; a->devicePointer = (void *)p; }
    else { a->type = cudaMemoryTypeHost; a->device = t_dev; a->devicePointer = (void *)p; a->hostPointer = (void *)p; }
    return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
    std::lock_guard<std::mutex> lk(g_mem_mu);
    auto it = g_allocs.find((uintptr_t)p);
    if (it == g_allocs.end() || it->second.kind != kDevOwn) return fail(cudaErrorInvalidValue);   // base of an own allocation only
    IpcHandle x{kIpcMagic, (int32_t)getpid(), it->second.fd, (uint32_t)it->second.dev, it->second.size, it->second.id};
    memset(h, 0, sizeof *h);
    memcpy(h, &x, sizeof x);
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
    EMU_UNSAFE("cudaIpcOpenMemHandle");
    IpcHandle x;
    memcpy(&x, &h, sizeof x);
    if (x.magic != kIpcMagic) return fail(cudaErrorInvalidValue);
    if (x.pid == (int32_t)getpid()) return fail(cudaErrorDeviceUninitialized);   // CUDA refuses to open a handle in the exporting process
    char path[64];
    snprintf(path, sizeof path, "/proc/%d/fd/%d", x.pid, x.fd);
    int fd = open(path, O_RDWR);
    if (fd < 0) return fail(cudaErrorInvalidValue);   // the exporter has freed the allocation (or is gone)
    void *m = mmap(nullptr, x.size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return fail(cudaErrorMemoryAllocation);
    std::lock_guard<std::mutex> lk(g_mem_mu);
    g_allocs[(uintptr_t)m] = Alloc{(char *)m, (size_t)x.size, g_next_id++, kDevIpc, -1, (int)x.dev};
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *p) {
    std::lock_guard<std::mutex> lk(g_mem_mu);
    auto it = g_allocs.find((uintptr_t)p);
    if (it == g_allocs.end() || it->second.kind != kDevIpc) return fail(cudaErrorInvalidValue);
    munmap(it->second.base, it->second.size);
    g_allocs.erase(it);
    return cudaSuccess;
}

// ---- streams ------------------------------------------------------------------------------------------------------
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags) { *s = (cudaStream_t)new_stream(flags, false); return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t *s) { return cudaStreamCreateWithFlags(s, 0); }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned flags, int) { return cudaStreamCreateWithFlags(s, flags); }
cudaError_t cudaStreamDestroy(cudaStream_t cs) {
    Stream *s = S(cs);
    if (s->is_null) return fail(cudaErrorInvalidResourceHandle);
    s->wait_for(s->tail());   // CUDA lets queued work drain; so do we (synchronously)
    { std::lock_guard<std::mutex> lk(s->mu); s->stop = true; }
    s->cv.notify_all();
    s->th.join();
    { std::lock_guard<std::mutex> lk(g_streams_mu); g_streams.erase(s); }
    delete s;
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t cs) {
    Stream *s = S(cs);
    { std::lock_guard<std::mutex> lk(g_graph_mu); if (s->cap) return fail(cudaErrorStreamCaptureUnsupported); }
    if (s->is_null) {   // legacy default stream: also waits for the blocking streams
        std::vector<Stream *> blocking;
        { std::lock_guard<std::mutex> lk(g_streams_mu); for (Stream *b : g_streams) if (!b->is_null && !(b->flags & cudaStreamNonBlocking)) blocking.push_back(b); }
        for (Stream *b : blocking) b->wait_for(b->tail());
    }
    s->wait_for(s->tail());
    return dead_or(cudaSuccess);
}
cudaError_t cudaStreamQuery(cudaStream_t cs) {
    Stream *s = S(cs);
    if (emu::g_trapped.load()) return fail(cudaErrorLaunchFailure);
    return s->idle() ? cudaSuccess : fail(cudaErrorNotReady);
}

// ---- events -------------------------------------------------------------------------------------------------------
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t) new Event(); return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventDestroy(cudaEvent_t ce) {
    Event *e = (Event *)ce;
    if (!e) return fail(cudaErrorInvalidResourceHandle);
    // a pending record/wait task may still reference the event and cudaEventDestroy must not block: park it
    static std::mutex mu;
    static std::vector<Event *> *graveyard = new std::vector<Event *>();   // never destroyed: stays reachable for leak checkers
    std::lock_guard<std::mutex> lk(mu);
    graveyard->push_back(e);
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t ce, cudaStream_t cs) {
    Event *e = (Event *)ce;
    Stream *s = S(cs);
    if (!e || wrong_device(s, "cudaEventRecord")) return fail(cudaErrorInvalidResourceHandle);
    {
        std::lock_guard<std::mutex> lk(g_graph_mu);
        if (s->cap) { e->cap = s->cap; e->cap_nodes = s->cap_deps; return cudaSuccess; }
        e->cap = nullptr; e->cap_nodes.clear();
    }
    uint64_t gen;
    { std::lock_guard<std::mutex> lk(e->mu); gen = ++e->recorded; }
    submit(s, "event-record", [e, gen] {
        { std::lock_guard<std::mutex> lk(e->mu); if (e->done < gen) e->done = gen; e->t_ns = emu::now_ns(); }
        e->cv.notify_all();
    });
    return cudaSuccess;
}
cudaError_t cudaEventRecordWithFlags(cudaEvent_t e, cudaStream_t s, unsigned) { return cudaEventRecord(e, s); }
cudaError_t cudaStreamWaitEvent(cudaStream_t cs, cudaEvent_t ce, unsigned) {
    Event *e = (Event *)ce;
    Stream *s = S(cs);
    if (!e) return fail(cudaErrorInvalidResourceHandle);
    {
        std::lock_guard<std::mutex> lk(g_graph_mu);
        if (e->cap) {   // the event was recorded inside a capture: the waiting stream joins that capture (fork) / merges frontiers (join)
            if (s->cap && s->cap != e->cap) return fail(cudaErrorStreamCaptureIsolation);
            if (!s->cap) { s->cap = e->cap; s->cap_deps.clear(); e->cap->members.insert(s); }
            for (Node *n : e->cap_nodes) if (std::find(s->cap_deps.begin(), s->cap_deps.end(), n) == s->cap_deps.end()) s->cap_deps.push_back(n);
            return cudaSuccess;
        }
        if (s->cap) return fail(cudaErrorStreamCaptureIsolation);   // a capturing stream may not depend on uncaptured work
    }
    uint64_t gen;
    { std::lock_guard<std::mutex> lk(e->mu); gen = e->recorded; if (gen == 0 || e->done >= gen) return cudaSuccess; }
    submit(s, "event-wait", [e, gen] { std::unique_lock<std::mutex> lk(e->mu); e->cv.wait(lk, [&] { return e->done >= gen; }); });
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t ce) {
    Event *e = (Event *)ce;
    if (!e) return fail(cudaErrorInvalidResourceHandle);
    std::unique_lock<std::mutex> lk(e->mu);
    const uint64_t gen = e->recorded;
    e->cv.wait(lk, [&] { return e->done >= gen; });
    return dead_or(cudaSuccess);
}
cudaError_t cudaEventQuery(cudaEvent_t ce) {
    Event *e = (Event *)ce;
    if (!e) return fail(cudaErrorInvalidResourceHandle);
    std::lock_guard<std::mutex> lk(e->mu);
    return e->done >= e->recorded ? cudaSuccess : fail(cudaErrorNotReady);   // like the real runtime: "not ready" becomes the thread's last error
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    Event *x = (Event *)a, *y = (Event *)b;
    if (!x || !y) return fail(cudaErrorInvalidResourceHandle);
    *ms = (float)(((double)y->t_ns - (double)x->t_ns) * 1e-6);
    return cudaSuccess;
}

// ---- work submission ------------------------------------------------------------------------------------------------
static cudaError_t submit_or_capture(Stream *s, Node *n, const char *what, std::function<void()> t) {
    if (s->dev != t_dev) {   // real CUDA: invalid resource handle — the calling thread has another device current than the stream's
        fprintf(stderr, "emu: %s submitted to a stream of device %d while device %d is current in this thread (missing cudaSetDevice?)\n", what, s->dev, t_dev);
        delete n;
        return fail(cudaErrorInvalidResourceHandle);
    }
    {
        std::lock_guard<std::mutex> lk(g_graph_mu);
        if (s->cap) { capture_node(s, n); return cudaSuccess; }
    }
    delete n;
    submit(s, what, std::move(t));
    return cudaSuccess;
}
cudaError_t cudaLaunchKernel(const void *func, dim3 grid, dim3 block, void **args, size_t, cudaStream_t cs) {
    if (emu::g_trapped.load()) return fail(cudaErrorLaunchFailure);
    auto kl = std::make_shared<KernelLaunch>();
    cudaError_t e = make_launch(func, grid, block, args, kl.get());
    if (e != cudaSuccess) return fail(e);
    Node *n = new Node();
    n->type = kNodeKernel; n->kl = *kl;
    return submit_or_capture(S(cs), n, kl->ki->name, [kl] { if (!emu::g_trapped.load()) kl->run(); });
}
cudaError_t cudaLaunchKernelExC(const cudaLaunchConfig_t *cfg, const void *func, void **args) {
    // launch attributes (programmatic stream serialization) only relax ordering on hardware; in-order execution is a legal schedule
    return cudaLaunchKernel(func, cfg->gridDim, cfg->blockDim, args, cfg->dynamicSmemBytes, cfg->stream);
}
// With unified addressing the runtime knows what each pointer is: an explicit direction that contradicts it is an error.
static bool is_device_memory(const void *p) {
    std::lock_guard<std::mutex> lk(g_mem_mu);
    const Alloc *a = find_alloc(p);
    return a && (a->kind == kDevOwn || a->kind == kDevIpc);
}
static bool memcpy_kind_ok(void *dst, const void *src, enum cudaMemcpyKind k) {
    switch (k) {
        case cudaMemcpyHostToDevice: return !is_device_memory(src) && is_device_memory(dst);
        case cudaMemcpyDeviceToHost: return is_device_memory(src) && !is_device_memory(dst);
        case cudaMemcpyDeviceToDevice: return is_device_memory(src) && is_device_memory(dst);
        case cudaMemcpyHostToHost: return !is_device_memory(src) && !is_device_memory(dst);
        default: return true;
    }
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, enum cudaMemcpyKind k, cudaStream_t cs) {
    if (n && !memcpy_kind_ok(dst, src, k)) { fprintf(stderr, "emu: cudaMemcpy direction %d contradicts the pointers (%p <- %p)\n", (int)k, dst, src); return fail(cudaErrorInvalidValue); }
    Node *nd = new Node();
    nd->type = kNodeMemcpy; nd->dst = dst; nd->src = src; nd->bytes = n;
    return submit_or_capture(S(cs), nd, "memcpy", [dst, src, n] { memmove(dst, src, n); });
}
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, enum cudaMemcpyKind k) {
    EMU_UNSAFE("cudaMemcpy (synchronous)");
    cudaError_t e = cudaMemcpyAsync(dst, src, n, k, nullptr);
    return e != cudaSuccess ? e : cudaStreamSynchronize(nullptr);
}
cudaError_t cudaMemsetAsync(void *dst, int v, size_t n, cudaStream_t cs) {
    Node *nd = new Node();
    nd->type = kNodeMemset; nd->dst = dst; nd->value = v; nd->bytes = n;
    return submit_or_capture(S(cs), nd, "memset", [dst, v, n] { memset(dst, v, n); });
}
cudaError_t cudaMemset(void *dst, int v, size_t n) {
    EMU_UNSAFE("cudaMemset (synchronous)");
    cudaError_t e = cudaMemsetAsync(dst, v, n, nullptr);
    return e != cudaSuccess ? e : cudaStreamSynchronize(nullptr);
}
cudaError_t cudaLaunchHostFunc(cudaStream_t cs, cudaHostFn_t fn, void *arg) {
    Node *nd = new Node();
    nd->type = kNodeHost; nd->fn = fn; nd->fn_arg = arg;
    return submit_or_capture(S(cs), nd, "host-fn", [fn, arg] { fn(arg); });
}

// ---- stream capture -------------------------------------------------------------------------------------------------
cudaError_t cudaStreamBeginCapture(cudaStream_t cs, enum cudaStreamCaptureMode mode) {
    Stream *s = S(cs);
    if (s->is_null) return fail(cudaErrorStreamCaptureUnsupported);   // the legacy stream cannot be captured
    std::lock_guard<std::mutex> lk(g_graph_mu);
    if (s->cap) return fail(cudaErrorIllegalState);
    Capture *c = new Capture();
    c->graph = new Graph(); c->origin = s; c->members.insert(s); c->id = g_capture_ids++;
    c->mode = mode;
    if (mode == cudaStreamCaptureModeGlobal) g_global_captures++; else if (mode == cudaStreamCaptureModeThreadLocal) t_local_captures++;
    g_graphs.insert(c->graph);
    s->cap = c; s->cap_deps.clear();
    return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t cs, cudaGraph_t *out) {
    Stream *s = S(cs);
    std::lock_guard<std::mutex> lk(g_graph_mu);
    Capture *c = s->cap;
    if (!c) return fail(cudaErrorIllegalState);
    if (c->origin != s) return fail(cudaErrorStreamCaptureUnmatched);
    // every forked stream must have been joined back: its frontier has to be an ancestor of the origin's frontier
    bool unjoined = false;
    for (Stream *m : c->members) {
        if (m == s) continue;
        for (Node *f : m->cap_deps) {
            bool ok = false;
            for (Node *o : s->cap_deps) { std::set<Node *> seen; if (reachable(o, f, seen)) { ok = true; break; } }
            if (!ok) unjoined = true;
        }
    }
    for (Stream *m : c->members) { m->cap = nullptr; m->cap_deps.clear(); }
    Graph *g = c->graph;
    if (c->mode == cudaStreamCaptureModeGlobal) g_global_captures--; else if (c->mode == cudaStreamCaptureModeThreadLocal) t_local_captures--;
    const bool invalid = c->invalid;
    delete c;
    if (invalid) { g_graphs.erase(g); delete g; if (out) *out = nullptr; return fail(cudaErrorStreamCaptureInvalidated); }
    if (unjoined) { g_graphs.erase(g); delete g; if (out) *out = nullptr; return fail(cudaErrorStreamCaptureUnjoined); }
    *out = (cudaGraph_t)g;
    return cudaSuccess;
}
cudaError_t cudaThreadExchangeStreamCaptureMode(enum cudaStreamCaptureMode *mode) {
    const enum cudaStreamCaptureMode old = t_capture_mode;
    t_capture_mode = *mode;
    *mode = old;
    return cudaSuccess;
}
cudaError_t cudaStreamIsCapturing(cudaStream_t cs, enum cudaStreamCaptureStatus *st) {
    Stream *s = S(cs);
    std::lock_guard<std::mutex> lk(g_graph_mu);
    *st = s->cap ? cudaStreamCaptureStatusActive : cudaStreamCaptureStatusNone;
    return cudaSuccess;
}
cudaError_t cudaStreamGetCaptureInfo_v2(cudaStream_t cs, enum cudaStreamCaptureStatus *st, unsigned long long *id, cudaGraph_t *graph, const cudaGraphNode_t **deps,
                                        size_t *ndeps) {
    Stream *s = S(cs);
    std::lock_guard<std::mutex> lk(g_graph_mu);
    *st = s->cap ? cudaStreamCaptureStatusActive : cudaStreamCaptureStatusNone;
    if (s->cap) {
        if (id) *id = s->cap->id;
        if (graph) *graph = (cudaGraph_t)s->cap->graph;
        if (deps) *deps = (const cudaGraphNode_t *)s->cap_deps.data();
        if (ndeps) *ndeps = s->cap_deps.size();
    }
    return cudaSuccess;
}

// ---- graphs ---------------------------------------------------------------------------------------------------------
cudaError_t cudaGraphCreate(cudaGraph_t *g, unsigned) {
    Graph *x = new Graph();
    std::lock_guard<std::mutex> lk(g_graph_mu);
    g_graphs.insert(x);
    *g = (cudaGraph_t)x;
    return cudaSuccess;
}
cudaError_t cudaGraphDestroy(cudaGraph_t cg) {
    Graph *g = (Graph *)cg;
    {
        std::lock_guard<std::mutex> lk(g_graph_mu);
        if (!g_graphs.erase(g)) return fail(cudaErrorInvalidResourceHandle);
    }
    delete g;   // outside the lock: user-object destructors run here
    return cudaSuccess;
}
static cudaError_t add_node(cudaGraphNode_t *out, cudaGraph_t cg, const cudaGraphNode_t *deps, size_t ndeps, Node *n) {
    Graph *g = (Graph *)cg;
    std::lock_guard<std::mutex> lk(g_graph_mu);
    if (!g_graphs.count(g)) { delete n; return fail(cudaErrorInvalidResourceHandle); }
    for (size_t i = 0; i < ndeps; i++) {
        Node *d = (Node *)deps[i];
        if (!d || d->owner != g) { delete n; return fail(cudaErrorInvalidValue); }   // dependencies must be nodes of the same graph
        n->deps.push_back(d);
    }
    n->owner = g;
    g->nodes.push_back(n);
    if (out) *out = (cudaGraphNode_t)n;
    return cudaSuccess;
}
cudaError_t cudaGraphAddKernelNode(cudaGraphNode_t *out, cudaGraph_t g, const cudaGraphNode_t *deps, size_t ndeps, const cudaKernelNodeParams *p) {
    Node *n = new Node();
    n->type = kNodeKernel;
    cudaError_t e = make_launch(p->func, p->gridDim, p->blockDim, p->kernelParams, &n->kl);
    if (e != cudaSuccess) { delete n; return fail(e); }
    return add_node(out, g, deps, ndeps, n);
}
cudaError_t cudaGraphAddEmptyNode(cudaGraphNode_t *out, cudaGraph_t g, const cudaGraphNode_t *deps, size_t ndeps) {
    return add_node(out, g, deps, ndeps, new Node());
}
cudaError_t cudaGraphAddChildGraphNode(cudaGraphNode_t *out, cudaGraph_t g, const cudaGraphNode_t *deps, size_t ndeps, cudaGraph_t child) {
    Node *n = new Node();
    n->type = kNodeChild;
    {
        std::lock_guard<std::mutex> lk(g_graph_mu);
        if (!g_graphs.count((Graph *)child)) { delete n; return fail(cudaErrorInvalidResourceHandle); }
        n->child = clone_graph((Graph *)child);   // the child is copied at this point, references to its user objects included
    }
    return add_node(out, g, deps, ndeps, n);
}
cudaError_t cudaGraphAddMemcpyNode1D(cudaGraphNode_t *out, cudaGraph_t g, const cudaGraphNode_t *deps, size_t ndeps, void *dst, const void *src, size_t n, enum cudaMemcpyKind) {
    Node *nd = new Node();
    nd->type = kNodeMemcpy; nd->dst = dst; nd->src = src; nd->bytes = n;
    return add_node(out, g, deps, ndeps, nd);
}
cudaError_t cudaGraphAddHostNode(cudaGraphNode_t *out, cudaGraph_t g, const cudaGraphNode_t *deps, size_t ndeps, const cudaHostNodeParams *p) {
    Node *nd = new Node();
    nd->type = kNodeHost; nd->fn = p->fn; nd->fn_arg = p->userData;
    return add_node(out, g, deps, ndeps, nd);
}
cudaError_t cudaGraphGetNodes(cudaGraph_t cg, cudaGraphNode_t *nodes, size_t *n) {
    Graph *g = (Graph *)cg;
    std::lock_guard<std::mutex> lk(g_graph_mu);
    if (!g_graphs.count(g)) return fail(cudaErrorInvalidResourceHandle);
    if (nodes) for (size_t i = 0; i < *n && i < g->nodes.size(); i++) nodes[i] = (cudaGraphNode_t)g->nodes[i];
    *n = g->nodes.size();
    return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *exec, cudaGraph_t cg, unsigned long long) {
    Graph *g = (Graph *)cg;
    std::lock_guard<std::mutex> lk(g_graph_mu);
    if (!g_graphs.count(g)) return fail(cudaErrorInvalidResourceHandle);
    Graph *x = clone_graph(g);
    x->is_exec = true;
    g_graphs.insert(x);
    *exec = (cudaGraphExec_t)x;
    return cudaSuccess;
}
cudaError_t cudaGraphInstantiateWithFlags(cudaGraphExec_t *exec, cudaGraph_t g, unsigned long long f) { return cudaGraphInstantiate(exec, g, f); }
cudaError_t cudaGraphLaunch(cudaGraphExec_t ce, cudaStream_t cs) {
    Graph *x = (Graph *)ce;
    Stream *s = S(cs);
    {
        std::lock_guard<std::mutex> lk(g_graph_mu);
        if (!g_graphs.count(x) || !x->is_exec || wrong_device(s, "cudaGraphLaunch")) return fail(cudaErrorInvalidResourceHandle);
        if (s->cap) return fail(cudaErrorNotSupported);
    }
    if (emu::g_trapped.load()) return fail(cudaErrorLaunchFailure);
    { std::lock_guard<std::mutex> lk(g_graph_mu); x->inflight++; }
    submit(s, "graph-launch", [x] {
        run_graph(x);
        bool last;
        { std::lock_guard<std::mutex> lk(g_graph_mu); last = --x->inflight == 0 && x->zombie; }
        if (last) delete x;
    });
    return cudaSuccess;
}
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t ce) {
    Graph *x = (Graph *)ce;
    {
        std::lock_guard<std::mutex> lk(g_graph_mu);
        if (!g_graphs.erase(x)) return fail(cudaErrorInvalidResourceHandle);
        if (x->inflight) { x->zombie = true; return cudaSuccess; }   // like CUDA: destruction is deferred behind launches in flight
    }
    delete x;
    return cudaSuccess;
}
cudaError_t cudaUserObjectCreate(cudaUserObject_t *out, void *ptr, cudaHostFn_t destroy, unsigned refs, unsigned) {
    if (!destroy || refs == 0) return fail(cudaErrorInvalidValue);
    UserObj *o = new UserObj{ptr, destroy, {(long)refs}};
    *out = (cudaUserObject_t)o;
    return cudaSuccess;
}
cudaError_t cudaUserObjectRetain(cudaUserObject_t o, unsigned n) { ((UserObj *)o)->refs.fetch_add(n); return cudaSuccess; }
cudaError_t cudaUserObjectRelease(cudaUserObject_t o, unsigned n) { uobj_release((UserObj *)o, n); return cudaSuccess; }
cudaError_t cudaGraphRetainUserObject(cudaGraph_t cg, cudaUserObject_t co, unsigned n, unsigned flags) {
    Graph *g = (Graph *)cg;
    UserObj *o = (UserObj *)co;
    std::lock_guard<std::mutex> lk(g_graph_mu);
    if (!g_graphs.count(g) || !o) return fail(cudaErrorInvalidResourceHandle);
    if (!(flags & cudaGraphUserObjectMove)) o->refs.fetch_add(n);   // Move: the caller's references become the graph's
    g->uobjs.push_back({o, (long)n});
    return cudaSuccess;
}
cudaError_t cudaGraphReleaseUserObject(cudaGraph_t cg, cudaUserObject_t co, unsigned n) {
    Graph *g = (Graph *)cg;
    UserObj *o = (UserObj *)co;
    {
        std::lock_guard<std::mutex> lk(g_graph_mu);
        if (!g_graphs.count(g)) return fail(cudaErrorInvalidResourceHandle);
        bool found = false;
        for (auto &u : g->uobjs) if (u.first == o && u.second >= (long)n) { u.second -= n; found = true; break; }
        if (!found) return fail(cudaErrorInvalidValue);
    }
    uobj_release(o, n);
    return cudaSuccess;
}

}  // extern "C"
