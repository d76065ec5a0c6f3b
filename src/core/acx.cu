// acx-b200 runtime: init/finalize, windows + CUDA-IPC bootstrap, allocators, and the
// host front-ends of the whole MPIX_* API (enqueue / wait / graph / partitioned).
//
// Behavioural contract mirrors NVIDIA/mpi-acx (reference src/init.cpp, src/sendrecv.cu,
// src/partitioned.cu, src/triggered.cpp); the mechanism does not: see internal.h.
#include <mpi.h>
#include "../../include/mpi-acx.h"
#include "internal.h"
#include "kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <thread>
#include <unistd.h>
#include <sched.h>
#include <fcntl.h>
#include <sys/syscall.h>
#include <ctype.h>
#include <nvtx3/nvToolsExt.h>

AcxState *g_acx = nullptr;

using namespace acxk;

// =======================================================================================
// small utilities
// =======================================================================================
int acx_env_int(const char *name, const char *alt, long dflt, long *out) {
    const char *v = getenv(name);
    if ((!v || !*v) && alt) v = getenv(alt);
    if (!v || !*v) { *out = dflt; return 0; }
    char *end = nullptr;
    long x = strtol(v, &end, 0);
    if (end && (*end == 'k' || *end == 'K')) x <<= 10;
    else if (end && (*end == 'm' || *end == 'M')) x <<= 20;
    else if (end && (*end == 'g' || *end == 'G')) x <<= 30;
    *out = x;
    return 1;
}

static uint64_t *g_diag = nullptr, *g_diag_d = nullptr;      // watchdog diagnostics (pinned)
static uint64_t *g_trace_cursor = nullptr;                   // device
static std::vector<int> g_pending_retire;                    // slots whose wait has been enqueued; freed once the device marks them retired
struct DeferredSlot { int slot; unsigned long long init_gen; };
static std::vector<DeferredSlot> g_deferred_slots;           // freed by graph user-object destructors (CUDA may run those on an internal thread, after cudaGraphDestroy has returned)
static std::atomic<int> g_graph_objs_live{0};                // user objects created and not yet destroyed
static std::mutex g_deferred_mu;

// Stream for the calling thread's host-blocking operations (MPIX_Wait, MPIX_Start(all), host Pready / Parrived, host waits
// of partitioned rounds). One per application thread: these calls launch kernels that may SPIN (a wait for a message
// that has not been sent yet), and on a shared stream one thread's spinning wait would hold back another thread's
// unrelated Pready or Start — with three or more ranks that closes a cycle (tests/c/threads.c found it on the software
// device). The thread that called MPIX_Init uses the library's first stream.
static std::mutex g_host_streams_mu;
static std::vector<cudaStream_t> g_host_streams;   // every stream handed out, destroyed at MPIX_Finalize
static unsigned long long g_init_generation = 0;   // bumped by MPIX_Init: streams of an earlier init are never reused
static cudaStream_t host_stream() {
    thread_local cudaStream_t mine = nullptr;
    thread_local unsigned long long mine_gen = 0;
    if (mine && mine_gen == g_init_generation) return mine;
    std::lock_guard<std::mutex> lk(g_host_streams_mu);
    cudaStream_t s = nullptr;
    if (g_host_streams.empty()) s = g_acx->istream;
    else {
        // CUDA's current device is per thread (a fresh thread starts on device 0): the stream must live on the library's device
        int cur = -1;
        const bool switched = cudaGetDevice(&cur) == cudaSuccess && cur != g_acx->dev && cudaSetDevice(g_acx->dev) == cudaSuccess;
        const cudaError_t ce = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        if (switched) cudaSetDevice(cur);
        if (ce != cudaSuccess) { cudaGetLastError(); return g_acx->istream; }
    }
    g_host_streams.push_back(s);
    mine = s; mine_gen = g_init_generation;
    return s;
}

// NVTX ranges around the API calls (ACX_NVTX=1); header-only NVTX, a no-op without a tool attached.
static int g_mover_tma = 0, g_mover_256 = 0, g_mover_block = 0;   // g_mover_block: interleaved block size in bytes (ACX_MOVER_BLOCK; 0 = one chunk per CTA)
static int g_apush = 0;                   // deterministic large-message mode: the sender always pushes, gated on its per-channel stream (ACX_SENDER_PUSH)
static std::map<std::pair<int, int>, cudaStream_t> g_chan_streams;   // (peer, channel) -> stream for sender-push movers
struct OverflowStream { cudaStream_t s; AcxRequest *owner; };
static std::vector<OverflowStream> g_ovf_pool;   // pool streams of overflow sends (send_overflow_begin); guarded by g_acx->mu
static int g_side_on = 1;                 // run large movers on per-peer side streams so transfers to different peers overlap (ACX_SIDE_STREAMS)
static std::vector<cudaStream_t> g_side;  // side streams, indexed by peer % size
static int g_wild = 0;                    // ACX_WILDCARDS: MPI_ANY_SOURCE / MPI_ANY_TAG receives on the device transport (experimental; matched at wait time)
static AnyTable *g_any_tbl_d = nullptr;   // device copy of the peer window bases, built on first wildcard use
static acx_anyres *g_anyres_d = nullptr;  // one resolution record per request slot
static int g_coop = 0;                    // ACX_COOP: cooperative moving with chunk stealing (both ends of a large rendezvous message copy it); must be set on every rank
static long g_coop_min = 4 << 20;          // ACX_COOP_MIN_BYTES
static long g_late_min = 4 << 20;          // ACX_LATE_PUSH_MIN: messages from this size up use the sender-moves protocol (RNDV_LATE); 0 = off
static long g_late_helpers = 0;            // ACX_LATE_HELPERS: helper CTAs all late sends of one wait kernel share (0 = one per SM)
static long g_ce_min = 64 << 20;           // ACX_CE_MIN: stream-mode messages from this size up travel by copy engine when the sender's host can see the matching receive's enqueue-time record (0 = off); measured break-even with the SM mover on a 2-GPU ring: 64 MiB (profiles/raw/r2_call11/sweep.txt)
static long g_wait_early_helpers = 0;      // ACX_WAIT_EARLY_HELPERS: a wait kernel with at most this many helper CTAs still releases its programmatic dependents early (measured: no gain on the ring step, profiles/raw/r2_call13/ring.txt — off)
static long g_ce_spingate = 1;             // ACX_CE_SPINGATE: the gate kernel waits for the send's trigger itself (acquire spin on the op record) instead of behind a cross-stream event
static long g_ce_spin_us = 40;             // ACX_CE_SPIN_US: how long MPIX_Isend_enqueue looks for that record before it takes the kernel path
static long g_ce_spin_busy_us = 300;       // ACX_CE_SPIN_BUSY_US: ... when the caller's stream still has work queued (the host is ahead of its GPU: waiting for the neighbour's host costs nothing)
static long g_late_fallback_us = 50;       // ACX_LATE_FALLBACK_US: grace period before a receive's wait pulls an unclaimed RNDV_LATE message itself
static int g_pbatch = 1;                  // ACX_PBATCH (default on since round 2): one kernel launch per MPIX_Startall / MPIX_Waitall group of partitioned requests instead of one per request
static int g_branch = 2;                  // ACX_PEER_BRANCH: run every operation with a peer on that peer's side stream (fork at enqueue, join at the wait): 0 never, 1 always, 2 (default) while the stream is capturing — there the forks are graph edges and cost nothing at run time
static int g_fuse_rndv = 1, g_fuse_min_ctas = 8, g_fuse_max_ctas = 32;   // single-launch rendezvous for grids up to this many CTAs (ACX_FUSE_RNDV, ACX_FUSE_MAX_CTAS)
static long g_send_overflow_max = 128;   // ACX_SEND_OVERFLOW_MAX: overflow sends that may be waiting for ring slots at one time, over all channels (each has a pool stream to itself while it waits)
static int g_send_overflow = 1;   // ACX_SEND_OVERFLOW: the triggers of stream-mode sends beyond a ring's worth of un-waited sends on one channel run on pool streams instead of holding up the caller's (send_overflow_begin)
static int g_pdl = 1;   // programmatic dependent launch between consecutive acx kernels on a stream (ACX_PDL)
static bool g_nvtx = false;
struct AcxRange {
    bool on;
    explicit AcxRange(const char *name) : on(g_nvtx) { if (on) nvtxRangePushA(name); }
    ~AcxRange() { if (on) nvtxRangePop(); }
};
// Fault injection for tests (ACX_FAULT_DROP_SEND=n): the n-th send is never triggered, so the
// matching receive can only end in the watchdog (failure-detection path).
static long g_fault_drop_send = 0;
static long g_fault_delay_us = 0;         // ACX_FAULT_DELAY_US=n: sleep a pseudo-random 0..n us on the host before every send / receive is enqueued (shakes arrival orders)
static std::atomic<long> g_send_counter{0};

// The caller may be inside a stream capture begun in Global (or ThreadLocal) mode: CUDA then refuses "potentially unsafe"
// calls from this thread — cudaHostRegister, cudaIpcOpenMemHandle, cudaMalloc, synchronous copies — and invalidates the
// capture. Library-internal housekeeping of that kind (pinning a host buffer, mapping a peer's registered allocation,
// first-use tables) is independent of what is being captured, so it runs with the thread's capture mode relaxed, as
// cudaThreadExchangeStreamCaptureMode is meant to be used by libraries.
struct CaptureRelax {
    cudaStreamCaptureMode saved = cudaStreamCaptureModeRelaxed;
    bool ok;
    CaptureRelax() { ok = cudaThreadExchangeStreamCaptureMode(&saved) == cudaSuccess; if (!ok) cudaGetLastError(); }
    ~CaptureRelax() { if (ok) cudaThreadExchangeStreamCaptureMode(&saved); }
};

static Watchdog make_wd() {
    Watchdog w;
    w.budget_ns = g_acx->watchdog_ms > 0 ? (uint64_t)g_acx->watchdog_ms * 1000000ull : 0;
    w.diag = g_diag_d;
    return w;
}
static TraceCtx make_tr() {
    TraceCtx t{nullptr, 0, nullptr};
    if (g_acx->trace) {
        t.ring = (acx_trace_rec *)(g_acx->ctrl + g_acx->geom.off_trace);
        t.mask = g_acx->geom.ntrace - 1;
        t.cursor = g_trace_cursor;
    }
    return t;
}

static void report_watchdog() {
    if (g_diag && g_diag[0]) {
        const char *what = "unknown";
        switch (g_diag[0]) {
            case 0x5e01: what = "send trigger: ring slot never freed (receiver not consuming; > ACX_DEPTH unconsumed sends on one channel?)"; break;
            case 0x5e0d: what = "receive matched a message with another tag / communicator (a[63:32] wanted tag, a[31:0] got; b = contexts): more distinct (communicator, tag) pairs between two ranks than ACX_NCHAN-1, posted in another order than sent — raise ACX_NCHAN"; break;
            case 0x5e02: what = "recv trigger: own ring slot still in use (> ACX_DEPTH outstanding receives on one channel)"; break;
            case 0x5e03: what = "wait(send): rendezvous never completed (no matching receive)"; break;
            case 0x5e04: what = "wait(recv): message never arrived (no matching send)"; break;
            case 0x5e05: what = "host wait: operation was never triggered on its stream/graph"; break;
            case 0x5e06: what = "partitioned wait(send): a partition was never marked ready"; break;
            case 0x5e07: what = "partitioned wait(recv): a partition never arrived"; break;
            case 0x5e08: what = "partitioned wait(recv): a device poller never finished its copy-out"; break;
            case 0x5e09: what = "fused rendezvous: CTA 0 never published the match decision"; break;
            case 0x5e0a: what = "sender-push gate: the matching receive was never posted"; break;
            case 0x5e0b: what = "wildcard receive: no matching send arrived"; break;
            case 0x5e0c: what = "cooperative-move gate: the peer never arrived"; break;
            case 0x5e0f: what = "host-predicted ticket (b) differs from the one the trigger drew (a): a memOp wait was built on the wrong slot"; break;
            case 0x5e10: what = "copy-engine send: host-predicted ticket (b) differs from the one the trigger drew (a)"; break;
            case 0x5e11: what = "copy-engine send: the matching receive (enqueued by its host) was never posted on the device"; break;
            case 0x5e12: what = "copy-engine send: the receive's device-side destination (a) is not the one its host published (b), or it is too small"; break;
            case 0x5e13: what = "copy-engine send: the send's trigger never ran (gate kernel without an event, ACX_CE_SPINGATE)"; break;
            case 0x5e14: what = "wait kernel: helper CTAs never checked out (a = helpers, b = checked out)"; break;
            case 0x5e15: what = "overflow send: the channel's ticket counter never reached the ticket the host counted for it (b)"; break;
            case 0x5e16: what = "overflow send: the ticket the host counted (a) was drawn by another send (counter at b)"; break;
            case 0x5e0e: what = "wait kernel helper: the request's waiter CTA never decided (a = ticket, b = state word)"; break;
        }
        ACX_ERR("WATCHDOG (code 0x%llx, a=0x%llx b=0x%llx): %s", (unsigned long long)g_diag[0], (unsigned long long)g_diag[1],
                (unsigned long long)g_diag[2], what);
    }
}

#define ACX_CUDA_WD(stmt)                                                  \
    do {                                                                   \
        cudaError_t _e = (stmt);                                           \
        if (_e != cudaSuccess) {                                           \
            report_watchdog();                                             \
            ACX_ERR("%s failed: %s", #stmt, cudaGetErrorString(_e));       \
            return MPI_ERR_OTHER;                                          \
        }                                                                  \
    } while (0)

// =======================================================================================
// allocators
// =======================================================================================
// Small host-to-device copies of library data go through a pinned scratch page: the library pins USER host buffers page by
// page, so a local variable or a heap object of ours may share a page with one, and CUDA rejects (invalid argument) a
// cudaMemcpy whose host range is only partly pinned — first seen on hardware with the wildcard tables.
static int acx_h2d(void *dst, const void *src, size_t n, cudaStream_t stream) {
    static char *scratch = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!scratch && cudaMallocHost((void **)&scratch, 4096) != cudaSuccess) { cudaGetLastError(); scratch = nullptr; }
    for (size_t off = 0; off < n; off += 4096) {
        const size_t len = std::min<size_t>(4096, n - off);
        const void *from = (const char *)src + off;
        if (scratch) { memcpy(scratch, from, len); from = scratch; }
        cudaError_t e = cudaMemcpyAsync((char *)dst + off, from, len, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { ACX_ERR("host-to-device copy of %zu bytes failed: %s", n, cudaGetErrorString(e)); return MPI_ERR_OTHER; }
    }
    return 0;
}
static std::vector<cudaEvent_t> g_event_pool;
static cudaEvent_t event_get() {
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return e;
}
static void send_flow_release(AcxRequest *r);   // sends in flight per channel (send_overflow_begin)
static void send_flow_forget(AcxRequest *r);
static void release_request_resources(AcxRequest *r) {
    send_flow_forget(r);
    if (r->bounce_off >= 0) { acx_heap_free(r->bounce_off); r->bounce_off = -1; }
    if (r->ev_trig) { g_event_pool.push_back(r->ev_trig); r->ev_trig = nullptr; }
    if (r->ev_done) { g_event_pool.push_back(r->ev_done); r->ev_done = nullptr; }
}

// A rendezvous-capable operation may have a mover (or gate + mover) on a side stream that the caller's stream never
// joins. That kernel reads the operation's op record when it runs — possibly after the wait kernel has retired the
// slot (an eager message into a rendezvous-sized receive completes without the mover) — so the slot and its bounce buffer
// may only be handed to a new owner once the side branch has drained. Found by the software-device stress run
// (tests/emu): a late mover read the NEXT owner's `mover` flag and published `done` early.
static bool side_branch_finished(const AcxRequest *r) {
    if (!r || !r->ev_done) return true;
    if (cudaEventQuery(r->ev_done) != cudaErrorNotReady) return true;   // success, or a dead device (nothing to protect then)
    cudaGetLastError();   // "not ready" is recorded as the thread's last error: clear it, or the caller's next cudaGetLastError() (torch checks every launch) reports it
    return false;
}

void acx_slot_free(int slot) {
    // caller holds g_acx->mu
    AcxRequest *r = g_acx->slot_req[slot];
    if (r) {
        send_flow_release(r);   // a send released without a wait (error exits)
        release_request_resources(r);
        r->magic = 0;
        delete r;
    }
    g_acx->slot_req[slot] = nullptr;
    g_acx->slot_busy[slot] = 0;
    g_acx->hstat[slot].retired = 0;
    g_acx->free_slots.push_back(slot);
    g_acx->slots_in_use--;
}

void acx_reclaim_retired() {
    // caller holds g_acx->mu. Stream-mode slots whose wait kernel has run mark themselves retired
    // in pinned host memory; graph-owned slots come back through the deferred list.
    {
        std::lock_guard<std::mutex> lk(g_deferred_mu);
        for (const DeferredSlot &d : g_deferred_slots)   // a destructor that outlived its MPIX_Init/Finalize epoch names a slot of the OLD table: drop it
            if (d.init_gen == g_init_generation && g_acx->slot_busy[d.slot] == 2) acx_slot_free(d.slot);
        g_deferred_slots.clear();
    }
    size_t keep = 0;
    for (size_t i = 0; i < g_pending_retire.size(); i++) {
        const int s = g_pending_retire[i];
        if (g_acx->slot_busy[s] == 3 && ACX_HOST_FLAG_GET(&g_acx->hstat[s].retired) && side_branch_finished(g_acx->slot_req[s])) acx_slot_free(s);
        else if (g_acx->slot_busy[s] == 3) g_pending_retire[keep++] = s;
    }
    g_pending_retire.resize(keep);
}

int acx_slot_alloc(AcxRequest *r, bool graph) {
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    if (g_acx->free_slots.empty() || g_pending_retire.size() >= 64) acx_reclaim_retired();
    if (g_acx->free_slots.empty()) {
        // Waited-on stream ops retire asynchronously; give in-flight waits a moment. Slots whose wait has run but whose
        // side-stream branch is still queued (a lagging side stream) are worth blocking for: that branch never spins.
        for (int s2 : g_pending_retire)
            if (g_acx->slot_busy[s2] == 3 && ACX_HOST_FLAG_GET(&g_acx->hstat[s2].retired) && !side_branch_finished(g_acx->slot_req[s2])) { cudaEventSynchronize(g_acx->slot_req[s2]->ev_done); break; }
        acx_reclaim_retired();
        for (int i = 0; i < 2000 && g_acx->free_slots.empty(); i++) { usleep(100); acx_reclaim_retired(); }
        if (g_acx->free_slots.empty()) {
            ACX_ERR("out of request slots (%d); raise ACX_NFLAGS / MPIACX_NFLAGS", g_acx->geom.nops);
            return -1;
        }
    }
    int s = g_acx->free_slots.back();
    g_acx->free_slots.pop_back();
    g_acx->slot_busy[s] = graph ? 2 : 1;
    g_acx->slot_req[s] = r;
    g_acx->hstat[s].retired = 0;
    g_acx->slots_in_use++;
    unsigned long long hw = g_acx->st.slots_hw.load();
    if ((unsigned long long)g_acx->slots_in_use > hw) g_acx->st.slots_hw.store(g_acx->slots_in_use);
    return s;
}

int64_t acx_heap_alloc(size_t bytes) {
    if (bytes == 0) bytes = 1;
    bytes = (bytes + 255) & ~(size_t)255;
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    for (int attempt = 0; attempt < 2; attempt++) {
        for (auto it = g_acx->heap_free.begin(); it != g_acx->heap_free.end(); ++it) {
            if (it->second >= bytes) {
                size_t off = it->first, sz = it->second;
                g_acx->heap_free.erase(it);
                if (sz > bytes) g_acx->heap_free[off + bytes] = sz - bytes;
                g_acx->heap_used[off] = bytes;
                g_acx->heap_in_use += bytes;
                if (g_acx->heap_in_use > g_acx->st.heap_hw.load()) g_acx->st.heap_hw.store(g_acx->heap_in_use);
                return (int64_t)off;
            }
        }
        acx_reclaim_retired();  // bounce buffers of retired requests come back here
    }
    // Bounce buffers of in-flight stream operations return when their wait kernels retire: poll briefly.
    for (int i = 0; i < 20000; i++) {
        usleep(100);
        acx_reclaim_retired();
        for (auto it = g_acx->heap_free.begin(); it != g_acx->heap_free.end(); ++it)
            if (it->second >= bytes) { i = 1 << 30; break; }
        if (i >= (1 << 30)) return acx_heap_alloc(bytes);
    }
    ACX_ERR("symmetric heap exhausted (%zu bytes requested, %zu in use of %zu); raise ACX_HEAP_BYTES", bytes, g_acx->heap_in_use,
            g_acx->heap_bytes);
    return -1;
}

void acx_heap_free(int64_t off) {
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    auto it = g_acx->heap_used.find((size_t)off);
    if (it != g_acx->heap_used.end()) {
        size_t o = it->first, sz = it->second;
        g_acx->heap_used.erase(it);
        g_acx->heap_in_use -= sz;
        auto nx = g_acx->heap_free.lower_bound(o);
        if (nx != g_acx->heap_free.end() && o + sz == nx->first) { sz += nx->second; nx = g_acx->heap_free.erase(nx); }
        if (nx != g_acx->heap_free.begin()) {
            auto pv = std::prev(nx);
            if (pv->first + pv->second == o) { pv->second += sz; sz = 0; }
        }
        if (sz) g_acx->heap_free[o] = sz;
    }
}

int64_t acx_heap_offset_of(const void *p) {
    const char *c = (const char *)p;
    if (g_acx->heap && c >= g_acx->heap && c < g_acx->heap + g_acx->heap_bytes) return c - g_acx->heap;
    return -1;
}

uint32_t acx_pflags_alloc(uint32_t n) {
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    if (g_acx->pflags_next + n > g_acx->geom.npflags) { ACX_ERR("out of partition flags (%u); raise ACX_NPFLAGS", g_acx->geom.npflags); return UINT32_MAX; }
    uint32_t r = g_acx->pflags_next;
    g_acx->pflags_next += n;
    return r;
}

// Registrations of pageable host memory are cached (page granular) — and the application may free a buffer and get the same
// addresses back from a later mmap: the cached registration then still maps the OLD, driver-pinned physical pages while the
// CPU sees new ones (sends would transmit stale data, receives land in orphaned pages; the reference hands host buffers to
// MPI and has no such cache). Pinned pages cannot migrate, so a registration is still good exactly when the first page's
// physical frame is the one recorded when it was pinned: /proc/self/pagemap tells (frames are hidden from unprivileged
// processes — then the check is skipped and cached host buffers must outlive the library, or be dropped explicitly with
// MPIX_Buffer_forget).
static uint64_t acx_pfn_of(uintptr_t va) {
#if defined(__linux__) && !defined(ACX_EMU)
    static int fd = -2;
    if (fd == -2) fd = open("/proc/self/pagemap", O_RDONLY | O_CLOEXEC);
    if (fd < 0) return 0;
    uint64_t v = 0;
    if (pread(fd, &v, 8, (off_t)(va / 4096) * 8) != 8) return 0;
    if (!(v >> 63)) return 0;                       // not present
    return v & ((1ull << 55) - 1);                  // 0 when the kernel hides frame numbers
#else
    (void)va;
    return 0;
#endif
}
// Drop cached registrations that no longer describe the memory behind their addresses. Caller holds g_acx->mu.
static void acx_host_regs_validate(uintptr_t lo, uintptr_t hi) {
    for (size_t i = 0; i < g_acx->host_regs.size();) {
        AcxHostReg &r = g_acx->host_regs[i];
        if (r.pfn && r.hi > lo && r.lo < hi) {
            const uint64_t now = acx_pfn_of(r.lo);
            if (now != r.pfn) {
                ACX_LOG(1, "host registration [%#lx, %#lx) went stale (the buffer was unmapped and its addresses reused): re-pinning", (unsigned long)r.lo, (unsigned long)r.hi);
                cudaHostUnregister((void *)r.lo);
                cudaGetLastError();
                g_acx->host_regs.erase(g_acx->host_regs.begin() + (long)i);
                continue;
            }
        }
        i++;
    }
}
extern "C" int MPIX_Buffer_forget(const void *buf, size_t bytes) {
    if (!g_acx || !g_acx->initialized || !g_acx->device) return MPI_SUCCESS;
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    const uintptr_t lo = (uintptr_t)buf & ~(uintptr_t)4095, hi = ((uintptr_t)buf + bytes + 4095) & ~(uintptr_t)4095;
    for (size_t i = 0; i < g_acx->host_regs.size();) {
        AcxHostReg &r = g_acx->host_regs[i];
        if (r.lo >= lo && r.hi <= hi) { cudaHostUnregister((void *)r.lo); cudaGetLastError(); g_acx->host_regs.erase(g_acx->host_regs.begin() + (long)i); }
        else i++;
    }
    return MPI_SUCCESS;
}

int acx_device_alias(const void *p, size_t bytes, char **out) {
    if (!p || bytes == 0) { *out = (char *)p; return 0; }
    if (acx_heap_offset_of(p) >= 0) { *out = (char *)p; return 0; }
    {
        // Host memory this library has already pinned (registrations are permanent): answered from the table, without the
        // cudaPointerGetAttributes driver call — status words and small host buffers come through here on every operation.
        std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
        const uintptr_t lo = (uintptr_t)p, hi = lo + bytes;
        if (!g_acx->host_regs.empty()) acx_host_regs_validate(lo, hi);
        if (!g_acx->host_ptr_is_dev_ptr) {
            for (auto &r : g_acx->host_regs) if (lo >= r.lo && hi <= r.hi) { *out = r.dptr + (lo - r.lo); return 0; }
        } else {
            uintptr_t cur = lo;
            bool progress = true;
            while (cur < hi && progress) {
                progress = false;
                for (auto &r : g_acx->host_regs) if (cur >= r.lo && cur < r.hi) { cur = r.hi; progress = true; break; }
            }
            if (cur >= hi) { *out = (char *)p; return 0; }
        }
    }
    CaptureRelax relax;   // cudaHostRegister below may run inside a user's capture
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); a.type = cudaMemoryTypeUnregistered; }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) { *out = (char *)p; return 0; }
    if (a.type == cudaMemoryTypeHost) {
        // Pinned host memory. If the page `p` starts in is one of OUR page-granular registrations, the buffer may still run
        // into pages nobody has pinned yet (the table lookup above found the range only partly covered): fall through and
        // register the rest. Trusting the attribute here sent kernels into unpinned pages — served by HMM page faults on this
        // platform, racing with the cudaHostRegister of that very page for the next buffer: an intermittent "illegal memory
        // access" on hardware (tests/c/status_errors.c, 1-KiB rows of a static array as message buffers).
        bool ours = false;
        {
            std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
            for (auto &r : g_acx->host_regs) if ((uintptr_t)p >= r.lo && (uintptr_t)p < r.hi) { ours = true; break; }
        }
        if (!ours) { *out = (char *)a.devicePointer; return 0; }   // the application's own pinned allocation (cudaMallocHost / cudaHostRegister): pinned as a whole
    }
    // pageable host memory: pin + map the covering pages. Registrations are never undone while the
    // library is up (device aliases are baked into persistent requests and graphs), so a new range that
    // partially overlaps earlier ones only registers the pages that are still uncovered.
    const uintptr_t page = 4096;
    uintptr_t lo = (uintptr_t)p & ~(page - 1), hi = ((uintptr_t)p + bytes + page - 1) & ~(page - 1);
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    if (!g_acx->host_ptr_is_dev_ptr) {
        // no unified addressing for registered memory: one registration must cover the buffer
        for (auto &r : g_acx->host_regs)
            if (lo >= r.lo && hi <= r.hi) { *out = r.dptr + ((uintptr_t)p - r.lo); return 0; }
        // partly covered by earlier registrations: replace them by one registration of the union (their aliases die with
        // them — acceptable on the rare platform without unified addressing of registered memory; nothing of theirs can be
        // in flight for THIS buffer, and other buffers on those pages are re-aliased on their next use)
        for (size_t i = 0; i < g_acx->host_regs.size();) {
            AcxHostReg &r = g_acx->host_regs[i];
            if (r.hi > lo && r.lo < hi) {
                lo = std::min(lo, r.lo); hi = std::max(hi, r.hi);
                cudaHostUnregister((void *)r.lo); cudaGetLastError();
                g_acx->host_regs.erase(g_acx->host_regs.begin() + (long)i);
                i = 0;   // the union grew: re-scan
                continue;
            }
            i++;
        }
        e = cudaHostRegister((void *)lo, hi - lo, cudaHostRegisterMapped | cudaHostRegisterPortable);
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (g_acx->pageable_access) { *out = (char *)p; return 0; }
            ACX_ERR("cannot make host buffer %p (+%zu) device-accessible: %s", p, bytes, cudaGetErrorString(e));
            return MPI_ERR_BUFFER;
        }
        char *d = nullptr;
        cudaHostGetDevicePointer((void **)&d, (void *)lo, 0);
        g_acx->host_regs.push_back({lo, hi, d, acx_pfn_of(lo)});
        *out = d + ((uintptr_t)p - lo);
        return 0;
    }
    // unified addressing: device alias == host pointer once every page is registered
    uintptr_t cur = lo;
    while (cur < hi) {
        // skip pages already covered
        bool covered = false;
        for (auto &r : g_acx->host_regs) if (cur >= r.lo && cur < r.hi) { cur = r.hi; covered = true; break; }
        if (covered) continue;
        uintptr_t end = hi;
        for (auto &r : g_acx->host_regs) if (r.lo > cur && r.lo < end) end = r.lo;
        e = cudaHostRegister((void *)cur, end - cur, cudaHostRegisterMapped | cudaHostRegisterPortable);
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (g_acx->pageable_access) {   // HMM/ATS: the GPU may be able to touch it as is
                static bool warned = false;
                if (!warned) { warned = true; ACX_LOG(0, "cudaHostRegister(%p, +%zu) failed (%s): relying on pageable-memory access for this buffer", (void *)cur, (size_t)(end - cur), cudaGetErrorString(e)); }
                *out = (char *)p; return 0;
            }
            ACX_ERR("cannot make host buffer %p (+%zu) device-accessible: %s", p, bytes, cudaGetErrorString(e));
            return MPI_ERR_BUFFER;
        }
        g_acx->host_regs.push_back({cur, end, (char *)cur, acx_pfn_of(cur)});
        cur = end;
    }
    *out = (char *)p;
    return 0;
}

int acx_comm_ctx(MPI_Comm comm, int *ctx, const AcxCommMap **map) {
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    for (size_t i = 0; i < g_acx->comms.size(); i++)
        if (g_acx->comms[i].comm == comm) { *ctx = g_acx->comms[i].ctx; if (map) *map = &g_acx->comms[i]; return 0; }
    AcxCommMap m;
    m.comm = comm;
    // The context id feeds the channel hash and the partitioned handshake tag, so it must be the SAME number on
    // every member. The bundled mini-MPI agrees on communicator handles collectively (identical on all members),
    // so the handle itself is used; with another MPI the order of first use is the only portable stand-in (ranks
    // must then first-use their communicators in the same order).
#ifdef ACX_MINI_MPI
    m.ctx = (int)comm - (int)MPI_COMM_WORLD;
#else
    m.ctx = (int)g_acx->comms.size();
#endif
    MPI_Group gc, gw;
    if (MPI_Comm_size(comm, &m.size) != MPI_SUCCESS || MPI_Comm_rank(comm, &m.me) != MPI_SUCCESS || MPI_Comm_group(comm, &gc) != MPI_SUCCESS) {
        ACX_ERR("invalid communicator");
        return MPI_ERR_COMM;
    }
    MPI_Comm_group(MPI_COMM_WORLD, &gw);
    std::vector<int> mine((size_t)m.size);
    for (int i = 0; i < m.size; i++) mine[i] = i;
    m.to_world.assign((size_t)m.size, -1);
    int rc = MPI_Group_translate_ranks(gc, m.size, mine.data(), gw, m.to_world.data());
    MPI_Group_free(&gc); MPI_Group_free(&gw);
    if (rc != MPI_SUCCESS) { ACX_ERR("cannot translate the communicator's ranks to MPI_COMM_WORLD"); return MPI_ERR_COMM; }
    m.identity = m.size == g_acx->size;
    for (int i = 0; i < m.size; i++) {
        if (m.to_world[i] < 0 || m.to_world[i] >= g_acx->size) { ACX_ERR("communicator has a member outside MPI_COMM_WORLD"); return MPI_ERR_COMM; }
        if (m.to_world[i] != i) m.identity = false;
    }
    g_acx->comms.reserve(64);   // callers keep pointers into this vector
    if (g_acx->comms.size() >= 64 && g_acx->comms.size() == g_acx->comms.capacity()) { ACX_ERR("too many communicators"); return MPI_ERR_COMM; }
    g_acx->comms.push_back(m);
    *ctx = m.ctx;
    if (map) *map = &g_acx->comms.back();
    return 0;
}

// =======================================================================================
// registration cache: zero-copy for plain cudaMalloc buffers (incl. torch tensors)
// =======================================================================================
// The owner publishes {buffer id, base, size, CUDA-IPC handle} in the job's shared scratch; every
// rank maps new entries of its peers during its own MPIX calls, writes the mapped base into its
// device-side xlat table and acks. The owner switches a buffer from the bounce path to the
// registered (zero-copy) path for a given peer once that peer has acked — nobody ever waits for a
// mapping, so there is no progress thread and no deadlock behind a spinning kernel.
#ifdef ACX_MINI_MPI
struct AcxRegEntry { uint64_t buffer_id, base, size; cudaIpcMemHandle_t handle; uint64_t pad; };
struct AcxRegTable {
    std::atomic<uint32_t> count;
    uint32_t pad[15];
    AcxRegEntry e[64];
    std::atomic<uint8_t> ack[64][64];   // [reader rank][entry]
};
static bool g_reg_on = false;
static AcxRegTable *g_reg_tables = nullptr;
static std::unordered_map<uint64_t, int> g_reg_mine;     // buffer id -> entry index (-1: not registrable)
static std::vector<uint32_t> g_reg_seen;
static std::vector<void *> g_reg_opened;
static std::vector<uint64_t> g_xlat_host;   // host copy of the device translation table: [owner rank][entry] -> address mapped here (copy-engine sends resolve destinations on the host)
typedef int (*cuMemGetAddressRangeFn)(unsigned long long *, size_t *, unsigned long long);
typedef int (*cuPointerGetAttributeFn)(void *, int, unsigned long long);
static cuMemGetAddressRangeFn g_cuMemGetAddressRange = nullptr;
static cuPointerGetAttributeFn g_cuPointerGetAttribute = nullptr;

static void reg_init() {
    long on = 1;
    acx_env_int("ACX_REGCACHE", nullptr, 1, &on);
    size_t sb = 0;
    void *scratch = ACX_Mpi_shared_scratch(&sb);
    if (!on || !scratch || g_acx->size > 64 || sb < sizeof(AcxRegTable) * (size_t)g_acx->size) return;
    void *f1 = nullptr, *f2 = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f1, cudaEnableDefault, &q) != cudaSuccess || !f1) return;
    if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &f2, cudaEnableDefault, &q) != cudaSuccess || !f2) return;
    g_cuMemGetAddressRange = (cuMemGetAddressRangeFn)f1;
    g_cuPointerGetAttribute = (cuPointerGetAttributeFn)f2;
    g_reg_tables = (AcxRegTable *)scratch;   // zero-initialised by the job's rank 0 before anyone runs
    {   // a re-initialization starts from an empty table (MPIX_Init ends with a barrier: nobody reads mine before that)
        AcxRegTable *mine = g_reg_tables + g_acx->rank;
        mine->count.store(0, std::memory_order_relaxed);
        for (auto &row : mine->ack) for (auto &a : row) a.store(0, std::memory_order_relaxed);
    }
    g_reg_seen.assign(g_acx->size, 0);
    g_xlat_host.assign((size_t)g_acx->size * g_acx->geom.maxreg, 0);
    g_reg_on = true;
}

// Map registrations our peers have published since the last call; ack them. Cheap when idle.
static void reg_progress() {
    if (!g_reg_on) return;
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);   // application threads enqueue concurrently; new entries are rare
    CaptureRelax relax;   // cudaIpcOpenMemHandle + a synchronous 8-byte copy: legal inside a user's capture only in relaxed mode
    bool wrote = false;
    std::vector<std::pair<int, uint32_t>> to_ack;
    for (int p = 0; p < g_acx->size; p++) {
        AcxRegTable *t = g_reg_tables + p;
        const uint32_t c = t->count.load(std::memory_order_acquire);
        for (uint32_t i = g_reg_seen[p]; i < c && i < g_acx->geom.maxreg; i++) {
            uint64_t mapped = 0;
            if (p == g_acx->rank) mapped = t->e[i].base;
            else {
                void *m = nullptr;
                if (cudaIpcOpenMemHandle(&m, t->e[i].handle, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) { mapped = (uint64_t)m; g_reg_opened.push_back(m); }
                else cudaGetLastError();
            }
            if (mapped) {
                acx_h2d(g_acx->ctrl + g_acx->geom.off_xlat + ((uint64_t)p * g_acx->geom.maxreg + i) * 8, &mapped, 8, host_stream());   // entries are rare
                g_xlat_host[(size_t)p * g_acx->geom.maxreg + i] = mapped;
                to_ack.push_back({p, i});
                wrote = true;
            }
        }
        g_reg_seen[p] = c;
    }
    if (wrote) for (auto &a : to_ack) g_reg_tables[a.first].ack[g_acx->rank][a.second].store(1, std::memory_order_release);
}

// Location code (ACX_OFF_REG_BIT set) of [p, p+bytes) if it lies in a registered allocation that `peer`
// has mapped; ~0 otherwise (caller bounces). Publishes the allocation on first sight.
static uint64_t reg_encode(const void *p, size_t bytes, int peer) {
    if (!g_reg_on) return ~0ull;
    CaptureRelax relax;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return ~0ull; }
    if (a.type != cudaMemoryTypeDevice || a.device != g_acx->dev) return ~0ull;
    unsigned long long id = 0;
    if (g_cuPointerGetAttribute(&id, 7 /*CU_POINTER_ATTRIBUTE_BUFFER_ID*/, (unsigned long long)p) != 0) return ~0ull;
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    AcxRegTable *mine = g_reg_tables + g_acx->rank;
    auto it = g_reg_mine.find(id);
    int idx;
    if (it != g_reg_mine.end()) idx = it->second;
    else {
        idx = -1;
        unsigned long long base = 0; size_t size = 0;
        const uint32_t c = mine->count.load(std::memory_order_relaxed);
        cudaIpcMemHandle_t h;
        if (c < g_acx->geom.maxreg && g_cuMemGetAddressRange(&base, &size, (unsigned long long)p) == 0 &&
            cudaIpcGetMemHandle(&h, (void *)base) == cudaSuccess) {
            mine->e[c].buffer_id = id; mine->e[c].base = base; mine->e[c].size = size; mine->e[c].handle = h;
            mine->count.store(c + 1, std::memory_order_release);
            idx = (int)c;
            ACX_LOG(2, "registered allocation %p (+%zu) as #%d", (void *)base, size, idx);
        } else cudaGetLastError();
        g_reg_mine[id] = idx;
    }
    if (idx < 0) return ~0ull;
    const AcxRegEntry &e = mine->e[idx];
    if ((uint64_t)p < e.base || (uint64_t)p + bytes > e.base + e.size) return ~0ull;
    if (!mine->ack[peer][idx].load(std::memory_order_acquire)) return ~0ull;   // peer has not mapped it yet
    return ACX_OFF_REG_BIT | ((uint64_t)idx << 48) | ((uint64_t)p - e.base);
}
static void reg_finalize() {
    for (void *m : g_reg_opened) cudaIpcCloseMemHandle(m);
    g_reg_opened.clear(); g_reg_mine.clear(); g_reg_on = false;
}
#else
static void reg_init() {}
static void reg_progress() {}
static uint64_t reg_encode(const void *, size_t, int) { return ~0ull; }
static void reg_finalize() {}
#endif

// =======================================================================================
// stream memory operations (ACX_MEMOPS=1)
// =======================================================================================
// Reference: the memOps shim wrapWriteValue / wrapWaitValue / wrapBatchMemOp (src/sendrecv.cu:34-42) pokes and polls flags in
// HOST memory for the CPU proxy. Point-to-point operations here cannot use it — a message's ring slot and generation come
// from a ticket the trigger kernel takes on the device, so the host has no value to wait for — but the partitioned
// protocol runs on epochs the HOST counts: MPIX_Start is "write epoch e to two words", completing a round is "wait until
// sent[p] / arrived[p] >= e for every partition". Those are exactly cuStreamWriteValue64 / cuStreamWaitValue64(GEQ), on
// flags in device HBM, and a whole MPIX_Startall / MPIX_Waitall group is ONE cuStreamBatchMemOp call: no kernel launch
// and no SM held by a spinning CTA. Opt-in until measured on hardware (tools/gpu_next_round.sh). A memOp wait has
// no watchdog, and receives that land in a private buffer still need the copy-out kernel.
static int g_memops = 0;
typedef int (*cuStreamBatchMemOpFn)(CUstream, unsigned int, CUstreamBatchMemOpParams *, unsigned int);
static cuStreamBatchMemOpFn g_cuStreamBatchMemOp = nullptr;
static const size_t kMemOpBatchMax = 256;   // per driver call

static long g_ce_memop = 1;                // ACX_CE_MEMOP: copy-engine sends publish completion with cuStreamWriteValue64 instead of a kernel (1), and also wait for the receive's mirror with cuStreamWaitValue64 instead of the gate kernel (2)
static void memops_init() {
    if (!g_memops && !g_ce_memop) return;
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamBatchMemOp", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
        cudaGetLastError();
        ACX_LOG(1, "ACX_MEMOPS: cuStreamBatchMemOp is not available; using kernels");
        g_memops = 0; g_ce_memop = 0;
        return;
    }
    g_cuStreamBatchMemOp = (cuStreamBatchMemOpFn)f;
}
struct MemOpBatch {
    std::vector<CUstreamBatchMemOpParams> ops;
    void write64(uint64_t *addr, uint64_t v) {
        CUstreamBatchMemOpParams p;
        memset(&p, 0, sizeof p);
        p.writeValue.operation = CU_STREAM_MEM_OP_WRITE_VALUE_64;
        p.writeValue.address = (CUdeviceptr)(uintptr_t)addr;
        p.writeValue.value64 = v;
        p.writeValue.flags = CU_STREAM_WRITE_VALUE_DEFAULT;   // default = memory barrier in front of the write (release)
        ops.push_back(p);
    }
    void wait_geq64(const uint64_t *addr, uint64_t v) {
        CUstreamBatchMemOpParams p;
        memset(&p, 0, sizeof p);
        p.waitValue.operation = CU_STREAM_MEM_OP_WAIT_VALUE_64;
        p.waitValue.address = (CUdeviceptr)(uintptr_t)addr;
        p.waitValue.value64 = v;
        p.waitValue.flags = CU_STREAM_WAIT_VALUE_GEQ;
        ops.push_back(p);
    }
    int flush(cudaStream_t stream) {   // operations of one call execute in array order
        for (size_t i = 0; i < ops.size(); i += kMemOpBatchMax) {
            const unsigned n = (unsigned)std::min(kMemOpBatchMax, ops.size() - i);
            const int rc = g_cuStreamBatchMemOp((CUstream)stream, n, ops.data() + i, 0);
            if (rc != 0) { ACX_ERR("cuStreamBatchMemOp failed (CUresult %d)", rc); return MPI_ERR_OTHER; }
            g_acx->st.kernels++;   // counted with the launches: one enqueued device-side operation group
        }
        ops.clear();
        return 0;
    }
};
static bool stream_is_capturing(cudaStream_t s) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
    return cs == cudaStreamCaptureStatusActive;
}

// =======================================================================================
// init / finalize
// =======================================================================================
struct Bootstrap {
    cudaIpcMemHandle_t ctrl, heap;
    int dev, pid;
    unsigned long long heap_bytes, ctrl_bytes;
    char host[64];
};

static void make_geom(acx_geom &g, int n) {
    long v;
    g.nranks = n;
    acx_env_int("ACX_NCHAN", nullptr, 32, &v); g.nchan = (int)std::max(1L, v);   // one channel per (communicator, tag) a pair of ranks uses (see chan_lookup); the last one is the overflow FIFO
    acx_env_int("ACX_DEPTH", nullptr, 32, &v); g.depth = (int)std::max(2L, v);
    acx_env_int("ACX_NFLAGS", "MPIACX_NFLAGS", 4096, &v); g.nops = (int)std::max(8L, v);
    acx_env_int("ACX_EAGER_MAX", nullptr, 8192, &v); g.eager_max = (uint32_t)((std::max(16L, v) + 15) & ~15L);
    acx_env_int("ACX_NPFLAGS", nullptr, 1 << 16, &v); g.npflags = (uint32_t)v;
    acx_env_int("ACX_TRACE_RECORDS", nullptr, 1 << 14, &v);
    uint32_t nt = 1; while (nt < (uint32_t)v) nt <<= 1; g.ntrace = nt;
    acx_env_int("ACX_MAXREG", nullptr, 64, &v); g.maxreg = (uint32_t)std::min(std::max(1L, v), 64L);
    acx_geom_layout(g);
}


// ---------------------------------------------------------------------------------------
// NUMA placement. On an HGX B200 node GPUs 0-3 hang off socket 0 and 4-7 off socket 1; a rank whose staging memory
// or launching thread sits on the other socket pays the inter-socket link on every H2D copy and every launch
// (round 1: 8-rank end-to-end step 9.1 ms against 5 ms of PCIe time). mpiexec-style launchers bind ranks; torchrun
// does not, so MPIX_Init does it itself: the calling thread is restricted to the CPUs of the GPU's NUMA node and its
// memory policy prefers that node (threads and pinned allocations made afterwards inherit both). ACX_NUMA_BIND=0
// turns it off; a process whose affinity was already narrowed by its launcher is left alone.
// ---------------------------------------------------------------------------------------
static int acx_numa_bind(int dev) {
#if defined(__linux__) && !defined(ACX_EMU)
    long on = 1;
    acx_env_int("ACX_NUMA_BIND", nullptr, 1, &on);
    if (!on) return -1;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char *c = bus; *c; c++) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    int node = -1;
    if (f) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    if (node < 0) return -1;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return -1;
    char list[4096] = {0};
    if (!fgets(list, sizeof list, f)) list[0] = 0;
    fclose(f);
    cpu_set_t want, have, both;
    CPU_ZERO(&want);
    for (char *tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b && c < CPU_SETSIZE; c++) CPU_SET(c, &want); }
        else if (sscanf(tok, "%d", &a) == 1 && a < CPU_SETSIZE) CPU_SET(a, &want);
    }
    if (sched_getaffinity(0, sizeof have, &have) != 0) return -1;
    CPU_AND(&both, &want, &have);
    if (CPU_COUNT(&both) == 0) return -1;                        // the launcher confined us elsewhere: respect it
    if (CPU_COUNT(&both) != CPU_COUNT(&have)) sched_setaffinity(0, sizeof both, &both);
    // set_mempolicy(MPOL_PREFERRED, {node}): no libnuma in the image, so the raw system call
    unsigned long mask[16] = {0};
    if (node < (int)(sizeof mask * 8)) {
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof mask * 8);
    }
    return node;
#else
    (void)dev;
    return -1;
#endif
}

static void ce_init();       // enqueue-time receive records (copy-engine rendezvous), defined with the front-ends
static void tagdir_init();   // channel directory (exact tag matching), defined with the front-ends
// Reference: MPIX_Init, src/init.cpp:157-253 (flag table in mapped host memory + proxy thread). Here: device windows,
// CUDA-IPC exchange, no thread.
extern "C" int MPIX_Init(void) {
    int initialized = 0, finalized = 0, thread_level = 0;
    MPI_Initialized(&initialized);
    MPI_Finalized(&finalized);
    if (!initialized || finalized) {
        fprintf(stderr, "[acx] MPIX_Init: MPI is not initialized (%d) or already finalized (%d)\n", initialized, finalized);
        return 1;
    }
    MPI_Query_thread(&thread_level);
    if (g_acx && g_acx->initialized) return 0;
    g_acx = new AcxState();
    MPI_Comm_rank(MPI_COMM_WORLD, &g_acx->rank);
    MPI_Comm_size(MPI_COMM_WORLD, &g_acx->size);
    long v;
    acx_env_int("ACX_VERBOSE", nullptr, 0, &v); g_acx->verbose = (int)v;
    acx_env_int("ACX_TRACE", nullptr, 0, &v); g_acx->trace = v != 0;
    acx_env_int("ACX_WATCHDOG_MS", nullptr, 0, &v); g_acx->watchdog_ms = v;
    acx_env_int("ACX_NVTX", nullptr, 0, &v); g_nvtx = v != 0;
    if (getenv("MPIACX_DISABLE_MEMOPS")) ACX_LOG(1, "MPIACX_DISABLE_MEMOPS: stream memOps stay off (they are opt-in here: ACX_MEMOPS=1, partitioned Start / Wait only)");
    { const char *m = getenv("ACX_MOVER"); g_mover_tma = m && !strcmp(m, "tma"); g_mover_256 = !m ? 0 : !strcmp(m, "ldst256") ? 1 : !strcmp(m, "ldstcs") ? 2 : 0; }   // copy-loop flavour, see copy_flavour()
#ifdef ACX_EMU
    g_mover_tma = 0;
#endif
    acx_env_int("ACX_PDL", nullptr, 1, &v); g_pdl = v != 0;
    acx_env_int("ACX_SEND_OVERFLOW", nullptr, 1, &v); g_send_overflow = v != 0;
    acx_env_int("ACX_SEND_OVERFLOW_MAX", nullptr, 128, &v); g_send_overflow_max = v < 0 ? 0 : v;
    acx_env_int("ACX_MOVER_BLOCK", nullptr, 0, &v); g_mover_block = (int)std::max(0L, v) & ~511;
    acx_env_int("ACX_FUSE_RNDV", nullptr, 1, &v); g_fuse_rndv = v != 0;
    acx_env_int("ACX_SIDE_STREAMS", nullptr, 1, &v); g_side_on = v != 0;
    acx_env_int("ACX_WILDCARDS", nullptr, 0, &v); g_wild = v != 0;   // experimental device-side wildcard matching (see recv_any_kernel)
    acx_env_int("ACX_COOP", nullptr, 0, &v); g_coop = v != 0;   // opt-in until measured on hardware
    acx_env_int("ACX_COOP_MIN_BYTES", nullptr, 4 << 20, &v); g_coop_min = std::max(65536L, v);
    acx_env_int("ACX_LATE_PUSH_MIN", nullptr, 4 << 20, &v); g_late_min = std::max(0L, v);   // must be the same on every rank (like every protocol knob)
    acx_env_int("ACX_LATE_FALLBACK_US", nullptr, 50, &v); g_late_fallback_us = std::max(0L, v);
    acx_env_int("ACX_LATE_HELPERS", nullptr, 0, &v); g_late_helpers = std::max(0L, v);
    acx_env_int("ACX_PBATCH", nullptr, 1, &v); g_pbatch = v != 0;   // one launch per MPIX_Startall / MPIX_Waitall group: measured 0.641 vs 0.776 ms per API-level halo step on 2x B200 (reference arm 0.700), profiles/raw/r2_call1/halo_api.txt
    acx_env_int("ACX_CE_MIN", nullptr, 64 << 20, &v); g_ce_min = std::max(0L, v);
    acx_env_int("ACX_WAIT_EARLY_HELPERS", nullptr, 0, &v); g_wait_early_helpers = std::max(0L, v);
    acx_env_int("ACX_CE_MEMOP", nullptr, 1, &v); g_ce_memop = getenv("MPIACX_DISABLE_MEMOPS") ? 0 : std::max(0L, v);
    acx_env_int("ACX_CE_SPINGATE", nullptr, 1, &v); g_ce_spingate = v != 0;
    acx_env_int("ACX_CE_SPIN_US", nullptr, 40, &v); g_ce_spin_us = std::max(0L, v);
    acx_env_int("ACX_CE_SPIN_BUSY_US", nullptr, 300, &v); g_ce_spin_busy_us = std::max(0L, v);
    acx_env_int("ACX_MEMOPS", nullptr, 0, &v); g_memops = v != 0 && !getenv("MPIACX_DISABLE_MEMOPS");   // partitioned Start / Wait as stream memOps (opt-in until measured)
    acx_env_int("ACX_PEER_BRANCH", nullptr, 2, &v); g_branch = (int)std::max(0L, std::min(2L, v));   // 8x B200 all-pairs under CUDA graphs, 1 MiB per pair: 79.6 GB/s per GPU with branches vs 22.3 without (reference 45.8); 16 MiB: 288.6 vs 290.4 (profiles/raw/r2_call9)
    acx_env_int("ACX_SENDER_PUSH", nullptr, 0, &v); g_apush = v != 0;   // opt-in: a spinning gate can deadlock against device-synchronizing host calls
    acx_env_int("ACX_FUSE_MAX_CTAS", nullptr, 32, &v); g_fuse_max_ctas = (int)v;
    acx_env_int("ACX_FUSE_MIN_CTAS", nullptr, 8, &v); g_fuse_min_ctas = (int)v;
    acx_env_int("ACX_FAULT_DROP_SEND", nullptr, 0, &v); g_fault_drop_send = v;
    acx_env_int("ACX_FAULT_DELAY_US", nullptr, 0, &v); g_fault_delay_us = std::max(0L, v);
    acx_env_int("ACX_MOVER_MAX_CTAS", nullptr, 0, &v); g_acx->mover_max_ctas = (int)v;
    acx_env_int("ACX_MOVER_CTA_BYTES", nullptr, 64 << 10, &v); g_acx->mover_min_cta_bytes = (size_t)std::max(4096L, v);
    { int c0; if (acx_comm_ctx(MPI_COMM_WORLD, &c0)) return 1; }
    if (MPI_Comm_dup(MPI_COMM_WORLD, &g_acx->icomm) != MPI_SUCCESS) { ACX_ERR("MPI_Comm_dup failed"); return 1; }

    int ndev = 0;
    long force_host = 0;
    acx_env_int("ACX_FORCE_HOST", nullptr, 0, &force_host);
    if (force_host || cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        // All ranks must agree on the transport.
        int have = 0, all = 0;
        MPI_Allreduce(&have, &all, 1, MPI_INT, MPI_MIN, MPI_COMM_WORLD);
        g_acx->device = false;
        g_acx->initialized = true;
        ACX_LOG(1, "no CUDA device: host transport (MPI p2p, ops trigger at enqueue time)");
        return 0;
    }
    {
        int have = 1, all = 0;
        MPI_Allreduce(&have, &all, 1, MPI_INT, MPI_MIN, MPI_COMM_WORLD);
        if (!all) { g_acx->device = false; g_acx->initialized = true; return 0; }
    }
    long want_dev;
    if (acx_env_int("ACX_DEVICE", nullptr, -1, &want_dev) && want_dev >= 0) {
        if (cudaSetDevice((int)(want_dev % ndev)) != cudaSuccess) { ACX_ERR("cudaSetDevice(%ld) failed", want_dev); return 1; }
    }
    if (cudaGetDevice(&g_acx->dev) != cudaSuccess) { ACX_ERR("cudaGetDevice failed"); return 1; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, g_acx->dev) != cudaSuccess) { ACX_ERR("cudaGetDeviceProperties failed"); return 1; }
    g_acx->sm_count = prop.multiProcessorCount;
    { const int node = acx_numa_bind(g_acx->dev); if (node >= 0) ACX_LOG(1, "bound to NUMA node %d (GPU %d's socket)", node, g_acx->dev); }
    cudaDeviceGetAttribute(&g_acx->pageable_access, cudaDevAttrPageableMemoryAccess, g_acx->dev);
    cudaDeviceGetAttribute(&g_acx->host_ptr_is_dev_ptr, cudaDevAttrCanUseHostPointerForRegisteredMem, g_acx->dev);

    make_geom(g_acx->geom, g_acx->size);
    acx_env_int("ACX_HEAP_BYTES", nullptr, 1024L << 20, &v);   // 1 GiB of 180: the fused halo op keeps its two 512^3 fields (with ghost shell) here
    g_acx->heap_bytes = ((size_t)v + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    cudaError_t e;
    if ((e = cudaMalloc((void **)&g_acx->ctrl, g_acx->geom.bytes)) != cudaSuccess || (e = cudaMalloc((void **)&g_acx->heap, g_acx->heap_bytes)) != cudaSuccess) {
        ACX_ERR("cudaMalloc of windows (%llu + %zu bytes) failed: %s", (unsigned long long)g_acx->geom.bytes, g_acx->heap_bytes, cudaGetErrorString(e));
        return 1;
    }
    cudaMemset(g_acx->ctrl, 0, g_acx->geom.bytes);
    cudaMalloc((void **)&g_trace_cursor, 8);
    cudaMemset(g_trace_cursor, 0, 8);
    cudaDeviceSynchronize();
    g_acx->heap_free[0] = g_acx->heap_bytes;

    // ---- exchange IPC handles, map every peer's windows ----
    Bootstrap mine;
    memset(&mine, 0, sizeof mine);
    if (g_acx->size > 1) {
        if ((e = cudaIpcGetMemHandle(&mine.ctrl, g_acx->ctrl)) != cudaSuccess || (e = cudaIpcGetMemHandle(&mine.heap, g_acx->heap)) != cudaSuccess) {
            ACX_ERR("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
            return 1;
        }
    }
    mine.dev = g_acx->dev; mine.pid = (int)getpid(); mine.heap_bytes = g_acx->heap_bytes; mine.ctrl_bytes = g_acx->geom.bytes;
    gethostname(mine.host, sizeof mine.host - 1);
    std::vector<Bootstrap> all(g_acx->size);
    MPI_Allgather(&mine, (int)sizeof(Bootstrap), MPI_BYTE, all.data(), (int)sizeof(Bootstrap), MPI_BYTE, MPI_COMM_WORLD);
    g_acx->peer_ctrl.assign(g_acx->size, nullptr);
    g_acx->peer_heap.assign(g_acx->size, nullptr);
    for (int p = 0; p < g_acx->size; p++) {
        if (p == g_acx->rank) { g_acx->peer_ctrl[p] = g_acx->ctrl; g_acx->peer_heap[p] = g_acx->heap; continue; }
        if (all[p].heap_bytes != g_acx->heap_bytes || all[p].ctrl_bytes != g_acx->geom.bytes) { ACX_ERR("rank %d has a different window geometry (env mismatch)", p); return 1; }
        if (strncmp(all[p].host, mine.host, sizeof mine.host) != 0) { ACX_ERR("rank %d is on another node (%s); multi-node is not supported", p, all[p].host); return 1; }
        if (all[p].dev != g_acx->dev) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g_acx->dev, all[p].dev);
            if (!can) { ACX_ERR("device %d cannot access peer device %d", g_acx->dev, all[p].dev); return 1; }
        }
        if ((e = cudaIpcOpenMemHandle((void **)&g_acx->peer_ctrl[p], all[p].ctrl, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess ||
            (e = cudaIpcOpenMemHandle((void **)&g_acx->peer_heap[p], all[p].heap, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) {
            ACX_ERR("cudaIpcOpenMemHandle(rank %d) failed: %s", p, cudaGetErrorString(e));
            return 1;
        }
    }
    // ---- host-side tables ----
    if ((e = cudaHostAlloc((void **)&g_acx->hstat, sizeof(acx_host_status) * g_acx->geom.nops, cudaHostAllocMapped)) != cudaSuccess ||
        (e = cudaHostAlloc((void **)&g_diag, 64, cudaHostAllocMapped)) != cudaSuccess) {
        ACX_ERR("cudaHostAlloc failed: %s", cudaGetErrorString(e));
        return 1;
    }
    memset(g_acx->hstat, 0, sizeof(acx_host_status) * g_acx->geom.nops);
    memset(g_diag, 0, 64);
    cudaHostGetDevicePointer((void **)&g_acx->hstat_d, g_acx->hstat, 0);
    cudaHostGetDevicePointer((void **)&g_diag_d, g_diag, 0);
    cudaStreamCreateWithFlags(&g_acx->istream, cudaStreamNonBlocking);
    { std::lock_guard<std::mutex> lk(g_host_streams_mu); g_host_streams.clear(); g_init_generation++; }
    host_stream();   // the initializing thread takes the first one
    g_side.assign(8, nullptr);
    for (auto &ss : g_side) cudaStreamCreateWithFlags(&ss, cudaStreamNonBlocking);
    g_acx->slot_busy.assign(g_acx->geom.nops, 0);
    g_acx->slot_req.assign(g_acx->geom.nops, nullptr);
    for (int s = g_acx->geom.nops - 1; s >= 0; s--) g_acx->free_slots.push_back(s);
    g_acx->recv_outstanding.assign(g_acx->size, std::vector<int>(g_acx->geom.nchan, 0));
    g_acx->recv_deferred.assign(g_acx->size, std::vector<std::vector<AcxRequest *>>(g_acx->geom.nchan));
    g_acx->send_count.assign(g_acx->size, std::vector<AcxState::ChanCount>(g_acx->geom.nchan));
    g_acx->send_flow.assign(g_acx->size, std::vector<AcxState::SendFlow>(g_acx->geom.nchan));
    g_acx->recv_count.assign(g_acx->size, std::vector<AcxState::ChanCount>(g_acx->geom.nchan));
    g_acx->device = true;
    g_acx->initialized = true;
    reg_init();
    tagdir_init();
    ce_init();
    memops_init();
#ifdef ACX_MINI_MPI
    ACX_Mpi_set_abort_hook(report_watchdog);
#endif
    MPI_Barrier(MPI_COMM_WORLD);
    ACX_LOG(1, "device transport up: dev %d (%s, %d SMs), %d ranks, ctrl %.1f MiB, heap %.0f MiB, eager<=%u, depth %d, chans %d", g_acx->dev, prop.name,
            g_acx->sm_count, g_acx->size, g_acx->geom.bytes / 1048576.0, g_acx->heap_bytes / 1048576.0, g_acx->geom.eager_max, g_acx->geom.depth, g_acx->geom.nchan);
    return 0;
}

// Reference: MPIX_Finalize, src/init.cpp:255-275 (join proxy, warn on leaked slots).
extern "C" int MPIX_Finalize(void) {
    if (!g_acx || !g_acx->initialized) return 1;
    int err = 0;
    if (g_acx->device) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { report_watchdog(); ACX_ERR("device error at finalize: %s", cudaGetErrorString(e)); err = 1; }
        MPI_Barrier(MPI_COMM_WORLD);  // nobody unmaps while a peer may still store into us
        {
            std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
            acx_reclaim_retired();
            // The destructors of graph user objects may still be on their way (CUDA runs them on an internal thread some time
            // after cudaGraphDestroy / cudaGraphExecDestroy return): give them a moment before calling a slot leaked.
            for (int i = 0; i < 1000 && g_graph_objs_live.load() > 0; i++) { usleep(1000); acx_reclaim_retired(); }   // (up to a second: on a loaded host the software device's destructor thread was seen to take longer than 200 ms — a LeakSanitizer report at exit)
            int leaked = 0;
            for (int s = 0; s < g_acx->geom.nops; s++) if (g_acx->slot_busy[s] == 1 || g_acx->slot_busy[s] == 2) leaked++;
            if (leaked) printf("Warning: %d request slot(s) still in use at MPIX_Finalize (un-waited operations or live graphs)\n", leaked);
        }
        if (g_acx->trace) {
            char path[256];
            snprintf(path, sizeof path, "%s/acx_trace_rank%d.jsonl", getenv("ACX_TRACE_DIR") ? getenv("ACX_TRACE_DIR") : ".", g_acx->rank);
            MPIX_Trace_dump(path);
        }
        if (g_acx->verbose >= 1) {
            MPIX_Acx_stats s;
            MPIX_Get_stats(&s);
            fprintf(stderr, "[acx %d] stats: sends %llu recvs %llu waits %llu | bytes tx %llu rx(cap) %llu | eager %llu rndv %llu bounce %llu | kernels %llu graph-nodes %llu | slots hw %llu heap hw %llu | copy-engine %llu overflow-sends %llu\n",
                    g_acx->rank, s.sends_enqueued, s.recvs_enqueued, s.waits_enqueued, s.bytes_sent, s.bytes_received, s.eager_msgs, s.rendezvous_msgs, s.bounce_copies,
                    s.kernels_launched, s.graph_nodes_added, s.slots_high_water, s.heap_bytes_high_water, s.copy_engine_msgs, s.overflow_sends);
        }
        reg_finalize();
        g_pending_retire.clear();
        for (int p = 0; p < g_acx->size; p++) {
            if (p == g_acx->rank) continue;
            if (g_acx->peer_ctrl[p]) cudaIpcCloseMemHandle(g_acx->peer_ctrl[p]);
            if (g_acx->peer_heap[p]) cudaIpcCloseMemHandle(g_acx->peer_heap[p]);
        }
        MPI_Barrier(MPI_COMM_WORLD);
        for (auto &r : g_acx->host_regs) cudaHostUnregister((void *)r.lo);
        {
            std::lock_guard<std::mutex> lk(g_host_streams_mu);
            for (cudaStream_t hs : g_host_streams) if (hs != g_acx->istream) cudaStreamDestroy(hs);
            g_host_streams.clear();
        }
        cudaStreamDestroy(g_acx->istream);
        for (auto ss : g_side) if (ss) cudaStreamDestroy(ss);
        g_side.clear();
        for (auto &kv : g_chan_streams) cudaStreamDestroy(kv.second);
        g_chan_streams.clear();
        for (OverflowStream &o : g_ovf_pool) cudaStreamDestroy(o.s);
        g_ovf_pool.clear();
        for (auto e : g_event_pool) cudaEventDestroy(e);
        g_event_pool.clear();
        cudaFreeHost(g_acx->hstat);
        cudaFreeHost(g_diag);
        cudaFree(g_trace_cursor);
        if (g_any_tbl_d) { cudaFree(g_any_tbl_d); g_any_tbl_d = nullptr; }
        if (g_anyres_d) { cudaFree(g_anyres_d); g_anyres_d = nullptr; }
        cudaFree(g_acx->ctrl);
        cudaFree(g_acx->heap);
        g_diag = g_diag_d = nullptr;
    }
    MPI_Comm_free(&g_acx->icomm);
    delete g_acx;
    g_acx = nullptr;
    return err;
}

extern "C" int MPIX_Is_device_transport(void) { return g_acx && g_acx->device ? 1 : 0; }

extern "C" int MPIX_Get_stats(MPIX_Acx_stats *s) {
    if (!g_acx || !s) return MPI_ERR_ARG;
    if (g_acx->device) { std::lock_guard<std::recursive_mutex> lk(g_acx->mu); acx_reclaim_retired(); }   // slots of retired operations / destroyed graphs count as free
    s->sends_enqueued = g_acx->st.sends; s->recvs_enqueued = g_acx->st.recvs; s->waits_enqueued = g_acx->st.waits;
    s->bytes_sent = g_acx->st.bytes_sent; s->bytes_received = g_acx->st.bytes_recv;
    s->eager_msgs = g_acx->st.eager; s->rendezvous_msgs = g_acx->st.rndv; s->bounce_copies = g_acx->st.bounce;
    s->kernels_launched = g_acx->st.kernels; s->graph_nodes_added = g_acx->st.graph_nodes;
    s->slots_in_use = g_acx->slots_in_use; s->slots_high_water = g_acx->st.slots_hw;
    s->heap_bytes_in_use = g_acx->heap_in_use; s->heap_bytes_high_water = g_acx->st.heap_hw;
    s->copy_engine_msgs = g_acx->st.ce;
    s->overflow_sends = g_acx->st.send_overflow;
    return MPI_SUCCESS;
}

extern "C" int MPIX_Trace_dump(const char *path) {
    if (!g_acx || !g_acx->device || !g_acx->trace) return MPI_ERR_OTHER;
    uint64_t cur = 0;
    cudaMemcpy(&cur, g_trace_cursor, 8, cudaMemcpyDeviceToHost);
    uint32_t n = g_acx->geom.ntrace;
    std::vector<acx_trace_rec> recs(n);
    cudaMemcpy(recs.data(), g_acx->ctrl + g_acx->geom.off_trace, sizeof(acx_trace_rec) * n, cudaMemcpyDeviceToHost);
    FILE *f = fopen(path, "w");
    if (!f) return MPI_ERR_OTHER;
    static const char *names[] = {"?", "send_trigger", "recv_trigger", "move_begin", "move_end", "wait_begin", "wait_end"};
    uint64_t first = cur > n ? cur - n : 0;
    for (uint64_t i = first; i < cur; i++) {
        const acx_trace_rec &r = recs[i & (n - 1)];
        fprintf(f, "{\"rank\": %d, \"t_ns\": %llu, \"ev\": \"%s\", \"peer\": %u, \"a\": %llu, \"b\": %llu}\n", g_acx->rank, (unsigned long long)r.t_ns,
                r.ev < 7 ? names[r.ev] : "?", r.peer, (unsigned long long)r.a, (unsigned long long)r.b);
    }
    fclose(f);
    return MPI_SUCCESS;
}

extern "C" int MPIX_Alloc_mem(size_t bytes, MPI_Info, void *baseptr) {
    if (!g_acx || !g_acx->initialized) return MPI_ERR_OTHER;
    if (!g_acx->device) {
        void *p = malloc(bytes ? bytes : 1);
        *(void **)baseptr = p;
        return p ? MPI_SUCCESS : MPI_ERR_OTHER;
    }
    int64_t off = acx_heap_alloc(bytes);
    if (off < 0) return MPI_ERR_OTHER;
    *(void **)baseptr = g_acx->heap + off;
    return MPI_SUCCESS;
}

extern "C" int MPIX_Free_mem(void *base) {
    if (!g_acx || !g_acx->initialized) return MPI_ERR_OTHER;
    if (!g_acx->device) { free(base); return MPI_SUCCESS; }
    int64_t off = acx_heap_offset_of(base);
    if (off < 0) return MPI_ERR_ARG;
    acx_heap_free(off);
    return MPI_SUCCESS;
}

extern "C" int MPIX_Peer_ptr(const void *local, int peer, void **peer_ptr) {
    if (!g_acx || !g_acx->device || !peer_ptr || peer < 0 || peer >= g_acx->size) return MPI_ERR_ARG;
    int64_t off = acx_heap_offset_of(local);
    if (off < 0) return MPI_ERR_BUFFER;
    *peer_ptr = g_acx->peer_heap[peer] + off;
    return MPI_SUCCESS;
}

extern "C" int MPIX_Heap_offset(const void *local, size_t *offset) {
    if (!g_acx || !g_acx->device || !offset) return MPI_ERR_ARG;
    int64_t off = acx_heap_offset_of(local);
    if (off < 0) return MPI_ERR_BUFFER;
    *offset = (size_t)off;
    return MPI_SUCCESS;
}
extern "C" int MPIX_Peer_ptr_at(size_t offset, int peer, void **peer_ptr) {
    if (!g_acx || !g_acx->device || !peer_ptr || peer < 0 || peer >= g_acx->size || offset >= g_acx->heap_bytes) return MPI_ERR_ARG;
    *peer_ptr = g_acx->peer_heap[peer] + offset;
    return MPI_SUCCESS;
}

// =======================================================================================
// emitter: the same op sequence goes to a stream, a capturing stream, or a fresh graph
// =======================================================================================
struct Emitter {
    int qtype = MPIX_QUEUE_CUDA_STREAM;
    cudaStream_t stream = nullptr;
    bool capturing = false;
    cudaGraph_t graph = nullptr;          // GRAPH mode: graph under construction; capture: the capture graph
    cudaGraphNode_t last = nullptr;
    bool parallel = false;                // GRAPH mode: do not chain nodes (independent waits)
    bool no_pdl = false;                  // plain launches (kernels that sit next to a copy-engine copy in their stream)

    int begin(int qt, void *queue) {
        qtype = qt;
        if (qt == MPIX_QUEUE_CUDA_STREAM) {
            stream = *(cudaStream_t *)queue;
            cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
            cudaGraph_t cg = nullptr;
            cudaError_t e = cudaStreamGetCaptureInfo_v2(stream, &st, nullptr, &cg, nullptr, nullptr);
            if (e != cudaSuccess) { ACX_ERR("cudaStreamGetCaptureInfo: %s", cudaGetErrorString(e)); return MPI_ERR_OTHER; }
            capturing = st == cudaStreamCaptureStatusActive;
            graph = capturing ? cg : nullptr;
            return 0;
        }
        if (qt == MPIX_QUEUE_CUDA_GRAPH) {
            cudaError_t e = cudaGraphCreate(&graph, 0);
            if (e != cudaSuccess) { ACX_ERR("cudaGraphCreate: %s", cudaGetErrorString(e)); return MPI_ERR_OTHER; }
            return 0;
        }
        ACX_ERR("invalid queue type %d", qt);
        return MPI_ERR_ARG;
    }
    bool in_graph() const { return qtype == MPIX_QUEUE_CUDA_GRAPH || capturing; }

    template <typename P>
    int launch(void (*kern)(P), dim3 grid, dim3 block, P &params, size_t smem = 0) {
        void *args[] = {&params};
        if (qtype == MPIX_QUEUE_CUDA_STREAM) {
            cudaError_t e;
            if (g_pdl && !capturing && !no_pdl) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                e = cudaLaunchKernelExC(&cfg, (const void *)kern, args);
            } else {
                e = cudaLaunchKernel((const void *)kern, grid, block, args, smem, stream);
            }
            if (e != cudaSuccess) { report_watchdog(); ACX_ERR("kernel launch failed: %s", cudaGetErrorString(e)); return MPI_ERR_OTHER; }
            g_acx->st.kernels++;
            if (capturing) g_acx->st.graph_nodes++;
            return 0;
        }
        cudaKernelNodeParams np = {};
        np.func = (void *)kern; np.gridDim = grid; np.blockDim = block; np.sharedMemBytes = (unsigned)smem; np.kernelParams = args; np.extra = nullptr;
        cudaGraphNode_t node;
        cudaError_t e = cudaGraphAddKernelNode(&node, graph, (last && !parallel) ? &last : nullptr, (last && !parallel) ? 1 : 0, &np);
        if (e != cudaSuccess) { ACX_ERR("cudaGraphAddKernelNode: %s", cudaGetErrorString(e)); return MPI_ERR_OTHER; }
        last = node;
        g_acx->st.graph_nodes++;
        return 0;
    }
    template <typename P, typename Q>
    int launch2(void (*kern)(P, Q), dim3 grid, dim3 block, P &p1, Q &p2) {
        void *args[] = {&p1, &p2};
        if (qtype == MPIX_QUEUE_CUDA_STREAM) {
            cudaError_t e;
            if (g_pdl && !capturing && !no_pdl) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = grid; cfg.blockDim = block; cfg.stream = stream;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                e = cudaLaunchKernelExC(&cfg, (const void *)kern, args);
            } else e = cudaLaunchKernel((const void *)kern, grid, block, args, 0, stream);
            if (e != cudaSuccess) { report_watchdog(); ACX_ERR("kernel launch failed: %s", cudaGetErrorString(e)); return MPI_ERR_OTHER; }
            g_acx->st.kernels++;
            if (capturing) g_acx->st.graph_nodes++;
            return 0;
        }
        cudaKernelNodeParams np = {};
        np.func = (void *)kern; np.gridDim = grid; np.blockDim = block; np.kernelParams = args;
        cudaGraphNode_t node;
        cudaError_t e = cudaGraphAddKernelNode(&node, graph, (last && !parallel) ? &last : nullptr, (last && !parallel) ? 1 : 0, &np);
        if (e != cudaSuccess) { ACX_ERR("cudaGraphAddKernelNode: %s", cudaGetErrorString(e)); return MPI_ERR_OTHER; }
        last = node;
        g_acx->st.graph_nodes++;
        return 0;
    }
    int launch3(void (*kern)(char *, const char *, uint64_t), dim3 grid, dim3 block, char *a, const char *b, uint64_t c) {
        void *args[] = {&a, &b, &c};
        if (qtype == MPIX_QUEUE_CUDA_STREAM) {
            cudaError_t e = cudaLaunchKernel((const void *)kern, grid, block, args, 0, stream);
            if (e != cudaSuccess) { ACX_ERR("kernel launch failed: %s", cudaGetErrorString(e)); return MPI_ERR_OTHER; }
            g_acx->st.kernels++;
            return 0;
        }
        cudaKernelNodeParams np = {};
        np.func = (void *)kern; np.gridDim = grid; np.blockDim = block; np.kernelParams = args;
        cudaGraphNode_t node;
        cudaError_t e = cudaGraphAddKernelNode(&node, graph, last ? &last : nullptr, last ? 1 : 0, &np);
        if (e != cudaSuccess) { ACX_ERR("cudaGraphAddKernelNode: %s", cudaGetErrorString(e)); return MPI_ERR_OTHER; }
        last = node;
        g_acx->st.graph_nodes++;
        return 0;
    }
};

// Resources that must live as long as a CUDA graph does (the reference attaches a
// cudaUserObject whose destructor host-spins and frees, src/sendrecv.cu:106-127,174-184;
// ours only queues the slots for reuse — destructors may not call CUDA).
struct AcxGraphRes { std::vector<int> slots; unsigned long long init_gen; };
static void CUDART_CB graph_res_destroy(void *p) {
    AcxGraphRes *g = (AcxGraphRes *)p;
    {
        std::lock_guard<std::mutex> lk(g_deferred_mu);
        for (int s : g->slots) g_deferred_slots.push_back({s, g->init_gen});
    }
    g_graph_objs_live.fetch_sub(1);
    delete g;
}
static int attach_to_graph(cudaGraph_t graph, int slot) {
    AcxGraphRes *res = new AcxGraphRes();
    res->slots.push_back(slot);
    res->init_gen = g_init_generation;
    cudaUserObject_t obj;
    cudaError_t e = cudaUserObjectCreate(&obj, res, graph_res_destroy, 1, cudaUserObjectNoDestructorSync);
    if (e != cudaSuccess) { delete res; ACX_ERR("cudaUserObjectCreate: %s", cudaGetErrorString(e)); return MPI_ERR_OTHER; }
    g_graph_objs_live.fetch_add(1);
    e = cudaGraphRetainUserObject(graph, obj, 1, cudaGraphUserObjectMove);
    if (e != cudaSuccess) { cudaUserObjectRelease(obj, 1); ACX_ERR("cudaGraphRetainUserObject: %s", cudaGetErrorString(e)); return MPI_ERR_OTHER; }
    return 0;
}

static int mover_grid(size_t bytes) {
    if (g_mover_tma) {
        size_t g = (bytes + (128u << 10) - 1) / (128u << 10);   // >= 8 chunks per CTA keeps the ring busy
        int cap = g_acx->mover_max_ctas > 0 ? g_acx->mover_max_ctas : 2 * g_acx->sm_count;
        return (int)std::max<size_t>(1, std::min<size_t>(g, cap));
    }
    return 0;
}
static int mover_grid_ldst(size_t bytes);
static int mover_grid_ldst(size_t bytes) {
    size_t g = (bytes + g_acx->mover_min_cta_bytes - 1) / g_acx->mover_min_cta_bytes;
    // Half an SM-count of 512-thread CTAs already saturates one NVLink direction (measured 647 vs 652 GB/s at 148) and
    // leaves room for a pull and a push mover to co-reside when one GPU ends up doing both.
    int cap = g_acx->mover_max_ctas > 0 ? g_acx->mover_max_ctas : std::max(1, g_acx->sm_count / 2);
    if (g < 1) g = 1;
    if (g > (size_t)cap) g = cap;
    return (int)g;
}

static int launch_mover(Emitter &em, MoverParams &mp, size_t bytes) {
#ifndef ACX_EMU   // the TMA mover is inline PTX; the software device runs the LDG/STG movers
    if (g_mover_tma) {
        static bool attr = false;
        if (!attr) {
            if (cudaFuncSetAttribute(mover_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmem) != cudaSuccess) { ACX_ERR("cannot raise dynamic smem for mover_tma"); return MPI_ERR_OTHER; }
            attr = true;
        }
        return em.launch(mover_tma, dim3(mover_grid(bytes)), dim3(kTmaThreads), mp, kTmaSmem);
    }
#endif
    mp.vec256 = g_mover_256;
    mp.block_bytes = (uint32_t)g_mover_block;
    int grid = mover_grid_ldst(bytes);
    if ((int)mp.peer == g_acx->rank && g_acx->mover_max_ctas <= 0) {
        // Self-transfer (local HBM copy, no link to share): a full-width grid sweeping interleaved 128-KiB blocks
        // measured 0.129 ms per 256 MiB vs 0.144 ms for the NVLink-sized grid (profiles/raw/self_copy_layouts.txt).
        size_t g = (bytes + g_acx->mover_min_cta_bytes - 1) / g_acx->mover_min_cta_bytes;
        grid = (int)std::max<size_t>(1, std::min<size_t>(g, (size_t)4 * g_acx->sm_count));
        if (!mp.block_bytes) mp.block_bytes = 128u << 10;
    }
    return em.launch(g_coop ? mover : g_mover_256 == 1 ? mover_v256 : g_mover_256 == 2 ? mover_cs : mover_plain, dim3(grid), dim3(kMoverThreads), mp);
}

// Whole-operation branching (ACX_PEER_BRANCH=1): every kernel of an operation with peer p — bounce-in copy, trigger,
// mover or fused rendezvous — goes to p's side stream, forked from the caller's stream at the enqueue point and
// joined at the wait. Operations with different peers then run concurrently (an all-pairs step is 2(N-1) kernels deep
// otherwise), while everything addressed to ONE peer stays on one stream, so device-assigned tickets keep the
// enqueue order per (peer, channel). Because eager sends also leave the caller's stream, their "complete when the
// trigger retires" shortcut needs the join too: the wait makes the caller's stream depend on ev_done.
struct BranchScope {
    Emitter &em;
    AcxRequest *r;
    cudaStream_t user = nullptr;
    bool on = false;
    BranchScope(Emitter &e, AcxRequest *req) : em(e), r(req) {}
    int begin() {
        if (!(g_branch == 1 || (g_branch == 2 && em.capturing)) || g_apush || em.qtype != MPIX_QUEUE_CUDA_STREAM || g_side.empty()) return 0;
        r->ev_trig = event_get(); r->ev_done = event_get();
        if (!r->ev_trig || !r->ev_done) return 0;   // out of events: stay on the caller's stream (still correct)
        cudaStream_t side = g_side[(size_t)r->peer % g_side.size()];
        ACX_CUDA(cudaEventRecord(r->ev_trig, em.stream));
        ACX_CUDA(cudaStreamWaitEvent(side, r->ev_trig, 0));
        user = em.stream;
        em.stream = side;
        on = true;
        r->branched = true;
        return 0;
    }
    int end() {
        if (!on) return 0;
        on = false;
        cudaStream_t side = em.stream;
        em.stream = user;
        ACX_CUDA(cudaEventRecord(r->ev_done, side));
        return 0;
    }
};

// Sends in flight beyond a channel's ring depth (ACX_SEND_OVERFLOW, default on).
// The trigger of a send draws a ticket and waits ON ITS STREAM until the ring slot of that ticket is free again, i.e. until the
// receiver has consumed the message `depth` tickets earlier. A program that enqueues more than `depth` sends on one channel
// before the operations the peer needs for progress (a burst of 100 MPIX_Isend_enqueue, then the receives, then one
// MPIX_Waitall_enqueue — legal MPI, and fine with the reference, whose triggers only set a flag for the proxy thread) used to
// stop its own stream there for good. Now the host counts, per channel, the stream-mode sends that have been enqueued and not
// yet waited on; a send that finds `depth` of them ahead of it — or earlier overflow sends still pending — becomes an
// OVERFLOW send: its trigger runs on a pool stream, forked from the caller's stream by an event (it still reads the buffer in
// stream order), waits for slot credit there and holds up nobody.
//  * Nothing is ever enqueued BEHIND a waiting trigger. CUDA feeds all streams of a process through a few in-order hardware
//    queues: an operation that waits for its stream predecessor holds up whatever other streams queued
//    behind it — and the predecessor here would be a trigger spinning for credit that only later work of this process (the
//    waits of its receives) can produce. A first version that chained a channel's overflow triggers on one stream and recorded
//    an event behind each deadlocked on the software device with one or two hardware queues. So every pending overflow trigger
//    has a pool stream to itself and is the last thing on it: a rendezvous message gets trigger and mover in one kernel
//    (send_rndv_fused, one or two CTAs) and there is no event behind it — it reports to the host through its slot's pinned
//    status record (trig_serial) and to its wait kernel through op->armed.
//  * Ticket order without stream order: the trigger gets the ticket the HOST counted for it (host_ticket) and takes it when the
//    channel's counter reaches that value. Channels whose tickets the host cannot count (graphs, several streams) keep the old
//    behaviour.
//  * A wait call completes its overflow sends LAST, behind the wait kernels of everything else it was given: the receives
//    completed there are what returns credit to the peer's overflow sends, and the peer is usually in the mirror image of the
//    same call.
// A steady-state exchange (send, receive, wait per step) never has more than a few un-waited sends per channel and never gets
// here: the cost on the hot path is one counter under the library mutex.
// What remains bounded: an overflow send completes only once the receiver has WAITED on the receive `depth` messages
// earlier (slot credit is returned by the receiver's wait, not by its trigger), and ACX_SEND_OVERFLOW_MAX of them can be
// pending at a time — see DESIGN.md section 8.
static bool overflow_trigger_pending(const AcxRequest *r) {   // has this overflow send's trigger drawn its ticket yet?
    return ACX_HOST_FLAG_GET(&g_acx->hstat[r->slot].trig_serial) != (int32_t)(uint32_t)r->serial;
}
static bool send_channel_backlogged_locked(int dest, int chan) {   // caller holds g_acx->mu
    AcxState::SendFlow &f = g_acx->send_flow[(size_t)dest][(size_t)chan];
    if (f.last) {
        if (overflow_trigger_pending(f.last)) return true;   // earlier overflow sends have not drawn their tickets yet (they draw in order)
        f.last = nullptr;
    }
    return f.unwaited >= g_acx->geom.depth;
}
static bool send_channel_backlogged(int dest, int chan) {   // a fused send (MPIX_Sendrecv_enqueue) must not jump the queue
    if (!g_send_overflow) return false;
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    return send_channel_backlogged_locked(dest, chan);
}
static void send_flow_forget(AcxRequest *r) {   // caller holds g_acx->mu; the request is going away
    if (!r->overflow || r->slot < 0 || !g_acx || (size_t)r->peer >= g_acx->send_flow.size()) return;
    AcxState::SendFlow &f = g_acx->send_flow[(size_t)r->peer][(size_t)r->chan];
    if (f.last == r) f.last = nullptr;
    for (OverflowStream &o : g_ovf_pool) if (o.owner == r) o.owner = nullptr;
}
static void send_flow_release(AcxRequest *r) {   // caller holds g_acx->mu; the request's wait has been enqueued (or it is being released)
    if (!r || !r->send_counted) return;
    r->send_counted = false;
    AcxState::SendFlow &f = g_acx->send_flow[(size_t)r->peer][(size_t)r->chan];
    if (f.unwaited > 0) f.unwaited--;
}
// Count a stream-mode send against its channel and decide whether it is an overflow send (needs the host's ticket).
// 0: ordinary send; 1: overflow send, *stream is the pool stream for its trigger; < 0: refused.
static int send_overflow_begin(AcxRequest *r, cudaStream_t *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    AcxState::SendFlow &f = g_acx->send_flow[(size_t)r->peer][(size_t)r->chan];
    const bool overflow = r->hticket != ~0ull && send_channel_backlogged_locked(r->peer, r->chan);
    if (!overflow) { f.unwaited++; r->send_counted = true; return 0; }
    OverflowStream *slot = nullptr;
    for (OverflowStream &o : g_ovf_pool) {
        if (o.owner && !overflow_trigger_pending(o.owner)) o.owner = nullptr;   // that trigger is past its waits: whatever follows it on the stream will run
        if (!o.owner) { slot = &o; break; }
    }
    if (!slot) {
        if ((long)g_ovf_pool.size() >= g_send_overflow_max) {
            ACX_ERR("%ld sends are waiting for ring slots on overflow streams (ACX_SEND_OVERFLOW_MAX) and %d sends to rank %d (tag %d) are not yet waited on: wait on earlier sends first", g_send_overflow_max, f.unwaited, r->peer, r->tag);
            return -1;
        }
        // grow the pool in blocks (the first overflow send of the process creates 16 streams, then the pool doubles): few calls
        // into the driver's stream allocator while triggers are already waiting on the device
        const size_t first_new = g_ovf_pool.size();
        const size_t grow = std::min<size_t>((size_t)g_send_overflow_max - first_new, std::max<size_t>(16, first_new));
        for (size_t i = 0; i < grow; i++) {
            cudaStream_t ns = nullptr;
            if (cudaStreamCreateWithFlags(&ns, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); break; }
            g_ovf_pool.push_back({ns, nullptr});
        }
        if (g_ovf_pool.size() == first_new) { ACX_ERR("cannot create an overflow stream"); return -1; }
        slot = &g_ovf_pool[first_new];
    }
    slot->owner = r;
    *stream = slot->s;
    f.unwaited++;
    r->send_counted = true;
    r->overflow = true;
    f.last = r;
    ACX_HOST_FLAG_SET(&g_acx->hstat[r->slot].trig_serial, 0);
    g_acx->st.send_overflow++;
    return 1;
}

// Unfused movers of stream-mode operations run on a per-peer side stream, forked after the trigger kernel: the
// user's stream is not held up by the transfer, and transfers to/from different peers overlap (all-pairs,
// halo-style neighbour exchanges). The wait joins the side stream (required when capturing).
static int launch_mover_maybe_side(Emitter &em, AcxRequest *r, MoverParams &mp, size_t bytes) {
    if (!g_side_on || em.qtype != MPIX_QUEUE_CUDA_STREAM || g_side.empty() || r->branched) return launch_mover(em, mp, bytes);
    r->ev_trig = event_get(); r->ev_done = event_get();
    if (!r->ev_trig || !r->ev_done) return launch_mover(em, mp, bytes);
    cudaStream_t side = g_side[(size_t)r->peer % g_side.size()];
    ACX_CUDA(cudaEventRecord(r->ev_trig, em.stream));
    ACX_CUDA(cudaStreamWaitEvent(side, r->ev_trig, 0));
    Emitter es = em;
    es.stream = side;
    int rc = launch_mover(es, mp, bytes);
    if (rc) return rc;
    ACX_CUDA(cudaEventRecord(r->ev_done, side));
    return 0;
}

// Deterministic large-message path: gate + push mover on the sender's (peer, channel) stream. FIFO order inside one
// channel is the matching order anyway, so a gate waiting for a late receive only ever delays messages that could not
// complete before it; every GPU moves exactly its own sends (no arrival-order lottery about who does the work).
static int launch_sender_push(Emitter &em, AcxRequest *r, SendParams &sp, MoverParams &mp, size_t bytes) {
    cudaStream_t cs;
    {
        std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
        auto key = std::make_pair(r->peer, r->chan);
        auto it = g_chan_streams.find(key);
        if (it == g_chan_streams.end()) {
            cudaStream_t ns = nullptr;
            ACX_CUDA(cudaStreamCreateWithFlags(&ns, cudaStreamNonBlocking));
            it = g_chan_streams.emplace(key, ns).first;
        }
        cs = it->second;
    }
    r->ev_trig = event_get(); r->ev_done = event_get();
    if (!r->ev_trig || !r->ev_done) return MPI_ERR_OTHER;
    ACX_CUDA(cudaEventRecord(r->ev_trig, em.stream));
    ACX_CUDA(cudaStreamWaitEvent(cs, r->ev_trig, 0));
    Emitter es = em;
    es.stream = cs;
    GateParams gp;
    memset(&gp, 0, sizeof gp);
    gp.op = sp.op; gp.sack = sp.sack; gp.peer_heap = sp.peer_heap; gp.xlat = sp.xlat; gp.bytes = bytes; gp.depth = sp.depth; gp.peer = sp.peer; gp.wd = sp.wd;
    int rc = es.launch(push_gate, dim3(1), dim3(32), gp);
    if (rc) return rc;
    if ((rc = launch_mover(es, mp, bytes))) return rc;
    ACX_CUDA(cudaEventRecord(r->ev_done, cs));
    return 0;
}

// Copy-engine variant of the above (see the record ring at ce_lookup): gate kernel, cudaMemcpyAsync, completion kernel on the
// sender's (peer, channel) stream, forked from the caller's stream behind the trigger. `dst` is the receive's buffer as mapped
// here, read from the receiver's enqueue-time record; the gate checks it against what the receive's trigger publishes.
static int launch_sender_ce(Emitter &em, AcxRequest *r, SendParams &sp, char *dst, uint64_t dst_off, const char *src, size_t bytes) {
    cudaStream_t cs;
    {
        std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
        auto key = std::make_pair(r->peer, r->chan);
        auto it = g_chan_streams.find(key);
        if (it == g_chan_streams.end()) {
            cudaStream_t ns = nullptr;
            ACX_CUDA(cudaStreamCreateWithFlags(&ns, cudaStreamNonBlocking));
            it = g_chan_streams.emplace(key, ns).first;
        }
        cs = it->second;
    }
    r->ev_done = event_get();
    if (!r->ev_done) return MPI_ERR_OTHER;
    const bool spingate = g_ce_spingate != 0;
    if (!spingate) {
        r->ev_trig = event_get();
        if (!r->ev_trig) return MPI_ERR_OTHER;
        ACX_CUDA(cudaEventRecord(r->ev_trig, em.stream));
        ACX_CUDA(cudaStreamWaitEvent(cs, r->ev_trig, 0));
    }
    Emitter es = em;
    es.stream = cs;
    es.no_pdl = true;   // the gate follows an event wait, the completion kernel a copy: full stream dependencies, nothing programmatic
    CeParams cp;
    memset(&cp, 0, sizeof cp);
    cp.op = sp.op; cp.ring = sp.ring; cp.sack = sp.sack; cp.ticket = r->hticket; cp.dst_off = dst_off; cp.bytes = bytes;
    cp.depth = sp.depth; cp.peer = sp.peer; cp.wd = sp.wd; cp.tr = sp.tr;
    cp.serial = spingate ? sp.serial : 0;
    // ACX_CE_MEMOP: the words on both sides of the copy are known on the host (ticket -> slot and generation), so the stream can
    // wait for the mirror and publish completion in its memOp unit, next to the copy engine's own queue, instead of in kernels.
    const uint32_t ce_e = (uint32_t)(r->hticket % (uint64_t)sp.depth);
    const uint64_t ce_g = r->hticket / (uint64_t)sp.depth;
    int rc = 0;
    if (g_ce_memop >= 2 && g_cuStreamBatchMemOp && !spingate) {
        MemOpBatch mb;
        mb.wait_geq64(&sp.sack[ce_e].r_gen, ce_g + 1);   // the drawn ticket is checked against the prediction by the wait kernel (0x5e0f)
        if ((rc = mb.flush(cs))) return rc;
    } else if ((rc = es.launch(ce_gate, dim3(1), dim3(32), cp))) return rc;
    ACX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, cs));
    if (g_ce_memop >= 1 && g_cuStreamBatchMemOp) {
        MemOpBatch mb;
        mb.write64(&sp.ring[ce_e].done, ce_g + 1);    // receiver may consume (peer-mapped word)
        mb.write64(&sp.sack[ce_e].sdone, ce_g + 1);   // sender's buffer is reusable
        if ((rc = mb.flush(cs))) return rc;
    } else if ((rc = es.launch(ce_complete, dim3(1), dim3(32), cp))) return rc;
    ACX_CUDA(cudaEventRecord(r->ev_done, cs));
    g_acx->st.ce++;
    return 0;
}

static void fault_delay() {
    if (g_fault_delay_us <= 0) return;
    static std::atomic<unsigned> seed{12345u};
    unsigned x = seed.fetch_add(0x9e3779b9u) * 2654435761u + (unsigned)g_acx->rank * 40503u;
    x ^= x >> 15;
    usleep((useconds_t)(x % (unsigned)(g_fault_delay_us + 1)));
}

// End-of-operation marker for stream-mode operations whose mover (or fused trigger+mover kernel) runs on the CALLER's
// stream: an event behind the operation's last kernel. A host MPIX_Wait completes on a different (internal) stream as
// soon as the data is in place — possibly while that kernel's tail still touches the op record (the fused kernels'
// last CTA re-arms the gate; a first-arriver's mover has not even started) — so the slot may only be released once
// this event has completed, exactly like the side branches' ev_done. Stream waits sit behind the kernel anyway.
static void mark_operation_end(Emitter &em, AcxRequest *r) {
    if (em.in_graph() || r->ev_done) return;
    r->ev_done = event_get();
    if (r->ev_done && cudaEventRecord(r->ev_done, em.stream) != cudaSuccess) { cudaGetLastError(); g_event_pool.push_back(r->ev_done); r->ev_done = nullptr; }
}

// ---------------------------------------------------------------------------------------
// Channel directory: exact (source, tag, communicator) matching.
// A message S -> R with (ctx, tag) travels on a per-pair channel (a FIFO ring, see internal.h). Round 1 hashed
// (tag + 7 ctx) mod NCHAN, so two tags that collided shared one FIFO and a receive order different from the send order
// swapped payloads (the reference hands tag and peer to MPI untouched, src/init.cpp:69-70,76-77, so any pattern matches).
// Now every (ctx, tag) a pair uses gets its OWN channel, assigned on first use by whichever side gets there first: the
// directory of the ordered pair lives in the job's shared-memory scratch (one 64-bit key per channel, claimed by
// compare-and-swap), so sender and receiver agree without exchanging a message. The last channel is the overflow FIFO
// for pairs that use more than NCHAN-1 distinct (ctx, tag) combinations: messages there carry ctx and tag in their
// header and a receive that finds another message at the head of the FIFO is a fatal, diagnosed error (raise ACX_NCHAN),
// never a silent mismatch. Without the shared scratch (an external MPI, > 64 ranks) everything hashes as before — with
// the same fatal check.
// ---------------------------------------------------------------------------------------
static std::atomic<uint64_t> *g_tagdir = nullptr;   // [sender][receiver][nchan - 1] keys, 0 = free
static inline uint64_t tagdir_key(int tag, int ctx) { return ((uint64_t)((uint32_t)ctx + 1u) << 32) | (uint32_t)tag; }
static inline std::atomic<uint64_t> *tagdir_row(int sender, int receiver) {
    return g_tagdir + ((size_t)sender * g_acx->size + receiver) * (size_t)(g_acx->geom.nchan - 1);
}
static inline int chan_hash(int tag, int ctx, int n) { return (int)(((unsigned)tag + 7u * (unsigned)ctx) % (unsigned)n); }
// Channel of messages `sender` -> `receiver` (world ranks) with (ctx, tag); *exact = false when it is a shared FIFO.
static int chan_lookup(int sender, int receiver, int tag, int ctx, bool *exact = nullptr) {
    const int n = g_acx->geom.nchan;
    if (exact) *exact = false;
    if (!g_tagdir || n < 2) return chan_hash(tag, ctx, n);
    const uint64_t key = tagdir_key(tag, ctx);
    std::atomic<uint64_t> *row = tagdir_row(sender, receiver);
    for (int i = 0; i < n - 1; i++) {
        uint64_t cur = row[i].load(std::memory_order_acquire);
        if (cur == 0 && row[i].compare_exchange_strong(cur, key, std::memory_order_acq_rel)) cur = key;
        if (cur == key) { if (exact) *exact = true; return i; }
    }
    static bool warned = false;
    if (!warned) { warned = true; ACX_LOG(0, "more than %d distinct (communicator, tag) pairs between ranks %d -> %d: the rest share the overflow channel (FIFO matching, checked); raise ACX_NCHAN", n - 1, sender, receiver); }
    return n - 1;
}
static void tagdir_init() {
    g_tagdir = nullptr;
#ifdef ACX_MINI_MPI
    long on = 1;
    acx_env_int("ACX_TAGDIR", nullptr, 1, &on);
    size_t sb = 0;
    char *scratch = (char *)ACX_Mpi_shared_scratch(&sb);
    const size_t off = (sizeof(AcxRegTable) * 64 + 4095) & ~(size_t)4095;   // behind the registration tables (always sized for 64 ranks)
    const size_t need = (size_t)g_acx->size * g_acx->size * (size_t)(g_acx->geom.nchan - 1) * sizeof(std::atomic<uint64_t>);
    if (!on || !scratch || g_acx->geom.nchan < 2 || off + need > sb) return;
    g_tagdir = (std::atomic<uint64_t> *)(scratch + off);
    // a re-initialization starts from an empty directory: every rank clears the rows it sends on; MPIX_Init ends with a
    // barrier, and nobody enqueues before that
    for (int r = 0; r < g_acx->size; r++)
        for (int i = 0; i < g_acx->geom.nchan - 1; i++) tagdir_row(g_acx->rank, r)[i].store(0, std::memory_order_release);
#endif
}

// Predict the ticket the trigger kernel of a stream-mode operation will draw (see AcxState::ChanCount). ~0 = unknown.
static uint64_t host_ticket(bool is_send, int peer, int chan, cudaStream_t stream, bool in_graph) {
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    AcxState::ChanCount &c = (is_send ? g_acx->send_count : g_acx->recv_count)[(size_t)peer][(size_t)chan];
    if (in_graph) c.dev = true;                                   // a graph's launches are invisible to the host
    else if (c.next > 0 && c.stream != stream) c.dev = true;      // two streams: the execution order of their triggers is anybody's guess
    if (c.dev) return ~0ull;
    c.stream = stream;
    return c.next++;
}
static void host_tickets_forget_recv(int source) {   // a wildcard receive claims tickets on whichever channel matches
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    for (int s = 0; s < g_acx->size; s++) if (source < 0 || s == source) for (auto &c : g_acx->recv_count[(size_t)s]) c.dev = true;
}


// ---------------------------------------------------------------------------------------
// Copy-engine rendezvous: enqueue-time records of posted receives, in the job's shared-memory scratch.
// The device protocol learns a receive's destination when its trigger kernel EXECUTES — too late for a cudaMemcpyAsync, whose
// addresses are fixed when it is enqueued. But where both hosts count tickets (host_ticket: stream-mode operations on one
// stream per channel) the k-th send of a channel matches the k-th receive, and the receiver's host knows the destination
// when it ENQUEUES the receive: it leaves {ticket, destination code, capacity} in a small ring per (sender, receiver,
// channel). A sender that finds the record of its ticket there hands the payload to the copy engines (launch_sender_ce);
// one that does not (receive not enqueued yet, device-counted channel, graph) takes the kernel path. Which path a message
// takes is the sender's call alone; the receiver's device side is the same either way (RNDV_PUSH: it publishes its mirror
// and waits for `done`).
// ---------------------------------------------------------------------------------------
struct AcxCeRec { std::atomic<uint64_t> seq; uint64_t dst_off, cap, pad; };   // seq = ticket + 1, stored last (release)
constexpr int kCeDepth = 16;
static AcxCeRec *g_ce_recs = nullptr;
static std::unique_ptr<std::atomic<int>[]> g_ce_miss;   // [dest][chan]: consecutive sends that found no record (stop waiting for one after a few)
static inline AcxCeRec *ce_row(int sender, int receiver, int chan) {
    return g_ce_recs + (((size_t)sender * g_acx->size + receiver) * (size_t)g_acx->geom.nchan + chan) * kCeDepth;
}
static void ce_init() {
    g_ce_recs = nullptr;
#ifdef ACX_MINI_MPI
    if (g_ce_min <= 0) return;
    size_t sb = 0;
    char *scratch = (char *)ACX_Mpi_shared_scratch(&sb);
    size_t off = (sizeof(AcxRegTable) * 64 + 4095) & ~(size_t)4095;
    off += (size_t)g_acx->size * g_acx->size * (size_t)std::max(0, (int)g_acx->geom.nchan - 1) * sizeof(std::atomic<uint64_t>);   // the channel directory
    off = (off + 4095) & ~(size_t)4095;
    const size_t need = (size_t)g_acx->size * g_acx->size * (size_t)g_acx->geom.nchan * kCeDepth * sizeof(AcxCeRec);
    if (!scratch || off + need > sb) return;
    g_ce_recs = (AcxCeRec *)(scratch + off);
    // a re-initialization starts from empty rings: every rank clears the rows it receives on (MPIX_Init ends with a barrier)
    for (int s = 0; s < g_acx->size; s++)
        for (int c = 0; c < (int)g_acx->geom.nchan; c++)
            for (int i = 0; i < kCeDepth; i++) ce_row(s, g_acx->rank, c)[i].seq.store(0, std::memory_order_release);
    g_ce_miss.reset(new std::atomic<int>[(size_t)g_acx->size * g_acx->geom.nchan]);
    for (size_t i = 0; i < (size_t)g_acx->size * g_acx->geom.nchan; i++) g_ce_miss[i].store(0, std::memory_order_relaxed);
#endif
}
// Receiver, at enqueue time: receive number `ticket` of (source -> me, chan) will land at `dst_off` (heap offset or registration code).
static void ce_publish(int source, int chan, uint64_t ticket, uint64_t dst_off, uint64_t cap) {
    if (!g_ce_recs || ticket == ~0ull || dst_off == ~0ull) return;
    AcxCeRec *rec = ce_row(source, g_acx->rank, chan) + ticket % kCeDepth;
    rec->seq.store(0, std::memory_order_relaxed);
    std::atomic_thread_fence(std::memory_order_release);
    rec->dst_off = dst_off; rec->cap = cap;
    rec->seq.store(ticket + 1, std::memory_order_release);
}
// Sender, at enqueue time: the address (mapped here) send number `ticket` of (me -> dest, chan) must be copied to, or null.
static char *ce_lookup(int dest, int chan, uint64_t ticket, size_t bytes, cudaStream_t stream, uint64_t *dst_off) {
    if (!g_ce_recs || ticket == ~0ull) return nullptr;
    AcxCeRec *rec = ce_row(g_acx->rank, dest, chan) + ticket % kCeDepth;
    std::atomic<int> &miss = g_ce_miss[(size_t)dest * g_acx->geom.nchan + chan];
    uint64_t off = 0, cap = 0;
    bool found = false, extended = false;
    long limit = g_ce_spin_us;
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        const uint64_t s1 = rec->seq.load(std::memory_order_acquire);
        if (s1 == ticket + 1) {
            off = rec->dst_off; cap = rec->cap;
            std::atomic_thread_fence(std::memory_order_acquire);
            if (rec->seq.load(std::memory_order_relaxed) == s1) { found = true; break; }   // not overwritten by a later receive meanwhile
            break;
        }
        if (s1 > ticket + 1) break;   // a later receive has taken the record: the one for this ticket is gone
        if (miss.load(std::memory_order_relaxed) >= 8 || g_ce_spin_us <= 0) break;   // this peer does not post its receives ahead of my sends: do not hold the caller up
        const long waited = (long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        if (waited >= limit) {
            // Hosts enqueue far ahead of their GPUs and drift apart; a host whose stream still has work queued loses nothing by
            // letting its neighbour's host catch up (which keeps the ranks' enqueue fronts together, and the copy engines in use).
            if (!extended && g_ce_spin_busy_us > limit && cudaStreamQuery(stream) == cudaErrorNotReady) { cudaGetLastError(); extended = true; limit = g_ce_spin_busy_us; continue; }
            cudaGetLastError();
            break;
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    if (!found) { if (miss.load(std::memory_order_relaxed) < 1 << 20) miss.fetch_add(1, std::memory_order_relaxed); return nullptr; }
    miss.store(0, std::memory_order_relaxed);
    if (cap < bytes) return nullptr;   // truncation: the kernel path reports it
    *dst_off = off;
    if (off & ACX_OFF_REG_BIT) {
#ifdef ACX_MINI_MPI
        const size_t idx = (size_t)dest * g_acx->geom.maxreg + ACX_OFF_REG_IDX(off);
        if (idx >= g_xlat_host.size() || !g_xlat_host[idx]) return nullptr;
        return (char *)(g_xlat_host[idx] + ACX_OFF_REG_OFF(off));
#else
        return nullptr;
#endif
    }
    return g_acx->peer_heap[dest] + off;
}

static AcxRequest *new_request(AcxReqKind kind, int peer, int tag, int ctx, size_t bytes, const void *buf) {
    AcxRequest *r = new AcxRequest();
    memset((void *)r, 0, sizeof *r);
    r->magic = ACX_REQ_MAGIC; r->kind = kind; r->peer = peer; r->comm_peer = peer; r->tag = tag; r->comm_ctx = ctx; r->bytes = bytes;
    r->user_buf = (void *)buf; r->bounce_off = -1; r->heap_off = -1; r->slot = -1; r->land_off = -1;
    r->serial = g_acx->serial++;
    r->hticket = ~0ull;
    r->mpi_req = MPI_REQUEST_NULL;
    return r;
}

// Layout of `count` items of `dt`. Contiguous types (the common case) come back with noncontig = false and are moved
// as count x size bytes; strided ones (vector / subarray / nestings, as halo faces are) are packed / unpacked by kernels.
static int describe_type(const char *fn, MPI_Datatype dt, int count, int sz, AcxLayout *out) {
    memset(out, 0, sizeof *out);
    out->n[0] = out->n[1] = out->n[2] = 1;
    out->run = sz; out->extent = sz; out->count = count; out->span = (long long)count * sz;
#ifdef ACX_MINI_MPI
    ACX_Mpi_layout L;
    if (ACX_Mpi_type_layout(dt, &L) != MPI_SUCCESS) { ACX_ERR("%s: invalid datatype", fn); return MPI_ERR_TYPE; }
    if (L.contiguous && L.disp == 0) return 0;
    long long reach = L.disp + L.run;
    for (int d = 0; d < 3; d++) {
        if (L.stride[d] < 0) { ACX_ERR("%s: datatypes with negative strides are not supported", fn); return MPI_ERR_TYPE; }
        out->n[d] = L.n[d]; out->stride[d] = L.stride[d];
        reach += (L.n[d] - 1) * L.stride[d];
    }
    if (L.disp < 0 || L.extent < 0) { ACX_ERR("%s: datatypes with negative displacement/extent are not supported", fn); return MPI_ERR_TYPE; }
    out->run = L.run; out->disp = L.disp; out->extent = L.extent;
    out->span = count > 0 ? (long long)(count - 1) * L.extent + reach : 0;
    out->noncontig = true;
    return 0;
#else
    MPI_Aint lb = 0, ext = 0;
    if (MPI_Type_get_extent(dt, &lb, &ext) != MPI_SUCCESS || lb != 0 || ext != (MPI_Aint)sz) {
        ACX_ERR("%s: non-contiguous datatypes need the bundled MPI (layout introspection); pack the buffer first", fn);
        return MPI_ERR_TYPE;
    }
    return 0;
#endif
}
static int pack_grid(size_t bytes, uint32_t vec) {
    size_t units = bytes / vec, ctas = (units + 256 * 8 - 1) / (256 * 8);   // ~8 units per thread
    return (int)std::max<size_t>(1, std::min<size_t>(ctas, (size_t)4 * g_acx->sm_count));
}

// Validates the arguments shared by the enqueue calls; *wpeer receives the peer's MPI_COMM_WORLD rank (wildcards and
// MPI_PROC_NULL pass through unchanged).
static int check_common(const char *fn, bool is_recv, int count, MPI_Datatype dt, int peer, int tag, MPI_Comm comm, MPIX_Request *request, void *queue, size_t *bytes, int *ctx, int *wpeer, AcxLayout *ty) {
    if (!g_acx || !g_acx->initialized) { fprintf(stderr, "[acx] %s: MPIX_Init has not been called\n", fn); return MPI_ERR_OTHER; }
    if (!request || !queue) { ACX_ERR("%s: null request/queue", fn); return MPI_ERR_ARG; }
    if (count < 0) { ACX_ERR("%s: negative count", fn); return MPI_ERR_COUNT; }
    int sz = 0;
    if (MPI_Type_size(dt, &sz) != MPI_SUCCESS || sz <= 0) { ACX_ERR("%s: invalid datatype", fn); return MPI_ERR_TYPE; }
    // MPI_PROC_NULL is valid everywhere. Wildcards (receive side only) are matched by the MPI library on the host
    // transport; the device transport pairs messages by (peer, channel) tickets and cannot honour them yet.
    const AcxCommMap *cm = nullptr;
    {
        int rc = acx_comm_ctx(comm, ctx, &cm);
        if (rc) return rc;
    }
    const bool wild_ok = is_recv && (!g_acx->device || g_wild);
    if (peer == MPI_ANY_SOURCE && wild_ok && g_acx->device && !cm->identity) {
        ACX_ERR("%s: MPI_ANY_SOURCE on the device transport is limited to communicators that span MPI_COMM_WORLD in order", fn);
        return MPI_ERR_RANK;
    }
    if (peer != MPI_PROC_NULL && !(wild_ok && peer == MPI_ANY_SOURCE) && (peer < 0 || peer >= cm->size)) {
        ACX_ERR("%s: peer %d out of range%s", fn, peer, peer == MPI_ANY_SOURCE ? " (MPI_ANY_SOURCE on the device transport needs ACX_WILDCARDS=1)" : "");
        return MPI_ERR_RANK;
    }
    if (tag < 0 && !(wild_ok && tag == MPI_ANY_TAG) && peer != MPI_PROC_NULL) {
        ACX_ERR("%s: tag %d invalid%s", fn, tag, tag == MPI_ANY_TAG ? " (MPI_ANY_TAG on the device transport needs ACX_WILDCARDS=1)" : "");
        return MPI_ERR_TAG;
    }
    *bytes = (size_t)count * (size_t)sz;
    *wpeer = peer >= 0 ? cm->to_world[(size_t)peer] : peer;
    return describe_type(fn, dt, count, sz, ty);
}

// A send to / receive from MPI_PROC_NULL: nothing is enqueued and no slot is taken; the request completes at the
// first wait with an empty status (source MPI_PROC_NULL, tag MPI_ANY_TAG, 0 bytes), as MPI prescribes. In
// explicit-graph mode the caller still gets a (one empty node) graph to embed.
static int make_null_op(AcxRequest *r, int qtype, void *queue, MPIX_Request *request) {
    r->null_op = true;
    r->in_graph = false;
    if (qtype == MPIX_QUEUE_CUDA_GRAPH) {
        cudaGraph_t g = nullptr;
        cudaGraphNode_t node;
        if (cudaGraphCreate(&g, 0) != cudaSuccess || cudaGraphAddEmptyNode(&node, g, nullptr, 0) != cudaSuccess) {
            ACX_ERR("cudaGraphCreate failed for a MPI_PROC_NULL operation");
            if (g) cudaGraphDestroy(g);
            r->magic = 0; delete r;
            return MPI_ERR_OTHER;
        }
        *(cudaGraph_t *)queue = g;
        r->in_graph = true;
    } else if (qtype != MPIX_QUEUE_CUDA_STREAM) {
        r->magic = 0; delete r;
        return MPI_ERR_ARG;
    } else {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(*(cudaStream_t *)queue, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive) r->in_graph = true;
        else cudaGetLastError();
    }
    *request = r;
    return MPI_SUCCESS;
}
static void complete_null_op(AcxRequest *r, MPI_Status *st) {
    if (st != MPI_STATUS_IGNORE) {
        st->MPI_SOURCE = MPI_PROC_NULL; st->MPI_TAG = MPI_ANY_TAG; st->MPI_ERROR = MPI_SUCCESS;
#ifdef ACX_MINI_MPI
        st->_bytes = 0;
#endif
    }
    r->magic = 0;
    delete r;
}

// Wildcard receives (experimental): device tables on first use, and the kernel that resolves + matches one at wait time.
static int any_tables() {
    if (g_any_tbl_d) return 0;
    CaptureRelax relax;
    if (g_acx->size > kAnyMaxRanks) { ACX_ERR("wildcard receives support up to %d ranks", kAnyMaxRanks); return MPI_ERR_OTHER; }
    AnyTable h;
    memset(&h, 0, sizeof h);
    for (int i = 0; i < g_acx->size; i++) { h.peer_ctrl[i] = g_acx->peer_ctrl[i]; h.peer_heap[i] = g_acx->peer_heap[i]; }
    ACX_CUDA(cudaMalloc((void **)&g_any_tbl_d, sizeof h));
    if (acx_h2d(g_any_tbl_d, &h, sizeof h, host_stream())) return MPI_ERR_OTHER;
    ACX_CUDA(cudaMalloc((void **)&g_anyres_d, sizeof(acx_anyres) * (size_t)g_acx->geom.nops));
    ACX_CUDA(cudaMemset(g_anyres_d, 0, sizeof(acx_anyres) * (size_t)g_acx->geom.nops));
    return 0;
}
static void fill_any(AcxRequest *r, AnyParams *ap) {
    const acx_geom &G = g_acx->geom;
    memset(ap, 0, sizeof *ap);
    ap->tbl = g_any_tbl_d; ap->ctrl = g_acx->ctrl; ap->res = g_anyres_d + r->slot;
    ap->op = (acx_op *)(g_acx->ctrl + G.off_ops) + r->slot;
    const bool rndv_capable = r->bytes > G.eager_max;
    ap->rndv_capable = rndv_capable ? 1 : 0;
    ap->dst = r->bounce_off >= 0 ? g_acx->heap + r->bounce_off : r->dev_buf;
    ap->dst_off = r->bounce_off >= 0 ? (uint64_t)r->bounce_off : (r->heap_off >= 0 ? (uint64_t)r->heap_off : ~0ull);
    ap->cap = r->bytes; ap->serial = r->serial; ap->geom = G;
    if (r->peer == MPI_ANY_SOURCE) { ap->src_lo = 0; ap->src_hi = g_acx->size - 1; } else { ap->src_lo = ap->src_hi = r->peer; }
    // channels are per (source, ctx, tag): a wildcard source or tag looks at the head of every channel and filters on the header
    if (r->tag == MPI_ANY_TAG || r->peer == MPI_ANY_SOURCE) { ap->chan_lo = 0; ap->chan_hi = G.nchan - 1; } else { ap->chan_lo = ap->chan_hi = chan_lookup(r->peer, g_acx->rank, r->tag, r->comm_ctx); }
    ap->want_ctx = (uint32_t)r->comm_ctx;
    ap->want_tag = r->tag; ap->me = (uint32_t)g_acx->rank; ap->vec256 = (uint32_t)g_mover_256;
    ap->wd = make_wd(); ap->tr = make_tr();
}
static int any_grid(const AcxRequest *r) {
    if (r->bytes <= g_acx->geom.eager_max) return 1;
    size_t g = (r->bytes + g_acx->mover_min_cta_bytes - 1) / g_acx->mover_min_cta_bytes;
    return (int)std::max<size_t>(1, std::min<size_t>(g, 32));   // all CTAs must be co-resident: they wait for CTA 0's match
}

static bool recv_channel_backlogged(int src, int chan) {   // full ring or deferred receives ahead: a fused receive must not jump the queue
    std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
    return g_acx->recv_outstanding[(size_t)src][(size_t)chan] >= g_acx->geom.depth || !g_acx->recv_deferred[(size_t)src][(size_t)chan].empty();
}
// The kernels of a posted receive: trigger (+ mover, or the single-launch rendezvous kernel). Called from MPIX_Irecv_enqueue,
// or later — for a receive that found its channel's ring full — from the wait that makes room.
struct DeferredRecv { RecvParams rp; MoverParams mp; int fgrid; bool rndv_capable; size_t bytes; };
static int launch_recv(Emitter &em, AcxRequest *r, RecvParams &rp, MoverParams &mp, int fgrid, bool rndv_capable, size_t bytes) {
    int rc;
    r->hticket = host_ticket(false, r->peer, r->chan, em.stream, r->in_graph);
    if (rndv_capable && rp.publish && !r->in_graph) ce_publish(r->peer, r->chan, r->hticket, rp.dst_off, rp.cap);   // a sender whose host sees this may use the copy engines
    if (fgrid >= g_fuse_min_ctas && fgrid <= g_fuse_max_ctas) {
        if ((rc = em.launch2(recv_rndv_fused, dim3(fgrid), dim3(kMoverThreads), rp, mp))) return rc;
    } else {
        if ((rc = em.launch(recv_trigger, dim3(1), dim3(32), rp))) return rc;
        if (rndv_capable && (rc = launch_mover_maybe_side(em, r, mp, bytes))) return rc;
    }
    if (rndv_capable) mark_operation_end(em, r);
    return 0;
}
// Launch the deferred receives of `upto`'s channel, oldest first, as far as the ring has room — at least up to and including
// `upto` (null: as many as fit). Called with the emitter of a wait that is about to complete receives of that channel.
static int launch_deferred_recvs(Emitter &em, int src, int chan, AcxRequest *upto) {
    for (;;) {
        AcxRequest *d = nullptr;
        {
            std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
            auto &q = g_acx->recv_deferred[(size_t)src][(size_t)chan];
            if (q.empty()) return 0;
            if (g_acx->recv_outstanding[(size_t)src][(size_t)chan] >= g_acx->geom.depth) {
                bool needed = false;
                if (upto) for (AcxRequest *x : q) if (x == upto) needed = true;
                if (!needed) return 0;
                return -2;   // the ring is full and `upto` is still queued: the caller completes what it has (making room) or reports the error
            }
            d = q.front();
            q.erase(q.begin());
            g_acx->recv_outstanding[(size_t)src][(size_t)chan]++;
        }
        DeferredRecv *dr = (DeferredRecv *)d->deferred;
        d->deferred = nullptr;
        d->stream = em.stream;
        const int rc = launch_recv(em, d, dr->rp, dr->mp, dr->fgrid, dr->rndv_capable, dr->bytes);
        delete dr;
        if (rc) return rc;
    }
}

// =======================================================================================
// MPIX_Isend_enqueue / MPIX_Irecv_enqueue
// =======================================================================================
// Reference: MPIX_Isend_enqueue, src/sendrecv.cu:129-228 (slot + WriteValue32(PENDING) / set<<<1,1>>>; proxy posts MPI_Isend).
extern "C" int MPIX_Isend_enqueue(const void *buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm, MPIX_Request *request,
                                  int qtype, void *queue) {
    AcxRange _r("MPIX_Isend_enqueue");
    size_t bytes; int ctx, wdest;
    AcxLayout ty;
    int rc = check_common("MPIX_Isend_enqueue", false, count, datatype, dest, tag, comm, request, queue, &bytes, &ctx, &wdest, &ty);
    if (rc) return rc;
    AcxRequest *r = new_request(ACX_REQ_SEND, wdest, tag, ctx, bytes, buf);
    r->comm_peer = dest;
    if (!g_acx->device) {
        if (qtype != MPIX_QUEUE_CUDA_STREAM) { ACX_ERR("graph queues need a GPU"); delete r; return MPI_ERR_OTHER; }
        r->host_transport = true;
        rc = MPI_Isend(buf, count, datatype, dest, tag, comm, &r->mpi_req);
        if (rc) { delete r; return rc; }
        g_acx->st.sends++; g_acx->st.bytes_sent += bytes;
        *request = r;
        return MPI_SUCCESS;
    }
    if (dest == MPI_PROC_NULL) return make_null_op(r, qtype, queue, request);
    fault_delay();
    dest = wdest;   // from here on: the world rank (windows, rings and side streams are indexed by it)
    Emitter em;
    BranchScope branch(em, r);
    if ((rc = em.begin(qtype, queue))) { delete r; return rc; }
    r->in_graph = em.in_graph();
    r->stream = qtype == MPIX_QUEUE_CUDA_STREAM ? em.stream : nullptr;
    r->chan = chan_lookup(g_acx->rank, dest, tag, ctx);
    if ((r->slot = acx_slot_alloc(r, r->in_graph)) < 0) { delete r; return MPI_ERR_OTHER; }
    if ((rc = acx_device_alias(buf, ty.noncontig ? (size_t)ty.span : bytes, &r->dev_buf))) goto fail;
    if ((rc = branch.begin())) goto fail;
    {
        const acx_geom &G = g_acx->geom;
        const bool eager = bytes <= G.eager_max;
        uint64_t src_off = ~0ull;
        const char *src_local = r->dev_buf;
        if (ty.noncontig && bytes > 0) {
            // strided layout: gather it into a heap staging buffer, which then travels like any contiguous heap buffer
            r->typed = true; r->layout = ty;
            if ((r->bounce_off = acx_heap_alloc(bytes)) < 0) { rc = MPI_ERR_OTHER; goto fail; }
            PackParams pp;
            memset(&pp, 0, sizeof pp);
            pp.packed = g_acx->heap + r->bounce_off; pp.user = r->dev_buf; pp.L = acx_device_layout(ty, r->dev_buf, pp.packed);
            if ((rc = em.launch(pack_kernel, dim3(pack_grid(bytes, pp.L.vec)), dim3(256), pp))) goto fail;
            src_off = (uint64_t)r->bounce_off;
            src_local = pp.packed;
            r->heap_off = -1;
        } else if (!eager) {
            reg_progress();
            r->heap_off = acx_heap_offset_of(buf);
            uint64_t enc = ~0ull;
            if (r->heap_off >= 0) src_off = (uint64_t)r->heap_off;
            else if ((enc = reg_encode(buf, bytes, dest)) != ~0ull) src_off = enc;   // registered allocation: zero-copy
            else {
                // private source: bounce into the symmetric heap first (local HBM copy)
                if ((r->bounce_off = acx_heap_alloc(bytes)) < 0) { rc = MPI_ERR_OTHER; goto fail; }
                src_off = (uint64_t)r->bounce_off;
                src_local = g_acx->heap + r->bounce_off;
                if ((rc = em.launch3(copy_kernel, dim3(mover_grid_ldst(bytes)), dim3(kMoverThreads), g_acx->heap + r->bounce_off, r->dev_buf, (uint64_t)bytes))) goto fail;
                g_acx->st.bounce++;
            }
        }
        const uint64_t ring_idx = acx_entry_index(G, g_acx->rank, r->chan, 0);  // at the receiver: [src = me][chan]
        const uint64_t sack_idx = acx_entry_index(G, dest, r->chan, 0);       // at me: [dst][chan]
        SendParams sp;
        memset(&sp, 0, sizeof sp);
        sp.ring = acx_entries_of(g_acx->peer_ctrl[dest]) + ring_idx;
        sp.eager = g_acx->peer_ctrl[dest] + G.off_eager + ring_idx * G.eager_max;
        sp.sack = acx_sacks_of(g_acx->ctrl) + sack_idx;
        sp.seq = (uint64_t *)(g_acx->ctrl + G.off_send_seq) + ((uint64_t)dest * G.nchan + r->chan);
        sp.op = (acx_op *)(g_acx->ctrl + G.off_ops) + r->slot;
        sp.hstat = (eager && !r->in_graph) ? g_acx->hstat_d + r->slot : nullptr;
        sp.src = src_local;   // eager payload source: the user's buffer, or the packed image of a strided one
        sp.peer_heap = g_acx->peer_heap[dest];
        sp.xlat = (const uint64_t *)(g_acx->ctrl + G.off_xlat) + (uint64_t)dest * G.maxreg;
        sp.src_off = src_off; sp.bytes = bytes; sp.serial = r->serial;
        sp.tag = (uint32_t)tag; sp.ctx = (uint32_t)ctx; sp.depth = G.depth; sp.eager_max = G.eager_max; sp.peer = dest;
        sp.wd = make_wd(); sp.tr = make_tr();
        const bool dropped = g_fault_drop_send > 0 && ++g_send_counter == g_fault_drop_send;
        if (dropped) ACX_ERR("FAULT INJECTION: dropping send #%ld (tag %d -> rank %d)", g_fault_drop_send, tag, dest);
        MoverParams mp;
        memset(&mp, 0, sizeof mp);
        mp.op = sp.op; mp.local = (char *)src_local; mp.ring = sp.ring; mp.sack = sp.sack; mp.depth = G.depth; mp.push = 1; mp.peer = dest; mp.tr = sp.tr;
        mp.vec256 = g_mover_256; mp.block_bytes = (uint32_t)g_mover_block;
        r->src_local = (char *)src_local;
        if (!dropped) r->hticket = host_ticket(true, dest, r->chan, r->stream, r->in_graph);
        // More un-waited sends on this channel than its ring holds: the trigger goes to a pool stream (see send_overflow_begin)
        cudaStream_t ovf_stream = nullptr;
        if (g_send_overflow && !dropped && !branch.on && qtype == MPIX_QUEUE_CUDA_STREAM && !em.capturing && !g_apush) {
            const int ov = send_overflow_begin(r, &ovf_stream);
            if (ov < 0) { rc = MPI_ERR_OTHER; goto fail; }
        }
        const int fgrid = (!eager && g_fuse_rndv && !g_mover_tma && !r->overflow) ? mover_grid_ldst(bytes) : 0;
        const bool fused = fgrid >= g_fuse_min_ctas && fgrid <= g_fuse_max_ctas;   // measured: pays off from ~512 KiB, costs ~1.5 us below
        const bool apush = !eager && !fused && g_apush && g_side_on && qtype == MPIX_QUEUE_CUDA_STREAM;
        sp.apush = apush ? 1 : 0;
        // Cooperative moving is a property of the MESSAGE, chosen here from its size and announced in the header
        r->coop = g_coop && !eager && !fused && !apush && !g_mover_tma && !r->overflow && (long)bytes >= g_coop_min;
        sp.coop = r->coop ? 1 : 0;   // travels in the message header: the receiver follows the sender's choice
        // Large messages: the sender moves them — the mover behind this trigger if the receive is posted by then, the helper
        // CTAs of this request's wait kernel otherwise. The receiver never pulls (except as a late fallback), so every GPU
        // copies exactly its own sends whatever order ranks arrive in (a ring step stays one NVLink push per GPU).
        // Copy-engine rendezvous: if the receiver's host has already enqueued the matching receive (its record is in shared
        // memory), the destination is known now and the copy engines move the payload; decided per message, by the sender.
        char *ce_dst = nullptr;
        uint64_t ce_off = 0;
        if (g_ce_min > 0 && !eager && !fused && !apush && !r->coop && !dropped && !r->overflow && (long)bytes >= g_ce_min && g_side_on && qtype == MPIX_QUEUE_CUDA_STREAM && !em.capturing && !r->branched)
            ce_dst = ce_lookup(dest, r->chan, r->hticket, bytes, em.stream, &ce_off);
        if (ce_dst) sp.apush = 1;   // announced as RNDV_PUSH: the receiver leaves the message alone
        r->late = g_late_min > 0 && !eager && !fused && !apush && !ce_dst && !r->coop && !r->overflow && (long)bytes >= g_late_min;
        sp.late = r->late ? 1 : 0;
        if (r->overflow) {
            // the trigger alone, on a pool stream, behind an event that marks this point of the caller's stream (what the caller's
            // stream has done to the buffer so far — and the pack / bounce copy above — comes first); nothing follows it there
            cudaEvent_t fork = event_get();
            if (!fork) { rc = MPI_ERR_OTHER; goto fail; }
            cudaError_t fe = cudaEventRecord(fork, em.stream);
            if (fe == cudaSuccess) fe = cudaStreamWaitEvent(ovf_stream, fork, 0);
            { std::lock_guard<std::recursive_mutex> lk(g_acx->mu); g_event_pool.push_back(fork); }   // the wait has captured the record: the event object may be reused
            if (fe != cudaSuccess) { ACX_ERR("cannot fork the overflow stream: %s", cudaGetErrorString(fe)); rc = MPI_ERR_OTHER; goto fail; }
            sp.hticket1 = r->hticket + 1;
            sp.hdone = &(g_acx->hstat_d + r->slot)->trig_serial; sp.hserial = (uint32_t)r->serial;
            sp.hstat = nullptr;      // its wait kernel retires the slot (an eager trigger that did would hand the op record on while that kernel is still to run)
            Emitter eo = em;
            eo.stream = ovf_stream;
            eo.no_pdl = true;        // behind an event wait: a full dependency, nothing programmatic
            // rendezvous: trigger and mover in ONE kernel (nothing may follow a waiting trigger on its stream), with a grid of at
            // most two CTAs — up to ACX_SEND_OVERFLOW_MAX of these sit resident while they wait for credit
            if (eager) rc = eo.launch(send_trigger, dim3(1), dim3(bytes > 512 ? kTriggerThreads : 32), sp);
            else rc = eo.launch2(send_rndv_fused, dim3(bytes > (1u << 20) ? 2 : 1), dim3(kMoverThreads), sp, mp);
            if (rc) goto fail;
            if (eager) g_acx->st.eager++; else g_acx->st.rndv++;
        } else if (!dropped && fused) {
            if ((rc = em.launch2(send_rndv_fused, dim3(fgrid), dim3(kMoverThreads), sp, mp))) goto fail;
        } else if (!dropped && (rc = em.launch(send_trigger, dim3(1), dim3(eager && bytes > 512 ? kTriggerThreads : 32), sp))) goto fail;
        if (r->overflow) {
            // complete as launched: no mover, no copy engine, no end-of-operation event
        } else if (!eager) {
            if (ce_dst) { if ((rc = launch_sender_ce(em, r, sp, ce_dst, ce_off, src_local, bytes))) goto fail; }
            else if (apush) { if (!dropped && (rc = launch_sender_push(em, r, sp, mp, bytes))) goto fail; }
            else if (!fused && (rc = launch_mover_maybe_side(em, r, mp, bytes))) goto fail;
            mark_operation_end(em, r);   // no-op when a side branch already carries the event
            g_acx->st.rndv++;
        } else {
            g_acx->st.eager++;
        }
    }
    if ((rc = branch.end())) goto fail;
    if (em.capturing && (rc = attach_to_graph(em.graph, r->slot))) goto fail;
    if (qtype == MPIX_QUEUE_CUDA_GRAPH) {
        if ((rc = attach_to_graph(em.graph, r->slot))) goto fail;
        *(cudaGraph_t *)queue = em.graph;
    }
    g_acx->st.sends++; g_acx->st.bytes_sent += bytes;
    *request = r;
    return MPI_SUCCESS;
fail:
    if (qtype == MPIX_QUEUE_CUDA_GRAPH && em.graph) cudaGraphDestroy(em.graph);
    { std::lock_guard<std::recursive_mutex> lk(g_acx->mu); acx_slot_free(r->slot); }
    return rc;
}

// Reference: MPIX_Irecv_enqueue, src/sendrecv.cu:231-327.
extern "C" int MPIX_Irecv_enqueue(void *buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPIX_Request *request, int qtype,
                                  void *queue) {
    AcxRange _r("MPIX_Irecv_enqueue");
    size_t bytes; int ctx, wsource;
    bool counted = false;   // this receive has been added to recv_outstanding (undone on the error exits)
    bool defer = false;     // the channel's ring is full: keep the receive, launch its trigger from a later wait
    AcxLayout ty;
    int rc = check_common("MPIX_Irecv_enqueue", true, count, datatype, source, tag, comm, request, queue, &bytes, &ctx, &wsource, &ty);
    if (rc) return rc;
    AcxRequest *r = new_request(ACX_REQ_RECV, wsource, tag, ctx, bytes, buf);
    r->comm_peer = source;
    if (!g_acx->device) {
        if (qtype != MPIX_QUEUE_CUDA_STREAM) { ACX_ERR("graph queues need a GPU"); delete r; return MPI_ERR_OTHER; }
        r->host_transport = true;
        rc = MPI_Irecv(buf, count, datatype, source, tag, comm, &r->mpi_req);
        if (rc) { delete r; return rc; }
        g_acx->st.recvs++; g_acx->st.bytes_recv += bytes;
        *request = r;
        return MPI_SUCCESS;
    }
    if (source == MPI_PROC_NULL) return make_null_op(r, qtype, queue, request);
    fault_delay();
    source = wsource;   // world rank from here on
    Emitter em;
    BranchScope branch(em, r);
    if ((rc = em.begin(qtype, queue))) { delete r; return rc; }
    r->in_graph = em.in_graph();
    r->stream = qtype == MPIX_QUEUE_CUDA_STREAM ? em.stream : nullptr;
    if (source == MPI_ANY_SOURCE || tag == MPI_ANY_TAG) {
        // Wildcard: nothing is launched now; the wait resolves, matches and (for a rendezvous message) pulls.
        if (ty.noncontig) { ACX_ERR("wildcard receives with non-contiguous datatypes are not supported on the device transport"); if (qtype == MPIX_QUEUE_CUDA_GRAPH) cudaGraphDestroy(em.graph); delete r; return MPI_ERR_TYPE; }
        r->wild = true; r->chan = -1;
        host_tickets_forget_recv(source == MPI_ANY_SOURCE ? -1 : wsource);   // it will claim a ticket on whichever channel matches: the host can no longer predict those
        if ((rc = any_tables())) { if (qtype == MPIX_QUEUE_CUDA_GRAPH) cudaGraphDestroy(em.graph); delete r; return rc; }
        if ((r->slot = acx_slot_alloc(r, r->in_graph)) < 0) { if (qtype == MPIX_QUEUE_CUDA_GRAPH) cudaGraphDestroy(em.graph); delete r; return MPI_ERR_OTHER; }
        if ((rc = acx_device_alias(buf, bytes, &r->dev_buf))) goto fail;
        r->heap_off = acx_heap_offset_of(buf);
        if (bytes > g_acx->geom.eager_max && r->heap_off < 0) {   // a sender may push: the landing zone must be in the heap
            if ((r->bounce_off = acx_heap_alloc(bytes)) < 0) { rc = MPI_ERR_OTHER; goto fail; }
            g_acx->st.bounce++;
        }
        if (qtype == MPIX_QUEUE_CUDA_GRAPH) { cudaGraphNode_t node; cudaGraphAddEmptyNode(&node, em.graph, nullptr, 0); }
        goto done;
    }
    r->chan = chan_lookup(wsource, g_acx->rank, tag, ctx);
    if (!r->in_graph) {
        // A channel's ring holds ACX_DEPTH receives between trigger and wait. The reference allows MPIACX_NFLAGS operations in
        // flight (include/mpi-acx-internal.h:141, src/triggered.cpp:35-62); here a receive beyond the ring's depth — or behind an
        // earlier deferred one: tickets are drawn in posting order — keeps its slot and buffers but its trigger is launched
        // later, from the wait that makes room (launch_deferred_recvs). Round 1 rejected it.
        std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
        defer = g_acx->recv_outstanding[source][r->chan] >= g_acx->geom.depth || !g_acx->recv_deferred[source][r->chan].empty();
        if (!defer) { g_acx->recv_outstanding[source][r->chan]++; counted = true; }
    }
    if ((r->slot = acx_slot_alloc(r, r->in_graph)) < 0) {
        if (counted) { std::lock_guard<std::recursive_mutex> lk(g_acx->mu); g_acx->recv_outstanding[source][r->chan]--; }
        delete r;
        return MPI_ERR_OTHER;
    }
    if ((rc = acx_device_alias(buf, ty.noncontig ? (size_t)ty.span : bytes, &r->dev_buf))) goto fail;
    if ((rc = branch.begin())) goto fail;
    {
        const acx_geom &G = g_acx->geom;
        const bool rndv_capable = bytes > G.eager_max;
        uint64_t dst_off = ~0ull;
        char *dst_local = r->dev_buf;
        r->heap_off = acx_heap_offset_of(buf);
        uint64_t enc = ~0ull;
        if (rndv_capable) reg_progress();
        if (ty.noncontig && bytes > 0) {
            // strided destination: the packed image lands in a heap staging buffer; unpack_kernel scatters it after the wait
            r->typed = true; r->layout = ty; r->heap_off = -1;
            if ((r->bounce_off = acx_heap_alloc(bytes)) < 0) { rc = MPI_ERR_OTHER; goto fail; }
            dst_off = (uint64_t)r->bounce_off;
            dst_local = g_acx->heap + r->bounce_off;
            g_acx->st.bounce++;
        } else if (r->heap_off >= 0) dst_off = (uint64_t)r->heap_off;
        else if (rndv_capable && (enc = reg_encode(buf, bytes, source)) != ~0ull) dst_off = enc;   // registered: zero-copy
        else if (rndv_capable) {
            if ((r->bounce_off = acx_heap_alloc(bytes)) < 0) { rc = MPI_ERR_OTHER; goto fail; }
            dst_off = (uint64_t)r->bounce_off;
            dst_local = g_acx->heap + r->bounce_off;
            g_acx->st.bounce++;
        }
        const uint64_t ring_idx = acx_entry_index(G, source, r->chan, 0);        // at me: [src][chan]
        const uint64_t sack_idx = acx_entry_index(G, g_acx->rank, r->chan, 0);     // at the sender: [dst = me][chan]
        RecvParams rp;
        memset(&rp, 0, sizeof rp);
        rp.ring = acx_entries_of(g_acx->ctrl) + ring_idx;
        rp.seq = (uint64_t *)(g_acx->ctrl + G.off_recv_seq) + ((uint64_t)source * G.nchan + r->chan);
        rp.op = (acx_op *)(g_acx->ctrl + G.off_ops) + r->slot;
        rp.sack = acx_sacks_of(g_acx->peer_ctrl[source]) + sack_idx;
        rp.rndv_capable = rndv_capable ? 1 : 0;
        rp.publish = (rndv_capable || g_apush) ? 1 : 0;   // sender-push gates key on the mirror's generation: every receive must publish then
        rp.peer_heap = g_acx->peer_heap[source];
        rp.xlat = (const uint64_t *)(g_acx->ctrl + G.off_xlat) + (uint64_t)source * G.maxreg;
        rp.dst_off = dst_off; rp.cap = bytes; rp.serial = r->serial; rp.depth = G.depth; rp.peer = source;
        rp.wd = make_wd(); rp.tr = make_tr();
        MoverParams mp;
        memset(&mp, 0, sizeof mp);
        mp.op = rp.op; mp.local = dst_local; mp.ring = rp.ring; mp.sack = acx_sacks_of(g_acx->peer_ctrl[source]) + sack_idx;
        mp.depth = G.depth; mp.push = 0; mp.peer = source; mp.tr = rp.tr; mp.vec256 = g_mover_256; mp.block_bytes = (uint32_t)g_mover_block;
        const int fgrid = (rndv_capable && g_fuse_rndv && !g_mover_tma) ? mover_grid_ldst(bytes) : 0;
        r->coop = g_coop && rndv_capable && !g_apush && !g_mover_tma && (long)bytes >= g_coop_min;   // "may be": the wait's gate + agent kernels help only if the message says RNDV_COOP
        if (defer) {
            DeferredRecv *d = new DeferredRecv{rp, mp, fgrid, rndv_capable, bytes};
            r->deferred = d;
            std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
            g_acx->recv_deferred[source][r->chan].push_back(r);
        } else if ((rc = launch_recv(em, r, rp, mp, fgrid, rndv_capable, bytes))) goto fail;
    }
    if ((rc = branch.end())) goto fail;
done:
    if (em.capturing && (rc = attach_to_graph(em.graph, r->slot))) goto fail;
    if (qtype == MPIX_QUEUE_CUDA_GRAPH) {
        if ((rc = attach_to_graph(em.graph, r->slot))) goto fail;
        *(cudaGraph_t *)queue = em.graph;
    }
    g_acx->st.recvs++; g_acx->st.bytes_recv += bytes;
    *request = r;
    return MPI_SUCCESS;
fail:
    if (qtype == MPIX_QUEUE_CUDA_GRAPH && em.graph) cudaGraphDestroy(em.graph);
    {
        std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
        if (counted) g_acx->recv_outstanding[r->peer][r->chan]--;   // the rejected receive does not occupy a ring slot
        acx_slot_free(r->slot);
    }
    return rc;
}

// =======================================================================================
// waits
// =======================================================================================
static AcxRequest *as_request(MPIX_Request h) {
    AcxRequest *r = (AcxRequest *)h;
    if (!r || r->magic != ACX_REQ_MAGIC) return nullptr;
    return r;
}

static int fill_wait(AcxRequest *r, MPI_Status *status, bool retire, bool host_wait, WaitOne *w) {
    const acx_geom &G = g_acx->geom;
    memset(w, 0, sizeof *w);
    w->op = (acx_op *)(g_acx->ctrl + G.off_ops) + r->slot;
    w->hstat = g_acx->hstat_d + r->slot;
    w->depth = G.depth; w->eager_max = G.eager_max;
    w->serial = (host_wait || r->overflow) ? r->serial : 0;   // an overflow send's trigger runs on a pool stream: no stream order between it and this wait
    w->trigger_only = (r->overflow && r->kind == ACX_REQ_SEND && r->bytes <= g_acx->geom.eager_max) ? 1 : 0;
    w->retire = retire ? 1 : 0;
    w->late_tag = (uint32_t)r->serial;
    w->hticket = r->hticket;
    w->src = r->comm_peer; w->tag = r->tag; w->ctx = (uint32_t)r->comm_ctx;   // status reports the peer in the caller's communicator
    w->off_source = w->off_tag = w->off_error = w->off_bytes = 0xffff;
    if (r->kind == ACX_REQ_SEND) {
        w->is_recv = 0;
        w->sack = acx_sacks_of(g_acx->ctrl) + acx_entry_index(G, r->peer, r->chan, 0);
        if (r->late) {   // may have to push from the wait: see late_helper
            w->late = 1;
            w->ring = acx_entries_of(g_acx->peer_ctrl[r->peer]) + acx_entry_index(G, g_acx->rank, r->chan, 0);   // the slot at the receiver (claim, done)
            w->late_local = r->src_local;
            w->peer_heap = g_acx->peer_heap[r->peer];
            w->xlat = (const uint64_t *)(g_acx->ctrl + G.off_xlat) + (uint64_t)r->peer * G.maxreg;
            w->late_bytes = r->bytes;
        }
    } else {
        w->is_recv = 1;
        if (r->wild) {
            w->any = g_anyres_d + r->slot;   // ring / sack / eager / source are resolved on the device
        } else {
            const uint64_t ring_idx = acx_entry_index(G, r->peer, r->chan, 0);
            w->ring = acx_entries_of(g_acx->ctrl) + ring_idx;
            w->eager = g_acx->ctrl + G.off_eager + ring_idx * G.eager_max;
            w->sack = acx_sacks_of(g_acx->peer_ctrl[r->peer]) + acx_entry_index(G, g_acx->rank, r->chan, 0);
        }
        w->dst = r->dev_buf;
        w->bounce = r->bounce_off >= 0 ? g_acx->heap + r->bounce_off : nullptr;
        w->coop = g_coop ? 1 : 0;   // the receiver's wait kernel owns the slot's chunk cursor reset
        if (r->typed) { w->dst = w->bounce; w->bounce = nullptr; }   // eager payloads are copied into the staging buffer too; unpack_kernel does the rest
        w->cap = r->bytes;
        if (g_late_min > 0 && !r->wild && (long)r->bytes >= g_late_min) {   // may receive an RNDV_LATE message: fallback pull from the wait
            w->late = 1;
            w->late_local = r->bounce_off >= 0 ? g_acx->heap + r->bounce_off : r->dev_buf;
            w->peer_heap = g_acx->peer_heap[r->peer];
            w->xlat = (const uint64_t *)(g_acx->ctrl + G.off_xlat) + (uint64_t)r->peer * G.maxreg;
        }
        if (status != MPI_STATUS_IGNORE && !host_wait) {
            char *ds = nullptr;
            int rc = acx_device_alias(status, sizeof(MPI_Status), &ds);
            if (rc) return rc;
            w->ustatus = ds;
            w->off_source = (uint16_t)offsetof(MPI_Status, MPI_SOURCE);
            w->off_tag = (uint16_t)offsetof(MPI_Status, MPI_TAG);
            w->off_error = (uint16_t)offsetof(MPI_Status, MPI_ERROR);
#ifdef ACX_MINI_MPI
            w->off_bytes = (uint16_t)offsetof(MPI_Status, _bytes);
#endif
        }
    }
    return 0;
}

static void host_status_send(MPI_Status *st, AcxRequest *r) {
    if (st == MPI_STATUS_IGNORE) return;
    st->MPI_SOURCE = r->comm_peer; st->MPI_TAG = r->tag; st->MPI_ERROR = MPI_SUCCESS;
#ifdef ACX_MINI_MPI
    st->_bytes = 0;
#endif
}

static void fill_unpack(AcxRequest *r, bool retire, UnpackParams *up) {
    memset(up, 0, sizeof *up);
    up->op = (acx_op *)(g_acx->ctrl + g_acx->geom.off_ops) + r->slot;
    up->user = r->dev_buf; up->packed = g_acx->heap + r->bounce_off; up->hstat = g_acx->hstat_d + r->slot;
    up->L = acx_device_layout(r->layout, r->dev_buf, up->packed);
    up->retire = retire ? 1 : 0;
}

// Grid of a wait kernel: one waiter CTA per request plus the helper CTAs of its RNDV_LATE requests (a late send's helpers
// push the message if nobody has moved it yet, a receive's — fewer: it is a fallback — pull it). Sends share one CTA per SM.
static int plan_wait_grid(WaitBatch &wb, int n) {
    wb.nreq = (uint32_t)n;
    wb.fallback_ns = (uint64_t)g_late_fallback_us * 1000ull;
    wb.early_release_max = (uint32_t)g_wait_early_helpers;
    int nsend = 0;
    for (int j = 0; j < n; j++) if (wb.w[j].late && !wb.w[j].is_recv) nsend++;
    const size_t unit = 128u << 10;
    int next = 0;
    for (int j = 0; j < n; j++) {
        WaitOne &w = wb.w[j];
        if (!w.late) continue;
        const size_t bytes = w.is_recv ? (size_t)w.cap : (size_t)w.late_bytes;
        size_t want = std::max<size_t>(1, (bytes + unit - 1) / unit);
        want = std::min<size_t>(want, w.is_recv ? 16 : (size_t)std::max(1, (int)(g_late_helpers > 0 ? g_late_helpers : g_acx->sm_count) / std::max(1, nsend)));   // 148 x 256 threads push as fast as the 74 x 512 mover (tools/bwprobe); a receive's helpers are a fallback
        w.help_lo = (uint16_t)next; w.help_n = (uint16_t)want;
        next += (int)want;
    }
    return n + next;
}

static int wait_enqueue_impl(int count, MPIX_Request *reqs, MPI_Status *statuses, int qtype, void *queue) {
    AcxRange _r("MPIX_Wait(all)_enqueue");
    if (!g_acx || !g_acx->initialized) return MPI_ERR_OTHER;
    if (count < 0 || (count > 0 && !reqs) || !queue) return MPI_ERR_ARG;
    std::vector<AcxRequest *> rs;
    for (int i = 0; i < count; i++) {
        if (reqs[i] == MPIX_REQUEST_NULL) continue;
        AcxRequest *r = as_request(reqs[i]);
        if (!r) { ACX_ERR("invalid request handle at index %d", i); return MPI_ERR_REQUEST; }
        if (r->kind == ACX_REQ_PSEND || r->kind == ACX_REQ_PRECV) {
            ACX_ERR("partitioned requests cannot be completed with MPIX_Wait(all)_enqueue; use MPIX_Wait / MPIX_Pwait_enqueue");
            return MPI_ERR_REQUEST;
        }
        if (r->null_op) {   // MPI_PROC_NULL: complete on the spot, nothing to enqueue
            complete_null_op(r, statuses == MPI_STATUSES_IGNORE ? MPI_STATUS_IGNORE : &statuses[i]);
            reqs[i] = MPIX_REQUEST_NULL;
            continue;
        }
        rs.push_back(r);
    }
    if (!g_acx->device) {
        // host transport: "the stream" is the calling thread; complete in order.
        if (qtype != MPIX_QUEUE_CUDA_STREAM) return MPI_ERR_OTHER;
        for (int i = 0; i < count; i++) {
            AcxRequest *r = as_request(reqs[i]);
            if (!r) continue;
            MPI_Status *st = statuses == MPI_STATUSES_IGNORE ? MPI_STATUS_IGNORE : &statuses[i];
            int rc = MPI_Wait(&r->mpi_req, st);
            if (rc) return rc;
            r->magic = 0; delete r;
            reqs[i] = MPIX_REQUEST_NULL;
            g_acx->st.waits++;
        }
        return MPI_SUCCESS;
    }
    Emitter em;
    int rc = em.begin(qtype, queue);
    if (rc) return rc;
    em.parallel = true;  // explicit graphs: independent wait nodes (reference: src/sendrecv.cu:544-566)
    for (AcxRequest *r : rs) {
        if (r->in_graph != em.in_graph()) {
            ACX_ERR("%s", r->in_graph ? "in-graph operations must be waited on in a graph" : "on-stream operations cannot be waited on from a graph");
            if (qtype == MPIX_QUEUE_CUDA_GRAPH) cudaGraphDestroy(em.graph);
            return MPI_ERR_REQUEST;
        }
    }
    const bool retire = !em.in_graph();
    // Eager sends complete when their trigger kernel retires: stream/graph order already gives the
    // guarantee a wait would, so no node is emitted for them (the trigger marks the slot retired).
    struct Item { AcxRequest *r; MPI_Status *st; };
    std::vector<Item> need;
    {
        int idx = 0;
        for (AcxRequest *r : rs) {
            while (reqs[idx] == MPIX_REQUEST_NULL) idx++;
            MPI_Status *st = statuses == MPI_STATUSES_IGNORE ? MPI_STATUS_IGNORE : &statuses[idx];
            idx++;
            if (r->kind == ACX_REQ_SEND && st != MPI_STATUS_IGNORE && !em.in_graph()) host_status_send(st, r);
            if (!em.in_graph() && qtype == MPIX_QUEUE_CUDA_STREAM && r->stream != em.stream) {
                // Waited on another stream than it was enqueued on (the reference allows it: its flags are global). Everything
                // below assumes stream order between the operation's kernels and its wait (the wait kernel reads the ticket the
                // trigger took; the wait of an eager send is elided altogether): create that order with an event.
                cudaEvent_t xe = event_get();
                cudaError_t xerr = xe ? cudaEventRecord(xe, r->stream) : cudaErrorMemoryAllocation;
                if (xerr == cudaSuccess) xerr = cudaStreamWaitEvent(em.stream, xe, 0);
                if (xe) { std::lock_guard<std::recursive_mutex> lk(g_acx->mu); g_event_pool.push_back(xe); }   // the wait has captured the record: the event object may be reused
                if (xerr != cudaSuccess) { ACX_ERR("cannot order the wait behind the stream the operation was enqueued on: %s", cudaGetErrorString(xerr)); return MPI_ERR_OTHER; }
            }
            if (r->ev_done && (em.capturing || r->branched)) {   // join the forked branch: required when capturing, and for branched operations (eager sends rely on it)
                cudaError_t je = cudaStreamWaitEvent(em.stream, r->ev_done, 0);
                if (je != cudaSuccess) { ACX_ERR("cudaStreamWaitEvent(join): %s", cudaGetErrorString(je)); return MPI_ERR_OTHER; }
            }
            if (r->kind == ACX_REQ_SEND && r->bytes <= g_acx->geom.eager_max && !r->overflow) continue;   // (an overflow send's trigger is on a pool stream: its wait is a kernel)
            need.push_back({r, st});
        }
    }
    // Wildcard receives: resolve + match (+ pull) kernel, then a single-entry wait that reads the resolution. They go FIRST:
    // a wildcard receive is only posted by its resolver, and a cooperative gate of one of this rank's sends (below) spins until
    // the peer's receive is posted — with every rank's resolver behind its own gates a ring of such exchanges never starts
    // (found on the software device: wildcards.c with ACX_COOP=1).
    {
        std::vector<Item> plain;
        for (const Item &it : need) {
            if (!it.r->wild) { plain.push_back(it); continue; }
            const bool par = em.parallel;
            em.parallel = false;   // the wait node must follow the resolver node
            AnyParams ap;
            fill_any(it.r, &ap);
            rc = em.launch(recv_any_kernel, dim3(any_grid(it.r)), dim3(kMoverThreads), ap);
            WaitBatch wb;
            memset(&wb, 0, sizeof wb);
            if (!rc) rc = fill_wait(it.r, it.st, retire, false, &wb.w[0]);
            wb.wd = make_wd(); wb.tr = make_tr();
            if (!rc) rc = em.launch(wait_kernel, dim3(plan_wait_grid(wb, 1)), dim3(kWaitThreads), wb);
            em.parallel = par;
            if (rc) return rc;
        }
        need.swap(plain);
    }
    // Cooperative moving: each large rendezvous request gets an agent kernel in front of the waits, so this GPU copies its
    // share of the message even when it arrived first (ACX_COOP).
    for (const Item &it : need) {
        AcxRequest *r = it.r;
        if (!r->coop || r->wild) continue;
        const acx_geom &G = g_acx->geom;
        CoopParams cp;
        memset(&cp, 0, sizeof cp);
        cp.op = (acx_op *)(g_acx->ctrl + G.off_ops) + r->slot;
        cp.is_send = r->kind == ACX_REQ_SEND ? 1 : 0;
        if (cp.is_send) {
            cp.ring = acx_entries_of(g_acx->peer_ctrl[r->peer]) + acx_entry_index(G, g_acx->rank, r->chan, 0);
            cp.sack = acx_sacks_of(g_acx->ctrl) + acx_entry_index(G, r->peer, r->chan, 0);
        } else {
            cp.ring = acx_entries_of(g_acx->ctrl) + acx_entry_index(G, r->peer, r->chan, 0);
            cp.sack = acx_sacks_of(g_acx->peer_ctrl[r->peer]) + acx_entry_index(G, g_acx->rank, r->chan, 0);
        }
        cp.local = r->bounce_off >= 0 ? g_acx->heap + r->bounce_off : r->dev_buf;
        cp.peer_heap = g_acx->peer_heap[r->peer];
        cp.xlat = (const uint64_t *)(g_acx->ctrl + G.off_xlat) + (uint64_t)r->peer * G.maxreg;
        cp.bytes = r->bytes; cp.depth = G.depth; cp.vec256 = (uint32_t)g_mover_256; cp.wd = make_wd();
        const bool par = em.parallel;
        em.parallel = false;
        rc = em.launch(coop_gate, dim3(1), dim3(32), cp);
        if (!rc) rc = em.launch(coop_kernel, dim3(mover_grid_ldst(r->bytes)), dim3(kMoverThreads), cp);
        em.parallel = par;
        if (rc) return rc;
    }
    // Overflow sends go last (see send_overflow_begin): their wait kernels hold this stream until the receiver has consumed
    // earlier messages of the channel, and the receiver may be waiting for exactly the receives this call completes.
    std::stable_partition(need.begin(), need.end(), [](const Item &it) { return !it.r->overflow; });
    size_t first_ovf = need.size();
    for (size_t i = 0; i < need.size(); i++) if (need[i].r->overflow) { first_ovf = i; break; }
    for (size_t base = 0, step = 0; base < need.size(); base += step) {
        WaitBatch wb;
        memset(&wb, 0, sizeof wb);
        int n = (int)std::min((size_t)kWaitBatch, need.size() - base);
        if (base < first_ovf) n = (int)std::min((size_t)n, first_ovf - base);   // no batch mixes the two kinds
        // receives that were posted while their channel's ring was full: their triggers go out now, in posting order (the waits
        // of earlier batches have made room)
        for (int j = 0; j < n; j++) {
            AcxRequest *r = need[base + j].r;
            if (r->kind == ACX_REQ_RECV && r->deferred) {
                const bool par = em.parallel;
                em.parallel = false;
                rc = launch_deferred_recvs(em, r->peer, r->chan, r);
                em.parallel = par;
                if (rc == -2 && j > 0) { n = j; break; }   // no room yet: complete the requests in front of it first (this batch ends here)
                if (rc == -2) {
                    ACX_ERR("receive from rank %d (tag %d) was posted behind %d others that are still un-waited on its channel: wait on earlier receives first (or raise ACX_DEPTH)", r->peer, r->tag, g_acx->geom.depth);
                    return MPI_ERR_OTHER;
                }
                if (rc) return rc;
            }
        }
        std::vector<AcxRequest *> bounced, typed;
        for (int j = 0; j < n; j++) {
            AcxRequest *r = need[base + j].r;
            // Large private destinations: copy-out at full width by a follow-up kernel, which then retires.
            // Strided destinations: unpack_kernel scatters the staged image, then retires.
            const bool is_typed = r->kind == ACX_REQ_RECV && r->typed;
            const bool big_bounce = !is_typed && r->kind == ACX_REQ_RECV && r->bounce_off >= 0 && r->bytes > (256u << 10);
            if ((rc = fill_wait(r, need[base + j].st, retire && !big_bounce && !is_typed, false, &wb.w[j]))) return rc;
            wb.w[j].bounce_kernel = big_bounce ? 1 : 0;
            if (big_bounce) bounced.push_back(r);
            if (is_typed) typed.push_back(r);
        }
        wb.wd = make_wd(); wb.tr = make_tr();
        if (g_memops && g_cuStreamBatchMemOp && qtype == MPIX_QUEUE_CUDA_STREAM && !em.capturing) {
            // Stream memOps on the on-stream path (the reference's cuStreamWaitValue32, src/sendrecv.cu:373-385, polls a host flag
            // over PCIe; here the words live in local HBM): where the host could predict an operation's ticket, the stream
            // waits for its completion word in the memOp unit — no SM spins — and the completion kernel behind it (copy-out of
            // eager payloads, status, slot bookkeeping) finds everything in place. Large sends (which may have to push from
            // their wait) and operations on device-counted channels keep the spinning wait.
            MemOpBatch mb;
            const acx_geom &G = g_acx->geom;
            // A batch with a large-message request keeps the spinning wait for ALL its requests: its wait kernel may have to move
            // data itself (a late send's helper CTAs push, a large receive's pull as a fallback), and a memOp in front of that
            // kernel waiting for another request of the batch would wait for exactly the data the kernel has not moved yet
            // (tests/c/ce_path.c part 3 with ACX_MEMOPS=1: every rank's send posted first — a deadlock on the software device).
            bool movers_in_wait = false;
            for (int j = 0; j < n; j++) if (wb.w[j].late) movers_in_wait = true;
            for (int j = 0; j < n && !movers_in_wait; j++) {
                AcxRequest *r = need[base + j].r;
                if (r->hticket == ~0ull || r->wild || r->overflow) continue;
                const uint32_t e = (uint32_t)(r->hticket % (uint64_t)G.depth);
                const uint64_t g = r->hticket / (uint64_t)G.depth;
                if (r->kind == ACX_REQ_RECV) mb.wait_geq64(&(acx_entries_of(g_acx->ctrl) + acx_entry_index(G, r->peer, r->chan, 0))[e].done, g + 1);
                else if (!r->late) mb.wait_geq64(&(acx_sacks_of(g_acx->ctrl) + acx_entry_index(G, r->peer, r->chan, 0))[e].sdone, g + 1);
            }
            if (!mb.ops.empty() && (rc = mb.flush(em.stream))) return rc;
        }
        if ((rc = em.launch(wait_kernel, dim3(plan_wait_grid(wb, n)), dim3(kWaitThreads), wb))) return rc;
        if (retire) {   // these receives no longer occupy ring slots once this kernel has run: later triggers on the stream may reuse them
            std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
            for (int j = 0; j < n; j++) {
                AcxRequest *r = need[base + j].r;
                if (r->kind == ACX_REQ_RECV && !r->wild && g_acx->recv_outstanding[r->peer][r->chan] > 0) g_acx->recv_outstanding[r->peer][r->chan]--;
            }
        }
        if (retire) for (int j = 0; j < n; j++) {   // post deferred receives of the channels just served as early as there is room
            AcxRequest *r = need[base + j].r;
            if (r->kind == ACX_REQ_RECV && !r->wild && (rc = launch_deferred_recvs(em, r->peer, r->chan, nullptr))) return rc;
        }
        for (AcxRequest *r : typed) {
            UnpackParams up;
            fill_unpack(r, retire, &up);
            const bool par = em.parallel;
            em.parallel = false;  // must follow the wait node
            rc = em.launch(unpack_kernel, dim3(pack_grid(r->bytes, up.L.vec)), dim3(256), up);
            em.parallel = par;
            if (rc) return rc;
        }
        for (AcxRequest *r : bounced) {
            BounceOutParams bp;
            memset(&bp, 0, sizeof bp);
            bp.op = (acx_op *)(g_acx->ctrl + g_acx->geom.off_ops) + r->slot;
            bp.dst = r->dev_buf; bp.bounce = g_acx->heap + r->bounce_off; bp.hstat = g_acx->hstat_d + r->slot; bp.retire = retire ? 1 : 0;
            const bool par = em.parallel;
            em.parallel = false;  // must follow the wait node
            rc = em.launch(bounce_out_kernel, dim3(mover_grid_ldst(r->bytes)), dim3(kMoverThreads), bp);
            em.parallel = par;
            if (rc) return rc;
        }
        step = (size_t)n;
    }
    if (qtype == MPIX_QUEUE_CUDA_GRAPH && !em.last) {
        cudaGraphNode_t node;
        cudaGraphAddEmptyNode(&node, em.graph, nullptr, 0);
    }
    {
        std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
        for (AcxRequest *r : rs) {
            send_flow_release(r);   // waited on: no longer counts against its channel's sends in flight
            if (retire) { g_acx->slot_busy[r->slot] = 3; g_pending_retire.push_back(r->slot); }  // reclaimed once the device marks it retired
        }
    }
    for (int i = 0; i < count; i++) reqs[i] = MPIX_REQUEST_NULL;  // handle is consumed at enqueue time (reference: src/sendrecv.cu:430-432)
    g_acx->st.waits += rs.size();
    if (qtype == MPIX_QUEUE_CUDA_GRAPH) *(cudaGraph_t *)queue = em.graph;
    return MPI_SUCCESS;
}

// Reference: MPIX_Wait_enqueue, src/sendrecv.cu:330-436 (WaitValue32(COMPLETED)+WriteValue32(CLEANUP) or wait kernels; GRAPH path waits
// for the wrong constant, defect D2). MPIX_Waitall_enqueue: src/sendrecv.cu:439-579.
extern "C" int MPIX_Wait_enqueue(MPIX_Request *req, MPI_Status *status, int qtype, void *queue) {
    if (!req) return MPI_ERR_ARG;
    return wait_enqueue_impl(1, req, status == MPI_STATUS_IGNORE ? MPI_STATUSES_IGNORE : status, qtype, queue);
}

extern "C" int MPIX_Waitall_enqueue(int count, MPIX_Request *reqs, MPI_Status *statuses, int qtype, void *queue) {
    return wait_enqueue_impl(count, reqs, statuses, qtype, queue);
}

// =======================================================================================
// MPIX_Send_enqueue / MPIX_Recv_enqueue (extension): blocking on the queue, no request handle
// =======================================================================================
extern "C" int MPIX_Send_enqueue(const void *buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm, int qtype, void *queue) {
    // An eager send is complete when its trigger retires, so the pair below launches exactly one kernel for it;
    // a rendezvous send gets its trigger (+ mover) and a wait, as the two calls would.
    MPIX_Request r = MPIX_REQUEST_NULL;
    int rc = MPIX_Isend_enqueue(buf, count, datatype, dest, tag, comm, &r, qtype, queue);
    if (rc) return rc;
    if (qtype == MPIX_QUEUE_CUDA_GRAPH) {
        // explicit-graph mode returns ONE graph per call: chain the wait's graph behind the send's as child nodes
        cudaGraph_t gs = *(cudaGraph_t *)queue, gw = nullptr, g = nullptr;
        rc = MPIX_Wait_enqueue(&r, MPI_STATUS_IGNORE, qtype, &gw);
        if (rc) { cudaGraphDestroy(gs); return rc; }
        cudaGraphNode_t a, b;
        if (cudaGraphCreate(&g, 0) != cudaSuccess || cudaGraphAddChildGraphNode(&a, g, nullptr, 0, gs) != cudaSuccess || cudaGraphAddChildGraphNode(&b, g, &a, 1, gw) != cudaSuccess) {
            ACX_ERR("MPIX_Send_enqueue: cannot compose the graph");
            if (g) cudaGraphDestroy(g);
            cudaGraphDestroy(gs); cudaGraphDestroy(gw);
            return MPI_ERR_OTHER;
        }
        cudaGraphDestroy(gs); cudaGraphDestroy(gw);   // the child nodes hold clones (with the slots' user objects)
        *(cudaGraph_t *)queue = g;
        return MPI_SUCCESS;
    }
    return MPIX_Wait_enqueue(&r, MPI_STATUS_IGNORE, qtype, queue);
}

extern "C" int MPIX_Recv_enqueue(void *buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Status *status, int qtype, void *queue) {
    AcxRange _r("MPIX_Recv_enqueue");
    // Fast path: device transport, stream queue (capturing or not), contiguous eager-size receive from a specific peer.
    bool fast = g_acx && g_acx->initialized && g_acx->device && qtype == MPIX_QUEUE_CUDA_STREAM && queue && source >= 0 && tag >= 0 && g_branch != 1;
    size_t bytes = 0; int ctx = 0, wsource = 0;
    AcxLayout ty;
    if (fast) {
        MPIX_Request dummy = MPIX_REQUEST_NULL;
        int rc = check_common("MPIX_Recv_enqueue", true, count, datatype, source, tag, comm, &dummy, queue, &bytes, &ctx, &wsource, &ty);
        if (rc) return rc;
        fast = !ty.noncontig && bytes <= g_acx->geom.eager_max && !recv_channel_backlogged(wsource, chan_lookup(wsource, g_acx->rank, tag, ctx));
    }
    if (!fast) {
        MPIX_Request r = MPIX_REQUEST_NULL;
        int rc = MPIX_Irecv_enqueue(buf, count, datatype, source, tag, comm, &r, qtype, queue);
        if (rc) return rc;
        if (qtype == MPIX_QUEUE_CUDA_GRAPH) {
            cudaGraph_t gr = *(cudaGraph_t *)queue, gw = nullptr, g = nullptr;
            rc = MPIX_Wait_enqueue(&r, status, qtype, &gw);
            if (rc) { cudaGraphDestroy(gr); return rc; }
            cudaGraphNode_t a, b;
            if (cudaGraphCreate(&g, 0) != cudaSuccess || cudaGraphAddChildGraphNode(&a, g, nullptr, 0, gr) != cudaSuccess || cudaGraphAddChildGraphNode(&b, g, &a, 1, gw) != cudaSuccess) {
                ACX_ERR("MPIX_Recv_enqueue: cannot compose the graph");
                if (g) cudaGraphDestroy(g);
                cudaGraphDestroy(gr); cudaGraphDestroy(gw);
                return MPI_ERR_OTHER;
            }
            cudaGraphDestroy(gr); cudaGraphDestroy(gw);
            *(cudaGraph_t *)queue = g;
            return MPI_SUCCESS;
        }
        return MPIX_Wait_enqueue(&r, status, qtype, queue);
    }
    AcxRequest *r = new_request(ACX_REQ_RECV, wsource, tag, ctx, bytes, buf);
    r->comm_peer = source;
    source = wsource;
    Emitter em;
    int rc = em.begin(qtype, queue);
    if (rc) { delete r; return rc; }
    r->in_graph = em.in_graph();
    r->chan = chan_lookup(source, g_acx->rank, tag, ctx);
    if (!r->in_graph) {   // the depth check of MPIX_Irecv_enqueue: this receive occupies a ring slot until its kernel has run
        std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
        if (g_acx->recv_outstanding[source][r->chan] >= g_acx->geom.depth) {
            ACX_ERR("more than ACX_DEPTH=%d receives outstanding from rank %d on the channel of tag %d", g_acx->geom.depth, source, tag);
            delete r;
            return MPI_ERR_OTHER;
        }
    }
    if ((r->slot = acx_slot_alloc(r, r->in_graph)) < 0) { delete r; return MPI_ERR_OTHER; }
    if ((rc = acx_device_alias(buf, bytes, &r->dev_buf))) goto fail;
    {
        const acx_geom &G = g_acx->geom;
        r->heap_off = acx_heap_offset_of(buf);
        const uint64_t ring_idx = acx_entry_index(G, source, r->chan, 0), sack_idx = acx_entry_index(G, g_acx->rank, r->chan, 0);
        RecvParams rp;
        memset(&rp, 0, sizeof rp);
        rp.ring = acx_entries_of(g_acx->ctrl) + ring_idx;
        rp.seq = (uint64_t *)(g_acx->ctrl + G.off_recv_seq) + ((uint64_t)source * G.nchan + r->chan);
        rp.op = (acx_op *)(g_acx->ctrl + G.off_ops) + r->slot;
        rp.sack = acx_sacks_of(g_acx->peer_ctrl[source]) + sack_idx;
        rp.peer_heap = g_acx->peer_heap[source];
        rp.xlat = (const uint64_t *)(g_acx->ctrl + G.off_xlat) + (uint64_t)source * G.maxreg;
        rp.dst_off = ~0ull; rp.cap = bytes; rp.serial = r->serial; rp.depth = G.depth; rp.peer = source;
        rp.rndv_capable = 0; rp.publish = g_apush ? 1 : 0;
        rp.wd = make_wd(); rp.tr = make_tr();
        WaitBatch wb;
        memset(&wb, 0, sizeof wb);
        const bool retire = !em.in_graph();
        if ((rc = fill_wait(r, status, retire, false, &wb.w[0]))) goto fail;
        wb.wd = make_wd(); wb.tr = make_tr();
        r->hticket = host_ticket(false, source, r->chan, em.stream, r->in_graph);
        wb.w[0].hticket = r->hticket;
        if ((rc = em.launch2(recv_wait_fused, dim3(1), dim3(kWaitThreads), rp, wb))) goto fail;
    }
    if (em.capturing && (rc = attach_to_graph(em.graph, r->slot))) goto fail;
    {
        std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
        if (!em.in_graph()) { g_acx->slot_busy[r->slot] = 3; g_pending_retire.push_back(r->slot); }   // reclaimed once the kernel marks it retired
    }
    g_acx->st.recvs++; g_acx->st.bytes_recv += bytes; g_acx->st.waits++;
    return MPI_SUCCESS;
fail:
    { std::lock_guard<std::recursive_mutex> lk(g_acx->mu); acx_slot_free(r->slot); }
    return rc;
}

// explicit-graph mode: one graph that runs `first[0..nfirst)` (independent of each other), then `last`; the inputs are consumed
static int compose_graphs(cudaGraph_t *first, int nfirst, cudaGraph_t last, cudaGraph_t *out) {
    cudaGraph_t g = nullptr;
    cudaGraphNode_t nodes[4], tail;
    cudaError_t e = cudaGraphCreate(&g, 0);
    for (int i = 0; i < nfirst && e == cudaSuccess; i++) e = cudaGraphAddChildGraphNode(&nodes[i], g, nullptr, 0, first[i]);
    if (e == cudaSuccess) e = cudaGraphAddChildGraphNode(&tail, g, nodes, (size_t)nfirst, last);
    for (int i = 0; i < nfirst; i++) cudaGraphDestroy(first[i]);   // the child nodes hold clones (with the slots' user objects)
    cudaGraphDestroy(last);
    if (e != cudaSuccess) { ACX_ERR("cannot compose the operation's graph: %s", cudaGetErrorString(e)); if (g) cudaGraphDestroy(g); return MPI_ERR_OTHER; }
    *out = g;
    return MPI_SUCCESS;
}

extern "C" int MPIX_Sendrecv_enqueue(const void *sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag, void *recvbuf, int recvcount, MPI_Datatype recvtype,
                                     int source, int recvtag, MPI_Comm comm, MPI_Status *status, int qtype, void *queue) {
    AcxRange _r("MPIX_Sendrecv_enqueue");
    bool fast = g_acx && g_acx->initialized && g_acx->device && qtype == MPIX_QUEUE_CUDA_STREAM && queue && dest >= 0 && source >= 0 && sendtag >= 0 && recvtag >= 0 &&
                g_branch != 1 && !g_fault_drop_send;
    size_t sbytes = 0, rbytes = 0; int sctx = 0, rctx = 0, wdest = 0, wsource = 0;
    AcxLayout sty, rty;
    if (fast) {
        MPIX_Request dummy = MPIX_REQUEST_NULL;
        int rc = check_common("MPIX_Sendrecv_enqueue", false, sendcount, sendtype, dest, sendtag, comm, &dummy, queue, &sbytes, &sctx, &wdest, &sty);
        if (!rc) rc = check_common("MPIX_Sendrecv_enqueue", true, recvcount, recvtype, source, recvtag, comm, &dummy, queue, &rbytes, &rctx, &wsource, &rty);
        if (rc) return rc;
        fast = !sty.noncontig && !rty.noncontig && sbytes <= g_acx->geom.eager_max && rbytes <= g_acx->geom.eager_max &&
               !recv_channel_backlogged(wsource, chan_lookup(wsource, g_acx->rank, recvtag, rctx)) &&
               !send_channel_backlogged(wdest, chan_lookup(g_acx->rank, wdest, sendtag, sctx));
    }
    if (!fast) {
        MPIX_Request rq[2] = {MPIX_REQUEST_NULL, MPIX_REQUEST_NULL};
        MPI_Status sts[2];
        if (qtype == MPIX_QUEUE_CUDA_GRAPH) {
            cudaGraph_t parts[2] = {nullptr, nullptr}, gw = nullptr;
            int rc = MPIX_Irecv_enqueue(recvbuf, recvcount, recvtype, source, recvtag, comm, &rq[0], qtype, &parts[0]);
            if (rc) return rc;
            rc = MPIX_Isend_enqueue(sendbuf, sendcount, sendtype, dest, sendtag, comm, &rq[1], qtype, &parts[1]);
            if (rc) { cudaGraphDestroy(parts[0]); return rc; }
            rc = MPIX_Waitall_enqueue(2, rq, status == MPI_STATUS_IGNORE ? MPI_STATUSES_IGNORE : sts, qtype, &gw);
            if (rc) { cudaGraphDestroy(parts[0]); cudaGraphDestroy(parts[1]); return rc; }
            if (status != MPI_STATUS_IGNORE) ACX_LOG(1, "MPIX_Sendrecv_enqueue in explicit-graph mode fills no status (the graph outlives the call)");
            return compose_graphs(parts, 2, gw, (cudaGraph_t *)queue);
        }
        int rc = MPIX_Irecv_enqueue(recvbuf, recvcount, recvtype, source, recvtag, comm, &rq[0], qtype, queue);
        if (rc) return rc;
        rc = MPIX_Isend_enqueue(sendbuf, sendcount, sendtype, dest, sendtag, comm, &rq[1], qtype, queue);
        if (rc) return rc;
        if (status == MPI_STATUS_IGNORE) return MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, qtype, queue);
        // the receive's status goes to the caller's word; the send's has no taker
        rc = MPIX_Wait_enqueue(&rq[0], status, qtype, queue);
        if (rc) return rc;
        return MPIX_Wait_enqueue(&rq[1], MPI_STATUS_IGNORE, qtype, queue);
    }
    AcxRequest *r = new_request(ACX_REQ_RECV, wsource, recvtag, rctx, rbytes, recvbuf);
    r->comm_peer = source;
    Emitter em;
    int rc = em.begin(qtype, queue);
    if (rc) { delete r; return rc; }
    r->in_graph = em.in_graph();
    r->stream = em.stream;
    r->chan = chan_lookup(wsource, g_acx->rank, recvtag, rctx);
    if (!r->in_graph) {
        std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
        if (g_acx->recv_outstanding[wsource][r->chan] >= g_acx->geom.depth) {
            ACX_ERR("more than ACX_DEPTH=%d receives outstanding from rank %d on the channel of tag %d", g_acx->geom.depth, wsource, recvtag);
            delete r;
            return MPI_ERR_OTHER;
        }
    }
    if ((r->slot = acx_slot_alloc(r, r->in_graph)) < 0) { delete r; return MPI_ERR_OTHER; }
    char *send_alias = nullptr;
    if ((rc = acx_device_alias(recvbuf, rbytes, &r->dev_buf)) || (rc = acx_device_alias(sendbuf, sbytes, &send_alias))) goto fail;
    fault_delay();
    {
        const acx_geom &G = g_acx->geom;
        r->heap_off = acx_heap_offset_of(recvbuf);
        const int schan = chan_lookup(g_acx->rank, wdest, sendtag, sctx);
        // receive half (as MPIX_Recv_enqueue)
        const uint64_t rring = acx_entry_index(G, wsource, r->chan, 0), rsack = acx_entry_index(G, g_acx->rank, r->chan, 0);
        RecvParams rp;
        memset(&rp, 0, sizeof rp);
        rp.ring = acx_entries_of(g_acx->ctrl) + rring;
        rp.seq = (uint64_t *)(g_acx->ctrl + G.off_recv_seq) + ((uint64_t)wsource * G.nchan + r->chan);
        rp.op = (acx_op *)(g_acx->ctrl + G.off_ops) + r->slot;
        rp.sack = acx_sacks_of(g_acx->peer_ctrl[wsource]) + rsack;
        rp.peer_heap = g_acx->peer_heap[wsource];
        rp.xlat = (const uint64_t *)(g_acx->ctrl + G.off_xlat) + (uint64_t)wsource * G.maxreg;
        rp.dst_off = ~0ull; rp.cap = rbytes; rp.serial = r->serial; rp.depth = G.depth; rp.peer = wsource;
        rp.rndv_capable = 0; rp.publish = g_apush ? 1 : 0;
        rp.wd = make_wd(); rp.tr = make_tr();
        // send half: eager only, so it owns no op record and retires nothing (the kernel's end is its completion)
        const uint64_t sring = acx_entry_index(G, g_acx->rank, schan, 0), ssack = acx_entry_index(G, wdest, schan, 0);
        SendParams sp;
        memset(&sp, 0, sizeof sp);
        sp.ring = acx_entries_of(g_acx->peer_ctrl[wdest]) + sring;
        sp.eager = g_acx->peer_ctrl[wdest] + G.off_eager + sring * G.eager_max;
        sp.sack = acx_sacks_of(g_acx->ctrl) + ssack;
        sp.seq = (uint64_t *)(g_acx->ctrl + G.off_send_seq) + ((uint64_t)wdest * G.nchan + schan);
        sp.op = nullptr; sp.hstat = nullptr;
        sp.src = send_alias; sp.peer_heap = g_acx->peer_heap[wdest];
        sp.xlat = (const uint64_t *)(g_acx->ctrl + G.off_xlat) + (uint64_t)wdest * G.maxreg;
        sp.src_off = ~0ull; sp.bytes = sbytes; sp.serial = 0;
        sp.tag = (uint32_t)sendtag; sp.ctx = (uint32_t)sctx; sp.depth = G.depth; sp.eager_max = G.eager_max; sp.peer = wdest;
        sp.wd = make_wd(); sp.tr = make_tr();
        WaitBatch wb;
        memset(&wb, 0, sizeof wb);
        const bool retire = !em.in_graph();
        if ((rc = fill_wait(r, status, retire, false, &wb.w[0]))) goto fail;
        wb.wd = make_wd(); wb.tr = make_tr();
        r->hticket = host_ticket(false, wsource, r->chan, em.stream, r->in_graph);   // both halves draw a ticket in this kernel
        wb.w[0].hticket = r->hticket;
        (void)host_ticket(true, wdest, schan, em.stream, r->in_graph);
        void *args[] = {&sp, &rp, &wb};
        cudaError_t le;
        if (g_pdl && !em.capturing) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(1); cfg.blockDim = dim3(kWaitThreads); cfg.stream = em.stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            le = cudaLaunchKernelExC(&cfg, (const void *)sendrecv_fused, args);
        } else le = cudaLaunchKernel((const void *)sendrecv_fused, dim3(1), dim3(kWaitThreads), args, 0, em.stream);
        if (le != cudaSuccess) { report_watchdog(); ACX_ERR("kernel launch failed: %s", cudaGetErrorString(le)); rc = MPI_ERR_OTHER; goto fail; }
        g_acx->st.kernels++;
        if (em.capturing) g_acx->st.graph_nodes++;
    }
    if (em.capturing && (rc = attach_to_graph(em.graph, r->slot))) goto fail;
    {
        std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
        if (!em.in_graph()) { g_acx->slot_busy[r->slot] = 3; g_pending_retire.push_back(r->slot); }
    }
    g_acx->st.sends++; g_acx->st.bytes_sent += sbytes; g_acx->st.eager++;
    g_acx->st.recvs++; g_acx->st.bytes_recv += rbytes; g_acx->st.waits += 2;
    return MPI_SUCCESS;
fail:
    { std::lock_guard<std::recursive_mutex> lk(g_acx->mu); acx_slot_free(r->slot); }
    return rc;
}

// ---- partitioned: forward declarations ----
static int pwait_host(AcxRequest *r, MPI_Status *status);
static int pwait_emit(AcxRequest *r, cudaStream_t stream);
static bool pwait_memops(AcxRequest *r, MemOpBatch &mb);

// Reference: MPIX_Wait / MPIX_Waitall, src/sendrecv.cu:582-651 (host spin on the flag; partitioned: spin per partition + MPI_Wait).
extern "C" int MPIX_Wait(MPIX_Request *req, MPI_Status *status) {
    if (!g_acx || !g_acx->initialized || !req) return MPI_ERR_OTHER;
    if (*req == MPIX_REQUEST_NULL) return MPI_SUCCESS;
    AcxRequest *r = as_request(*req);
    if (!r) { ACX_ERR("invalid request handle"); return MPI_ERR_REQUEST; }
    if (r->kind == ACX_REQ_PSEND || r->kind == ACX_REQ_PRECV) return pwait_host(r, status);
    if (r->host_transport) {
        int rc = MPI_Wait(&r->mpi_req, status);
        r->magic = 0; delete r;
        *req = MPIX_REQUEST_NULL;
        return rc;
    }
    if (r->null_op) { complete_null_op(r, status); *req = MPIX_REQUEST_NULL; return MPI_SUCCESS; }
    if (r->in_graph) { ACX_ERR("in-graph operations must be completed by a wait inside a graph"); return MPI_ERR_REQUEST; }
    if (r->kind == ACX_REQ_SEND && r->bytes <= g_acx->geom.eager_max && !r->overflow) {   // (an overflow send's slot may only be released once a wait kernel has seen the trigger's LAST store, op->armed)
        // eager send: complete once its trigger kernel has run; it says so in pinned host memory
        const int32_t *flag = &g_acx->hstat[r->slot].retired;
        auto t0 = std::chrono::steady_clock::now();
        while (!ACX_HOST_FLAG_GET(flag)) {
            if (g_acx->watchdog_ms > 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(g_acx->watchdog_ms)) {
                ACX_ERR("MPIX_Wait: send was never triggered on its stream (waited %ld ms)", g_acx->watchdog_ms);
                return MPI_ERR_OTHER;
            }
            std::this_thread::yield();
        }
    } else {
        if (r->wild) {   // resolve + match first (same internal stream, so the wait kernel below follows it)
            AnyParams ap;
            fill_any(r, &ap);
            void *aargs[] = {&ap};
            ACX_CUDA_WD(cudaLaunchKernel((const void *)recv_any_kernel, dim3(any_grid(r)), dim3(kMoverThreads), aargs, 0, host_stream()));
            g_acx->st.kernels++;
        }
        if (r->kind == ACX_REQ_RECV && r->deferred) {   // posted while its channel's ring was full: its trigger goes out now (internal stream)
            Emitter hem;
            cudaStream_t hs = host_stream();
            int drc = hem.begin(MPIX_QUEUE_CUDA_STREAM, &hs);
            if (!drc) drc = launch_deferred_recvs(hem, r->peer, r->chan, r);
            if (drc == -2) { ACX_ERR("receive from rank %d (tag %d) was posted behind %d others that are still un-waited on its channel: wait on earlier receives first (or raise ACX_DEPTH)", r->peer, r->tag, g_acx->geom.depth); drc = MPI_ERR_OTHER; }
            if (drc) return drc;
        }
        WaitBatch wb;
        memset(&wb, 0, sizeof wb);
        int rc = fill_wait(r, MPI_STATUS_IGNORE, false, true, &wb.w[0]);
        if (rc) return rc;
        wb.wd = make_wd(); wb.tr = make_tr();
        void *args[] = {&wb};
        const int wgrid = plan_wait_grid(wb, 1);   // a late send may have to be pushed from here, a late receive pulled
        ACX_CUDA_WD(cudaLaunchKernel((const void *)wait_kernel, dim3(wgrid), dim3(kWaitThreads), args, 0, host_stream()));
        g_acx->st.kernels++;
        if (r->kind == ACX_REQ_RECV && r->typed) {   // strided destination: scatter the staged image
            UnpackParams up;
            fill_unpack(r, false, &up);
            void *uargs[] = {&up};
            ACX_CUDA_WD(cudaLaunchKernel((const void *)unpack_kernel, dim3(pack_grid(r->bytes, up.L.vec)), dim3(256), uargs, 0, host_stream()));
            g_acx->st.kernels++;
        }
        ACX_CUDA_WD(cudaStreamSynchronize(host_stream()));
    }
    if (status != MPI_STATUS_IGNORE) {
        if (r->kind == ACX_REQ_SEND) host_status_send(status, r);
        else {
            const acx_host_status &h = g_acx->hstat[r->slot];
            status->MPI_SOURCE = h.source; status->MPI_TAG = h.tag; status->MPI_ERROR = h.error;
#ifdef ACX_MINI_MPI
            status->_bytes = h.bytes;
#endif
        }
    }
    if (r->ev_done) ACX_CUDA_WD(cudaEventSynchronize(r->ev_done));   // the side-stream mover still reads the op record (see side_branch_finished)
    {
        std::lock_guard<std::recursive_mutex> lk(g_acx->mu);
        if (r->kind == ACX_REQ_RECV && !r->wild && g_acx->recv_outstanding[r->peer][r->chan] > 0) g_acx->recv_outstanding[r->peer][r->chan]--;
        acx_slot_free(r->slot);
    }
    *req = MPIX_REQUEST_NULL;
    g_acx->st.waits++;
    return MPI_SUCCESS;
}

extern "C" int MPIX_Waitall(int count, MPIX_Request *reqs, MPI_Status *statuses) {
    // partitioned requests on the device transport: emit every completion kernel first, synchronize once
    if (g_acx && g_acx->device) {
        bool any = false;
        MemOpBatch wait_ops;
        PWaitBatch pb;
        int nb = 0;
        auto flush = [&]() -> int {
            if (!nb) return 0;
            pb.wd = make_wd();
            void *args[] = {&pb};
            ACX_CUDA_WD(cudaLaunchKernel((const void *)pwait_batch_kernel, dim3(nb), dim3(kWaitThreads), args, 0, host_stream()));
            g_acx->st.kernels++;
            nb = 0;
            return 0;
        };
        for (int i = 0; i < count; i++) {
            AcxRequest *r = reqs[i] ? as_request(reqs[i]) : nullptr;
            if (r && (r->kind == ACX_REQ_PSEND || r->kind == ACX_REQ_PRECV) && !r->host_transport && r->active) {
                if (pwait_memops(r, wait_ops)) {   // one cuStreamBatchMemOp for every request of the group that needs no copy-out
                    r->active = false;
                } else if (g_pbatch) {
                    pb.ch[nb] = r->pchan_d; pb.e[nb] = r->epoch;
                    r->active = false;
                    int rc;
                    if (++nb == kPBatch && (rc = flush())) return rc;
                } else {
                    int rc = pwait_emit(r, host_stream());
                    if (rc) return rc;
                }
                any = true;
            }
        }
        { int rc = flush(); if (rc) return rc; }
        if (!wait_ops.ops.empty()) { int rc = wait_ops.flush(host_stream()); if (rc) return rc; }
        if (any) ACX_CUDA_WD(cudaStreamSynchronize(host_stream()));
    }
    // Overflow sends last, as in the stream waits (send_overflow_begin): they complete only once the peer has consumed earlier
    // messages of their channel, and the peer's host may be sitting in the mirror image of this very call.
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < count; i++) {
            AcxRequest *r = reqs[i] ? as_request(reqs[i]) : nullptr;
            if (pass == 0 && r && r->overflow) continue;
            int rc = MPIX_Wait(&reqs[i], statuses == MPI_STATUSES_IGNORE ? MPI_STATUS_IGNORE : &statuses[i]);
            if (rc) return rc;
        }
    }
    return MPI_SUCCESS;
}

// =======================================================================================
// partitioned
// =======================================================================================
struct PInfo {  // handshake record exchanged over the private communicator
    int32_t partitions;
    int32_t pad;
    uint64_t part_bytes;
    int64_t land_off;       // receiver: landing buffer offset in its heap
    uint32_t pf_arrived;    // receiver: arrived[] index in its pflags
    uint32_t pf_cts;        // sender: cts word index in its pflags
};

struct PHandshake { PInfo mine, theirs; MPI_Request sreq, rreq; };
static std::unordered_map<AcxRequest *, PHandshake *> g_hs;

static int pinit_common(AcxReqKind kind, const void *buf, int partitions, MPI_Count count, MPI_Datatype dt, int peer, int tag, MPI_Comm comm,
                        MPIX_Request *request) {
    if (!g_acx || !g_acx->initialized) return MPI_ERR_OTHER;
    if (!request) return MPI_ERR_ARG;
    if (partitions <= 0) { ACX_ERR("partitions must be > 0"); return MPI_ERR_COUNT; }
    int sz = 0, ctx = 0;
    if (MPI_Type_size(dt, &sz) != MPI_SUCCESS || sz <= 0) return MPI_ERR_TYPE;
    if (tag < 0) return MPI_ERR_TAG;
    const AcxCommMap *cm = nullptr;
    int rc = acx_comm_ctx(comm, &ctx, &cm);
    if (rc) return rc;
    {
        AcxLayout ty;
        if ((rc = describe_type("MPIX_Psend/Precv_init", dt, (int)count, sz, &ty))) return rc;
        if (ty.noncontig) { ACX_ERR("partitioned operations take contiguous datatypes (a partition is a byte range of the buffer)"); return MPI_ERR_TYPE; }
    }
    if (peer < 0 || peer >= cm->size) return MPI_ERR_RANK;
    const int cpeer = peer;                 // as named by the caller (host transport, status)
    const int wpeer = cm->to_world[(size_t)peer];
    size_t part_bytes = (size_t)count * (size_t)sz;
    AcxRequest *r = new_request(kind, wpeer, tag, ctx, part_bytes * partitions, buf);
    r->comm_peer = cpeer;
    r->partitions = partitions; r->part_bytes = part_bytes;
    if (!g_acx->device) {
        r->host_transport = true;
        rc = kind == ACX_REQ_PSEND ? MPI_Psend_init(buf, partitions, count, dt, cpeer, tag, comm, MPI_INFO_NULL, &r->mpi_req)
                                   : MPI_Precv_init((void *)buf, partitions, count, dt, cpeer, tag, comm, MPI_INFO_NULL, &r->mpi_req);
        if (rc) { delete r; return rc; }
        *request = r;
        return MPI_SUCCESS;
    }
    peer = wpeer;   // device transport: world rank from here on (the handshake runs on a dup of MPI_COMM_WORLD)
    if ((rc = acx_device_alias(buf, r->bytes, &r->dev_buf))) { delete r; return rc; }
    PHandshake *hs = new PHandshake();
    memset(hs, 0, sizeof *hs);
    hs->mine.partitions = partitions; hs->mine.part_bytes = part_bytes; hs->mine.land_off = -1;
    if (kind == ACX_REQ_PSEND) {
        uint32_t base = acx_pflags_alloc(2 * partitions + 2);   // sent[p], cts, epoch, work[p] (CTA counters of the grid-cooperative Pready)
        if (base == UINT32_MAX) { delete r; delete hs; return MPI_ERR_OTHER; }
        r->pf_local = base; r->pf_cts = base + partitions; r->pf_epoch = base + partitions + 1; r->pf_work = base + partitions + 2;
        hs->mine.pf_cts = r->pf_cts;
    } else {
        uint32_t base = acx_pflags_alloc(2 * partitions + 1);
        if (base == UINT32_MAX) { delete r; delete hs; return MPI_ERR_OTHER; }
        r->pf_arrived = base; r->pf_local = base + partitions; r->pf_epoch = base + 2 * partitions;
        r->heap_off = acx_heap_offset_of(buf);
        if (r->heap_off >= 0) { r->land_off = r->heap_off; r->land_is_user = true; }
        else {
            if ((r->land_off = acx_heap_alloc(r->bytes)) < 0) { delete r; delete hs; return MPI_ERR_OTHER; }
            g_acx->st.bounce++;
        }
        hs->mine.land_off = r->land_off; hs->mine.pf_arrived = r->pf_arrived;
    }
    // Exchange descriptors on the private communicator; matching order per (peer, tag)
    // follows MPI's own non-overtaking rule, which is exactly the init-order pairing we need.
    // (per communicator: the context id occupies the tag bits above the user's tag range)
    const int t_base = (int)(((unsigned)ctx & 0x3fu) << 24) | ((tag & 0x7fffff) << 1);
    const int t_out = t_base + (kind == ACX_REQ_PSEND ? 0 : 1), t_in = t_base + (kind == ACX_REQ_PSEND ? 1 : 0);
    MPI_Irecv(&hs->theirs, (int)sizeof(PInfo), MPI_BYTE, peer, t_in, g_acx->icomm, &hs->rreq);
    MPI_Isend(&hs->mine, (int)sizeof(PInfo), MPI_BYTE, peer, t_out, g_acx->icomm, &hs->sreq);
    { std::lock_guard<std::recursive_mutex> lk(g_acx->mu); g_hs[r] = hs; }
    *request = r;
    return MPI_SUCCESS;
}

// Reference: MPIX_Psend_init / MPIX_Precv_init, src/partitioned.cu:36-123 (MPI_Psend_init/MPI_Precv_init + one flag slot per partition).
extern "C" int MPIX_Psend_init(const void *buf, int partitions, MPI_Count count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm, MPI_Info,
                               MPIX_Request *request) {
    return pinit_common(ACX_REQ_PSEND, buf, partitions, count, datatype, dest, tag, comm, request);
}
extern "C" int MPIX_Precv_init(void *buf, int partitions, MPI_Count count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Info,
                               MPIX_Request *request) {
    return pinit_common(ACX_REQ_PRECV, buf, partitions, count, datatype, source, tag, comm, request);
}

// Complete the handshake (blocks until the peer has called its matching *_init) and build
// the device descriptor.
static int presolve(AcxRequest *r) {
    if (r->resolved) return 0;
    CaptureRelax relax;   // cudaMalloc + cudaMemcpy of the device descriptor
    PHandshake *hs;
    { std::lock_guard<std::recursive_mutex> lk(g_acx->mu); hs = g_hs[r]; }
    MPI_Wait(&hs->sreq, MPI_STATUS_IGNORE);
    MPI_Wait(&hs->rreq, MPI_STATUS_IGNORE);
    if (hs->theirs.partitions != r->partitions || hs->theirs.part_bytes != r->part_bytes) {
        ACX_ERR("partitioned send/recv geometry mismatch with rank %d: mine %d x %zu B, theirs %d x %llu B (equal partitioning is required)", r->peer,
                r->partitions, r->part_bytes, hs->theirs.partitions, (unsigned long long)hs->theirs.part_bytes);
        return MPI_ERR_OTHER;
    }
    acx_pchan &c = r->pchan_h;
    memset(&c, 0, sizeof c);
    c.partitions = r->partitions; c.part_bytes = r->part_bytes; c.user_buf = r->dev_buf; c.peer = r->peer; c.tag = r->tag;
    uint64_t *mypf = acx_pflags_of(g_acx->ctrl), *peerpf = acx_pflags_of(g_acx->peer_ctrl[r->peer]);
    c.epoch = mypf + r->pf_epoch;
    if (r->kind == ACX_REQ_PSEND) {
        c.kind = ACX_PKIND_SEND;
        c.land = g_acx->peer_heap[r->peer] + hs->theirs.land_off;
        c.arrived = peerpf + hs->theirs.pf_arrived;
        c.local_flags = mypf + r->pf_local;
        c.cts = mypf + r->pf_cts;
        c.work = mypf + r->pf_work;
    } else {
        c.kind = ACX_PKIND_RECV;
        c.land = r->land_is_user ? r->dev_buf : g_acx->heap + r->land_off;
        c.arrived = mypf + r->pf_arrived;
        c.local_flags = mypf + r->pf_local;
        c.cts = peerpf + hs->theirs.pf_cts;  // remote: written at Start
    }
    ACX_CUDA(cudaMalloc((void **)&r->pchan_d, sizeof(acx_pchan)));
    if (acx_h2d(r->pchan_d, &c, sizeof c, host_stream())) return MPI_ERR_OTHER;
    { std::lock_guard<std::recursive_mutex> lk(g_acx->mu); g_hs.erase(r); }
    delete hs;
    r->resolved = true;
    return 0;
}

// Reference: MPIX_Prequest_create / _free, src/partitioned.cu:160-197 (device copy of {idx[], flags}).
extern "C" int MPIX_Prequest_create(MPIX_Request request, MPIX_Prequest *prequest) {
    AcxRequest *r = as_request(request);
    if (!r || (r->kind != ACX_REQ_PSEND && r->kind != ACX_REQ_PRECV) || !prequest) return MPI_ERR_REQUEST;
    if (!g_acx->device) { ACX_ERR("device partitioned handles need a GPU"); return MPI_ERR_OTHER; }
    int rc = presolve(r);
    if (rc) return rc;
    *prequest = (MPIX_Prequest)r->pchan_d;
    return MPI_SUCCESS;
}

extern "C" int MPIX_Prequest_free(MPIX_Prequest *preq) {
    if (!preq) return MPI_ERR_ARG;
    *preq = MPIX_PREQUEST_NULL;  // the descriptor is owned by the request (freed in MPIX_Request_free)
    return MPI_SUCCESS;
}

static int pstart_emit(AcxRequest *r, cudaStream_t stream) {
    int rc = presolve(r);
    if (rc) return rc;
    if (r->active) { ACX_ERR("MPIX_Start on an active partitioned request"); return MPI_ERR_REQUEST; }
    r->epoch++;
    r->active = true;
    uint64_t *epoch_local = r->pchan_h.epoch;
    uint64_t *cts_remote = r->kind == ACX_REQ_PRECV ? r->pchan_h.cts : nullptr;
    uint64_t e = r->epoch;
    if (g_memops && !stream_is_capturing(stream)) {
        MemOpBatch mb;
        mb.write64(epoch_local, e);
        if (cts_remote) mb.write64(cts_remote, e);   // peer-mapped: clear-to-send at the sender
        return mb.flush(stream);
    }
    void *args[] = {&epoch_local, &cts_remote, &e};
    ACX_CUDA_WD(cudaLaunchKernel((const void *)pstart_kernel, dim3(1), dim3(1), args, 0, stream));
    g_acx->st.kernels++;
    return 0;
}

// Reference: MPIX_Start / MPIX_Startall, src/partitioned.cu:125-157 (MPI_Start; recv flags -> ISSUED so the proxy polls MPI_Parrived).
extern "C" int MPIX_Startall(int count, MPIX_Request *reqs) {
    if (!g_acx || !g_acx->initialized) return MPI_ERR_OTHER;
    bool any_dev = false;
    MemOpBatch start_ops;
    PStartBatch sb;
    memset(&sb, 0, sizeof sb);
    auto flush = [&]() -> int {
        if (!sb.n) return 0;
        void *args[] = {&sb};
        ACX_CUDA_WD(cudaLaunchKernel((const void *)pstart_batch_kernel, dim3(1), dim3(32), args, 0, host_stream()));
        g_acx->st.kernels++;
        sb.n = 0;
        return 0;
    };
    for (int i = 0; i < count; i++) {
        AcxRequest *r = as_request(reqs[i]);
        if (!r || (r->kind != ACX_REQ_PSEND && r->kind != ACX_REQ_PRECV)) { ACX_ERR("MPIX_Start needs a partitioned request"); return MPI_ERR_REQUEST; }
        if (r->host_transport) { int rc = MPI_Start(&r->mpi_req); if (rc) return rc; r->active = true; continue; }
        if (g_memops) {   // the whole group becomes one cuStreamBatchMemOp (flushed below)
            int rc = presolve(r);
            if (rc) return rc;
            if (r->active) { ACX_ERR("MPIX_Start on an active partitioned request"); return MPI_ERR_REQUEST; }
            r->epoch++;
            r->active = true;
            start_ops.write64(r->pchan_h.epoch, r->epoch);
            if (r->kind == ACX_REQ_PRECV) start_ops.write64(r->pchan_h.cts, r->epoch);
            any_dev = true;
            continue;
        }
        if (g_pbatch) {
            int rc = presolve(r);
            if (rc) return rc;
            if (r->active) { ACX_ERR("MPIX_Start on an active partitioned request"); return MPI_ERR_REQUEST; }
            r->epoch++;
            r->active = true;
            sb.epoch_local[sb.n] = r->pchan_h.epoch;
            sb.cts_remote[sb.n] = r->kind == ACX_REQ_PRECV ? r->pchan_h.cts : nullptr;
            sb.e[sb.n] = r->epoch;
            if (++sb.n == kPBatch && (rc = flush())) return rc;
            any_dev = true;
            continue;
        }
        int rc = pstart_emit(r, host_stream());
        if (rc) return rc;
        any_dev = true;
    }
    { int rc = flush(); if (rc) return rc; }
    if (!start_ops.ops.empty()) { int rc = start_ops.flush(host_stream()); if (rc) return rc; }
    // Kernels launched by the user right after Start (on any stream) must observe the new epoch.
    if (any_dev) ACX_CUDA_WD(cudaStreamSynchronize(host_stream()));
    return MPI_SUCCESS;
}
extern "C" int MPIX_Start(MPIX_Request *request) { return MPIX_Startall(1, request); }

extern "C" int MPIX_Start_enqueue(MPIX_Request *request, int qtype, void *queue) {
    AcxRequest *r = request ? as_request(*request) : nullptr;
    if (!r || r->host_transport || qtype != MPIX_QUEUE_CUDA_STREAM) return MPI_ERR_REQUEST;
    return pstart_emit(r, *(cudaStream_t *)queue);
}

// Completion of a round as memOp waits, when nothing has to be copied: a send waits for sent[p] >= e, a receive whose
// landing zone IS the user's buffer for arrived[p] >= e. Returns false (nothing appended) when a kernel is needed.
static bool pwait_memops(AcxRequest *r, MemOpBatch &mb) {
    if (!g_memops || (size_t)r->partitions > kMemOpBatchMax) return false;
    const acx_pchan &c = r->pchan_h;
    if (r->kind == ACX_REQ_PSEND) { for (int p = 0; p < r->partitions; p++) mb.wait_geq64(c.local_flags + p, r->epoch); return true; }
    if (c.land != c.user_buf) return false;   // private destination: pwait_kernel copies out what device pollers did not claim
    for (int p = 0; p < r->partitions; p++) mb.wait_geq64(c.arrived + p, r->epoch);
    return true;
}

static int pwait_emit(AcxRequest *r, cudaStream_t stream) {
    if (!r->active) return 0;
    if (g_memops && !stream_is_capturing(stream)) {
        MemOpBatch mb;
        if (pwait_memops(r, mb)) { r->active = false; return mb.flush(stream); }
    }
    const acx_pchan *ch = r->pchan_d;
    uint64_t e = r->epoch;
    Watchdog wd = make_wd();
    void *args[] = {&ch, &e, &wd};
    ACX_CUDA_WD(cudaLaunchKernel((const void *)pwait_kernel, dim3(1), dim3(kWaitThreads), args, 0, stream));
    g_acx->st.kernels++;
    r->active = false;
    return 0;
}

extern "C" int MPIX_Pwait_enqueue(MPIX_Request *request, int qtype, void *queue) {
    AcxRequest *r = request ? as_request(*request) : nullptr;
    if (!r || r->host_transport || qtype != MPIX_QUEUE_CUDA_STREAM) return MPI_ERR_REQUEST;
    return pwait_emit(r, *(cudaStream_t *)queue);
}

static int pwait_host(AcxRequest *r, MPI_Status *status) {
    if (r->host_transport) {
        int rc = MPI_Wait(&r->mpi_req, status);
        r->active = false;
        return rc;
    }
    int rc = pwait_emit(r, host_stream());
    if (rc) return rc;
    ACX_CUDA_WD(cudaStreamSynchronize(host_stream()));
    if (status != MPI_STATUS_IGNORE) {
        status->MPI_SOURCE = r->comm_peer; status->MPI_TAG = r->tag; status->MPI_ERROR = MPI_SUCCESS;
#ifdef ACX_MINI_MPI
        status->_bytes = r->kind == ACX_REQ_PRECV ? (long long)r->bytes : 0;
#endif
    }
    return MPI_SUCCESS;
}

// Reference: MPIX_Request_free, src/sendrecv.cu:654-682 (partitioned only; basic requests silently ignored, defect D11).
extern "C" int MPIX_Request_free(MPIX_Request *request) {
    if (!request || *request == MPIX_REQUEST_NULL) return MPI_SUCCESS;
    AcxRequest *r = as_request(*request);
    if (!r) return MPI_ERR_REQUEST;
    if (r->kind == ACX_REQ_SEND || r->kind == ACX_REQ_RECV) {
        // The reference silently ignores basic requests here (src/sendrecv.cu:667); we do the
        // sensible thing: an un-waited basic request is completed and released.
        return MPIX_Wait(request, MPI_STATUS_IGNORE);
    }
    if (r->active) { ACX_ERR("MPIX_Request_free on an active partitioned request"); return MPI_ERR_REQUEST; }
    if (r->host_transport) {
        MPI_Request_free(&r->mpi_req);
    } else {
        if (!r->resolved) presolve(r);  // drain the handshake so the peer's messages are matched
        if (r->pchan_d) cudaFree(r->pchan_d);
        if (r->kind == ACX_REQ_PRECV && !r->land_is_user && r->land_off >= 0) acx_heap_free(r->land_off);
    }
    r->magic = 0;
    delete r;
    *request = MPIX_REQUEST_NULL;
    return MPI_SUCCESS;
}

// Reference: host side of MPIX_Pready / MPIX_Parrived, src/partitioned.cu:206-231.
extern "C" int MPIX_Pready_host(int partition, MPIX_Request *request) {
    AcxRequest *r = request ? as_request(*request) : nullptr;
    if (!r || r->kind != ACX_REQ_PSEND) return MPI_ERR_REQUEST;
    if (partition < 0 || partition >= r->partitions) return MPI_ERR_ARG;
    if (r->host_transport) return MPI_Pready(partition, r->mpi_req);
    const acx_pchan *ch = r->pchan_d;
    void *args[] = {&ch, &partition};
    ACX_CUDA_WD(cudaLaunchKernel((const void *)pready_host_kernel, dim3(1), dim3(256), args, 0, host_stream()));
    ACX_CUDA_WD(cudaStreamSynchronize(host_stream()));
    return MPI_SUCCESS;
}

extern "C" int MPIX_Parrived_host(MPIX_Request *request, int partition, int *flag) {
    AcxRequest *r = request ? as_request(*request) : nullptr;
    if (!r || r->kind != ACX_REQ_PRECV || !flag) return MPI_ERR_REQUEST;
    if (partition < 0 || partition >= r->partitions) return MPI_ERR_ARG;
    if (r->host_transport) return MPI_Parrived(r->mpi_req, partition, flag);
    if (!r->active) { *flag = 1; return MPI_SUCCESS; }
    thread_local int *dflag = nullptr;   // one pinned word per calling thread (never freed: a few bytes per thread)
    if (!dflag) ACX_CUDA(cudaHostAlloc((void **)&dflag, sizeof(int), cudaHostAllocMapped));
    *dflag = 0;
    const acx_pchan *ch = r->pchan_d;
    int *dp = nullptr;
    cudaHostGetDevicePointer((void **)&dp, dflag, 0);
    void *args[] = {&ch, &partition, &dp};
    ACX_CUDA_WD(cudaLaunchKernel((const void *)parrived_host_kernel, dim3(1), dim3(1), args, 0, host_stream()));
    ACX_CUDA_WD(cudaStreamSynchronize(host_stream()));
    *flag = *dflag;
    return MPI_SUCCESS;
}

#ifdef ACX_EMU
// Software device (tests/emu): kernels are plain functions; tell its cudaLaunchKernel / cudaGraphAddKernelNode about
// their argument lists. All of them are written for any blockDim (second argument), so the emulation may shrink blocks.
ACX_EMU_KERNEL(copy_kernel, true) ACX_EMU_KERNEL(send_trigger, true) ACX_EMU_KERNEL(recv_trigger, true) ACX_EMU_KERNEL(push_gate, true) ACX_EMU_KERNEL(ce_gate, true) ACX_EMU_KERNEL(ce_complete, true)
ACX_EMU_KERNEL(mover, true) ACX_EMU_KERNEL(mover_plain, true) ACX_EMU_KERNEL(mover_v256, true) ACX_EMU_KERNEL(mover_cs, true) ACX_EMU_KERNEL(send_rndv_fused, true) ACX_EMU_KERNEL(recv_rndv_fused, true) ACX_EMU_KERNEL(recv_any_kernel, true)
ACX_EMU_KERNEL(coop_gate, true) ACX_EMU_KERNEL(coop_kernel, true) ACX_EMU_KERNEL(wait_kernel, true) ACX_EMU_KERNEL(recv_wait_fused, true) ACX_EMU_KERNEL(sendrecv_fused, true) ACX_EMU_KERNEL(bounce_out_kernel, true)
ACX_EMU_KERNEL(pack_kernel, true) ACX_EMU_KERNEL(unpack_kernel, true) ACX_EMU_KERNEL(pstart_kernel, true) ACX_EMU_KERNEL(pstart_batch_kernel, true)
ACX_EMU_KERNEL(pwait_kernel, true) ACX_EMU_KERNEL(pwait_batch_kernel, true) ACX_EMU_KERNEL(pready_host_kernel, true) ACX_EMU_KERNEL(parrived_host_kernel, true)
#endif
