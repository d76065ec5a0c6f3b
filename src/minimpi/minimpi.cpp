// acx mini-MPI implementation (see include/mpi.h for the rationale).
//
// Single node, N processes, one POSIX shm segment per job:
//   [Header | per-rank blobs | shared scratch | N*N SPSC cell rings]
// P2P: every ordered pair (src,dst) has a single-producer/single-consumer ring of
// fixed-size cells. A message is a HDR cell followed by DATA cells streamed through
// the ring (pipelined copy-in/copy-out, like the shm BTL/CH4 shm eager path). Device
// buffers: small -> staged through the cells with cudaMemcpy; large -> CUDA-IPC
// rendezvous (RTS cell carries the cudaIpcMemHandle, receiver issues cudaMemcpyAsync
// from the mapped peer buffer on a private non-blocking stream, then FINs).
// Matching: MPI (source, tag, comm) with wildcards, posted-receive and unexpected
// queues, non-overtaking per pair. Progress is made only inside MPI calls (no helper
// thread), under one global mutex (MPI_THREAD_MULTIPLE is honoured that way).
//
// All CUDA use is optional and resolved at run time (dlopen of libcudart) so the same
// library serves GPU-less CI.
#include <algorithm>
#include <mpi.h>

#include <atomic>
#include <cerrno>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <list>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include <dlfcn.h>
#include <fcntl.h>
#include <sched.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <time.h>
#include <unistd.h>

namespace {

// ---------------------------------------------------------------------------------
// Optional CUDA runtime (dlopen) — only what a CUDA-aware intra-node MPI needs.
// ---------------------------------------------------------------------------------
struct CudaRt {
    bool tried = false, ok = false;
    void *h = nullptr;
    int (*PointerGetAttributes)(void *attr, const void *ptr) = nullptr;
    int (*Memcpy)(void *, const void *, size_t, int) = nullptr;
    int (*MemcpyAsync)(void *, const void *, size_t, int, void *) = nullptr;
    int (*StreamCreateWithFlags)(void **, unsigned) = nullptr;
    int (*StreamSynchronize)(void *) = nullptr;
    int (*StreamQuery)(void *) = nullptr;
    int (*EventCreateWithFlags)(void **, unsigned) = nullptr;
    int (*EventRecord)(void *, void *) = nullptr;
    int (*EventQuery)(void *) = nullptr;
    int (*EventDestroy)(void *) = nullptr;
    int (*IpcGetMemHandle)(void *, void *) = nullptr;
    int (*GetDeviceCount)(int *) = nullptr;
    int (*GetLastError)() = nullptr;
    int (*GetDevice)(int *) = nullptr;
    int (*SetDevice)(int) = nullptr;
    int (*Free)(void *) = nullptr;
    void *stream = nullptr;
};
CudaRt g_cuda;

struct IpcHandle { char b[64]; };
typedef int (*IpcOpenFn)(void **, IpcHandle, unsigned);
IpcOpenFn g_ipc_open = nullptr;
typedef int (*MemGetAddressRangeFn)(unsigned long long *, size_t *, unsigned long long);

// cudaPointerAttributes layout (CUDA 11+/12): {int type; int device; void* devicePointer; void* hostPointer;}
struct PtrAttr { int type; int device; void *dptr; void *hptr; long pad[4]; };

bool cuda_load() {
    if (g_cuda.tried) return g_cuda.ok;
    g_cuda.tried = true;
    if (getenv("ACX_MPI_NO_CUDA")) return false;
    const char *names[] = {"libcudart.so.12", "libcudart.so", "libcudart.so.13", nullptr};
    for (int i = 0; names[i] && !g_cuda.h; i++) g_cuda.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!g_cuda.h) return false;
#define L(field, sym) *(void **)(&g_cuda.field) = dlsym(g_cuda.h, sym); if (!g_cuda.field) return false;
    L(PointerGetAttributes, "cudaPointerGetAttributes")
    L(Memcpy, "cudaMemcpy")
    L(MemcpyAsync, "cudaMemcpyAsync")
    L(StreamCreateWithFlags, "cudaStreamCreateWithFlags")
    L(StreamSynchronize, "cudaStreamSynchronize")
    L(StreamQuery, "cudaStreamQuery")
    L(EventCreateWithFlags, "cudaEventCreateWithFlags")
    L(EventRecord, "cudaEventRecord")
    L(EventQuery, "cudaEventQuery")
    L(EventDestroy, "cudaEventDestroy")
    L(IpcGetMemHandle, "cudaIpcGetMemHandle")
    L(GetDeviceCount, "cudaGetDeviceCount")
    L(GetLastError, "cudaGetLastError")
    L(GetDevice, "cudaGetDevice")
    L(SetDevice, "cudaSetDevice")
    L(Free, "cudaFree")
#undef L
    g_ipc_open = (IpcOpenFn)dlsym(g_cuda.h, "cudaIpcOpenMemHandle");
    if (!g_ipc_open) return false;
    int n = 0;
    if (g_cuda.GetDeviceCount(&n) != 0 || n <= 0) { g_cuda.GetLastError(); return false; }
    g_cuda.ok = true;
    return true;
}

enum { kMemcpyDefault = 4 };

// MPI calls may come from threads that never touched CUDA (e.g. the reference's proxy
// thread): bind the calling thread to the device that owns the buffer, as CUDA-aware MPIs do.
thread_local int t_bound_dev = -1;
void bind_thread_to(int dev) {
    if (dev < 0 || dev == t_bound_dev) return;
    g_cuda.SetDevice(dev);
    g_cuda.Free(nullptr);  // force the primary context current on this thread (driver calls follow)
    t_bound_dev = dev;
}

// Returns true if p is device (or managed) memory the CPU must not touch directly.
bool is_device_ptr(const void *p) {
    if (!p || !cuda_load()) return false;
    PtrAttr a;
    memset(&a, 0, sizeof(a));
    if (g_cuda.PointerGetAttributes(&a, p) != 0) { g_cuda.GetLastError(); return false; }
    if (a.type == 2 /*cudaMemoryTypeDevice*/) { bind_thread_to(a.device); return true; }
    return false;
}

void *cuda_stream() {
    static void *streams[64] = {nullptr};
    int d = t_bound_dev < 0 ? 0 : t_bound_dev & 63;
    if (t_bound_dev < 0) { int cur = 0; if (g_cuda.GetDevice(&cur) == 0) { d = cur & 63; } }
    if (!streams[d]) g_cuda.StreamCreateWithFlags(&streams[d], 1 /*NonBlocking*/);
    return streams[d];
}

void die(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "[acx-mpi] fatal: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    _exit(86);
}

// ---------------------------------------------------------------------------------
// Shared segment layout
// ---------------------------------------------------------------------------------
constexpr uint32_t kMagic = 0xAC0B2001u;
constexpr size_t kBlobBytes = 8192;         // per-rank exchange blob
constexpr size_t kScratchBytes = 8u << 20;  // libacx control area (directories, registries)

enum CellKind : uint32_t { CELL_HDR = 1, CELL_DATA = 2, CELL_RTS_IPC = 3, CELL_FIN = 4 };

struct CellHdr {
    uint32_t kind;
    uint32_t len;       // payload bytes in this cell
    int32_t tag;
    int32_t comm;
    int32_t part;       // partition index for partitioned traffic, -1 otherwise
    uint32_t pad;
    uint64_t total;     // message bytes (HDR / RTS)
    uint64_t cookie;    // sender request cookie (RTS/FIN)
};

struct Ring {
    alignas(64) std::atomic<uint64_t> head;  // producer cursor (cells written)
    alignas(64) std::atomic<uint64_t> tail;  // consumer cursor (cells consumed)
};

struct Header {
    std::atomic<uint32_t> magic;
    uint32_t nranks;
    uint32_t ncell;
    uint32_t cell_bytes;  // payload bytes per cell
    int32_t creator_pid;
    std::atomic<int32_t> abort_code;
    alignas(64) std::atomic<uint32_t> checkin;
    alignas(64) std::atomic<uint32_t> bar_count;
    alignas(64) std::atomic<uint32_t> bar_sense;
    alignas(64) std::atomic<int32_t> pids[1024];
};

struct Seg {
    Header *hdr = nullptr;
    char *blobs = nullptr;
    char *scratch = nullptr;
    Ring *rings = nullptr;
    char *cells = nullptr;
    size_t bytes = 0;
    size_t cell_stride = 0;
    std::string name;
};

struct State {
    bool initialized = false, finalized = false;
    int rank = 0, size = 1;
    int thread_level = MPI_THREAD_MULTIPLE;
    std::string jobid;
    Seg seg;
    uint32_t my_sense = 0;
    int next_comm = 16;
    int next_group = 16;
};
// A communicator: matching context + ordered member list (world ranks). Handles are small integers agreed on by all
// members when the communicator is created (max over the members' private counters), so a handle doubles as a
// job-wide context id: libacx hashes it into its channel index.
struct CommInfo {
    int ctx = 0;
    std::vector<int> ranks;      // comm rank -> world rank
    int me = 0;                  // my rank in this communicator
    bool world_like = false;     // all of COMM_WORLD in world order: collectives take the shared-memory fast path
    std::vector<int> dims, periods;   // Cartesian topology (empty: none)
};
std::unordered_map<int, CommInfo> g_comms;
std::unordered_map<int, std::vector<int>> g_groups;   // group handle -> world ranks
State g;
std::recursive_mutex g_mu;

inline Ring *ring_of(int src, int dst) { return &g.seg.rings[(size_t)src * g.size + dst]; }
inline char *cell_of(int src, int dst, uint64_t idx) {
    size_t r = (size_t)src * g.size + dst;
    return g.seg.cells + (r * g.seg.hdr->ncell + (idx % g.seg.hdr->ncell)) * g.seg.cell_stride;
}

int env_int(const char *a, const char *b = nullptr, const char *c = nullptr, int dflt = -1) {
    const char *names[] = {a, b, c};
    for (const char *n : names) {
        if (!n) continue;
        const char *v = getenv(n);
        if (v && *v) return atoi(v);
    }
    return dflt;
}

size_t seg_size(int n, uint32_t ncell, uint32_t cell_bytes, size_t *off_blobs, size_t *off_scratch,
                size_t *off_rings, size_t *off_cells, size_t *stride) {
    auto up = [](size_t x) { return (x + 4095) & ~(size_t)4095; };
    size_t o = up(sizeof(Header));
    *off_blobs = o;   o = up(o + kBlobBytes * n);
    *off_scratch = o; o = up(o + kScratchBytes);
    *off_rings = o;   o = up(o + sizeof(Ring) * n * n);
    *off_cells = o;
    *stride = (sizeof(CellHdr) + cell_bytes + 63) & ~(size_t)63;
    o = up(o + (size_t)n * n * ncell * *stride);
    return o;
}

void seg_attach() {
    int n = g.size;
    uint32_t ncell = (uint32_t)env_int("ACX_MPI_NCELL", nullptr, nullptr, n > 16 ? 8 : 32);
    uint32_t cell_bytes = (uint32_t)env_int("ACX_MPI_CELL_BYTES", nullptr, nullptr, 16384);
    size_t ob, os, orr, oc, stride;
    size_t bytes = seg_size(n, ncell, cell_bytes, &ob, &os, &orr, &oc, &stride);
    g.seg.name = "/acx_" + g.jobid;
    int fd = -1;
    if (n == 1 && !getenv("ACX_JOBID")) {
        // Singleton: private anonymous mapping, no file.
        void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) die("mmap singleton failed");
        g.seg.hdr = (Header *)p;
        g.seg.name.clear();
    } else if (g.rank == 0) {
        shm_unlink(g.seg.name.c_str());
        fd = shm_open(g.seg.name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0) die("shm_open(create %s): %s", g.seg.name.c_str(), strerror(errno));
        if (ftruncate(fd, (off_t)bytes) != 0) die("ftruncate(%zu): %s", bytes, strerror(errno));
        void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (p == MAP_FAILED) die("mmap: %s", strerror(errno));
        close(fd);
        g.seg.hdr = (Header *)p;
    } else {
        // Wait for rank 0 to create + publish a segment for THIS job (creator alive, size matches).
        double t0 = MPI_Wtime();
        for (;;) {
            fd = shm_open(g.seg.name.c_str(), O_RDWR, 0600);
            if (fd >= 0) {
                struct stat st;
                if (fstat(fd, &st) == 0 && (size_t)st.st_size == bytes) {
                    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
                    if (p != MAP_FAILED) {
                        Header *h = (Header *)p;
                        bool good = false;
                        for (int spin = 0; spin < 2000; spin++) {
                            if (h->magic.load(std::memory_order_acquire) == kMagic && h->nranks == (uint32_t)n &&
                                h->creator_pid > 0 && kill(h->creator_pid, 0) == 0) { good = true; break; }
                            usleep(100);
                        }
                        if (good) { close(fd); g.seg.hdr = h; break; }
                        munmap(p, bytes);
                    }
                }
                close(fd);
            }
            if (MPI_Wtime() - t0 > 120.0) die("rank %d: timed out waiting for shm segment %s", g.rank, g.seg.name.c_str());
            usleep(1000);
        }
    }
    char *base = (char *)g.seg.hdr;
    g.seg.bytes = bytes;
    g.seg.blobs = base + ob;
    g.seg.scratch = base + os;
    g.seg.rings = (Ring *)(base + orr);
    g.seg.cells = base + oc;
    g.seg.cell_stride = stride;
    if (g.rank == 0) {
        Header *h = g.seg.hdr;
        h->nranks = n; h->ncell = ncell; h->cell_bytes = cell_bytes; h->creator_pid = getpid();
        h->abort_code.store(0); h->checkin.store(0); h->bar_count.store(0); h->bar_sense.store(0);
        h->magic.store(kMagic, std::memory_order_release);
    }
    g.seg.hdr->pids[g.rank].store(getpid());
    g.seg.hdr->checkin.fetch_add(1);
    double t0 = MPI_Wtime();
    while (g.seg.hdr->checkin.load() < (uint32_t)n) {
        if (MPI_Wtime() - t0 > 120.0) die("rank %d: only %u/%d ranks checked in", g.rank, g.seg.hdr->checkin.load(), n);
        usleep(200);
    }
}

void check_abort() {
    int c = g.seg.hdr ? g.seg.hdr->abort_code.load(std::memory_order_relaxed) : 0;
    if (c) { fprintf(stderr, "[acx-mpi] rank %d exiting: job aborted with code %d\n", g.rank, c); _exit(c & 0xff ? c & 0xff : 1); }
}

// ---------------------------------------------------------------------------------
// Requests and matching
// ---------------------------------------------------------------------------------
enum ReqKind { RQ_SEND, RQ_RECV, RQ_PSEND, RQ_PRECV };

}  // namespace

struct acx_mpi_request_s {
    ReqKind kind;
    void *buf;
    size_t bytes;       // send: message bytes; recv: capacity
    int peer, tag, comm, part;
    int cid = 1;        // communicator handle (status.MPI_SOURCE is reported in its rank space)
    std::vector<char> packed;   // non-contiguous datatype: the wire image (send: packed at post time; recv: unpacked at completion)
    void *user_buf = nullptr; long long user_count = 0; int user_dt = 0;   // recv: where and how to unpack
    bool done = false, active = true, persistent = false, freed = false;
    MPI_Status st{};
    // send progress
    size_t sent = 0;
    bool hdr_sent = false, is_dev = false, wait_fin = false;
    // recv progress
    size_t got = 0, total = 0;
    // ipc rendezvous on the receive side
    void *ipc_event = nullptr;
    uint64_t ipc_cookie = 0;
    int ipc_src = -1;
    // partitioned
    int partitions = 0;
    size_t part_bytes = 0;
    std::vector<acx_mpi_request_s *> parts;  // child requests (one per partition)
    std::vector<char> part_ready;
};
typedef acx_mpi_request_s Req;

namespace {

struct Unexpected {
    int src, tag, comm, part;
    size_t total, got;
    std::vector<char> data;  // host staging
    bool complete;
    bool is_ipc = false;
    IpcHandle ipc{};
    size_t ipc_off = 0;
    uint64_t cookie = 0;
    Req *attached = nullptr;  // posted recv that adopted this message while streaming
};

std::list<Req *> g_posted;            // posted receives, in post order
std::list<Unexpected *> g_unexp;      // unexpected messages, in arrival order
std::vector<std::deque<Req *>> g_sendq;  // per-destination FIFO of sends being streamed
std::vector<Unexpected *> g_cur_in;   // per-source: message currently being streamed in (unexpected)
std::vector<Req *> g_cur_req;         // per-source: ... or directly into a matched posted recv
std::unordered_map<uint64_t, Req *> g_await_fin;
std::list<Req *> g_ipc_inflight;
uint64_t g_cookie = 1;
std::unordered_map<std::string, void *> g_ipc_cache;

bool match(int rsrc, int rtag, int rcomm, int rpart, int src, int tag, int comm, int part) {
    if (rcomm != comm || rpart != part) return false;
    if (rsrc != MPI_ANY_SOURCE && rsrc != src) return false;
    if (rtag != MPI_ANY_TAG && rtag != tag) return false;
    return true;
}

void copy_in(void *dst, bool dst_dev, const void *src, size_t n) {
    if (!n) return;
    if (dst_dev) {
        is_device_ptr(dst);
        g_cuda.MemcpyAsync(dst, src, n, kMemcpyDefault, cuda_stream());
        g_cuda.StreamSynchronize(cuda_stream());
    } else {
        memcpy(dst, src, n);
    }
}

void *ipc_map(const IpcHandle &h) {
    std::string key(h.b, 64);
    auto it = g_ipc_cache.find(key);
    if (it != g_ipc_cache.end()) return it->second;
    void *p = nullptr;
    int e = g_ipc_open(&p, h, 1 /*cudaIpcMemLazyEnablePeerAccess*/);
    if (e != 0) die("cudaIpcOpenMemHandle failed (%d)", e);
    g_ipc_cache[key] = p;
    return p;
}

void complete_recv(Req *r, int src, int tag, size_t total) {
    r->st.MPI_SOURCE = src;
    r->st.MPI_TAG = tag;
    r->st.MPI_ERROR = total > r->bytes ? MPI_ERR_TRUNCATE : MPI_SUCCESS;
    r->st._bytes = (long long)(total > r->bytes ? r->bytes : total);
    r->done = true;
}

bool ring_push(int dst, const CellHdr &h, const void *payload) {
    Ring *rg = ring_of(g.rank, dst);
    uint64_t head = rg->head.load(std::memory_order_relaxed);
    uint64_t tail = rg->tail.load(std::memory_order_acquire);
    if (head - tail >= g.seg.hdr->ncell) return false;
    char *c = cell_of(g.rank, dst, head);
    memcpy(c, &h, sizeof(h));
    if (h.len && payload) memcpy(c + sizeof(CellHdr), payload, h.len);
    rg->head.store(head + 1, std::memory_order_release);
    return true;
}

void start_ipc_pull(Req *r, int src, const IpcHandle &h, size_t off, size_t total, uint64_t cookie, int tag) {
    is_device_ptr(r->buf);  // binds this thread to the destination's device (progress may run on any thread)
    char *peer = src == g.rank ? (char *)off : (char *)ipc_map(h) + off;
    size_t n = total < r->bytes ? total : r->bytes;
    g_cuda.MemcpyAsync(r->buf, peer, n, kMemcpyDefault, cuda_stream());
    g_cuda.EventCreateWithFlags(&r->ipc_event, 2 /*DisableTiming*/);
    g_cuda.EventRecord(r->ipc_event, cuda_stream());
    r->ipc_cookie = cookie;
    r->ipc_src = src;
    r->total = total;
    r->st.MPI_TAG = tag;
    g_ipc_inflight.push_back(r);
}

// Pull cells from one incoming ring.
void drain_from(int src) {
    Ring *rg = ring_of(src, g.rank);
    for (;;) {
        uint64_t tail = rg->tail.load(std::memory_order_relaxed);
        uint64_t head = rg->head.load(std::memory_order_acquire);
        if (tail == head) return;
        char *c = cell_of(src, g.rank, tail);
        CellHdr h;
        memcpy(&h, c, sizeof(h));
        const char *payload = c + sizeof(CellHdr);
        if (h.kind == CELL_FIN) {
            auto it = g_await_fin.find(h.cookie);
            if (it != g_await_fin.end()) { it->second->done = true; g_await_fin.erase(it); }
        } else if (h.kind == CELL_HDR || h.kind == CELL_RTS_IPC) {
            Req *m = nullptr;
            for (auto it = g_posted.begin(); it != g_posted.end(); ++it) {
                if (match((*it)->peer, (*it)->tag, (*it)->comm, (*it)->part, src, h.tag, h.comm, h.part)) {
                    m = *it; g_posted.erase(it); break;
                }
            }
            if (h.kind == CELL_RTS_IPC) {
                IpcHandle ih; size_t off;
                memcpy(&ih, payload, 64); memcpy(&off, payload + 64, sizeof(size_t));
                if (m) {
                    start_ipc_pull(m, src, ih, off, h.total, h.cookie, h.tag);
                } else {
                    Unexpected *u = new Unexpected{src, h.tag, h.comm, h.part, (size_t)h.total, 0, {}, true};
                    u->is_ipc = true; u->ipc = ih; u->ipc_off = off; u->cookie = h.cookie;
                    g_unexp.push_back(u);
                }
            } else if (m) {
                m->total = h.total; m->got = 0; m->st.MPI_TAG = h.tag;
                if (h.total == 0) complete_recv(m, src, h.tag, 0);
                else g_cur_req[src] = m;
            } else {
                Unexpected *u = new Unexpected{src, h.tag, h.comm, h.part, (size_t)h.total, 0, {}, h.total == 0};
                u->data.resize(h.total);
                g_unexp.push_back(u);
                if (h.total) g_cur_in[src] = u;
            }
        } else if (h.kind == CELL_DATA) {
            if (Req *m = g_cur_req[src]) {
                size_t room = m->got < m->bytes ? m->bytes - m->got : 0;
                size_t n = h.len < room ? h.len : room;
                copy_in((char *)m->buf + m->got, m->is_dev, payload, n);
                m->got += h.len;
                if (m->got >= m->total) { complete_recv(m, src, m->st.MPI_TAG, m->total); g_cur_req[src] = nullptr; }
            } else if (Unexpected *u = g_cur_in[src]) {
                memcpy(u->data.data() + u->got, payload, h.len);
                u->got += h.len;
                if (u->got >= u->total) {
                    u->complete = true; g_cur_in[src] = nullptr;
                    if (u->attached) {
                        Req *m = u->attached;
                        size_t n = u->total < m->bytes ? u->total : m->bytes;
                        copy_in(m->buf, m->is_dev, u->data.data(), n);
                        complete_recv(m, u->src, u->tag, u->total);
                        g_unexp.remove(u); delete u;
                    }
                }
            } else {
                die("DATA cell with no message in flight from %d", src);
            }
        }
        rg->tail.store(tail + 1, std::memory_order_release);
    }
}

// Push as much of the head-of-line send to `dst` as the ring accepts.
void pump_sends(int dst) {
    auto &q = g_sendq[dst];
    while (!q.empty()) {
        Req *s = q.front();
        if (s->is_dev) is_device_ptr(s->buf);  // bind the progressing thread to the source's device
        if (!s->hdr_sent) {
            CellHdr h{};
            h.tag = s->tag; h.comm = s->comm; h.part = s->part; h.total = s->bytes;
            size_t ipc_min = (size_t)env_int("ACX_MPI_IPC_THRESHOLD", nullptr, nullptr, 32768);
            if (s->is_dev && s->bytes >= ipc_min) {
                // CUDA-IPC rendezvous: ship handle+offset; receiver copies and FINs.
                static MemGetAddressRangeFn get_range = nullptr;
                static bool tried = false;
                if (!tried) { tried = true; void *cu = dlopen("libcuda.so.1", RTLD_NOW); if (cu) get_range = (MemGetAddressRangeFn)dlsym(cu, "cuMemGetAddressRange_v2"); }
                unsigned long long base = 0; size_t sz = 0;
                if (!get_range || get_range(&base, &sz, (unsigned long long)s->buf) != 0) die("cuMemGetAddressRange failed");
                IpcHandle ih;
                memset(&ih, 0, sizeof ih);
                size_t off = (size_t)((unsigned long long)s->buf - base);
                if (dst == g.rank) off = (size_t)s->buf;  // self-send: same address space, no IPC
                else if (g_cuda.IpcGetMemHandle(&ih, (void *)base) != 0) die("cudaIpcGetMemHandle failed");
                char pl[64 + sizeof(size_t)];
                memcpy(pl, &ih, 64); memcpy(pl + 64, &off, sizeof(size_t));
                h.kind = CELL_RTS_IPC; h.len = sizeof(pl); h.cookie = g_cookie++;
                if (!ring_push(dst, h, pl)) return;
                g_await_fin[h.cookie] = s;
                s->hdr_sent = true; s->wait_fin = true; s->sent = s->bytes;
                q.pop_front();
                continue;
            }
            h.kind = CELL_HDR; h.len = 0;
            if (!ring_push(dst, h, nullptr)) return;
            s->hdr_sent = true;
        }
        while (s->sent < s->bytes) {
            size_t n = s->bytes - s->sent;
            if (n > g.seg.hdr->cell_bytes) n = g.seg.hdr->cell_bytes;
            CellHdr h{};
            h.kind = CELL_DATA; h.len = (uint32_t)n;
            if (s->is_dev) {
                Ring *rg = ring_of(g.rank, dst);
                uint64_t head = rg->head.load(std::memory_order_relaxed);
                if (head - rg->tail.load(std::memory_order_acquire) >= g.seg.hdr->ncell) return;
                char *c = cell_of(g.rank, dst, head);
                memcpy(c, &h, sizeof(h));
                g_cuda.MemcpyAsync(c + sizeof(CellHdr), (char *)s->buf + s->sent, n, kMemcpyDefault, cuda_stream());
                g_cuda.StreamSynchronize(cuda_stream());
                rg->head.store(head + 1, std::memory_order_release);
            } else if (!ring_push(dst, h, (char *)s->buf + s->sent)) {
                return;
            }
            s->sent += n;
        }
        s->done = true;  // buffered-send semantics: data fully handed to the ring
        q.pop_front();
    }
}

void progress() {
    check_abort();
    for (int p = 0; p < g.size; p++) drain_from(p);
    for (int p = 0; p < g.size; p++) if (!g_sendq[p].empty()) pump_sends(p);
    for (auto it = g_ipc_inflight.begin(); it != g_ipc_inflight.end();) {
        Req *r = *it;
        if (!r->ipc_event || g_cuda.EventQuery(r->ipc_event) == 0) {   // null: the copy finished on an earlier pass, only the FIN is still owed
            if (r->ipc_event) { g_cuda.EventDestroy(r->ipc_event); r->ipc_event = nullptr; }
            CellHdr h{}; h.kind = CELL_FIN; h.cookie = r->ipc_cookie;
            if (!ring_push(r->ipc_src, h, nullptr)) { ++it; continue; }  // ring full: retry the FIN next round (the event is gone, see above)
            complete_recv(r, r->ipc_src, r->st.MPI_TAG, r->total);
            it = g_ipc_inflight.erase(it);
        } else {
            g_cuda.GetLastError();
            ++it;
        }
    }
}

// ---- datatypes -----------------------------------------------------------------------
// Basic handles carry their size in the low byte. Derived handles index g_types; every derived type is kept in the
// canonical form of ACX_Mpi_layout (<= 3 strided dimensions around one contiguous run), which covers contiguous,
// (h)vector and 1-3-D subarray types and their nestings — the layouts of halo faces.
std::vector<ACX_Mpi_layout> g_types;
std::vector<char> g_type_live;
bool is_derived(MPI_Datatype dt) { return dt >= ACX_DT_DERIVED_BASE; }
const ACX_Mpi_layout *layout_of(MPI_Datatype dt) {
    if (!is_derived(dt)) return nullptr;
    size_t i = (size_t)(dt - ACX_DT_DERIVED_BASE);
    return i < g_types.size() && g_type_live[i] ? &g_types[i] : nullptr;
}
bool get_layout(MPI_Datatype dt, ACX_Mpi_layout *L) {
    if (is_derived(dt)) { const ACX_Mpi_layout *p = layout_of(dt); if (!p) return false; *L = *p; return true; }
    const long long sz = dt & 0xff;
    if (!sz) return false;
    for (int i = 0; i < 3; i++) { L->n[i] = 1; L->stride[i] = 0; }
    L->run = sz; L->disp = 0; L->extent = sz; L->size = sz; L->contiguous = 1;
    return true;
}
size_t dt_size(MPI_Datatype dt) {
    if (is_derived(dt)) { const ACX_Mpi_layout *p = layout_of(dt); return p ? (size_t)p->size : 0; }
    return (size_t)(dt & 0xff);
}
bool dt_contig(MPI_Datatype dt) { const ACX_Mpi_layout *p = layout_of(dt); return !is_derived(dt) || (p && p->contiguous && p->disp == 0); }
MPI_Datatype add_type(const ACX_Mpi_layout &L) {
    g_types.push_back(L);
    g_type_live.push_back(1);
    return (MPI_Datatype)(ACX_DT_DERIVED_BASE + g_types.size() - 1);
}
void normalize(ACX_Mpi_layout &L) {
    // drop unit dimensions, then fold the innermost dimension into the run while it continues it seamlessly
    long long n[3], st[3];
    int k = 0;
    for (int d = 0; d < 3; d++) if (L.n[d] > 1) { n[k] = L.n[d]; st[k] = L.stride[d]; k++; }
    int lo = 0;
    while (lo < k && st[lo] == L.run) { L.run *= n[lo]; lo++; }
    for (int d = 0; d < 3; d++) {
        if (lo + d < k) { L.n[d] = n[lo + d]; L.stride[d] = st[lo + d]; }
        else { L.n[d] = 1; L.stride[d] = 0; }
    }
    L.size = L.n[0] * L.n[1] * L.n[2] * L.run;
    L.contiguous = (L.n[0] == 1 && L.n[1] == 1 && L.n[2] == 1 && L.extent == L.size) ? 1 : 0;
}
// `count` blocks of `blocklen` old items, `stride_bytes` apart: adds (up to) two dimensions around `old`.
bool make_strided(const ACX_Mpi_layout &old, long long count, long long blocklen, long long stride_bytes, ACX_Mpi_layout *out) {
    ACX_Mpi_layout L = old;
    int used = 0;
    for (int d = 0; d < 3; d++) if (L.n[d] > 1) used = d + 1;
    auto push = [&](long long n, long long st) {
        if (n == 1) return true;
        if (used == 3) return false;
        L.n[used] = n; L.stride[used] = st; used++;
        return true;
    };
    if (old.contiguous && old.disp == 0) {   // a block of contiguous items is one longer run
        L.run = old.size * blocklen;
        if (!push(count, stride_bytes)) return false;
    } else {
        if (!push(blocklen, old.extent) || !push(count, stride_bytes)) return false;
    }
    L.extent = (count - 1) * stride_bytes + blocklen * old.extent;
    if (count == 0 || blocklen == 0) { L.n[0] = L.n[1] = L.n[2] = 1; L.run = 0; L.extent = 0; }
    normalize(L);
    *out = L;
    return true;
}
template <bool kPack>
void pack_unpack(char *packed, char *user, long long count, const ACX_Mpi_layout &L, size_t limit) {
    size_t done = 0;
    for (long long c = 0; c < count; c++) {
        char *base = user + c * L.extent + L.disp;
        for (long long i2 = 0; i2 < L.n[2]; i2++)
            for (long long i1 = 0; i1 < L.n[1]; i1++)
                for (long long i0 = 0; i0 < L.n[0]; i0++) {
                    if (done >= limit) return;
                    size_t n = (size_t)L.run;
                    if (done + n > limit) n = limit - done;
                    char *u = base + i2 * L.stride[2] + i1 * L.stride[1] + i0 * L.stride[0];
                    if (kPack) memcpy(packed + done, u, n); else memcpy(u, packed + done, n);
                    done += n;
                }
    }
}

CommInfo *comm_of(MPI_Comm c) {
    auto it = g_comms.find(c);
    return it == g_comms.end() ? nullptr : &it->second;
}
// comm rank -> world rank (wildcard / PROC_NULL pass through); -1000 = out of range
int to_world(const CommInfo *ci, int r) {
    if (r == MPI_PROC_NULL || r == MPI_ANY_SOURCE) return r;
    if (r < 0 || r >= (int)ci->ranks.size()) return -1000;
    return ci->ranks[r];
}
int from_world(const CommInfo *ci, int w) {
    if (w < 0 || !ci) return w;
    if (ci->world_like) return w;
    for (size_t i = 0; i < ci->ranks.size(); i++) if (ci->ranks[i] == w) return (int)i;
    return w;
}
void register_comm(int id, int ctx, const std::vector<int> &ranks) {
    CommInfo ci;
    ci.ctx = ctx; ci.ranks = ranks;
    ci.world_like = (int)ranks.size() == g.size;
    for (size_t i = 0; i < ranks.size(); i++) {
        if (ranks[i] == g.rank) ci.me = (int)i;
        if (ranks[i] != (int)i) ci.world_like = false;
    }
    g_comms[id] = ci;
}

Req *post_send(const void *buf, size_t bytes, int dest, int tag, int comm, int part) {
    Req *s = new Req();
    s->kind = RQ_SEND; s->buf = (void *)buf; s->bytes = bytes; s->peer = dest; s->tag = tag; s->comm = comm; s->part = part;
    s->is_dev = is_device_ptr(buf);
    if (dest == MPI_PROC_NULL) { s->done = true; return s; }
    g_sendq[dest].push_back(s);
    pump_sends(dest);
    return s;
}

Req *post_recv(void *buf, size_t bytes, int src, int tag, int comm, int part) {
    Req *r = new Req();
    r->kind = RQ_RECV; r->buf = buf; r->bytes = bytes; r->peer = src; r->tag = tag; r->comm = comm; r->part = part;
    r->is_dev = is_device_ptr(buf);
    if (src == MPI_PROC_NULL) { r->done = true; r->st.MPI_SOURCE = MPI_PROC_NULL; r->st.MPI_TAG = MPI_ANY_TAG; return r; }
    for (auto it = g_unexp.begin(); it != g_unexp.end(); ++it) {
        Unexpected *u = *it;
        if (u->attached || !match(src, tag, comm, part, u->src, u->tag, u->comm, u->part)) continue;
        if (u->is_ipc) {
            start_ipc_pull(r, u->src, u->ipc, u->ipc_off, u->total, u->cookie, u->tag);
            g_unexp.erase(it); delete u;
        } else if (u->complete) {
            size_t n = u->total < bytes ? u->total : bytes;
            copy_in(buf, r->is_dev, u->data.data(), n);
            complete_recv(r, u->src, u->tag, u->total);
            g_unexp.erase(it); delete u;
        } else {
            u->attached = r;  // finishes when the remaining cells arrive
        }
        return r;
    }
    g_posted.push_back(r);
    return r;
}

void wait_req(Req *r) {
    while (!r->done) {
        progress();
        if (!r->done) { g_mu.unlock(); sched_yield(); g_mu.lock(); }
    }
}

void shm_barrier() {
    Header *h = g.seg.hdr;
    g.my_sense ^= 1;
    if (h->bar_count.fetch_add(1) + 1 == (uint32_t)g.size) {
        h->bar_count.store(0);
        h->bar_sense.store(g.my_sense, std::memory_order_release);
    } else {
        while (h->bar_sense.load(std::memory_order_acquire) != g.my_sense) {
            progress();
            g_mu.unlock(); sched_yield(); g_mu.lock();
        }
    }
}

template <typename T> void reduce_typed(T *acc, const T *in, int n, MPI_Op op) {
    for (int i = 0; i < n; i++) {
        switch (op) {
            case MPI_SUM: acc[i] = acc[i] + in[i]; break;
            case MPI_PROD: acc[i] = acc[i] * in[i]; break;
            case MPI_MAX: acc[i] = acc[i] > in[i] ? acc[i] : in[i]; break;
            case MPI_MIN: acc[i] = acc[i] < in[i] ? acc[i] : in[i]; break;
            case MPI_LAND: acc[i] = (T)(acc[i] && in[i]); break;
            case MPI_LOR: acc[i] = (T)(acc[i] || in[i]); break;
            default: break;
        }
    }
}
template <typename T> void reduce_bits(T *acc, const T *in, int n, MPI_Op op) {
    for (int i = 0; i < n; i++) {
        if (op == MPI_BAND) acc[i] &= in[i];
        else if (op == MPI_BOR) acc[i] |= in[i];
    }
}

void reduce_into(void *acc, const void *in, int n, MPI_Datatype dt, MPI_Op op) {
    if (op == MPI_BAND || op == MPI_BOR) {
        switch (dt_size(dt)) {
            case 1: reduce_bits((uint8_t *)acc, (const uint8_t *)in, n, op); return;
            case 2: reduce_bits((uint16_t *)acc, (const uint16_t *)in, n, op); return;
            case 4: reduce_bits((uint32_t *)acc, (const uint32_t *)in, n, op); return;
            default: reduce_bits((uint64_t *)acc, (const uint64_t *)in, n, op); return;
        }
    }
    switch (dt) {
        case MPI_FLOAT: reduce_typed((float *)acc, (const float *)in, n, op); break;
        case MPI_DOUBLE: reduce_typed((double *)acc, (const double *)in, n, op); break;
        case MPI_INT: case MPI_INT32_T: reduce_typed((int32_t *)acc, (const int32_t *)in, n, op); break;
        case MPI_UNSIGNED: case MPI_UINT32_T: reduce_typed((uint32_t *)acc, (const uint32_t *)in, n, op); break;
        case MPI_LONG: case MPI_LONG_LONG: case MPI_INT64_T: case MPI_COUNT: case MPI_AINT:
            reduce_typed((int64_t *)acc, (const int64_t *)in, n, op); break;
        case MPI_UNSIGNED_LONG: case MPI_UNSIGNED_LONG_LONG: case MPI_UINT64_T:
            reduce_typed((uint64_t *)acc, (const uint64_t *)in, n, op); break;
        case MPI_SHORT: case MPI_INT16_T: reduce_typed((int16_t *)acc, (const int16_t *)in, n, op); break;
        case MPI_UNSIGNED_SHORT: case MPI_UINT16_T: reduce_typed((uint16_t *)acc, (const uint16_t *)in, n, op); break;
        case MPI_CHAR: case MPI_SIGNED_CHAR: case MPI_INT8_T: reduce_typed((int8_t *)acc, (const int8_t *)in, n, op); break;
        case MPI_UNSIGNED_CHAR: case MPI_BYTE: case MPI_UINT8_T: case MPI_C_BOOL:
            reduce_typed((uint8_t *)acc, (const uint8_t *)in, n, op); break;
        default: die("MPI reduction on unsupported datatype 0x%x", dt);
    }
}

void atexit_cleanup() {
    if (g.initialized && g.rank == 0 && !g.seg.name.empty()) shm_unlink(g.seg.name.c_str());
}

// Collectives on communicators that are not "all of COMM_WORLD in order" cannot use the job-wide shared-memory barrier;
// they run over point-to-point messages in a reserved context (linear fan-in/fan-out through comm rank 0: these
// communicators are small, and nothing here is on a data path).
constexpr int kCollCtx = 0x20000000;
void coll_send(const CommInfo &ci, int dst, const void *buf, size_t n, int tag) {
    Req *q = post_send(buf, n, ci.ranks[dst], tag, ci.ctx | kCollCtx, -1);
    wait_req(q);
    delete q;
}
void coll_recv(const CommInfo &ci, int src, void *buf, size_t n, int tag) {
    Req *q = post_recv(buf, n, ci.ranks[src], tag, ci.ctx | kCollCtx, -1);
    wait_req(q);
    delete q;
}
void sub_bcast(const CommInfo &ci, void *buf, size_t bytes, int root, int tag) {
    const int n = (int)ci.ranks.size();
    if (ci.me == root) { for (int r = 0; r < n; r++) if (r != root) coll_send(ci, r, buf, bytes, tag); }
    else coll_recv(ci, root, buf, bytes, tag);
}
void sub_barrier(const CommInfo &ci) {
    const int n = (int)ci.ranks.size();
    char b = 0;
    if (ci.me == 0) { for (int r = 1; r < n; r++) coll_recv(ci, r, &b, 1, 1); }
    else coll_send(ci, 0, &b, 1, 1);
    sub_bcast(ci, &b, 1, 0, 2);
}
void sub_allgather(const CommInfo &ci, const void *mine, void *all, size_t bytes) {
    const int n = (int)ci.ranks.size();
    if (ci.me == 0) {
        if (mine != all) memmove(all, mine, bytes);
        for (int r = 1; r < n; r++) coll_recv(ci, r, (char *)all + bytes * r, bytes, 3);
    } else {
        coll_send(ci, 0, mine, bytes, 3);
    }
    sub_bcast(ci, all, bytes * n, 0, 4);
}

#define LOCK std::lock_guard<std::recursive_mutex> _lk(g_mu)

}  // namespace

// =====================================================================================
// Public API
// =====================================================================================
extern "C" {

int MPI_Init_thread(int *, char ***, int required, int *provided) {
    LOCK;
    if (g.initialized) { if (provided) *provided = g.thread_level; return MPI_SUCCESS; }
    g.rank = env_int("ACX_RANK", "RANK", "PMI_RANK", 0);
    g.size = env_int("ACX_SIZE", "WORLD_SIZE", "PMI_SIZE", 1);
    if (g.size > 1024 || g.rank < 0 || g.rank >= g.size) die("bad rank/size %d/%d", g.rank, g.size);
    const char *jid = getenv("ACX_JOBID");
    if (jid && *jid) g.jobid = jid;
    else {
        const char *port = getenv("MASTER_PORT");
        const char *run = getenv("TORCHELASTIC_RUN_ID");
        char tmp[256];
        if (g.size == 1 && !port) snprintf(tmp, sizeof tmp, "solo_%d", (int)getpid());
        else snprintf(tmp, sizeof tmp, "u%d_p%s_%s", (int)getuid(), port ? port : "0", run ? run : "x");
        for (char *p = tmp; *p; p++) if (*p == '/' || *p == ' ') *p = '_';
        g.jobid = tmp;
    }
    seg_attach();
    g_sendq.assign(g.size, {});
    g_cur_in.assign(g.size, nullptr);
    g_cur_req.assign(g.size, nullptr);
    {
        std::vector<int> all(g.size);
        for (int i = 0; i < g.size; i++) all[i] = i;
        register_comm(MPI_COMM_WORLD, 1, all);
        register_comm(MPI_COMM_SELF, 2, std::vector<int>(1, g.rank));
    }
    g.thread_level = MPI_THREAD_MULTIPLE;
    (void)required;
    if (provided) *provided = MPI_THREAD_MULTIPLE;
    g.initialized = true;
    atexit(atexit_cleanup);
    return MPI_SUCCESS;
}

int MPI_Init(int *argc, char ***argv) { int p; return MPI_Init_thread(argc, argv, MPI_THREAD_SINGLE, &p); }

int MPI_Finalize(void) {
    LOCK;
    if (!g.initialized || g.finalized) return MPI_ERR_OTHER;
    // Drain our outgoing traffic, then rendezvous so nobody unmaps a live ring.
    for (;;) {
        bool busy = false;
        for (int p = 0; p < g.size; p++) if (!g_sendq[p].empty()) busy = true;
        if (!busy) break;
        progress();
        g_mu.unlock(); sched_yield(); g_mu.lock();
    }
    shm_barrier();
    g.finalized = true;
    if (g.rank == 0 && !g.seg.name.empty()) shm_unlink(g.seg.name.c_str());
    return MPI_SUCCESS;
}

int MPI_Initialized(int *flag) { *flag = g.initialized; return MPI_SUCCESS; }
int MPI_Finalized(int *flag) { *flag = g.finalized; return MPI_SUCCESS; }
int MPI_Query_thread(int *provided) { *provided = g.thread_level; return MPI_SUCCESS; }

static void (*g_abort_hook)(void) = nullptr;
void ACX_Mpi_set_abort_hook(void (*fn)(void)) { g_abort_hook = fn; }

int MPI_Abort(MPI_Comm, int code) {
    if (g_abort_hook) { void (*h)(void) = g_abort_hook; g_abort_hook = nullptr; h(); }   // e.g. libacx's watchdog diagnostics
    fprintf(stderr, "[acx-mpi] MPI_Abort(%d) on rank %d\n", code, g.rank);
    if (g.seg.hdr) {
        g.seg.hdr->abort_code.store(code ? code : 1);
        for (int r = 0; r < g.size; r++) {
            int pid = g.seg.hdr->pids[r].load();
            if (r != g.rank && pid > 0) kill(pid, SIGTERM);
        }
        if (g.rank == 0 && !g.seg.name.empty()) shm_unlink(g.seg.name.c_str());
    }
    _exit(code ? (code & 0xff ? code & 0xff : 1) : 1);
}

int MPI_Error_string(int code, char *s, int *len) {
    const char *m = "unknown error";
    switch (code) {
        case MPI_SUCCESS: m = "no error"; break;
        case MPI_ERR_BUFFER: m = "invalid buffer"; break;
        case MPI_ERR_COUNT: m = "invalid count"; break;
        case MPI_ERR_TYPE: m = "invalid datatype"; break;
        case MPI_ERR_TAG: m = "invalid tag"; break;
        case MPI_ERR_COMM: m = "invalid communicator"; break;
        case MPI_ERR_RANK: m = "invalid rank"; break;
        case MPI_ERR_REQUEST: m = "invalid request"; break;
        case MPI_ERR_TRUNCATE: m = "message truncated"; break;
        case MPI_ERR_ARG: m = "invalid argument"; break;
        case MPI_ERR_OTHER: m = "other error"; break;
        case MPI_ERR_INTERN: m = "internal error"; break;
    }
    snprintf(s, MPI_MAX_ERROR_STRING, "%s", m);
    if (len) *len = (int)strlen(s);
    return MPI_SUCCESS;
}

int MPI_Comm_set_errhandler(MPI_Comm, MPI_Errhandler) { return MPI_SUCCESS; }

double MPI_Wtime(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Get_processor_name(char *name, int *len) {
    if (gethostname(name, MPI_MAX_PROCESSOR_NAME) != 0) snprintf(name, MPI_MAX_PROCESSOR_NAME, "localhost");
    if (len) *len = (int)strlen(name);
    return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm c, int *rank) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci) return MPI_ERR_COMM;
    *rank = ci->me;
    return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm c, int *size) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci) return MPI_ERR_COMM;
    *size = (int)ci->ranks.size();
    return MPI_SUCCESS;
}

// Handle for a communicator being created collectively over `parent`: the maximum of the members' private counters,
// so every member (and therefore both ends of any message on the new communicator) uses the same number.
static int agree_comm_id(MPI_Comm parent) {
    int id = g.next_comm;
    MPI_Allreduce(MPI_IN_PLACE, &id, 1, MPI_INT, MPI_MAX, parent);
    g.next_comm = id + 1;
    return id;
}

int MPI_Comm_dup(MPI_Comm c, MPI_Comm *out) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci) return MPI_ERR_COMM;
    const int id = agree_comm_id(c);
    CommInfo copy = *comm_of(c);
    copy.ctx = id;
    g_comms[id] = copy;
    *out = id;
    return MPI_SUCCESS;
}

int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm *out) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci) return MPI_ERR_COMM;
    const int n = (int)ci->ranks.size();
    int mine[2] = {color, key};
    std::vector<int> all(2 * (size_t)n);
    MPI_Allgather(mine, 2, MPI_INT, all.data(), 2, MPI_INT, c);
    const int id = agree_comm_id(c);
    if (color == MPI_UNDEFINED) { *out = MPI_COMM_NULL; return MPI_SUCCESS; }
    ci = comm_of(c);
    std::vector<std::pair<std::pair<int, int>, int>> members;   // ((key, parent rank), world rank)
    for (int r = 0; r < n; r++) if (all[2 * r] == color) members.push_back({{all[2 * r + 1], r}, ci->ranks[r]});
    std::sort(members.begin(), members.end());
    std::vector<int> ranks;
    for (auto &m : members) ranks.push_back(m.second);
    register_comm(id, id, ranks);
    *out = id;
    return MPI_SUCCESS;
}

int MPI_Comm_free(MPI_Comm *c) {
    LOCK;
    if (*c != MPI_COMM_WORLD && *c != MPI_COMM_SELF) g_comms.erase(*c);
    *c = MPI_COMM_NULL;
    return MPI_SUCCESS;
}

int MPI_Comm_group(MPI_Comm c, MPI_Group *group) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci) return MPI_ERR_COMM;
    int id = g.next_group++;
    g_groups[id] = ci->ranks;
    *group = id;
    return MPI_SUCCESS;
}
int MPI_Group_size(MPI_Group group, int *size) {
    LOCK;
    if (group == MPI_GROUP_EMPTY) { *size = 0; return MPI_SUCCESS; }
    auto it = g_groups.find(group);
    if (it == g_groups.end()) return MPI_ERR_ARG;
    *size = (int)it->second.size();
    return MPI_SUCCESS;
}
int MPI_Group_rank(MPI_Group group, int *rank) {
    LOCK;
    *rank = MPI_UNDEFINED;
    auto it = g_groups.find(group);
    if (it == g_groups.end()) return group == MPI_GROUP_EMPTY ? MPI_SUCCESS : MPI_ERR_ARG;
    for (size_t i = 0; i < it->second.size(); i++) if (it->second[i] == g.rank) *rank = (int)i;
    return MPI_SUCCESS;
}
int MPI_Group_translate_ranks(MPI_Group g1, int n, const int r1[], MPI_Group g2, int r2[]) {
    LOCK;
    auto a = g_groups.find(g1), b = g_groups.find(g2);
    if (a == g_groups.end() || b == g_groups.end()) return MPI_ERR_ARG;
    for (int i = 0; i < n; i++) {
        r2[i] = MPI_UNDEFINED;
        if (r1[i] == MPI_PROC_NULL) { r2[i] = MPI_PROC_NULL; continue; }
        if (r1[i] < 0 || r1[i] >= (int)a->second.size()) return MPI_ERR_RANK;
        const int w = a->second[r1[i]];
        for (size_t j = 0; j < b->second.size(); j++) if (b->second[j] == w) { r2[i] = (int)j; break; }
    }
    return MPI_SUCCESS;
}
int MPI_Group_free(MPI_Group *group) {
    LOCK;
    g_groups.erase(*group);
    *group = MPI_GROUP_NULL;
    return MPI_SUCCESS;
}

// ---- Cartesian topologies ------------------------------------------------------------
int MPI_Dims_create(int nnodes, int ndims, int dims[]) {
    if (nnodes <= 0 || ndims <= 0) return MPI_ERR_ARG;
    int fixed = 1, free_dims = 0;
    for (int i = 0; i < ndims; i++) { if (dims[i] < 0) return MPI_ERR_ARG; if (dims[i] > 0) fixed *= dims[i]; else free_dims++; }
    if (nnodes % fixed) return MPI_ERR_ARG;
    int rest = nnodes / fixed;
    std::vector<int> f(free_dims > 0 ? free_dims : 0, 1);
    // distribute prime factors, largest first, always onto the currently smallest dimension: as square as possible
    std::vector<int> primes;
    for (int p = 2; p * p <= rest; p++) while (rest % p == 0) { primes.push_back(p); rest /= p; }
    if (rest > 1) primes.push_back(rest);
    if (free_dims == 0) return primes.empty() ? MPI_SUCCESS : MPI_ERR_ARG;
    std::sort(primes.rbegin(), primes.rend());
    for (int p : primes) *std::min_element(f.begin(), f.end()) *= p;
    std::sort(f.rbegin(), f.rend());   // non-increasing order, as the standard asks
    int k = 0;
    for (int i = 0; i < ndims; i++) if (dims[i] == 0) dims[i] = f[k++];
    return MPI_SUCCESS;
}
int MPI_Cart_create(MPI_Comm c, int ndims, const int dims[], const int periods[], int, MPI_Comm *out) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci || ndims <= 0) return MPI_ERR_COMM;
    int total = 1;
    for (int i = 0; i < ndims; i++) total *= dims[i];
    if (total > (int)ci->ranks.size()) return MPI_ERR_ARG;
    // ranks beyond the grid get MPI_COMM_NULL (split keeps parent order, i.e. row-major coordinates)
    MPI_Comm nc;
    int e = MPI_Comm_split(c, ci->me < total ? 0 : MPI_UNDEFINED, ci->me, &nc);
    if (e) return e;
    if (nc != MPI_COMM_NULL) {
        CommInfo *ni = comm_of(nc);
        ni->dims.assign(dims, dims + ndims);
        ni->periods.assign(periods, periods + ndims);
    }
    *out = nc;
    return MPI_SUCCESS;
}
int MPI_Cart_coords(MPI_Comm c, int rank, int maxdims, int coords[]) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci || ci->dims.empty()) return MPI_ERR_COMM;
    if (rank < 0 || rank >= (int)ci->ranks.size()) return MPI_ERR_RANK;
    const int nd = (int)ci->dims.size();
    for (int i = nd - 1; i >= 0; i--) { if (i < maxdims) coords[i] = rank % ci->dims[i]; rank /= ci->dims[i]; }
    return MPI_SUCCESS;
}
int MPI_Cart_rank(MPI_Comm c, const int coords[], int *rank) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci || ci->dims.empty()) return MPI_ERR_COMM;
    int r = 0;
    for (size_t i = 0; i < ci->dims.size(); i++) {
        int x = coords[i];
        if (ci->periods[i]) x = ((x % ci->dims[i]) + ci->dims[i]) % ci->dims[i];
        else if (x < 0 || x >= ci->dims[i]) return MPI_ERR_RANK;
        r = r * ci->dims[i] + x;
    }
    *rank = r;
    return MPI_SUCCESS;
}
int MPI_Cart_shift(MPI_Comm c, int direction, int disp, int *src, int *dst) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci || ci->dims.empty() || direction < 0 || direction >= (int)ci->dims.size()) return MPI_ERR_COMM;
    std::vector<int> me(ci->dims.size());
    MPI_Cart_coords(c, ci->me, (int)me.size(), me.data());
    auto at = [&](int delta) {
        std::vector<int> x = me;
        x[direction] += delta;
        int r;
        return MPI_Cart_rank(c, x.data(), &r) == MPI_SUCCESS ? r : MPI_PROC_NULL;
    };
    *src = at(-disp);
    *dst = at(disp);
    return MPI_SUCCESS;
}

int MPI_Type_size(MPI_Datatype dt, int *size) { LOCK; *size = (int)dt_size(dt); return (*size || layout_of(dt)) ? MPI_SUCCESS : MPI_ERR_TYPE; }
int MPI_Get_count(const MPI_Status *st, MPI_Datatype dt, int *count) {
    size_t s = dt_size(dt);
    if (!s) return MPI_ERR_TYPE;
    *count = (st->_bytes % (long long)s) ? MPI_UNDEFINED : (int)(st->_bytes / (long long)s);
    return MPI_SUCCESS;
}

int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype *nt) {
    LOCK;
    ACX_Mpi_layout o, L;
    if (count < 0 || !get_layout(old, &o)) return MPI_ERR_TYPE;
    if (!make_strided(o, 1, count, 0, &L)) return MPI_ERR_TYPE;
    *nt = add_type(L);
    return MPI_SUCCESS;
}
int MPI_Type_create_hvector(int count, int blocklen, MPI_Aint stride_bytes, MPI_Datatype old, MPI_Datatype *nt) {
    LOCK;
    ACX_Mpi_layout o, L;
    if (count < 0 || blocklen < 0 || !get_layout(old, &o)) return MPI_ERR_TYPE;
    if (!make_strided(o, count, blocklen, stride_bytes, &L)) return MPI_ERR_TYPE;   // more than 3 strided dimensions
    *nt = add_type(L);
    return MPI_SUCCESS;
}
int MPI_Type_vector(int count, int blocklen, int stride, MPI_Datatype old, MPI_Datatype *nt) {
    ACX_Mpi_layout o;
    { LOCK; if (!get_layout(old, &o)) return MPI_ERR_TYPE; }
    return MPI_Type_create_hvector(count, blocklen, (MPI_Aint)stride * o.extent, old, nt);
}
int MPI_Type_create_subarray(int ndims, const int sizes[], const int subsizes[], const int starts[], int order, MPI_Datatype old,
                             MPI_Datatype *nt) {
    LOCK;
    ACX_Mpi_layout o;
    if (ndims < 1 || ndims > 4 || !get_layout(old, &o) || !(o.contiguous && o.disp == 0)) return MPI_ERR_TYPE;
    // walk the dimensions from fastest to slowest varying
    ACX_Mpi_layout L = o;
    long long pitch = o.extent, disp = 0, total = 1;
    int used = 0;
    for (int k = 0; k < ndims; k++) {
        const int d = order == MPI_ORDER_FORTRAN ? k : ndims - 1 - k;
        if (subsizes[d] < 1 || starts[d] < 0 || starts[d] + subsizes[d] > sizes[d]) return MPI_ERR_ARG;
        disp += (long long)starts[d] * pitch;
        if (k == 0) L.run = o.size * subsizes[d];
        else if (subsizes[d] > 1) {
            if (used == 3) return MPI_ERR_TYPE;
            L.n[used] = subsizes[d]; L.stride[used] = pitch; used++;
        }
        pitch *= sizes[d];
        total *= sizes[d];
    }
    L.disp = disp;
    L.extent = total * o.extent;
    normalize(L);
    if (L.disp != 0) L.contiguous = 0;
    *nt = add_type(L);
    return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype *dt) { LOCK; return (!is_derived(*dt) || layout_of(*dt)) ? MPI_SUCCESS : MPI_ERR_TYPE; }
int MPI_Type_free(MPI_Datatype *dt) {
    LOCK;
    if (is_derived(*dt) && layout_of(*dt)) g_type_live[(size_t)(*dt - ACX_DT_DERIVED_BASE)] = 0;
    *dt = MPI_DATATYPE_NULL;
    return MPI_SUCCESS;
}
int MPI_Type_get_extent(MPI_Datatype dt, MPI_Aint *lb, MPI_Aint *extent) {
    LOCK;
    ACX_Mpi_layout L;
    if (!get_layout(dt, &L)) return MPI_ERR_TYPE;
    if (lb) *lb = 0;
    if (extent) *extent = L.extent;
    return MPI_SUCCESS;
}
int MPI_Pack_size(int incount, MPI_Datatype dt, MPI_Comm, int *size) {
    LOCK;
    size_t s = dt_size(dt);
    if (!s && !is_derived(dt)) return MPI_ERR_TYPE;
    *size = (int)(s * (size_t)incount);
    return MPI_SUCCESS;
}
int ACX_Mpi_type_layout(MPI_Datatype dt, ACX_Mpi_layout *L) { LOCK; return get_layout(dt, L) ? MPI_SUCCESS : MPI_ERR_TYPE; }

int MPI_Barrier(MPI_Comm c) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci) return MPI_ERR_COMM;
    if (ci->ranks.size() == 1) return MPI_SUCCESS;
    if (ci->world_like) shm_barrier();
    else { CommInfo snap = *ci; sub_barrier(snap); }
    return MPI_SUCCESS;
}

int ACX_Mpi_exchange(const void *mine, void *all, size_t bytes) {
    LOCK;
    if (bytes > kBlobBytes) return MPI_ERR_ARG;
    shm_barrier();  // previous readers are done with the blobs
    memcpy(g.seg.blobs + (size_t)g.rank * kBlobBytes, mine, bytes);
    shm_barrier();
    for (int r = 0; r < g.size; r++) memcpy((char *)all + (size_t)r * bytes, g.seg.blobs + (size_t)r * kBlobBytes, bytes);
    return MPI_SUCCESS;
}

void *ACX_Mpi_shared_scratch(size_t *bytes) { if (bytes) *bytes = kScratchBytes; return g.seg.scratch; }
int ACX_Mpi_local_rank(void) { return env_int("ACX_LOCAL_RANK", "LOCAL_RANK", nullptr, g.rank); }
const char *ACX_Mpi_jobid(void) { return g.jobid.c_str(); }

int MPI_Allgather(const void *sb, int sc, MPI_Datatype sdt, void *rb, int rc, MPI_Datatype rdt, MPI_Comm c) {
    LOCK;
    size_t sbytes = (size_t)sc * dt_size(sdt), rbytes = (size_t)rc * dt_size(rdt);
    CommInfo *ci = comm_of(c);
    if (!ci) return MPI_ERR_COMM;
    if (ci->ranks.size() == 1) { if (sb != MPI_IN_PLACE) memcpy(rb, sb, sbytes); return MPI_SUCCESS; }
    const char *src = sb == MPI_IN_PLACE ? (char *)rb + rbytes * ci->me : (const char *)sb;
    std::vector<char> mine(src, src + rbytes);
    if (!ci->world_like) {
        CommInfo snap = *ci;
        if (snap.me == 0) memcpy(rb, mine.data(), rbytes);
        sub_allgather(snap, snap.me == 0 ? rb : (void *)mine.data(), rb, rbytes);
        return MPI_SUCCESS;
    }
    for (size_t off = 0; off < rbytes; off += kBlobBytes) {
        size_t n = rbytes - off < kBlobBytes ? rbytes - off : kBlobBytes;
        std::vector<char> tmp(n * g.size);
        ACX_Mpi_exchange(mine.data() + off, tmp.data(), n);
        for (int r = 0; r < g.size; r++) memcpy((char *)rb + rbytes * r + off, tmp.data() + n * r, n);
    }
    return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sb, void *rb, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm c) {
    LOCK;
    size_t es = dt_size(dt), bytes = es * (size_t)count;
    if (sb != MPI_IN_PLACE) memcpy(rb, sb, bytes);
    CommInfo *ci = comm_of(c);
    if (!ci) return MPI_ERR_COMM;
    if (ci->ranks.size() == 1) return MPI_SUCCESS;
    if (!ci->world_like) {
        CommInfo snap = *ci;
        const int n = (int)snap.ranks.size();
        if (snap.me == 0) {
            std::vector<char> in(bytes);
            for (int r = 1; r < n; r++) { coll_recv(snap, r, in.data(), bytes, 5); reduce_into(rb, in.data(), count, dt, op); }
        } else {
            coll_send(snap, 0, rb, bytes, 5);
        }
        sub_bcast(snap, rb, bytes, 0, 6);
        return MPI_SUCCESS;
    }
    int per = (int)(kBlobBytes / es);
    std::vector<char> tmp(kBlobBytes * g.size);
    for (int off = 0; off < count; off += per) {
        int n = count - off < per ? count - off : per;
        ACX_Mpi_exchange((char *)rb + es * off, tmp.data(), es * n);
        memcpy((char *)rb + es * off, tmp.data(), es * n);
        for (int r = 1; r < g.size; r++) reduce_into((char *)rb + es * off, tmp.data() + es * n * r, n, dt, op);
    }
    return MPI_SUCCESS;
}

int MPI_Reduce(const void *sb, void *rb, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm c) {
    std::vector<char> tmp(dt_size(dt) * (size_t)count);
    const void *in = sb == MPI_IN_PLACE ? rb : sb;
    memcpy(tmp.data(), in, tmp.size());
    int e = MPI_Allreduce(MPI_IN_PLACE, tmp.data(), count, dt, op, c);
    int rank; MPI_Comm_rank(c, &rank);
    if (rank == root) memcpy(rb, tmp.data(), tmp.size());
    return e;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm c) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci) return MPI_ERR_COMM;
    if (ci->ranks.size() == 1) return MPI_SUCCESS;
    size_t bytes = dt_size(dt) * (size_t)count;
    if (root < 0 || root >= (int)ci->ranks.size()) return MPI_ERR_RANK;
    if (!ci->world_like) { CommInfo snap = *ci; sub_bcast(snap, buf, bytes, root, 7); return MPI_SUCCESS; }
    std::vector<char> tmp(kBlobBytes * g.size);
    for (size_t off = 0; off < bytes; off += kBlobBytes) {
        size_t n = bytes - off < kBlobBytes ? bytes - off : kBlobBytes;
        ACX_Mpi_exchange((char *)buf + off, tmp.data(), n);
        memcpy((char *)buf + off, tmp.data() + n * root, n);
    }
    return MPI_SUCCESS;
}

int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm c, MPI_Request *req) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci) return MPI_ERR_COMM;
    if (dest == MPI_ANY_SOURCE || (dest = to_world(ci, dest)) == -1000) return MPI_ERR_RANK;
    if (!dt_contig(dt)) {
        // non-contiguous layout: pack now (sends have buffered semantics anyway), ship the packed image
        ACX_Mpi_layout L;
        if (!get_layout(dt, &L) || is_device_ptr(buf)) return MPI_ERR_TYPE;
        std::vector<char> img((size_t)L.size * (size_t)count);
        pack_unpack<true>(img.data(), (char *)buf, count, L, img.size());
        const char *data = img.data();
        *req = post_send(data, img.size(), dest, tag, ci->ctx, -1);
        (*req)->packed = std::move(img);   // the buffer (and `data`) stays where it is
        (*req)->cid = c;
        return MPI_SUCCESS;
    }
    *req = post_send(buf, dt_size(dt) * (size_t)count, dest, tag, ci->ctx, -1);
    (*req)->cid = c;
    return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Request *req) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci) return MPI_ERR_COMM;
    if ((src = to_world(ci, src)) == -1000) return MPI_ERR_RANK;
    if (!dt_contig(dt)) {
        ACX_Mpi_layout L;
        if (!get_layout(dt, &L) || is_device_ptr(buf)) return MPI_ERR_TYPE;
        std::vector<char> img((size_t)L.size * (size_t)count);
        char *data = img.data();
        if (src == MPI_PROC_NULL) { *req = post_recv(buf, 0, src, tag, ci->ctx, -1); (*req)->cid = c; return MPI_SUCCESS; }
        *req = post_recv(data, img.size(), src, tag, ci->ctx, -1);
        (*req)->packed = std::move(img);
        (*req)->user_buf = buf; (*req)->user_count = count; (*req)->user_dt = dt;
        (*req)->cid = c;
        progress();
        return MPI_SUCCESS;
    }
    *req = post_recv(buf, dt_size(dt) * (size_t)count, src, tag, ci->ctx, -1);
    (*req)->cid = c;
    progress();
    return MPI_SUCCESS;
}

static void finish_req(MPI_Request *req, MPI_Status *st) {
    Req *r = *req;
    if (r->user_buf && r->kind == RQ_RECV) {   // non-contiguous receive: scatter the packed image into the user's layout
        ACX_Mpi_layout L;
        if (get_layout(r->user_dt, &L)) pack_unpack<false>(r->packed.data(), (char *)r->user_buf, r->user_count, L, (size_t)r->st._bytes);
        r->user_buf = nullptr;
    }
    if (st) {
        *st = r->st;
        if (st->MPI_SOURCE >= 0) st->MPI_SOURCE = from_world(comm_of(r->cid), st->MPI_SOURCE);   // report in the communicator's rank space
    }
    if (r->persistent) { r->active = false; return; }
    delete r;
    *req = MPI_REQUEST_NULL;
}

static bool test_partitioned(Req *r) {
    bool all = true;
    for (Req *p : r->parts) if (p && !p->done) all = false;
    if (r->kind == RQ_PSEND) for (int i = 0; i < r->partitions; i++) if (!r->part_ready[i]) all = false;
    return all;
}

int MPI_Test(MPI_Request *req, int *flag, MPI_Status *st) {
    LOCK;
    if (!*req) { *flag = 1; return MPI_SUCCESS; }
    progress();
    Req *r = *req;
    if (r->persistent) {
        if (!r->active) { *flag = 1; return MPI_SUCCESS; }
        *flag = test_partitioned(r);
        if (*flag) {
            for (Req *&p : r->parts) { delete p; p = nullptr; }
            r->st.MPI_SOURCE = r->peer; r->st.MPI_TAG = r->tag; r->st.MPI_ERROR = MPI_SUCCESS;
            finish_req(req, st);
        }
        return MPI_SUCCESS;
    }
    *flag = r->done;
    if (r->done) finish_req(req, st);
    return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *req, MPI_Status *st) {
    LOCK;
    if (!*req) return MPI_SUCCESS;
    for (;;) {
        int flag = 0;
        MPI_Test(req, &flag, st);
        if (flag) return MPI_SUCCESS;
        g_mu.unlock(); sched_yield(); g_mu.lock();
    }
}

int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts) {
    for (int i = 0; i < n; i++) MPI_Wait(&reqs[i], sts ? &sts[i] : MPI_STATUS_IGNORE);
    return MPI_SUCCESS;
}

int MPI_Testall(int n, MPI_Request *reqs, int *flag, MPI_Status *sts) {
    LOCK;
    progress();
    for (int i = 0; i < n; i++) {
        Req *r = reqs[i];
        if (!r) continue;
        bool d = r->persistent ? (!r->active || test_partitioned(r)) : r->done;
        if (!d) { *flag = 0; return MPI_SUCCESS; }
    }
    *flag = 1;
    for (int i = 0; i < n; i++) { int f; MPI_Test(&reqs[i], &f, sts ? &sts[i] : MPI_STATUS_IGNORE); }
    return MPI_SUCCESS;
}

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm c) {
    MPI_Request r;
    int e = MPI_Isend(buf, count, dt, dest, tag, c, &r);
    return e ? e : MPI_Wait(&r, MPI_STATUS_IGNORE);
}
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Status *st) {
    MPI_Request r;
    int e = MPI_Irecv(buf, count, dt, src, tag, c, &r);
    return e ? e : MPI_Wait(&r, st);
}
int MPI_Sendrecv(const void *sb, int sc, MPI_Datatype sdt, int dest, int stag, void *rb, int rc, MPI_Datatype rdt,
                 int src, int rtag, MPI_Comm c, MPI_Status *st) {
    MPI_Request r[2];
    int e = MPI_Irecv(rb, rc, rdt, src, rtag, c, &r[0]);
    if (e) return e;
    e = MPI_Isend(sb, sc, sdt, dest, stag, c, &r[1]);
    if (e) return e;
    MPI_Wait(&r[1], MPI_STATUS_IGNORE);
    return MPI_Wait(&r[0], st);
}

int MPI_Request_free(MPI_Request *req) {
    LOCK;
    Req *r = *req;
    if (!r) return MPI_SUCCESS;
    if (r->persistent) {
        for (Req *p : r->parts) if (p) wait_req(p), delete p;
        delete r;
    } else {
        // Non-persistent: let it finish in the background (sends only); simple approach: wait.
        wait_req(r);
        delete r;
    }
    *req = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}

// ---- partitioned -------------------------------------------------------------------
static int pinit(ReqKind k, void *buf, int partitions, MPI_Count count, MPI_Datatype dt, int peer, int tag, MPI_Comm c,
                 MPI_Request *req) {
    LOCK;
    CommInfo *ci = comm_of(c);
    if (!ci) return MPI_ERR_COMM;
    const int ctx = ci->ctx;
    if (partitions <= 0 || count < 0) return MPI_ERR_COUNT;
    if (!dt_contig(dt)) return MPI_ERR_TYPE;   // partitioned transfers: contiguous datatypes only
    if (peer == MPI_ANY_SOURCE || (peer = to_world(ci, peer)) == -1000) return MPI_ERR_RANK;
    Req *r = new Req();
    r->cid = c;
    r->kind = k; r->buf = buf; r->partitions = partitions; r->part_bytes = dt_size(dt) * (size_t)count;
    r->bytes = r->part_bytes * partitions; r->peer = peer; r->tag = tag; r->comm = ctx | 0x40000000; r->part = -1;
    r->persistent = true; r->active = false; r->done = false;
    r->parts.assign(partitions, nullptr);
    r->part_ready.assign(partitions, 0);
    *req = r;
    return MPI_SUCCESS;
}

int MPI_Psend_init(const void *buf, int partitions, MPI_Count count, MPI_Datatype dt, int dest, int tag, MPI_Comm c,
                   MPI_Info, MPI_Request *req) {
    return pinit(RQ_PSEND, (void *)buf, partitions, count, dt, dest, tag, c, req);
}
int MPI_Precv_init(void *buf, int partitions, MPI_Count count, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Info,
                   MPI_Request *req) {
    return pinit(RQ_PRECV, buf, partitions, count, dt, src, tag, c, req);
}

int MPI_Start(MPI_Request *req) {
    LOCK;
    Req *r = *req;
    if (!r || !r->persistent || r->active) return MPI_ERR_REQUEST;
    r->active = true;
    std::fill(r->part_ready.begin(), r->part_ready.end(), 0);
    if (r->kind == RQ_PRECV)
        for (int p = 0; p < r->partitions; p++)
            r->parts[p] = post_recv((char *)r->buf + r->part_bytes * p, r->part_bytes, r->peer, r->tag, r->comm, p);
    return MPI_SUCCESS;
}
int MPI_Startall(int n, MPI_Request *reqs) {
    for (int i = 0; i < n; i++) { int e = MPI_Start(&reqs[i]); if (e) return e; }
    return MPI_SUCCESS;
}

int MPI_Pready(int partition, MPI_Request r) {
    LOCK;
    if (!r || r->kind != RQ_PSEND || !r->active || partition < 0 || partition >= r->partitions) return MPI_ERR_REQUEST;
    if (r->part_ready[partition]) return MPI_ERR_ARG;
    r->part_ready[partition] = 1;
    r->parts[partition] = post_send((char *)r->buf + r->part_bytes * partition, r->part_bytes, r->peer, r->tag, r->comm, partition);
    return MPI_SUCCESS;
}

int MPI_Parrived(MPI_Request r, int partition, int *flag) {
    LOCK;
    if (!r || r->kind != RQ_PRECV || partition < 0 || partition >= r->partitions) return MPI_ERR_REQUEST;
    if (!r->active) { *flag = 1; return MPI_SUCCESS; }  // inactive request: completed round
    progress();
    *flag = r->parts[partition] && r->parts[partition]->done;
    return MPI_SUCCESS;
}

}  // extern "C"
