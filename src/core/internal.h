// acx-b200 internal definitions: control-window layout shared by host and device,
// global state, request objects, logging/error macros.
//
// Design (vs the reference's flags-in-host-memory + CPU proxy):
//  * every rank owns two device allocations, both CUDA-IPC mapped into every peer at
//    MPIX_Init: a CONTROL window (rendezvous rings, acks, op records, eager slots,
//    partitioned flags) and a symmetric HEAP (user-visible via MPIX_Alloc_mem, also
//    used for bounce buffers);
//  * a message S->R on channel c (= hash(tag,comm)) takes ticket k from a device-side
//    counter when its trigger kernel EXECUTES (stream/graph order == matching order,
//    relaunch-safe); slot e = k % depth, generation g = k / depth of the ring that
//    lives in R's control window;
//  * eager (<= eager_max): the send kernel stores payload+header into R's slot and
//    bumps the slot's arrival count (red.release.sys); the receiver's wait kernel copies
//    out locally;
//  * rendezvous: both triggers add to the slot's 64-bit state word (hi32 = sends arrived,
//    lo32 = recvs posted). Whoever arrives SECOND moves the data: a sender pushes
//    (st over NVLink), a receiver pulls (ld over NVLink), then publishes done (to R)
//    and sdone (to S). Nobody spins on remote memory, nothing waits for the peer's
//    *stream* to make progress (MPI progress semantics without a progress thread).
#ifndef ACX_INTERNAL_H
#define ACX_INTERNAL_H

#include <mpi.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/acx_device.cuh"

// ---- shared host/device layout -----------------------------------------------------
enum { ACX_KIND_EAGER = 1, ACX_KIND_RNDV = 2, ACX_KIND_RNDV_PUSH = 3, ACX_KIND_RNDV_COOP = 4, ACX_KIND_RNDV_LATE = 5 };   // RNDV_LATE: large messages — the SENDER moves the data, at its trigger if the receive is posted by then, else from its wait kernel (see late_helper)   // RNDV: second arriver moves; RNDV_PUSH: the sender always pushes; RNDV_COOP: second arriver moves, both ends may help (chunk cursor)

// Buffer locations travel as 64-bit "offsets": bit 63 clear = offset into the owner's symmetric heap;
// bit 63 set = registered user allocation: bits [62:48] registration index, bits [47:0] offset inside it
// (translated through the reader's xlat table, filled when it opened the owner's CUDA-IPC handle).
#define ACX_OFF_REG_BIT (1ull << 63)
#define ACX_OFF_REG_IDX(o) ((uint32_t)(((o) >> 48) & 0x7fff))
#define ACX_OFF_REG_OFF(o) ((o) & 0xffffffffffffull)

struct alignas(64) acx_entry {  // one slot of a (src -> me, channel) ring; lives at the receiver
    uint64_t state;      // hi32: #sends arrived on this slot, lo32: #recvs posted (atomic add, never reset)
    uint64_t done;       // #generations whose data has fully landed (mover: release store)
    uint64_t consumed;   // #generations the receiver has finished with (receiver-local bookkeeping)
    uint64_t s_src_off;  // sender: heap offset of the source (rendezvous)
    uint64_t s_bytes;    // sender: message bytes
    uint64_t coop_cursor; // cooperative moving (ACX_COOP): next unclaimed chunk of this generation's payload (atomic add by agents on either GPU)
    uint64_t coop_done;   // cooperative moving: chunks fully copied; both are reset by the receiver's wait kernel before the slot is recycled
    uint64_t claim;      // RNDV_LATE: who moves this generation's payload, ((generation + 1) << 2) | 1 sender / 2 receiver (fallback pull); CAS from either GPU
    uint32_t s_tag;
    uint32_t s_kind;     // low 8 bits: ACX_KIND_*; upper 24 bits: the sender's communicator context (checked by the receive: see ACX_ENT_CTX)
};
#define ACX_ENT_KIND(x) ((uint32_t)(x) & 0xffu)
#define ACX_ENT_CTX(x) ((uint32_t)(x) >> 8)
#define ACX_ENT_HEADER(tag, kind, ctx) ((uint64_t)(uint32_t)(tag) | ((uint64_t)((uint32_t)(kind) | ((uint32_t)(ctx) << 8)) << 32))

struct alignas(32) acx_sack {  // per slot, lives at the SENDER, written by the receiver / mover
    uint64_t sdone;      // #generations whose send completed (source buffer reusable)
    uint64_t efree;      // #generations the receiver has consumed (slot reusable by generation efree)
    uint64_t r_dst_off;  // mirror of the receiver's posted destination (so a late sender reads it locally)
    uint64_t r_cap;
    uint64_t r_gen;      // #generations for which the mirror above has been published (release-stored last)
    uint64_t pad;
};

struct alignas(64) acx_op {  // per request slot, local
    uint64_t ticket;     // k assigned by the trigger kernel
    uint64_t armed;      // request serial written by the trigger after `ticket` (host waits key on it)
    uint64_t peer_ptr;   // mover: translated remote address
    uint64_t n;          // mover: bytes to move
    uint32_t mover;      // 1 if this side arrived second and must move the data
    uint32_t eager;      // send side: message went eager (nothing to wait for)
    uint32_t cta_count;  // last-CTA detection
    uint32_t pad;        // fused rendezvous: CTAs that have left the kernel (last one re-arms the gate)
    uint64_t recv_bytes; // receive side: actual message size
    uint64_t bounce_n;   // receive side: bytes the bounce-out kernel still has to copy (0 if eager/zero-copy)
    uint64_t fused_decided;  // single-launch rendezvous gate: 1 once CTA 0 has published the match of this launch
    uint64_t coop;       // this message is moved cooperatively (kind RNDV_COOP): the mover claims chunks from the slot's cursor
    uint64_t late_state; // RNDV_LATE at wait time: ((ticket + 1) << 2) | 1 nothing to move here / 2 move now (n, peer_ptr valid); written by the waiter CTA, polled by the helper CTAs
    uint64_t helpers_out; // helper CTAs of the current wait kernel that are done with this record; the waiter CTA retires the slot only when all of them are (and resets the count)
};

struct acx_anyres {  // resolution of a wildcard receive (device memory, one per request slot): which ring the matched message sits in
    acx_entry *ring;     // local ring of the matched (source, channel)
    acx_sack *sack;      // peer-mapped acks at the matched sender
    char *eager;         // local eager slots of that ring
    int32_t src, pad;
};

struct acx_host_status {  // pinned host memory, one per request slot; written by wait kernels
    int32_t source, tag, error, retired;
    int64_t bytes;
    int32_t trig_serial;      // overflow sends: low 32 bits of the request serial, stored by the trigger once it has drawn its ticket and deposited / announced the message
    int32_t pad;
};

// `retired` (and the status fields next to it) live in pinned host memory: kernels store, the host polls. On hardware
// that is a plain store into mapped memory and a volatile load; compiled for the software device (tests/emu) the same
// accesses are atomics, so ThreadSanitizer sees them as the synchronization they are.
#ifdef ACX_EMU
#define ACX_HOST_FLAG_SET(p, v) __atomic_store_n((p), (v), __ATOMIC_RELEASE)
#define ACX_HOST_FLAG_GET(p) __atomic_load_n((p), __ATOMIC_ACQUIRE)
#else
#define ACX_HOST_FLAG_SET(p, v) (*(p) = (v))
#define ACX_HOST_FLAG_GET(p) (*(const volatile int32_t *)(p))
#endif

struct acx_trace_rec {  // device-side trace ring (ACX_TRACE=1)
    uint64_t t_ns;      // %globaltimer
    uint32_t ev, peer;
    uint64_t a, b;
};
enum { ACX_EV_SEND_TRIGGER = 1, ACX_EV_RECV_TRIGGER, ACX_EV_MOVE_BEGIN, ACX_EV_MOVE_END, ACX_EV_WAIT_BEGIN, ACX_EV_WAIT_END };

// Geometry of a control window; identical on every rank (derived from env + nranks).
struct acx_geom {
    int nranks, nchan, depth, nops;
    uint32_t eager_max;
    uint64_t off_entries, off_sack, off_send_seq, off_recv_seq, off_ops, off_pflags, off_trace, off_xlat, off_eager, bytes;
    uint32_t npflags, ntrace, maxreg;
};

#ifdef __CUDACC__
#define ACX_HDI __host__ __device__ __forceinline__
#else
#define ACX_HDI inline
#endif
ACX_HDI uint64_t acx_entry_index(const acx_geom &g, int src, int chan, int e) { return ((uint64_t)src * g.nchan + chan) * g.depth + e; }

// Offsets of a control window from its configuration (nranks, nchan, depth, nops, eager_max, npflags, ntrace, maxreg
// must be filled in). One definition for the runtime and for the host-side protocol emulation (tests/emu).
inline void acx_geom_layout(acx_geom &g) {
    auto up = [](uint64_t x) { return (x + 255) & ~255ull; };
    const uint64_t n = (uint64_t)g.nranks, nslots = n * g.nchan * g.depth;
    uint64_t o = 0;
    g.off_entries = o;  o = up(o + nslots * sizeof(acx_entry));
    g.off_sack = o;     o = up(o + nslots * sizeof(acx_sack));
    g.off_send_seq = o; o = up(o + n * g.nchan * 8);
    g.off_recv_seq = o; o = up(o + n * g.nchan * 8);
    g.off_ops = o;      o = up(o + (uint64_t)g.nops * sizeof(acx_op));
    g.off_pflags = o;   o = up(o + (uint64_t)g.npflags * 8);
    g.off_trace = o;    o = up(o + (uint64_t)g.ntrace * sizeof(acx_trace_rec));
    g.off_xlat = o;     o = up(o + n * g.maxreg * 8);
    g.off_eager = o;    o = up(o + nslots * g.eager_max);
    g.bytes = o;
}

// Strided datatype layout as the pack / unpack kernels consume it. Work is split in `vec`-byte units (the widest power of
// two <= 16 dividing the run, every stride in use and both base addresses); unit j of the packed image belongs at
// acx_layout_unit_offset(L, j) bytes from the user's buffer address.
struct TypeLayout {
    long long n[3], stride[3];
    long long run, disp, extent, count;
    uint32_t vec, pad;
};
ACX_HDI uint64_t acx_layout_units(const TypeLayout &L, uint64_t limit_bytes) {
    const uint64_t units = (uint64_t)(L.count * L.n[2] * L.n[1] * L.n[0]) * ((uint64_t)L.run / L.vec);
    return units * L.vec > limit_bytes ? limit_bytes / L.vec : units;
}
ACX_HDI long long acx_layout_unit_offset(const TypeLayout &L, uint64_t j) {
    const uint64_t upr = (uint64_t)L.run / L.vec;   // units per contiguous run
    uint64_t r = j / upr;
    const uint64_t within = (j - r * upr) * L.vec;
    const uint64_t i0 = r % (uint64_t)L.n[0]; r /= (uint64_t)L.n[0];
    const uint64_t i1 = r % (uint64_t)L.n[1]; r /= (uint64_t)L.n[1];
    const uint64_t i2 = r % (uint64_t)L.n[2]; r /= (uint64_t)L.n[2];
    return (long long)r * L.extent + L.disp + (long long)i2 * L.stride[2] + (long long)i1 * L.stride[1] + (long long)i0 * L.stride[0] + (long long)within;
}

// Cooperative moving (ACX_COOP=1): chunk size for an n-byte rendezvous payload. Both ends evaluate this, so it may only
// depend on n. Small enough that two GPUs can share a message, large enough that the remote atomic per claim is noise.
ACX_HDI uint64_t acx_coop_chunk(uint64_t n) {
    const uint64_t unit = 64ull << 10;
    if (n <= (8ull << 20)) return unit;
    uint64_t c = (n / 128 + unit - 1) / unit * unit;
    return c > (1ull << 20) ? (1ull << 20) : c;
}

// ---- host-side state -------------------------------------------------------------------
#ifdef __cplusplus
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

enum AcxReqKind { ACX_REQ_SEND = 1, ACX_REQ_RECV = 2, ACX_REQ_PSEND = 3, ACX_REQ_PRECV = 4 };

// Datatype layout in canonical form (see ACX_Mpi_layout in the bundled mpi.h): element (c,i2,i1,i0) is `run` contiguous
// bytes at c*extent + disp + i2*stride[2] + i1*stride[1] + i0*stride[0].
struct AcxLayout {
    long long n[3], stride[3];
    long long run, disp, extent, count;
    long long span;           // bytes of user memory touched, measured from the buffer address
    bool noncontig;
};

inline TypeLayout acx_device_layout(const AcxLayout &a, const void *user, const void *packed) {
    TypeLayout t{};
    for (int d = 0; d < 3; d++) { t.n[d] = a.n[d]; t.stride[d] = a.stride[d]; }
    t.run = a.run; t.disp = a.disp; t.extent = a.extent; t.count = a.count;
    unsigned long long bits = (unsigned long long)a.run | (unsigned long long)a.disp | (uintptr_t)user | (uintptr_t)packed;
    for (int d = 0; d < 3; d++) if (a.n[d] > 1) bits |= (unsigned long long)a.stride[d];
    if (a.count > 1) bits |= (unsigned long long)a.extent;
    uint32_t vec = 16;
    while (vec > 1 && (bits & (vec - 1))) vec >>= 1;
    t.vec = vec;
    return t;
}

struct AcxGraphRes;  // resources owned by a CUDA graph (released by its user object)

struct AcxRequest {
    uint32_t magic;
    AcxReqKind kind;
    int peer, tag, comm_ctx, chan;   // peer: WORLD rank (indexes windows/rings)
    int comm_peer;            // the peer as the caller named it (rank in the operation's communicator): what MPI_Status reports
    int slot;                 // op slot (basic) / -1
    uint64_t serial;
    size_t bytes;             // message bytes (send) / capacity (recv)
    void *user_buf;           // as passed by the user
    char *dev_buf;            // device-accessible alias of user_buf
    int64_t bounce_off;       // heap offset of bounce buffer, -1 if none
    int64_t heap_off;         // heap offset of user_buf if it lives in the heap, else -1
    cudaStream_t stream;      // stream-mode operations: where the operation was enqueued (a wait on another stream is ordered behind it with an event)
    cudaEvent_t ev_trig, ev_done;  // large transfers: mover runs on a per-peer side stream (fork after the trigger / join at the wait)
    bool in_graph;
    bool host_transport;      // no-GPU mode: backed by an MPI_Request
    bool null_op;             // peer was MPI_PROC_NULL: completes immediately, moves nothing, owns no slot
    void *deferred;           // RECV posted while its channel's ring was full (> ACX_DEPTH receives outstanding): the trigger is launched later, in posting order, from a wait (DeferredRecv in acx.cu)
    uint64_t hticket;         // ticket this operation's trigger will draw, predicted on the host (~0 = unknown: graphs, several streams or wildcards on its channel); lets a stream wait be a memOp
    bool late;                // large message (>= ACX_LATE_PUSH_MIN): sender-moves protocol, the wait kernel carries helper CTAs
    char *src_local;          // SEND: device address the message is read from (user buffer, bounce or packed image)
    bool coop;                // ACX_COOP: large rendezvous operation whose mover claims chunks cooperatively; the wait launches a coop_kernel in front of the wait kernel
    bool typed;               // non-contiguous datatype: the wire image is staged in the heap (bounce_off), packed before the send / unpacked after the wait
    AcxLayout layout;
    bool wild;                // wildcard receive (MPI_ANY_SOURCE / MPI_ANY_TAG): matched by recv_any_kernel when the wait runs
    bool branched;            // ACX_PEER_BRANCH: all of this operation's kernels ran on the peer's side stream (the wait joins ev_done)
    bool overflow;            // SEND enqueued while its channel already had a ring's worth of un-waited sends (or earlier overflow sends still pending): its trigger runs on a pool stream with the host's ticket, its wait is a kernel that watches op->armed and comes AFTER the waits of everything else in the same call
    bool send_counted;        // SEND that is counted in AcxState::send_flow[peer][chan].unwaited (until its wait is enqueued / it is released)
    MPI_Request mpi_req;
    MPI_Status *user_status;
    AcxGraphRes *gres;
    // partitioned
    int partitions;
    size_t part_bytes;
    uint64_t epoch;
    bool active;
    bool resolved;
    int pair_rec;             // index of the handshake record
    int64_t land_off;         // RECV: landing buffer heap offset (may alias user buffer)
    bool land_is_user;
    uint32_t pf_arrived;      // RECV: index of arrived[] in my pflags; SEND: in the peer's
    uint32_t pf_local;        // SEND: sent[] ; RECV: copied[]
    uint32_t pf_cts;          // SEND: cts word index (mine); RECV: the sender's cts index
    uint32_t pf_epoch;        // my epoch word
    uint32_t pf_work;         // SEND: work[partitions], CTA counters of acx::pready_grid
    acx_pchan *pchan_d;       // device descriptor (lives in my pflags-adjacent area via cudaMalloc)
    acx_pchan pchan_h;
};

struct AcxHostReg { uintptr_t lo, hi; char *dptr; uint64_t pfn; };   // pfn: physical frame of the first page when it was pinned (0 = unknown): a different frame later means the application unmapped and re-mapped the range

// A communicator the library has seen: its matching context and the translation between its ranks and COMM_WORLD's
// (all windows, rings and side streams are indexed by world rank; the API speaks communicator ranks).
struct AcxCommMap {
    MPI_Comm comm;
    int ctx;                    // same number on every member (see acx_comm_ctx)
    int size, me;
    bool identity;              // every world rank, in world order
    std::vector<int> to_world;  // comm rank -> world rank
};

struct AcxState {
    bool initialized = false;
    bool device = false;      // GPU transport active
    int rank = 0, size = 1, dev = -1;
    int sm_count = 148;
    acx_geom geom{};
    // my windows
    char *ctrl = nullptr, *heap = nullptr;
    size_t heap_bytes = 0;
    // peer mappings (index = rank; self = own pointers)
    std::vector<char *> peer_ctrl, peer_heap;
    // pinned host mirrors
    acx_host_status *hstat = nullptr, *hstat_d = nullptr;
    cudaStream_t istream = nullptr;  // internal non-blocking stream (host waits, start/finish)
    // allocators
    std::recursive_mutex mu;
    std::vector<int> free_slots;
    std::vector<char> slot_busy;     // 0 free, 1 in use (stream), 2 in use (graph)
    std::vector<AcxRequest *> slot_req;
    std::map<size_t, size_t> heap_free;      // offset -> size (free blocks)
    std::unordered_map<size_t, size_t> heap_used;  // offset -> size
    uint32_t pflags_next = 0;
    std::vector<AcxHostReg> host_regs;
    std::vector<AcxCommMap> comms;           // known communicators (context id + rank translation)
    MPI_Comm icomm;                          // private dup of COMM_WORLD for handshakes
    std::atomic<uint64_t> serial{1};         // request serials (application threads enqueue concurrently)
    std::vector<std::vector<int>> recv_outstanding;  // [src][chan] posted-not-yet-waited recvs
    // Host-side ticket prediction per (peer, channel, direction): exact while every operation on the channel is enqueued in
    // stream mode on ONE stream (tickets are drawn by the trigger kernels in execution order); a graph, a second stream or a
    // wildcard receive makes the channel device-counted for good.
    std::vector<std::vector<std::vector<AcxRequest *>>> recv_deferred;   // [src][chan]: receives whose trigger has not been launched yet, in posting order
    struct ChanCount { uint64_t next = 0; cudaStream_t stream = nullptr; bool dev = false; };
    std::vector<std::vector<ChanCount>> send_count, recv_count;   // [peer][chan]
    // Sends in flight beyond a channel's ring depth (ACX_SEND_OVERFLOW, acx.cu: send_overflow_begin): stream-mode sends enqueued and not
    // yet waited on, per channel, and the newest overflow send whose trigger may still be waiting on its pool stream.
    struct SendFlow { int unwaited = 0; AcxRequest *last = nullptr; };   // last: the newest overflow send of the channel while its trigger may still be pending (acx_host_status::trig_serial tells); cleared when that request is released
    std::vector<std::vector<SendFlow>> send_flow;   // [peer][chan]
    // config
    bool trace = false;
    int verbose = 0;
    long watchdog_ms = 0;
    size_t mover_min_cta_bytes = 64 << 10;
    int mover_max_ctas = 0;
    // stats
    struct {
        std::atomic<unsigned long long> sends{0}, recvs{0}, waits{0}, bytes_sent{0}, bytes_recv{0}, eager{0}, rndv{0}, bounce{0},
            kernels{0}, graph_nodes{0}, slots_hw{0}, heap_hw{0}, ce{0}, send_overflow{0};
    } st;
    size_t heap_in_use = 0;
    int slots_in_use = 0;
    int pageable_access = 0;
    int host_ptr_is_dev_ptr = 0;
};

extern AcxState *g_acx;

#define ACX_REQ_MAGIC 0xACB200A1u

#define ACX_ERR(...)                                                                         \
    do {                                                                                     \
        fprintf(stderr, "[acx %d] %s:%d %s: ", g_acx ? g_acx->rank : -1, __FILE__, __LINE__, __func__); \
        fprintf(stderr, __VA_ARGS__);                                                        \
        fprintf(stderr, "\n");                                                               \
    } while (0)
#define ACX_LOG(lvl, ...)                                                  \
    do {                                                                   \
        if (g_acx && g_acx->verbose >= (lvl)) {                                \
            fprintf(stderr, "[acx %d] ", g_acx->rank);                       \
            fprintf(stderr, __VA_ARGS__);                                  \
            fprintf(stderr, "\n");                                         \
        }                                                                  \
    } while (0)
#define ACX_CUDA(stmt)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (stmt);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ACX_ERR("%s failed: %s", #stmt, cudaGetErrorString(_e));                         \
            return MPI_ERR_OTHER;                                                            \
        }                                                                                    \
    } while (0)

// allocators (init.cu)
int acx_slot_alloc(AcxRequest *r, bool graph);
void acx_slot_free(int slot);
void acx_reclaim_retired();
int64_t acx_heap_alloc(size_t bytes);
void acx_heap_free(int64_t off);
int64_t acx_heap_offset_of(const void *p);  // -1 if not inside my heap
uint32_t acx_pflags_alloc(uint32_t n);
// Device-accessible alias of an arbitrary user pointer (host memory gets pinned+mapped).
int acx_device_alias(const void *p, size_t bytes, char **out);
int acx_comm_ctx(MPI_Comm comm, int *ctx, const AcxCommMap **map = nullptr);
int acx_env_int(const char *name, const char *alt, long dflt, long *out);

// layout helpers
inline acx_entry *acx_entries_of(char *ctrl) { return (acx_entry *)(ctrl + g_acx->geom.off_entries); }
inline acx_sack *acx_sacks_of(char *ctrl) { return (acx_sack *)(ctrl + g_acx->geom.off_sack); }
inline uint64_t *acx_pflags_of(char *ctrl) { return (uint64_t *)(ctrl + g_acx->geom.off_pflags); }

#endif  // __cplusplus
#endif
