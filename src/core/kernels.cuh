// acx-b200 point-to-point kernels: triggers, data movers, waits.
//
// These replace the reference's four <<<1,1>>> flag kernels (set / wait / wait_and_set /
// waitall, src/sendrecv.cu:44-72), which only poke a host-memory flag for a CPU proxy.
// Here the kernels ARE the transport: they match, move payload over NVLink and signal.
#ifndef ACX_KERNELS_CUH
#define ACX_KERNELS_CUH

#include "internal.h"

namespace acxk {

using namespace acx;

constexpr int kTriggerThreads = 256;
constexpr int kMoverThreads = 512;
constexpr int kWaitThreads = 256;
constexpr int kWaitBatch = 12;

struct TraceCtx {
    acx_trace_rec *ring;  // null when tracing is off
    uint32_t mask;        // ring size - 1
    uint64_t *cursor;
};

// Programmatic dependent launch: let the NEXT kernel in the stream get scheduled while this one still runs,
// then wait for the PREVIOUS kernel's completion + memory flush before touching anything. Both instructions
// are no-ops when the launch did not opt in (ACX_PDL=0, graph nodes, user kernels in between).
#ifdef ACX_EMU   // host-side protocol emulation (tests/emu): kernels are called as functions in stream order
__device__ __forceinline__ void pdl_prologue() {}
__device__ __forceinline__ void pdl_wait_only() {}
__device__ __forceinline__ uint64_t globaltimer_ns() { return emu::now_ns(); }
#else
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ void pdl_wait_only() { asm volatile("griddepcontrol.wait;" ::: "memory"); }   // dependents start when this grid has finished

__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif
__device__ __forceinline__ void trace(const TraceCtx &t, uint32_t ev, uint32_t peer, uint64_t a, uint64_t b) {
    if (!t.ring) return;
    uint64_t i = atomicAdd((unsigned long long *)t.cursor, 1ull);
    acx_trace_rec *r = t.ring + (i & t.mask);
    r->t_ns = globaltimer_ns(); r->ev = ev; r->peer = peer; r->a = a; r->b = b;
}

// Watchdog: a spin that exceeds the budget records a diagnostic and traps, instead of
// hanging the GPU forever (the reference's spins have no timeout, src/sendrecv.cu:51).
struct Watchdog {
    uint64_t budget_ns;   // 0 = disabled
    uint64_t *diag;       // [0]=code [1]=a [2]=b (pinned host memory)
};
__device__ __forceinline__ void spin_geq(const uint64_t *p, uint64_t v, const Watchdog &wd, uint64_t code) {
    unsigned ns = 0;
    uint64_t t0 = wd.budget_ns ? globaltimer_ns() : 0;
    while (ld_acquire_sys(p) < v) {
        if (ns) __nanosleep(ns);
        ns = ns < 256 ? ns * 2 + 20 : 256;
        if (wd.budget_ns && globaltimer_ns() - t0 > wd.budget_ns) {
            wd.diag[1] = (uint64_t)p; wd.diag[2] = v; wd.diag[0] = code;
            __threadfence_system();
            __trap();
        }
    }
}
__device__ __forceinline__ void wd_fail(const Watchdog &wd, uint64_t code, uint64_t a, uint64_t b) {   // diagnosed fatal error
    if (wd.diag) { wd.diag[1] = a; wd.diag[2] = b; wd.diag[0] = code; }
    __threadfence_system();
    __trap();
}

// 256-bit (LDG.256 / STG.256, new on sm_100) copy loop: half the LSU instructions per byte and
// 1-KiB contiguous runs per warp instruction. Needs 32-B aligned pointers; tail via copy_bytes.
struct alignas(32) v8u32 { uint32_t r[8]; };
#ifdef ACX_EMU
__device__ __forceinline__ v8u32 ld256(const char *p) { v8u32 v; memcpy(&v, p, 32); return v; }
__device__ __forceinline__ void st256(char *p, const v8u32 &v) { memcpy(p, &v, 32); }
#else
__device__ __forceinline__ v8u32 ld256(const char *p) {
    v8u32 v;
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.r[0]), "=r"(v.r[1]), "=r"(v.r[2]), "=r"(v.r[3]), "=r"(v.r[4]), "=r"(v.r[5]), "=r"(v.r[6]), "=r"(v.r[7]) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st256(char *p, const v8u32 &v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(v.r[0]), "r"(v.r[1]), "r"(v.r[2]), "r"(v.r[3]), "r"(v.r[4]), "r"(v.r[5]), "r"(v.r[6]), "r"(v.r[7]) : "memory");
}
#endif
__device__ __forceinline__ void copy_bytes_256(char *dst, const char *src, uint64_t n, unsigned tid, unsigned nthr) {
    if ((((uintptr_t)dst | (uintptr_t)src) & 31) != 0) { copy_bytes(dst, src, n, tid, nthr); return; }
    const uint64_t nv = n >> 5;
    uint64_t i = tid;
    for (; i + 3ull * nthr < nv; i += 4ull * nthr) {
        v8u32 a = ld256(src + (i << 5)), b = ld256(src + ((i + nthr) << 5)), c = ld256(src + ((i + 2ull * nthr) << 5)), d = ld256(src + ((i + 3ull * nthr) << 5));
        st256(dst + (i << 5), a); st256(dst + ((i + nthr) << 5), b); st256(dst + ((i + 2ull * nthr) << 5), c); st256(dst + ((i + 3ull * nthr) << 5), d);
    }
    for (; i < nv; i += nthr) st256(dst + (i << 5), ld256(src + (i << 5)));
    const uint64_t done = nv << 5;
    if (done < n) copy_bytes(dst + done, src + done, n - done, tid, nthr);
}

// Streaming flavour (ACX_MOVER=ldstcs): ld.global.cs / st.global.cs mark every line evict-first, for copies whose
// working set is far beyond L2 (the 1-GPU self-transfer is HBM-bound: nothing it touches is read again).
#ifdef ACX_EMU
__device__ __forceinline__ int4 ld_cs(const int4 *p) { return *p; }
__device__ __forceinline__ void st_cs(int4 *p, int4 v) { *p = v; }
#else
__device__ __forceinline__ int4 ld_cs(const int4 *p) { return __ldcs(p); }
__device__ __forceinline__ void st_cs(int4 *p, int4 v) { __stcs(p, v); }
#endif
__device__ __forceinline__ void copy_bytes_cs(char *dst, const char *src, uint64_t n, unsigned tid, unsigned nthr) {
    if ((((uintptr_t)dst | (uintptr_t)src) & 15) != 0) { copy_bytes(dst, src, n, tid, nthr); return; }
    const uint64_t nv = n >> 4;
    const int4 *s = (const int4 *)src;
    int4 *d = (int4 *)dst;
    uint64_t i = tid;
    for (; i + 3ull * nthr < nv; i += 4ull * nthr) {
        int4 a = ld_cs(s + i), b = ld_cs(s + i + nthr), c = ld_cs(s + i + 2ull * nthr), e = ld_cs(s + i + 3ull * nthr);
        st_cs(d + i, a); st_cs(d + i + nthr, b); st_cs(d + i + 2ull * nthr, c); st_cs(d + i + 3ull * nthr, e);
    }
    for (; i < nv; i += nthr) st_cs(d + i, ld_cs(s + i));
    const uint64_t done = nv << 4;
    if (done < n) copy_bytes(dst + done, src + done, n - done, tid, nthr);
}
// Copy-loop flavour of the LDG/STG movers: 0 = 16-B vectors, 1 = 256-bit vectors, 2 = 16-B streaming.
__device__ __forceinline__ void copy_flavour(uint32_t flavour, char *dst, const char *src, uint64_t n, unsigned tid, unsigned nthr) {
    if (flavour == 1) copy_bytes_256(dst, src, n, tid, nthr);
    else if (flavour == 2) copy_bytes_cs(dst, src, n, tid, nthr);
    else copy_bytes(dst, src, n, tid, nthr);
}

// ---------------------------------------------------------------------------------------
// Plain copy (bounce-in for private send buffers). Grid-stride, 16-B vectors.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMoverThreads) copy_kernel(char *dst, const char *src, uint64_t n) {
    pdl_prologue();
    copy_bytes(dst, src, n, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// Resolve a peer buffer location (see ACX_OFF_REG_BIT) to an address this GPU can dereference.
__device__ __forceinline__ char *translate_peer(char *peer_heap, const uint64_t *xlat_row, uint64_t off) {
    if (off & ACX_OFF_REG_BIT) return (char *)(xlat_row[ACX_OFF_REG_IDX(off)] + ACX_OFF_REG_OFF(off));
    return peer_heap + off;
}

// ---------------------------------------------------------------------------------------
// SEND trigger: take a ticket, wait for slot credit, then either deposit eagerly or do
// the rendezvous handshake (remote atomic on the slot's state word).
// ---------------------------------------------------------------------------------------
struct SendParams {
    acx_entry *ring;        // peer-mapped: receiver's ring for (me -> dst, chan)
    char *eager;            // peer-mapped: receiver's eager slots for that ring
    acx_sack *sack;         // local: acks for (me -> dst, chan)
    uint64_t *seq;          // local: send ticket counter
    acx_op *op;             // local op record
    acx_host_status *hstat; // pinned host record (null: do not signal the host)
    const char *src;        // device-accessible source
    char *peer_heap;        // peer-mapped heap base of dst
    const uint64_t *xlat;   // local: translation row for registrations owned by dst
    uint64_t src_off;       // heap offset of the (possibly bounced) source; rendezvous only
    uint64_t bytes;
    uint64_t serial;
    uint32_t tag, depth, eager_max, peer;
    uint32_t ctx, late;     // ctx: communicator context, travels in the slot header (exact (source, tag, comm) matching: a receive checks it); late: announce RNDV_LATE (sender moves, now or at its wait)
    // Overflow sends (acx.cu: send_overflow_begin) run on a pool stream of their own, several of one channel at a time: the ticket is the
    // one the HOST counted (hticket1 - 1; 0 = draw from the counter), taken when the channel's counter reaches it, so tickets
    // keep the enqueue order without any stream order between the triggers; the trigger reports to the host (hdone) and to the
    // wait kernel (op->armed, for eager messages too). Nothing follows it on its stream: a rendezvous message gets the
    // single-launch kernel (send_rndv_fused) with a small grid.
    uint64_t hticket1;
    int32_t *hdone;         // pinned host word of the slot (acx_host_status::trig_serial): set to hserial once the ticket is drawn and the message deposited / announced
    uint32_t hserial, pad_ov;
    uint32_t apush, coop;   // apush: deterministic mode — this sender pushes the payload itself (gate + mover on its channel stream); coop: announce the message as RNDV_COOP (ACX_COOP)
    Watchdog wd;
    TraceCtx tr;
};

// Rendezvous half of the send trigger (one thread): publish the source, add to the slot's state word,
// and if the receive is already posted resolve the destination (this side will push).
__device__ __forceinline__ void send_rndv_match(const SendParams &p, uint64_t k) {
    const uint32_t e = (uint32_t)(k % p.depth);
    const uint64_t g = k / p.depth;
    acx_entry *ent = p.ring + e;

    ent->s_src_off = p.src_off; ent->s_bytes = p.bytes;
    *reinterpret_cast<uint64_t *>(&ent->s_tag) = ACX_ENT_HEADER(p.tag, p.apush ? ACX_KIND_RNDV_PUSH : p.coop ? ACX_KIND_RNDV_COOP : p.late ? ACX_KIND_RNDV_LATE : ACX_KIND_RNDV, p.ctx);
    uint64_t old = atom_add_acqrel_sys(&ent->state, 1ull << 32);
    p.op->ticket = k; p.op->eager = 0; p.op->cta_count = 0; p.op->coop = p.coop;
    if (p.apush) {
        p.op->mover = 0;   // decided by push_gate on the channel stream once the receive is posted
    } else if ((uint32_t)old == (uint32_t)(g + 1)) {  // receive already posted: I move the data (push)
        // the receiver mirrored its destination into MY sack before it bumped the state word — if it can take a rendezvous
        // message at all: a mirror of another generation means "too small", i.e. capacity 0 for this purpose
        const bool fresh = ld_relaxed_sys(&p.sack[e].r_gen) == g + 1;
        uint64_t cap = fresh ? ld_relaxed_sys(&p.sack[e].r_cap) : 0, off = fresh ? ld_relaxed_sys(&p.sack[e].r_dst_off) : 0;
        p.op->peer_ptr = (uint64_t)translate_peer(p.peer_heap, p.xlat, off);
        p.op->n = p.bytes < cap ? p.bytes : cap;
        if (p.op->n == 0) p.op->coop = 0;   // nothing to chunk (receive too small for a rendezvous message): the plain mover publishes completion
        if (p.late) st_relaxed_sys(&ent->claim, ((g + 1) << 2) | 1ull);   // RNDV_LATE: the mover behind this trigger pushes now — the receiver's fallback pull must stay out
        p.op->mover = 1;
    } else {
        p.op->mover = 0;
    }
    if (p.hdone) ACX_HOST_FLAG_SET(p.hdone, (int32_t)p.hserial);
    st_release_sys(&p.op->armed, p.serial);
    trace(p.tr, ACX_EV_SEND_TRIGGER, p.peer, k, p.bytes | (1ull << 63));
}

// A send's ticket on its channel (one thread): drawn from the channel's counter — or, for an overflow send, the one the host
// counted for it, taken when the counter gets there (the overflow triggers of a channel run on different streams).
__device__ __forceinline__ uint64_t send_take_ticket(const SendParams &p) {
    if (!p.hticket1) return atomicAdd((unsigned long long *)p.seq, 1ull);
    const uint64_t k = p.hticket1 - 1;
    spin_geq(p.seq, k, p.wd, 0x5e15);
    if (ld_relaxed_sys(p.seq) != k) wd_fail(p.wd, 0x5e16, k, ld_relaxed_sys(p.seq));   // somebody drew the ticket the host had counted for this send
    st_release_sys(p.seq, k + 1);
    return k;
}

__device__ __forceinline__ void send_body(const SendParams &p) {
    __shared__ uint64_t s_k;
    const unsigned tid = threadIdx.x;
    if (tid == 0) {
        const uint64_t k = send_take_ticket(p);
        uint32_t e = (uint32_t)(k % p.depth);
        uint64_t g = k / p.depth;
        spin_geq(&p.sack[e].efree, g, p.wd, 0x5e01);  // slot's previous generation consumed
        s_k = k;
    }
    __syncthreads();
    const uint64_t k = s_k;
    const uint32_t e = (uint32_t)(k % p.depth);
    acx_entry *ent = p.ring + e;
    if (p.bytes <= p.eager_max) {
        // Eager: payload + header into the receiver's slot, then ONE release (the barrier makes the
        // other threads' stores cumulative under thread 0's release).
        copy_bytes(p.eager + (uint64_t)e * p.eager_max, p.src, p.bytes, tid, blockDim.x);
        __syncthreads();
        if (tid == 0) {
            ent->s_bytes = p.bytes;
            *reinterpret_cast<uint64_t *>(&ent->s_tag) = ACX_ENT_HEADER(p.tag, ACX_KIND_EAGER, p.ctx);
            red_release_add_sys(&ent->state, 1ull << 32);
            st_release_sys(&ent->done, k / p.depth + 1);   // one completion word for every kind of message: what a stream memOp (or a wait kernel) watches
            // The source buffer is reusable as soon as this kernel retires: nothing to wait for later.
            if (p.hstat) ACX_HOST_FLAG_SET(&p.hstat->retired, 1);
            if (p.hdone) {   // overflow send: its wait is a kernel on another stream, which watches the op record
                ACX_HOST_FLAG_SET(p.hdone, (int32_t)p.hserial);
                p.op->ticket = k;
                st_release_sys(&p.op->armed, p.serial);
            }
            trace(p.tr, ACX_EV_SEND_TRIGGER, p.peer, k, p.bytes);
        }
    } else if (tid == 0) {
        send_rndv_match(p, k);
    }
}
__global__ void __launch_bounds__(kTriggerThreads) send_trigger(SendParams p) {
    pdl_prologue();
    send_body(p);
}
// ---------------------------------------------------------------------------------------
// RECV trigger: take a ticket, publish destination, add to the state word.
// ---------------------------------------------------------------------------------------
struct RecvParams {
    acx_entry *ring;        // local ring for (src -> me, chan)
    acx_sack *sack;         // peer-mapped: the sender's sack for (src -> me, chan); null if not rendezvous-capable
    uint64_t *seq;          // local recv ticket counter
    acx_op *op;
    char *peer_heap;        // peer-mapped heap base of src
    const uint64_t *xlat;   // local: translation row for registrations owned by src
    uint64_t dst_off;       // heap offset of destination (or bounce); ~0 when not rendezvous-capable
    uint64_t cap;
    uint64_t serial;
    uint32_t depth, peer;
    uint32_t rndv_capable;  // 0: capacity <= eager_max — a rendezvous-size message can only be a truncation error: it is completed with no bytes moved
    uint32_t publish;       // mirror {destination, capacity} into the sender's sack (rendezvous-capable receives; every receive in sender-push mode)
    Watchdog wd;
    TraceCtx tr;
};

__device__ __forceinline__ void recv_match_at(const RecvParams &p, uint64_t k) {   // one thread; k = this receive's ticket
    uint32_t e = (uint32_t)(k % p.depth);
    uint64_t g = k / p.depth;
    acx_entry *ent = p.ring + e;
    spin_geq(&ent->consumed, g, p.wd, 0x5e02);  // my own previous use of this slot is finished
    if (p.publish) {  // mirror the destination into the sender's memory: a late sender then needs no remote load
        // Small receives (capacity <= eager_max) stay silent — three remote stores less on the latency path. A late
        // rendezvous sender notices: the mirror's generation (r_gen) is not this slot generation's, so it moves nothing and
        // completes the (truncated) message instead of acting on what an earlier generation left behind.
        st_relaxed_sys(&p.sack[e].r_dst_off, p.rndv_capable ? p.dst_off : 0);
        st_relaxed_sys(&p.sack[e].r_cap, p.rndv_capable ? p.cap : 0);
        st_release_sys(&p.sack[e].r_gen, g + 1);   // a pushing sender's gate keys on this
    }
    uint64_t old = atom_add_acqrel_sys(&ent->state, 1ull);   // release orders the mirror stores before it
    p.op->ticket = k; p.op->cta_count = 0; p.op->mover = 0; p.op->bounce_n = 0; p.op->coop = 0;
    if ((uint32_t)(old >> 32) == (uint32_t)(g + 1) && (ACX_ENT_KIND(ent->s_kind) == ACX_KIND_RNDV || ACX_ENT_KIND(ent->s_kind) == ACX_KIND_RNDV_COOP || ACX_ENT_KIND(ent->s_kind) == ACX_KIND_RNDV_LATE) && !p.rndv_capable) {
        // a rendezvous message is waiting for a receive that cannot hold it (no mover was launched for this receive):
        // complete it here with nothing moved; the wait reports MPI_ERR_TRUNCATE and 0 bytes
        __threadfence_system();
        st_release_sys(&ent->done, g + 1);
        if (p.sack) st_release_sys(&p.sack[e].sdone, g + 1);
    } else if ((uint32_t)(old >> 32) == (uint32_t)(g + 1) && (ACX_ENT_KIND(ent->s_kind) == ACX_KIND_RNDV || ACX_ENT_KIND(ent->s_kind) == ACX_KIND_RNDV_COOP)) {
        p.op->coop = ACX_ENT_KIND(ent->s_kind) == ACX_KIND_RNDV_COOP;   // the sender decides how its message is moved
        // the send is already here and waiting: I move the data (pull)
        uint64_t sb = ent->s_bytes;
        p.op->peer_ptr = (uint64_t)translate_peer(p.peer_heap, p.xlat, ent->s_src_off);
        p.op->n = sb < p.cap ? sb : p.cap;
        p.op->mover = 1;
    }
    st_release_sys(&p.op->armed, p.serial);
    trace(p.tr, ACX_EV_RECV_TRIGGER, p.peer, k, p.cap);
}

__device__ __forceinline__ void recv_match(const RecvParams &p) {   // one thread
    recv_match_at(p, atomicAdd((unsigned long long *)p.seq, 1ull));
}

__global__ void recv_trigger(RecvParams p) {
    pdl_prologue();
    if (threadIdx.x == 0) recv_match(p);
}

// ---------------------------------------------------------------------------------------
// Mover: launched right after a rendezvous-capable trigger with a bandwidth-sized grid.
// Exits immediately unless the trigger found this side to be the second arriver.
//   push (sender second):  local src  -> peer dst   (STG over NVLink)
//   pull (receiver second): peer src  -> local dst  (LDG over NVLink)
// The last CTA to finish publishes completion to both sides.
// ---------------------------------------------------------------------------------------
struct MoverParams {
    acx_op *op;
    char *local;            // push: local source; pull: local destination
    acx_entry *ring;        // push: peer-mapped ring; pull: local ring
    acx_sack *sack;         // push: local sack; pull: peer-mapped sack at the sender
    uint32_t depth, push, peer, vec256;   // vec256: copy-loop flavour (copy_flavour)
    uint32_t block_bytes, pad1;   // 0: one contiguous chunk per CTA; else CTAs sweep interleaved blocks of this size (DRAM-page friendly)
    TraceCtx tr;
};

// Gate of the deterministic "sender always pushes" mode: runs on the sender's per-channel stream right before the
// mover; waits (on LOCAL memory) until the receiver has published its destination for this generation, then
// resolves it. Keeping the big mover grid out of the wait leaves the SMs free while the receive is late.
struct GateParams {
    acx_op *op;
    acx_sack *sack;          // local
    char *peer_heap;
    const uint64_t *xlat;
    uint64_t bytes;
    uint32_t depth, peer;
    Watchdog wd;
};
__global__ void push_gate(GateParams p) {
    pdl_prologue();
    if (threadIdx.x != 0) return;
    const uint64_t k = p.op->ticket;
    const uint32_t e = (uint32_t)(k % p.depth);
    const uint64_t g = k / p.depth;
    spin_geq(&p.sack[e].r_gen, g + 1, p.wd, 0x5e0a);
    const uint64_t cap = ld_relaxed_sys(&p.sack[e].r_cap), off = ld_relaxed_sys(&p.sack[e].r_dst_off);
    p.op->peer_ptr = (uint64_t)translate_peer(p.peer_heap, p.xlat, off);
    p.op->n = p.bytes < cap ? p.bytes : cap;
    if (p.op->n == 0) p.op->coop = 0;
    p.op->mover = 1;
}

// Copy-engine rendezvous (ACX_CE_MIN; host side: launch_sender_ce in acx.cu). The sender's host found the receiver's
// enqueue-time record of the matching receive (destination, capacity — both ranks count tickets on the host), so the
// destination address is known when the operation is enqueued and the payload can travel by cudaMemcpyAsync — the copy
// engines reach ~760 GB/s per direction where SM-issued stores stop at ~690 (tools/bwprobe.cu) and no SM is occupied.
// On the sender's (peer, channel) stream: ce_gate — wait until the receive has really been posted on the device (its
// mirror in MY sack carries this generation), check the host's prediction — then the copy, then ce_complete, which
// publishes completion to both sides exactly as the last CTA of a mover does. The message travels as RNDV_PUSH: the
// receiver never moves it.
struct CeParams {
    acx_op *op;
    acx_entry *ring;         // peer-mapped: the receiver's ring for (me -> dst, chan)
    acx_sack *sack;          // local
    uint64_t ticket;         // host-predicted ticket of this send
    uint64_t dst_off, bytes; // host-read destination code and message size
    uint64_t serial;         // != 0: no event orders this kernel behind the send's trigger — wait until the trigger has armed the op record for this request
    uint32_t depth, peer;
    Watchdog wd;
    TraceCtx tr;
};
__global__ void ce_gate(CeParams p) {
    pdl_prologue();
    if (threadIdx.x != 0) return;
    if (p.serial) {   // ACX_CE_SPINGATE: the cross-stream event hop (record + wait, several microseconds on the step's critical path) is replaced by this acquire spin
        unsigned ns = 0;
        const uint64_t t0 = p.wd.budget_ns ? globaltimer_ns() : 0;
        while (ld_acquire_sys(&p.op->armed) != p.serial) {
            if (ns) __nanosleep(ns);
            ns = ns < 256 ? ns * 2 + 20 : 256;
            if (p.wd.budget_ns && globaltimer_ns() - t0 > p.wd.budget_ns) wd_fail(p.wd, 0x5e13, (uint64_t)&p.op->armed, p.serial);
        }
    }
    const uint64_t k = p.op->ticket;
    if (k != p.ticket) { wd_fail(p.wd, 0x5e10, k, p.ticket); }   // the host mis-predicted the ticket: the copy behind this kernel would land in the wrong receive's buffer
    const uint32_t e = (uint32_t)(k % p.depth);
    const uint64_t g = k / p.depth;
    spin_geq(&p.sack[e].r_gen, g + 1, p.wd, 0x5e11);
    const uint64_t cap = ld_relaxed_sys(&p.sack[e].r_cap), off = ld_relaxed_sys(&p.sack[e].r_dst_off);
    if (off != p.dst_off || cap < p.bytes) { wd_fail(p.wd, 0x5e12, off, p.dst_off); }
    trace(p.tr, ACX_EV_MOVE_BEGIN, p.peer, k, p.bytes);
}
__global__ void ce_complete(CeParams p) {
    pdl_prologue();
    if (threadIdx.x != 0) return;
    const uint64_t k = p.ticket;
    const uint32_t e = (uint32_t)(k % p.depth);
    const uint64_t g = k / p.depth;
    __threadfence_system();
    st_release_sys(&p.ring[e].done, g + 1);   // receiver may consume (the copy before this kernel in stream order has landed)
    st_release_sys(&p.sack[e].sdone, g + 1);  // sender's buffer is reusable
    trace(p.tr, ACX_EV_MOVE_END, p.peer, k, p.bytes);
}

// Cooperative moving (ACX_COOP=1): the payload is cut into acx_coop_chunk(n)-byte chunks behind an atomic cursor in the
// RECEIVER's ring entry. Any CTA on either GPU may act as an agent: claim a chunk, copy it from its side (a sender pushes,
// a receiver pulls), count it done; whoever completes the last chunk publishes done / sdone. The second arriver's mover
// runs this loop, and so does the coop_kernel each side launches in front of its wait kernel, so a GPU that arrived
// first — and would otherwise idle while its neighbour copies — takes its share.
//   ent: the slot's entry at the receiver (local for a receiver, peer-mapped for a sender)
__device__ __forceinline__ void coop_agent(acx_entry *ent, acx_sack *sk, char *dst, const char *src, uint64_t n, uint64_t g, uint32_t vec256) {
    const uint64_t cb = acx_coop_chunk(n), C = (n + cb - 1) / cb;
    __shared__ uint64_t s_claim;
    for (;;) {
        if (threadIdx.x == 0) {
            // The cursor carries the generation in its high half: a straggling agent of a finished generation (the
            // receiver's wait kernel has already re-armed the slot) must not claim from the next one.
            uint64_t cur = ld_acquire_sys(&ent->coop_cursor), got = ~0ull;
            while ((uint32_t)(cur >> 32) == (uint32_t)g && (uint32_t)cur < C) {
                const uint64_t seen = atom_cas_acqrel_sys(&ent->coop_cursor, cur, cur + 1);
                if (seen == cur) { got = (uint32_t)cur; break; }
                cur = seen;
            }
            s_claim = got;
        }
        __syncthreads();
        const uint64_t i = s_claim;
        __syncthreads();
        if (i >= C) return;
        const uint64_t lo = i * cb, len = n - lo < cb ? n - lo : cb;
        copy_flavour(vec256, dst + lo, src + lo, len, threadIdx.x, blockDim.x);
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (atom_add_acqrel_sys(&ent->coop_done, 1ull) == C - 1) {   // every chunk is in place (each agent released its copies)
                st_release_sys(&ent->done, g + 1);    // receiver may consume
                st_release_sys(&sk->sdone, g + 1);    // sender's buffer is reusable
            }
        }
    }
}

// FL >= 0: one copy-loop flavour, no cooperative moving, compiled on its own — 64 registers, no spills, two 512-thread
// CTAs per SM (what a bandwidth-sized grid needs: the copy is latency-bound, profiles/mover_ncu.md). FL < 0: the general
// body (run-time flavour + the chunk-stealing agent), 108-112 registers = one CTA per SM — fine for the 8-32-CTA fused
// kernels and the cooperative mode, and what the plain mover had silently become once the 256-bit loop was added.
template <int FL>
__device__ __forceinline__ void mover_body_t(const MoverParams &p) {
    acx_op *op = p.op;
    if (!op->mover) return;
    const uint64_t n = op->n;
    char *peer = (char *)op->peer_ptr;
    const uint32_t flavour = FL < 0 ? p.vec256 : (uint32_t)FL;
    if (blockIdx.x == 0 && threadIdx.x == 0) trace(p.tr, ACX_EV_MOVE_BEGIN, p.peer, op->ticket, n);
    if (FL >= 0) {
        if (op->coop) __trap();   // a peer announced a cooperative message while this rank runs without ACX_COOP: the environment must be the same on every rank
    } else if (op->coop) {
        const uint64_t k = op->ticket;
        const uint32_t e = (uint32_t)(k % p.depth);
        coop_agent(p.ring + e, p.sack + e, p.push ? peer : p.local, p.push ? p.local : peer, n, k / p.depth, p.vec256);
        return;
    }
    if (p.block_bytes) {
        // Interleaved blocks: at any instant the resident CTAs cover one contiguous window of the buffers, so the
        // DRAM pages they touch (source and destination) stay few — the access pattern of a plain elementwise copy.
        const uint64_t bs = p.block_bytes;
        for (uint64_t lo = (uint64_t)blockIdx.x * bs; lo < n; lo += (uint64_t)gridDim.x * bs) {
            const uint64_t len = n - lo < bs ? n - lo : bs;
            char *d = p.push ? peer + lo : p.local + lo;
            const char *s = p.push ? p.local + lo : peer + lo;
            copy_flavour(flavour, d, s, len, threadIdx.x, blockDim.x);
        }
    }
    // Contiguous chunk per CTA: keeps each CTA's NVLink stream sequential.
    const uint64_t per = ((n + gridDim.x - 1) / gridDim.x + 127) & ~127ull;
    const uint64_t lo = per * blockIdx.x;
    if (!p.block_bytes && lo < n) {
        const uint64_t len = n - lo < per ? n - lo : per;
        char *d = p.push ? peer + lo : p.local + lo;
        const char *s = p.push ? p.local + lo : peer + lo;
        copy_flavour(flavour, d, s, len, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();   // one fence per CTA, cumulative over the CTA's stores (ordered by the barrier)
        unsigned t = atomicAdd(&op->cta_count, 1u);
        if (t == gridDim.x - 1) {
            // Every CTA fenced its own stores at system scope before it counted itself in; observing the full count (a
            // gpu-scope atomic) makes them cumulative HERE, and the flags below must carry that to the other GPU: release
            // at SYSTEM scope (a gpu-scope fence + relaxed.sys store is not a system-scope release in the PTX model).
            const uint64_t k = op->ticket;
            const uint32_t e = (uint32_t)(k % p.depth);
            const uint64_t g = k / p.depth;
            op->cta_count = 0;
            __threadfence_system();
            st_release_sys(&p.ring[e].done, g + 1);   // receiver may consume
            st_release_sys(&p.sack[e].sdone, g + 1);  // sender's buffer is reusable
            trace(p.tr, ACX_EV_MOVE_END, p.peer, k, n);
        }
    }
}

__device__ __forceinline__ void mover_body(const MoverParams &p) { mover_body_t<-1>(p); }

__global__ void __launch_bounds__(kMoverThreads) mover(MoverParams p) {   // general: ACX_COOP=1
    pdl_prologue();
    mover_body(p);
}
// What launch_mover() picks unless ACX_COOP is set: one kernel per ACX_MOVER flavour (ldst = default, ldst256, ldstcs).
__global__ void __launch_bounds__(kMoverThreads, 2) mover_plain(MoverParams p) { pdl_prologue(); mover_body_t<0>(p); }
__global__ void __launch_bounds__(kMoverThreads) mover_v256(MoverParams p) { pdl_prologue(); mover_body_t<1>(p); }
__global__ void __launch_bounds__(kMoverThreads, 2) mover_cs(MoverParams p) { pdl_prologue(); mover_body_t<2>(p); }

// Single-launch rendezvous for small grids (all CTAs co-resident): CTA 0 runs the match and publishes the
// decision; the others wait for it on a launch-indexed word (relaunch-safe in graphs), then everyone moves.
// Saves one kernel launch + inter-kernel gap per message on each side.
__device__ __forceinline__ void fused_exit(acx_op *op) {
    // Every CTA is past the gate once it gets here: the last one out re-arms the gate for the next launch
    // (graph relaunch) or the next owner of this request slot.
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned old = atomicAdd(&op->pad, 1u);
        if (old == gridDim.x - 1) { op->pad = 0; st_relaxed_sys(&op->fused_decided, 0); }
    }
}
__global__ void __launch_bounds__(kMoverThreads) send_rndv_fused(SendParams sp, MoverParams mp) {
    pdl_prologue();
    acx_op *op = sp.op;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            const uint64_t k = send_take_ticket(sp);
            spin_geq(&sp.sack[k % sp.depth].efree, k / sp.depth, sp.wd, 0x5e01);
            send_rndv_match(sp, k);
            __threadfence();
            st_release_sys(&op->fused_decided, 1);
        } else {
            spin_geq(&op->fused_decided, 1, sp.wd, 0x5e09);
        }
    }
    __syncthreads();
    mover_body(mp);
    fused_exit(op);
}
__global__ void __launch_bounds__(kMoverThreads) recv_rndv_fused(RecvParams rp, MoverParams mp) {
    pdl_prologue();
    acx_op *op = rp.op;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            recv_match(rp);
            __threadfence();
            st_release_sys(&op->fused_decided, 1);
        } else {
            spin_geq(&op->fused_decided, 1, rp.wd, 0x5e09);
        }
    }
    __syncthreads();
    mover_body(mp);
    fused_exit(op);
}

// ---------------------------------------------------------------------------------------
// Wildcard receive (MPI_ANY_SOURCE and/or MPI_ANY_TAG; experimental, ACX_WILDCARDS=1). Matching is deferred to the
// WAIT: nothing runs when the receive is enqueued (so no kernel ever spins at Irecv time), and the wait emits this
// kernel in front of the ordinary wait kernel. CTA 0 scans the candidate (source, channel) rings for the next
// unclaimed ticket whose send has already arrived (and whose tag fits), claims that ticket with a CAS on the same
// counter ordinary receives fetch-add, and from there behaves like an ordinary receive that arrived second: it
// publishes its destination, and pulls if the message is a rendezvous. The resolved ring/ack/eager pointers and the
// source are left in `res` for the wait kernel, whose own parameters cannot know them.
// ---------------------------------------------------------------------------------------
constexpr int kAnyMaxRanks = 64;
struct AnyTable {   // device copy of the peer window bases (built on first wildcard use)
    char *peer_ctrl[kAnyMaxRanks];
    char *peer_heap[kAnyMaxRanks];
};
struct AnyParams {
    const AnyTable *tbl;
    char *ctrl;                 // my control window
    acx_anyres *res;            // per-slot resolution record (device memory)
    acx_op *op;
    char *dst;                  // local landing address for a rendezvous pull (heap buffer or bounce)
    uint64_t dst_off, cap, serial;
    acx_geom geom;
    int32_t src_lo, src_hi, chan_lo, chan_hi;   // inclusive candidate ranges
    int32_t want_tag;           // MPI_ANY_TAG = any
    uint32_t want_ctx;          // the receive's communicator context: channels are per (ctx, tag), the header says which
    uint32_t me, vec256, rndv_capable;
    Watchdog wd;
    TraceCtx tr;
};

__global__ void __launch_bounds__(kMoverThreads) recv_any_kernel(AnyParams p) {
    pdl_prologue();
    acx_op *op = p.op;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            const acx_geom &G = p.geom;
            acx_entry *entries = (acx_entry *)(p.ctrl + G.off_entries);
            uint64_t *seqs = (uint64_t *)(p.ctrl + G.off_recv_seq);
            unsigned ns = 0;
            const uint64_t t0 = p.wd.budget_ns ? globaltimer_ns() : 0;
            int src = -1, chan = -1;
            uint64_t k = 0;
            for (;;) {
                for (int s = p.src_lo; s <= p.src_hi && src < 0; s++) {
                    for (int c = p.chan_lo; c <= p.chan_hi; c++) {
                        uint64_t *seq = seqs + ((uint64_t)s * G.nchan + c);
                        const uint64_t kk = ld_relaxed_sys(seq);
                        const uint32_t e = (uint32_t)(kk % G.depth);
                        const uint64_t g = kk / G.depth;
                        acx_entry *ent = entries + acx_entry_index(G, s, c, e);
                        if (ld_acquire_sys(&ent->consumed) < g) continue;                                   // my previous use of the slot is not finished
                        if ((uint32_t)(ld_acquire_sys(&ent->state) >> 32) < (uint32_t)(g + 1)) continue;    // the send with ticket kk is not here yet
                        if (ACX_ENT_CTX(ent->s_kind) != p.want_ctx) continue;                               // a message of another communicator
                        if (p.want_tag != MPI_ANY_TAG && (int32_t)ent->s_tag != p.want_tag) continue;       // head of this channel carries another tag
                        if (atomicCAS((unsigned long long *)seq, (unsigned long long)kk, (unsigned long long)(kk + 1)) != kk) continue;   // a concurrent receive took it
                        src = s; chan = c; k = kk;
                        break;
                    }
                }
                if (src >= 0) break;
                if (ns) __nanosleep(ns);
                ns = ns < 1024 ? ns * 2 + 20 : 1024;
                if (p.wd.budget_ns && globaltimer_ns() - t0 > p.wd.budget_ns) {
                    p.wd.diag[1] = (uint64_t)(int64_t)p.src_lo; p.wd.diag[2] = (uint64_t)(int64_t)p.want_tag; p.wd.diag[0] = 0x5e0b;
                    __threadfence_system();
                    __trap();
                }
            }
            const uint64_t ring_idx = acx_entry_index(G, src, chan, 0);
            acx_sack *peer_sack = (acx_sack *)(p.tbl->peer_ctrl[src] + G.off_sack) + acx_entry_index(G, (int)p.me, chan, 0);
            RecvParams rp;
            rp.ring = entries + ring_idx;
            rp.sack = peer_sack; rp.rndv_capable = p.rndv_capable; rp.publish = p.rndv_capable;
            rp.seq = nullptr;
            rp.op = op;
            rp.peer_heap = p.tbl->peer_heap[src];
            rp.xlat = (const uint64_t *)(p.ctrl + G.off_xlat) + (uint64_t)src * G.maxreg;
            rp.dst_off = p.dst_off; rp.cap = p.cap; rp.serial = p.serial; rp.depth = (uint32_t)G.depth; rp.peer = (uint32_t)src;
            rp.wd = p.wd; rp.tr = p.tr;
            p.res->ring = rp.ring; p.res->sack = peer_sack; p.res->eager = p.ctrl + G.off_eager + ring_idx * G.eager_max; p.res->src = src;
            recv_match_at(rp, k);
            {   // A sender-moves (RNDV_LATE) message: this resolver runs at wait time, where moving data is legitimate, and the
                // sender's own wait may be ordered behind something this receive gates — claim it and pull now unless the sender
                // already has (no grace period: a wildcard receive only ever matches messages that have arrived).
                acx_entry *ent = rp.ring + (uint32_t)(k % G.depth);
                const uint64_t g = k / G.depth;
                if (!op->mover && p.rndv_capable && ACX_ENT_KIND(ent->s_kind) == ACX_KIND_RNDV_LATE && ld_acquire_sys(&ent->done) < g + 1) {
                    uint64_t cur = ld_relaxed_sys(&ent->claim);
                    bool mine = false;
                    while ((cur >> 2) != g + 1) {
                        const uint64_t seen = atom_cas_acqrel_sys(&ent->claim, cur, ((g + 1) << 2) | 2ull);
                        if (seen == cur) { mine = true; break; }
                        cur = seen;
                    }
                    if (mine) {
                        const uint64_t sb = ent->s_bytes;
                        op->peer_ptr = (uint64_t)translate_peer(rp.peer_heap, rp.xlat, ent->s_src_off);
                        op->n = sb < p.cap ? sb : p.cap;
                        op->mover = 1;
                    }
                }
            }
            __threadfence();
            st_release_sys(&op->fused_decided, 1);
        } else {
            spin_geq(&op->fused_decided, 1, p.wd, 0x5e09);
        }
    }
    __syncthreads();
    MoverParams mp;
    mp.op = op; mp.local = p.dst; mp.ring = p.res->ring; mp.sack = p.res->sack;
    mp.depth = (uint32_t)p.geom.depth; mp.push = 0; mp.peer = (uint32_t)p.res->src; mp.vec256 = p.vec256; mp.block_bytes = 0; mp.pad1 = 0; mp.tr = p.tr;
    mover_body(mp);
    fused_exit(op);
}

// ---------------------------------------------------------------------------------------
// TMA flavour of the mover (ACX_MOVER=tma): one elected thread per CTA drives a 4-stage ring of
// 16-KiB bulk copies through shared memory — cp.async.bulk global->smem (mbarrier complete_tx),
// then cp.async.bulk smem->global (bulk_group). The SM issues two instructions per 16 KiB instead
// of 2048 LDG/STG pairs; peer addresses are served over NVLink by the TMA unit directly.
// Falls back to the LDG/STG loop for pointers/sizes that are not 16-B aligned.
// ---------------------------------------------------------------------------------------
#ifndef ACX_EMU   // inline PTX throughout; the emulation exercises the LDG/STG mover
constexpr int kTmaStages = 4;
constexpr uint32_t kTmaChunk = 16384;   // 4 x 16 KiB per CTA: three CTAs fit per SM, so a pushing and a pulling mover can co-reside
constexpr int kTmaThreads = 128;
constexpr int kTmaSmem = kTmaStages * (int)kTmaChunk + 64;

__device__ __forceinline__ uint32_t smem_addr_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kTmaThreads) mover_tma(MoverParams p) {
    extern __shared__ __align__(128) unsigned char tsm[];
    pdl_prologue();
    acx_op *op = p.op;
    if (!op->mover) return;
    const uint64_t n = op->n;
    char *peer = (char *)op->peer_ptr;
    const char *src = p.push ? p.local : peer;
    char *dst = p.push ? peer : p.local;
    if (blockIdx.x == 0 && threadIdx.x == 0) trace(p.tr, ACX_EV_MOVE_BEGIN, p.peer, op->ticket, n);
    const uint64_t per = ((n + gridDim.x - 1) / gridDim.x + kTmaChunk - 1) / kTmaChunk * kTmaChunk;
    const uint64_t lo = per * blockIdx.x;
    if (lo < n) {
        const uint64_t len = n - lo < per ? n - lo : per;
        const bool aligned = ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) && (len % 16 == 0);
        if (!aligned) {
            copy_bytes(dst + lo, src + lo, len, threadIdx.x, blockDim.x);
        } else if (threadIdx.x == 0) {
            uint64_t *bars = reinterpret_cast<uint64_t *>(tsm + kTmaStages * kTmaChunk);
            for (int i = 0; i < kTmaStages; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&bars[i])) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint64_t nchunks = (len + kTmaChunk - 1) / kTmaChunk;
            uint32_t phase_bits = 0;
            auto load = [&](uint64_t c) {
                const int st = (int)(c % kTmaStages);
                const uint32_t bytes = (uint32_t)(len - c * kTmaChunk < kTmaChunk ? len - c * kTmaChunk : kTmaChunk);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(&bars[st])), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_addr_u32(tsm + st * kTmaChunk)), "l"(src + lo + c * kTmaChunk), "r"(bytes), "r"(smem_addr_u32(&bars[st])) : "memory");
            };
            const uint64_t pre = nchunks < (uint64_t)kTmaStages ? nchunks : (uint64_t)kTmaStages;
            for (uint64_t c = 0; c < pre; c++) load(c);
            for (uint64_t c = 0; c < nchunks; c++) {
                const int st = (int)(c % kTmaStages);
                const uint32_t parity = (phase_bits >> st) & 1;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tTW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni TD;\n\tbra.uni TW;\n\tTD:\n\t}"
                    ::"r"(smem_addr_u32(&bars[st])), "r"(parity) : "memory");
                phase_bits ^= 1u << st;
                const uint32_t bytes = (uint32_t)(len - c * kTmaChunk < kTmaChunk ? len - c * kTmaChunk : kTmaChunk);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + lo + c * kTmaChunk), "r"(smem_addr_u32(tsm + st * kTmaChunk)), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (c >= 1 && c - 1 + kTmaStages < nchunks) {
                    // store(c-1) has finished READING its stage once at most one group (store c) is pending
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    load(c - 1 + kTmaStages);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete (writes performed)
            asm volatile("fence.proxy.async;" ::: "memory");
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        unsigned t = atomicAdd(&op->cta_count, 1u);
        if (t == gridDim.x - 1) {
            const uint64_t k = op->ticket;
            const uint32_t e = (uint32_t)(k % p.depth);
            const uint64_t g = k / p.depth;
            op->cta_count = 0;
            __threadfence_system();
            st_release_sys(&p.ring[e].done, g + 1);
            st_release_sys(&p.sack[e].sdone, g + 1);
            trace(p.tr, ACX_EV_MOVE_END, p.peer, k, n);
        }
    }
}

#endif  // !ACX_EMU

// Launched in front of the wait kernel for a large rendezvous request when ACX_COOP=1: coop_gate (one thread; waiting
// here is legitimate, the user asked for a wait) returns once the peer has arrived or the transfer is complete, then
// coop_kernel's CTAs join the agents of this message.
struct CoopParams {
    acx_op *op;
    acx_entry *ring;      // the (src -> dst, chan) ring at the receiver: local for a receive, peer-mapped for a send
    acx_sack *sack;       // acks at the sender: peer-mapped for a receive, local for a send
    char *local;          // send: local source; receive: local landing address
    char *peer_heap;
    const uint64_t *xlat;
    uint64_t bytes;       // send: message bytes; receive: capacity
    uint32_t depth, is_send, vec256, pad;
    Watchdog wd;
};
// One look at the slot (one thread): has the peer arrived, and if so what is there to move? Returns false while neither
// the peer has arrived nor the transfer completed; *n == 0 on return true means "nothing for an agent to do".
__device__ __forceinline__ bool coop_peek(const CoopParams &p, acx_entry *ent, acx_sack *sk, uint64_t g, uint64_t *n, uint64_t *peer) {
    *n = 0; *peer = 0;
    if (p.is_send) {
        if (ld_acquire_sys(&sk->sdone) >= g + 1) return true;                     // already complete
        if (ld_acquire_sys(&sk->r_gen) < g + 1) return false;                     // the receive is not posted yet
        const uint64_t cap = ld_relaxed_sys(&sk->r_cap);                          // its destination is mirrored into my acks
        *peer = (uint64_t)translate_peer(p.peer_heap, p.xlat, ld_relaxed_sys(&sk->r_dst_off));
        *n = p.bytes < cap ? p.bytes : cap;
        return true;
    }
    if (ld_acquire_sys(&ent->done) >= g + 1) return true;
    if ((uint32_t)(ld_acquire_sys(&ent->state) >> 32) < (uint32_t)(g + 1)) return false;   // the send has not arrived yet
    if (ACX_ENT_KIND(ent->s_kind) == ACX_KIND_RNDV_COOP) {                                       // its header is in my entry; only messages announced as cooperative have a cursor
        const uint64_t sb = ent->s_bytes;
        *peer = (uint64_t)translate_peer(p.peer_heap, p.xlat, ent->s_src_off);
        *n = sb < p.bytes ? sb : p.bytes;
    }
    return true;
}
// The waiting half, one thread: returns once the peer has arrived (or the transfer is complete), so the agent grid that
// follows never spins and a late peer costs one resident warp, not half the GPU.
__global__ void coop_gate(CoopParams p) {
    pdl_prologue();
    if (threadIdx.x != 0) return;
    const uint64_t k = p.op->ticket;
    const uint32_t e = (uint32_t)(k % p.depth);
    uint64_t n, peer;
    unsigned ns = 0;
    const uint64_t t0 = p.wd.budget_ns ? globaltimer_ns() : 0;
    while (!coop_peek(p, p.ring + e, p.sack + e, k / p.depth, &n, &peer)) {
        if (ns) __nanosleep(ns);
        ns = ns < 256 ? ns * 2 + 20 : 256;
        if (p.wd.budget_ns && globaltimer_ns() - t0 > p.wd.budget_ns) { p.wd.diag[1] = k; p.wd.diag[2] = p.is_send; p.wd.diag[0] = 0x5e0c; __threadfence_system(); __trap(); }
    }
}
__global__ void __launch_bounds__(kMoverThreads) coop_kernel(CoopParams p) {
    pdl_prologue();
    __shared__ uint64_t s_n, s_peer;
    const uint64_t k = p.op->ticket;
    const uint32_t e = (uint32_t)(k % p.depth);
    const uint64_t g = k / p.depth;
    acx_entry *ent = p.ring + e;
    acx_sack *sk = p.sack + e;
    if (threadIdx.x == 0) {
        uint64_t n, peer;
        if (!coop_peek(p, ent, sk, g, &n, &peer)) n = 0;   // cannot happen behind coop_gate; never spin here
        s_n = n; s_peer = peer;
    }
    __syncthreads();
    const uint64_t n = s_n;
    if (!n) return;
    char *peer = (char *)s_peer;
    coop_agent(ent, sk, p.is_send ? peer : p.local, p.is_send ? p.local : peer, n, g, p.vec256);
}

// ---------------------------------------------------------------------------------------
// Waits. One CTA per request, up to kWaitBatch requests per launch (Waitall = 1 launch).
// ---------------------------------------------------------------------------------------
struct WaitOne {
    acx_op *op;
    acx_entry *ring;        // recv: local ring;  send: unused
    acx_sack *sack;         // recv: peer-mapped sack at the sender; send: local sack
    char *eager;            // recv: local eager slots of the ring
    char *dst;              // recv: final destination (device-accessible)
    char *bounce;           // recv: bounce buffer the mover filled (null if dst is in the heap)
    acx_host_status *hstat; // pinned host record of this slot (device alias)
    char *ustatus;          // device alias of the user's MPI_Status (null = ignore)
    uint64_t cap;
    uint64_t serial;        // != 0: first wait until op->armed == serial (host-side waits)
    int32_t src, tag;
    uint32_t depth, eager_max;
    uint32_t ctx, pad0;     // the receive's communicator context (checked against the message header)
    // RNDV_LATE (large messages; see late_helper): `late` marks a send that may have to push from this wait, or a receive
    // that may have to fall back to pulling; helper CTAs help_lo .. help_lo+help_n-1 (counted behind the waiter CTAs) serve it
    char *late_local;       // send: local source; recv: local landing address (heap buffer or bounce)
    char *peer_heap;        // the peer's heap, peer-mapped
    const uint64_t *xlat;   // translation row for registrations owned by the peer
    uint64_t late_bytes;    // send: message bytes
    uint16_t late, help_lo, help_n, trigger_only;   // trigger_only: eager overflow send — complete once its trigger (on a pool stream) has run, i.e. op->armed == serial
    uint64_t hticket;          // host-predicted ticket (~0: none): checked against the ticket the trigger drew
    uint32_t late_tag, pad2;   // low bits of the request serial: together with the ticket it makes the op record's late_state word unique per (request, launch)
    uint16_t is_recv, retire;
    uint16_t bounce_kernel, coop;   // bounce_kernel: a bounce_out_kernel follows (large private destination); coop: reset the slot's chunk cursor when consuming (ACX_COOP)
    uint16_t off_source, off_tag, off_error, off_bytes;  // field offsets inside MPI_Status (0xffff = absent)
    const acx_anyres *any;  // wildcard receive: ring / sack / eager / src come from here (written by recv_any_kernel)
};
struct WaitBatch {
    WaitOne w[kWaitBatch];
    Watchdog wd;
    TraceCtx tr;
    uint32_t nreq, pad;       // CTAs 0 .. nreq-1 are waiters (one per request); the rest are RNDV_LATE helpers
    uint64_t fallback_ns;     // RNDV_LATE: how long a receive's wait watches an unclaimed message before it pulls it itself
    uint32_t early_release_max, pad_er;   // helper CTAs up to which the kernel still releases its programmatic dependents early
};

// late_state = [request serial, 30 bits | ticket + 1, 32 bits | code, 2 bits]: an op record is reused by later requests (which may
// draw the same ticket on another channel) and by graph relaunches (same request, next ticket) — a stale word must never match.
__device__ __forceinline__ uint64_t late_tagv(const WaitOne &w, uint64_t k) { return ((uint64_t)(w.late_tag & 0x3fffffffu) << 34) | (((k + 1) & 0xffffffffull) << 2); }

__device__ __forceinline__ void wait_body(const WaitBatch &b, const WaitOne &w) {
    __shared__ uint64_t s_n;
    __shared__ int s_eager;
    const unsigned tid = threadIdx.x;
    acx_op *op = w.op;
    if (tid == 0) {
        if (w.serial) {
            unsigned ns = 0;
            uint64_t t0 = b.wd.budget_ns ? globaltimer_ns() : 0;
            while (ld_acquire_sys(&op->armed) != w.serial) {
                __nanosleep(ns); ns = ns < 256 ? ns * 2 + 20 : 256;
                if (b.wd.budget_ns && globaltimer_ns() - t0 > b.wd.budget_ns) { b.wd.diag[1] = w.serial; b.wd.diag[0] = 0x5e05; __threadfence_system(); __trap(); }
            }
        }
        const uint64_t k = op->ticket;
        if (w.hticket != ~0ull && k != w.hticket) {   // the host predicted this ticket and built memOp waits on it: a wrong prediction must never pass silently
            b.wd.diag[1] = k; b.wd.diag[2] = w.hticket; b.wd.diag[0] = 0x5e0f;
            __threadfence_system();
            __trap();
        }
        const uint32_t e = (uint32_t)(k % w.depth);
        const uint64_t g = k / w.depth;
        trace(b.tr, ACX_EV_WAIT_BEGIN, w.src, k, w.is_recv);
        if (!w.is_recv && w.trigger_only) {
            s_n = 0; s_eager = 0;   // the trigger has run (armed): the payload is in the receiver's slot, the buffer is reusable
        } else if (!w.is_recv) {
            if (w.late) {
                // RNDV_LATE: if the receive was posted when the trigger ran, the mover launched behind it is pushing (or has
                // pushed) — nothing to do here. Otherwise nobody has moved anything yet and the RECEIVER will not pull (its
                // trigger leaves LATE messages alone, so that every GPU copies exactly its own sends, whatever the arrival
                // order): as soon as the receive is posted (its mirror lands in my acks), claim the message and let this
                // kernel's helper CTAs push it. Spinning here is what the user asked for; nothing spins at Isend time.
                const uint64_t tagv = late_tagv(w, k);
                if (op->mover) {
                    st_release_sys(&op->late_state, tagv | 1ull);
                    spin_geq(&w.sack[e].sdone, g + 1, b.wd, 0x5e03);
                } else {
                    unsigned ns = 0;
                    const uint64_t t0 = b.wd.budget_ns ? globaltimer_ns() : 0;
                    for (;;) {
                        if (ld_acquire_sys(&w.sack[e].sdone) >= g + 1) { st_release_sys(&op->late_state, tagv | 1ull); break; }   // the receiver's fallback pull (or a truncation) completed it
                        if (ld_acquire_sys(&w.sack[e].r_gen) == g + 1) {   // the receive is posted: its destination is mirrored into my acks
                            acx_entry *rent = w.ring + e;                 // peer-mapped: the slot at the receiver
                            uint64_t cur = ld_relaxed_sys(&rent->claim);
                            bool mine = false;
                            while ((cur >> 2) != g + 1) {
                                const uint64_t seen = atom_cas_acqrel_sys(&rent->claim, cur, ((g + 1) << 2) | 1ull);
                                if (seen == cur) { mine = true; break; }
                                cur = seen;
                            }
                            if (mine) {
                                const uint64_t cap = ld_relaxed_sys(&w.sack[e].r_cap), off = ld_relaxed_sys(&w.sack[e].r_dst_off);
                                op->peer_ptr = (uint64_t)translate_peer(w.peer_heap, w.xlat, off);
                                op->n = w.late_bytes < cap ? w.late_bytes : cap;
                                op->cta_count = 0;
                                st_release_sys(&op->late_state, tagv | 2ull);
                            } else {
                                st_release_sys(&op->late_state, tagv | 1ull);   // the receiver claimed it (fallback pull in progress)
                            }
                            spin_geq(&w.sack[e].sdone, g + 1, b.wd, 0x5e03);
                            break;
                        }
                        if (ns) __nanosleep(ns);
                        ns = ns < 256 ? ns * 2 + 20 : 256;
                        if (b.wd.budget_ns && globaltimer_ns() - t0 > b.wd.budget_ns) { b.wd.diag[1] = k; b.wd.diag[2] = 1; b.wd.diag[0] = 0x5e03; __threadfence_system(); __trap(); }
                    }
                }
            } else {
                spin_geq(&w.sack[e].sdone, g + 1, b.wd, 0x5e03);
            }
            s_n = 0; s_eager = 0;
        } else {
            acx_entry *ent = (w.any ? w.any->ring : w.ring) + e;
            // arrival: either an eager deposit (state.hi reached g+1 with kind EAGER) or mover done
            unsigned ns = 0;
            uint64_t t0 = b.wd.budget_ns ? globaltimer_ns() : 0;
            int eager = 0;
            uint64_t t_seen = 0;          // RNDV_LATE fallback: when the unclaimed message was first seen
            bool late_decided = false;
            for (;;) {
                if (ld_acquire_sys(&ent->done) >= g + 1) break;
                const bool here = (uint32_t)(ld_acquire_sys(&ent->state) >> 32) >= (uint32_t)(g + 1);
                if (here && ACX_ENT_KIND(ent->s_kind) == ACX_KIND_EAGER) { eager = 1; break; }
                if (w.late && here && !late_decided && ACX_ENT_KIND(ent->s_kind) == ACX_KIND_RNDV_LATE) {
                    // The sender moves LATE messages — from the mover behind its trigger, or from ITS wait kernel. If that wait is
                    // not running (the sender waits for something else first: MPI lets it), nobody would ever move the message:
                    // after a grace period this wait claims it and pulls it with its own helper CTAs (progress, at a price).
                    const uint64_t now = globaltimer_ns();
                    if (!t_seen) t_seen = now;
                    else if (now - t_seen > b.fallback_ns) {
                        uint64_t cur = ld_relaxed_sys(&ent->claim);
                        bool mine = false;
                        while ((cur >> 2) != g + 1) {
                            const uint64_t seen = atom_cas_acqrel_sys(&ent->claim, cur, ((g + 1) << 2) | 2ull);
                            if (seen == cur) { mine = true; break; }
                            cur = seen;
                        }
                        if (mine) {
                            const uint64_t sb = ent->s_bytes;
                            op->peer_ptr = (uint64_t)translate_peer(w.peer_heap, w.xlat, ent->s_src_off);
                            op->n = sb < w.cap ? sb : w.cap;
                            op->cta_count = 0;
                            st_release_sys(&op->late_state, late_tagv(w, k) | 2ull);
                        }
                        late_decided = true;   // mine: the helpers pull and publish `done`; theirs: the sender is on it
                    }
                }
                if (ns) __nanosleep(ns);
                ns = ns < 256 ? ns * 2 + 20 : 256;
                if (b.wd.budget_ns && globaltimer_ns() - t0 > b.wd.budget_ns) { b.wd.diag[1] = k; b.wd.diag[2] = ent->state; b.wd.diag[0] = 0x5e04; __threadfence_system(); __trap(); }
            }
            if (!eager && (uint32_t)(ld_acquire_sys(&ent->state) >> 32) >= (uint32_t)(g + 1) && ACX_ENT_KIND(ent->s_kind) == ACX_KIND_EAGER) eager = 1;   // left the loop through `done`: eager deposits set it too
            if (w.late && (ld_relaxed_sys(&op->late_state) >> 2) != (late_tagv(w, k) >> 2)) st_release_sys(&op->late_state, late_tagv(w, k) | 1ull);   // let the helper CTAs go: nothing to pull here
            uint64_t sb = ent->s_bytes;
            s_n = sb < w.cap ? sb : w.cap;
            if (!eager && w.cap <= w.eager_max) s_n = 0;   // rendezvous message into an eager-size receive: truncated, nothing was moved
            s_eager = eager;
        }
    }
    __syncthreads();
    if (w.is_recv) {
        const uint64_t k = op->ticket;
        const uint32_t e = (uint32_t)(k % w.depth);
        const uint64_t g = k / w.depth;
        const uint64_t n = s_n;
        acx_entry *const ring = w.any ? w.any->ring : w.ring;
        acx_sack *const sack = w.any ? w.any->sack : w.sack;
        const char *const eager = w.any ? w.any->eager : w.eager;
        const int32_t src = w.any ? w.any->src : w.src;
        if (s_eager) copy_bytes(w.dst, eager + (uint64_t)e * w.eager_max, n, tid, blockDim.x);
        else if (w.bounce && !w.bounce_kernel) copy_bytes(w.dst, w.bounce, n, tid, blockDim.x);
        __syncthreads();
        if (tid == 0) {
            acx_entry *ent = ring + e;
            const uint64_t sb = ent->s_bytes;
            const int32_t stag = (int32_t)ent->s_tag;
            // Channels are exact per (communicator, tag) (chan_lookup in acx.cu), so the header can only disagree on the shared
            // overflow channel (or with the directory disabled) when receives were posted in another order than the sends.
            // That is a mis-delivery, whether or not anybody looks at the status: diagnose it and stop.
            const bool mismatch = (w.tag != MPI_ANY_TAG && stag != w.tag) || ACX_ENT_CTX(ent->s_kind) != w.ctx;
            if (mismatch && b.wd.diag) {
                b.wd.diag[1] = ((uint64_t)(uint32_t)w.tag << 32) | (uint32_t)stag; b.wd.diag[2] = ((uint64_t)w.ctx << 32) | ACX_ENT_CTX(ent->s_kind); b.wd.diag[0] = 0x5e0d;
                __threadfence_system();
                __trap();
            }
            const int32_t err = (sb > w.cap) ? MPI_ERR_TRUNCATE : (mismatch ? MPI_ERR_TAG : MPI_SUCCESS);
            op->recv_bytes = n;
            op->bounce_n = (w.bounce && w.bounce_kernel && !s_eager) ? n : 0;
            if (w.ustatus) {
                if (w.off_source != 0xffff) *(int32_t *)(w.ustatus + w.off_source) = src;
                if (w.off_tag != 0xffff) *(int32_t *)(w.ustatus + w.off_tag) = stag;
                if (w.off_error != 0xffff) *(int32_t *)(w.ustatus + w.off_error) = err;
                if (w.off_bytes != 0xffff) *(int64_t *)(w.ustatus + w.off_bytes) = (int64_t)n;
            }
            if (w.serial) { w.hstat->source = src; w.hstat->tag = stag; w.hstat->error = err; w.hstat->bytes = (int64_t)n; }
            if (w.coop) { st_relaxed_sys(&ent->coop_cursor, (g + 1) << 32); st_relaxed_sys(&ent->coop_done, 0); __threadfence_system(); }   // atomic stores: a straggling agent of generation g may still be reading the cursor   // re-arm the chunk cursor for generation g+1 before the slot is handed on
            // All reads of the slot are complete (data dependence + barrier): plain stores suffice.
            st_relaxed_sys(&ent->consumed, g + 1);
            st_relaxed_sys(&sack[e].efree, g + 1);  // remote: slot may be reused by generation g+1
        }
    }
    if (tid == 0) {
        if (w.late && w.help_n) {
            // The helper CTAs of this request read the op record (ticket, decision word, source / destination) — and a helper may
            // be scheduled late. Once this CTA says "retired" the host hands the slot to the next request, whose trigger (with
            // several application threads: on another stream) rewrites the record: a helper that only starts then never sees its
            // decision and spins for good (ThreadSanitizer on the software device: tests/c/threads.c with large messages). So the
            // slot is not retired before every helper has checked out.
            unsigned ns = 0;
            const uint64_t t0 = b.wd.budget_ns ? globaltimer_ns() : 0;
            while (ld_acquire_sys(&op->helpers_out) < (uint64_t)w.help_n) {
                if (ns) __nanosleep(ns);
                ns = ns < 256 ? ns * 2 + 20 : 256;
                if (b.wd.budget_ns && globaltimer_ns() - t0 > b.wd.budget_ns) wd_fail(b.wd, 0x5e14, (uint64_t)w.help_n, ld_relaxed_sys(&op->helpers_out));
            }
            st_relaxed_sys(&op->helpers_out, 0);   // the next wait kernel on this record (next owner, or a relaunch of this graph) counts from zero
        }
        if (w.retire) ACX_HOST_FLAG_SET(&w.hstat->retired, 1);
        trace(b.tr, ACX_EV_WAIT_END, w.src, op->ticket, s_n);
    }
}
// A helper CTA's last access to the op record (thread 0, after the CTA's other threads are done with it).
__device__ __forceinline__ void helper_checkout(acx_op *op) {
    __threadfence();
    atomicAdd((unsigned long long *)&op->helpers_out, 1ull);
}
// RNDV_LATE helper CTAs of a wait kernel: they idle (on a LOCAL word) until the request's waiter CTA has decided, then
// either leave or move the payload — a send's helpers push (STG over NVLink), a receive's fallback helpers pull. Static
// slices, no helper waits for another, the last one out publishes completion to both ends (as the movers do).
__device__ __forceinline__ void late_helper(const WaitBatch &b, unsigned h) {
    int j = -1;
#pragma unroll 1
    for (int i = 0; i < (int)b.nreq && i < kWaitBatch; i++)
        if (b.w[i].late && h >= b.w[i].help_lo && h < (unsigned)b.w[i].help_lo + b.w[i].help_n) { j = i; break; }
    if (j < 0) return;
    const WaitOne &w = b.w[j];
    acx_op *op = w.op;
    __shared__ uint64_t s_state, s_k;
    if (threadIdx.x == 0) {
        unsigned ns = 0;
        const uint64_t t0 = b.wd.budget_ns ? globaltimer_ns() : 0;
        if (w.serial) while (ld_acquire_sys(&op->armed) != w.serial) { __nanosleep(ns); ns = ns < 256 ? ns * 2 + 20 : 256; }   // host-side wait: the trigger may not have run yet
        const uint64_t k = op->ticket;
        uint64_t st;
        ns = 0;
        while (((st = ld_acquire_sys(&op->late_state)) >> 2) != (late_tagv(w, k) >> 2)) {
            __nanosleep(ns); ns = ns < 512 ? ns * 2 + 40 : 512;
            if (b.wd.budget_ns && globaltimer_ns() - t0 > b.wd.budget_ns) { b.wd.diag[1] = k; b.wd.diag[2] = st; b.wd.diag[0] = 0x5e0e; __threadfence_system(); __trap(); }
        }
        s_state = st; s_k = k;
    }
    __syncthreads();
    if ((s_state & 3ull) != 2ull) { if (threadIdx.x == 0) helper_checkout(op); return; }
    const uint64_t n = op->n, k = s_k;
    char *peer = (char *)op->peer_ptr;
    const unsigned rank = h - w.help_lo, nh = w.help_n;
    if (rank == 0 && threadIdx.x == 0) trace(b.tr, ACX_EV_MOVE_BEGIN, (uint32_t)w.src, k, n);
    const uint64_t per = ((n + nh - 1) / nh + 127) & ~127ull, lo = per * rank;
    if (lo < n) {
        const uint64_t len = n - lo < per ? n - lo : per;
        if (w.is_recv) copy_bytes(w.late_local + lo, peer + lo, len, threadIdx.x, blockDim.x);
        else copy_bytes(peer + lo, w.late_local + lo, len, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(&op->cta_count, 1u) == nh - 1) {
            const uint32_t e = (uint32_t)(k % w.depth);
            const uint64_t g = k / w.depth;
            op->cta_count = 0;
            __threadfence_system();
            // recv: w.ring is my ring, w.sack the sender's acks (peer-mapped); send: w.ring is the receiver's ring (peer-mapped), w.sack my acks
            st_release_sys(&w.ring[e].done, g + 1);
            st_release_sys(&w.sack[e].sdone, g + 1);
            trace(b.tr, ACX_EV_MOVE_END, (uint32_t)w.src, k, n);
        }
        helper_checkout(op);   // after the barrier above: every thread of this CTA has read what it needed from the record
    }
}
__global__ void __launch_bounds__(kWaitThreads) wait_kernel(WaitBatch b) {
    // With helper CTAs on board this kernel must NOT release its programmatic dependents early: the next step's kernels —
    // among them the next wait kernel with its few hundred helper CTAs — would become resident at once and sit in
    // griddepcontrol.wait holding registers and thread slots while the mover this kernel waits for starves (measured on
    // B200: the 256-MiB self-ring slowed down in proportion to the helper count and hung at 296; fine with ACX_PDL=0).
    // ACX_WAIT_EARLY_HELPERS lets a kernel with only a few (receive-fallback) helpers release early all the same; measured on
    // the 256-MiB ring step it buys nothing (0.3897 vs 0.3899-0.3903 ms), so the limit is 0 by default.
    if (gridDim.x > b.nreq + b.early_release_max) pdl_wait_only(); else pdl_prologue();
    if (blockIdx.x < b.nreq) wait_body(b, b.w[blockIdx.x]);
    else late_helper(b, blockIdx.x - b.nreq);
}

// MPIX_Recv_enqueue for eager-size receives: post + complete in ONE launch (trigger, then the wait's body). The small-
// message exchange is bound by the host's launch rate (three launches per ping-pong iteration and rank); this makes it two.
__global__ void __launch_bounds__(kWaitThreads) recv_wait_fused(RecvParams rp, WaitBatch b) {
    pdl_prologue();
    if (threadIdx.x == 0) recv_match(rp);
    __syncthreads();
    wait_body(b, b.w[0]);
}
// MPIX_Sendrecv_enqueue for eager sizes on both halves: the whole exchange step — post the receive, deposit the send,
// complete the receive — in ONE launch (a ring or halo step with small messages costs three launches otherwise).
__global__ void __launch_bounds__(kWaitThreads) sendrecv_fused(SendParams sp, RecvParams rp, WaitBatch b) {
    pdl_prologue();
    if (threadIdx.x == 0) recv_match(rp);
    __syncthreads();
    send_body(sp);          // eager: never waits for the peer beyond the slot credit
    __syncthreads();
    wait_body(b, b.w[0]);
}

// Copy-out of a bounced (private-destination) rendezvous receive at full width; the last CTA
// retires the slot. Launched right after the wait kernel when the destination is not in the heap.
struct BounceOutParams {
    acx_op *op;
    char *dst;
    const char *bounce;
    acx_host_status *hstat;
    uint32_t retire, pad;
};
__global__ void __launch_bounds__(kMoverThreads) bounce_out_kernel(BounceOutParams p) {
    pdl_prologue();
    const uint64_t n = p.op->bounce_n;
    if (n) copy_bytes(p.dst, p.bounce, n, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    if (!p.retire) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicAdd(&p.op->cta_count, 1u);
        if (t == gridDim.x - 1) { p.op->cta_count = 0; ACX_HOST_FLAG_SET(&p.hstat->retired, 1); }
    }
}

// ---------------------------------------------------------------------------------------
// Non-contiguous datatypes (vector / subarray layouts, e.g. halo faces): the wire format is the packed image.
// Send: pack_kernel gathers the user's strided layout into a heap staging buffer, which then travels like any
// contiguous heap buffer. Receive: the message lands (eager copy-out or rendezvous mover) in a heap staging buffer and
// unpack_kernel, launched behind the wait kernel, scatters it; its last CTA retires the slot.
// Element (c, i2, i1, i0) of the layout is `run` contiguous bytes at  c*extent + disp + i2*s2 + i1*s1 + i0*s0.
// Work is split in `vec`-byte units (the widest power of two <= 16 that divides run, strides and both base addresses).
// ---------------------------------------------------------------------------------------
template <bool kPack>
__device__ __forceinline__ void pack_unpack_body(char *packed, char *user, const TypeLayout &L, uint64_t limit) {
    const uint64_t units = acx_layout_units(L, limit);                  // short message: only what arrived
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < units; j += (uint64_t)gridDim.x * blockDim.x) {
        char *u = user + acx_layout_unit_offset(L, j);
        char *p = packed + j * L.vec;
        char *d = kPack ? p : u;
        const char *s = kPack ? u : p;
        switch (L.vec) {
            case 16: *(int4 *)d = *(const int4 *)s; break;
            case 8: *(int2 *)d = *(const int2 *)s; break;
            case 4: *(int *)d = *(const int *)s; break;
            case 2: *(short *)d = *(const short *)s; break;
            default: *d = *s; break;
        }
    }
}
struct PackParams {
    char *packed;        // heap staging buffer
    const char *user;    // device-accessible alias of the user's buffer
    TypeLayout L;
};
__global__ void __launch_bounds__(256) pack_kernel(PackParams p) {
    pdl_prologue();
    pack_unpack_body<true>(p.packed, (char *)p.user, p.L, ~0ull);
}
struct UnpackParams {
    acx_op *op;          // recv_bytes: how much actually arrived
    char *user;
    const char *packed;
    acx_host_status *hstat;
    TypeLayout L;
    uint32_t retire, pad;
};
__global__ void __launch_bounds__(256) unpack_kernel(UnpackParams p) {
    pdl_prologue();
    pack_unpack_body<false>((char *)p.packed, p.user, p.L, p.op->recv_bytes);
    if (!p.retire) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicAdd(&p.op->cta_count, 1u);
        if (t == gridDim.x - 1) { p.op->cta_count = 0; ACX_HOST_FLAG_SET(&p.hstat->retired, 1); }
    }
}

// ---------------------------------------------------------------------------------------
// Partitioned helpers (host Start / Wait paths; the device-side API is in acx_device.cuh)
// ---------------------------------------------------------------------------------------
__global__ void pstart_kernel(uint64_t *epoch_local, uint64_t *cts_remote, uint64_t e) {
    st_release_sys(epoch_local, e);
    if (cts_remote) { __threadfence_system(); st_release_sys(cts_remote, e); }
}
// Batched Start (ACX_PBATCH=1): one launch starts up to kPBatch requests, one thread each.
constexpr int kPBatch = 16;
struct PStartBatch {
    uint64_t *epoch_local[kPBatch];
    uint64_t *cts_remote[kPBatch];
    uint64_t e[kPBatch];
    int n, pad;
};
__global__ void pstart_batch_kernel(PStartBatch b) {
    const int i = (int)threadIdx.x;
    if (i >= b.n) return;
    st_release_sys(b.epoch_local[i], b.e[i]);
    if (b.cts_remote[i]) { __threadfence_system(); st_release_sys(b.cts_remote[i], b.e[i]); }
}

// Completes one round: SEND waits until every partition was marked ready (sent[p]==e);
// RECV waits for every arrival and copies out what device pollers have not claimed.
__device__ __forceinline__ void pwait_body(const acx_pchan *ch, uint64_t e, const Watchdog &wd) {
    __shared__ int s_copy;
    for (int p = 0; p < ch->partitions; p++) {
        if (ch->kind == ACX_PKIND_SEND) {
            if (threadIdx.x == 0) spin_geq(ch->local_flags + p, e, wd, 0x5e06);
            continue;
        }
        if (threadIdx.x == 0) {
            spin_geq(ch->arrived + p, e, wd, 0x5e07);
            int do_copy = 0;
            if (ch->land != ch->user_buf) {
                const uint64_t done = 2 * e, claimed = 2 * e - 1;
                uint64_t c = ld_acquire_sys(ch->local_flags + p);
                if (c != done && c != claimed && atom_cas_acqrel_sys(ch->local_flags + p, c, claimed) == c) do_copy = 1;
                else spin_geq(ch->local_flags + p, done, wd, 0x5e08);  // a device poller owns the copy
            }
            s_copy = do_copy;
        }
        __syncthreads();
        if (s_copy) {
            uint64_t off = (uint64_t)p * ch->part_bytes;
            copy_bytes(ch->user_buf + off, ch->land + off, ch->part_bytes, threadIdx.x, blockDim.x);
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0) st_release_sys(ch->local_flags + p, 2 * e);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kWaitThreads) pwait_kernel(const acx_pchan *ch, uint64_t e, Watchdog wd) { pwait_body(ch, e, wd); }
// Batched completion (ACX_PBATCH=1): one CTA per request, one launch for a whole MPIX_Waitall.
struct PWaitBatch {
    const acx_pchan *ch[kPBatch];
    uint64_t e[kPBatch];
    Watchdog wd;
};
__global__ void __launch_bounds__(kWaitThreads) pwait_batch_kernel(PWaitBatch b) { pwait_body(b.ch[blockIdx.x], b.e[blockIdx.x], b.wd); }

// Host-callable single-partition operations (MPIX_Pready_host / MPIX_Parrived_host).
__global__ void pready_host_kernel(const acx_pchan *ch, int p) { pready_cta(p, ch); }
__global__ void parrived_host_kernel(const acx_pchan *ch, int p, int *flag) { parrived(ch, p, flag); }

}  // namespace acxk
#endif
