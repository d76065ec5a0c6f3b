// Protocol emulation: the device kernels of src/core/kernels.cuh run on CPU threads (tests/emu/cuda_emu.h) for a few
// "ranks" inside one process. Exercises, with real concurrency and randomized arrival orders, the paths that were
// written after the round's GPU budget was spent — cooperative moving (chunk stealing), the wildcard resolver, the
// pack/unpack kernels — next to the hardware-validated eager / rendezvous paths, which double as a check of the
// emulation itself. Plain g++ build, no GPU, no libacx.so (everything is header-inline).
#include "cuda_emu.h"
#include "../../src/core/kernels.cuh"

#include <atomic>
#include <mutex>
#include <random>

AcxState *g_acx = nullptr;
using namespace acxk;

static int g_errs = 0;
#define CHECK(c, ...) do { if (!(c)) { fprintf(stderr, "%s:%d CHECK(%s): ", __FILE__, __LINE__, #c); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); __atomic_fetch_add(&g_errs, 1, __ATOMIC_RELAXED); } } while (0)

struct Rank {
    char *ctrl = nullptr, *heap = nullptr;
    std::vector<acx_host_status> hstat;
    std::atomic<uint64_t> serial{1};   // two receive streams of one rank share it in run_wildcard_vs_specific
    uint64_t diag[4] = {0, 0, 0, 0};
};
static std::vector<Rank> R;
static acx_geom G;
static const size_t kHeap = 8u << 20;

static void world_init(int n) {
    static AcxState st;
    g_acx = &st;
    memset(&G, 0, sizeof G);
    G.nranks = n; G.nchan = 2; G.depth = 2; G.nops = 64; G.eager_max = 64; G.npflags = 64; G.ntrace = 2; G.maxreg = 1;
    acx_geom_layout(G);
    st.geom = G; st.size = n;
    R = std::vector<Rank>(n);
    for (auto &r : R) {
        r.ctrl = (char *)aligned_alloc(256, G.bytes); memset(r.ctrl, 0, G.bytes);
        r.heap = (char *)aligned_alloc(256, kHeap); memset(r.heap, 0, kHeap);
        r.hstat.assign(G.nops, acx_host_status());
    }
}
static void world_free() { for (auto &r : R) { free(r.ctrl); free(r.heap); } R.clear(); }
static Watchdog wd_of(int me) { Watchdog w; w.budget_ns = 20ull * 1000000000ull; w.diag = R[me].diag; return w; }
static acx_op *op_of(int me, int slot) { return (acx_op *)(R[me].ctrl + G.off_ops) + slot; }
static TraceCtx no_trace() { TraceCtx t{nullptr, 0, nullptr}; return t; }

// ---- parameter builders: the same arithmetic as MPIX_Isend_enqueue / MPIX_Irecv_enqueue / fill_wait in src/core/acx.cu ----
static void build_send(int me, int dst, int chan, uint64_t src_off, uint64_t bytes, int tag, int slot, bool coop, SendParams *sp, MoverParams *mp, uint64_t *serial) {
    const uint64_t ring_idx = acx_entry_index(G, me, chan, 0), sack_idx = acx_entry_index(G, dst, chan, 0);
    memset(sp, 0, sizeof *sp); memset(mp, 0, sizeof *mp);
    sp->ring = acx_entries_of(R[dst].ctrl) + ring_idx;
    sp->eager = R[dst].ctrl + G.off_eager + ring_idx * G.eager_max;
    sp->sack = acx_sacks_of(R[me].ctrl) + sack_idx;
    sp->seq = (uint64_t *)(R[me].ctrl + G.off_send_seq) + ((uint64_t)dst * G.nchan + chan);
    sp->op = op_of(me, slot);
    sp->hstat = bytes <= G.eager_max ? &R[me].hstat[slot] : nullptr;
    sp->src = R[me].heap + src_off; sp->peer_heap = R[dst].heap;
    sp->xlat = (const uint64_t *)(R[me].ctrl + G.off_xlat) + (uint64_t)dst * G.maxreg;
    sp->src_off = src_off; sp->bytes = bytes; sp->serial = *serial = R[me].serial++;
    sp->tag = (uint32_t)tag; sp->depth = G.depth; sp->eager_max = G.eager_max; sp->peer = dst; sp->coop = coop ? 1 : 0;
    sp->wd = wd_of(me); sp->tr = no_trace();
    mp->op = sp->op; mp->local = R[me].heap + src_off; mp->ring = sp->ring; mp->sack = sp->sack; mp->depth = G.depth; mp->push = 1; mp->peer = dst; mp->tr = no_trace();
}
static void build_recv(int me, int src, int chan, uint64_t dst_off, uint64_t cap, int slot, RecvParams *rp, MoverParams *mp, uint64_t *serial) {
    const uint64_t ring_idx = acx_entry_index(G, src, chan, 0), sack_idx = acx_entry_index(G, me, chan, 0);
    const bool rndv = cap > G.eager_max;
    memset(rp, 0, sizeof *rp); memset(mp, 0, sizeof *mp);
    rp->ring = acx_entries_of(R[me].ctrl) + ring_idx;
    rp->seq = (uint64_t *)(R[me].ctrl + G.off_recv_seq) + ((uint64_t)src * G.nchan + chan);
    rp->op = op_of(me, slot);
    rp->sack = acx_sacks_of(R[src].ctrl) + sack_idx; rp->rndv_capable = rndv ? 1 : 0; rp->publish = rp->rndv_capable;
    rp->peer_heap = R[src].heap;
    rp->xlat = (const uint64_t *)(R[me].ctrl + G.off_xlat) + (uint64_t)src * G.maxreg;
    rp->dst_off = rndv ? dst_off : ~0ull; rp->cap = cap; rp->serial = *serial = R[me].serial++; rp->depth = G.depth; rp->peer = src;
    rp->wd = wd_of(me); rp->tr = no_trace();
    mp->op = rp->op; mp->local = R[me].heap + dst_off; mp->ring = rp->ring; mp->sack = acx_sacks_of(R[src].ctrl) + sack_idx;
    mp->depth = G.depth; mp->push = 0; mp->peer = src; mp->tr = no_trace();
}
static void build_wait(int me, bool is_recv, int peer, int chan, uint64_t dst_off, uint64_t cap, int tag, int slot, bool coop, const acx_anyres *any, WaitOne *w) {
    memset(w, 0, sizeof *w);
    w->hticket = ~0ull;   // no host-side ticket prediction here
    w->op = op_of(me, slot); w->hstat = &R[me].hstat[slot];
    w->depth = G.depth; w->eager_max = G.eager_max; w->retire = 1; w->src = peer; w->tag = tag;
    w->off_source = w->off_tag = w->off_error = w->off_bytes = 0xffff;
    if (!is_recv) { w->sack = acx_sacks_of(R[me].ctrl) + acx_entry_index(G, peer, chan, 0); return; }
    w->is_recv = 1;
    if (any) w->any = any;
    else {
        const uint64_t ring_idx = acx_entry_index(G, peer, chan, 0);
        w->ring = acx_entries_of(R[me].ctrl) + ring_idx;
        w->eager = R[me].ctrl + G.off_eager + ring_idx * G.eager_max;
        w->sack = acx_sacks_of(R[peer].ctrl) + acx_entry_index(G, me, chan, 0);
    }
    w->dst = R[me].heap + dst_off; w->cap = cap; w->coop = coop ? 1 : 0; w->hticket = ~0ull;
}
static void build_coop(int me, bool is_send, int peer, int chan, uint64_t off, uint64_t bytes, int slot, CoopParams *cp) {
    memset(cp, 0, sizeof *cp);
    cp->op = op_of(me, slot); cp->is_send = is_send ? 1 : 0;
    if (is_send) { cp->ring = acx_entries_of(R[peer].ctrl) + acx_entry_index(G, me, chan, 0); cp->sack = acx_sacks_of(R[me].ctrl) + acx_entry_index(G, peer, chan, 0); }
    else { cp->ring = acx_entries_of(R[me].ctrl) + acx_entry_index(G, peer, chan, 0); cp->sack = acx_sacks_of(R[peer].ctrl) + acx_entry_index(G, me, chan, 0); }
    cp->local = R[me].heap + off; cp->peer_heap = R[peer].heap;
    cp->xlat = (const uint64_t *)(R[me].ctrl + G.off_xlat) + (uint64_t)peer * G.maxreg;
    cp->bytes = bytes; cp->depth = G.depth; cp->wd = wd_of(me);
}

static const unsigned kBlk = 8;   // "CUDA threads" per CTA in the emulation
static void fill(char *p, size_t n, uint32_t seed) { for (size_t i = 0; i < n; i++) p[i] = (char)((i * 2654435761u + seed * 40503u) >> 7); }
static size_t mismatches(const char *p, size_t n, uint32_t seed) { size_t bad = 0; for (size_t i = 0; i < n; i++) if (p[i] != (char)((i * 2654435761u + seed * 40503u) >> 7)) bad++; return bad; }
static void jitter(std::mt19937 &rng, int max_us) { if (max_us > 0) { timespec t{0, (long)(rng() % (unsigned)max_us) * 1000}; nanosleep(&t, nullptr); } }

// ---- scenario 1: ring exchange, `iters` rounds, random arrival order; eager / classic rendezvous / cooperative -------------
static void ring_rank(int me, int n, size_t bytes, int iters, bool coop, int mover_ctas, int jitter_us, unsigned seed) {
    std::mt19937 rng(seed + 977 * me);
    const int left = (me + n - 1) % n, right = (me + 1) % n, chan = 1, tag = 5;
    const uint64_t s_off = 0, r_off = kHeap / 2;
    const bool rndv = bytes > G.eager_max;
    for (int it = 0; it < iters; it++) {
        fill(R[me].heap + s_off, bytes, 1000u * me + it);
        memset(R[me].heap + r_off, 0xAB, bytes);
        const int s_slot = (2 * it) % G.nops, r_slot = (2 * it + 1) % G.nops;
        SendParams sp; RecvParams rp; MoverParams smp, rmp; uint64_t sser, rser;
        build_send(me, right, chan, s_off, bytes, tag, s_slot, coop && rndv, &sp, &smp, &sser);
        build_recv(me, left, chan, r_off, bytes, r_slot, &rp, &rmp, &rser);
        auto do_recv = [&] {
            jitter(rng, jitter_us);
            emu::launch(1, 1, [=] { recv_trigger(rp); });
            if (rndv) emu::launch(mover_ctas, kBlk, [=] { mover(rmp); });
        };
        auto do_send = [&] {
            jitter(rng, jitter_us);
            emu::launch(1, kBlk, [=] { send_trigger(sp); });
            if (rndv) emu::launch(mover_ctas, kBlk, [=] { mover(smp); });
        };
        if (rng() & 1) { do_recv(); do_send(); } else { do_send(); do_recv(); }
        jitter(rng, jitter_us);
        if (coop && rndv) {   // what MPIX_Waitall_enqueue emits in front of the wait kernel
            CoopParams cr, cs;
            build_coop(me, false, left, chan, r_off, bytes, r_slot, &cr);
            build_coop(me, true, right, chan, s_off, bytes, s_slot, &cs);
            emu::launch(1, 1, [=] { coop_gate(cr); });
            emu::launch(mover_ctas, kBlk, [=] { coop_kernel(cr); });
            emu::launch(1, 1, [=] { coop_gate(cs); });
            emu::launch(mover_ctas, kBlk, [=] { coop_kernel(cs); });
        }
        WaitBatch wb;
        memset(&wb, 0, sizeof wb);
        build_wait(me, true, left, chan, r_off, bytes, tag, r_slot, coop, nullptr, &wb.w[0]);
        int nw = 1;
        if (rndv) { build_wait(me, false, right, chan, 0, 0, tag, s_slot, coop, nullptr, &wb.w[1]); nw = 2; }
        wb.wd = wd_of(me); wb.tr = no_trace();
        wb.nreq = (uint32_t)nw;
        emu::launch(nw, kBlk, [=] { wait_kernel(wb); });
        const size_t bad = mismatches(R[me].heap + r_off, bytes, 1000u * left + it);
        CHECK(bad == 0, "ring n=%d bytes=%zu coop=%d iter %d rank %d: %zu wrong bytes", n, bytes, (int)coop, it, me, bad);
        CHECK(op_of(me, r_slot)->recv_bytes == bytes, "recv_bytes %llu", (unsigned long long)op_of(me, r_slot)->recv_bytes);
        if (bad) return;
    }
}
static void run_ring(int n, size_t bytes, int iters, bool coop, int mover_ctas, int jitter_us, unsigned seed) {
    world_init(n);
    std::vector<std::thread> ts;
    for (int r = 0; r < n; r++) ts.emplace_back(ring_rank, r, n, bytes, iters, coop, mover_ctas, jitter_us, seed);
    for (auto &t : ts) t.join();
    world_free();
    printf("ring n=%d bytes=%zu iters=%d coop=%d ctas=%d: %s\n", n, bytes, iters, (int)coop, mover_ctas, g_errs ? "FAIL" : "ok");
}

// ---- scenario 2: wildcard receives (recv_any_kernel + wait with the resolution record) ------------------------------------
static AnyTable g_tbl;
static std::vector<acx_anyres> g_anyres;   // rank 0's per-slot resolution records
static void build_any(int me, int src, int tag, uint64_t dst_off, uint64_t cap, int slot, AnyParams *ap) {
    memset(ap, 0, sizeof *ap);
    ap->tbl = &g_tbl; ap->ctrl = R[me].ctrl; ap->res = &g_anyres[slot]; ap->op = op_of(me, slot);
    ap->rndv_capable = cap > G.eager_max ? 1 : 0;
    ap->dst = R[me].heap + dst_off; ap->dst_off = cap > G.eager_max ? dst_off : ~0ull; ap->cap = cap; ap->serial = R[me].serial++; ap->geom = G;
    if (src == MPI_ANY_SOURCE) { ap->src_lo = 0; ap->src_hi = G.nranks - 1; } else ap->src_lo = ap->src_hi = src;
    if (tag == MPI_ANY_TAG) { ap->chan_lo = 0; ap->chan_hi = G.nchan - 1; } else ap->chan_lo = ap->chan_hi = tag % G.nchan;
    ap->want_tag = tag; ap->me = (uint32_t)me; ap->wd = wd_of(me); ap->tr = no_trace();
}
struct Msg { uint32_t id, src, tag, bytes; };
static std::atomic<int> g_slot_rr{0};
static bool g_wild_coop = false;   // receiver-side wait kernels re-arm the chunk cursor (job-wide ACX_COOP)
// one wildcard (or specific, when src/tag are given) receive on rank 0, completed by a wait; returns the header found in the payload
static Msg recv_one(int me, int src, int tag, uint64_t dst_off, uint64_t cap, bool wildcard_kernel, std::mt19937 &rng, int jitter_us) {
    const int slot = g_slot_rr.fetch_add(1) % G.nops;
    jitter(rng, jitter_us);
    WaitBatch wb;
    memset(&wb, 0, sizeof wb);
    if (wildcard_kernel) {
        AnyParams ap;
        build_any(me, src, tag, dst_off, cap, slot, &ap);
        emu::launch(cap > G.eager_max ? 2 : 1, kBlk, [=] { recv_any_kernel(ap); });
        build_wait(me, true, src, 0, dst_off, cap, tag, slot, g_wild_coop, &g_anyres[slot], &wb.w[0]);
    } else {
        RecvParams rp; MoverParams mp; uint64_t ser;
        build_recv(me, src, tag % G.nchan, dst_off, cap, slot, &rp, &mp, &ser);
        emu::launch(1, 1, [=] { recv_trigger(rp); });
        if (cap > G.eager_max) emu::launch(2, kBlk, [=] { mover(mp); });
        build_wait(me, true, src, tag % G.nchan, dst_off, cap, tag, slot, false, nullptr, &wb.w[0]);
    }
    wb.w[0].serial = 0;
    wb.wd = wd_of(me); wb.tr = no_trace();
    R[me].hstat[slot] = acx_host_status();
    // read the status the way a host wait would: ask the kernel to fill hstat (serial != 0 makes it wait for `armed` first)
    wb.w[0].serial = op_of(me, slot)->armed;
    wb.nreq = 1;
    emu::launch(1, kBlk, [=] { wait_kernel(wb); });
    Msg m;
    memcpy(&m, R[me].heap + dst_off, sizeof m);
    const acx_host_status &h = R[me].hstat[slot];
    CHECK(h.source == (int)m.src && h.tag == (int)m.tag && h.bytes == (int64_t)m.bytes && h.error == MPI_SUCCESS,
          "status (%d,%d,%lld,err %d) vs payload header (src %u tag %u bytes %u)", h.source, h.tag, (long long)h.bytes, h.error, m.src, m.tag, m.bytes);
    const size_t bad = mismatches(R[me].heap + dst_off + sizeof m, m.bytes - sizeof m, m.id);
    CHECK(bad == 0, "wildcard payload id %u from %u: %zu wrong bytes", m.id, m.src, bad);
    return m;
}
static void sender_rank(int me, int count, int jitter_us, unsigned seed, bool one_tag, bool coop) {
    std::mt19937 rng(seed + 31 * me);
    for (int i = 0; i < count; i++) {
        const uint32_t bytes = (rng() & 1) ? 48 : 5000 + (rng() % 3000);   // eager or rendezvous
        const int tag = one_tag ? 4 : (int)(rng() % 6);
        Msg m{(uint32_t)(me * 100000 + i), (uint32_t)me, (uint32_t)tag, bytes};
        const uint64_t off = 0;
        memcpy(R[me].heap + off, &m, sizeof m);
        fill(R[me].heap + off + sizeof m, bytes - sizeof m, m.id);
        const int slot = i % G.nops;
        SendParams sp; MoverParams mp; uint64_t ser;
        build_send(me, 0, tag % G.nchan, off, bytes, tag, slot, coop && bytes > G.eager_max, &sp, &mp, &ser);
        jitter(rng, jitter_us);
        emu::launch(1, kBlk, [=] { send_trigger(sp); });
        if (bytes > G.eager_max) {
            emu::launch(2, kBlk, [=] { mover(mp); });
            if (coop) {
                CoopParams cs;
                build_coop(me, true, 0, tag % G.nchan, off, bytes, slot, &cs);
                emu::launch(1, 1, [=] { coop_gate(cs); });
                emu::launch(2, kBlk, [=] { coop_kernel(cs); });
            }
            WaitBatch wb;
            memset(&wb, 0, sizeof wb);
            build_wait(me, false, 0, tag % G.nchan, 0, 0, tag, slot, false, nullptr, &wb.w[0]);
            wb.wd = wd_of(me); wb.tr = no_trace();
            wb.nreq = 1;
    emu::launch(1, kBlk, [=] { wait_kernel(wb); });   // the source buffer is reused by the next message
        } else {
            // eager: the trigger has copied the payload out; nothing to wait for
        }
    }
}
static void run_wildcards(int n, int per_sender, int jitter_us, unsigned seed, bool coop) {
    world_init(n);
    memset(&g_tbl, 0, sizeof g_tbl);
    for (int r = 0; r < n; r++) { g_tbl.peer_ctrl[r] = R[r].ctrl; g_tbl.peer_heap[r] = R[r].heap; }
    g_anyres.assign(G.nops, acx_anyres());
    g_wild_coop = coop;
    std::vector<std::thread> ts;
    for (int r = 1; r < n; r++) ts.emplace_back(sender_rank, r, per_sender, jitter_us, seed, false, coop);
    std::vector<int> seen((size_t)n * 100000, 0);
    std::mt19937 rng(seed);
    std::vector<uint32_t> last(n, 0);
    for (int i = 0; i < (n - 1) * per_sender; i++) {
        Msg m = recv_one(0, MPI_ANY_SOURCE, MPI_ANY_TAG, kHeap / 2, 16384, true, rng, jitter_us);
        CHECK(m.src >= 1 && m.src < (uint32_t)n && m.id / 100000 == m.src, "bogus message id %u src %u", m.id, m.src);
        if (m.src < (uint32_t)n) seen[m.id]++;
    }
    for (auto &t : ts) t.join();
    for (int r = 1; r < n; r++) for (int i = 0; i < per_sender; i++) CHECK(seen[(size_t)r * 100000 + i] == 1, "message %d of rank %d received %d times", i, r, seen[(size_t)r * 100000 + i]);
    world_free();
    g_wild_coop = false;
    printf("wildcards n=%d per_sender=%d coop=%d: %s\n", n, per_sender, (int)coop, g_errs ? "FAIL" : "ok");
}
// a wildcard stream and a specific-receive stream compete for the messages of ONE (source, channel): every message is
// delivered exactly once, whichever receive flavour gets it
static void run_wildcard_vs_specific(int msgs, int jitter_us, unsigned seed) {
    world_init(2);
    memset(&g_tbl, 0, sizeof g_tbl);
    for (int r = 0; r < 2; r++) { g_tbl.peer_ctrl[r] = R[r].ctrl; g_tbl.peer_heap[r] = R[r].heap; }
    g_anyres.assign(G.nops, acx_anyres());
    std::thread snd(sender_rank, 1, 2 * msgs, jitter_us, seed, true, false);
    std::vector<int> seen(200000 + 2 * msgs, 0);
    std::mutex mu;
    auto stream = [&](bool wildcard, uint64_t off, unsigned sd) {
        std::mt19937 rng(sd);
        for (int i = 0; i < msgs; i++) {
            Msg m = recv_one(0, wildcard ? MPI_ANY_SOURCE : 1, 4, off, 16384, wildcard, rng, jitter_us);
            std::lock_guard<std::mutex> lk(mu);
            if (m.id >= 100000 && m.id < 100000 + 2u * msgs) seen[m.id]++; else CHECK(0, "bogus id %u", m.id);
        }
    };
    std::thread a(stream, true, kHeap / 2, seed + 1), b(stream, false, kHeap / 2 + 65536, seed + 2);
    a.join(); b.join(); snd.join();
    for (int i = 0; i < 2 * msgs; i++) CHECK(seen[100000 + i] == 1, "message %d delivered %d times", i, seen[100000 + i]);
    world_free();
    printf("wildcard vs specific on one channel, %d msgs each: %s\n", msgs, g_errs ? "FAIL" : "ok");
}

// ---- scenario 3: strided datatypes: pack_kernel -> (eager | rendezvous) -> wait -> unpack_kernel ---------------------------
static void typed_rank(int me, int n, int nx, int ny, int nz, int iters, unsigned seed) {
    std::mt19937 rng(seed + me);
    const int left = (me + n - 1) % n, right = (me + 1) % n, chan = 0, tag = 2;
    const size_t cells = (size_t)nx * ny * nz;
    int *a = (int *)(R[me].heap);                       // the 3-D block lives at heap offset 0 (any device-accessible memory would do)
    const uint64_t stage_s = kHeap / 4, stage_r = kHeap / 2;
    // x faces: send interior x = nx-2, receive into ghost x = 0
    AcxLayout ls, lr;
    memset(&ls, 0, sizeof ls);
    ls.n[0] = ny; ls.stride[0] = (long long)nx * 4; ls.n[1] = nz; ls.stride[1] = (long long)nx * ny * 4; ls.n[2] = 1;
    ls.run = 4; ls.extent = (long long)cells * 4; ls.count = 1; ls.noncontig = true;
    lr = ls;
    ls.disp = (long long)(nx - 2) * 4; lr.disp = 0;
    const uint64_t bytes = (uint64_t)ny * nz * 4;
    for (int it = 0; it < iters; it++) {
        for (size_t i = 0; i < cells; i++) a[i] = (int)(1000000 * (me + 1) + 1000 * it + i);
        for (int z = 0; z < nz; z++) for (int y = 0; y < ny; y++) a[((size_t)z * ny + y) * nx] = -1;
        const int s_slot = (2 * it) % G.nops, r_slot = (2 * it + 1) % G.nops;
        PackParams pp;
        memset(&pp, 0, sizeof pp);
        pp.packed = R[me].heap + stage_s; pp.user = (const char *)a; pp.L = acx_device_layout(ls, a, pp.packed);
        SendParams sp; RecvParams rp; MoverParams smp, rmp; uint64_t sser, rser;
        build_send(me, right, chan, stage_s, bytes, tag, s_slot, false, &sp, &smp, &sser);
        build_recv(me, left, chan, stage_r, bytes, r_slot, &rp, &rmp, &rser);
        const bool rndv = bytes > G.eager_max;
        jitter(rng, 100);
        emu::launch(1, 1, [=] { recv_trigger(rp); });
        if (rndv) emu::launch(2, kBlk, [=] { mover(rmp); });
        jitter(rng, 100);
        emu::launch(2, kBlk, [=] { pack_kernel(pp); });
        emu::launch(1, kBlk, [=] { send_trigger(sp); });
        if (rndv) emu::launch(2, kBlk, [=] { mover(smp); });
        WaitBatch wb;
        memset(&wb, 0, sizeof wb);
        build_wait(me, true, left, chan, stage_r, bytes, tag, r_slot, false, nullptr, &wb.w[0]);   // dst = the staging buffer, as fill_wait does for typed receives
        wb.w[0].retire = 0;
        int nw = 1;
        if (rndv) { build_wait(me, false, right, chan, 0, 0, tag, s_slot, false, nullptr, &wb.w[1]); nw = 2; }
        wb.wd = wd_of(me); wb.tr = no_trace();
        wb.nreq = (uint32_t)nw;
        emu::launch(nw, kBlk, [=] { wait_kernel(wb); });
        UnpackParams up;
        memset(&up, 0, sizeof up);
        up.op = op_of(me, r_slot); up.user = (char *)a; up.packed = R[me].heap + stage_r; up.hstat = &R[me].hstat[r_slot];
        up.L = acx_device_layout(lr, a, up.packed); up.retire = 1;
        R[me].hstat[r_slot].retired = 0;
        emu::launch(2, kBlk, [=] { unpack_kernel(up); });
        CHECK(R[me].hstat[r_slot].retired == 1, "unpack kernel did not retire the slot");
        size_t bad = 0;
        for (int z = 0; z < nz; z++) for (int y = 0; y < ny; y++) for (int x = 0; x < nx; x++) {
            const size_t i = ((size_t)z * ny + y) * nx + x;
            const int want = x == 0 ? (int)(1000000 * (left + 1) + 1000 * it + (i + nx - 2)) : (int)(1000000 * (me + 1) + 1000 * it + i);
            if (a[i] != want) bad++;
        }
        CHECK(bad == 0, "typed exchange %dx%dx%d iter %d rank %d: %zu wrong cells", nx, ny, nz, it, me, bad);
        if (bad) return;
    }
}
static void run_typed(int n, int nx, int ny, int nz, int iters, unsigned seed) {
    world_init(n);
    std::vector<std::thread> ts;
    for (int r = 0; r < n; r++) ts.emplace_back(typed_rank, r, n, nx, ny, nz, iters, seed);
    for (auto &t : ts) t.join();
    world_free();
    printf("typed faces %dx%dx%d n=%d: %s\n", nx, ny, nz, n, g_errs ? "FAIL" : "ok");
}

int main() {
    run_ring(2, 16, 40, false, 1, 50, 1);           // eager
    run_ring(3, 3000, 40, false, 3, 50, 2);         // classic rendezvous, depth-2 ring wraps 20 times
    run_ring(1, 3000, 10, false, 2, 0, 3);          // self
    run_ring(2, 300000, 30, true, 3, 200, 4);       // cooperative: 5 chunks of 64 KiB
    run_ring(3, 200000, 30, true, 2, 200, 5);
    run_ring(1, 200000, 8, true, 2, 0, 6);
    run_wildcards(3, 40, 100, 7, false);
    run_wildcards(4, 25, 0, 8, false);
    run_wildcards(3, 30, 60, 12, true);     // wildcard receives of messages announced as cooperative
    run_wildcard_vs_specific(40, 80, 9);
    run_typed(2, 6, 3, 4, 12, 10);      // 48-byte faces: eager
    run_typed(3, 8, 16, 12, 12, 11);    // 768-byte faces: rendezvous
    printf("protocol_emu: %s\n", g_errs ? "FAIL" : "PASS");
    return g_errs != 0;
}
