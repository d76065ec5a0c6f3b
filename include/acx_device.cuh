// acx-b200 device API: flag primitives (release/acquire at system scope across NVLink),
// the partitioned-channel device handle, and Pready / Parrived in single-thread,
// warp- and CTA-cooperative flavours.
//
// Reference behaviour being replaced (src/partitioned.cu:200-231): Pready = one
// volatile int store to a host-pinned flag (no fence, defect D5), a CPU proxy then
// calls MPI_Pready; Parrived = one volatile load of a host flag over PCIe.
// Here: Pready moves the partition's bytes into the peer GPU's HBM (or the producer
// already stored them there) and publishes a release flag in the peer's HBM; Parrived
// is an acquire load from local HBM. No host involvement.
#ifndef ACX_DEVICE_CUH
#define ACX_DEVICE_CUH

#include <stdint.h>

#if defined(__CUDACC__)
#define ACX_HD __host__ __device__
#define ACX_D __device__ __forceinline__
#elif defined(ACX_EMU)   // host-side protocol emulation (tests/emu): device code compiled as plain C++
#define ACX_HD
#define ACX_D inline
#else
#define ACX_HD
#endif

enum { ACX_PKIND_SEND = 1, ACX_PKIND_RECV = 2 };

// Device-resident descriptor of one partitioned request (persistent channel endpoint).
// MPIX_Prequest is a device pointer to this struct; kernels receive it by value.
struct acx_pchan {
    int32_t kind;            // ACX_PKIND_*
    int32_t partitions;
    uint64_t part_bytes;
    char *user_buf;          // local user buffer (device-accessible: heap / device / mapped host)
    char *land;              // SEND: peer-mapped landing buffer in the receiver's heap
                             // RECV: local landing buffer (== user_buf when the user buffer is in the heap)
    uint64_t *arrived;       // SEND: peer-mapped arrived[partitions] in the receiver's HBM; RECV: same array, local
    uint64_t *local_flags;   // SEND: sent[partitions] (local);  RECV: copied[partitions] (copy-out claim/complete)
    uint64_t *cts;           // SEND: local clear-to-send epoch, written remotely by the receiver at Start
    uint64_t *epoch;         // local current epoch (Start bumps it)
    int32_t zero_copy;       // SEND: user_buf needs no copy because producer wrote `land` directly (set by acx::remote_ptr users)
    int32_t peer;
    int32_t tag;
    int32_t pad;
    uint64_t *work;          // SEND: work[partitions], local CTA counters of the grid-cooperative Pready (pready_grid); null on RECV
};

#if defined(__CUDACC__) || defined(ACX_EMU)
namespace acx {

// ---- system-scope flag primitives ---------------------------------------------------
#ifdef ACX_EMU   // same names, C++ atomics: lets tests/emu run the protocol kernels on CPU threads
ACX_D uint64_t ld_acquire_sys(const uint64_t *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
ACX_D uint64_t ld_relaxed_sys(const uint64_t *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
ACX_D void st_release_sys(uint64_t *p, uint64_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#if defined(__SANITIZE_THREAD__)
// ThreadSanitizer does not model stand-alone fences, and the device code publishes with "__threadfence_system(); relaxed
// store" in places (legal on the GPU and in C++: fence-atomic synchronization). Under TSan a relaxed flag store is
// therefore treated as a release — i.e. TSan assumes the fence in front of it exists and checks everything else
// (reader-side acquires, plain accesses to shared records, slot reuse).
ACX_D void st_relaxed_sys(uint64_t *p, uint64_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#else
ACX_D void st_relaxed_sys(uint64_t *p, uint64_t v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
#endif
ACX_D void red_release_add_sys(uint64_t *p, uint64_t v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
ACX_D uint64_t atom_add_acqrel_sys(uint64_t *p, uint64_t v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
ACX_D uint64_t atom_cas_acqrel_sys(uint64_t *p, uint64_t cmp, uint64_t val) {
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
    return cmp;   // the value found (== the expected one on success), like atom.cas
}
#else
ACX_D uint64_t ld_acquire_sys(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
ACX_D uint64_t ld_relaxed_sys(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
ACX_D void st_release_sys(uint64_t *p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
ACX_D void st_relaxed_sys(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
ACX_D void red_release_add_sys(uint64_t *p, uint64_t v) {
    asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
ACX_D uint64_t atom_add_acqrel_sys(uint64_t *p, uint64_t v) {
    uint64_t old;
    asm volatile("atom.acq_rel.sys.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
    return old;
}
ACX_D uint64_t atom_cas_acqrel_sys(uint64_t *p, uint64_t cmp, uint64_t val) {
    uint64_t old;
    asm volatile("atom.acq_rel.sys.global.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "l"(p), "l"(cmp), "l"(val) : "memory");
    return old;
}
#endif  // ACX_EMU
// Spin until *p >= v (acquire). Exponential nanosleep backoff keeps the LSU quiet.
ACX_D void wait_geq_sys(const uint64_t *p, uint64_t v) {
    unsigned ns = 0;
    while (ld_acquire_sys(p) < v) {
        if (ns) __nanosleep(ns);
        ns = ns < 256 ? ns * 2 + 20 : 256;
    }
}

// ---- byte movers used by the single-thread / warp / CTA flavours ---------------------
// Copy n bytes with the widest access both pointers allow; `tid`/`nthr` describe the
// cooperating group (1 thread, a warp, or a CTA).
ACX_D void copy_bytes(char *dst, const char *src, uint64_t n, unsigned tid, unsigned nthr) {
    if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
        uint64_t nv = n >> 4;
        const int4 *s = (const int4 *)src;
        int4 *d = (int4 *)dst;
        uint64_t i = tid;
        // 4 independent 16-B loads in flight per thread before the stores (MLP over NVLink)
        for (; i + 3ull * nthr < nv; i += 4ull * nthr) {
            int4 a = s[i], b = s[i + nthr], c = s[i + 2ull * nthr], e = s[i + 3ull * nthr];
            d[i] = a; d[i + nthr] = b; d[i + 2ull * nthr] = c; d[i + 3ull * nthr] = e;
        }
        for (; i < nv; i += nthr) d[i] = s[i];
        for (uint64_t b = (nv << 4) + tid; b < n; b += nthr) dst[b] = src[b];
    } else if ((((uintptr_t)dst | (uintptr_t)src) & 3) == 0) {
        uint64_t nv = n >> 2;
        for (uint64_t i = tid; i < nv; i += nthr) ((int *)dst)[i] = ((const int *)src)[i];
        for (uint64_t b = (nv << 2) + tid; b < n; b += nthr) dst[b] = src[b];
    } else {
        for (uint64_t b = tid; b < n; b += nthr) dst[b] = src[b];
    }
}

// ---- partitioned device API ------------------------------------------------------------
ACX_D const acx_pchan *pchan(const void *preq) { return (const acx_pchan *)preq; }

// Pointer into the RECEIVER's memory where partition `p` of this send request lands.
// A producer kernel may store its output tile there directly (zero-copy over NVLink)
// and then call pready_signal(). Blocks until the receiver has started the round.
ACX_D char *remote_partition_ptr(const void *preq, int p) {
    const acx_pchan *ch = pchan(preq);
    wait_geq_sys(ch->cts, ld_relaxed_sys(ch->epoch));
    return ch->land + (uint64_t)p * ch->part_bytes;
}

// Flag-only Pready: the partition's bytes were already stored to remote_partition_ptr().
// Call from ONE thread after the writers' stores are ordered before it (e.g. after
// __syncthreads() in the producing CTA); the release makes them visible to the peer.
ACX_D int pready_signal(int p, const void *preq) {
    const acx_pchan *ch = pchan(preq);
    uint64_t e = ld_relaxed_sys(ch->epoch);
    __threadfence_system();
    st_release_sys(ch->arrived + p, e);
    st_release_sys(ch->local_flags + p, e);
    return 0;
}

// Explicit-epoch variants for persistent multi-step kernels that advance rounds on the
// device (no host MPIX_Start between rounds; the caller guarantees buffer reuse safety, e.g.
// by double buffering with a data dependence as the halo-stencil kernel does).
ACX_D char *remote_partition_ptr_nowait(const void *preq, int p) {
    const acx_pchan *ch = pchan(preq);
    return ch->land + (uint64_t)p * ch->part_bytes;
}
ACX_D void pready_signal_epoch(int p, const void *preq, uint64_t e) {
    const acx_pchan *ch = pchan(preq);
    __threadfence_system();
    st_release_sys(ch->arrived + p, e);
}
ACX_D bool parrived_epoch(const void *preq, int p, uint64_t e) { return ld_acquire_sys(pchan(preq)->arrived + p) >= e; }
ACX_D void wait_arrived_epoch(const void *preq, int p, uint64_t e) { wait_geq_sys(pchan(preq)->arrived + p, e); }

// Cooperative Pready: `nthr` threads (tid in [0,nthr)) of one warp or CTA copy the
// partition from the local send buffer into the peer's landing buffer, then thread 0
// publishes the flag. `sync` must be a barrier over exactly those threads.
template <typename Sync>
ACX_D int pready_group(int p, const void *preq, unsigned tid, unsigned nthr, Sync sync) {
    const acx_pchan *ch = pchan(preq);
    uint64_t e = ld_relaxed_sys(ch->epoch);
    if (tid == 0) wait_geq_sys(ch->cts, e);
    sync();
    uint64_t off = (uint64_t)p * ch->part_bytes;
    copy_bytes(ch->land + off, ch->user_buf + off, ch->part_bytes, tid, nthr);
    __threadfence_system();
    sync();
    if (tid == 0) {
        st_release_sys(ch->arrived + p, e);
        st_release_sys(ch->local_flags + p, e);
    }
    return 0;
}
ACX_D int pready_warp(int p, const void *preq) {
    return pready_group(p, preq, threadIdx.x & 31, 32, [] { __syncwarp(); });
}
ACX_D int pready_cta(int p, const void *preq) {
    unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    return pready_group(p, preq, tid, blockDim.x * blockDim.y * blockDim.z, [] { __syncthreads(); });
}

// Grid-cooperative Pready for LARGE partitions: `ncta` CTAs (this one is number `cta`, 0 <= cta < ncta) copy disjoint slices
// of partition p at NVLink bandwidth — one CTA moves ~20 GB/s, a 16-MiB activation panel wants tens of them — and the last
// one to finish publishes the flag (monotonic per-partition counter in local HBM: epoch-safe, nothing to reset).
// Launch e.g. ready<<<partitions * 32, 512>>> with p = blockIdx.x / 32, cta = blockIdx.x % 32, ncta = 32.
ACX_D int pready_grid(int p, const void *preq, unsigned cta, unsigned ncta) {
    const acx_pchan *ch = pchan(preq);
    const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z), nthr = blockDim.x * blockDim.y * blockDim.z;
    const uint64_t e = ld_relaxed_sys(ch->epoch);
    if (tid == 0) wait_geq_sys(ch->cts, e);
    __syncthreads();
    const uint64_t per = ((ch->part_bytes + ncta - 1) / ncta + 127) & ~127ull, lo = per * cta;
    if (lo < ch->part_bytes) {
        const uint64_t len = ch->part_bytes - lo < per ? ch->part_bytes - lo : per, off = (uint64_t)p * ch->part_bytes + lo;
        copy_bytes(ch->land + off, ch->user_buf + off, len, tid, nthr);
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        // work[p] counts CTA arrivals over all rounds; every round contributes exactly ncta of them (one grid shape per request)
        const uint64_t seen = atom_add_acqrel_sys(ch->work + p, 1ull) + 1;
        if (seen % ncta == 0) {
            __threadfence_system();
            st_release_sys(ch->arrived + p, e);
            st_release_sys(ch->local_flags + p, e);
        }
    }
    return 0;
}

// MPI-conformant single-thread Pready (what MPIX_Pready maps to on the device):
// the calling thread moves the partition itself, then releases the flag.
ACX_D int pready(int p, const void *preq) {
    return pready_group(p, preq, 0, 1, [] {});
}

// Non-blocking arrival test (MPIX_Parrived on the device). When the user's receive
// buffer is not in the symmetric heap the first poller to see the arrival copies the
// partition out of the landing buffer (claimed with a CAS); others see flag=0 until
// the copy is published.
ACX_D int parrived(const void *preq, int p, int *flag) {
    const acx_pchan *ch = pchan(preq);
    uint64_t e = ld_relaxed_sys(ch->epoch);
    if (ld_acquire_sys(ch->arrived + p) < e) { *flag = 0; return 0; }
    if (ch->land == ch->user_buf) { *flag = 1; return 0; }
    uint64_t done = 2 * e, claimed = 2 * e - 1;
    uint64_t c = ld_acquire_sys(ch->local_flags + p);
    if (c == done) { *flag = 1; return 0; }
    if (c == claimed) { *flag = 0; return 0; }
    if (atom_cas_acqrel_sys(ch->local_flags + p, c, claimed) != c) { *flag = 0; return 0; }
    uint64_t off = (uint64_t)p * ch->part_bytes;
    copy_bytes(ch->user_buf + off, ch->land + off, ch->part_bytes, 0, 1);
    __threadfence_system();
    st_release_sys(ch->local_flags + p, done);
    *flag = 1;
    return 0;
}

// Blocking convenience: spin (with backoff) until partition p has arrived. One thread.
ACX_D void wait_arrived(const void *preq, int p) {
    int f = 0;
    unsigned ns = 0;
    for (;;) {
        parrived(preq, p, &f);
        if (f) return;
        if (ns) __nanosleep(ns);
        ns = ns < 256 ? ns * 2 + 20 : 256;
    }
}

// Local pointer to where partition p's data is readable after arrival (RECV side).
ACX_D const char *local_partition_ptr(const void *preq, int p) {
    const acx_pchan *ch = pchan(preq);
    return ch->land + (uint64_t)p * ch->part_bytes;
}

}  // namespace acx
#endif  // __CUDACC__
#endif
