// Ring-bandwidth probe: which copy kernel moves 256 MiB per GPU around a ring of N GPUs fastest?
// One process, peer access, every GPU runs the same variant at the same time (bidirectional NVLink load, as in
// bench.py's headline). Prints one line per variant: GB/s per GPU (max time over GPUs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/bwprobe.cu -o tools/bwprobe
//   tools/bwprobe [ngpus] [bytes] [uni]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// ---- LDG/STG flavours --------------------------------------------------------------------------------
template <int UNROLL, int MODE>   // MODE 0: default, 1: ld.nc.no_allocate + st, 2: .cs both
__device__ __forceinline__ void copy16(int4 *__restrict__ d, const int4 *__restrict__ s, uint64_t nv, uint64_t tid, uint64_t nthr) {
    uint64_t i = tid;
    for (; i + (uint64_t)(UNROLL - 1) * nthr < nv; i += (uint64_t)UNROLL * nthr) {
        int4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int4 *p = s + i + (uint64_t)u * nthr;
            if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p));
            else if (MODE == 2) v[u] = __ldcs(p);
            else v[u] = *p;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            int4 *p = d + i + (uint64_t)u * nthr;
            if (MODE == 2) __stcs(p, v[u]); else *p = v[u];
        }
    }
    for (; i < nv; i += nthr) d[i] = s[i];
}
// chunked: each CTA owns a contiguous range (blocked) or sweeps interleaved blocks of `blk` bytes
template <int UNROLL, int MODE>
__global__ void __launch_bounds__(512) k_ldst(char *d, const char *s, uint64_t n, uint64_t blk) {
    if (blk == 0) {
        const uint64_t per = ((n + gridDim.x - 1) / gridDim.x + 127) & ~127ull, lo = per * blockIdx.x;
        if (lo >= n) return;
        const uint64_t len = n - lo < per ? n - lo : per;
        copy16<UNROLL, MODE>((int4 *)(d + lo), (const int4 *)(s + lo), len >> 4, threadIdx.x, blockDim.x);
    } else {
        for (uint64_t lo = (uint64_t)blockIdx.x * blk; lo < n; lo += (uint64_t)gridDim.x * blk) {
            const uint64_t len = n - lo < blk ? n - lo : blk;
            copy16<UNROLL, MODE>((int4 *)(d + lo), (const int4 *)(s + lo), len >> 4, threadIdx.x, blockDim.x);
        }
    }
}

// ---- TMA bulk ring: global -> smem -> global, one elected thread per CTA ------------------------------
__device__ __forceinline__ uint32_t sa(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int STAGES>
__global__ void __launch_bounds__(128) k_tma(char *dst, const char *src, uint64_t n, uint32_t chunk, uint64_t blk) {
    extern __shared__ __align__(128) unsigned char sm[];
    if (threadIdx.x != 0) return;
    uint64_t *bars = (uint64_t *)(sm + (size_t)STAGES * chunk);
    for (int i = 0; i < STAGES; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&bars[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the list of chunks this CTA moves: interleaved blocks of `blk` bytes (blk multiple of chunk) or one contiguous range
    uint64_t per, lo0, stride;
    if (blk) { per = blk; lo0 = (uint64_t)blockIdx.x * blk; stride = (uint64_t)gridDim.x * blk; }
    else { per = ((n + gridDim.x - 1) / gridDim.x + chunk - 1) / chunk * chunk; lo0 = per * blockIdx.x; stride = ~0ull >> 1; }
    const uint64_t cpb = per / chunk;   // chunks per block
    // total chunks for this CTA
    uint64_t total = 0;
    for (uint64_t lo = lo0; lo < n; lo += stride) { const uint64_t len = n - lo < per ? n - lo : per; total += (len + chunk - 1) / chunk; if (!blk) break; }
    auto addr_of = [&](uint64_t c, uint32_t &bytes) -> uint64_t {
        const uint64_t b = c / cpb, w = c % cpb;
        const uint64_t off = lo0 + b * stride + w * chunk;
        const uint64_t rem = n - off;
        bytes = (uint32_t)(rem < chunk ? rem : chunk);
        return off;
    };
    uint32_t phase = 0;
    auto load = [&](uint64_t c) {
        const int st = (int)(c % STAGES);
        uint32_t bytes; const uint64_t off = addr_of(c, bytes);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(&bars[st])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(sa(sm + (size_t)st * chunk)), "l"(src + off), "r"(bytes), "r"(sa(&bars[st])) : "memory");
    };
    const uint64_t pre = total < (uint64_t)STAGES ? total : (uint64_t)STAGES;
    for (uint64_t c = 0; c < pre; c++) load(c);
    for (uint64_t c = 0; c < total; c++) {
        const int st = (int)(c % STAGES);
        const uint32_t parity = (phase >> st) & 1;
        asm volatile("{\n\t.reg .pred p;\n\tTW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni TD;\n\tbra.uni TW;\n\tTD:\n\t}" ::"r"(sa(&bars[st])), "r"(parity) : "memory");
        phase ^= 1u << st;
        uint32_t bytes; const uint64_t off = addr_of(c, bytes);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(sa(sm + (size_t)st * chunk)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (c >= 1 && c - 1 + STAGES < total) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            load(c - 1 + STAGES);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct Dev { char *src, *dst; cudaStream_t st, st2; cudaEvent_t e0, e1, ef, ej; };
static int N; static uint64_t BYTES; static std::vector<Dev> D; static bool UNI = false;

// fn(dev, push?) launches the work of one GPU for one step
template <class F> static void run(const char *name, F fn, int iters = 8) {
    for (int rep = 0; rep < 2; rep++) {   // rep 0 = warm-up
        for (int d = 0; d < N; d++) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
        for (int d = 0; d < N; d++) { if (UNI && d) continue; CK(cudaSetDevice(d)); CK(cudaEventRecord(D[d].e0, D[d].st)); }
        for (int it = 0; it < (rep ? iters : 2); it++)
            for (int d = 0; d < N; d++) { if (UNI && d) continue; CK(cudaSetDevice(d)); fn(d); }
        for (int d = 0; d < N; d++) { if (UNI && d) continue; CK(cudaSetDevice(d)); CK(cudaEventRecord(D[d].e1, D[d].st)); }
        float worst = 0;
        for (int d = 0; d < N; d++) { if (UNI && d) continue; CK(cudaSetDevice(d)); CK(cudaEventSynchronize(D[d].e1)); float ms; CK(cudaEventElapsedTime(&ms, D[d].e0, D[d].e1)); worst = std::max(worst, ms); }
        CK(cudaGetLastError());
        if (rep) printf("%-52s %8.1f GB/s per GPU  (%.4f ms/step)\n", name, BYTES / (worst / iters) / 1e6, worst / iters);
    }
    fflush(stdout);
}

int main(int argc, char **argv) {
    CK(cudaGetDeviceCount(&N));
    if (argc > 1 && atoi(argv[1]) > 0) N = std::min(N, atoi(argv[1]));
    BYTES = argc > 2 ? strtoull(argv[2], 0, 0) : (256ull << 20);
    UNI = argc > 3 && std::string(argv[3]) == "uni";
    D.resize(N);
    for (int d = 0; d < N; d++) {
        CK(cudaSetDevice(d));
        for (int p = 0; p < N; p++) if (p != d) { cudaError_t e = cudaDeviceEnablePeerAccess(p, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { fprintf(stderr, "peer %d->%d: %s\n", d, p, cudaGetErrorString(e)); return 1; } cudaGetLastError(); }
        CK(cudaMalloc(&D[d].src, BYTES)); CK(cudaMalloc(&D[d].dst, BYTES));
        CK(cudaMemset(D[d].src, d + 1, BYTES)); CK(cudaMemset(D[d].dst, 0, BYTES));
        CK(cudaStreamCreateWithFlags(&D[d].st, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&D[d].st2, cudaStreamNonBlocking));
        CK(cudaEventCreate(&D[d].e0)); CK(cudaEventCreate(&D[d].e1)); CK(cudaEventCreateWithFlags(&D[d].ef, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&D[d].ej, cudaEventDisableTiming));
    }
    printf("# bwprobe ngpus=%d bytes=%llu %s\n", N, (unsigned long long)BYTES, UNI ? "unidirectional (GPU0 only)" : "ring (every GPU sends to d+1)");
    auto right = [&](int d) { return (d + 1) % N; };
    auto left = [&](int d) { return (d + N - 1) % N; };

    run("copy engine: cudaMemcpyAsync push", [&](int d) { CK(cudaMemcpyAsync(D[right(d)].dst, D[d].src, BYTES, cudaMemcpyDeviceToDevice, D[d].st)); });
    if (N > 1) run("copy engine: cudaMemcpyAsync pull", [&](int d) { CK(cudaMemcpyAsync(D[d].dst, D[left(d)].src, BYTES, cudaMemcpyDeviceToDevice, D[d].st)); });

    char nm[160];
    const int grids[] = {74, 148, 296, 592};
#define LDST(U, M, tag) \
    for (int push = 1; push >= 0; push--) for (int g : grids) for (uint64_t blk : {0ull, 131072ull}) { \
        if (N == 1 && !push) continue; \
        snprintf(nm, sizeof nm, "ldst %s unroll=%d grid=%d blk=%llu %s", tag, U, g, (unsigned long long)blk, push ? "push" : "pull"); \
        run(nm, [&](int d) { if (push) k_ldst<U, M><<<g, 512, 0, D[d].st>>>(D[right(d)].dst, D[d].src, BYTES, blk); else k_ldst<U, M><<<g, 512, 0, D[d].st>>>(D[d].dst, D[left(d)].src, BYTES, blk); }); }
    LDST(4, 0, "plain")
    LDST(8, 0, "plain")
    LDST(4, 1, "nc.noalloc")
    LDST(8, 2, "cs")
    // half push + half pull, concurrently on two streams (every GPU moves half of its outgoing and half of its incoming message)
    if (N > 1) for (int g : {74, 148, 296}) {
        snprintf(nm, sizeof nm, "ldst plain unroll=8 half-push+half-pull grid=2x%d", g);
        run(nm, [&](int d) {
            const uint64_t h = BYTES / 2;
            CK(cudaEventRecord(D[d].ef, D[d].st)); CK(cudaStreamWaitEvent(D[d].st2, D[d].ef, 0));
            k_ldst<8, 0><<<g, 512, 0, D[d].st>>>(D[right(d)].dst, D[d].src, h, 0);
            k_ldst<8, 0><<<g, 512, 0, D[d].st2>>>(D[d].dst + h, D[left(d)].src + h, BYTES - h, 0);
            CK(cudaEventRecord(D[d].ej, D[d].st2)); CK(cudaStreamWaitEvent(D[d].st, D[d].ej, 0));
        });
    }
    // TMA
#define TMA(S) \
    for (int push = 1; push >= 0; push--) for (uint32_t chunk : {16384u, 32768u}) for (int g : {148, 296, 444}) for (uint64_t blk : {0ull, 4ull}) { \
        if (N == 1 && !push) continue; \
        const size_t smem = (size_t)S * chunk + 64; if (smem > 227 * 1024) continue; \
        if ((size_t)(g / 148) * smem > 227 * 1024) continue; \
        snprintf(nm, sizeof nm, "tma stages=%d chunk=%u grid=%d blk=%llu %s", S, chunk, g, (unsigned long long)(blk * chunk), push ? "push" : "pull"); \
        run(nm, [&](int d) { for (int x = 0; x < 1; x++) { if (push) k_tma<S><<<g, 128, smem, D[d].st>>>(D[right(d)].dst, D[d].src, BYTES, chunk, blk * chunk); else k_tma<S><<<g, 128, smem, D[d].st>>>(D[d].dst, D[left(d)].src, BYTES, chunk, blk * chunk); } }); }
    for (int d = 0; d < N; d++) { CK(cudaSetDevice(d)); CK(cudaFuncSetAttribute(k_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); CK(cudaFuncSetAttribute(k_tma<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); }
    TMA(4)
    TMA(6)
    // verify last variant's data
    for (int d = 0; d < N; d++) { if (UNI && d) continue; CK(cudaSetDevice(d)); unsigned char b[2]; CK(cudaMemcpy(b, D[d].dst + BYTES - 2, 2, cudaMemcpyDeviceToHost)); if (b[0] != (unsigned char)(left(d) + 1) && !UNI) printf("verify FAIL dev %d got %d\n", d, b[0]); }
    printf("# done\n");
    return 0;
}
