// fp8 (e4m3) GEMM for sm_100a with a partition-gated TMA producer and a communicating epilogue:
//     Y[M][N] (e4m3) = scale * X[M][K] (e4m3) * W[N][K]^T (e4m3),  fp32 accumulation in TMEM
//
// Pipeline-parallel use (BASELINE.json config #4): X is the previous stage's output, delivered
// by that stage's epilogue straight into this GPU's HBM one row-panel (partition) at a time;
// this kernel's TMA warp waits on the panel's arrival flag (acquire on LOCAL memory) before it
// loads tiles of that panel, and this kernel's epilogue stores Y straight into the NEXT stage's
// input buffer (peer-mapped, over NVLink) and releases a flag per completed panel. One kernel,
// no host in the loop, compute and transfer overlapped tile by tile.
//
// Structure (one CTA per SM, persistent over 128x256 output tiles, 256 threads):
//   warp 0      : TMA producer   (cp.async.bulk.tensor 2D, SWIZZLE_128B, 4-stage mbarrier ring)
//   warp 1      : MMA issuer     (one elected lane: tcgen05.mma.cta_group::1.kind::f8f6f4, M128 N256 K32)
//   warp 2      : TMEM allocator (512 columns = two 128x256 fp32 accumulators, double-buffered)
//   warps 4..7  : epilogue       (tcgen05.ld 32x32b -> scale -> e4m3 -> swizzled smem -> 16-B row-contiguous
//                                 stores to local or peer memory -> per-panel completion count -> release flag)
// Nothing here comes from the reference (it has no compute kernels at all).
#ifndef ACX_GEMM_FP8_SM100_CUH
#define ACX_GEMM_FP8_SM100_CUH

#include <cuda.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>

#ifndef ACX_GEMM_TMA_STORE_DEFAULT
#define ACX_GEMM_TMA_STORE_DEFAULT 1   // local outputs go out by TMA; 2 = the next stage's buffer too
#endif
#ifndef ACX_GEMM_PAIR_DEFAULT
#define ACX_GEMM_PAIR_DEFAULT 1   // the CTA-pair kernel: 3.28 PFLOP/s at 8192^3 against 2.82 for the single-CTA kernel (profiles/gemm_pair_ncu.md)
#endif

namespace acxg {
namespace {   // internal linkage: every module that includes this header owns a private copy of the kernel and its
              // host stub (a shared copy would be interposed across libacx.so / executables, each of which
              // carries its own static CUDA runtime)

constexpr int BM = 128, BN = 256, BK = 128;      // tile; BK in bytes == elements (fp8)
constexpr int UMMA_K = 32;                       // bytes of K per tcgen05.mma (8-bit types)
constexpr int STAGES = 4;
constexpr int A_STAGE = BM * BK, B_STAGE = BN * BK;          // 16 KiB + 32 KiB
constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
constexpr int OUT_TILE = BM * BN;                             // 32 KiB of e4m3
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_TILE + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int THREADS = 256;
constexpr int TMEM_COLS = 512;

struct Gating {            // all pointers may be null (= feature off)
    const unsigned long long *in_arrived;   // local: arrival flag per input row-panel
    unsigned long long in_epoch;            // wait until in_arrived[p] >= in_epoch
    unsigned long long *out_arrived;        // peer-mapped: arrival flag per output row-panel at the next stage
    unsigned long long out_epoch;           // value to publish
    unsigned long long *panel_count;        // local: completed-tile counters per panel (monotonic)
    unsigned long long panel_target;        // publish when count reaches this
    const unsigned long long *credit;       // local: written by the next stage when it finished reading its input buffer
    unsigned long long credit_need;         // wait until *credit >= credit_need before the first remote store
    unsigned long long *done_count;         // local: CTAs finished (monotonic)
    unsigned long long done_target;         // the CTA that reaches it ...
    unsigned long long *credit_out;         // ... publishes credit_out_val to the PREVIOUS stage (peer-mapped)
    unsigned long long credit_out_val;
    int rows_per_panel;
};

struct Params {
    CUtensorMap tmA, tmB;   // X [M][K] and W [N][K], both K-major, box {128 B, BM|BN rows}, SWIZZLE_128B
    CUtensorMap tmC;        // Y [M][N], box {128 B, 128 rows}, SWIZZLE_128B — used when tma_store is set
    int tma_store;          // the epilogue hands the staged tile to TMA (two cp.async.bulk.tensor stores per tile) instead of storing it with 128 threads
    uint8_t *out;           // Y [M][N] e4m3 (local or peer-mapped)
    int M, N, K;
    float scale;
    Gating g;
};

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *tm, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, const void *src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // the staged tile has been read: the buffer may be overwritten
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }             // the stores have been performed
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm) { asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in smem, 128 B of K per row, SWIZZLE_128B: 8-row x 128-B atoms, 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);        // start address
    d |= (uint64_t)0 << 16;                              // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}
// instruction descriptor: D=f32, A=B=e4m3, both K-major, N=BN, M=BM
__device__ __forceinline__ uint32_t umma_idesc_e4m3() {
    uint32_t d = 0;
    d |= 1u << 4;                     // c_format = F32
    d |= 0u << 7;                     // a_format = E4M3
    d |= 0u << 10;                    // b_format = E4M3
    d |= 0u << 15;                    // a_major = K
    d |= 0u << 16;                    // b_major = K
    d |= (uint32_t)(BN >> 3) << 17;   // n_dim
    d |= (uint32_t)(BM >> 4) << 24;   // m_dim
    return d;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void spin_geq_sys(const unsigned long long *p, unsigned long long v) {
    unsigned ns = 0;
    while (ld_acquire_sys_u64(p) < v) { if (ns) __nanosleep(ns); ns = ns < 512 ? ns * 2 + 32 : 512; }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 1) gemm_fp8_kernel(const __grid_constant__ Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024-B alignment
    uint8_t *sA = smem, *sB = smem + STAGES * A_STAGE, *sOut = smem + STAGES * STAGE_BYTES;
    uint64_t *bars = (uint64_t *)(sOut + OUT_TILE);
    uint64_t *full = bars, *empty = bars + STAGES, *tfull = bars + 2 * STAGES, *tempty = bars + 2 * STAGES + 2;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = P.N / BN, tiles_m = P.M / BM, ntiles = tiles_m * tiles_n, kblocks = P.K / BK;

    if (warp == 0 && elect_one()) { tma_prefetch_desc(&P.tmA); tma_prefetch_desc(&P.tmB); if (P.tma_store) tma_prefetch_desc(&P.tmC); }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int tm = t / tiles_n, tn = t % tiles_n;
                if (P.g.in_arrived) {
                    // this tile's rows belong to one input partition: wait until the previous stage delivered it
                    spin_geq_sys(P.g.in_arrived + (tm * BM) / P.g.rows_per_panel, P.g.in_epoch);
                    asm volatile("fence.proxy.async;" ::: "memory");   // remote generic-proxy stores -> our async-proxy (TMA) reads
                }
                for (int kb = 0; kb < kblocks; kb++) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_2d(&P.tmA, &full[stage], sA + stage * A_STAGE, kb * BK, tm * BM);
                    tma_load_2d(&P.tmB, &full[stage], sB + stage * B_STAGE, kb * BK, tn * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = umma_idesc_e4m3();
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);          // epilogue drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)acc * BN;
            for (int kb = 0; kb < kblocks; kb++) {
                mbar_wait(&full[stage], phase);               // TMA landed this stage
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t ad = umma_desc_sw128(smem_u32(sA + stage * A_STAGE));
                    const uint64_t bd = umma_desc_sw128(smem_u32(sB + stage * B_STAGE));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; k++)
                        tc_mma_f8(tmem_d, ad + (uint64_t)(k * UMMA_K >> 4), bd + (uint64_t)(k * UMMA_K >> 4), idesc, (kb | k) != 0);
                    tc_commit(&empty[stage]);                  // frees the smem stage when these MMAs retire
                    if (kb == kblocks - 1) tc_commit(&tfull[acc]);   // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===== epilogue: warps 4..7 own TMEM lanes 32*(warp-4) .. +31 (== tile rows) =====
        const int ew = warp - 4, row = ew * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        bool credit_ok = P.g.credit == nullptr;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int tm = t / tiles_n, tn = t % tiles_n;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            // all four warps (or the TMA store of the previous tile) must be done reading sOut before it is overwritten
            if (P.tma_store && threadIdx.x == 128) bulk_wait_read0();
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < BN / 32; c++) {
                uint32_t v[32];
                tc_ld_32x32b_x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
                tc_wait_ld();
                uint32_t packed[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float f0 = __uint_as_float(v[4 * i]) * P.scale, f1 = __uint_as_float(v[4 * i + 1]) * P.scale;
                    const float f2 = __uint_as_float(v[4 * i + 2]) * P.scale, f3 = __uint_as_float(v[4 * i + 3]) * P.scale;
                    const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(f0, f1), __NV_SATFINITE, __NV_E4M3);
                    const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(f2, f3), __NV_SATFINITE, __NV_E4M3);
                    packed[i] = lo | (hi << 16);
                }
                // row-major staging, 16-B chunks XOR-swizzled by row so both this write and the read below are conflict-free
                const int ch0 = 2 * c, ch1 = 2 * c + 1;
                if (P.tma_store) {   // two 128-row x 128-B boxes in the TMA SWIZZLE_128B layout (16-B chunk index XOR row mod 8)
                    *reinterpret_cast<uint4 *>(sOut + (ch0 >> 3) * (BM * 128) + row * 128 + (((ch0 & 7) ^ (row & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    *reinterpret_cast<uint4 *>(sOut + (ch1 >> 3) * (BM * 128) + row * 128 + (((ch1 & 7) ^ (row & 7)) << 4)) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                } else {
                    *reinterpret_cast<uint4 *>(sOut + row * BN + ((ch0 ^ (row & 15)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    *reinterpret_cast<uint4 *>(sOut + row * BN + ((ch1 ^ (row & 15)) << 4)) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                }
            }
            // accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (P.tma_store) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this thread's staging writes before the async proxy's reads
            asm volatile("bar.sync 1, 128;" ::: "memory");      // staging tile complete
            if (!credit_ok) {                                    // next stage still reading the buffer we are about to overwrite?
                if (threadIdx.x == 128) spin_geq_sys(P.g.credit, P.g.credit_need);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                credit_ok = true;
            }
            // write-out: TMA (two bulk tensor stores per tile), or coalesced — each warp instruction stores two full 256-B rows
            if (P.tma_store) {
                // (the writers' fence.proxy.async sits in front of the "staging tile complete" barrier above)
                if (threadIdx.x == 128) {
                    tma_store_2d(&P.tmC, sOut, tn * BN, tm * BM);
                    tma_store_2d(&P.tmC, sOut + BM * 128, tn * BN + 128, tm * BM);
                    bulk_commit();
                    if (P.g.panel_count) bulk_wait0();   // the tile must have landed (at the next stage) before it is counted
                }
            } else {
                uint8_t *gout = P.out + (size_t)(tm * BM) * P.N + (size_t)tn * BN;
                const int et = threadIdx.x - 128;
#pragma unroll 4
                for (int i = 0; i < (BM * BN / 16) / 128; i++) {
                    const int idx = i * 128 + et, r = idx >> 4, ch = idx & 15;
                    const uint4 val = *reinterpret_cast<const uint4 *>(sOut + r * BN + ((ch ^ (r & 15)) << 4));
                    *reinterpret_cast<uint4 *>(gout + (size_t)r * P.N + (ch << 4)) = val;
                }
            }
            if (P.g.panel_count) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (threadIdx.x == 128) {
                    __threadfence_system();
                    const int panel = (tm * BM) / P.g.rows_per_panel;
                    const unsigned long long old = atomicAdd(P.g.panel_count + panel, 1ull);
                    if (old + 1 == P.g.panel_target && P.g.out_arrived) {
                        __threadfence_system();
                        st_release_sys_u64(P.g.out_arrived + panel, P.g.out_epoch);   // acx Pready: partition has landed at the next stage
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (P.tma_store && threadIdx.x == 128) bulk_wait0();   // shared memory must outlive the last tile's stores
    }
    // ===== teardown =====
    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    if (threadIdx.x == 0 && P.g.done_count) {
        // all of this CTA's TMA reads of the input buffer have completed (their data was consumed by retired MMAs)
        __threadfence();
        const unsigned long long old = atomicAdd(P.g.done_count, 1ull);
        if (old + 1 == P.g.done_target && P.g.credit_out) st_release_sys_u64(P.g.credit_out, P.g.credit_out_val);
    }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs on the two SMs of a TPC owns one 256x256 output tile.
// CTA r of the pair loads rows [128 r, 128 r + 128) of the tile's X panel and rows [128 r, +128) of its W panel — half of
// each operand — and ONE thread of the leader CTA (rank 0) issues tcgen05.mma.cta_group::2 M256 N256 K32, which reads both
// CTAs' shared memory and accumulates rows 128 r .. +127 into CTA r's tensor memory. Per SM and MMA the shared-memory read
// traffic drops from 4 KiB (A) + 8 KiB (B) to 4 + 4 KiB and the TMA fill traffic from 48 to 32 KiB per k-block, which is
// what the single-CTA kernel is bound by (tensor pipe 87 % active, profiles/gemm_fp8_ncu.md); the smaller stages also buy a
// 6-deep ring. Synchronisation across the pair:
//   full[s]   lives in the LEADER: its producer arms it with the bytes of BOTH CTAs' loads; the peer's TMA loads are
//             cp.async.bulk.tensor ... .cta_group::2 pointing at the leader's barrier (mapa);
//   empty[s]  in each CTA: the leader's tcgen05.commit is multicast to both, so each producer sees its stage freed;
//   tfull[a]  in each CTA, same multicast commit: both epilogues start on their own half of the accumulator;
//   tempty[a] in the leader, 8 arrivals: the four epilogue warps of both CTAs (remote mbarrier.arrive for the peer's).
// The epilogue and the partition gating are per CTA and identical to the single-CTA kernel's (each CTA owns 128 rows).
constexpr int STAGES2 = 6;
constexpr int A_STAGE2 = BM * BK, B_STAGE2 = (BN / 2) * BK;          // 16 KiB + 16 KiB per CTA
constexpr int STAGE_BYTES2 = A_STAGE2 + B_STAGE2;
constexpr int SMEM_BYTES2 = STAGES2 * STAGE_BYTES2 + OUT_TILE + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {   // address of the same shared-memory location in CTA `rank` of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on a barrier that may live in the PEER CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap *tm, uint32_t bar_cluster_addr, void *dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar) {   // arrives on the barrier at this offset in BOTH CTAs when the MMAs issued so far retire
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t umma_idesc_e4m3_pair() {   // D=f32, A=B=e4m3, K-major, N=256, M=256 (both CTAs)
    uint32_t d = 0;
    d |= 1u << 4;
    d |= (uint32_t)(BN >> 3) << 17;
    d |= (uint32_t)((2 * BM) >> 4) << 24;
    return d;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) gemm_fp8_pair_kernel(const __grid_constant__ Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem, *sB = smem + STAGES2 * A_STAGE2, *sOut = smem + STAGES2 * STAGE_BYTES2;
    uint64_t *bars = (uint64_t *)(sOut + OUT_TILE);
    uint64_t *full = bars, *empty = bars + STAGES2, *tfull = bars + 2 * STAGES2, *tempty = bars + 2 * STAGES2 + 2;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * STAGES2 + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int tiles_n = P.N / BN, tiles_m2 = P.M / (2 * BM), ntiles = tiles_m2 * tiles_n, kblocks = P.K / BK;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

    if (warp == 0 && elect_one()) { tma_prefetch_desc(&P.tmA); tma_prefetch_desc(&P.tmB); if (P.tma_store) tma_prefetch_desc(&P.tmC); }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < STAGES2; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // one warp of EACH CTA of the pair issues the pair-wide allocation
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer's barriers are initialised and its tensor memory allocated before anything is signalled across
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs: own half of each operand; completion lands on the leader's full barrier) =====
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int t = cluster_id; t < ntiles; t += nclusters) {
                const int tm = 2 * (t / tiles_n) + (int)rank, tn = t % tiles_n;
                if (P.g.in_arrived) {
                    spin_geq_sys(P.g.in_arrived + (tm * BM) / P.g.rows_per_panel, P.g.in_epoch);
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                for (int kb = 0; kb < kblocks; kb++) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (leader) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES2);
                    const uint32_t fb = mapa_shared(smem_u32(&full[stage]), 0);
                    tma_load_2d_pair(&P.tmA, fb, sA + stage * A_STAGE2, kb * BK, tm * BM);
                    tma_load_2d_pair(&P.tmB, fb, sB + stage * B_STAGE2, kb * BK, tn * BN + (int)rank * (BN / 2));
                    if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (leader) {
            const uint32_t idesc = umma_idesc_e4m3_pair();
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int t = cluster_id; t < ntiles; t += nclusters) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);          // both CTAs' epilogues drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)acc * BN;
                for (int kb = 0; kb < kblocks; kb++) {
                    mbar_wait(&full[stage], phase);               // both CTAs' loads landed
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = umma_desc_sw128(smem_u32(sA + stage * A_STAGE2));
                        const uint64_t bd = umma_desc_sw128(smem_u32(sB + stage * B_STAGE2));
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; k++)
                            tc_mma_f8_pair(tmem_d, ad + (uint64_t)(k * UMMA_K >> 4), bd + (uint64_t)(k * UMMA_K >> 4), idesc, (kb | k) != 0);
                        tc_commit_pair(&empty[stage]);
                        if (kb == kblocks - 1) tc_commit_pair(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): warps 4..7 own TMEM lanes 32*(warp-4) .. +31 == rows of this CTA's half of the tile =====
        const int ew = warp - 4, row = ew * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        bool credit_ok = P.g.credit == nullptr;
        for (int t = cluster_id; t < ntiles; t += nclusters) {
            const int tm = 2 * (t / tiles_n) + (int)rank, tn = t % tiles_n;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            if (P.tma_store && threadIdx.x == 128) bulk_wait_read0();
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < BN / 32; c++) {
                uint32_t v[32];
                tc_ld_32x32b_x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
                tc_wait_ld();
                uint32_t packed[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float f0 = __uint_as_float(v[4 * i]) * P.scale, f1 = __uint_as_float(v[4 * i + 1]) * P.scale;
                    const float f2 = __uint_as_float(v[4 * i + 2]) * P.scale, f3 = __uint_as_float(v[4 * i + 3]) * P.scale;
                    const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(f0, f1), __NV_SATFINITE, __NV_E4M3);
                    const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(f2, f3), __NV_SATFINITE, __NV_E4M3);
                    packed[i] = lo | (hi << 16);
                }
                const int ch0 = 2 * c, ch1 = 2 * c + 1;
                if (P.tma_store) {   // two 128-row x 128-B boxes in the TMA SWIZZLE_128B layout (16-B chunk index XOR row mod 8)
                    *reinterpret_cast<uint4 *>(sOut + (ch0 >> 3) * (BM * 128) + row * 128 + (((ch0 & 7) ^ (row & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    *reinterpret_cast<uint4 *>(sOut + (ch1 >> 3) * (BM * 128) + row * 128 + (((ch1 & 7) ^ (row & 7)) << 4)) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                } else {
                    *reinterpret_cast<uint4 *>(sOut + row * BN + ((ch0 ^ (row & 15)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    *reinterpret_cast<uint4 *>(sOut + row * BN + ((ch1 ^ (row & 15)) << 4)) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[acc]), 0));   // the leader's MMA warp may overwrite this accumulator (in both CTAs) once all 8 warps are here
            if (P.tma_store) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (!credit_ok) {
                if (threadIdx.x == 128) spin_geq_sys(P.g.credit, P.g.credit_need);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                credit_ok = true;
            }
            if (P.tma_store) {
                // (the writers' fence.proxy.async sits in front of the "staging tile complete" barrier above)
                if (threadIdx.x == 128) {
                    tma_store_2d(&P.tmC, sOut, tn * BN, tm * BM);
                    tma_store_2d(&P.tmC, sOut + BM * 128, tn * BN + 128, tm * BM);
                    bulk_commit();
                    if (P.g.panel_count) bulk_wait0();   // the tile must have landed (at the next stage) before it is counted
                }
            } else {
                uint8_t *gout = P.out + (size_t)(tm * BM) * P.N + (size_t)tn * BN;
                const int et = threadIdx.x - 128;
#pragma unroll 4
                for (int i = 0; i < (BM * BN / 16) / 128; i++) {
                    const int idx = i * 128 + et, r = idx >> 4, ch = idx & 15;
                    const uint4 val = *reinterpret_cast<const uint4 *>(sOut + r * BN + ((ch ^ (r & 15)) << 4));
                    *reinterpret_cast<uint4 *>(gout + (size_t)r * P.N + (ch << 4)) = val;
                }
            }
            if (P.g.panel_count) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (threadIdx.x == 128) {
                    __threadfence_system();
                    const int panel = (tm * BM) / P.g.rows_per_panel;
                    const unsigned long long old = atomicAdd(P.g.panel_count + panel, 1ull);
                    if (old + 1 == P.g.panel_target && P.g.out_arrived) {
                        __threadfence_system();
                        st_release_sys_u64(P.g.out_arrived + panel, P.g.out_epoch);
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (P.tma_store && threadIdx.x == 128) bulk_wait0();
    }
    // ===== teardown: neither CTA may leave (or free tensor memory) while the other can still signal its barriers or read its smem =====
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    if (threadIdx.x == 0 && P.g.done_count) {
        __threadfence();
        const unsigned long long old = atomicAdd(P.g.done_count, 1ull);
        if (old + 1 == P.g.done_target && P.g.credit_out) st_release_sys_u64(P.g.credit_out, P.g.credit_out_val);
    }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    }
    return fn;
}
// [rows][K] uint8, K contiguous; box = {128 bytes of K, box_rows}
inline int make_tmap_kmajor(CUtensorMap *tm, const void *base, uint64_t rows, uint64_t K, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return 1;
    cuuint64_t dims[2] = {K, rows};
    cuuint64_t strides[1] = {K};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 2;
}

// Tensor maps are pure functions of (base, rows, K, box rows): encode once per operand and keep them — a pipeline stage
// launches the same two or three operands for every microbatch (round 1 re-encoded both maps on every launch).
inline bool cached_tmap_copy(CUtensorMap *out, const void *base, uint64_t rows, uint64_t K, uint32_t box_rows);
inline const CUtensorMap *cached_tmap(const void *base, uint64_t rows, uint64_t K, uint32_t box_rows) {
    struct Key { const void *base; uint64_t rows, K; uint32_t box; };
    struct Slot { Key k; CUtensorMap tm; };
    static Slot cache[16];
    static int used = 0, next = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    for (int i = 0; i < used; i++)
        if (cache[i].k.base == base && cache[i].k.rows == rows && cache[i].k.K == K && cache[i].k.box == box_rows) return &cache[i].tm;
    Slot &s = cache[used < 16 ? used++ : (next++ % 16)];
    if (make_tmap_kmajor(&s.tm, base, rows, K, box_rows)) { s.k.base = nullptr; return nullptr; }
    s.k = Key{base, rows, K, box_rows};
    return &s.tm;
}

inline bool cached_tmap_copy(CUtensorMap *out, const void *base, uint64_t rows, uint64_t K, uint32_t box_rows) {
    static std::mutex mu2;
    std::lock_guard<std::mutex> lk(mu2);   // the copy happens under this lock: a concurrent launch may evict the slot
    const CUtensorMap *t = cached_tmap(base, rows, K, box_rows);
    if (!t) return false;
    *out = *t;
    return true;
}

// CTA-pair kernel (cta_group::2) when the shape allows it (M a multiple of 256); ACX_GEMM_PAIR=0 keeps the single-CTA kernel.
inline bool gemm_use_pair(int M) {
    static int on = -1;
    if (on < 0) { const char *e = getenv("ACX_GEMM_PAIR"); on = e ? atoi(e) != 0 : ACX_GEMM_PAIR_DEFAULT; }
    return on && M % (2 * BM) == 0;
}
// Number of CTAs launch_gemm_fp8 starts for this shape (pipeline stages count finished CTAs).
inline int gemm_grid(int M, int N, int num_sms) {
    if (gemm_use_pair(M)) {
        const int ntiles = (M / (2 * BM)) * (N / BN);
        return 2 * (ntiles < num_sms / 2 ? ntiles : num_sms / 2);
    }
    const int ntiles = (M / BM) * (N / BN);
    return ntiles < num_sms ? ntiles : num_sms;
}

inline int launch_gemm_fp8(const uint8_t *X, const uint8_t *W, uint8_t *out, int M, int N, int K, float scale, const Gating &g, int num_sms, cudaStream_t stream) {
    if (M % BM || N % BN || K % BK) return 3;
    const bool pair = gemm_use_pair(M);
    Params P;
    if (!cached_tmap_copy(&P.tmA, X, (uint64_t)M, (uint64_t)K, BM) || !cached_tmap_copy(&P.tmB, W, (uint64_t)N, (uint64_t)K, pair ? BN / 2 : BN)) return 4;
    P.out = out; P.M = M; P.N = N; P.K = K; P.scale = scale; P.g = g;
    // ACX_GEMM_TMA_STORE: 1 = TMA-store epilogue for local outputs, 2 = also when the output is the next pipeline stage's input buffer (peer memory)
    static int tma_store_mode = -1;
    if (tma_store_mode < 0) { const char *e = getenv("ACX_GEMM_TMA_STORE"); tma_store_mode = e ? atoi(e) : ACX_GEMM_TMA_STORE_DEFAULT; }
    P.tma_store = (tma_store_mode >= 2 || (tma_store_mode == 1 && !g.out_arrived)) ? 1 : 0;
    memset(&P.tmC, 0, sizeof P.tmC);
    if (P.tma_store && !cached_tmap_copy(&P.tmC, out, (uint64_t)M, (uint64_t)N, BM)) return 4;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_fp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return 5;
        if (cudaFuncSetAttribute(gemm_fp8_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES2) != cudaSuccess) return 5;
        attr_set = true;
    }
    const int grid = gemm_grid(M, N, num_sms);
    if (pair) gemm_fp8_pair_kernel<<<grid, THREADS, SMEM_BYTES2, stream>>>(P);
    else gemm_fp8_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : 6;
}

}  // namespace
}  // namespace acxg
#endif
