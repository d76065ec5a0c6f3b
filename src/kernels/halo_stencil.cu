// Fused halo-exchange + 3-D 7-point stencil for sm_100a (BASELINE.json config #3), round 2: TMA-fed.
//
// There is nothing like this in the reference (it never touches payload).
// Each rank owns an nx*ny*nz bf16 block of a periodic px*py*pz decomposition, stored WITH its ghost shell:
//   row (y, z) = nx interior cells followed by 16 pad cells; pad[0] is the row's right ghost (x = nx) and pad[15] the
//   NEXT row's left ghost (x = -1), so every interior row starts on a 32-B sector and "8 cells before the row" is a plain
//   16-B vector; rows y = -1 and y = ny and planes z = -1 and z = nz are the y / z ghosts.
// One kernel per time step does everything:
//   * the field is read ONLY by TMA: cp.async.bulk.tensor.3d boxes of (whole padded row) x (TY + 2 rows) x (1 plane) —
//     one contiguous 10.5 KB run of memory each — into an 8-stage shared-memory ring with full/empty mbarriers; the
//     loads run 5 planes ahead of the math and straight through tile boundaries, so ~100 KB per SM are in flight at all
//     times without a single register holding a prefetched operand (round 1: 113 registers, latency-bound at 34 % of HBM);
//   * a tile is TY = 8 rows x all of x x TZ = 32 planes; thread (row pair, 8-cell vector) z-marches with the previous /
//     current centre vectors of its two rows in registers (each row is the other's y-neighbour), reads the row above /
//     below the pair and the next plane's two centres from shared memory (4 LDS.128 per 16 cells), x-neighbours by warp
//     shuffle (ghost / neighbouring-warp cells from shared memory), packed fp32 arithmetic (FADD2 / FFMA2);
//   * ghost cells are written by the NEIGHBOURS: a tile's face cells are stored twice, into the local output field and —
//     16-B vectors over NVLink — straight into the neighbour's ghost cells of ITS output field (peer-mapped symmetric
//     heap), followed by one release flag per (face, tile) partition (acx::pready_signal_epoch on the partitioned
//     channel); a tile's producer acquires the flags of the partitions its boxes cover (acx::wait_arrived_epoch, LOCAL
//     HBM) and issues fence.proxy.async before its first TMA load;
//   * tiles touching a y / z face are scheduled first so their NVLink stores overlap the interior sweep.
// The 12 + 12 partitioned channels are ordinary MPIX_Psend_init / MPIX_Precv_init requests (flag-only use: the
// producer stores the payload in place, the device API's zero-copy form); rounds advance on the device (explicit
// epochs), the two fields alternate, and the stencil's own data dependence is the flow control (see acx_halo_step).
// Host involvement per step: one kernel launch (capturable in a CUDA graph).
#include <mpi.h>
#include "../../include/mpi-acx.h"
#include "../../include/acx_kernels.h"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#ifndef ACX_HALO_TILE16_DEFAULT
#define ACX_HALO_TILE16_DEFAULT 1   // 16 x 16 tiles: 0.1434 vs 0.1533 ms per 512^3 step (profiles/raw/r2_call16)
#endif

namespace {

// tile: TY rows x all of x, z-march of TZ planes. Two geometries are compiled: 8 x 32 (8 compute warps + the producer warp: nine
// warps are allocated as twelve, so 108 registers leave room for ONE CTA per SM) and 16 x 16 (16 compute warps + the producer warp
// in one CTA per SM at 96 registers: twice the warps to hide the math's latency behind, ACX_HALO_TILE=16).
constexpr int STAGES = 8;           // shared-memory ring of planes (8 x 10.6 KB per CTA, two CTAs per SM)
constexpr int XPAD = 16;            // pad cells behind every row (see the layout comment above)

struct HaloDev {
    __nv_bfloat16 *un;              // interior origin (x = y = z = 0) of the local OUTPUT field
    __nv_bfloat16 *nb[6];           // peer-mapped interior origins of the neighbours' output fields (face order X- X+ Y- Y+ Z- Z+)
    int nx, ny, nz, nty, ntz, ntiles;
    long long rp, pp;               // row / plane pitch in cells
    const acx_pchan *send[6];       // flags: my face f -> neighbour in direction f
    const acx_pchan *recv[6];       // flags: ghost cells behind face f have been written
    unsigned long long wait_val, signal_val;
    float c0, c1;
    const int *order;
    const unsigned char *in_box0;   // != null (ACX_HALO_BULK=1): address of box (0, 0, 0) of the INPUT field — a box is one contiguous run of memory, so it can travel as a 1-D bulk copy instead of a 3-D tensor box
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "HW_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni HW_DONE;\n\t"
        "bra.uni HW_LOOP;\n\t"
        "HW_DONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(const void *src, uint64_t *bar, void *dst, uint32_t bytes) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap *tm, uint64_t *bar, void *dst, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// bf16x8 <-> four fp32 pairs. The stencil arithmetic runs on packed pairs (FADD2 / FMUL2 / FFMA2: two fp32 lanes per
// instruction on sm_100): the kernel is issue-bound, not memory-bound, once TMA keeps the planes coming.
__device__ __forceinline__ void unpack8(const uint4 &v, float2 f[4]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) f[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
}
__device__ __forceinline__ uint4 pack8(const float2 f[4]) {
    uint4 v;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __float22bfloat162_rn(f[i]);
    return v;
}
// out = c0*cur + c1*(((xm + xp) + (ym + yp)) + (zm + zp)) for the 8 cells of one vector; left / right are the x-neighbours
// of cells 0 and 7.
__device__ __forceinline__ uint4 stencil8(const float2 cur[4], const float2 ym[4], const float2 yp[4], const float2 zm[4], const float2 zp[4], float left, float right,
                                          float2 c0, float2 c1) {
    float2 t[4], o[4];
    t[0] = make_float2(left + cur[0].y, cur[0].x + cur[1].x);
    t[1] = make_float2(cur[0].y + cur[1].y, cur[1].x + cur[2].x);
    t[2] = make_float2(cur[1].y + cur[2].y, cur[2].x + cur[3].x);
    t[3] = make_float2(cur[2].y + cur[3].y, cur[3].x + right);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 s = __fadd2_rn(__fadd2_rn(t[i], __fadd2_rn(ym[i], yp[i])), __fadd2_rn(zm[i], zp[i]));
        o[i] = __ffma2_rn(c1, s, __fmul2_rn(c0, cur[i]));
    }
    return pack8(o);
}

// Face ids: 0 X-, 1 X+, 2 Y-, 3 Y+, 4 Z-, 5 Z+. Partition index of tile (ty, tc): x faces ty + nty * tc (every tile
// touches both), y faces tc (tiles with ty == 0 / nty-1), z faces ty (tiles with tc == 0 / ntz-1).
// Thread (rp, vx) owns rows 2*rp and 2*rp+1 of the tile's TY rows and the 8 cells vx*8 .. vx*8+7: two vectors per plane,
// each the other's y-neighbour (registers), so a plane costs 4 LDS.128 per thread: the two next-plane centres, the row
// above the pair and the row below it.
template <int TY, int TZ>
__device__ __forceinline__ void halo_stencil_body(const CUtensorMap &tm, const HaloDev &P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int ncw = (int)(blockDim.x >> 5) - 1;                  // compute warps; the last warp is the producer
    const bool producer = (tid >> 5) == ncw;
    const int vpr = P.nx >> 3;                                   // 16-B vectors per row
    const int rp2 = tid / vpr, vx = tid - rp2 * vpr;             // row pair, vector (compute threads)
    const uint32_t row_bytes = (uint32_t)(P.nx + XPAD) * 2u;     // one padded row
    const uint32_t stage_bytes = row_bytes * (TY + 2);           // what one box delivers
    const uint32_t stage_stride = (stage_bytes + 127u) & ~127u;  // TMA destinations are 128-B aligned
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * stage_stride);
    uint64_t *empty = full + STAGES;
    uint64_t *tile_done = empty + STAGES;                        // one per parity of the tile counter
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], (uint32_t)ncw); }
        mbar_init(&tile_done[0], (uint32_t)ncw); mbar_init(&tile_done[1], (uint32_t)ncw);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (producer) {
        // ---- producer warp (one lane): an uninterrupted stream of plane loads, straight through tile boundaries, as far ahead
        // of the math as the ring allows; between loads it publishes the face flags of tiles the compute warps have finished ----
        if (lane != 0) return;
        uint32_t p_stage = 0, p_par = 1;       // ring position of the next load; parity to wait for on its empty barrier (1 on the first pass: passes at once)
        uint32_t sig_no = 0;                   // tiles whose flags have been published
        int sig_it = blockIdx.x;
        auto try_signal = [&](bool block) {
            // tile number sig_no is complete when every compute warp has arrived on its tile barrier
            const uint32_t par = (sig_no >> 1) & 1u;
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&tile_done[sig_no & 1])), "r"(par) : "memory");
            if (!ok) { if (!block) return false; mbar_wait(&tile_done[sig_no & 1], par); }
            const int t = P.order[sig_it];
            const int ty = t % P.nty, tc = t / P.nty;
            __threadfence_system();            // ONE system-scope fence, then relaxed flag stores (fence + strong store = release)
            acx::st_relaxed_sys(P.send[0]->arrived + t, P.signal_val);
            acx::st_relaxed_sys(P.send[1]->arrived + t, P.signal_val);
            if (ty == 0) acx::st_relaxed_sys(P.send[2]->arrived + tc, P.signal_val);
            if (ty == P.nty - 1) acx::st_relaxed_sys(P.send[3]->arrived + tc, P.signal_val);
            if (tc == 0) acx::st_relaxed_sys(P.send[4]->arrived + ty, P.signal_val);
            if (tc == P.ntz - 1) acx::st_relaxed_sys(P.send[5]->arrived + ty, P.signal_val);
            sig_no++; sig_it += gridDim.x;
            return true;
        };
        uint32_t loaded_tiles = 0;
        for (int it = blockIdx.x; it < P.ntiles; it += gridDim.x, loaded_tiles++) {
            const int t = P.order[it];
            const int ty = t % P.nty, tc = t / P.nty;
            // the ghost cells this tile's boxes cover were written by the neighbours' previous step: poll their flags (relaxed),
            // then ONE acquire fence for all of them and the generic-to-async proxy fence in front of the TMA loads
            auto poll = [&](const acx_pchan *ch, int part) { const uint64_t *f = ch->arrived + part; unsigned ns = 0; while (acx::ld_relaxed_sys(f) < P.wait_val) { __nanosleep(ns); ns = ns < 256 ? ns * 2 + 20 : 256; } };
            poll(P.recv[0], t); poll(P.recv[1], t);
            if (ty == 0) poll(P.recv[2], tc);
            if (ty == P.nty - 1) poll(P.recv[3], tc);
            if (tc == 0) poll(P.recv[4], ty);
            if (tc == P.ntz - 1) poll(P.recv[5], ty);
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes (remote stores, made visible by the acquire) before async-proxy reads
            for (int q = 0; q < TZ + 2; q++) {
                mbar_wait(&empty[p_stage], p_par);                 // every compute warp is done with the stage's previous plane
                mbar_expect_tx(&full[p_stage], stage_bytes);
                // box origin in the padded tensor: u64 column 0 (= 16 B before the interior row), padded row ty*TY, padded plane tc*TZ + q
                if (P.in_box0) bulk_load_1d(P.in_box0 + ((long long)(tc * TZ + q) * P.pp + (long long)(ty * TY) * P.rp) * 2, &full[p_stage], smem + (size_t)p_stage * stage_stride, stage_bytes);
                else tma_load_3d(&tm, &full[p_stage], smem + (size_t)p_stage * stage_stride, 0, ty * TY, tc * TZ + q);
                if (++p_stage == STAGES) { p_stage = 0; p_par ^= 1u; }
                if ((q & 7) == 7 && sig_no < loaded_tiles) try_signal(false);
            }
        }
        while (sig_no < loaded_tiles) try_signal(true);
        return;
    }

    // ---- compute warps ----
    uint32_t c_stage = 0, c_par = 0;           // ring position / phase of the plane being consumed
    auto advance = [&](uint32_t &st, uint32_t &par) { if (++st == STAGES) { st = 0; par ^= 1u; } };
    auto release = [&](uint32_t st) { __syncwarp(); if (lane == 0) mbar_arrive(&empty[st]); };
    uint32_t tile_no = 0;
    const float2 c0v = make_float2(P.c0, P.c0), c1v = make_float2(P.c1, P.c1);
    for (int it = blockIdx.x; it < P.ntiles; it += gridDim.x, tile_no++) {
        const int t = P.order[it];
        const int ty = t % P.nty, tc = t / P.nty;
        const int ya = ty * TY + 2 * rp2, z0 = tc * TZ, x0 = vx * 8;
        const uint32_t offA = (uint32_t)(2 * rp2 + 1) * row_bytes + 16u + (uint32_t)x0 * 2u;   // row A's vector inside a stage; row B is one row further
        const uint32_t offB = offA + row_bytes;

        float2 prevA[4], curA[4], prevB[4], curB[4];
        mbar_wait(&full[c_stage], c_par);                         // plane z0-1: only its centre rows are needed
        {
            const unsigned char *s = smem + (size_t)c_stage * stage_stride;
            unpack8(*reinterpret_cast<const uint4 *>(s + offA), prevA);
            unpack8(*reinterpret_cast<const uint4 *>(s + offB), prevB);
        }
        release(c_stage);
        advance(c_stage, c_par);
        mbar_wait(&full[c_stage], c_par);                         // plane z0
        {
            const unsigned char *s = smem + (size_t)c_stage * stage_stride;
            unpack8(*reinterpret_cast<const uint4 *>(s + offA), curA);
            unpack8(*reinterpret_cast<const uint4 *>(s + offB), curB);
        }
        // which faces this thread's two rows / vector touch in this tile (most threads: none)
        const uint32_t fmask = (vx == 0 ? 1u : 0u) | (vx == vpr - 1 ? 2u : 0u) | (ya == 0 ? 4u : 0u) | (ya + 1 == P.ny - 1 ? 8u : 0u) | (tc == 0 ? 16u : 0u) | (tc == P.ntz - 1 ? 32u : 0u);
        long long off = (long long)z0 * P.pp + (long long)ya * P.rp;   // row A of the current plane, relative to an interior origin
        uint32_t n_stage = c_stage, n_par = c_par;                 // ring position of the NEXT plane
        advance(n_stage, n_par);
#pragma unroll 3   // the prev / cur / next register roles rotate with period 3: unrolled, the rotation is renaming instead of 32 moves per plane
        for (int zl = 0; zl < TZ; zl++) {
            mbar_wait(&full[n_stage], n_par);
            const unsigned char *sc = smem + (size_t)c_stage * stage_stride, *sn = smem + (size_t)n_stage * stage_stride;
            const uint4 vnA = *reinterpret_cast<const uint4 *>(sn + offA), vnB = *reinterpret_cast<const uint4 *>(sn + offB);
            const uint4 vup = *reinterpret_cast<const uint4 *>(sc + offA - row_bytes), vdn = *reinterpret_cast<const uint4 *>(sc + offB + row_bytes);
            float lA = __shfl_up_sync(0xffffffffu, curA[3].y, 1), rA = __shfl_down_sync(0xffffffffu, curA[0].x, 1);
            float lB = __shfl_up_sync(0xffffffffu, curB[3].y, 1), rB = __shfl_down_sync(0xffffffffu, curB[0].x, 1);
            if (lane == 0) {    // cell x0-1: the left ghost (vx == 0) or the neighbouring warp's last cell
                lA = __bfloat162float(*reinterpret_cast<const __nv_bfloat16 *>(sc + offA - 2));
                lB = __bfloat162float(*reinterpret_cast<const __nv_bfloat16 *>(sc + offB - 2));
            }
            if (lane == 31) {   // cell x0+8: the right ghost (last vector) or the neighbouring warp's first cell
                rA = __bfloat162float(*reinterpret_cast<const __nv_bfloat16 *>(sc + offA + 16));
                rB = __bfloat162float(*reinterpret_cast<const __nv_bfloat16 *>(sc + offB + 16));
            }
            float2 nxtA[4], nxtB[4], edge[4];
            unpack8(vnA, nxtA); unpack8(vup, edge);
            const uint4 voA = stencil8(curA, edge, curB, prevA, nxtA, lA, rA, c0v, c1v);
            unpack8(vnB, nxtB); unpack8(vdn, edge);
            const uint4 voB = stencil8(curB, curA, edge, prevB, nxtB, lB, rB, c0v, c1v);
            {
                __nv_bfloat16 *o = P.un + off + x0;
                *reinterpret_cast<uint4 *>(o) = voA;
                *reinterpret_cast<uint4 *>(o + P.rp) = voB;
            }
            // face cells also go straight into the neighbours' ghost cells (NVLink stores; self for a periodic dimension of extent 1)
            if (fmask) {
                if (fmask & 1u) { __nv_bfloat16 *g = P.nb[0] + off + P.nx; *reinterpret_cast<uint4 *>(g) = voA; *reinterpret_cast<uint4 *>(g + P.rp) = voB; }   // their right ghost vector [nx, nx+8): cell 0 is the one they read
                if (fmask & 2u) { __nv_bfloat16 *g = P.nb[1] + off - 8; *reinterpret_cast<uint4 *>(g) = voA; *reinterpret_cast<uint4 *>(g + P.rp) = voB; }      // their left ghost vector [-8, 0): cell 7 is x = -1
                if (fmask & 4u) *reinterpret_cast<uint4 *>(P.nb[2] + off + (long long)P.ny * P.rp + x0) = voA;                 // row A is y = 0 -> their row y = ny
                if (fmask & 8u) *reinterpret_cast<uint4 *>(P.nb[3] + off + P.rp - (long long)P.ny * P.rp + x0) = voB;          // row B is y = ny-1 -> their row y = -1
                if ((fmask & 16u) && zl == 0) { __nv_bfloat16 *g = P.nb[4] + off + (long long)P.nz * P.pp + x0; *reinterpret_cast<uint4 *>(g) = voA; *reinterpret_cast<uint4 *>(g + P.rp) = voB; }        // plane z = 0 -> their plane z = nz
                if ((fmask & 32u) && zl == TZ - 1) { __nv_bfloat16 *g = P.nb[5] + off - (long long)P.nz * P.pp + x0; *reinterpret_cast<uint4 *>(g) = voA; *reinterpret_cast<uint4 *>(g + P.rp) = voB; }   // plane z = nz-1 -> their plane z = -1
            }
#pragma unroll
            for (int i = 0; i < 4; i++) { prevA[i] = curA[i]; curA[i] = nxtA[i]; prevB[i] = curB[i]; curB[i] = nxtB[i]; }
            off += P.pp;
            release(c_stage);                                      // rows above / below of the current plane have been read
            c_stage = n_stage; c_par = n_par;
            advance(n_stage, n_par);
        }
        release(c_stage);                                          // plane z0+TZ: centres only
        advance(c_stage, c_par);
        // the tile's stores are done: tell the producer warp, which publishes the face flags (release / acquire at CTA scope
        // through the barrier, one system-scope fence on its side)
        __syncwarp();
        if (lane == 0) mbar_arrive(&tile_done[tile_no & 1]);
    }
}

__global__ void __maxnreg__(112) halo_stencil_tma(const __grid_constant__ CUtensorMap tm, HaloDev P) { halo_stencil_body<8, 32>(tm, P); }
__global__ void __maxnreg__(96) halo_stencil_tma16(const __grid_constant__ CUtensorMap tm, HaloDev P) { halo_stencil_body<16, 16>(tm, P); }

// Initial exchange: faces of u^0 -> the neighbours' ghost cells of the SAME field, flag value 1 on the parity-0 channels.
// One CTA per (face, partition).
struct PackDev {
    const __nv_bfloat16 *u;         // interior origin of the local current field
    __nv_bfloat16 *nb[6];           // neighbours' interior origins of THEIR current field
    int nx, ny, nz, nty, ntz, TY, TZ;
    long long rp, pp;
    const acx_pchan *send[6];
    unsigned long long signal_val;
};
__global__ void halo_pack_faces(PackDev P) {
    const int TY = P.TY, TZ = P.TZ;
    const int f = blockIdx.y, p = blockIdx.x;
    const int nparts = f < 2 ? P.nty * P.ntz : f < 4 ? P.ntz : P.nty;
    if (p >= nparts) return;
    const int vpr = P.nx / 8;
    if (f < 2) {            // x faces: partition = tile (ty, tc); one 16-B vector per (row, plane)
        const int ty = p % P.nty, tc = p / P.nty;
        for (int i = threadIdx.x; i < TY * TZ; i += blockDim.x) {
            const int yl = i % TY, zl = i / TY;
            const long long off = (long long)(tc * TZ + zl) * P.pp + (long long)(ty * TY + yl) * P.rp;
            if (f == 0) *reinterpret_cast<uint4 *>(P.nb[0] + off + P.nx) = *reinterpret_cast<const uint4 *>(P.u + off);
            else *reinterpret_cast<uint4 *>(P.nb[1] + off - 8) = *reinterpret_cast<const uint4 *>(P.u + off + P.nx - 8);
        }
    } else if (f < 4) {     // y faces: partition = z-chunk tc; row y = 0 (f == 2) or ny-1 of TZ planes
        const int tc = p, y = f == 2 ? 0 : P.ny - 1;
        for (int i = threadIdx.x; i < TZ * vpr; i += blockDim.x) {
            const int zl = i / vpr, v = i % vpr;
            const long long off = (long long)(tc * TZ + zl) * P.pp + (long long)y * P.rp + v * 8;
            const long long shift = f == 2 ? (long long)P.ny * P.rp : -(long long)P.ny * P.rp;
            *reinterpret_cast<uint4 *>(P.nb[f] + off + shift) = *reinterpret_cast<const uint4 *>(P.u + off);
        }
    } else {                // z faces: partition = strip ty; plane z = 0 (f == 4) or nz-1, TY rows
        const int ty = p, z = f == 4 ? 0 : P.nz - 1;
        for (int i = threadIdx.x; i < TY * vpr; i += blockDim.x) {
            const int yl = i / vpr, v = i % vpr;
            const long long off = (long long)z * P.pp + (long long)(ty * TY + yl) * P.rp + v * 8;
            const long long shift = f == 4 ? (long long)P.nz * P.pp : -(long long)P.nz * P.pp;
            *reinterpret_cast<uint4 *>(P.nb[f] + off + shift) = *reinterpret_cast<const uint4 *>(P.u + off);
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) acx::pready_signal_epoch(p, P.send[f], P.signal_val);
}

// dense [nz][ny][nx] <-> the padded field (import / export for callers that keep a plain array)
__global__ void halo_copy_dense(__nv_bfloat16 *padded, __nv_bfloat16 *dense, int nx, int ny, int nz, long long rp, long long pp, int to_padded) {
    const long long nvec = (long long)nx / 8 * ny * nz;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % (nx / 8));
        const long long r = i / (nx / 8);
        const int y = (int)(r % ny), z = (int)(r / ny);
        uint4 *a = reinterpret_cast<uint4 *>(padded + (long long)z * pp + (long long)y * rp + v * 8);
        uint4 *b = reinterpret_cast<uint4 *>(dense + ((long long)z * ny + y) * nx + v * 8);
        if (to_padded) *a = *b; else *b = *a;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    }
    return fn;
}

}  // namespace

struct acx_halo {
    int nx, ny, nz, px, py, pz, rank, size;
    int nty, ntz, ntiles, ty, tz;  // ty x tz: tile geometry (rows, planes)
    long long rp, pp;              // pitches in cells
    size_t field_bytes;
    float c0, c1;
    char *alloc[2];                // symmetric-heap allocations
    __nv_bfloat16 *u[2];           // interior origins
    __nv_bfloat16 *nbu[2][6];      // neighbours' interior origins (peer-mapped), per field and face
    CUtensorMap tm[2];
    int cur;                       // u[cur] holds u^step
    unsigned long long step;       // number of completed steps
    bool exchanged;
    void *land[2][6];              // flag-only channels: tiny receive buffers the partitioned API wants (symmetric heap)
    MPIX_Request sreq[2][6], rreq[2][6];
    MPIX_Prequest spre[2][6], rpre[2][6];
    int *order_d;
    int grid, threads, bulk;
    size_t smem;
};

#define HCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "[acx halo] %s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define HOK(x) do { int r_ = (x); if (r_) { fprintf(stderr, "[acx halo] %s:%d %s -> %d\n", __FILE__, __LINE__, #x, r_); return 1; } } while (0)

extern "C" int acx_halo_create(int nx, int ny, int nz, int px, int py, int pz, float c0, float c1, acx_halo **out) {
    int rank, size;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    if (px * py * pz != size) { fprintf(stderr, "[acx halo] decomposition %dx%dx%d != %d ranks\n", px, py, pz, size); return 1; }
    const char *tile_env = getenv("ACX_HALO_TILE");
    int tile16 = tile_env ? atoi(tile_env) == 16 : ACX_HALO_TILE16_DEFAULT;
    if (tile16 && (ny % 16 || nz % 16) && ny % 8 == 0 && nz % 32 == 0) tile16 = 0;        // the block only fits the other geometry
    else if (!tile16 && (ny % 8 || nz % 32) && ny % 16 == 0 && nz % 16 == 0) tile16 = 1;
    const int TY = tile16 ? 16 : 8, TZ = tile16 ? 16 : 32;
    if ((nx != 256 && nx != 512) || ny % TY || nz % TZ) { fprintf(stderr, "[acx halo] block %dx%dx%d: nx must be 256 or 512 (one thread per 8 cells, %d rows per CTA), ny a multiple of %d, nz of %d\n", nx, ny, nz, TY, TY, TZ); return 1; }
    acx_halo *h = new acx_halo();
    h->ty = TY; h->tz = TZ;
    h->nx = nx; h->ny = ny; h->nz = nz; h->px = px; h->py = py; h->pz = pz; h->rank = rank; h->size = size; h->c0 = c0; h->c1 = c1;
    h->nty = ny / TY; h->ntz = nz / TZ; h->ntiles = h->nty * h->ntz;
    h->rp = nx + XPAD; h->pp = h->rp * (ny + 2);
    h->cur = 0; h->step = 0; h->exchanged = false;
    // padded field = (nz+2) planes of (ny+2) rows, plus one lead-in and one tail row (the tensor map starts 16 B early)
    h->field_bytes = ((size_t)h->pp * (nz + 2) + 2 * (size_t)h->rp) * 2;
    const int cx = rank % px, cy = (rank / px) % py, cz = rank / (px * py);
    auto rk = [&](int x, int y, int z) { return (((z + pz) % pz) * py + ((y + py) % py)) * px + ((x + px) % px); };
    const int nbr[6] = {rk(cx - 1, cy, cz), rk(cx + 1, cy, cz), rk(cx, cy - 1, cz), rk(cx, cy + 1, cz), rk(cx, cy, cz - 1), rk(cx, cy, cz + 1)};
    EncodeTiledFn enc = encode_fn();
    if (!enc) { fprintf(stderr, "[acx halo] cuTensorMapEncodeTiled is unavailable\n"); return 1; }
    for (int b = 0; b < 2; b++) {
        void *p = nullptr;
        if (MPIX_Alloc_mem(h->field_bytes, MPI_INFO_NULL, &p)) {
            fprintf(stderr, "[acx halo] the two padded fields need 2 x %.0f MiB of the symmetric heap (neighbours write ghost cells in place): raise ACX_HEAP_BYTES\n", h->field_bytes / 1048576.0);
            return 1;
        }
        h->alloc[b] = (char *)p;
        HCK(cudaMemset(p, 0, h->field_bytes));
        // interior origin: skip the lead-in row, the z = -1 plane and the y = -1 row
        h->u[b] = (__nv_bfloat16 *)(h->alloc[b]) + h->rp + h->pp + h->rp;
        for (int f = 0; f < 6; f++) {
            void *q = nullptr;
            HOK(MPIX_Peer_ptr(h->u[b], nbr[f], &q));
            h->nbu[b][f] = (__nv_bfloat16 *)q;
        }
        // 3-D tensor map over 8-byte elements (4 cells): {row, rows of a plane, planes}. Its origin sits 16 B before padded
        // row 0 of padded plane 0, so a box that starts at column 0 begins with the row's left ghost vector.
        char *base = h->alloc[b] + (size_t)h->rp * 2 - 16;
        cuuint64_t dims[3] = {(cuuint64_t)(h->rp / 4), (cuuint64_t)(ny + 2), (cuuint64_t)(nz + 2)};
        cuuint64_t strides[2] = {(cuuint64_t)h->rp * 2, (cuuint64_t)h->pp * 2};
        cuuint32_t box[3] = {(cuuint32_t)(h->rp / 4), (cuuint32_t)(h->ty + 2), 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&h->tm[b], CU_TENSOR_MAP_DATA_TYPE_INT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { fprintf(stderr, "[acx halo] cuTensorMapEncodeTiled failed (%d)\n", (int)r); return 1; }
    }
    const int nparts[6] = {h->ntiles, h->ntiles, h->ntz, h->ntz, h->nty, h->nty};
    for (int par = 0; par < 2; par++)
        for (int f = 0; f < 6; f++) {
            // flag-only channels: the payload is stored in place by the producer (zero-copy form of the device API), the
            // partitioned request carries the per-partition arrival flags; 16 bytes per partition keep the API honest
            HOK(MPIX_Alloc_mem((size_t)nparts[f] * 16, MPI_INFO_NULL, &h->land[par][f]));
            HCK(cudaMemset(h->land[par][f], 0, (size_t)nparts[f] * 16));
            // My face f travels to nbr[f] and lands behind ITS face f^1: tag by the sender's face id.
            const int tag_send = 100 + par * 8 + f, tag_recv = 100 + par * 8 + (f ^ 1);
            HOK(MPIX_Psend_init(nullptr, nparts[f], (MPI_Count)16, MPI_BYTE, nbr[f], tag_send, MPI_COMM_WORLD, MPI_INFO_NULL, &h->sreq[par][f]));
            HOK(MPIX_Precv_init(h->land[par][f], nparts[f], (MPI_Count)16, MPI_BYTE, nbr[f], tag_recv, MPI_COMM_WORLD, MPI_INFO_NULL, &h->rreq[par][f]));
        }
    for (int par = 0; par < 2; par++)
        for (int f = 0; f < 6; f++) {
            HOK(MPIX_Prequest_create(h->sreq[par][f], &h->spre[par][f]));
            HOK(MPIX_Prequest_create(h->rreq[par][f], &h->rpre[par][f]));
        }
    // tiles on a y / z face first: their NVLink stores overlap the interior sweep
    std::vector<int> order(h->ntiles);
    for (int i = 0; i < h->ntiles; i++) order[i] = i;
    auto nfaces = [&](int t) { int ty = t % h->nty, tc = t / h->nty; return (ty == 0) + (ty == h->nty - 1) + (tc == 0) + (tc == h->ntz - 1); };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nfaces(a) > nfaces(b); });
    HCK(cudaMalloc((void **)&h->order_d, sizeof(int) * h->ntiles));
    HCK(cudaMemcpy(h->order_d, order.data(), sizeof(int) * h->ntiles, cudaMemcpyHostToDevice));
    h->threads = nx / 8 * (h->ty / 2) + 32;   // two rows per thread, plus the producer warp
    h->smem = (size_t)STAGES * ((((size_t)(h->ty + 2) * (size_t)h->rp * 2) + 127) & ~(size_t)127) + (2 * STAGES + 2) * 8 + 16;
    HCK(cudaFuncSetAttribute(h->ty == 16 ? halo_stencil_tma16 : halo_stencil_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->ty == 16 ? halo_stencil_tma16 : halo_stencil_tma, h->threads, h->smem);
    if (occ < 1) occ = 1;
    h->bulk = getenv("ACX_HALO_BULK") && atoi(getenv("ACX_HALO_BULK")) != 0;
    const char *g = getenv("ACX_HALO_CTAS_PER_SM");
    if (g && atoi(g) > 0) occ = std::min(occ, atoi(g));
    // equal-size tiles: a grid that does not divide them leaves most CTAs idle during the last wave (1024 tiles on 296 CTAs:
    // four waves, the last one a quarter full). Use the fewest waves the resident CTAs allow and spread the tiles evenly.
    const int slots = std::max(1, sms * occ), waves = (h->ntiles + slots - 1) / slots;
    h->grid = std::min(h->ntiles, (h->ntiles + waves - 1) / waves);
    MPI_Barrier(MPI_COMM_WORLD);
    *out = h;
    return 0;
}

extern "C" void *acx_halo_field(acx_halo *h) { return h->u[h->cur]; }
extern "C" void acx_halo_field_pitch(acx_halo *h, long long *row_pitch_cells, long long *plane_pitch_cells) { *row_pitch_cells = h->rp; *plane_pitch_cells = h->pp; }
extern "C" unsigned long long acx_halo_steps_done(acx_halo *h) { return h->step; }
extern "C" int acx_halo_import_dense(acx_halo *h, const void *dense, cudaStream_t stream) {
    halo_copy_dense<<<1184, 256, 0, stream>>>(h->u[h->cur], (__nv_bfloat16 *)dense, h->nx, h->ny, h->nz, h->rp, h->pp, 1);
    HCK(cudaGetLastError());
    return 0;
}
extern "C" int acx_halo_export_dense(acx_halo *h, void *dense, cudaStream_t stream) {
    halo_copy_dense<<<1184, 256, 0, stream>>>(h->u[h->cur], (__nv_bfloat16 *)dense, h->nx, h->ny, h->nz, h->rp, h->pp, 0);
    HCK(cudaGetLastError());
    return 0;
}

extern "C" int acx_halo_exchange_initial(acx_halo *h, cudaStream_t stream) {
    PackDev P;
    P.u = h->u[h->cur]; P.nx = h->nx; P.ny = h->ny; P.nz = h->nz; P.nty = h->nty; P.ntz = h->ntz; P.TY = h->ty; P.TZ = h->tz; P.rp = h->rp; P.pp = h->pp;
    const int par = (int)(h->step & 1);
    for (int f = 0; f < 6; f++) { P.send[f] = (const acx_pchan *)h->spre[par][f]; P.nb[f] = h->nbu[h->cur][f]; }
    P.signal_val = h->step + 1;
    halo_pack_faces<<<dim3(h->ntiles, 6), 256, 0, stream>>>(P);
    HCK(cudaGetLastError());
    h->exchanged = true;
    return 0;
}

// Flow control without credits: step s reads U[s&1] (ghosts included) and writes U[(s+1)&1], mine and — for face cells —
// the neighbours'. A neighbour's ghost cells of U[(s+1)&1] were last read by its step s-1, by exactly the tile whose
// face flag (epoch s+1, parity s&1) this rank's mirror tile acquires before it computes anything in step s: when the
// flag is visible that tile has finished, loads included. So nothing is overwritten while still being read.
extern "C" int acx_halo_step(acx_halo *h, int nsteps, cudaStream_t stream) {
    if (!h->exchanged) HOK(acx_halo_exchange_initial(h, stream));
    for (int s = 0; s < nsteps; s++) {
        HaloDev P;
        const int cur = h->cur, nxt = cur ^ 1;
        P.un = h->u[nxt];
        for (int f = 0; f < 6; f++) P.nb[f] = h->nbu[nxt][f];
        P.nx = h->nx; P.ny = h->ny; P.nz = h->nz; P.nty = h->nty; P.ntz = h->ntz; P.ntiles = h->ntiles; P.rp = h->rp; P.pp = h->pp;
        const int rp = (int)(h->step & 1), sp = rp ^ 1;
        for (int f = 0; f < 6; f++) { P.recv[f] = (const acx_pchan *)h->rpre[rp][f]; P.send[f] = (const acx_pchan *)h->spre[sp][f]; }
        P.wait_val = h->step + 1; P.signal_val = h->step + 2;
        P.c0 = h->c0; P.c1 = h->c1; P.order = h->order_d;
        P.in_box0 = h->bulk ? (const unsigned char *)h->alloc[cur] + (size_t)h->rp * 2 - 16 : nullptr;
        if (h->ty == 16) halo_stencil_tma16<<<h->grid, h->threads, h->smem, stream>>>(h->tm[cur], P);
        else halo_stencil_tma<<<h->grid, h->threads, h->smem, stream>>>(h->tm[cur], P);
        HCK(cudaGetLastError());
        h->cur ^= 1;
        h->step++;
    }
    return 0;
}

extern "C" int acx_halo_destroy(acx_halo *h) {
    cudaDeviceSynchronize();
    MPI_Barrier(MPI_COMM_WORLD);
    for (int par = 0; par < 2; par++)
        for (int f = 0; f < 6; f++) {
            MPIX_Prequest_free(&h->spre[par][f]); MPIX_Prequest_free(&h->rpre[par][f]);
            MPIX_Request_free(&h->sreq[par][f]); MPIX_Request_free(&h->rreq[par][f]);
            MPIX_Free_mem(h->land[par][f]);
        }
    MPIX_Free_mem(h->alloc[0]); MPIX_Free_mem(h->alloc[1]);
    cudaFree(h->order_d);
    delete h;
    return 0;
}
