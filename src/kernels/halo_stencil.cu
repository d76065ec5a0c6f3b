// Fused halo-exchange + 3-D 7-point stencil for sm_100a (BASELINE.json config #3).
//
// There is nothing like this in the reference (it never touches payload).
// Each rank owns an nx*ny*nz bf16 block of a periodic px*py*pz decomposition. One kernel
// per time step does everything:
//   * tiles on a block face first wait (ld.acquire on LOCAL HBM flags) for the neighbour's
//     face partition of the previous step, which the neighbour's kernel stored straight
//     into this GPU's landing buffers over NVLink;
//   * every tile z-marches through its 256x16x32 cells (16-B vector loads, fp32 math,
//     y-neighbours through shared memory, x-neighbours through warp shuffles);
//   * face cells are stored twice: into the local output array and — packed, 16-B vectors —
//     into the neighbour's landing buffer (peer-mapped partitioned channel), followed by
//     a release flag per partition (acx::pready_signal_epoch);
//   * boundary tiles are scheduled first so the NVLink traffic overlaps the interior.
// The 12+12 partitioned channels are ordinary MPIX_Psend_init/Precv_init requests on
// symmetric-heap buffers; rounds advance on the device (explicit epochs), double-buffered.
// Host involvement per step: one kernel launch (capturable in a CUDA graph).
#include <mpi.h>
#include "../../include/mpi-acx.h"
#include "../../include/acx_kernels.h"
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

namespace {

constexpr int TX = 256, TY = 16, TZ = 32;   // tile; 32 lanes x 8 cells along x, 16 rows, z-march of 32
constexpr int LX = 32;

struct HaloDev {
    const __nv_bfloat16 *u;
    __nv_bfloat16 *un;
    int nx, ny, nz, ntx, nty, ntz;
    const acx_pchan *send[6];   // my face f -> neighbour in direction f (buffers of parity (s+1)&1)
    const acx_pchan *recv[6];   // halo from direction f (parity s&1)
    unsigned long long wait_val, signal_val;
    float c0, c1;
    const int *order;
    int ntiles;
};

__device__ __forceinline__ void unpack8(const uint4 &v, float f[8]) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float f[8]) {
    uint4 v;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
}
__device__ __forceinline__ void unpack8x2(const uint4 &v, float2 f[4]) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) f[i] = __bfloat1622float2(h[i]);
}
__device__ __forceinline__ uint4 pack8x2(const float2 f[4]) {
    uint4 v;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __float22bfloat162_rn(f[i]);
    return v;
}
__device__ __forceinline__ float bf(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

// Face ids: 0 X-, 1 X+, 2 Y-, 3 Y+, 4 Z-, 5 Z+.
// PD = software-pipeline depth of the z-march (planes in flight per thread); MINB = CTAs per SM the
// register allocation is tuned for. Variants are selected at run time (ACX_HALO_VARIANT).
template <int PD, int MINB>
__global__ void __launch_bounds__(LX * TY, MINB) halo_stencil_step(HaloDev P) {
    static_assert(TZ % PD == 0, "TZ must be a multiple of the prefetch depth");
    __shared__ uint4 rows[2][TY][LX];                 // current plane, double-buffered (16 KiB)
    __shared__ __align__(16) __nv_bfloat16 xface[2][TZ * TY];  // packed X-/X+ face of this tile (2 KiB)
    const int lx = threadIdx.x, ly = threadIdx.y;
    const int tid = ly * LX + lx;
    const size_t plane = (size_t)P.nx * P.ny;
    for (int it = blockIdx.x; it < P.ntiles; it += gridDim.x) {
        const int t = P.order[it];
        const int tx = t % P.ntx, ty = (t / P.ntx) % P.nty, tz = t / (P.ntx * P.nty);
        const bool fxm = tx == 0, fxp = tx == P.ntx - 1, fym = ty == 0, fyp = ty == P.nty - 1, fzm = tz == 0, fzp = tz == P.ntz - 1;
        const int px_ = tz * P.nty + ty, py_ = tz * P.ntx + tx, pz_ = ty * P.ntx + tx;   // partition index per face family
        if (tid == 0) {
            if (fxm) acx::wait_arrived_epoch(P.recv[0], px_, P.wait_val);
            if (fxp) acx::wait_arrived_epoch(P.recv[1], px_, P.wait_val);
            if (fym) acx::wait_arrived_epoch(P.recv[2], py_, P.wait_val);
            if (fyp) acx::wait_arrived_epoch(P.recv[3], py_, P.wait_val);
            if (fzm) acx::wait_arrived_epoch(P.recv[4], pz_, P.wait_val);
            if (fzp) acx::wait_arrived_epoch(P.recv[5], pz_, P.wait_val);
        }
        __syncthreads();
        const int x0 = tx * TX + lx * 8, y = ty * TY + ly, z0 = tz * TZ;
        // landing buffers (local) and remote partitions for this tile
        const __nv_bfloat16 *hxm = fxm ? (const __nv_bfloat16 *)acx::local_partition_ptr(P.recv[0], px_) : nullptr;
        const __nv_bfloat16 *hxp = fxp ? (const __nv_bfloat16 *)acx::local_partition_ptr(P.recv[1], px_) : nullptr;
        const uint4 *hym = fym ? (const uint4 *)acx::local_partition_ptr(P.recv[2], py_) : nullptr;
        const uint4 *hyp = fyp ? (const uint4 *)acx::local_partition_ptr(P.recv[3], py_) : nullptr;
        const uint4 *hzm = fzm ? (const uint4 *)acx::local_partition_ptr(P.recv[4], pz_) : nullptr;
        const uint4 *hzp = fzp ? (const uint4 *)acx::local_partition_ptr(P.recv[5], pz_) : nullptr;
        uint4 *rym = fym ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[2], py_) : nullptr;
        uint4 *ryp = fyp ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[3], py_) : nullptr;
        uint4 *rzm = fzm ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[4], pz_) : nullptr;
        uint4 *rzp = fzp ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[5], pz_) : nullptr;

        const __nv_bfloat16 *col = P.u + (size_t)y * P.nx + x0;   // + z*plane
        auto load_plane = [&](int z) -> uint4 {
            if (z < 0) return hzm[ly * LX + lx];
            if (z >= P.nz) return hzp[ly * LX + lx];
            return *reinterpret_cast<const uint4 *>(col + (size_t)z * plane);
        };
        // Edge operands come from global memory / landing buffers instead of shared memory; they are
        // software-pipelined PD planes ahead exactly like the centre column, so no thread ever issues a
        // load in the same iteration that consumes it.
        const bool yedge = ly == 0 || ly == TY - 1, xedge = lx == 0 || lx == LX - 1;
        const bool ylow = ly == 0, xlow = lx == 0;
        auto load_yrow = [&](int zl) -> uint4 {   // row y-1 (ly==0) or y+1 (ly==TY-1) of plane z0+zl
            const size_t zoff = (size_t)(z0 + zl) * plane;
            if (ylow) return fym ? hym[zl * LX + lx] : *reinterpret_cast<const uint4 *>(col + zoff - P.nx);
            return fyp ? hyp[zl * LX + lx] : *reinterpret_cast<const uint4 *>(col + zoff + P.nx);
        };
        auto load_xedge = [&](int zl) -> __nv_bfloat16 {   // cell x0-1 (lane 0) or x0+8 (lane 31) of plane z0+zl
            const size_t zoff = (size_t)(z0 + zl) * plane;
            if (xlow) return fxm ? hxm[zl * TY + ly] : col[zoff - 1];
            return fxp ? hxp[zl * TY + ly] : col[zoff + 8];
        };
        uint4 q[PD], qy[PD];
        __nv_bfloat16 qx[PD];
#pragma unroll
        for (int j = 0; j < PD; j++) {
            q[j] = load_plane(z0 + 1 + j);
            if (yedge) qy[j] = load_yrow(j);
            if (xedge) qx[j] = load_xedge(j);
        }
        float prev[8], cur[8];
        uint4 vcur = load_plane(z0);
        unpack8(load_plane(z0 - 1), prev);
        unpack8(vcur, cur);
#pragma unroll 1
        for (int zb = 0; zb < TZ; zb += PD) {
#pragma unroll
            for (int j = 0; j < PD; j++) {
                const int zl = zb + j, z = z0 + zl;
                const uint4 vnext = q[j];
                const uint4 vy = qy[j];
                const __nv_bfloat16 vx = qx[j];
                if (zl + 1 + PD <= TZ) q[j] = load_plane(z + 1 + PD);
                if (zl + PD < TZ) {
                    if (yedge) qy[j] = load_yrow(zl + PD);
                    if (xedge) qx[j] = load_xedge(zl + PD);
                }
                rows[zl & 1][ly][lx] = vcur;
                float left = __shfl_up_sync(0xffffffffu, cur[7], 1), right = __shfl_down_sync(0xffffffffu, cur[0], 1);
                if (lx == 0) left = __bfloat162float(vx);
                if (lx == LX - 1) right = __bfloat162float(vx);
                __syncthreads();
                float ym[8], yp[8], nxt[8];
                unpack8(ly > 0 ? rows[zl & 1][ly - 1][lx] : vy, ym);
                unpack8(ly < TY - 1 ? rows[zl & 1][ly + 1][lx] : vy, yp);
                unpack8(vnext, nxt);
                float out[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float xm = i == 0 ? left : cur[i - 1], xp = i == 7 ? right : cur[i + 1];
                    out[i] = P.c0 * cur[i] + P.c1 * (((xm + xp) + (ym[i] + yp[i])) + (prev[i] + nxt[i]));
                }
                const uint4 vo = pack8(out);
                *reinterpret_cast<uint4 *>(P.un + (size_t)z * plane + (size_t)y * P.nx + x0) = vo;
                // face cells also go straight into the neighbours' landing buffers (NVLink stores)
                if (fzm && zl == 0) rzm[ly * LX + lx] = vo;
                if (fzp && zl == TZ - 1) rzp[ly * LX + lx] = vo;
                if (fym && ly == 0) rym[zl * LX + lx] = vo;
                if (fyp && ly == TY - 1) ryp[zl * LX + lx] = vo;
                if (fxm && lx == 0) xface[0][zl * TY + ly] = reinterpret_cast<const __nv_bfloat16 *>(&vo)[0];
                if (fxp && lx == LX - 1) xface[1][zl * TY + ly] = reinterpret_cast<const __nv_bfloat16 *>(&vo)[7];
#pragma unroll
                for (int i = 0; i < 8; i++) { prev[i] = cur[i]; cur[i] = nxt[i]; }
                vcur = vnext;
            }
        }
        __syncthreads();
        if (fxm && tid < TZ * TY / 8) ((uint4 *)acx::remote_partition_ptr_nowait(P.send[0], px_))[tid] = ((const uint4 *)xface[0])[tid];
        if (fxp && tid >= 64 && tid < 64 + TZ * TY / 8) ((uint4 *)acx::remote_partition_ptr_nowait(P.send[1], px_))[tid - 64] = ((const uint4 *)xface[1])[tid - 64];
        if (fxm | fxp | fym | fyp | fzm | fzp) {
            __syncthreads();
            if (tid == 0) {
                __threadfence_system();   // one cumulative fence (the barrier ordered the other threads' stores before it)
                if (fxm) acx::pready_signal_epoch(px_, P.send[0], P.signal_val);
                if (fxp) acx::pready_signal_epoch(px_, P.send[1], P.signal_val);
                if (fym) acx::pready_signal_epoch(py_, P.send[2], P.signal_val);
                if (fyp) acx::pready_signal_epoch(py_, P.send[3], P.signal_val);
                if (fzm) acx::pready_signal_epoch(pz_, P.send[4], P.signal_val);
                if (fzp) acx::pready_signal_epoch(pz_, P.send[5], P.signal_val);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Barrier-free variant (default): 4 warps per CTA, each warp owns 4 rows x 256 columns of the tile
// and z-marches on its own — y-neighbours of the strip's inner rows are already in registers, the two
// rows above/below come from global memory (the adjacent warp loaded them a moment ago: L1/L2 hits)
// or from the Y landing buffers, software-pipelined one plane ahead. No shared-memory exchange and no
// __syncthreads inside the march; ~3x fewer instructions per cell than the shared-memory variants.
// Same tile geometry (256x16x32), same partitions, same flags.
// ---------------------------------------------------------------------------------------------
constexpr int RW = 4;              // rows per warp
constexpr int NW = TY / RW;        // warps per CTA

template <int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) halo_stencil_step_rows(HaloDev P) {
    __shared__ __align__(16) __nv_bfloat16 xface[2][TZ * TY];
    const int lx = threadIdx.x & 31, w = threadIdx.x >> 5, tid = threadIdx.x;
    const size_t plane = (size_t)P.nx * P.ny;
    for (int it = blockIdx.x; it < P.ntiles; it += gridDim.x) {
        const int t = P.order[it];
        const int tx = t % P.ntx, ty = (t / P.ntx) % P.nty, tz = t / (P.ntx * P.nty);
        const bool fxm = tx == 0, fxp = tx == P.ntx - 1, fym = ty == 0, fyp = ty == P.nty - 1, fzm = tz == 0, fzp = tz == P.ntz - 1;
        const int px_ = tz * P.nty + ty, py_ = tz * P.ntx + tx, pz_ = ty * P.ntx + tx;
        if (tid == 0) {
            if (fxm) acx::wait_arrived_epoch(P.recv[0], px_, P.wait_val);
            if (fxp) acx::wait_arrived_epoch(P.recv[1], px_, P.wait_val);
            if (fym) acx::wait_arrived_epoch(P.recv[2], py_, P.wait_val);
            if (fyp) acx::wait_arrived_epoch(P.recv[3], py_, P.wait_val);
            if (fzm) acx::wait_arrived_epoch(P.recv[4], pz_, P.wait_val);
            if (fzp) acx::wait_arrived_epoch(P.recv[5], pz_, P.wait_val);
        }
        __syncthreads();
        const int x0 = tx * TX + lx * 8, yl0 = w * RW, y0 = ty * TY + yl0, z0 = tz * TZ;
        const __nv_bfloat16 *hxm = fxm ? (const __nv_bfloat16 *)acx::local_partition_ptr(P.recv[0], px_) : nullptr;
        const __nv_bfloat16 *hxp = fxp ? (const __nv_bfloat16 *)acx::local_partition_ptr(P.recv[1], px_) : nullptr;
        const uint4 *hym = fym ? (const uint4 *)acx::local_partition_ptr(P.recv[2], py_) : nullptr;
        const uint4 *hyp = fyp ? (const uint4 *)acx::local_partition_ptr(P.recv[3], py_) : nullptr;
        const uint4 *hzm = fzm ? (const uint4 *)acx::local_partition_ptr(P.recv[4], pz_) : nullptr;
        const uint4 *hzp = fzp ? (const uint4 *)acx::local_partition_ptr(P.recv[5], pz_) : nullptr;
        uint4 *rym = fym ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[2], py_) : nullptr;
        uint4 *ryp = fyp ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[3], py_) : nullptr;
        uint4 *rzm = fzm ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[4], pz_) : nullptr;
        uint4 *rzp = fzp ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[5], pz_) : nullptr;
        const __nv_bfloat16 *col = P.u + (size_t)y0 * P.nx + x0;            // row r, plane z: col + z*plane + r*nx
        __nv_bfloat16 *ocol = P.un + (size_t)y0 * P.nx + x0;
        const bool top = w == 0, bot = w == NW - 1;
        auto load_own = [&](int z, int r) -> uint4 {                          // own row r of plane z (z may be the z-halo)
            if (z < 0) return hzm[(yl0 + r) * LX + lx];
            if (z >= P.nz) return hzp[(yl0 + r) * LX + lx];
            return *reinterpret_cast<const uint4 *>(col + (size_t)z * plane + (size_t)r * P.nx);
        };
        auto load_halo_lo = [&](int zl) -> uint4 {                            // row y0-1 of plane z0+zl
            if (top && fym) return hym[zl * LX + lx];
            return *reinterpret_cast<const uint4 *>(col + (size_t)(z0 + zl) * plane - P.nx);
        };
        auto load_halo_hi = [&](int zl) -> uint4 {                            // row y0+RW of plane z0+zl
            if (bot && fyp) return hyp[zl * LX + lx];
            return *reinterpret_cast<const uint4 *>(col + (size_t)(z0 + zl) * plane + (size_t)RW * P.nx);
        };
        auto load_xe = [&](int zl, int r) -> __nv_bfloat16 {                  // lane 0: cell x0-1; lane 31: cell x0+8
            const size_t o = (size_t)(z0 + zl) * plane + (size_t)r * P.nx;
            if (lx == 0) return fxm ? hxm[zl * TY + yl0 + r] : col[o - 1];
            return fxp ? hxp[zl * TY + yl0 + r] : col[o + 8];
        };
        const bool xedge = lx == 0 || lx == LX - 1;
        float prev[RW][8], cur[RW][8];
        uint4 nxt[RW], hlo, hhi;
        __nv_bfloat16 xe[RW];
#pragma unroll
        for (int r = 0; r < RW; r++) {
            unpack8(load_own(z0 - 1, r), prev[r]);
            unpack8(load_own(z0, r), cur[r]);
            nxt[r] = load_own(z0 + 1, r);
            xe[r] = xedge ? load_xe(0, r) : __float2bfloat16(0.f);
        }
        hlo = load_halo_lo(0); hhi = load_halo_hi(0);
#pragma unroll 1
        for (int zl = 0; zl < TZ; zl++) {
            const int z = z0 + zl;
            // consume this plane's prefetched operands ...
            uint4 vn[RW];
            float edge[RW];
#pragma unroll
            for (int r = 0; r < RW; r++) { vn[r] = nxt[r]; edge[r] = __bfloat162float(xe[r]); }
            const uint4 vlo = hlo, vhi = hhi;
            // ... and put the next plane's in flight before the math
            if (zl + 1 < TZ) {
#pragma unroll
                for (int r = 0; r < RW; r++) { nxt[r] = load_own(z + 2, r); if (xedge) xe[r] = load_xe(zl + 1, r); }
                hlo = load_halo_lo(zl + 1); hhi = load_halo_hi(zl + 1);
            }
            float below[8];                       // old value of the row beneath the one being updated (y-1)
            unpack8(vlo, below);
#pragma unroll
            for (int r = 0; r < RW; r++) {
                float above[8], nf[8], out[8];
                if (r < RW - 1) {
#pragma unroll
                    for (int i = 0; i < 8; i++) above[i] = cur[r + 1][i];
                } else unpack8(vhi, above);
                unpack8(vn[r], nf);
                float left = __shfl_up_sync(0xffffffffu, cur[r][7], 1), right = __shfl_down_sync(0xffffffffu, cur[r][0], 1);
                if (lx == 0) left = edge[r];
                if (lx == LX - 1) right = edge[r];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float xm = i == 0 ? left : cur[r][i - 1], xp = i == 7 ? right : cur[r][i + 1];
                    out[i] = P.c0 * cur[r][i] + P.c1 * (((xm + xp) + (below[i] + above[i])) + (prev[r][i] + nf[i]));
                }
                const uint4 vo = pack8(out);
                *reinterpret_cast<uint4 *>(ocol + (size_t)z * plane + (size_t)r * P.nx) = vo;
                if (fzm && zl == 0) rzm[(yl0 + r) * LX + lx] = vo;
                if (fzp && zl == TZ - 1) rzp[(yl0 + r) * LX + lx] = vo;
                if (fym && top && r == 0) rym[zl * LX + lx] = vo;
                if (fyp && bot && r == RW - 1) ryp[zl * LX + lx] = vo;
                if (fxm && lx == 0) xface[0][zl * TY + yl0 + r] = reinterpret_cast<const __nv_bfloat16 *>(&vo)[0];
                if (fxp && lx == LX - 1) xface[1][zl * TY + yl0 + r] = reinterpret_cast<const __nv_bfloat16 *>(&vo)[7];
#pragma unroll
                for (int i = 0; i < 8; i++) { below[i] = cur[r][i]; prev[r][i] = cur[r][i]; cur[r][i] = nf[i]; }
            }
        }
        __syncthreads();
        if (fxm && tid < TZ * TY / 8) ((uint4 *)acx::remote_partition_ptr_nowait(P.send[0], px_))[tid] = ((const uint4 *)xface[0])[tid];
        if (fxp && tid >= 64 && tid < 64 + TZ * TY / 8) ((uint4 *)acx::remote_partition_ptr_nowait(P.send[1], px_))[tid - 64] = ((const uint4 *)xface[1])[tid - 64];
        if (fxm | fxp | fym | fyp | fzm | fzp) {
            __syncthreads();
            if (tid == 0) {
                __threadfence_system();
                if (fxm) acx::pready_signal_epoch(px_, P.send[0], P.signal_val);
                if (fxp) acx::pready_signal_epoch(px_, P.send[1], P.signal_val);
                if (fym) acx::pready_signal_epoch(py_, P.send[2], P.signal_val);
                if (fyp) acx::pready_signal_epoch(py_, P.send[3], P.signal_val);
                if (fzm) acx::pready_signal_epoch(pz_, P.send[4], P.signal_val);
                if (fzp) acx::pready_signal_epoch(pz_, P.send[5], P.signal_val);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// "direct" variant (default): one row of 8 cells per thread, 512 threads per CTA, no shared-memory
// exchange and no barrier inside the z-march. y-neighbours are read straight from global memory (the
// warps above/below fetch the same rows as their centre: L1/L2 hits, DRAM traffic stays 1x), or from
// the packed Y landing buffers on face tiles. Every operand source is resolved ONCE per tile into a
// (pointer, plane-stride) pair, so the loop body is pointer bumps + 3 vector loads issued a full plane
// ahead + fp32 math: ~2x fewer instructions than the other variants and <= 64 registers (2 CTAs/SM).
// ---------------------------------------------------------------------------------------------
template <int MINB>
__global__ void __launch_bounds__(LX * TY, MINB) halo_stencil_step_direct(HaloDev P) {
    __shared__ __align__(16) __nv_bfloat16 xface[2][TZ * TY];
    const int lx = threadIdx.x, ly = threadIdx.y, tid = ly * LX + lx;
    const size_t plane = (size_t)P.nx * P.ny;
    for (int it = blockIdx.x; it < P.ntiles; it += gridDim.x) {
        const int t = P.order[it];
        const int tx = t % P.ntx, ty = (t / P.ntx) % P.nty, tz = t / (P.ntx * P.nty);
        const bool fxm = tx == 0, fxp = tx == P.ntx - 1, fym = ty == 0, fyp = ty == P.nty - 1, fzm = tz == 0, fzp = tz == P.ntz - 1;
        const int px_ = tz * P.nty + ty, py_ = tz * P.ntx + tx, pz_ = ty * P.ntx + tx;
        if (tid == 0) {
            if (fxm) acx::wait_arrived_epoch(P.recv[0], px_, P.wait_val);
            if (fxp) acx::wait_arrived_epoch(P.recv[1], px_, P.wait_val);
            if (fym) acx::wait_arrived_epoch(P.recv[2], py_, P.wait_val);
            if (fyp) acx::wait_arrived_epoch(P.recv[3], py_, P.wait_val);
            if (fzm) acx::wait_arrived_epoch(P.recv[4], pz_, P.wait_val);
            if (fzp) acx::wait_arrived_epoch(P.recv[5], pz_, P.wait_val);
        }
        __syncthreads();
        const int x0 = tx * TX + lx * 8, y = ty * TY + ly, z0 = tz * TZ;
        const __nv_bfloat16 *c0p = P.u + (size_t)z0 * plane + (size_t)y * P.nx + x0;   // centre cell of plane z0
        // --- operand sources, resolved once per tile: (pointer at plane z0, stride per plane in 16-B units or elements)
        const uint4 *pc = reinterpret_cast<const uint4 *>(c0p);                 // own row
        const size_t sc = plane / 8;
        const uint4 *pym, *pyp; size_t sym, syp;                                // rows y-1 / y+1
        if (ly == 0 && fym) { pym = (const uint4 *)acx::local_partition_ptr(P.recv[2], py_) + lx; sym = LX; }
        else { pym = reinterpret_cast<const uint4 *>(c0p - P.nx); sym = sc; }
        if (ly == TY - 1 && fyp) { pyp = (const uint4 *)acx::local_partition_ptr(P.recv[3], py_) + lx; syp = LX; }
        else { pyp = reinterpret_cast<const uint4 *>(c0p + P.nx); syp = sc; }
        const __nv_bfloat16 *pxe = c0p; size_t sxe = 0;                         // x-edge cell (lanes 0 and 31 only)
        const bool xedge = lx == 0 || lx == LX - 1;
        if (lx == 0) { if (fxm) { pxe = (const __nv_bfloat16 *)acx::local_partition_ptr(P.recv[0], px_) + ly; sxe = TY; } else { pxe = c0p - 1; sxe = plane; } }
        if (lx == LX - 1) { if (fxp) { pxe = (const __nv_bfloat16 *)acx::local_partition_ptr(P.recv[1], px_) + ly; sxe = TY; } else { pxe = c0p + 8; sxe = plane; } }
        uint4 *po = reinterpret_cast<uint4 *>(P.un + (size_t)z0 * plane + (size_t)y * P.nx + x0);
        // remote landing buffers (packed): y faces advance one row of LX vectors per plane
        uint4 *rym = (fym && ly == 0) ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[2], py_) + lx : nullptr;
        uint4 *ryp = (fyp && ly == TY - 1) ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[3], py_) + lx : nullptr;
        uint4 *rzm = fzm ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[4], pz_) + ly * LX + lx : nullptr;
        uint4 *rzp = fzp ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[5], pz_) + ly * LX + lx : nullptr;
        __nv_bfloat16 *sxm = (fxm && lx == 0) ? &xface[0][ly] : nullptr, *sxp = (fxp && lx == LX - 1) ? &xface[1][ly] : nullptr;
        // --- prologue: planes z0-1, z0, z0+1 and the first neighbours
        float2 prev[4], cur[4];
        unpack8x2(fzm ? ((const uint4 *)acx::local_partition_ptr(P.recv[4], pz_))[ly * LX + lx] : *(pc - sc), prev);
        unpack8x2(*pc, cur);
        uint4 vn = (TZ == 1 && fzp) ? ((const uint4 *)acx::local_partition_ptr(P.recv[5], pz_))[ly * LX + lx] : *(pc + sc);   // plane z0+1
        uint4 vym = *pym, vyp = *pyp;
        __nv_bfloat16 vxe = xedge ? *pxe : __float2bfloat16(0.f);
        pc += 2 * sc; pym += sym; pyp += syp; pxe += sxe;
#pragma unroll 2
        for (int zl = 0; zl < TZ; zl++) {
            // operands of THIS plane (loaded one iteration ago) ...
            const uint4 un = vn, uym = vym, uyp = vyp;
            const float edge = __bfloat162float(vxe);
            // ... next plane's loads go out before the math
            if (zl + 2 < TZ) vn = *pc;                                           // own row, plane z+2
            else if (zl + 2 == TZ) vn = fzp ? ((const uint4 *)acx::local_partition_ptr(P.recv[5], pz_))[ly * LX + lx] : *pc;
            if (zl + 1 < TZ) { vym = *pym; vyp = *pyp; if (xedge) vxe = *pxe; }
            pc += sc; pym += sym; pyp += syp; pxe += sxe;
            float2 ym[4], yp[4], nf[4], out[4];
            unpack8x2(uym, ym); unpack8x2(uyp, yp); unpack8x2(un, nf);
            float left = __shfl_up_sync(0xffffffffu, cur[3].y, 1), right = __shfl_down_sync(0xffffffffu, cur[0].x, 1);
            if (lx == 0) left = edge;
            if (lx == LX - 1) right = edge;
            const float2 c0v = make_float2(P.c0, P.c0), c1v = make_float2(P.c1, P.c1);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // x-neighbour sums stay scalar (they straddle register pairs); everything else is FADD2/FMUL2/FFMA2
                const float xs0 = (j == 0 ? left : cur[j - 1].y) + cur[j].y;
                const float xs1 = cur[j].x + (j == 3 ? right : cur[j + 1].x);
                const float2 t = __fadd2_rn(__fadd2_rn(make_float2(xs0, xs1), __fadd2_rn(ym[j], yp[j])), __fadd2_rn(prev[j], nf[j]));
                out[j] = __ffma2_rn(c0v, cur[j], __fmul2_rn(c1v, t));
            }
            const uint4 vo = pack8x2(out);
            *po = vo; po += sc;
            if (rym) { *rym = vo; rym += LX; }
            if (ryp) { *ryp = vo; ryp += LX; }
            if (zl == 0 && rzm) *rzm = vo;
            if (zl == TZ - 1 && rzp) *rzp = vo;
            if (sxm) { *sxm = reinterpret_cast<const __nv_bfloat16 *>(&vo)[0]; sxm += TY; }
            if (sxp) { *sxp = reinterpret_cast<const __nv_bfloat16 *>(&vo)[7]; sxp += TY; }
#pragma unroll
            for (int j = 0; j < 4; j++) { prev[j] = cur[j]; cur[j] = nf[j]; }
        }
        __syncthreads();
        if (fxm && tid < TZ * TY / 8) ((uint4 *)acx::remote_partition_ptr_nowait(P.send[0], px_))[tid] = ((const uint4 *)xface[0])[tid];
        if (fxp && tid >= 64 && tid < 64 + TZ * TY / 8) ((uint4 *)acx::remote_partition_ptr_nowait(P.send[1], px_))[tid - 64] = ((const uint4 *)xface[1])[tid - 64];
        if (fxm | fxp | fym | fyp | fzm | fzp) {
            __syncthreads();
            if (tid == 0) {
                __threadfence_system();
                if (fxm) acx::pready_signal_epoch(px_, P.send[0], P.signal_val);
                if (fxp) acx::pready_signal_epoch(px_, P.send[1], P.signal_val);
                if (fym) acx::pready_signal_epoch(py_, P.send[2], P.signal_val);
                if (fyp) acx::pready_signal_epoch(py_, P.send[3], P.signal_val);
                if (fzm) acx::pready_signal_epoch(pz_, P.send[4], P.signal_val);
                if (fzp) acx::pready_signal_epoch(pz_, P.send[5], P.signal_val);
            }
        }
        __syncthreads();
    }
}

// Deeper-prefetch flavour of the direct kernel (ACX_HALO_VARIANT=7: PD=2, 8: PD=3; not the default — written after the
// round's GPU budget was spent, to be measured next). ncu on the direct kernel shows ideal DRAM traffic but too few
// DRAM-missing loads in flight per SM: only the own-row load of the next plane misses (the y+-1 rows and x edges hit in
// L1/L2), one per thread. Here the own-row loads run PD planes ahead in a register ring (4 registers per stage), so
// each SM keeps PD x 8 KiB of DRAM reads in flight instead of 8 KiB; everything else is the direct kernel unchanged.
template <int PD>
__global__ void __launch_bounds__(LX * TY, 1) halo_stencil_step_direct_pd(HaloDev P) {
    static_assert(PD >= 2 && PD < TZ, "prefetch depth");
    __shared__ __align__(16) __nv_bfloat16 xface[2][TZ * TY];
    const int lx = threadIdx.x, ly = threadIdx.y, tid = ly * LX + lx;
    const size_t plane = (size_t)P.nx * P.ny;
    for (int it = blockIdx.x; it < P.ntiles; it += gridDim.x) {
        const int t = P.order[it];
        const int tx = t % P.ntx, ty = (t / P.ntx) % P.nty, tz = t / (P.ntx * P.nty);
        const bool fxm = tx == 0, fxp = tx == P.ntx - 1, fym = ty == 0, fyp = ty == P.nty - 1, fzm = tz == 0, fzp = tz == P.ntz - 1;
        const int px_ = tz * P.nty + ty, py_ = tz * P.ntx + tx, pz_ = ty * P.ntx + tx;
        if (tid == 0) {
            if (fxm) acx::wait_arrived_epoch(P.recv[0], px_, P.wait_val);
            if (fxp) acx::wait_arrived_epoch(P.recv[1], px_, P.wait_val);
            if (fym) acx::wait_arrived_epoch(P.recv[2], py_, P.wait_val);
            if (fyp) acx::wait_arrived_epoch(P.recv[3], py_, P.wait_val);
            if (fzm) acx::wait_arrived_epoch(P.recv[4], pz_, P.wait_val);
            if (fzp) acx::wait_arrived_epoch(P.recv[5], pz_, P.wait_val);
        }
        __syncthreads();
        const int x0 = tx * TX + lx * 8, y = ty * TY + ly, z0 = tz * TZ;
        const __nv_bfloat16 *c0p = P.u + (size_t)z0 * plane + (size_t)y * P.nx + x0;   // centre cell of plane z0
        // --- operand sources, resolved once per tile: (pointer at plane z0, stride per plane in 16-B units or elements)
        const uint4 *pc = reinterpret_cast<const uint4 *>(c0p);                 // own row
        const size_t sc = plane / 8;
        const uint4 *pym, *pyp; size_t sym, syp;                                // rows y-1 / y+1
        if (ly == 0 && fym) { pym = (const uint4 *)acx::local_partition_ptr(P.recv[2], py_) + lx; sym = LX; }
        else { pym = reinterpret_cast<const uint4 *>(c0p - P.nx); sym = sc; }
        if (ly == TY - 1 && fyp) { pyp = (const uint4 *)acx::local_partition_ptr(P.recv[3], py_) + lx; syp = LX; }
        else { pyp = reinterpret_cast<const uint4 *>(c0p + P.nx); syp = sc; }
        const __nv_bfloat16 *pxe = c0p; size_t sxe = 0;                         // x-edge cell (lanes 0 and 31 only)
        const bool xedge = lx == 0 || lx == LX - 1;
        if (lx == 0) { if (fxm) { pxe = (const __nv_bfloat16 *)acx::local_partition_ptr(P.recv[0], px_) + ly; sxe = TY; } else { pxe = c0p - 1; sxe = plane; } }
        if (lx == LX - 1) { if (fxp) { pxe = (const __nv_bfloat16 *)acx::local_partition_ptr(P.recv[1], px_) + ly; sxe = TY; } else { pxe = c0p + 8; sxe = plane; } }
        uint4 *po = reinterpret_cast<uint4 *>(P.un + (size_t)z0 * plane + (size_t)y * P.nx + x0);
        // remote landing buffers (packed): y faces advance one row of LX vectors per plane
        uint4 *rym = (fym && ly == 0) ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[2], py_) + lx : nullptr;
        uint4 *ryp = (fyp && ly == TY - 1) ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[3], py_) + lx : nullptr;
        uint4 *rzm = fzm ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[4], pz_) + ly * LX + lx : nullptr;
        uint4 *rzp = fzp ? (uint4 *)acx::remote_partition_ptr_nowait(P.send[5], pz_) + ly * LX + lx : nullptr;
        __nv_bfloat16 *sxm = (fxm && lx == 0) ? &xface[0][ly] : nullptr, *sxp = (fxp && lx == LX - 1) ? &xface[1][ly] : nullptr;
        // --- prologue: planes z0-1, z0, z0+1 and the first neighbours
        float2 prev[4], cur[4];
        unpack8x2(fzm ? ((const uint4 *)acx::local_partition_ptr(P.recv[4], pz_))[ly * LX + lx] : *(pc - sc), prev);
        unpack8x2(*pc, cur);
        // own-row ring: vq[i] holds plane z0 + 1 + i; plane z0 + TZ comes from the z+ landing buffer on a face tile
        const uint4 *zp_land = fzp ? (const uint4 *)acx::local_partition_ptr(P.recv[5], pz_) + ly * LX + lx : nullptr;
        uint4 vq[PD];
#pragma unroll
        for (int i = 0; i < PD; i++) vq[i] = *(pc + (size_t)(1 + i) * sc);       // PD < TZ: all inside the tile
        uint4 vym = *pym, vyp = *pyp;
        __nv_bfloat16 vxe = xedge ? *pxe : __float2bfloat16(0.f);
        pym += sym; pyp += syp; pxe += sxe;
#pragma unroll 2
        for (int zl = 0; zl < TZ; zl++) {
            // operands of THIS plane (own row loaded PD iterations ago, neighbours one iteration ago) ...
            const uint4 un = vq[0], uym = vym, uyp = vyp;
            const float edge = __bfloat162float(vxe);
#pragma unroll
            for (int i = 0; i + 1 < PD; i++) vq[i] = vq[i + 1];
            // ... the loads of later planes go out before the math
            const int q = zl + 1 + PD;                                           // own row, plane z0 + q
            if (q < TZ) vq[PD - 1] = *(pc + (size_t)q * sc);
            else if (q == TZ) vq[PD - 1] = zp_land ? *zp_land : *(pc + (size_t)q * sc);
            if (zl + 1 < TZ) { vym = *pym; vyp = *pyp; if (xedge) vxe = *pxe; }
            pym += sym; pyp += syp; pxe += sxe;
            float2 ym[4], yp[4], nf[4], out[4];
            unpack8x2(uym, ym); unpack8x2(uyp, yp); unpack8x2(un, nf);
            float left = __shfl_up_sync(0xffffffffu, cur[3].y, 1), right = __shfl_down_sync(0xffffffffu, cur[0].x, 1);
            if (lx == 0) left = edge;
            if (lx == LX - 1) right = edge;
            const float2 c0v = make_float2(P.c0, P.c0), c1v = make_float2(P.c1, P.c1);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // x-neighbour sums stay scalar (they straddle register pairs); everything else is FADD2/FMUL2/FFMA2
                const float xs0 = (j == 0 ? left : cur[j - 1].y) + cur[j].y;
                const float xs1 = cur[j].x + (j == 3 ? right : cur[j + 1].x);
                const float2 t = __fadd2_rn(__fadd2_rn(make_float2(xs0, xs1), __fadd2_rn(ym[j], yp[j])), __fadd2_rn(prev[j], nf[j]));
                out[j] = __ffma2_rn(c0v, cur[j], __fmul2_rn(c1v, t));
            }
            const uint4 vo = pack8x2(out);
            *po = vo; po += sc;
            if (rym) { *rym = vo; rym += LX; }
            if (ryp) { *ryp = vo; ryp += LX; }
            if (zl == 0 && rzm) *rzm = vo;
            if (zl == TZ - 1 && rzp) *rzp = vo;
            if (sxm) { *sxm = reinterpret_cast<const __nv_bfloat16 *>(&vo)[0]; sxm += TY; }
            if (sxp) { *sxp = reinterpret_cast<const __nv_bfloat16 *>(&vo)[7]; sxp += TY; }
#pragma unroll
            for (int j = 0; j < 4; j++) { prev[j] = cur[j]; cur[j] = nf[j]; }
        }
        __syncthreads();
        if (fxm && tid < TZ * TY / 8) ((uint4 *)acx::remote_partition_ptr_nowait(P.send[0], px_))[tid] = ((const uint4 *)xface[0])[tid];
        if (fxp && tid >= 64 && tid < 64 + TZ * TY / 8) ((uint4 *)acx::remote_partition_ptr_nowait(P.send[1], px_))[tid - 64] = ((const uint4 *)xface[1])[tid - 64];
        if (fxm | fxp | fym | fyp | fzm | fzp) {
            __syncthreads();
            if (tid == 0) {
                __threadfence_system();
                if (fxm) acx::pready_signal_epoch(px_, P.send[0], P.signal_val);
                if (fxp) acx::pready_signal_epoch(px_, P.send[1], P.signal_val);
                if (fym) acx::pready_signal_epoch(py_, P.send[2], P.signal_val);
                if (fyp) acx::pready_signal_epoch(py_, P.send[3], P.signal_val);
                if (fzm) acx::pready_signal_epoch(pz_, P.send[4], P.signal_val);
                if (fzp) acx::pready_signal_epoch(pz_, P.send[5], P.signal_val);
            }
        }
        __syncthreads();
    }
}


// Initial exchange: faces of u^0 -> neighbours' parity-0 landing buffers, flag value 1.
// One CTA per (face, partition).
struct PackDev {
    const __nv_bfloat16 *u;
    int nx, ny, nz, ntx, nty, ntz;
    const acx_pchan *send[6];
    unsigned long long signal_val;
};
__global__ void halo_pack_faces(PackDev P) {
    const int f = blockIdx.y, p = blockIdx.x;
    const int nparts = f < 2 ? P.nty * P.ntz : f < 4 ? P.ntx * P.ntz : P.ntx * P.nty;
    if (p >= nparts) return;
    __nv_bfloat16 *dst = (__nv_bfloat16 *)acx::remote_partition_ptr_nowait(P.send[f], p);
    const size_t plane = (size_t)P.nx * P.ny;
    if (f < 2) {
        const int tz = p / P.nty, ty = p % P.nty, x = f == 0 ? 0 : P.nx - 1;
        for (int i = threadIdx.x; i < TZ * TY; i += blockDim.x) {
            const int zl = i / TY, yl = i % TY;
            dst[i] = P.u[(size_t)(tz * TZ + zl) * plane + (size_t)(ty * TY + yl) * P.nx + x];
        }
    } else if (f < 4) {
        const int tz = p / P.ntx, tx = p % P.ntx, y = f == 2 ? 0 : P.ny - 1;
        for (int i = threadIdx.x; i < TZ * TX; i += blockDim.x) {
            const int zl = i / TX, xl = i % TX;
            dst[i] = P.u[(size_t)(tz * TZ + zl) * plane + (size_t)y * P.nx + tx * TX + xl];
        }
    } else {
        const int ty = p / P.ntx, tx = p % P.ntx, z = f == 4 ? 0 : P.nz - 1;
        for (int i = threadIdx.x; i < TY * TX; i += blockDim.x) {
            const int yl = i / TX, xl = i % TX;
            dst[i] = P.u[(size_t)z * plane + (size_t)(ty * TY + yl) * P.nx + tx * TX + xl];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) acx::pready_signal_epoch(p, P.send[f], P.signal_val);
}

}  // namespace

struct acx_halo {
    int nx, ny, nz, px, py, pz, rank, size;
    int ntx, nty, ntz, ntiles;
    float c0, c1;
    __nv_bfloat16 *u[2];
    int cur;                       // u[cur] holds u^step
    unsigned long long step;       // number of completed steps
    bool exchanged;
    void *land[2][6];              // recv landing buffers per parity/face (symmetric heap)
    MPIX_Request sreq[2][6], rreq[2][6];
    MPIX_Prequest spre[2][6], rpre[2][6];
    int *order_d;
    int grid, variant;
};

#define HCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "[acx halo] %s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define HOK(x) do { int r_ = (x); if (r_) { fprintf(stderr, "[acx halo] %s:%d %s -> %d\n", __FILE__, __LINE__, #x, r_); return 1; } } while (0)

extern "C" int acx_halo_create(int nx, int ny, int nz, int px, int py, int pz, float c0, float c1, acx_halo **out) {
    int rank, size;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    if (px * py * pz != size) { fprintf(stderr, "[acx halo] decomposition %dx%dx%d != %d ranks\n", px, py, pz, size); return 1; }
    if (nx % TX || ny % TY || nz % TZ) { fprintf(stderr, "[acx halo] block %dx%dx%d must be a multiple of the %dx%dx%d tile\n", nx, ny, nz, TX, TY, TZ); return 1; }
    acx_halo *h = new acx_halo();
    h->nx = nx; h->ny = ny; h->nz = nz; h->px = px; h->py = py; h->pz = pz; h->rank = rank; h->size = size; h->c0 = c0; h->c1 = c1;
    h->ntx = nx / TX; h->nty = ny / TY; h->ntz = nz / TZ; h->ntiles = h->ntx * h->nty * h->ntz;
    h->cur = 0; h->step = 0; h->exchanged = false;
    const size_t cells = (size_t)nx * ny * nz;
    HCK(cudaMalloc((void **)&h->u[0], cells * 2));
    HCK(cudaMalloc((void **)&h->u[1], cells * 2));
    const int cx = rank % px, cy = (rank / px) % py, cz = rank / (px * py);
    auto rk = [&](int x, int y, int z) { return (((z + pz) % pz) * py + ((y + py) % py)) * px + ((x + px) % px); };
    const int nbr[6] = {rk(cx - 1, cy, cz), rk(cx + 1, cy, cz), rk(cx, cy - 1, cz), rk(cx, cy + 1, cz), rk(cx, cy, cz - 1), rk(cx, cy, cz + 1)};
    const int nparts[6] = {h->nty * h->ntz, h->nty * h->ntz, h->ntx * h->ntz, h->ntx * h->ntz, h->ntx * h->nty, h->ntx * h->nty};
    const size_t pbytes[6] = {(size_t)TY * TZ * 2, (size_t)TY * TZ * 2, (size_t)TX * TZ * 2, (size_t)TX * TZ * 2, (size_t)TX * TY * 2, (size_t)TX * TY * 2};
    for (int par = 0; par < 2; par++)
        for (int f = 0; f < 6; f++) {
            HOK(MPIX_Alloc_mem(nparts[f] * pbytes[f], MPI_INFO_NULL, &h->land[par][f]));
            HCK(cudaMemset(h->land[par][f], 0, nparts[f] * pbytes[f]));
            // My face f travels to nbr[f] and lands in ITS "halo from direction f^1" buffer: tag by the sender's face id.
            const int tag_send = 100 + par * 8 + f, tag_recv = 100 + par * 8 + (f ^ 1);
            HOK(MPIX_Psend_init(nullptr, nparts[f], (MPI_Count)pbytes[f], MPI_BYTE, nbr[f], tag_send, MPI_COMM_WORLD, MPI_INFO_NULL, &h->sreq[par][f]));
            HOK(MPIX_Precv_init(h->land[par][f], nparts[f], (MPI_Count)pbytes[f], MPI_BYTE, nbr[f], tag_recv, MPI_COMM_WORLD, MPI_INFO_NULL, &h->rreq[par][f]));
        }
    for (int par = 0; par < 2; par++)
        for (int f = 0; f < 6; f++) {
            HOK(MPIX_Prequest_create(h->sreq[par][f], &h->spre[par][f]));
            HOK(MPIX_Prequest_create(h->rreq[par][f], &h->rpre[par][f]));
        }
    // boundary tiles first: their NVLink stores overlap the interior sweep
    std::vector<int> order(h->ntiles);
    for (int i = 0; i < h->ntiles; i++) order[i] = i;
    auto nfaces = [&](int t) {
        int tx = t % h->ntx, ty = (t / h->ntx) % h->nty, tz = t / (h->ntx * h->nty);
        return (tx == 0) + (tx == h->ntx - 1) + (ty == 0) + (ty == h->nty - 1) + (tz == 0) + (tz == h->ntz - 1);
    };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nfaces(a) > nfaces(b); });
    HCK(cudaMalloc((void **)&h->order_d, sizeof(int) * h->ntiles));
    HCK(cudaMemcpy(h->order_d, order.data(), sizeof(int) * h->ntiles, cudaMemcpyHostToDevice));
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const char *var = getenv("ACX_HALO_VARIANT");
    h->variant = var ? atoi(var) : 5;
    if (h->variant == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, halo_stencil_step<1, 2>, LX * TY, 0);
    else if (h->variant == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, halo_stencil_step<2, 1>, LX * TY, 0);
    else if (h->variant == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, halo_stencil_step<4, 1>, LX * TY, 0);
    else if (h->variant == 3) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, halo_stencil_step_rows<3>, NW * 32, 0);
    else if (h->variant == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, halo_stencil_step_rows<2>, NW * 32, 0);
    else if (h->variant == 5) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, halo_stencil_step_direct<1>, LX * TY, 0);
    else if (h->variant == 7) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, halo_stencil_step_direct_pd<2>, LX * TY, 0);
    else if (h->variant == 8) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, halo_stencil_step_direct_pd<3>, LX * TY, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, halo_stencil_step_direct<2>, LX * TY, 0);
    if (occ < 1) occ = 1;
    h->grid = std::min(h->ntiles, sms * occ);
    MPI_Barrier(MPI_COMM_WORLD);
    *out = h;
    return 0;
}

extern "C" void *acx_halo_field(acx_halo *h) { return h->u[h->cur]; }
extern "C" unsigned long long acx_halo_steps_done(acx_halo *h) { return h->step; }

extern "C" int acx_halo_exchange_initial(acx_halo *h, cudaStream_t stream) {
    PackDev P;
    P.u = h->u[h->cur]; P.nx = h->nx; P.ny = h->ny; P.nz = h->nz; P.ntx = h->ntx; P.nty = h->nty; P.ntz = h->ntz;
    const int par = (int)(h->step & 1);
    for (int f = 0; f < 6; f++) P.send[f] = (const acx_pchan *)h->spre[par][f];
    P.signal_val = h->step + 1;
    int maxp = std::max(h->nty * h->ntz, std::max(h->ntx * h->ntz, h->ntx * h->nty));
    halo_pack_faces<<<dim3(maxp, 6), 256, 0, stream>>>(P);
    HCK(cudaGetLastError());
    h->exchanged = true;
    return 0;
}

extern "C" int acx_halo_step(acx_halo *h, int nsteps, cudaStream_t stream) {
    if (!h->exchanged) HOK(acx_halo_exchange_initial(h, stream));
    for (int s = 0; s < nsteps; s++) {
        HaloDev P;
        P.u = h->u[h->cur]; P.un = h->u[h->cur ^ 1];
        P.nx = h->nx; P.ny = h->ny; P.nz = h->nz; P.ntx = h->ntx; P.nty = h->nty; P.ntz = h->ntz;
        const int rp = (int)(h->step & 1), sp = rp ^ 1;
        for (int f = 0; f < 6; f++) { P.recv[f] = (const acx_pchan *)h->rpre[rp][f]; P.send[f] = (const acx_pchan *)h->spre[sp][f]; }
        P.wait_val = h->step + 1; P.signal_val = h->step + 2;
        P.c0 = h->c0; P.c1 = h->c1; P.order = h->order_d; P.ntiles = h->ntiles;
        if (h->variant == 0) halo_stencil_step<1, 2><<<h->grid, dim3(LX, TY), 0, stream>>>(P);
        else if (h->variant == 1) halo_stencil_step<2, 1><<<h->grid, dim3(LX, TY), 0, stream>>>(P);
        else if (h->variant == 2) halo_stencil_step<4, 1><<<h->grid, dim3(LX, TY), 0, stream>>>(P);
        else if (h->variant == 3) halo_stencil_step_rows<3><<<h->grid, NW * 32, 0, stream>>>(P);
        else if (h->variant == 4) halo_stencil_step_rows<2><<<h->grid, NW * 32, 0, stream>>>(P);
        else if (h->variant == 5) halo_stencil_step_direct<1><<<h->grid, dim3(LX, TY), 0, stream>>>(P);
        else if (h->variant == 7) halo_stencil_step_direct_pd<2><<<h->grid, dim3(LX, TY), 0, stream>>>(P);
        else if (h->variant == 8) halo_stencil_step_direct_pd<3><<<h->grid, dim3(LX, TY), 0, stream>>>(P);
        else halo_stencil_step_direct<2><<<h->grid, dim3(LX, TY), 0, stream>>>(P);
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
    cudaFree(h->u[0]); cudaFree(h->u[1]); cudaFree(h->order_d);
    delete h;
    return 0;
}
