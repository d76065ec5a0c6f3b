// Halo-exchange + 7-point stencil benchmark (BASELINE.json config #3), bf16, periodic PXxPYxPZ.
//   api   : what a user of the MPIX_* partitioned API writes (and what the reference arm runs):
//           per step host MPIX_Startall -> pack kernel (+ device Pready per partition) ->
//           arrival-poll kernel (device Parrived) -> stencil kernel -> sync -> host MPIX_Waitall.
//           Same source for both arms; -DACX_BENCH_REFERENCE=1 selects plain MPIX_Pready.
//   fused : acx-b200's single-kernel-per-step path (src/kernels/halo_stencil.cu), no host in the loop.
// Both verify against the other / a host reference on request and print one JSON line.
#include <mpi.h>
#include <mpi-acx.h>
#ifndef ACX_BENCH_REFERENCE
#include <acx_kernels.h>
#endif
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "[%d] %s:%d %s: %s\n", g_rank, __FILE__, __LINE__, #x, cudaGetErrorString(e_)); MPI_Abort(MPI_COMM_WORLD, 5); } } while (0)
#define OK(x) do { int r_ = (x); if (r_) { fprintf(stderr, "[%d] %s:%d %s -> %d\n", g_rank, __FILE__, __LINE__, #x, r_); MPI_Abort(MPI_COMM_WORLD, 6); } } while (0)
static int g_rank, g_size;
#ifdef ACX_BENCH_REFERENCE
static const char *kImpl = "reference";
#else
static const char *kImpl = "acx-b200";
#endif
typedef __nv_bfloat16 bf16;

static double max_over_ranks(double v) { double m = v; MPI_Allreduce(&v, &m, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD); return m; }

__host__ __device__ inline float init_value(long gx, long gy, long gz) {
    // smooth-ish deterministic field in [-1,1], exactly representable products avoided on purpose
    unsigned h = (unsigned)(gx * 73856093L ^ gy * 19349663L ^ gz * 83492791L);
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    return ((int)(h & 0xffff) - 32768) / 32768.0f;
}
__global__ void init_field(bf16 *u, int nx, int ny, int nz, int ox, int oy, int oz) {
    size_t n = (size_t)nx * ny * nz;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int x = i % nx, y = (i / nx) % ny, z = i / ((size_t)nx * ny);
        u[i] = __float2bfloat16(init_value(ox + x, oy + y, oz + z));
    }
}

// ---------------------------------------------------------------------------------------------
// API-level variant
// ---------------------------------------------------------------------------------------------
struct Faces { bf16 *f[6]; };                 // packed face buffers: X: [nz][ny], Y: [nz][nx], Z: [ny][nx]
struct Preqs { MPIX_Prequest q[6]; };
constexpr int NPART = 4;                      // partitions per face (slabs along the slow face dimension)

__device__ __forceinline__ size_t face_elems(int f, int nx, int ny, int nz) { return f < 2 ? (size_t)ny * nz : f < 4 ? (size_t)nx * nz : (size_t)nx * ny; }

// grid = (NPART, 6): CTA packs one partition of one face from u, then marks it ready.
__global__ void pack_and_ready(const bf16 *u, Faces send, Preqs sq, int nx, int ny, int nz) {
    const int f = blockIdx.y, p = blockIdx.x;
    const size_t per = face_elems(f, nx, ny, nz) / NPART, lo = per * p;
    bf16 *dst = send.f[f];
    for (size_t i = lo + threadIdx.x; i < lo + per; i += blockDim.x) {
        size_t src;
        if (f < 2) { int z = i / ny, y = i % ny; src = ((size_t)z * ny + y) * nx + (f == 0 ? 0 : nx - 1); }
        else if (f < 4) { int z = i / nx, x = i % nx; src = ((size_t)z * ny + (f == 2 ? 0 : ny - 1)) * nx + x; }
        else { int y = i / nx, x = i % nx; src = ((size_t)(f == 4 ? 0 : nz - 1) * ny + y) * nx + x; }
        dst[i] = u[src];
    }
    __threadfence_system();
    __syncthreads();
#ifdef ACX_BENCH_REFERENCE
    if (threadIdx.x == 0) MPIX_Pready(p, sq.q[f]);
#else
    acx::pready_cta(p, sq.q[f]);   // CTA-cooperative copy into the peer's HBM + release flag
#endif
}

__global__ void poll_arrivals(Preqs rq) {
    const int f = threadIdx.x / NPART, p = threadIdx.x % NPART;
    if (f >= 6) return;
    int flag = 0;
    while (!flag) MPIX_Parrived(rq.q[f], p, &flag);
}

__device__ __forceinline__ float ldbf(const bf16 *p) { return __bfloat162float(*p); }

// One thread per 8 x-cells; neighbours straight from global (L1/L2) or the packed halo buffers.
__global__ void __launch_bounds__(256) stencil_api(const bf16 *u, bf16 *un, Faces halo, int nx, int ny, int nz, float c0, float c1) {
    const int vx = nx / 8;
    const size_t nvec = (size_t)vx * ny * nz;
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        const int x0 = (int)(v % vx) * 8, y = (int)((v / vx) % ny), z = (int)(v / ((size_t)vx * ny));
        const size_t base = ((size_t)z * ny + y) * nx + x0;
        auto row = [&](int yy, int zz, int which) -> const bf16 * {  // pointer to 8 cells at (x0.., yy, zz) incl. halos
            if (yy < 0) return halo.f[2] + (size_t)zz * nx + x0;
            if (yy >= ny) return halo.f[3] + (size_t)zz * nx + x0;
            if (zz < 0) return halo.f[4] + (size_t)yy * nx + x0;
            if (zz >= nz) return halo.f[5] + (size_t)yy * nx + x0;
            (void)which;
            return u + ((size_t)zz * ny + yy) * nx + x0;
        };
        const bf16 *c = u + base, *ym = row(y - 1, z, 0), *yp = row(y + 1, z, 0), *zm = row(y, z - 1, 1), *zp = row(y, z + 1, 1);
        float cc[10];
        cc[0] = x0 == 0 ? ldbf(halo.f[0] + (size_t)z * ny + y) : ldbf(c - 1);
        for (int i = 0; i < 8; i++) cc[i + 1] = ldbf(c + i);
        cc[9] = x0 + 8 == nx ? ldbf(halo.f[1] + (size_t)z * ny + y) : ldbf(c + 8);
        bf16 out[8];
        for (int i = 0; i < 8; i++)
            out[i] = __float2bfloat16(c0 * cc[i + 1] + c1 * (((cc[i] + cc[i + 2]) + (ldbf(ym + i) + ldbf(yp + i))) + (ldbf(zm + i) + ldbf(zp + i))));
        *reinterpret_cast<uint4 *>(un + base) = *reinterpret_cast<const uint4 *>(out);
    }
}

static int run_api(int nx, int ny, int nz, int px, int py, int pz, int steps, int warmup, bool heap, bf16 **result) {
    const int cx = g_rank % px, cy = (g_rank / px) % py, cz = g_rank / (px * py);
    auto rk = [&](int x, int y, int z) { return (((z + pz) % pz) * py + ((y + py) % py)) * px + ((x + px) % px); };
    const int nbr[6] = {rk(cx - 1, cy, cz), rk(cx + 1, cy, cz), rk(cx, cy - 1, cz), rk(cx, cy + 1, cz), rk(cx, cy, cz - 1), rk(cx, cy, cz + 1)};
    const size_t cells = (size_t)nx * ny * nz;
    bf16 *u, *un;
    CK(cudaMalloc(&u, cells * 2)); CK(cudaMalloc(&un, cells * 2));
    init_field<<<1024, 256>>>(u, nx, ny, nz, cx * nx, cy * ny, cz * nz);
    Faces send, recv;
    Preqs sq, rq;
    MPIX_Request reqs[12];
    for (int f = 0; f < 6; f++) {
        const size_t fe = f < 2 ? (size_t)ny * nz : f < 4 ? (size_t)nx * nz : (size_t)nx * ny;
#ifndef ACX_BENCH_REFERENCE
        if (heap) { OK(MPIX_Alloc_mem(fe * 2, MPI_INFO_NULL, &send.f[f])); OK(MPIX_Alloc_mem(fe * 2, MPI_INFO_NULL, &recv.f[f])); } else
#endif
        { CK(cudaMalloc(&send.f[f], fe * 2)); CK(cudaMalloc(&recv.f[f], fe * 2)); }
        OK(MPIX_Psend_init(send.f[f], NPART, (MPI_Count)(fe / NPART), MPI_SHORT, nbr[f], 40 + f, MPI_COMM_WORLD, MPI_INFO_NULL, &reqs[f]));
        OK(MPIX_Precv_init(recv.f[f], NPART, (MPI_Count)(fe / NPART), MPI_SHORT, nbr[f], 40 + (f ^ 1), MPI_COMM_WORLD, MPI_INFO_NULL, &reqs[6 + f]));
    }
    for (int f = 0; f < 6; f++) { OK(MPIX_Prequest_create(reqs[f], &sq.q[f])); OK(MPIX_Prequest_create(reqs[6 + f], &rq.q[f])); }
    cudaStream_t st = 0;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    for (int s = 0; s < warmup + steps; s++) {
        if (s == warmup) { CK(cudaDeviceSynchronize()); MPI_Barrier(MPI_COMM_WORLD); CK(cudaEventRecord(e0, st)); }
        OK(MPIX_Startall(12, reqs));
        pack_and_ready<<<dim3(NPART, 6), 256, 0, st>>>(u, send, sq, nx, ny, nz);
        if (s == 0) {
            // first round: poll from the host so that a lazily connecting transport (CUDA-IPC handle opens in the
            // MPI under the reference) never has to make progress behind a spinning kernel (reference README, "Caveats")
            for (int f = 0; f < 6; f++) for (int p = 0; p < NPART; p++) { int fl = 0; while (!fl) OK(MPIX_Parrived(&reqs[6 + f], p, &fl)); }
        } else poll_arrivals<<<1, 32, 0, st>>>(rq);
        stencil_api<<<1184, 256, 0, st>>>(u, un, recv, nx, ny, nz, 0.4f, 0.1f);
        CK(cudaStreamSynchronize(st));
        OK(MPIX_Waitall(12, reqs, MPI_STATUSES_IGNORE));
        bf16 *t = u; u = un; un = t;
    }
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    for (int f = 0; f < 6; f++) { MPIX_Prequest_free(&sq.q[f]); MPIX_Prequest_free(&rq.q[f]); }
    for (int i = 0; i < 12; i++) OK(MPIX_Request_free(&reqs[i]));
    double t = max_over_ranks(ms) / steps;
    const double bytes = 2.0 * cells * 2;
    if (g_rank == 0)
        printf("{\"bench\": \"halo_api\", \"impl\": \"%s\", \"n_gpus\": %d, \"block\": [%d, %d, %d], \"decomp\": [%d, %d, %d], \"steps\": %d, \"ms_per_step\": %.6f, "
               "\"hbm_gbs_per_gpu\": %.1f, \"cell_updates_per_s_total\": %.4e, \"heap\": %d}\n",
               kImpl, g_size, nx, ny, nz, px, py, pz, steps, t, bytes / t / 1e6, g_size * cells / (t * 1e-3), (int)heap);
    *result = u;
    cudaFree(un);
    return 0;
}

#ifndef ACX_BENCH_REFERENCE
static int run_fused(int nx, int ny, int nz, int px, int py, int pz, int steps, int warmup, bool graph, bf16 **result, acx_halo **hout) {
    acx_halo *h;
    OK(acx_halo_create(nx, ny, nz, px, py, pz, 0.4f, 0.1f, &h));
    const int cx = g_rank % px, cy = (g_rank / px) % py, cz = g_rank / (px * py);
    bf16 *dense = nullptr;   // the op's field is pitched (ghost shell in place): initialise a dense array and import it
    CK(cudaMalloc(&dense, (size_t)nx * ny * nz * 2));
    init_field<<<1024, 256>>>(dense, nx, ny, nz, cx * nx, cy * ny, cz * nz);
    OK(acx_halo_import_dense(h, dense, 0));
    CK(cudaDeviceSynchronize());
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    OK(acx_halo_exchange_initial(h, st));
    OK(acx_halo_step(h, warmup, st));
    CK(cudaStreamSynchronize(st));
    MPI_Barrier(MPI_COMM_WORLD);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms = 0;
    if (graph) {
        // the per-step launches advance host-side state, so capture exactly `steps` steps once
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
        OK(acx_halo_step(h, steps, st));
        CK(cudaStreamEndCapture(st, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        CK(cudaEventRecord(e0, st));
        CK(cudaGraphLaunch(ge, st));
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    } else {
        CK(cudaEventRecord(e0, st));
        OK(acx_halo_step(h, steps, st));
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
    }
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double t = max_over_ranks(ms) / steps;
    const size_t cells = (size_t)nx * ny * nz;
    const double bytes = 2.0 * cells * 2;
    if (g_rank == 0)
        printf("{\"bench\": \"halo_fused\", \"impl\": \"%s\", \"n_gpus\": %d, \"block\": [%d, %d, %d], \"decomp\": [%d, %d, %d], \"steps\": %d, \"ms_per_step\": %.6f, "
               "\"hbm_gbs_per_gpu\": %.1f, \"frac_of_measured_hbm_6571\": %.3f, \"cell_updates_per_s_total\": %.4e, \"graph\": %d}\n",
               kImpl, g_size, nx, ny, nz, px, py, pz, steps, t, bytes / t / 1e6, bytes / t / 1e6 / 6570.9, g_size * cells / (t * 1e-3), (int)graph);
    OK(acx_halo_export_dense(h, dense, 0));
    CK(cudaDeviceSynchronize());
    *result = dense;
    *hout = h;
    cudaStreamDestroy(st);
    return 0;
}
#endif

int main(int argc, char **argv) {
    int provided, ndev = 0;
    MPI_Init_thread(&argc, &argv, MPI_THREAD_MULTIPLE, &provided);
    MPI_Comm_rank(MPI_COMM_WORLD, &g_rank);
    MPI_Comm_size(MPI_COMM_WORLD, &g_size);
    CK(cudaGetDeviceCount(&ndev));
    const char *lr = getenv("LOCAL_RANK") ? getenv("LOCAL_RANK") : getenv("ACX_LOCAL_RANK");
    CK(cudaSetDevice((lr ? atoi(lr) : g_rank) % ndev));
    if (MPIX_Init()) { fprintf(stderr, "MPIX_Init failed\n"); MPI_Abort(MPI_COMM_WORLD, 2); }
    std::string mode = argc > 1 ? argv[1] : "fused";
    int n = 512, steps = 20, warmup = 3, px = 0, py = 1, pz = 1;
    bool heap = true, graph = true, verify = false;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() { return i + 1 < argc ? atoi(argv[++i]) : 0; };
        if (a == "--n") n = val();
        else if (a == "--steps") steps = val();
        else if (a == "--warmup") warmup = val();
        else if (a == "--px") px = val();
        else if (a == "--py") py = val();
        else if (a == "--pz") pz = val();
        else if (a == "--no-heap") heap = false;
        else if (a == "--no-graph") graph = false;
        else if (a == "--verify") verify = true;
    }
    if (px == 0) {  // default decomposition: split x, then y, then z (8 -> 2x2x2)
        px = py = pz = 1;
        int r = g_size, d = 0;
        while (r > 1) { if (r % 2) { px *= r; break; } (d % 3 == 0 ? px : d % 3 == 1 ? py : pz) *= 2; r /= 2; d++; }
    }
    int rc = 0;
    bf16 *res = nullptr;
    if (mode == "api") rc = run_api(n, n, n, px, py, pz, steps, warmup, heap, &res);
#ifndef ACX_BENCH_REFERENCE
    else if (mode == "fused" || mode == "both") {
        acx_halo *h = nullptr;
        rc = run_fused(n, n, n, px, py, pz, steps, warmup, graph, &res, &h);
        if (mode == "both" || verify) {
            // cross-check the fused path against the API-level path (same number of steps, same initial field)
            bf16 *ref = nullptr;
            run_api(n, n, n, px, py, pz, steps, warmup, heap, &ref);
            size_t cells = (size_t)n * n * n;
            std::vector<bf16> a(cells), b(cells);
            CK(cudaMemcpy(a.data(), res, cells * 2, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(b.data(), ref, cells * 2, cudaMemcpyDeviceToHost));
            double maxd = 0; size_t bad = 0;
            for (size_t i = 0; i < cells; i++) { double d = fabs((double)__bfloat162float(a[i]) - (double)__bfloat162float(b[i])); if (d > maxd) maxd = d; if (d > 0.02) bad++; }
            maxd = max_over_ranks(maxd);
            if (g_rank == 0) printf("{\"bench\": \"halo_verify\", \"max_abs_diff_fused_vs_api\": %.6f, \"bad\": %zu}\n", maxd, bad);
            if (bad) rc = 1;
        }
        if (h) acx_halo_destroy(h);
    }
#endif
    else { fprintf(stderr, "unknown mode %s\n", mode.c_str()); rc = 2; }
    fflush(stdout);
    MPIX_Finalize();
    MPI_Finalize();
    return rc;
}
