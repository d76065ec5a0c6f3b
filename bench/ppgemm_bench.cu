// Pipeline-parallel fp8 GEMM chain benchmark (BASELINE.json config #4): R ranks = R stages,
// each stage Y = scale * X * W_r^T with X [M][D], W_r [D][D] in e4m3 (tcgen05/TMEM kernel,
// src/kernels/gemm_fp8_sm100.cuh), activations handed stage to stage in `P` row-panel partitions.
//   fused : acx-b200 — epilogue stores panels into the next stage's HBM + release flags, consumer's
//           TMA producer gated per panel; one kernel per microbatch per stage, no host in the loop.
//   api   : what the MPIX_* partitioned API allows without fusion (and what the reference arm runs):
//           host Startall -> arrival-poll kernel -> GEMM (local output) -> Pready kernel -> sync ->
//           host Waitall, per microbatch. Same GEMM kernel in both arms.
#include <mpi.h>
#include <mpi-acx.h>
#ifndef ACX_BENCH_REFERENCE
#include <acx_kernels.h>
#endif
#include "../src/kernels/gemm_fp8_sm100.cuh"
#include <cuda_fp8.h>
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
static double max_over_ranks(double v) { double m = v; MPI_Allreduce(&v, &m, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD); return m; }

__global__ void fill_e4m3(uint8_t *p, size_t n, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u + seed; h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
        p[i] = (uint8_t)(((h & 1) << 7) | ((h >> 1) % 0x48));   // |v| < 4, never NaN
    }
}
static float e4m3_to_float(uint8_t v) {
    int s = v >> 7, e = (v >> 3) & 15, m = v & 7;
    float f = e == 0 ? ldexpf((float)m, -9) : ldexpf(1.0f + m / 8.0f, e - 7);
    return s ? -f : f;
}
// sampled check of Y = scale * X W^T
static int verify(const uint8_t *dX, const uint8_t *dW, const uint8_t *dY, int M, int D, float scale) {
    std::vector<uint8_t> x((size_t)M * D), w((size_t)D * D), y((size_t)M * D);
    CK(cudaMemcpy(x.data(), dX, x.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(w.data(), dW, w.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(y.data(), dY, y.size(), cudaMemcpyDeviceToHost));
    unsigned s = 777 + g_rank;
    int wrong = 0, near = 0;
    for (int t = 0; t < 512; t++) {
        s = s * 1664525u + 1013904223u; int i = (s >> 8) % M;
        s = s * 1664525u + 1013904223u; int j = (s >> 8) % D;
        double acc = 0;
        for (int k = 0; k < D; k++) acc += (double)e4m3_to_float(x[(size_t)i * D + k]) * (double)e4m3_to_float(w[(size_t)j * D + k]);
        float ref = (float)(acc * scale);
        uint8_t want = (uint8_t)__nv_cvt_float_to_fp8(ref, __NV_SATFINITE, __NV_E4M3), got = y[(size_t)i * D + j];
        if (got != want) {
            float gf = e4m3_to_float(got), wf = e4m3_to_float(want);
            if (fabsf(gf - wf) <= fabsf(wf) * 0.13f + 2e-3f) near++;
            else { if (wrong < 4) fprintf(stderr, "[%d] mismatch (%d,%d) got %g want %g\n", g_rank, i, j, gf, wf); wrong++; }
        }
    }
    return wrong + (near > 50 ? 1 : 0);
}

#ifndef ACX_BENCH_REFERENCE
static int run_fused(int M, int D, int P, int steps, int warmup, bool check) {
    acx_ppgemm *h;
    const float scale = 1.0f / sqrtf((float)D) * 1.5f;
    OK(acx_ppgemm_create(M, D, P, scale, &h));
    fill_e4m3<<<1024, 256>>>((uint8_t *)acx_ppgemm_weights(h), (size_t)D * D, 1000u + g_rank);
    if (g_rank == 0) fill_e4m3<<<1024, 256>>>((uint8_t *)acx_ppgemm_input(h, 0), (size_t)M * D, 7u);
    CK(cudaDeviceSynchronize());
    MPI_Barrier(MPI_COMM_WORLD);
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < warmup; i++) OK(acx_ppgemm_step(h, st));
    CK(cudaStreamSynchronize(st));
    MPI_Barrier(MPI_COMM_WORLD);
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < steps; i++) OK(acx_ppgemm_step(h, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    MPI_Barrier(MPI_COMM_WORLD);
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double t = max_over_ranks(ms) / steps;
    int bad = 0;
    if (check) {
        const int last_mb = warmup + steps - 1, b = last_mb & 1;
        bad = verify((const uint8_t *)acx_ppgemm_input(h, g_rank == 0 ? 0 : b), (const uint8_t *)acx_ppgemm_weights(h), (const uint8_t *)acx_ppgemm_output(h, b), M, D, scale);
    }
    double badmax = max_over_ranks(bad);
    const double flop = 2.0 * M * (double)D * D;
    if (g_rank == 0)
        printf("{\"bench\": \"ppgemm_fused\", \"impl\": \"%s\", \"n_gpus\": %d, \"M\": %d, \"D\": %d, \"partitions\": %d, \"steps\": %d, \"ms_per_step\": %.5f, "
               "\"tflops_per_gpu\": %.1f, \"tflops_total\": %.1f, \"nvlink_gbs_per_link\": %.1f, \"verify_wrong\": %.0f}\n",
               kImpl, g_size, M, D, P, steps, t, flop / t / 1e9, g_size * flop / t / 1e9, g_size > 1 ? (double)M * D / t / 1e6 : 0.0, badmax);
    acx_ppgemm_destroy(h);
    cudaStreamDestroy(st);
    return badmax != 0;
}
#endif

struct PQ { MPIX_Prequest q; int n; };
__global__ void poll_all(PQ p) { if ((int)threadIdx.x < p.n) { int f = 0; while (!f) MPIX_Parrived(p.q, threadIdx.x, &f); } }
#ifdef ACX_BENCH_REFERENCE
__global__ void ready_all(PQ p) { if ((int)threadIdx.x < p.n) MPIX_Pready(threadIdx.x, p.q); }
#else
__global__ void __launch_bounds__(512) ready_all(PQ p) { acx::pready_grid((int)(blockIdx.x / 32), p.q, blockIdx.x % 32, 32); }   // 16-MiB panels: a grid per partition, not one CTA
#endif

static int run_api(int M, int D, int P, int steps, int warmup, bool check) {
    const float scale = 1.0f / sqrtf((float)D) * 1.5f;
    const size_t act = (size_t)M * D;
    const bool first = g_rank == 0, last = g_rank == g_size - 1;
    uint8_t *W, *xin, *ysend;
    CK(cudaMalloc(&W, (size_t)D * D));
#ifndef ACX_BENCH_REFERENCE
    OK(MPIX_Alloc_mem(act, MPI_INFO_NULL, &xin)); OK(MPIX_Alloc_mem(act, MPI_INFO_NULL, &ysend));   // symmetric heap: partitions land in place
#else
    CK(cudaMalloc(&xin, act)); CK(cudaMalloc(&ysend, act));
#endif
    fill_e4m3<<<1024, 256>>>(W, (size_t)D * D, 1000u + g_rank);
    if (first) fill_e4m3<<<1024, 256>>>(xin, act, 7u);
    CK(cudaDeviceSynchronize());
    MPIX_Request reqs[2]; int nreq = 0, ri = -1, si = -1;
    MPIX_Prequest rq = nullptr, sq = nullptr;
    if (!first) { ri = nreq; OK(MPIX_Precv_init(xin, P, (MPI_Count)(act / P), MPI_BYTE, g_rank - 1, 61, MPI_COMM_WORLD, MPI_INFO_NULL, &reqs[nreq++])); }
    if (!last) { si = nreq; OK(MPIX_Psend_init(ysend, P, (MPI_Count)(act / P), MPI_BYTE, g_rank + 1, 61, MPI_COMM_WORLD, MPI_INFO_NULL, &reqs[nreq++])); }
    if (!first) OK(MPIX_Prequest_create(reqs[ri], &rq));
    if (!last) OK(MPIX_Prequest_create(reqs[si], &sq));
    int dev = 0, sms = 0;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t st = 0;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    acxg::Gating g = {};
    MPI_Barrier(MPI_COMM_WORLD);
    for (int mb = 0; mb < warmup + steps; mb++) {
        if (mb == warmup) { CK(cudaDeviceSynchronize()); MPI_Barrier(MPI_COMM_WORLD); CK(cudaEventRecord(e0, st)); }
        if (nreq) OK(MPIX_Startall(nreq, reqs));
        if (!first) {
            if (mb == 0) { for (int p = 0; p < P; p++) { int fl = 0; while (!fl) OK(MPIX_Parrived(&reqs[ri], p, &fl)); } }   // see halo_bench: lazy connect
            else { PQ p{rq, P}; poll_all<<<1, 32, 0, st>>>(p); }
        }
        int rc = acxg::launch_gemm_fp8(xin, W, ysend, M, D, D, scale, g, sms, st);
        if (rc) { fprintf(stderr, "gemm launch rc=%d\n", rc); MPI_Abort(MPI_COMM_WORLD, 7); }
        if (!last) {
            PQ p{sq, P};
#ifdef ACX_BENCH_REFERENCE
            ready_all<<<1, 32, 0, st>>>(p);
#else
            ready_all<<<P * 32, 512, 0, st>>>(p);
            CK(cudaGetLastError());
#endif
        }
        CK(cudaStreamSynchronize(st));
        if (nreq) OK(MPIX_Waitall(nreq, reqs, MPI_STATUSES_IGNORE));
    }
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double t = max_over_ranks(ms) / steps;
    int bad = check ? verify(xin, W, ysend, M, D, scale) : 0;
    double badmax = max_over_ranks(bad);
    const double flop = 2.0 * M * (double)D * D;
    if (g_rank == 0)
        printf("{\"bench\": \"ppgemm_api\", \"impl\": \"%s\", \"n_gpus\": %d, \"M\": %d, \"D\": %d, \"partitions\": %d, \"steps\": %d, \"ms_per_step\": %.5f, "
               "\"tflops_per_gpu\": %.1f, \"tflops_total\": %.1f, \"verify_wrong\": %.0f}\n",
               kImpl, g_size, M, D, P, steps, t, flop / t / 1e9, g_size * flop / t / 1e9, badmax);
    if (rq) MPIX_Prequest_free(&rq);
    if (sq) MPIX_Prequest_free(&sq);
    for (int i = 0; i < nreq; i++) OK(MPIX_Request_free(&reqs[i]));
    cudaFree(W);
#ifndef ACX_BENCH_REFERENCE
    MPIX_Free_mem(xin); MPIX_Free_mem(ysend);
#else
    cudaFree(xin); cudaFree(ysend);
#endif
    return badmax != 0;
}

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
    int M = 8192, D = 8192, P = 4, steps = 10, warmup = 3, rc = 0;
    bool check = true;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() { return i + 1 < argc ? atoi(argv[++i]) : 0; };
        if (a == "--m") M = val();
        else if (a == "--d") D = val();
        else if (a == "--partitions") P = val();
        else if (a == "--steps") steps = val();
        else if (a == "--warmup") warmup = val();
        else if (a == "--no-check") check = false;
    }
    if (mode == "api") rc = run_api(M, D, P, steps, warmup, check);
#ifndef ACX_BENCH_REFERENCE
    else if (mode == "fused") rc = run_fused(M, D, P, steps, warmup, check);
#endif
    else { fprintf(stderr, "unknown mode %s\n", mode.c_str()); rc = 2; }
    fflush(stdout);
    MPIX_Finalize();
    MPI_Finalize();
    return rc;
}
