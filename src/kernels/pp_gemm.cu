// Pipeline-parallel fp8 GEMM chain (BASELINE.json config #4): one stage per rank, stage-to-stage
// hand-off through acx partitioned channels driven entirely from inside the GEMM kernel
// (gemm_fp8_sm100.cuh): epilogue = Pready with the data already stored remotely, TMA producer =
// Parrived gate. The reference can only express this as GEMM -> host Start/Wait + flag Pready,
// with the CPU proxy and MPI moving the activation afterwards (README "How does it work").
#include <mpi.h>
#include "../../include/mpi-acx.h"
#include "../../include/acx_kernels.h"
#include "gemm_fp8_sm100.cuh"
#include <stdio.h>
#include <vector>

struct acx_ppgemm {
    int M, D, parts, rank, size, sms, grid;
    float scale;
    uint8_t *W;
    uint8_t *xin[2];                 // heap: inputs (double-buffered)
    uint8_t *yout[2];                // last stage: local outputs
    MPIX_Request sreq[2], rreq[2];
    acx_pchan sch[2], rch[2];        // host copies of the device channel descriptors
    unsigned long long *credit;      // heap: credit[2], written by the next stage
    unsigned long long *credit_prev; // peer-mapped: previous stage's credit[2]
    unsigned long long *counters;    // device: panel_count[2][parts], done_count[2]
    unsigned long long mb;
};

#define PCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "[acx ppgemm] %s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define POK(x) do { int r_ = (x); if (r_) { fprintf(stderr, "[acx ppgemm] %s:%d %s -> %d\n", __FILE__, __LINE__, #x, r_); return 1; } } while (0)

// Plain (non-communicating) use of the same kernel: Y = scale * X * W^T, all local.
extern "C" int acx_gemm_fp8(const void *X, const void *W, void *Y, int M, int N, int K, float scale, cudaStream_t stream) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    acxg::Gating g = {};
    return acxg::launch_gemm_fp8((const uint8_t *)X, (const uint8_t *)W, (uint8_t *)Y, M, N, K, scale, g, sms, stream);
}

extern "C" int acx_ppgemm_create(int M, int D, int partitions, float scale, acx_ppgemm **out) {
    acx_ppgemm *h = new acx_ppgemm();
    memset(h, 0, sizeof *h);
    MPI_Comm_rank(MPI_COMM_WORLD, &h->rank);
    MPI_Comm_size(MPI_COMM_WORLD, &h->size);
    h->M = M; h->D = D; h->parts = partitions; h->scale = scale;
    if (M % acxg::BM || D % acxg::BN || D % acxg::BK || M % partitions || (M / partitions) % acxg::BM) {
        fprintf(stderr, "[acx ppgemm] M=%d D=%d partitions=%d not tile-aligned\n", M, D, partitions);
        return 1;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, dev);
    h->grid = acxg::gemm_grid(M, D, h->sms);
    const size_t act = (size_t)M * D;
    PCK(cudaMalloc((void **)&h->W, (size_t)D * D));
    for (int b = 0; b < 2; b++) { POK(MPIX_Alloc_mem(act, MPI_INFO_NULL, &h->xin[b])); PCK(cudaMemset(h->xin[b], 0, act)); }
    POK(MPIX_Alloc_mem(64, MPI_INFO_NULL, &h->credit));
    PCK(cudaMemset(h->credit, 0, 64));
    PCK(cudaMalloc((void **)&h->counters, sizeof(unsigned long long) * (2 * partitions + 2)));
    PCK(cudaMemset(h->counters, 0, sizeof(unsigned long long) * (2 * partitions + 2)));
    const bool first = h->rank == 0, last = h->rank == h->size - 1;
    if (last) for (int b = 0; b < 2; b++) PCK(cudaMalloc((void **)&h->yout[b], act));
    for (int b = 0; b < 2; b++) {
        if (!last) POK(MPIX_Psend_init(nullptr, partitions, (MPI_Count)(act / partitions), MPI_BYTE, h->rank + 1, 60 + b, MPI_COMM_WORLD, MPI_INFO_NULL, &h->sreq[b]));
        if (!first) POK(MPIX_Precv_init(h->xin[b], partitions, (MPI_Count)(act / partitions), MPI_BYTE, h->rank - 1, 60 + b, MPI_COMM_WORLD, MPI_INFO_NULL, &h->rreq[b]));
    }
    for (int b = 0; b < 2; b++) {
        MPIX_Prequest q;
        if (!last) { POK(MPIX_Prequest_create(h->sreq[b], &q)); PCK(cudaMemcpy(&h->sch[b], q, sizeof(acx_pchan), cudaMemcpyDeviceToHost)); }
        if (!first) { POK(MPIX_Prequest_create(h->rreq[b], &q)); PCK(cudaMemcpy(&h->rch[b], q, sizeof(acx_pchan), cudaMemcpyDeviceToHost)); }
    }
    {
        size_t off = 0;
        POK(MPIX_Heap_offset(h->credit, &off));
        std::vector<unsigned long long> offs(h->size);
        unsigned long long mine = off;
        MPI_Allgather(&mine, 1, MPI_UNSIGNED_LONG_LONG, offs.data(), 1, MPI_UNSIGNED_LONG_LONG, MPI_COMM_WORLD);
        if (!first) POK(MPIX_Peer_ptr_at((size_t)offs[h->rank - 1], h->rank - 1, (void **)&h->credit_prev));
    }
    PCK(cudaDeviceSynchronize());
    MPI_Barrier(MPI_COMM_WORLD);
    *out = h;
    return 0;
}

extern "C" void *acx_ppgemm_weights(acx_ppgemm *h) { return h->W; }
extern "C" void *acx_ppgemm_input(acx_ppgemm *h, int buf) { return h->xin[buf & 1]; }
extern "C" void *acx_ppgemm_output(acx_ppgemm *h, int buf) { return h->rank == h->size - 1 ? (void *)h->yout[buf & 1] : (void *)h->sch[buf & 1].land; }
extern "C" unsigned long long acx_ppgemm_microbatches(acx_ppgemm *h) { return h->mb; }

extern "C" int acx_ppgemm_step(acx_ppgemm *h, cudaStream_t stream) {
    const bool first = h->rank == 0, last = h->rank == h->size - 1;
    const int b = (int)(h->mb & 1);
    const unsigned long long e = h->mb / 2 + 1;
    acxg::Gating g = {};
    g.rows_per_panel = h->M / h->parts;
    const uint8_t *X = first ? h->xin[0] : h->xin[b];
    uint8_t *out = last ? h->yout[b] : (uint8_t *)h->sch[b].land;
    if (!first) {
        g.in_arrived = (const unsigned long long *)h->rch[b].arrived; g.in_epoch = e;
        g.done_count = h->counters + 2 * h->parts + b; g.done_target = (unsigned long long)h->grid * e;
        g.credit_out = h->credit_prev + b; g.credit_out_val = e;
    }
    if (!last) {
        g.out_arrived = (unsigned long long *)h->sch[b].arrived; g.out_epoch = e;
        g.panel_count = h->counters + (size_t)b * h->parts;
        g.panel_target = (unsigned long long)((h->M / h->parts / acxg::BM) * (h->D / acxg::BN)) * e;
        g.credit = h->credit + b; g.credit_need = e - 1;
    }
    int rc = acxg::launch_gemm_fp8(X, h->W, out, h->M, h->D, h->D, h->scale, g, h->sms, stream);
    if (rc) { fprintf(stderr, "[acx ppgemm] launch failed (%d)\n", rc); return 1; }
    h->mb++;
    return 0;
}

extern "C" int acx_ppgemm_destroy(acx_ppgemm *h) {
    cudaDeviceSynchronize();
    MPI_Barrier(MPI_COMM_WORLD);
    for (int b = 0; b < 2; b++) {
        if (h->sreq[b]) MPIX_Request_free(&h->sreq[b]);
        if (h->rreq[b]) MPIX_Request_free(&h->rreq[b]);
        MPIX_Free_mem(h->xin[b]);
        if (h->yout[b]) cudaFree(h->yout[b]);
    }
    MPIX_Free_mem(h->credit);
    cudaFree(h->W); cudaFree(h->counters);
    delete h;
    return 0;
}
