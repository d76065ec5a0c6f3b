// Point-to-point benchmarks written against the MPIX_* API that acx-b200 and the reference
// (NVIDIA/mpi-acx) share, so the SAME source measures both arms:
//   ring      : every rank Isend->right / Irecv<-left on its stream, Waitall_enqueue; K steps
//               (config "send/recv GB/s", also the e2e variant with per-step H2D + D2H)
//   pingpong  : 2 ranks, on-stream ping-pong 8 B - 1 GiB (BASELINE.json config #2)
//   allpairs  : CUDA-graph captured all-pairs exchange, relaunched (config #5)
// Device-timed with CUDA events on the launching stream, max over ranks, JSON on rank 0.
// Build flags: -DACX_BENCH_REFERENCE=1 when compiled against the reference library.
#include <mpi.h>
#include <mpi-acx.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "[%d] %s:%d %s: %s\n", g_rank, __FILE__, __LINE__, #x, cudaGetErrorString(e_)); MPI_Abort(MPI_COMM_WORLD, 5); } } while (0)
#define OK(x) do { int r_ = (x); if (r_) { fprintf(stderr, "[%d] %s:%d %s -> %d\n", g_rank, __FILE__, __LINE__, #x, r_); MPI_Abort(MPI_COMM_WORLD, 6); } } while (0)

static int g_rank, g_size;
static bool g_sendrecv = false;   // ring --sendrecv (acx-b200 only)
// kernel<<<grid, block, 0, stream>>>(...); on the software CUDA device (tests/emu) kernels are plain functions and the
// grid-stride helpers below run with a small geometry
#ifdef ACX_EMU
#define LAUNCH(kern, grid, block, stream, ...) CK(emu::launch_user(kern, (grid) > 4 ? 4 : (grid), (block) > 8 ? 8 : (block), stream, __VA_ARGS__))
#else
#define LAUNCH(kern, grid, block, stream, ...) kern<<<grid, block, 0, stream>>>(__VA_ARGS__)
#endif
#ifdef ACX_BENCH_REFERENCE
static const char *kImpl = "reference";
#else
static const char *kImpl = "acx-b200";
#endif

static void *comm_alloc(size_t bytes, bool heap) {
    void *p = nullptr;
#ifndef ACX_BENCH_REFERENCE
    if (heap) { OK(MPIX_Alloc_mem(bytes, MPI_INFO_NULL, &p)); return p; }
#endif
    (void)heap;
    CK(cudaMalloc(&p, bytes));
    return p;
}
static void comm_free(void *p, bool heap) {
#ifndef ACX_BENCH_REFERENCE
    if (heap) { MPIX_Free_mem(p); return; }
#endif
    (void)heap;
    cudaFree(p);
}

static double max_over_ranks(double v) {
    double m = v;
    MPI_Allreduce(&v, &m, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
    return m;
}

__global__ void fill_kernel(unsigned *p, size_t n, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = seed + (unsigned)i * 2654435761u;
}
__global__ void check_kernel(const unsigned *p, size_t n, unsigned seed, unsigned long long *bad) {
    unsigned long long b = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b += p[i] != seed + (unsigned)i * 2654435761u;
    if (b) atomicAdd(bad, b);
}

// ------------------------------------------------------------------------------------------
static int run_ring(size_t bytes, int steps, int warmup, bool heap, bool e2e, bool verify, bool prepost) {
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    char *sbuf = (char *)comm_alloc(bytes, heap), *rbuf = (char *)comm_alloc(bytes, heap);
    char *rbuf2 = prepost ? (char *)comm_alloc(bytes, heap) : nullptr;   // --prepost: receives are posted one step ahead into alternating buffers
    char *h_in = nullptr, *h_out = nullptr;
    const size_t d2h = 4096 < bytes ? 4096 : bytes;
    if (e2e) { CK(cudaMallocHost((void **)&h_in, bytes)); CK(cudaMallocHost((void **)&h_out, d2h)); memset(h_in, g_rank + 1, bytes); }
    LAUNCH(fill_kernel, 592, 256, st, (unsigned *)sbuf, bytes / 4, 77u * g_rank);
    CK(cudaMemsetAsync(rbuf, 0, bytes, st));
    CK(cudaStreamSynchronize(st));
    int right = (g_rank + 1) % g_size, left = (g_rank + g_size - 1) % g_size;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto step = [&](int tag) {
        MPIX_Request rq[2];
        if (e2e) CK(cudaMemcpyAsync(sbuf, h_in, bytes, cudaMemcpyHostToDevice, st));
#ifndef ACX_BENCH_REFERENCE
        if (g_sendrecv) {   // --sendrecv: the step as one MPIX_Sendrecv_enqueue (one kernel for eager sizes)
            OK(MPIX_Sendrecv_enqueue(sbuf, (int)bytes, MPI_BYTE, right, tag, rbuf, (int)bytes, MPI_BYTE, left, tag, MPI_COMM_WORLD, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
            if (e2e) CK(cudaMemcpyAsync(h_out, rbuf, d2h, cudaMemcpyDeviceToHost, st));
            return;
        }
#endif
        OK(MPIX_Irecv_enqueue(rbuf, (int)bytes, MPI_BYTE, left, tag, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &st));
        OK(MPIX_Isend_enqueue(sbuf, (int)bytes, MPI_BYTE, right, tag, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &st));
        OK(MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        if (e2e) CK(cudaMemcpyAsync(h_out, rbuf, d2h, cudaMemcpyDeviceToHost, st));
    };
    // Pre-posted variant: K steps = first receive, then per step {send i, receive i+1, wait(recv i, send i)}.
    auto run_prepost = [&](int k) {
        MPIX_Request pend, nxt = MPIX_REQUEST_NULL;
        OK(MPIX_Irecv_enqueue(rbuf, (int)bytes, MPI_BYTE, left, 0, MPI_COMM_WORLD, &pend, MPIX_QUEUE_CUDA_STREAM, &st));
        for (int i = 0; i < k; i++) {
            MPIX_Request rq[2];
            if (e2e) CK(cudaMemcpyAsync(sbuf, h_in, bytes, cudaMemcpyHostToDevice, st));
            OK(MPIX_Isend_enqueue(sbuf, (int)bytes, MPI_BYTE, right, 0, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &st));
            if (i + 1 < k) OK(MPIX_Irecv_enqueue((i + 1) & 1 ? rbuf2 : rbuf, (int)bytes, MPI_BYTE, left, 0, MPI_COMM_WORLD, &nxt, MPIX_QUEUE_CUDA_STREAM, &st));
            rq[0] = pend;
            OK(MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
            if (e2e) CK(cudaMemcpyAsync(h_out, i & 1 ? rbuf2 : rbuf, d2h, cudaMemcpyDeviceToHost, st));
            pend = nxt;
        }
    };
    if (prepost) run_prepost(warmup); else for (int i = 0; i < warmup; i++) step(0);
    CK(cudaStreamSynchronize(st));
    unsigned long long bad = 0;
    if (verify && !e2e) {
        unsigned long long *dbad;
        CK(cudaMalloc(&dbad, 8)); CK(cudaMemset(dbad, 0, 8));
        LAUNCH(check_kernel, 592, 256, st, (const unsigned *)((prepost && ((warmup - 1) & 1)) ? rbuf2 : rbuf), bytes / 4, 77u * left, dbad);
        CK(cudaMemcpyAsync(&bad, dbad, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        cudaFree(dbad);
    }
    MPI_Barrier(MPI_COMM_WORLD);
    CK(cudaEventRecord(e0, st));
    if (prepost) run_prepost(steps); else for (int i = 0; i < steps; i++) step(0);
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    MPI_Barrier(MPI_COMM_WORLD);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double tmax = max_over_ranks(ms) / steps;
    double badmax = max_over_ranks((double)bad);
    if (rbuf2) comm_free(rbuf2, heap);
    if (g_rank == 0)
        printf("{\"bench\": \"ring%s%s\", \"impl\": \"%s\", \"n_gpus\": %d, \"bytes\": %zu, \"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.6f, "
               "\"gbs_per_gpu\": %.3f, \"gbs_total\": %.3f, \"heap\": %d, \"bad_words\": %.0f, \"h2d_bytes_per_step\": %zu, \"d2h_bytes_per_step\": %zu}\n",
               e2e ? "_e2e" : "", prepost ? "_prepost" : "", kImpl, g_size, bytes, steps, warmup, tmax, bytes / tmax / 1e6, g_size * (bytes / tmax / 1e6), (int)heap, badmax,
               e2e ? bytes : (size_t)0, e2e ? d2h : (size_t)0);
    comm_free(sbuf, heap); comm_free(rbuf, heap);
    if (e2e) { cudaFreeHost(h_in); cudaFreeHost(h_out); }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return badmax != 0;
}

// ------------------------------------------------------------------------------------------
// Same ping-pong, but `unroll` iterations are captured once into a CUDA graph and the graph is relaunched: the host's
// per-call launch cost (the limiter of the plain on-stream loop) is paid once, what remains is the device-side path.
// One ping-pong iteration. `blocking` (acx-b200 only): MPIX_Send_enqueue / MPIX_Recv_enqueue — an eager receive is one fused
// post+wait kernel, so an iteration is two launches per rank instead of three.
static bool g_blocking = false;
static int pingpong_iter(char *sbuf, char *rbuf, size_t b, cudaStream_t *st) {
    const int peer = 1 - g_rank, t_out = g_rank == 0 ? 1 : 2, t_in = g_rank == 0 ? 2 : 1;
#ifndef ACX_BENCH_REFERENCE
    if (g_blocking) {
        if (g_rank == 0) {
            OK(MPIX_Send_enqueue(sbuf, (int)b, MPI_BYTE, peer, t_out, MPI_COMM_WORLD, MPIX_QUEUE_CUDA_STREAM, st));
            OK(MPIX_Recv_enqueue(rbuf, (int)b, MPI_BYTE, peer, t_in, MPI_COMM_WORLD, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, st));
        } else {
            OK(MPIX_Recv_enqueue(rbuf, (int)b, MPI_BYTE, peer, t_in, MPI_COMM_WORLD, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, st));
            OK(MPIX_Send_enqueue(sbuf, (int)b, MPI_BYTE, peer, t_out, MPI_COMM_WORLD, MPIX_QUEUE_CUDA_STREAM, st));
        }
        return 0;
    }
#endif
    MPIX_Request q;
    if (g_rank == 0) {
        OK(MPIX_Isend_enqueue(sbuf, (int)b, MPI_BYTE, peer, t_out, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, st));
        OK(MPIX_Wait_enqueue(&q, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, st));
        OK(MPIX_Irecv_enqueue(rbuf, (int)b, MPI_BYTE, peer, t_in, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, st));
        OK(MPIX_Wait_enqueue(&q, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, st));
    } else {
        OK(MPIX_Irecv_enqueue(rbuf, (int)b, MPI_BYTE, peer, t_in, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, st));
        OK(MPIX_Wait_enqueue(&q, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, st));
        OK(MPIX_Isend_enqueue(sbuf, (int)b, MPI_BYTE, peer, t_out, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, st));
        OK(MPIX_Wait_enqueue(&q, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, st));
    }
    return 0;
}

static int run_pingpong_graph(size_t min_b, size_t max_b, bool heap, int launches) {
    if (g_size < 2) { if (g_rank == 0) printf("{\"bench\": \"pingpong_graph\", \"skipped\": \"needs >= 2 ranks\"}\n"); return 0; }
    const int unroll = 16;
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    char *sbuf = (char *)comm_alloc(max_b, heap), *rbuf = (char *)comm_alloc(max_b, heap);
    CK(cudaMemset(sbuf, 1, max_b)); CK(cudaMemset(rbuf, 0, max_b));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    if (g_rank == 0) printf("{\"bench\": \"pingpong_graph\", \"impl\": \"%s\", \"heap\": %d, \"blocking_calls\": %d, \"unroll\": %d, \"rows\": [", kImpl, (int)heap, (int)g_blocking, unroll);
    bool first = true;
    for (size_t b = min_b; b <= max_b; b *= 4) {
        cudaGraph_t g; cudaGraphExec_t ge;
        if (g_rank < 2) {
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
            for (int i = 0; i < unroll; i++) {
                pingpong_iter(sbuf, rbuf, b, &st);
            }
            CK(cudaStreamEndCapture(st, &g));
            CK(cudaGraphInstantiate(&ge, g, 0));
        }
        float ms = 0;
        for (int phase = 0; phase < 2; phase++) {
            int n = phase == 0 ? 2 : launches;
            MPI_Barrier(MPI_COMM_WORLD);
            if (phase == 1) CK(cudaEventRecord(e0, st));
            if (g_rank < 2) for (int i = 0; i < n; i++) CK(cudaGraphLaunch(ge, st));
            if (phase == 1) CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
        }
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double rtt_us = max_over_ranks(g_rank < 2 ? ms : 0) * 1e3 / (launches * unroll);
        if (g_rank == 0) { printf("%s{\"bytes\": %zu, \"oneway_us\": %.3f, \"gbs\": %.3f}", first ? "" : ", ", b, rtt_us / 2, b / (rtt_us / 2) / 1e3); first = false; fflush(stdout); }
        if (g_rank < 2) { CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g)); }
    }
    if (g_rank == 0) printf("]}\n");
    comm_free(sbuf, heap); comm_free(rbuf, heap);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return 0;
}

static int run_pingpong(size_t min_b, size_t max_b, bool heap, int max_iters) {
    if (g_size < 2) { if (g_rank == 0) printf("{\"bench\": \"pingpong\", \"skipped\": \"needs >= 2 ranks\"}\n"); return 0; }
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    char *sbuf = (char *)comm_alloc(max_b, heap), *rbuf = (char *)comm_alloc(max_b, heap);
    CK(cudaMemset(sbuf, 1, max_b)); CK(cudaMemset(rbuf, 0, max_b));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    if (g_rank == 0) printf("{\"bench\": \"pingpong\", \"impl\": \"%s\", \"heap\": %d, \"blocking_calls\": %d, \"rows\": [", kImpl, (int)heap, (int)g_blocking);
    bool first = true;
    for (size_t b = min_b; b <= max_b; b *= 2) {
        int iters = b <= (64u << 10) ? max_iters : b <= (16u << 20) ? max_iters / 4 : 6;
        if (iters < 3) iters = 3;
        int warm = 3;
        float ms = 0;
        for (int phase = 0; phase < 2; phase++) {
            int n = phase == 0 ? warm : iters;
            MPI_Barrier(MPI_COMM_WORLD);
            if (phase == 1) CK(cudaEventRecord(e0, st));
            if (g_rank < 2) {
                for (int i = 0; i < n; i++) {
                    pingpong_iter(sbuf, rbuf, b, &st);
                }
            }
            if (phase == 1) CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
        }
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double rtt_us = max_over_ranks(g_rank < 2 ? ms : 0) * 1e3 / iters;
        if (g_rank == 0) {
            printf("%s{\"bytes\": %zu, \"iters\": %d, \"oneway_us\": %.3f, \"gbs\": %.3f}", first ? "" : ", ", b, iters, rtt_us / 2, b / (rtt_us / 2) / 1e3);
            first = false;
            fflush(stdout);
        }
        if (b == 0) break;
    }
    if (g_rank == 0) printf("]}\n");
    comm_free(sbuf, heap); comm_free(rbuf, heap);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return 0;
}

// ------------------------------------------------------------------------------------------
// All-pairs: every rank sends `bytes` to every other rank and receives from every other rank,
// captured once into a CUDA graph and relaunched `steps` times.
static int run_allpairs(size_t bytes, int steps, int warmup, bool heap) {
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int npeer = g_size > 1 ? g_size - 1 : 1;
    char *sbuf = (char *)comm_alloc(bytes * npeer, heap), *rbuf = (char *)comm_alloc(bytes * npeer, heap);
    CK(cudaMemset(sbuf, 3, bytes * npeer)); CK(cudaMemset(rbuf, 0, bytes * npeer));
    std::vector<MPIX_Request> rq(2 * npeer);
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
    for (int k = 0; k < npeer; k++) {
        int to = g_size > 1 ? (g_rank + 1 + k) % g_size : 0, from = g_size > 1 ? (g_rank + g_size - 1 - k) % g_size : 0;
        OK(MPIX_Irecv_enqueue(rbuf + bytes * k, (int)bytes, MPI_BYTE, from, 7, MPI_COMM_WORLD, &rq[2 * k], MPIX_QUEUE_CUDA_STREAM, &st));
        OK(MPIX_Isend_enqueue(sbuf + bytes * k, (int)bytes, MPI_BYTE, to, 7, MPI_COMM_WORLD, &rq[2 * k + 1], MPIX_QUEUE_CUDA_STREAM, &st));
    }
    OK(MPIX_Waitall_enqueue(2 * npeer, rq.data(), MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
    CK(cudaStreamEndCapture(st, &graph));
    CK(cudaGraphInstantiate(&exec, graph, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < warmup; i++) CK(cudaGraphLaunch(exec, st));
    CK(cudaStreamSynchronize(st));
    MPI_Barrier(MPI_COMM_WORLD);
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < steps; i++) CK(cudaGraphLaunch(exec, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double t = max_over_ranks(ms) / steps;
    if (g_rank == 0)
        printf("{\"bench\": \"allpairs_graph\", \"impl\": \"%s\", \"n_gpus\": %d, \"bytes_per_pair\": %zu, \"steps\": %d, \"ms_per_step\": %.6f, \"gbs_per_gpu_out\": %.3f, \"gbs_total\": %.3f, \"heap\": %d}\n",
               kImpl, g_size, bytes, steps, t, bytes * (double)npeer / t / 1e6, g_size * (bytes * (double)npeer / t / 1e6), (int)heap);
    CK(cudaGraphExecDestroy(exec));
    CK(cudaGraphDestroy(graph));
    comm_free(sbuf, heap); comm_free(rbuf, heap);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return 0;
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
    std::string mode = argc > 1 ? argv[1] : "ring";
    size_t bytes = 256u << 20, minb = 8, maxb = 1u << 30;
    int steps = 20, warmup = 5, iters = 200, rc = 0;
    bool heap = true, e2e = false, verify = true, prepost = false;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() { return i + 1 < argc ? atoll(argv[++i]) : 0LL; };
        if (a == "--bytes") bytes = (size_t)val();
        else if (a == "--min") minb = (size_t)val();
        else if (a == "--max") maxb = (size_t)val();
        else if (a == "--steps") steps = (int)val();
        else if (a == "--warmup") warmup = (int)val();
        else if (a == "--iters") iters = (int)val();
        else if (a == "--no-heap") heap = false;
        else if (a == "--e2e") e2e = true;
        else if (a == "--no-verify") verify = false;
        else if (a == "--prepost") prepost = true;
        else if (a == "--sendrecv") g_sendrecv = true;
        else if (a == "--blocking") g_blocking = true;   // pingpong / pingpong_graph through MPIX_Send_enqueue / MPIX_Recv_enqueue (not in the reference build)
    }
    if (mode == "ring") rc = run_ring(bytes, steps, warmup, heap, e2e, verify, prepost);
    else if (mode == "pingpong") rc = run_pingpong(minb, maxb, heap, iters);
    else if (mode == "pingpong_graph") rc = run_pingpong_graph(minb, maxb, heap, iters > 40 ? 20 : iters);
    else if (mode == "allpairs") rc = run_allpairs(bytes, steps, warmup, heap);
    else { fprintf(stderr, "unknown mode %s\n", mode.c_str()); rc = 2; }
    fflush(stdout);
    MPIX_Finalize();
    MPI_Finalize();
    return rc;
}
