/* MPIX_Init / MPIX_Finalize twice in one process (MPI stays up): windows, peer mappings, internal streams, event pool,
 * registration cache and per-thread state must all come back clean. Each session runs eager, rendezvous (heap and
 * private buffers), a partitioned round with host Pready / Parrived and a captured graph. */
#include "acx_test.h"

static unsigned char pat(size_t i, int salt) { return (unsigned char)((i * 17u + (size_t)salt * 5u) & 0xff); }

static void session(int round) {
    cudaStream_t st;
    if (t_gpu) T_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); else st = 0;
    const size_t sizes[3] = {100, 9000, 300000};
    for (int heap = 1; heap >= 0; heap--) for (int s = 0; s < 3; s++) {
        const size_t n = sizes[s];
        unsigned char *out, *in, *h = (unsigned char *)malloc(n);
        if (heap || !t_gpu) { T_OK(MPIX_Alloc_mem(n, MPI_INFO_NULL, &out)); T_OK(MPIX_Alloc_mem(n, MPI_INFO_NULL, &in)); }
        else { T_CUDA(cudaMalloc((void **)&out, n)); T_CUDA(cudaMalloc((void **)&in, n)); }
        for (size_t i = 0; i < n; i++) h[i] = pat(i, t_rank + round * 10 + s);
        if (t_gpu) T_CUDA(cudaMemcpy(out, h, n, cudaMemcpyHostToDevice)); else memcpy(out, h, n);
        MPIX_Request rq[2];
        T_OK(MPIX_Irecv_enqueue(in, (int)n, MPI_BYTE, T_LEFT, 3, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Isend_enqueue(out, (int)n, MPI_BYTE, T_RIGHT, 3, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &st));
        if (s & 1) { T_OK(MPIX_Wait(&rq[0], MPI_STATUS_IGNORE)); T_OK(MPIX_Wait(&rq[1], MPI_STATUS_IGNORE)); }
        else T_OK(MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        if (t_gpu) { T_CUDA(cudaStreamSynchronize(st)); T_CUDA(cudaMemcpy(h, in, n, cudaMemcpyDeviceToHost)); } else memcpy(h, in, n);
        size_t bad = 0;
        for (size_t i = 0; i < n; i++) if (h[i] != pat(i, T_LEFT + round * 10 + s)) bad++;
        T_CHECK(bad == 0, "session %d heap %d n=%zu: %zu wrong bytes", round, heap, n, bad);
        if (heap || !t_gpu) { T_OK(MPIX_Free_mem(out)); T_OK(MPIX_Free_mem(in)); } else { T_CUDA(cudaFree(out)); T_CUDA(cudaFree(in)); }
        free(h);
    }
    /* partitioned round, host-side Pready / Parrived */
    {
        static int ps[4 * 8], pr[4 * 8];
        MPIX_Request q[2];
        for (int i = 0; i < 32; i++) { ps[i] = round * 1000 + t_rank * 100 + i; pr[i] = -1; }
        T_OK(MPIX_Psend_init(ps, 4, 8, MPI_INT, T_RIGHT, 5, MPI_COMM_WORLD, MPI_INFO_NULL, &q[0]));
        T_OK(MPIX_Precv_init(pr, 4, 8, MPI_INT, T_LEFT, 5, MPI_COMM_WORLD, MPI_INFO_NULL, &q[1]));
        T_OK(MPIX_Startall(2, q));
        for (int p = 0; p < 4; p++) T_OK(MPIX_Pready_host(p, &q[0]));
        for (int p = 0; p < 4; p++) { int f = 0; while (!f) T_OK(MPIX_Parrived_host(&q[1], p, &f)); }
        T_OK(MPIX_Waitall(2, q, MPI_STATUSES_IGNORE));
        int bad = 0;
        for (int i = 0; i < 32; i++) if (pr[i] != round * 1000 + T_LEFT * 100 + i) bad++;
        T_CHECK(bad == 0, "session %d partitioned: %d wrong", round, bad);
        T_OK(MPIX_Request_free(&q[0])); T_OK(MPIX_Request_free(&q[1]));
    }
    if (t_gpu) {   /* a captured exchange, launched twice, destroyed before the session ends */
        int *a, *b;
        T_CUDA(cudaMallocHost((void **)&a, 4)); T_CUDA(cudaMallocHost((void **)&b, 4));
        *a = 40 + t_rank + round; *b = -1;
        cudaGraph_t g; cudaGraphExec_t ge; MPIX_Request rq[2];
        T_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
        T_OK(MPIX_Irecv_enqueue(b, 1, MPI_INT, T_LEFT, 6, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Isend_enqueue(a, 1, MPI_INT, T_RIGHT, 6, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        T_CUDA(cudaStreamEndCapture(st, &g));
        T_CUDA(cudaGraphInstantiate(&ge, g, 0));
        for (int k = 0; k < 2; k++) T_CUDA(cudaGraphLaunch(ge, st));
        T_CUDA(cudaStreamSynchronize(st));
        T_CHECK(*b == 40 + T_LEFT + round, "session %d graph: got %d", round, *b);
        T_CUDA(cudaGraphExecDestroy(ge)); T_CUDA(cudaGraphDestroy(g));
        cudaFreeHost(a); cudaFreeHost(b);
        T_CUDA(cudaStreamDestroy(st));
    }
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    session(0);
    MPI_Barrier(MPI_COMM_WORLD);
    T_CHECK(MPIX_Finalize() == 0, "first MPIX_Finalize failed");
    T_CHECK(MPIX_Init() == 0, "second MPIX_Init failed");
    T_CHECK(MPIX_Is_device_transport() == t_gpu, "transport changed across re-initialization");
    session(1);
    return t_finish("reinit");
}
