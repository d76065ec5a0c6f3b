/* CUDA graphs carrying rendezvous-size messages, relaunched more often than the ring is deep.
 * The small-token graph tests only ever send eager messages; here every launch of one instantiated graph runs the
 * rendezvous handshake, the mover (or the single-launch fused rendezvous), the wait and — for private buffers — the
 * bounce copies again, so everything a launch consumes must be re-armed by the launch itself: CTA counters, the fused
 * gate, ring generations (LAUNCHES > ACX_DEPTH wraps every slot), chunk cursors (ACX_COOP).
 *   mode 0  stream capture (side-stream movers fork and must be joined back into the capture)
 *   mode 1  explicit graphs composed as child-graph nodes
 * for symmetric-heap buffers (zero-copy) and plain cudaMalloc buffers (bounce, then registration cache). The payload of
 * launch i is derived from i on the device side of the stream (a memset node would be constant), by copying a fresh
 * host pattern in before each launch and checking the result after each launch. */
#include "acx_test.h"
#include <unistd.h>

#define LAUNCHES 40

static unsigned char pat(size_t i, int salt) { return (unsigned char)((i * 131u + (size_t)salt * 29u + (i >> 11)) & 0xff); }

static void run(int mode, int heap, size_t n, int tag) {
    unsigned char *d_out, *d_in, *h = (unsigned char *)malloc(n);
    if (heap) { T_OK(MPIX_Alloc_mem(n, MPI_INFO_NULL, &d_out)); T_OK(MPIX_Alloc_mem(n, MPI_INFO_NULL, &d_in)); }
    else { T_CUDA(cudaMalloc((void **)&d_out, n)); T_CUDA(cudaMalloc((void **)&d_in, n)); }
    cudaStream_t stream;
    T_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    cudaGraph_t graph, parts[3];
    cudaGraphExec_t exec;
    MPIX_Request rq[2];
    MPI_Status st[2];
    if (mode == 0) {
        T_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeGlobal));
        T_OK(MPIX_Irecv_enqueue(d_in, (int)n, MPI_BYTE, T_LEFT, tag, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Isend_enqueue(d_out, (int)n, MPI_BYTE, T_RIGHT, tag, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
        T_CUDA(cudaStreamEndCapture(stream, &graph));
    } else {
        cudaGraphNode_t a, b, c;
        T_OK(MPIX_Isend_enqueue(d_out, (int)n, MPI_BYTE, T_RIGHT, tag, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_GRAPH, &parts[0]));
        T_OK(MPIX_Irecv_enqueue(d_in, (int)n, MPI_BYTE, T_LEFT, tag, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_GRAPH, &parts[1]));
        T_OK(MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_GRAPH, &parts[2]));
        T_CUDA(cudaGraphCreate(&graph, 0));
        T_CUDA(cudaGraphAddChildGraphNode(&a, graph, NULL, 0, parts[0]));
        T_CUDA(cudaGraphAddChildGraphNode(&b, graph, NULL, 0, parts[1]));      /* send and receive are independent branches */
        cudaGraphNode_t deps[2] = {a, b};
        T_CUDA(cudaGraphAddChildGraphNode(&c, graph, deps, 2, parts[2]));
    }
    T_CHECK(rq[0] == MPIX_REQUEST_NULL && rq[1] == MPIX_REQUEST_NULL, "handles must be consumed by the in-graph wait");
    T_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    size_t bad_total = 0;
    for (int it = 0; it < LAUNCHES; it++) {
        for (size_t i = 0; i < n; i++) h[i] = pat(i, t_rank * 64 + it);
        T_CUDA(cudaMemcpyAsync(d_out, h, n, cudaMemcpyHostToDevice, stream));
        T_CUDA(cudaMemsetAsync(d_in, 0xa5, n, stream));
        T_CUDA(cudaGraphLaunch(exec, stream));
        T_CUDA(cudaMemcpyAsync(h, d_in, n, cudaMemcpyDeviceToHost, stream));
        T_CUDA(cudaStreamSynchronize(stream));
        size_t bad = 0;
        for (size_t i = 0; i < n; i++) if (h[i] != pat(i, T_LEFT * 64 + it)) bad++;
        if (bad) { T_CHECK(0, "mode %d heap %d n=%zu launch %d: %zu wrong bytes (first got 0x%02x)", mode, heap, n, it, bad, h[0]); bad_total += bad; break; }
    }
    (void)st; (void)bad_total;
    T_CUDA(cudaGraphExecDestroy(exec));
    T_CUDA(cudaGraphDestroy(graph));
    if (mode == 1) for (int i = 0; i < 3; i++) T_CUDA(cudaGraphDestroy(parts[i]));
    T_CUDA(cudaStreamDestroy(stream));
    if (heap) { T_OK(MPIX_Free_mem(d_out)); T_OK(MPIX_Free_mem(d_in)); } else { T_CUDA(cudaFree(d_out)); T_CUDA(cudaFree(d_in)); }
    free(h);
}

/* Token ring captured with PAGEABLE host buffers (static storage, like the stack variables of the reference's
 * ring-all-graph test): the library pins them in place inside the capture. cudaHostRegister is a "potentially unsafe" call
 * in a Global-mode capture; the library relaxes the thread's capture mode around its own housekeeping calls. */
static void pageable_in_capture(void) {
    static int pageable[2 * 1024];   /* two words a page apart */
    int *send = &pageable[0], *recv = &pageable[1024];
    *send = 700 + t_rank; *recv = -1;
    cudaStream_t stream;
    T_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    MPIX_Request rq[2];
    T_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeGlobal));
    T_OK(MPIX_Isend_enqueue(send, 1, MPI_INT, T_RIGHT, 39, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Irecv_enqueue(recv, 1, MPI_INT, T_LEFT, 39, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
    T_CUDA(cudaStreamEndCapture(stream, &graph));
    T_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    for (int hop = 0; hop < t_size; hop++) {
        T_CUDA(cudaGraphLaunch(exec, stream));
        T_CUDA(cudaMemcpyAsync(send, recv, sizeof(int), cudaMemcpyHostToHost, stream));
    }
    T_CUDA(cudaStreamSynchronize(stream));
    T_CHECK(*recv == 700 + t_rank, "pageable buffers in a capture: token came back as %d", *recv);
    T_CUDA(cudaGraphExecDestroy(exec));
    T_CUDA(cudaGraphDestroy(graph));
    T_CUDA(cudaStreamDestroy(stream));
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 1);
    const size_t sizes[] = {20000, 200000, 1u << 20};   /* above eager: trigger + mover; larger ones take the fused kernel on a full-size GPU */
    int tag = 40;
    for (int mode = 0; mode < 2; mode++)
        for (int heap = 1; heap >= 0; heap--)
            for (unsigned s = 0; s < sizeof sizes / sizeof sizes[0]; s++) run(mode, heap, sizes[s], tag++);
    pageable_in_capture();
    /* graph-owned slots come back through cudaUserObject destructors, which CUDA may run on an internal thread some time
       after cudaGraphDestroy / cudaGraphExecDestroy have returned (first seen on the round-1 driver box): poll, bounded */
    MPIX_Acx_stats s;
    MPIX_Get_stats(&s);
    for (int i = 0; i < 5000 && s.slots_in_use != 0; i++) { usleep(1000); MPIX_Get_stats(&s); }
    T_CHECK(s.slots_in_use == 0, "%llu request slots still held after every graph was destroyed", s.slots_in_use);
    return t_finish("graph_large");
}
