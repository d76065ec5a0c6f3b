/* Explicit graph construction: the library returns one graph per operation
 * (MPIX_QUEUE_CUDA_GRAPH), the user composes them as child-graph nodes with explicit
 * dependencies send -> recv -> wait and relaunches (behaviour of the reference's
 * test/src/ring-all-graph-construction.c). Also covers MPIX_Wait_enqueue in GRAPH mode,
 * which the reference gets wrong (defect D2) and never tests. */
#include "acx_test.h"

static void run(int use_waitall) {
    int *send, *recv;
    T_CUDA(cudaMallocHost((void **)&send, sizeof(int)));
    T_CUDA(cudaMallocHost((void **)&recv, sizeof(int)));
    *send = 900 + t_rank; *recv = -1;
    cudaStream_t stream;
    T_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    cudaGraph_t g_send, g_recv, g_wait[2], graph;
    cudaGraphNode_t n_send, n_recv, n_wait;
    cudaGraphExec_t exec;
    MPIX_Request rq[2];
    int nwait = use_waitall ? 1 : 2;
    T_OK(MPIX_Isend_enqueue(send, 1, MPI_INT, T_RIGHT, 4, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_GRAPH, &g_send));
    T_OK(MPIX_Irecv_enqueue(recv, 1, MPI_INT, T_LEFT, 4, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_GRAPH, &g_recv));
    if (use_waitall) T_OK(MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_GRAPH, &g_wait[0]));
    else {
        T_OK(MPIX_Wait_enqueue(&rq[0], MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_GRAPH, &g_wait[0]));
        T_OK(MPIX_Wait_enqueue(&rq[1], MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_GRAPH, &g_wait[1]));
    }
    T_CUDA(cudaGraphCreate(&graph, 0));
    T_CUDA(cudaGraphAddChildGraphNode(&n_send, graph, NULL, 0, g_send));
    T_CUDA(cudaGraphAddChildGraphNode(&n_recv, graph, &n_send, 1, g_recv));
    cudaGraphNode_t dep = n_recv;
    for (int i = 0; i < nwait; i++) { T_CUDA(cudaGraphAddChildGraphNode(&n_wait, graph, &dep, 1, g_wait[i])); dep = n_wait; }
    T_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    for (int hop = 0; hop < t_size; hop++) {
        T_CUDA(cudaGraphLaunch(exec, stream));
        T_CUDA(cudaMemcpyAsync(send, recv, sizeof(int), cudaMemcpyHostToHost, stream));
    }
    T_CUDA(cudaStreamSynchronize(stream));
    T_CHECK(*recv == 900 + t_rank, "token came back as %d (waitall=%d)", *recv, use_waitall);
    T_CUDA(cudaGraphExecDestroy(exec));
    T_CUDA(cudaGraphDestroy(graph));
    T_CUDA(cudaGraphDestroy(g_send));
    T_CUDA(cudaGraphDestroy(g_recv));
    for (int i = 0; i < nwait; i++) T_CUDA(cudaGraphDestroy(g_wait[i]));
    T_CUDA(cudaStreamDestroy(stream));
    cudaFreeHost(send); cudaFreeHost(recv);
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 1);
    run(1);
    run(0);
    return t_finish("ring_graph_construction");
}
