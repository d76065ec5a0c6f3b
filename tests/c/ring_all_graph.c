/* Stream capture of Isend + Irecv + Waitall into a CUDA graph, relaunched world_size times
 * with a host-to-host copy of the received token into the send slot between launches, so
 * each token travels the whole ring and comes home (the behaviour of the reference's
 * test/src/ring-all-graph.c). Graph destruction returns the library resources. */
#include "acx_test.h"

int main(int argc, char **argv) {
    t_init(&argc, &argv, 1);
    int *send, *recv;   /* pinned so the in-graph H2H copy is a plain memcpy node */
    T_CUDA(cudaMallocHost((void **)&send, sizeof(int)));
    T_CUDA(cudaMallocHost((void **)&recv, sizeof(int)));
    *send = 500 + t_rank; *recv = -1;
    cudaStream_t stream;
    T_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    MPIX_Request rq[2];
    T_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeGlobal));
    T_OK(MPIX_Isend_enqueue(send, 1, MPI_INT, T_RIGHT, 0, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Irecv_enqueue(recv, 1, MPI_INT, T_LEFT, 0, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
    T_CUDA(cudaStreamEndCapture(stream, &graph));
    T_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    for (int hop = 0; hop < t_size; hop++) {
        T_CUDA(cudaGraphLaunch(exec, stream));
        T_CUDA(cudaMemcpyAsync(send, recv, sizeof(int), cudaMemcpyHostToHost, stream));
    }
    T_CUDA(cudaStreamSynchronize(stream));
    T_CHECK(*recv == 500 + t_rank, "token came back as %d", *recv);
    T_CUDA(cudaGraphExecDestroy(exec));
    T_CUDA(cudaGraphDestroy(graph));
    T_CUDA(cudaStreamDestroy(stream));
    cudaFreeHost(send); cudaFreeHost(recv);
    return t_finish("ring_all_graph");
}
