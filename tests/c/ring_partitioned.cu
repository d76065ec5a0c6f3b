// Partitioned ring: Psend/Precv init, device Pready / Parrived from user kernels, 10
// partitions x 10 iterations of Startall -> ready kernel -> arrival-poll kernel -> sync ->
// host Waitall, persistent reuse, frees (behaviour of the reference's
// test/src/ring-partitioned.cu). Run twice: host stack buffers (copy-in/copy-out through
// the landing zone) and symmetric-heap buffers (zero-copy), plus the host Pready/Parrived
// entry points the reference never tests.
#include "acx_test.h"

#define NPART 10
#define NITER 10
#define PER 3   /* ints per partition */

__global__ void mark_all_ready(MPIX_Prequest preq) { MPIX_Pready(threadIdx.x, preq); }

__global__ void poll_until_arrived(MPIX_Prequest preq) {
    int flag = 0;
    while (!flag) MPIX_Parrived(preq, threadIdx.x, &flag);
}

static void run(int use_heap, int host_calls) {
    int stack_send[NPART * PER], stack_recv[NPART * PER];
    int *send = stack_send, *recv = stack_recv;
    if (use_heap) { T_OK(MPIX_Alloc_mem(sizeof stack_send, MPI_INFO_NULL, &send)); T_OK(MPIX_Alloc_mem(sizeof stack_recv, MPI_INFO_NULL, &recv)); }
    /* staging copies live off the stack: the library pins the pages of host communication buffers in place, and CUDA
     * rejects a cudaMemcpy whose host range is only partially pinned (a neighbour on the same stack page could be) */
    static int h_send[NPART * PER], h_recv[NPART * PER];
    MPIX_Request rq[2];
    MPIX_Prequest pq[2];
    MPI_Status sts[2];
    T_OK(MPIX_Psend_init(send, NPART, PER, MPI_INT, T_RIGHT, 11 + use_heap, MPI_COMM_WORLD, MPI_INFO_NULL, &rq[0]));
    T_OK(MPIX_Precv_init(recv, NPART, PER, MPI_INT, T_LEFT, 11 + use_heap, MPI_COMM_WORLD, MPI_INFO_NULL, &rq[1]));
    T_OK(MPIX_Prequest_create(rq[0], &pq[0]));
    T_OK(MPIX_Prequest_create(rq[1], &pq[1]));
    for (int it = 0; it < NITER; it++) {
        for (int i = 0; i < NPART * PER; i++) { h_send[i] = t_rank * 1000 + it * 100 + i; h_recv[i] = -1; }
        if (use_heap) { T_CUDA(cudaMemcpy(send, h_send, sizeof h_send, cudaMemcpyHostToDevice)); T_CUDA(cudaMemcpy(recv, h_recv, sizeof h_recv, cudaMemcpyHostToDevice)); }
        else { memcpy(send, h_send, sizeof h_send); memcpy(recv, h_recv, sizeof h_recv); }
        T_OK(MPIX_Startall(2, rq));
        if (host_calls) {
            for (int p = NPART - 1; p >= 0; p--) T_OK(MPIX_Pready(p, &rq[0]));
            for (int p = 0; p < NPART; p++) { int f = 0; while (!f) T_OK(MPIX_Parrived(&rq[1], p, &f)); }
        } else {
            T_LAUNCH(mark_all_ready, 1, NPART, pq[0]);
            T_LAUNCH(poll_until_arrived, 1, NPART, pq[1]);
            T_CUDA(cudaStreamSynchronize(0));
        }
        T_OK(MPIX_Waitall(2, rq, sts));
        if (use_heap) T_CUDA(cudaMemcpy(h_recv, recv, sizeof h_recv, cudaMemcpyDeviceToHost)); else memcpy(h_recv, recv, sizeof h_recv);
        int bad = 0;
        for (int i = 0; i < NPART * PER; i++) if (h_recv[i] != T_LEFT * 1000 + it * 100 + i) bad++;
        T_CHECK(bad == 0, "heap=%d host=%d iter %d: %d wrong (recv[0]=%d)", use_heap, host_calls, it, bad, h_recv[0]);
    }
    T_OK(MPIX_Prequest_free(&pq[0]));
    T_OK(MPIX_Prequest_free(&pq[1]));
    T_OK(MPIX_Request_free(&rq[0]));
    T_OK(MPIX_Request_free(&rq[1]));
    T_CHECK(rq[0] == MPIX_REQUEST_NULL && pq[0] == MPIX_PREQUEST_NULL, "handles not nulled");
    if (use_heap) { T_OK(MPIX_Free_mem(send)); T_OK(MPIX_Free_mem(recv)); }
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 1);
    run(0, 0);
    run(1, 0);
    run(0, 1);
    return t_finish("ring_partitioned");
}
