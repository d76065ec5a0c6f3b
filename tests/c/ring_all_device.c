/* Device (cudaMalloc) buffers: H2D copy ordered before the enqueued send on the same
 * stream, completion by MPIX_Waitall_enqueue ON THE STREAM (the reference's
 * ring-all-device.c has to fall back to a host wait because its MPI may need the blocked
 * stream, README caveat D8 — the kernel transport here has no such dependency), D2H after.
 * Also exercises symmetric-heap buffers from MPIX_Alloc_mem. */
#include "acx_test.h"

#define N 4096

int main(int argc, char **argv) {
    t_init(&argc, &argv, 1);
    cudaStream_t stream;
    T_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int *h_out = (int *)malloc(N * sizeof(int)), *h_in = (int *)malloc(N * sizeof(int));
    for (int pass = 0; pass < 2; pass++) {
        int *d_out, *d_in;
        if (pass == 0) { T_CUDA(cudaMalloc((void **)&d_out, N * sizeof(int))); T_CUDA(cudaMalloc((void **)&d_in, N * sizeof(int))); }
        else { T_OK(MPIX_Alloc_mem(N * sizeof(int), MPI_INFO_NULL, &d_out)); T_OK(MPIX_Alloc_mem(N * sizeof(int), MPI_INFO_NULL, &d_in)); }
        for (int count = 1; count <= N; count *= 8) {   /* 4 B .. 16 KiB: eager and rendezvous */
            for (int i = 0; i < count; i++) { h_out[i] = t_rank * 100000 + count * 3 + i; h_in[i] = -1; }
            MPIX_Request rq[2];
            T_CUDA(cudaMemcpyAsync(d_out, h_out, count * sizeof(int), cudaMemcpyHostToDevice, stream));
            T_CUDA(cudaMemsetAsync(d_in, 0xff, count * sizeof(int), stream));
            T_OK(MPIX_Isend_enqueue(d_out, count, MPI_INT, T_RIGHT, 3, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
            T_OK(MPIX_Irecv_enqueue(d_in, count, MPI_INT, T_LEFT, 3, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
            if (count & 8) { T_OK(MPIX_Waitall(2, rq, MPI_STATUSES_IGNORE)); }   /* host wait variant */
            else T_OK(MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
            T_CUDA(cudaMemcpyAsync(h_in, d_in, count * sizeof(int), cudaMemcpyDeviceToHost, stream));
            T_CUDA(cudaStreamSynchronize(stream));
            int bad = 0;
            for (int i = 0; i < count; i++) if (h_in[i] != T_LEFT * 100000 + count * 3 + i) bad++;
            T_CHECK(bad == 0, "pass %d count %d: %d wrong elements (first %d)", pass, count, bad, h_in[0]);
        }
        if (pass == 0) { T_CUDA(cudaFree(d_out)); T_CUDA(cudaFree(d_in)); }
        else { T_OK(MPIX_Free_mem(d_out)); T_OK(MPIX_Free_mem(d_in)); }
    }
    free(h_out); free(h_in);
    T_CUDA(cudaStreamDestroy(stream));
    return t_finish("ring_all_device");
}
