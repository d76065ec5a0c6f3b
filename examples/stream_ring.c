/* Minimal stream-triggered exchange: every rank sends one buffer to its right neighbour on a CUDA stream and waits
 * for the one from its left — the calls return immediately, the GPU does the rest.
 *   make examples && bin/acxrun -np 2 build/examples/stream_ring */
#include <mpi.h>
#include <mpi-acx.h>
#include <stdio.h>

int main(int argc, char **argv) {
    int provided, rank, size, ndev;
    MPI_Init_thread(&argc, &argv, MPI_THREAD_MULTIPLE, &provided);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) cudaSetDevice(rank % ndev);   /* no GPU: the library falls back to its host transport */
    if (MPIX_Init()) return 1;

    const int n = 1 << 20;
    float *out, *in;                       /* symmetric-heap buffers: the send kernel stores straight into the peer's HBM */
    MPIX_Alloc_mem(n * sizeof(float), MPI_INFO_NULL, &out);
    MPIX_Alloc_mem(n * sizeof(float), MPI_INFO_NULL, &in);
    cudaStream_t s;
    cudaStreamCreate(&s);
    cudaMemsetAsync(out, rank + 1, n * sizeof(float), s);      /* producer work, ordered before the send on the same stream */

    MPIX_Request req[2];
    MPIX_Irecv_enqueue(in, n, MPI_FLOAT, (rank + size - 1) % size, 0, MPI_COMM_WORLD, &req[0], MPIX_QUEUE_CUDA_STREAM, &s);
    MPIX_Isend_enqueue(out, n, MPI_FLOAT, (rank + 1) % size, 0, MPI_COMM_WORLD, &req[1], MPIX_QUEUE_CUDA_STREAM, &s);
    MPIX_Waitall_enqueue(2, req, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &s);
    /* consumer kernels enqueued on `s` from here on see the received data */
    cudaStreamSynchronize(s);

    unsigned char first;
    cudaMemcpy(&first, in, 1, cudaMemcpyDeviceToHost);
    printf("rank %d received bytes of value %d from rank %d\n", rank, first, (rank + size - 1) % size);
    MPIX_Free_mem(out); MPIX_Free_mem(in);
    MPIX_Finalize();
    MPI_Finalize();
    return 0;
}
