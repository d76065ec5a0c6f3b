/* A token travels round the ring `laps` times with the blocking-on-the-stream calls: MPIX_Recv_enqueue (one fused
 * post+wait kernel for a small message) and MPIX_Send_enqueue (one trigger kernel). The host enqueues everything up
 * front and synchronizes once; between hops only the GPUs talk.
 *   make examples && bin/acxrun -np 4 build/examples/token_ring */
#include <mpi.h>
#include <mpi-acx.h>
#include <stdio.h>
#include <string.h>

int main(int argc, char **argv) {
    int provided, rank, size, ndev;
    MPI_Init_thread(&argc, &argv, MPI_THREAD_MULTIPLE, &provided);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) cudaSetDevice(rank % ndev);   /* no GPU: the library falls back to its host transport */
    if (MPIX_Init()) return 1;

    const int laps = 10, left = (rank + size - 1) % size, right = (rank + 1) % size;
    long long *token;                     /* one 8-byte token in the symmetric heap */
    MPIX_Alloc_mem(sizeof *token, MPI_INFO_NULL, &token);
    if (MPIX_Is_device_transport()) cudaMemset(token, 0, sizeof *token); else *token = 0;
    cudaStream_t s;
    cudaStreamCreate(&s);
    for (int lap = 0; lap < laps; lap++) {
        if (rank == 0) {                  /* rank 0 starts each lap, everybody else forwards what arrives */
            MPIX_Send_enqueue(token, 1, MPI_LONG_LONG, right, lap, MPI_COMM_WORLD, MPIX_QUEUE_CUDA_STREAM, &s);
            MPIX_Recv_enqueue(token, 1, MPI_LONG_LONG, left, lap, MPI_COMM_WORLD, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &s);
        } else {
            MPIX_Recv_enqueue(token, 1, MPI_LONG_LONG, left, lap, MPI_COMM_WORLD, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &s);
            MPIX_Send_enqueue(token, 1, MPI_LONG_LONG, right, lap, MPI_COMM_WORLD, MPIX_QUEUE_CUDA_STREAM, &s);
        }
    }
    cudaStreamSynchronize(s);
    MPIX_Acx_stats st;
    MPIX_Get_stats(&st);
    printf("rank %d: %d laps done, %llu kernels launched (%d hops)\n", rank, laps, st.kernels_launched, laps);
    MPIX_Free_mem(token);
    MPIX_Finalize();
    MPI_Finalize();
    return 0;
}
