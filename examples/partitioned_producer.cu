// Kernel-triggered partitioned communication: a producer kernel writes each output tile straight into the NEXT rank's
// receive buffer and marks the partition ready; the consumer kernel on the next rank starts on a partition as soon as
// it has arrived. No host involvement between MPIX_Startall and MPIX_Waitall.
//   make examples && bin/acxrun -np 2 build/examples/partitioned_producer
#include <mpi.h>
#include <mpi-acx.h>
#include <stdio.h>

constexpr int PARTS = 8, PER = 4096;   // floats per partition

__global__ void produce(MPIX_Prequest send, int rank) {            // one CTA per partition
    float *dst = (float *)acx::remote_partition_ptr(send, blockIdx.x);   // peer memory, mapped over NVLink
    for (int i = threadIdx.x; i < PER; i += blockDim.x) dst[i] = rank * 1000.f + blockIdx.x;
    __syncthreads();
    if (threadIdx.x == 0) acx::pready_signal(blockIdx.x, send);    // release flag in the peer's HBM
}
__global__ void consume(MPIX_Prequest recv, float *sums) {         // one CTA per partition
    if (threadIdx.x == 0) acx::wait_arrived(recv, blockIdx.x);     // acquire on local HBM
    __syncthreads();
    const float *src = (const float *)acx::local_partition_ptr(recv, blockIdx.x);
    float s = 0;
    for (int i = threadIdx.x; i < PER; i += blockDim.x) s += src[i];
    atomicAdd(&sums[blockIdx.x], s);
}

int main(int argc, char **argv) {
    int provided, rank, size, ndev;
    MPI_Init_thread(&argc, &argv, MPI_THREAD_MULTIPLE, &provided);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    cudaGetDeviceCount(&ndev);
    cudaSetDevice(rank % ndev);
    if (MPIX_Init()) return 1;
    float *inbox, *sums;
    MPIX_Alloc_mem(PARTS * PER * sizeof(float), MPI_INFO_NULL, &inbox);   // heap buffer: partitions land in place
    cudaMalloc(&sums, PARTS * sizeof(float));
    cudaMemset(sums, 0, PARTS * sizeof(float));
    MPIX_Request req[2];
    MPIX_Prequest preq[2];
    MPIX_Psend_init(NULL, PARTS, PER, MPI_FLOAT, (rank + 1) % size, 0, MPI_COMM_WORLD, MPI_INFO_NULL, &req[0]);   // zero-copy producer: no send buffer
    MPIX_Precv_init(inbox, PARTS, PER, MPI_FLOAT, (rank + size - 1) % size, 0, MPI_COMM_WORLD, MPI_INFO_NULL, &req[1]);
    MPIX_Prequest_create(req[0], &preq[0]);
    MPIX_Prequest_create(req[1], &preq[1]);
    MPIX_Startall(2, req);
    produce<<<PARTS, 256>>>(preq[0], rank);
    consume<<<PARTS, 256>>>(preq[1], sums);
    cudaDeviceSynchronize();
    MPIX_Waitall(2, req, MPI_STATUSES_IGNORE);
    float h[PARTS];
    cudaMemcpy(h, sums, sizeof h, cudaMemcpyDeviceToHost);
    printf("rank %d: partition sums %.0f ... %.0f (expected %.0f ... %.0f)\n", rank, h[0], h[PARTS - 1],
           PER * (((rank + size - 1) % size) * 1000.f), PER * (((rank + size - 1) % size) * 1000.f + PARTS - 1));
    MPIX_Prequest_free(&preq[0]); MPIX_Prequest_free(&preq[1]);
    MPIX_Request_free(&req[0]); MPIX_Request_free(&req[1]);
    MPIX_Free_mem(inbox); cudaFree(sums);
    MPIX_Finalize();
    MPI_Finalize();
    return 0;
}
