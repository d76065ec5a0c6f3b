/* Registration cache: plain cudaMalloc buffers start on the bounce path and switch to zero-copy
 * (CUDA-IPC mapped at the peer) once the peer has acknowledged the mapping. Checks payload on every
 * round and that bounce copies stop growing after the first rounds. */
#include "acx_test.h"

int main(int argc, char **argv) {
    t_init(&argc, &argv, 1);
    const size_t n = 4u << 20;
    unsigned *d_out, *d_in, *h = (unsigned *)malloc(n);
    T_CUDA(cudaMalloc((void **)&d_out, n));
    T_CUDA(cudaMalloc((void **)&d_in, n));
    cudaStream_t stream;
    T_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    unsigned long long bounce_after3 = 0;
    for (int round = 0; round < 8; round++) {
        for (size_t i = 0; i < n / 4; i++) h[i] = (unsigned)(i * 2654435761u) ^ (unsigned)(t_rank * 977 + round);
        T_CUDA(cudaMemcpyAsync(d_out, h, n, cudaMemcpyHostToDevice, stream));
        T_CUDA(cudaMemsetAsync(d_in, 0, n, stream));
        MPIX_Request rq[2];
        T_OK(MPIX_Irecv_enqueue(d_in, (int)n, MPI_BYTE, T_LEFT, 8, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Isend_enqueue(d_out, (int)n, MPI_BYTE, T_RIGHT, 8, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
        T_CUDA(cudaMemcpyAsync(h, d_in, n, cudaMemcpyDeviceToHost, stream));
        T_CUDA(cudaStreamSynchronize(stream));
        size_t bad = 0;
        for (size_t i = 0; i < n / 4; i++) if (h[i] != ((unsigned)(i * 2654435761u) ^ (unsigned)(T_LEFT * 977 + round))) bad++;
        T_CHECK(bad == 0, "round %d: %zu wrong words", round, bad);
        MPI_Barrier(MPI_COMM_WORLD);   /* lets both sides observe each other's acks between rounds */
        MPIX_Acx_stats s;
        MPIX_Get_stats(&s);
        if (round == 3) bounce_after3 = s.bounce_copies;
        if (round == 7) T_CHECK(s.bounce_copies == bounce_after3, "still bouncing after registration: %llu -> %llu", bounce_after3, s.bounce_copies);
    }
    free(h); cudaFree(d_out); cudaFree(d_in);
    T_CUDA(cudaStreamDestroy(stream));
    return t_finish("regcache");
}
