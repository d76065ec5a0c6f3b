/* Message sizes from 0 B to 64 MiB across the eager / rendezvous boundary, symmetric-heap
 * (zero-copy) and private cudaMalloc (bounced) buffers, both arrival orders (send first /
 * recv first), several messages in flight per peer, two tags interleaved out of order.
 * None of this is covered by the reference's tests (largest message there: 40 B). */
#include "acx_test.h"

static unsigned char pat(size_t i, int salt) { return (unsigned char)((i * 131 + (i >> 8) * 7 + salt) & 0xff); }

int main(int argc, char **argv) {
    t_init(&argc, &argv, 1);
    size_t maxb = 64u << 20;
    if (argc > 1) maxb = (size_t)atol(argv[1]);
    unsigned char *h = (unsigned char *)malloc(maxb + 16);
    cudaStream_t stream;
    T_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    size_t sizes[] = {0, 1, 7, 8, 100, 4096, 8192, 8193, 65536, 1000003, 16u << 20, maxb};
    for (int heap = 0; heap < 2; heap++) {
        unsigned char *d_out, *d_in;
        if (heap) { T_OK(MPIX_Alloc_mem(maxb + 16, MPI_INFO_NULL, &d_out)); T_OK(MPIX_Alloc_mem(maxb + 16, MPI_INFO_NULL, &d_in)); }
        else { T_CUDA(cudaMalloc((void **)&d_out, maxb + 16)); T_CUDA(cudaMalloc((void **)&d_in, maxb + 16)); }
        for (unsigned si = 0; si < sizeof sizes / sizeof sizes[0]; si++) {
            size_t n = sizes[si] > maxb ? maxb : sizes[si];
            int salt = t_rank * 17 + (int)si;
            for (size_t i = 0; i < n; i++) h[i] = pat(i, salt);
            T_CUDA(cudaMemcpyAsync(d_out, h, n, cudaMemcpyHostToDevice, stream));   /* same stream as the ops: ordered before them */
            T_CUDA(cudaMemsetAsync(d_in, 0xee, n + 16, stream));
            MPIX_Request rq[2];
            MPI_Status st;
            if (si & 1) {
                T_OK(MPIX_Irecv_enqueue(d_in, (int)n, MPI_BYTE, T_LEFT, 5, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
                T_OK(MPIX_Isend_enqueue(d_out, (int)n, MPI_BYTE, T_RIGHT, 5, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
            } else {
                T_OK(MPIX_Isend_enqueue(d_out, (int)n, MPI_BYTE, T_RIGHT, 5, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
                T_OK(MPIX_Irecv_enqueue(d_in, (int)n, MPI_BYTE, T_LEFT, 5, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
            }
            T_OK(MPIX_Wait_enqueue(&rq[0], MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
            T_OK(MPIX_Wait_enqueue(&rq[1], &st, MPIX_QUEUE_CUDA_STREAM, &stream));
            T_CUDA(cudaStreamSynchronize(stream));
            T_CUDA(cudaMemcpy(h, d_in, n + 16, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            int lsalt = T_LEFT * 17 + (int)si;
            for (size_t i = 0; i < n; i++) if (h[i] != pat(i, lsalt)) bad++;
            for (size_t i = n; i < n + 16; i++) if (h[i] != 0xee) bad++;   /* no overrun */
            int cnt = -1;
            MPI_Get_count(&st, MPI_BYTE, &cnt);
            T_CHECK(bad == 0 && cnt == (int)n && st.MPI_SOURCE == T_LEFT && st.MPI_TAG == 5, "heap=%d n=%zu: %zu bad bytes, count %d", heap, n, bad, cnt);
        }
        /* several in flight, two tags waited out of order, shorter message than capacity */
        {
            enum { K = 6, LEN = 3000, CAP = 5000 };
            MPIX_Request sq[2 * K], rq[2 * K];
            for (int i = 0; i < 2 * K * LEN; i++) h[i] = pat(i, t_rank + 99);
            T_CUDA(cudaMemcpyAsync(d_out, h, 2 * K * LEN, cudaMemcpyHostToDevice, stream));
            T_CUDA(cudaMemsetAsync(d_in, 0, 2 * K * CAP, stream));
            for (int k = 0; k < K; k++) for (int tg = 0; tg < 2; tg++)
                T_OK(MPIX_Isend_enqueue(d_out + (size_t)(2 * k + tg) * LEN, LEN, MPI_BYTE, T_RIGHT, 20 + tg, MPI_COMM_WORLD, &sq[2 * k + tg], MPIX_QUEUE_CUDA_STREAM, &stream));
            for (int tg = 1; tg >= 0; tg--) for (int k = 0; k < K; k++)
                T_OK(MPIX_Irecv_enqueue(d_in + (size_t)(2 * k + tg) * CAP, CAP, MPI_BYTE, T_LEFT, 20 + tg, MPI_COMM_WORLD, &rq[2 * k + tg], MPIX_QUEUE_CUDA_STREAM, &stream));
            T_OK(MPIX_Waitall_enqueue(2 * K, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
            T_OK(MPIX_Waitall_enqueue(2 * K, sq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
            T_CUDA(cudaStreamSynchronize(stream));
            unsigned char *hb = (unsigned char *)malloc(2 * K * CAP);
            T_CUDA(cudaMemcpy(hb, d_in, 2 * K * CAP, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            for (int m = 0; m < 2 * K; m++) for (int i = 0; i < LEN; i++) if (hb[(size_t)m * CAP + i] != pat((size_t)m * LEN + i, T_LEFT + 99)) bad++;
            T_CHECK(bad == 0, "heap=%d multi-in-flight: %zu bad bytes", heap, bad);
            free(hb);
        }
        if (heap) { T_OK(MPIX_Free_mem(d_out)); T_OK(MPIX_Free_mem(d_in)); } else { T_CUDA(cudaFree(d_out)); T_CUDA(cudaFree(d_in)); }
    }
    MPIX_Acx_stats s;
    MPIX_Get_stats(&s);
    T_CHECK(s.eager_msgs > 0 && s.rendezvous_msgs > 0 && s.bounce_copies > 0, "stats eager %llu rndv %llu bounce %llu", s.eager_msgs, s.rendezvous_msgs, s.bounce_copies);
    free(h);
    T_CUDA(cudaStreamDestroy(stream));
    return t_finish("large_msgs");
}
