/* MPI non-overtaking on one (peer, tag) channel with mixed protocols: an eager message, a rendezvous
 * message and another eager message sent back to back must be received in that order even though
 * the eager ones land immediately (the reference does not guarantee this at all: README "Known
 * Limitations"). Also: two streams sending on different tags concurrently. */
#include "acx_test.h"

int main(int argc, char **argv) {
    t_init(&argc, &argv, 1);
    const int big = 1 << 18;   /* ints: 1 MiB, rendezvous */
    int *s0, *s1, *s2, *r0, *r1, *r2;
    T_OK(MPIX_Alloc_mem(64, MPI_INFO_NULL, &s0)); T_OK(MPIX_Alloc_mem(big * 4, MPI_INFO_NULL, &s1)); T_OK(MPIX_Alloc_mem(64, MPI_INFO_NULL, &s2));
    T_OK(MPIX_Alloc_mem(big * 4, MPI_INFO_NULL, &r0)); T_OK(MPIX_Alloc_mem(big * 4, MPI_INFO_NULL, &r1)); T_OK(MPIX_Alloc_mem(big * 4, MPI_INFO_NULL, &r2));
    int *t_s, *t_r;   /* the second stream's buffers: separate allocations (stream a memsets all of r0) */
    T_OK(MPIX_Alloc_mem(64, MPI_INFO_NULL, &t_s)); T_OK(MPIX_Alloc_mem(64, MPI_INFO_NULL, &t_r));
    int *h = (int *)malloc(big * 4);
    cudaStream_t a, b;
    T_CUDA(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    T_CUDA(cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking));
    for (int round = 0; round < 4; round++) {
        int v0 = 11 + round + 100 * t_rank, v2 = 33 + round + 100 * t_rank;
        for (int i = 0; i < big; i++) h[i] = 22 + round + 100 * t_rank + i;
        T_CUDA(cudaMemcpyAsync(s0, &v0, 4, cudaMemcpyHostToDevice, a));
        T_CUDA(cudaMemcpyAsync(s1, h, big * 4, cudaMemcpyHostToDevice, a));
        T_CUDA(cudaMemcpyAsync(s2, &v2, 4, cudaMemcpyHostToDevice, a));
        T_CUDA(cudaMemsetAsync(r0, 0, big * 4, a)); T_CUDA(cudaMemsetAsync(r1, 0, big * 4, a)); T_CUDA(cudaMemsetAsync(r2, 0, big * 4, a));
        MPIX_Request sq[3], rq[3];
        MPI_Status st[3];
        /* sends: eager(4 B), rendezvous(1 MiB), eager(4 B) — same tag */
        T_OK(MPIX_Isend_enqueue(s0, 1, MPI_INT, T_RIGHT, 9, MPI_COMM_WORLD, &sq[0], MPIX_QUEUE_CUDA_STREAM, &a));
        T_OK(MPIX_Isend_enqueue(s1, big, MPI_INT, T_RIGHT, 9, MPI_COMM_WORLD, &sq[1], MPIX_QUEUE_CUDA_STREAM, &a));
        T_OK(MPIX_Isend_enqueue(s2, 1, MPI_INT, T_RIGHT, 9, MPI_COMM_WORLD, &sq[2], MPIX_QUEUE_CUDA_STREAM, &a));
        /* receives all with full capacity: sizes come from the matching order alone */
        for (int k = 0; k < 3; k++)
            T_OK(MPIX_Irecv_enqueue(k == 0 ? r0 : k == 1 ? r1 : r2, big, MPI_INT, T_LEFT, 9, MPI_COMM_WORLD, &rq[k], MPIX_QUEUE_CUDA_STREAM, &a));
        T_OK(MPIX_Waitall_enqueue(3, rq, st, MPIX_QUEUE_CUDA_STREAM, &a));
        T_OK(MPIX_Waitall_enqueue(3, sq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &a));
        /* a second stream talks on another tag at the same time */
        MPIX_Request q2[2];
        int v9 = 900 + round + t_rank;
        T_CUDA(cudaMemcpyAsync(t_s, &v9, 4, cudaMemcpyHostToDevice, b));
        T_OK(MPIX_Isend_enqueue(t_s, 1, MPI_INT, T_RIGHT, 10, MPI_COMM_WORLD, &q2[0], MPIX_QUEUE_CUDA_STREAM, &b));
        T_OK(MPIX_Irecv_enqueue(t_r, 1, MPI_INT, T_LEFT, 10, MPI_COMM_WORLD, &q2[1], MPIX_QUEUE_CUDA_STREAM, &b));
        T_OK(MPIX_Waitall_enqueue(2, q2, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &b));
        T_CUDA(cudaStreamSynchronize(a));
        T_CUDA(cudaStreamSynchronize(b));
        int g0, g2, g9, cnt[3];
        T_CUDA(cudaMemcpy(&g0, r0, 4, cudaMemcpyDeviceToHost));
        T_CUDA(cudaMemcpy(h, r1, big * 4, cudaMemcpyDeviceToHost));
        T_CUDA(cudaMemcpy(&g2, r2, 4, cudaMemcpyDeviceToHost));
        T_CUDA(cudaMemcpy(&g9, t_r, 4, cudaMemcpyDeviceToHost));
        for (int k = 0; k < 3; k++) MPI_Get_count(&st[k], MPI_INT, &cnt[k]);
        int bad = 0;
        for (int i = 0; i < big; i++) if (h[i] != 22 + round + 100 * T_LEFT + i) bad++;
        T_CHECK(g0 == 11 + round + 100 * T_LEFT && g2 == 33 + round + 100 * T_LEFT && bad == 0, "round %d order broken: %d %d bad=%d", round, g0, g2, bad);
        T_CHECK(cnt[0] == 1 && cnt[1] == big && cnt[2] == 1, "counts %d %d %d", cnt[0], cnt[1], cnt[2]);
        T_CHECK(g9 == 900 + round + T_LEFT, "second stream got %d", g9);
    }
    free(h);
    MPIX_Free_mem(t_s); MPIX_Free_mem(t_r);
    MPIX_Free_mem(s0); MPIX_Free_mem(s1); MPIX_Free_mem(s2); MPIX_Free_mem(r0); MPIX_Free_mem(r1); MPIX_Free_mem(r2);
    cudaStreamDestroy(a); cudaStreamDestroy(b);
    return t_finish("ordering");
}
