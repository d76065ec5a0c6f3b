/* More receives in flight to one peer on one tag than a channel's ring holds (ACX_DEPTH = 32). The reference allows
 * MPIACX_NFLAGS = 4096 operations in flight (include/mpi-acx-internal.h:141); round 1 of this library rejected the 33rd
 * receive. Now a receive beyond the ring's depth is kept and its trigger goes out — in posting order — from the wait
 * that makes room:
 *   1. 100 receives posted up front, then the matching sends in bursts of 25, each burst followed by the waits of its receives
 *      and sends (sizes cross the eager limit) — messages SENT and not yet consumed stay within the ring: a send's trigger
 *      waits for slot credit like a standard-mode MPI_Send waits for buffer space;
 *   2. the same with host MPIX_Waitall;
 *   3. waiting FIRST on a receive that sits behind 32 un-waited ones is an error (nothing can make room), not a deadlock;
 *      waiting in order afterwards works. */
#include "acx_test.h"

#define K 100
static int len_of(int k) { return k % 17 == 0 ? 5000 : 1 + (k % 7); }   /* ints: a few rendezvous messages among eager ones */

static void round_trip(int host_wait, cudaStream_t st) {
    static int *sbuf[K], *rbuf[K];
    MPIX_Request rq[2 * K];
    for (int k = 0; k < K; k++) {
        if (!sbuf[k]) { sbuf[k] = (int *)malloc(sizeof(int) * 5000); rbuf[k] = (int *)malloc(sizeof(int) * 5000); }
        for (int i = 0; i < len_of(k); i++) { sbuf[k][i] = t_rank * 1000000 + k * 100 + (i % 89) + host_wait; rbuf[k][i] = -1; }
    }
    for (int k = 0; k < K; k++) T_OK(MPIX_Irecv_enqueue(rbuf[k], 5000, MPI_INT, T_LEFT, 4, MPI_COMM_WORLD, &rq[k], MPIX_QUEUE_CUDA_STREAM, &st));
    for (int lo = 0; lo < K; lo += 25) {
        for (int k = lo; k < lo + 25; k++) T_OK(MPIX_Isend_enqueue(sbuf[k], len_of(k), MPI_INT, T_RIGHT, 4, MPI_COMM_WORLD, &rq[K + k], MPIX_QUEUE_CUDA_STREAM, &st));
        if (host_wait) { T_OK(MPIX_Waitall(25, rq + lo, MPI_STATUSES_IGNORE)); T_OK(MPIX_Waitall(25, rq + K + lo, MPI_STATUSES_IGNORE)); }
        else { T_OK(MPIX_Waitall_enqueue(25, rq + lo, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st)); T_OK(MPIX_Waitall_enqueue(25, rq + K + lo, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st)); }
    }
    if (!host_wait && t_gpu) T_CUDA(cudaStreamSynchronize(st));
    int bad = 0;
    for (int k = 0; k < K; k++) for (int i = 0; i < len_of(k); i++) if (rbuf[k][i] != T_LEFT * 1000000 + k * 100 + (i % 89) + host_wait) { bad++; break; }
    T_CHECK(bad == 0, "%d of %d messages wrong with %d receives outstanding on one channel (%s wait)", bad, K, K, host_wait ? "host" : "stream");
}

static void out_of_order_wait(cudaStream_t st) {
    if (!t_gpu) return;   /* the host transport has no ring */
    enum { N = 40 };
    static int r[N], s[N];
    MPIX_Request rq[N], sq[N];
    for (int k = 0; k < N; k++) { r[k] = -1; s[k] = 500 + k + t_rank; }
    for (int k = 0; k < N; k++) T_OK(MPIX_Irecv_enqueue(&r[k], 1, MPI_INT, T_LEFT, 6, MPI_COMM_WORLD, &rq[k], MPIX_QUEUE_CUDA_STREAM, &st));
    for (int k = 0; k < 20; k++) T_OK(MPIX_Isend_enqueue(&s[k], 1, MPI_INT, T_RIGHT, 6, MPI_COMM_WORLD, &sq[k], MPIX_QUEUE_CUDA_STREAM, &st));
    MPIX_Request last = rq[N - 1];
    int rc = MPIX_Wait_enqueue(&last, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st);
    T_CHECK(rc != MPI_SUCCESS, "waiting first on receive #%d of %d on one channel should be refused (rc %d)", N, N, rc);
    T_OK(MPIX_Waitall_enqueue(20, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
    for (int k = 20; k < N; k++) T_OK(MPIX_Isend_enqueue(&s[k], 1, MPI_INT, T_RIGHT, 6, MPI_COMM_WORLD, &sq[k], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Waitall_enqueue(N - 20, rq + 20, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Waitall_enqueue(N, sq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
    T_CUDA(cudaStreamSynchronize(st));
    int bad = 0;
    for (int k = 0; k < N; k++) if (r[k] != 500 + k + T_LEFT) bad++;
    T_CHECK(bad == 0, "in-order waits after the refused one: %d wrong", bad);
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    cudaStream_t st = 0;
    if (t_gpu) T_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    round_trip(0, st);
    round_trip(1, st);
    out_of_order_wait(st);
    if (t_gpu) T_CUDA(cudaStreamDestroy(st));
    return t_finish("many_outstanding");
}
