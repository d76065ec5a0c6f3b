/* More receives in flight to one peer on one tag than a channel's ring holds (ACX_DEPTH = 32). The reference allows
 * MPIACX_NFLAGS = 4096 operations in flight (include/mpi-acx-internal.h:141); round 1 of this library rejected the 33rd
 * receive. Now a receive beyond the ring's depth is kept and its trigger goes out — in posting order — from the wait
 * that makes room:
 *   1. 100 receives posted up front, then the matching sends in bursts of 25, each burst followed by the waits of its receives
 *      and sends (sizes cross the eager limit) — messages SENT and not yet consumed stay within the ring: a send's trigger
 *      waits for slot credit like a standard-mode MPI_Send waits for buffer space;
 *   2. the same with host MPIX_Waitall;
 *   3. waiting FIRST on a receive that sits behind 32 un-waited ones is an error (nothing can make room), not a deadlock;
 *      waiting in order afterwards works;
 *   4. the mirror image for SENDS: 80 sends enqueued up front on one channel, then the 80 receives, then ONE wait call for all
 *      160 requests (sends listed first) — on the stream, and on the host. A send's trigger waits for slot credit, which the
 *      receiver returns from the wait of the message 32 tickets earlier: the triggers of the sends beyond the ring wait on pool
 *      streams (MPIX_Acx_stats::overflow_sends) instead of stopping the caller's stream in front of the receives, and a wait
 *      call completes its overflow sends last. Then a second round on the same channel, waiting receives and sends in two calls. */
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

static unsigned long long overflow_count(void) { MPIX_Acx_stats s; MPIX_Get_stats(&s); return s.overflow_sends; }

static void send_burst(int host_wait, int two_calls, cudaStream_t st) {
    enum { KB = 80 };   /* 48 overflow sends */
    static int *sbuf[KB], *rbuf[KB];
    MPIX_Request rq[2 * KB];   /* sends first, then receives */
    const int salt = 7 * host_wait + 3 * two_calls;
    for (int k = 0; k < KB; k++) {
        if (!sbuf[k]) { sbuf[k] = (int *)malloc(sizeof(int) * 5000); rbuf[k] = (int *)malloc(sizeof(int) * 5000); }
        for (int i = 0; i < len_of(k); i++) { sbuf[k][i] = t_rank * 1000000 + k * 100 + (i % 83) + salt; rbuf[k][i] = -1; }
    }
    /* first use of a host buffer pins it (cudaHostRegister): do that in a lock-step round now, not in the middle of the burst
     * while triggers wait on the device (DESIGN.md section 8 on host calls next to waiting kernels) */
    static int warmed;
    if (!warmed) {
        warmed = 1;
        for (int k = 0; k < KB; k++) {
            MPIX_Request w[2];
            T_OK(MPIX_Irecv_enqueue(rbuf[k], 5000, MPI_INT, T_LEFT, 8, MPI_COMM_WORLD, &w[0], MPIX_QUEUE_CUDA_STREAM, &st));
            T_OK(MPIX_Isend_enqueue(sbuf[k], len_of(k), MPI_INT, T_RIGHT, 8, MPI_COMM_WORLD, &w[1], MPIX_QUEUE_CUDA_STREAM, &st));
            T_OK(MPIX_Waitall_enqueue(2, w, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        }
        if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < KB; k++) for (int i = 0; i < len_of(k); i++) rbuf[k][i] = -1;
    }
    const unsigned long long before = overflow_count();
    for (int k = 0; k < KB; k++) T_OK(MPIX_Isend_enqueue(sbuf[k], len_of(k), MPI_INT, T_RIGHT, 8, MPI_COMM_WORLD, &rq[k], MPIX_QUEUE_CUDA_STREAM, &st));
    for (int k = 0; k < KB; k++) T_OK(MPIX_Irecv_enqueue(rbuf[k], 5000, MPI_INT, T_LEFT, 8, MPI_COMM_WORLD, &rq[KB + k], MPIX_QUEUE_CUDA_STREAM, &st));
    const int depth = getenv("ACX_DEPTH") ? atoi(getenv("ACX_DEPTH")) : 32;
    if (t_gpu) T_CHECK(overflow_count() - before == (unsigned long long)(KB - depth), "%llu sends took the overflow path, expected %d (%d un-waited sends on one channel, ring depth %d)", overflow_count() - before, KB - depth, KB, depth);
    if (two_calls) {   /* receives, then sends: the order a program has to use when it splits the wait */
        if (host_wait) { T_OK(MPIX_Waitall(KB, rq + KB, MPI_STATUSES_IGNORE)); T_OK(MPIX_Waitall(KB, rq, MPI_STATUSES_IGNORE)); }
        else { T_OK(MPIX_Waitall_enqueue(KB, rq + KB, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st)); T_OK(MPIX_Waitall_enqueue(KB, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st)); }
    } else if (host_wait) T_OK(MPIX_Waitall(2 * KB, rq, MPI_STATUSES_IGNORE));
    else T_OK(MPIX_Waitall_enqueue(2 * KB, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
    if (!host_wait && t_gpu) T_CUDA(cudaStreamSynchronize(st));
    int bad = 0;
    for (int k = 0; k < KB; k++) for (int i = 0; i < len_of(k); i++) if (rbuf[k][i] != T_LEFT * 1000000 + k * 100 + (i % 83) + salt) { bad++; break; }
    T_CHECK(bad == 0, "%d of %d messages wrong with %d sends in flight on one channel (%s wait, %s)", bad, KB, KB, host_wait ? "host" : "stream", two_calls ? "two calls" : "one call");
    /* the channel is back to normal: a plain exchange behind the burst stays on the caller's stream */
    const unsigned long long mid = overflow_count();
    int a = 4242 + t_rank, b = -1;
    T_OK(MPIX_Sendrecv_enqueue(&a, 1, MPI_INT, T_RIGHT, 8, &b, 1, MPI_INT, T_LEFT, 8, MPI_COMM_WORLD, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
    if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
    T_CHECK(b == 4242 + T_LEFT, "exchange behind the burst delivered %d, expected %d", b, 4242 + T_LEFT);
    if (t_gpu && !host_wait) T_CHECK(overflow_count() == mid, "a single send on a drained channel took the overflow path");
}

/* The same with rendezvous messages above 1 MiB: an overflow send of that size is ONE two-CTA kernel (trigger + mover) on its
 * pool stream; CTA 1 waits for CTA 0's match while that waits for slot credit. 36 messages of 1.25 MiB, one wait call. */
static void big_burst(cudaStream_t st) {
    enum { KG = 36, N = 327680 };
    static int *sbuf[KG], *rbuf[KG];
    MPIX_Request rq[2 * KG];
    for (int k = 0; k < KG; k++) {
        if (!sbuf[k]) { sbuf[k] = (int *)malloc(sizeof(int) * N); rbuf[k] = (int *)malloc(sizeof(int) * N); }
        for (int i = 0; i < N; i++) { sbuf[k][i] = t_rank * 7919 + k * 131 + i; rbuf[k][i] = -1; }
    }
    for (int k = 0; k < KG; k++) {   /* pin the buffers in a lock-step round first (see send_burst) */
        MPIX_Request w[2];
        T_OK(MPIX_Irecv_enqueue(rbuf[k], N, MPI_INT, T_LEFT, 9, MPI_COMM_WORLD, &w[0], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Isend_enqueue(sbuf[k], N, MPI_INT, T_RIGHT, 9, MPI_COMM_WORLD, &w[1], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Waitall_enqueue(2, w, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
    }
    if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < KG; k++) for (int i = 0; i < N; i += 997) rbuf[k][i] = -1;
    const unsigned long long before = overflow_count();
    for (int k = 0; k < KG; k++) T_OK(MPIX_Isend_enqueue(sbuf[k], N, MPI_INT, T_RIGHT, 9, MPI_COMM_WORLD, &rq[k], MPIX_QUEUE_CUDA_STREAM, &st));
    for (int k = 0; k < KG; k++) T_OK(MPIX_Irecv_enqueue(rbuf[k], N, MPI_INT, T_LEFT, 9, MPI_COMM_WORLD, &rq[KG + k], MPIX_QUEUE_CUDA_STREAM, &st));
    const int depth = getenv("ACX_DEPTH") ? atoi(getenv("ACX_DEPTH")) : 32;
    if (t_gpu && depth < KG) T_CHECK(overflow_count() - before == (unsigned long long)(KG - depth), "%llu large sends took the overflow path, expected %d", overflow_count() - before, KG - depth);
    T_OK(MPIX_Waitall_enqueue(2 * KG, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
    if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
    int bad = 0;
    for (int k = 0; k < KG; k++) for (int i = 0; i < N; i += 997) if (rbuf[k][i] != T_LEFT * 7919 + k * 131 + i) { bad++; break; }
    T_CHECK(bad == 0, "%d of %d large messages wrong with %d sends in flight on one channel", bad, KG, KG);
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    cudaStream_t st = 0;
    if (t_gpu) T_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    round_trip(0, st);
    round_trip(1, st);
    out_of_order_wait(st);
    /* part 4 needs the overflow path: not with it switched off, nor in the modes that keep every operation of a peer on one
     * side stream (there the trigger of send 33 waits on that stream, in front of the receives); the GPU suite runs it as a
     * test of its own */
    const int no_overflow = (getenv("ACX_SEND_OVERFLOW") && atoi(getenv("ACX_SEND_OVERFLOW")) == 0) || (getenv("ACX_PEER_BRANCH") && atoi(getenv("ACX_PEER_BRANCH")) == 1) ||
                            (getenv("ACX_SENDER_PUSH") && atoi(getenv("ACX_SENDER_PUSH")) != 0);
    if (!getenv("ACX_TEST_NO_SEND_BURST") && !no_overflow) {
        send_burst(0, 0, st);
        send_burst(1, 0, st);
        send_burst(0, 1, st);
        send_burst(1, 1, st);
        big_burst(st);
    }
    if (t_gpu) T_CUDA(cudaStreamDestroy(st));
    return t_finish("many_outstanding");
}
