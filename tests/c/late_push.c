/* Large messages (>= ACX_LATE_PUSH_MIN): the SENDER moves the payload — from the mover behind its trigger when the receive
 * is already posted, otherwise from the helper CTAs of its wait kernel — and the receiver pulls only as a fallback, when the
 * sender's wait is not running. Covered here, payload-checked, on every ring link:
 *   1. Isend / Irecv / Waitall with the receive posted first and with the send posted first (both movers);
 *   2. Wait_enqueue(recv) BEFORE Wait_enqueue(send) with every rank sending first: each sender's wait sits behind its
 *      receive's wait, so only the receiver-side fallback pull can make progress (MPI allows this order);
 *   3. host MPIX_Wait on the send while the receive is posted late from another stream;
 *   4. a captured graph with a late-size message, relaunched.
 * usage: late_push [bytes]   (default 8 MiB; the emulation matrix passes a small size with ACX_LATE_PUSH_MIN lowered) */
#include "acx_test.h"
#include <unistd.h>

static size_t N;
static unsigned char *hbuf;

static void fill(unsigned char *d, int salt, cudaStream_t st) {
    for (size_t i = 0; i < N; i++) hbuf[i] = (unsigned char)(i * 13 + salt);
    T_CUDA(cudaMemcpyAsync(d, hbuf, N, cudaMemcpyHostToDevice, st));
    T_CUDA(cudaStreamSynchronize(st));
}
static size_t check(const unsigned char *d, int salt) {
    T_CUDA(cudaMemcpy(hbuf, d, N, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < N; i++) if (hbuf[i] != (unsigned char)(i * 13 + salt)) bad++;
    return bad;
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 1);
    N = argc > 1 ? (size_t)atoll(argv[1]) : (8u << 20);
    hbuf = (unsigned char *)malloc(N);
    unsigned char *s, *r;
    T_OK(MPIX_Alloc_mem(N, MPI_INFO_NULL, &s));
    T_OK(MPIX_Alloc_mem(N, MPI_INFO_NULL, &r));
    cudaStream_t st, st2;
    T_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    T_CUDA(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
    MPIX_Request q[2];

    /* 1. both enqueue orders, several rounds (slot reuse, both parities of who arrives first) */
    for (int round = 0; round < 6; round++) {
        fill(s, 10 * round + t_rank, st);
        T_CUDA(cudaMemsetAsync(r, 0, N, st));
        if (round & 1) {
            T_OK(MPIX_Isend_enqueue(s, (int)N, MPI_BYTE, T_RIGHT, 5, MPI_COMM_WORLD, &q[1], MPIX_QUEUE_CUDA_STREAM, &st));
            if (round == 3) usleep(2000);   /* the send certainly arrives first */
            T_OK(MPIX_Irecv_enqueue(r, (int)N, MPI_BYTE, T_LEFT, 5, MPI_COMM_WORLD, &q[0], MPIX_QUEUE_CUDA_STREAM, &st));
        } else {
            T_OK(MPIX_Irecv_enqueue(r, (int)N, MPI_BYTE, T_LEFT, 5, MPI_COMM_WORLD, &q[0], MPIX_QUEUE_CUDA_STREAM, &st));
            T_OK(MPIX_Isend_enqueue(s, (int)N, MPI_BYTE, T_RIGHT, 5, MPI_COMM_WORLD, &q[1], MPIX_QUEUE_CUDA_STREAM, &st));
        }
        T_OK(MPIX_Waitall_enqueue(2, q, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        T_CUDA(cudaStreamSynchronize(st));
        size_t bad = check(r, 10 * round + T_LEFT);
        T_CHECK(bad == 0, "waitall round %d: %zu wrong bytes", round, bad);
    }

    /* 2. every rank sends first and waits for its RECEIVE first: the senders' waits are all stuck behind it */
    for (int round = 0; round < 3; round++) {
        fill(s, 77 + round + t_rank, st);
        T_CUDA(cudaMemsetAsync(r, 0, N, st));
        MPI_Status stt;
        T_OK(MPIX_Isend_enqueue(s, (int)N, MPI_BYTE, T_RIGHT, 6, MPI_COMM_WORLD, &q[1], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Irecv_enqueue(r, (int)N, MPI_BYTE, T_LEFT, 6, MPI_COMM_WORLD, &q[0], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Wait_enqueue(&q[0], &stt, MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Wait_enqueue(&q[1], MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        T_CUDA(cudaStreamSynchronize(st));
        size_t bad = check(r, 77 + round + T_LEFT);
        int cnt = -1;
        MPI_Get_count(&stt, MPI_BYTE, &cnt);
        T_CHECK(bad == 0 && cnt == (int)N && stt.MPI_SOURCE == T_LEFT, "recv-wait-first round %d: %zu wrong bytes, count %d", round, bad, cnt);
    }

    /* 3. host wait on the send; the receive comes late, from another stream */
    fill(s, 140 + t_rank, st);
    T_CUDA(cudaMemsetAsync(r, 0, N, st2));
    T_CUDA(cudaStreamSynchronize(st2));
    T_OK(MPIX_Isend_enqueue(s, (int)N, MPI_BYTE, T_RIGHT, 7, MPI_COMM_WORLD, &q[1], MPIX_QUEUE_CUDA_STREAM, &st));
    usleep(1000);
    T_OK(MPIX_Irecv_enqueue(r, (int)N, MPI_BYTE, T_LEFT, 7, MPI_COMM_WORLD, &q[0], MPIX_QUEUE_CUDA_STREAM, &st2));
    T_OK(MPIX_Wait(&q[1], MPI_STATUS_IGNORE));
    T_OK(MPIX_Wait(&q[0], MPI_STATUS_IGNORE));
    { size_t bad = check(r, 140 + T_LEFT); T_CHECK(bad == 0, "host waits: %zu wrong bytes", bad); }

    /* 4. captured graph, relaunched: the helper CTAs' decision word must be per launch */
    {
        cudaGraph_t g; cudaGraphExec_t ge;
        T_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
        T_OK(MPIX_Isend_enqueue(s, (int)N, MPI_BYTE, T_RIGHT, 8, MPI_COMM_WORLD, &q[1], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Irecv_enqueue(r, (int)N, MPI_BYTE, T_LEFT, 8, MPI_COMM_WORLD, &q[0], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Waitall_enqueue(2, q, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        T_CUDA(cudaStreamEndCapture(st, &g));
        T_CUDA(cudaGraphInstantiate(&ge, g, 0));
        for (int launch = 0; launch < 4; launch++) {
            fill(s, 200 + launch + t_rank, st2);
            T_CUDA(cudaMemsetAsync(r, 0, N, st2));
            T_CUDA(cudaStreamSynchronize(st2));
            MPI_Barrier(MPI_COMM_WORLD);   /* nobody overwrites a source a neighbour may still be reading */
            T_CUDA(cudaGraphLaunch(ge, st));
            T_CUDA(cudaStreamSynchronize(st));
            size_t bad = check(r, 200 + launch + T_LEFT);
            T_CHECK(bad == 0, "graph launch %d: %zu wrong bytes", launch, bad);
            MPI_Barrier(MPI_COMM_WORLD);
        }
        T_CUDA(cudaGraphExecDestroy(ge)); T_CUDA(cudaGraphDestroy(g));
    }
    MPIX_Free_mem(s); MPIX_Free_mem(r);
    cudaStreamDestroy(st); cudaStreamDestroy(st2);
    free(hbuf);
    return t_finish("late_push");
}
