/* Copy-engine rendezvous (ACX_CE_MIN): where the sender's host can see that the matching receive has been enqueued — both
 * hosts count tickets, the receiver leaves {ticket, destination, capacity} in shared memory — the payload travels by
 * cudaMemcpyAsync behind a gate kernel instead of by an SM mover. The path is the sender's choice per message and must be
 * invisible apart from MPIX_Acx_stats::copy_engine_msgs. Payload-checked on every ring link:
 *   1. receive posted first, barrier, send: the record is there — the copy engines MUST have been used (when enabled);
 *   2. a ring of steps in the reference's order (irecv, isend, waitall) with no synchronization between hosts — either path;
 *   3. send enqueued before the receive exists: kernel path, no hang, tickets stay aligned for the next message;
 *   4. small and large messages interleaved on ONE channel (one tag): host tickets and device tickets agree;
 *   5. a receive larger than the message, a cudaMalloc'd (registered or bounced) destination, a status;
 *   6. a message larger than the receive (truncation): the record says so, the kernel path reports MPI_ERR_TRUNCATE;
 *   7. a second stream on the channel: the host stops predicting tickets, later messages take the kernel path and arrive.
 * usage: ce_path [bytes]   (default 8 MiB — below the default ACX_CE_MIN of 64 MiB, so the GPU suite sets ACX_CE_MIN; the emulation matrix passes a small size) */
#include "acx_test.h"
#include <unistd.h>

static size_t N;
static unsigned char *hbuf;

static void fill(unsigned char *d, size_t n, int salt, cudaStream_t st) {
    for (size_t i = 0; i < n; i++) hbuf[i] = (unsigned char)(i * 7 + salt);
    T_CUDA(cudaMemcpyAsync(d, hbuf, n, cudaMemcpyHostToDevice, st));
    T_CUDA(cudaStreamSynchronize(st));
}
static size_t check(const unsigned char *d, size_t n, int salt) {
    T_CUDA(cudaMemcpy(hbuf, d, n, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < n; i++) if (hbuf[i] != (unsigned char)(i * 7 + salt)) bad++;
    return bad;
}
static int verbose;
#define SECTION(n) do { if (verbose && t_rank == 0) printf("ce_path: section %d\n", n); } while (0)
static unsigned long long ce_count(void) { MPIX_Acx_stats s; MPIX_Get_stats(&s); return s.copy_engine_msgs; }

int main(int argc, char **argv) {
    t_init(&argc, &argv, 1);
    verbose = getenv("CE_PATH_VERBOSE") != NULL;
    N = argc > 1 ? (size_t)atoll(argv[1]) : (8u << 20);
    const char *cm = getenv("ACX_CE_MIN");
    const long ce_min = cm ? atol(cm) : (64 << 20);
    const int expect_ce = ce_min > 0 && (long)N >= ce_min && !getenv("ACX_SENDER_PUSH") && !getenv("ACX_COOP") && !getenv("ACX_PEER_BRANCH") &&
                          !(getenv("ACX_SIDE_STREAMS") && atoi(getenv("ACX_SIDE_STREAMS")) == 0);
    hbuf = (unsigned char *)malloc(2 * N);
    unsigned char *s, *r, *priv;
    T_OK(MPIX_Alloc_mem(N, MPI_INFO_NULL, &s));
    T_OK(MPIX_Alloc_mem(2 * N, MPI_INFO_NULL, &r));
    T_CUDA(cudaMalloc((void **)&priv, N));
    cudaStream_t st, st2;
    T_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    T_CUDA(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
    MPIX_Request q[4];

    SECTION(1);
    /* 1. the receive is certainly enqueued before the send */
    for (int round = 0; round < 3; round++) {
        const unsigned long long before = ce_count();
        fill(s, N, 3 * round + t_rank, st);
        T_CUDA(cudaMemsetAsync(r, 0, N, st));
        T_OK(MPIX_Irecv_enqueue(r, (int)N, MPI_BYTE, T_LEFT, 5, MPI_COMM_WORLD, &q[0], MPIX_QUEUE_CUDA_STREAM, &st));
        MPI_Barrier(MPI_COMM_WORLD);
        T_OK(MPIX_Isend_enqueue(s, (int)N, MPI_BYTE, T_RIGHT, 5, MPI_COMM_WORLD, &q[1], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Waitall_enqueue(2, q, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        T_CUDA(cudaStreamSynchronize(st));
        size_t bad = check(r, N, 3 * round + T_LEFT);
        T_CHECK(bad == 0, "posted-first round %d: %zu wrong bytes", round, bad);
        if (expect_ce) T_CHECK(ce_count() == before + 1, "posted-first round %d: the copy engines were not used (%llu -> %llu)", round, before, ce_count());
        else T_CHECK(ce_count() == before, "copy-engine path taken although it is off");
    }

    SECTION(2);
    /* 2. free-running ring steps, several in flight on the stream */
    {
        fill(s, N, 50 + t_rank, st);
        for (int step = 0; step < 12; step++) {
            T_OK(MPIX_Irecv_enqueue(r + (step & 1) * N, (int)N, MPI_BYTE, T_LEFT, 6, MPI_COMM_WORLD, &q[0], MPIX_QUEUE_CUDA_STREAM, &st));
            T_OK(MPIX_Isend_enqueue(s, (int)N, MPI_BYTE, T_RIGHT, 6, MPI_COMM_WORLD, &q[1], MPIX_QUEUE_CUDA_STREAM, &st));
            T_OK(MPIX_Waitall_enqueue(2, q, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        }
        T_CUDA(cudaStreamSynchronize(st));
        size_t bad = check(r, N, 50 + T_LEFT) + check(r + N, N, 50 + T_LEFT);
        T_CHECK(bad == 0, "free-running ring: %zu wrong bytes", bad);
    }

    SECTION(3);
    /* 3. the send goes out before the receive exists anywhere */
    MPI_Barrier(MPI_COMM_WORLD);
    for (int round = 0; round < 2; round++) {
        fill(s, N, 90 + round + t_rank, st);
        T_CUDA(cudaMemsetAsync(r, 0, N, st));
        T_OK(MPIX_Isend_enqueue(s, (int)N, MPI_BYTE, T_RIGHT, 7, MPI_COMM_WORLD, &q[1], MPIX_QUEUE_CUDA_STREAM, &st));
        MPI_Barrier(MPI_COMM_WORLD);
        T_OK(MPIX_Irecv_enqueue(r, (int)N, MPI_BYTE, T_LEFT, 7, MPI_COMM_WORLD, &q[0], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Waitall_enqueue(2, q, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        T_CUDA(cudaStreamSynchronize(st));
        size_t bad = check(r, N, 90 + round + T_LEFT);
        T_CHECK(bad == 0, "send-first round %d: %zu wrong bytes", round, bad);
    }

    SECTION(4);
    /* 4. one tag, sizes alternating between eager, small rendezvous and copy-engine size: tickets must stay aligned */
    {
        const size_t sizes[6] = {64, N, 100000 < N ? 100000 : N / 2, N, 8, N};
        for (int i = 0; i < 6; i++) {
            fill(s, sizes[i], 120 + i + t_rank, st);
            T_CUDA(cudaMemsetAsync(r, 0, N, st));
            T_OK(MPIX_Irecv_enqueue(r, (int)N, MPI_BYTE, T_LEFT, 8, MPI_COMM_WORLD, &q[0], MPIX_QUEUE_CUDA_STREAM, &st));
            if (i & 1) MPI_Barrier(MPI_COMM_WORLD);
            T_OK(MPIX_Isend_enqueue(s, (int)sizes[i], MPI_BYTE, T_RIGHT, 8, MPI_COMM_WORLD, &q[1], MPIX_QUEUE_CUDA_STREAM, &st));
            MPI_Status sts[2];
            T_OK(MPIX_Waitall_enqueue(2, q, sts, MPIX_QUEUE_CUDA_STREAM, &st));
            T_CUDA(cudaStreamSynchronize(st));
            int cnt = -1;
            MPI_Get_count(&sts[0], MPI_BYTE, &cnt);
            size_t bad = check(r, sizes[i], 120 + i + T_LEFT);
            T_CHECK(bad == 0 && cnt == (int)sizes[i], "mixed sizes, message %d (%zu bytes): %zu wrong bytes, count %d", i, sizes[i], bad, cnt);
        }
    }

    SECTION(5);
    /* 5. receive twice the size of the message into a cudaMalloc'd buffer, status checked */
    for (int round = 0; round < 3; round++) {
        fill(s, N, 160 + round + t_rank, st);
        T_CUDA(cudaMemsetAsync(priv, 0, N, st));
        MPI_Status stt;
        T_OK(MPIX_Irecv_enqueue(priv, (int)N, MPI_BYTE, T_LEFT, 9, MPI_COMM_WORLD, &q[0], MPIX_QUEUE_CUDA_STREAM, &st));
        MPI_Barrier(MPI_COMM_WORLD);
        T_OK(MPIX_Isend_enqueue(s, (int)(N / 2), MPI_BYTE, T_RIGHT, 9, MPI_COMM_WORLD, &q[1], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Wait_enqueue(&q[1], MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Wait_enqueue(&q[0], &stt, MPIX_QUEUE_CUDA_STREAM, &st));
        T_CUDA(cudaStreamSynchronize(st));
        int cnt = -1;
        MPI_Get_count(&stt, MPI_BYTE, &cnt);
        size_t bad = check(priv, N / 2, 160 + round + T_LEFT);
        T_CHECK(bad == 0 && cnt == (int)(N / 2) && stt.MPI_SOURCE == T_LEFT && stt.MPI_TAG == 9, "private destination round %d: %zu wrong bytes, count %d", round, bad, cnt);
    }

    SECTION(6);
    /* 6. truncation: the receive holds half the message */
    {
        fill(s, N, 200 + t_rank, st);
        MPI_Status stt;
        memset(&stt, 0, sizeof stt);
        T_OK(MPIX_Irecv_enqueue(r, (int)(N / 2), MPI_BYTE, T_LEFT, 10, MPI_COMM_WORLD, &q[0], MPIX_QUEUE_CUDA_STREAM, &st));
        MPI_Barrier(MPI_COMM_WORLD);
        T_OK(MPIX_Isend_enqueue(s, (int)N, MPI_BYTE, T_RIGHT, 10, MPI_COMM_WORLD, &q[1], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Wait_enqueue(&q[1], MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Wait_enqueue(&q[0], &stt, MPIX_QUEUE_CUDA_STREAM, &st));
        T_CUDA(cudaStreamSynchronize(st));
        T_CHECK(stt.MPI_ERROR == MPI_ERR_TRUNCATE, "truncated message: status error %d, expected MPI_ERR_TRUNCATE", stt.MPI_ERROR);
    }

    SECTION(7);
    /* 7. the channel is used from a second stream: tickets are drawn in device order from here on */
    {
        fill(s, N, 230 + t_rank, st);
        T_CUDA(cudaMemsetAsync(r, 0, N, st2));
        T_CUDA(cudaStreamSynchronize(st2));
        const unsigned long long before = ce_count();
        for (int round = 0; round < 2; round++) {
            T_OK(MPIX_Irecv_enqueue(r, (int)N, MPI_BYTE, T_LEFT, 5, MPI_COMM_WORLD, &q[0], MPIX_QUEUE_CUDA_STREAM, &st2));
            T_OK(MPIX_Isend_enqueue(s, (int)N, MPI_BYTE, T_RIGHT, 5, MPI_COMM_WORLD, &q[1], MPIX_QUEUE_CUDA_STREAM, &st2));
            T_OK(MPIX_Waitall_enqueue(2, q, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st2));
            T_CUDA(cudaStreamSynchronize(st2));
            MPI_Barrier(MPI_COMM_WORLD);
        }
        size_t bad = check(r, N, 230 + T_LEFT);
        T_CHECK(bad == 0, "second stream: %zu wrong bytes", bad);
        T_CHECK(ce_count() == before, "copy-engine path on a channel whose tickets the host can no longer predict");
    }
    if (t_rank == 0) printf("ce_path: %llu of this rank's messages went by copy engine\n", ce_count());
    MPIX_Free_mem(s); MPIX_Free_mem(r); cudaFree(priv);
    cudaStreamDestroy(st); cudaStreamDestroy(st2);
    free(hbuf);
    return t_finish("ce_path");
}
