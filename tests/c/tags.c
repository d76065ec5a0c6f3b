/* Exact (source, tag, communicator) matching on the device transport — what the reference gets from the MPI library
 * underneath (it passes tag and peer through untouched, reference src/init.cpp:69-70,76-77) and what round 1 of this
 * library only approximated with `tag mod 16` FIFOs:
 *   1. 24 tags to one peer, receives posted in increasing tag order, sends issued in DECREASING order, eager and
 *      rendezvous sizes mixed, among them tags that collide modulo 16 and modulo 32: every receive gets its own tag's payload;
 *   2. the same tag on two communicators, sent in one order and received in the other;
 *   3. more than ACX_DEPTH messages per tag still flow (slot recycling per channel);
 *   4. with ACX_WILDCARDS=1: MPI_ANY_TAG picks up a message whatever its tag, MPI_ANY_SOURCE whatever its source, and
 *      specific receives keep matching next to them.
 * Runs on the host transport too (there the MPI library matches). */
#include "acx_test.h"

#define NT 24
#define BIG 30000   /* ints: 120 KB, rendezvous */

static int tag_of(int k) { static const int base[NT] = {3, 19, 35, 51, 67, 4, 20, 36, 7, 1000, 1016, 1032, 5, 37, 69, 2, 18, 34, 66, 98, 130, 162, 9, 41}; return base[k]; }
static int len_of(int k) { return (k % 3 == 0) ? BIG : (k % 3 == 1 ? 1 : 257); }

static void many_tags(cudaStream_t st) {
    int *sbuf[NT], *rbuf[NT];
    MPIX_Request rq[2 * NT];
    MPI_Status sts[2 * NT];
    for (int k = 0; k < NT; k++) {
        sbuf[k] = (int *)malloc(sizeof(int) * BIG); rbuf[k] = (int *)malloc(sizeof(int) * BIG);
        for (int i = 0; i < BIG; i++) { sbuf[k][i] = t_rank * 1000000 + tag_of(k) * 100 + (i % 97); rbuf[k][i] = -1; }
    }
    for (int k = 0; k < NT; k++) T_OK(MPIX_Irecv_enqueue(rbuf[k], BIG, MPI_INT, T_LEFT, tag_of(k), MPI_COMM_WORLD, &rq[k], MPIX_QUEUE_CUDA_STREAM, &st));
    for (int k = NT - 1; k >= 0; k--) T_OK(MPIX_Isend_enqueue(sbuf[k], len_of(k), MPI_INT, T_RIGHT, tag_of(k), MPI_COMM_WORLD, &rq[NT + k], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Waitall_enqueue(2 * NT, rq, sts, MPIX_QUEUE_CUDA_STREAM, &st));
    if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
    int bad = 0;
    for (int k = 0; k < NT; k++) {
        int cnt = -1;
        MPI_Get_count(&sts[k], MPI_INT, &cnt);
        if (cnt != len_of(k) || sts[k].MPI_TAG != tag_of(k) || sts[k].MPI_SOURCE != T_LEFT || sts[k].MPI_ERROR != MPI_SUCCESS) { bad++; printf("[%d] tag %d: count %d (want %d) status tag %d src %d err %d\n", t_rank, tag_of(k), cnt, len_of(k), sts[k].MPI_TAG, sts[k].MPI_SOURCE, sts[k].MPI_ERROR); }
        for (int i = 0; i < len_of(k); i++) if (rbuf[k][i] != T_LEFT * 1000000 + tag_of(k) * 100 + (i % 97)) { bad++; printf("[%d] tag %d: payload[%d] = %d\n", t_rank, tag_of(k), i, rbuf[k][i]); break; }
        if (len_of(k) < BIG && rbuf[k][len_of(k)] != -1) bad++;
    }
    T_CHECK(bad == 0, "%d tags mis-matched with sends in reverse order", bad);
    for (int k = 0; k < NT; k++) { free(sbuf[k]); free(rbuf[k]); }
}

static void same_tag_two_communicators(cudaStream_t st) {
    MPI_Comm dup;
    MPI_Comm_dup(MPI_COMM_WORLD, &dup);
    int sa = 111 + t_rank, sb = 222 + t_rank, ra = -1, rb = -1;
    MPIX_Request rq[4];
    T_OK(MPIX_Irecv_enqueue(&ra, 1, MPI_INT, T_LEFT, 12, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Irecv_enqueue(&rb, 1, MPI_INT, T_LEFT, 12, dup, &rq[1], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Isend_enqueue(&sb, 1, MPI_INT, T_RIGHT, 12, dup, &rq[2], MPIX_QUEUE_CUDA_STREAM, &st));             /* the duplicate's message goes first */
    T_OK(MPIX_Isend_enqueue(&sa, 1, MPI_INT, T_RIGHT, 12, MPI_COMM_WORLD, &rq[3], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Waitall_enqueue(4, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
    if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
    T_CHECK(ra == 111 + T_LEFT && rb == 222 + T_LEFT, "same tag on two communicators: world got %d, dup got %d", ra, rb);
    MPI_Comm_free(&dup);
}

static void deep_channel(cudaStream_t st) {
    /* 100 messages on one tag, 10 in flight at a time: tickets wrap around the channel's ring several times */
    int bad = 0;
    for (int round = 0; round < 10; round++) {
        int s[10], r[10];
        MPIX_Request rq[20];
        for (int i = 0; i < 10; i++) { s[i] = round * 10 + i + 1000 * t_rank; r[i] = -1; }
        for (int i = 0; i < 10; i++) T_OK(MPIX_Irecv_enqueue(&r[i], 1, MPI_INT, T_LEFT, 77, MPI_COMM_WORLD, &rq[i], MPIX_QUEUE_CUDA_STREAM, &st));
        for (int i = 0; i < 10; i++) T_OK(MPIX_Isend_enqueue(&s[i], 1, MPI_INT, T_RIGHT, 77, MPI_COMM_WORLD, &rq[10 + i], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Waitall_enqueue(20, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < 10; i++) if (r[i] != round * 10 + i + 1000 * T_LEFT) bad++;
    }
    T_CHECK(bad == 0, "deep channel: %d of 100 messages out of order", bad);
}

static void wildcards(cudaStream_t st) {
    int s1 = 500 + t_rank, s2 = 600 + t_rank, s3 = 700 + t_rank, r1 = -1, r2 = -1, r3 = -1;
    MPIX_Request rq[6];
    MPI_Status sts[6];
    T_OK(MPIX_Isend_enqueue(&s1, 1, MPI_INT, T_RIGHT, 301, MPI_COMM_WORLD, &rq[3], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Isend_enqueue(&s2, 1, MPI_INT, T_RIGHT, 302, MPI_COMM_WORLD, &rq[4], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Isend_enqueue(&s3, 1, MPI_INT, T_RIGHT, 303, MPI_COMM_WORLD, &rq[5], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Irecv_enqueue(&r2, 1, MPI_INT, T_LEFT, 302, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &st));            /* specific */
    T_OK(MPIX_Irecv_enqueue(&r1, 1, MPI_INT, T_LEFT, MPI_ANY_TAG, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &st));    /* any tag from the left */
    T_OK(MPIX_Irecv_enqueue(&r3, 1, MPI_INT, MPI_ANY_SOURCE, MPI_ANY_TAG, MPI_COMM_WORLD, &rq[2], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Waitall_enqueue(6, rq, sts, MPIX_QUEUE_CUDA_STREAM, &st));
    if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
    T_CHECK(r2 == 600 + T_LEFT && sts[1].MPI_TAG == 302, "specific receive next to wildcards got %d (tag %d)", r2, sts[1].MPI_TAG);
    /* the two wildcard receives share the messages tagged 301 and 303, one each, and the statuses say which */
    const int ok13 = (r1 == 500 + T_LEFT && sts[0].MPI_TAG == 301 && r3 == 700 + T_LEFT && sts[2].MPI_TAG == 303);
    const int ok31 = (r1 == 700 + T_LEFT && sts[0].MPI_TAG == 303 && r3 == 500 + T_LEFT && sts[2].MPI_TAG == 301);
    T_CHECK(ok13 || ok31, "wildcards: r1=%d (tag %d) r3=%d (tag %d, source %d)", r1, sts[0].MPI_TAG, r3, sts[2].MPI_TAG, sts[2].MPI_SOURCE);
    T_CHECK(sts[0].MPI_SOURCE == T_LEFT && sts[2].MPI_SOURCE == T_LEFT, "wildcard sources %d %d", sts[0].MPI_SOURCE, sts[2].MPI_SOURCE);
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    cudaStream_t st = 0;
    if (t_gpu) T_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    many_tags(st);
    same_tag_two_communicators(st);
    deep_channel(st);
    if (!t_gpu || (getenv("ACX_WILDCARDS") && atoi(getenv("ACX_WILDCARDS")))) wildcards(st);
    if (t_gpu) T_CUDA(cudaStreamDestroy(st));
    return t_finish("tags");
}
