/* What the reference leaves to its MPI library and never tests: truncation, status contents, argument validation,
 * large request arrays.
 *   - a message longer than the receive: only `capacity` bytes land (guard bytes behind the buffer stay intact), the
 *     status says MPI_ERR_TRUNCATE and counts the bytes delivered — eager and rendezvous sizes, stream and host waits;
 *   - device transport only: two tags that share a channel (tag mod ACX_NCHAN) are one FIFO, and a receive that meets
 *     (exact per-tag channels since round 2) is matched by tag, not by arrival order;
 *   - MPIX_Waitall_enqueue over more requests than one wait kernel takes (30 > 12), statuses for every one of them;
 *   - invalid arguments come back as MPI error codes, the request stays MPIX_REQUEST_NULL and nothing leaks;
 *   - MPIX_Request_free on an un-waited basic request completes it (the reference ignores it, defect D11). */
#include "acx_test.h"
#include <time.h>

static unsigned char pat(size_t i, int salt) { return (unsigned char)((i * 13u + (size_t)salt * 7u + 1) & 0xff); }

static void put(void *dst, const void *src, size_t n) { if (t_gpu) { T_CUDA(cudaMemcpy(dst, src, n, cudaMemcpyDefault)); } else memcpy(dst, src, n); }

static void truncation(size_t send_n, size_t cap, int host_wait, int tag) {
    unsigned char *out, *in, *h = (unsigned char *)malloc(send_n + 64);
    cudaStream_t st = 0;
    T_OK(MPIX_Alloc_mem(send_n, MPI_INFO_NULL, &out));
    T_OK(MPIX_Alloc_mem(cap + 64, MPI_INFO_NULL, &in));
    for (size_t i = 0; i < send_n; i++) h[i] = pat(i, t_rank + tag);
    put(out, h, send_n);
    memset(h, 0xcd, cap + 64);
    put(in, h, cap + 64);
    MPIX_Request rq[2];
    MPI_Status sts[2];
    memset(sts, 0, sizeof sts);
    T_OK(MPIX_Irecv_enqueue(in, (int)cap, MPI_BYTE, T_LEFT, tag, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Isend_enqueue(out, (int)send_n, MPI_BYTE, T_RIGHT, tag, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &st));
    if (host_wait) {
        int rc0 = MPIX_Wait(&rq[0], &sts[0]);
        T_CHECK(rc0 == MPI_SUCCESS || rc0 == MPI_ERR_TRUNCATE, "host wait returned %d", rc0);
        T_OK(MPIX_Wait(&rq[1], &sts[1]));
    } else {
        int rc = MPIX_Waitall_enqueue(2, rq, sts, MPIX_QUEUE_CUDA_STREAM, &st);
        T_CHECK(rc == MPI_SUCCESS || rc == MPI_ERR_TRUNCATE, "waitall_enqueue returned %d", rc);
        if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
    }
    put(h, in, cap + 64);
    /* A message that needs the rendezvous path cannot be squeezed into a receive smaller than the eager limit: the device
     * transport completes it with nothing delivered (count 0). Everywhere else the first `cap` bytes arrive. */
    int cnt = -1;
    MPI_Get_count(&sts[0], MPI_BYTE, &cnt);
    const size_t delivered = (t_gpu && cnt == 0 && cap <= 8192 && send_n > 8192) ? 0 : cap;
    size_t bad = 0, overrun = 0;
    for (size_t i = 0; i < delivered; i++) if (h[i] != pat(i, T_LEFT + tag)) bad++;
    for (size_t i = delivered; i < cap + 64; i++) if (h[i] != 0xcd) overrun++;
    T_CHECK(bad == 0 && overrun == 0, "send %zu into %zu (host_wait %d): %zu wrong bytes, %zu guard bytes overwritten", send_n, cap, host_wait, bad, overrun);
    T_CHECK(sts[0].MPI_ERROR == MPI_ERR_TRUNCATE && cnt == (int)delivered && sts[0].MPI_SOURCE == T_LEFT && sts[0].MPI_TAG == tag,
            "send %zu into %zu (host_wait %d): status error %d count %d source %d tag %d", send_n, cap, host_wait, sts[0].MPI_ERROR, cnt, sts[0].MPI_SOURCE, sts[0].MPI_TAG);
    T_OK(MPIX_Free_mem(out)); T_OK(MPIX_Free_mem(in));
    free(h);
}

static void many_requests(void) {
    enum { K = 15, LEN = 257 };   /* 30 requests in one Waitall_enqueue */
    static int sendv[K][LEN], recvv[K][LEN];
    cudaStream_t st = 0;
    MPIX_Request rq[2 * K];
    MPI_Status sts[2 * K];
    for (int k = 0; k < K; k++) for (int i = 0; i < LEN; i++) { sendv[k][i] = t_rank * 100000 + k * 1000 + i; recvv[k][i] = -1; }
    for (int k = 0; k < K; k++) T_OK(MPIX_Irecv_enqueue(recvv[k], LEN, MPI_INT, T_LEFT, 50 + (k % 4), MPI_COMM_WORLD, &rq[k], MPIX_QUEUE_CUDA_STREAM, &st));
    for (int k = 0; k < K; k++) T_OK(MPIX_Isend_enqueue(sendv[k], LEN - k, MPI_INT, T_RIGHT, 50 + (k % 4), MPI_COMM_WORLD, &rq[K + k], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Waitall_enqueue(2 * K, rq, sts, MPIX_QUEUE_CUDA_STREAM, &st));
    if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
    int bad = 0;
    for (int k = 0; k < K; k++) {
        int cnt = -1;
        MPI_Get_count(&sts[k], MPI_INT, &cnt);
        if (cnt != LEN - k || sts[k].MPI_SOURCE != T_LEFT || sts[k].MPI_TAG != 50 + (k % 4) || sts[k].MPI_ERROR != MPI_SUCCESS) bad++;
        for (int i = 0; i < LEN - k; i++) if (recvv[k][i] != T_LEFT * 100000 + k * 1000 + i) { bad++; break; }
        if (rq[k] != MPIX_REQUEST_NULL || rq[K + k] != MPIX_REQUEST_NULL) bad++;
    }
    T_CHECK(bad == 0, "30-request waitall: %d problems", bad);
}

static void tag_collision(void) {
    /* tags 2 and 2 + nchan hashed onto one FIFO in round 1 and the receive for tag 2 met the other tag's message
       (MPI_ERR_TAG). Channels are exact per (communicator, tag) now: posted in the "wrong" order, both still match. */
    const int nchan = getenv("ACX_NCHAN") ? atoi(getenv("ACX_NCHAN")) : 16;
    int a = 1234 + t_rank, a2 = 4321 + t_rank, b = -1, b2 = -1;
    cudaStream_t st = 0;
    MPIX_Request rq[4];
    MPI_Status sts[4];
    T_OK(MPIX_Isend_enqueue(&a, 1, MPI_INT, T_RIGHT, 2 + nchan, MPI_COMM_WORLD, &rq[2], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Isend_enqueue(&a2, 1, MPI_INT, T_RIGHT, 2, MPI_COMM_WORLD, &rq[3], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Irecv_enqueue(&b2, 1, MPI_INT, T_LEFT, 2, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Irecv_enqueue(&b, 1, MPI_INT, T_LEFT, 2 + nchan, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Waitall_enqueue(4, rq, sts, MPIX_QUEUE_CUDA_STREAM, &st));
    T_CUDA(cudaStreamSynchronize(st));
    T_CHECK(sts[0].MPI_ERROR == MPI_SUCCESS && sts[0].MPI_TAG == 2 && b2 == 4321 + T_LEFT, "tag 2: error %d tag %d value %d", sts[0].MPI_ERROR, sts[0].MPI_TAG, b2);
    T_CHECK(sts[1].MPI_ERROR == MPI_SUCCESS && sts[1].MPI_TAG == 2 + nchan && b == 1234 + T_LEFT, "tag 2+nchan: error %d tag %d value %d", sts[1].MPI_ERROR, sts[1].MPI_TAG, b);
}

static void bad_arguments(void) {
    int v = 0;
    cudaStream_t st = 0;
    MPIX_Request q = MPIX_REQUEST_NULL;
    MPIX_Acx_stats s0, s1;
    MPIX_Get_stats(&s0);
    T_CHECK(MPIX_Isend_enqueue(&v, -1, MPI_INT, T_RIGHT, 1, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, &st) == MPI_ERR_COUNT && q == MPIX_REQUEST_NULL, "negative count");
    T_CHECK(MPIX_Isend_enqueue(&v, 1, MPI_INT, t_size, 1, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, &st) == MPI_ERR_RANK && q == MPIX_REQUEST_NULL, "rank out of range");
    T_CHECK(MPIX_Irecv_enqueue(&v, 1, MPI_INT, -7, 1, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, &st) == MPI_ERR_RANK && q == MPIX_REQUEST_NULL, "negative rank");
    T_CHECK(MPIX_Isend_enqueue(&v, 1, MPI_INT, T_RIGHT, -5, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, &st) == MPI_ERR_TAG && q == MPIX_REQUEST_NULL, "negative tag");
    T_CHECK(MPIX_Isend_enqueue(&v, 1, MPI_INT, T_RIGHT, 1, MPI_COMM_WORLD, NULL, MPIX_QUEUE_CUDA_STREAM, &st) == MPI_ERR_ARG, "null request pointer");
    T_CHECK(MPIX_Isend_enqueue(&v, 1, MPI_INT, T_RIGHT, 1, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, NULL) == MPI_ERR_ARG && q == MPIX_REQUEST_NULL, "null queue");
    if (t_gpu) T_CHECK(MPIX_Isend_enqueue(&v, 1, MPI_INT, T_RIGHT, 1, MPI_COMM_WORLD, &q, 7, &st) == MPI_ERR_ARG && q == MPIX_REQUEST_NULL, "unknown queue type");
    T_CHECK(MPIX_Psend_init(&v, 0, 1, MPI_INT, T_RIGHT, 1, MPI_COMM_WORLD, MPI_INFO_NULL, &q) == MPI_ERR_COUNT, "zero partitions");
    int junk = 5;
    MPIX_Request bogus = (MPIX_Request)&junk;
    T_CHECK(MPIX_Wait_enqueue(&bogus, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st) == MPI_ERR_REQUEST, "bogus request handle");
    q = MPIX_REQUEST_NULL;
    T_CHECK(MPIX_Wait(&q, MPI_STATUS_IGNORE) == MPI_SUCCESS && MPIX_Wait_enqueue(&q, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st) == MPI_SUCCESS, "waiting on MPIX_REQUEST_NULL is a no-op");
    MPIX_Get_stats(&s1);
    T_CHECK(s1.slots_in_use == s0.slots_in_use && s1.heap_bytes_in_use == s0.heap_bytes_in_use, "rejected calls leaked: slots %llu -> %llu, heap %llu -> %llu", s0.slots_in_use, s1.slots_in_use,
            s0.heap_bytes_in_use, s1.heap_bytes_in_use);
}

static void request_free_completes(void) {
    static int a[3000], b[3000];
    cudaStream_t st = 0;
    MPIX_Request rq[2];
    for (int i = 0; i < 3000; i++) { a[i] = t_rank * 7 + i; b[i] = -1; }
    T_OK(MPIX_Irecv_enqueue(b, 3000, MPI_INT, T_LEFT, 8, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Isend_enqueue(a, 3000, MPI_INT, T_RIGHT, 8, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Request_free(&rq[0]));
    T_OK(MPIX_Request_free(&rq[1]));
    int bad = 0;
    for (int i = 0; i < 3000; i++) if (b[i] != T_LEFT * 7 + i) bad++;
    T_CHECK(bad == 0 && rq[0] == MPIX_REQUEST_NULL && rq[1] == MPIX_REQUEST_NULL, "Request_free on un-waited requests: %d wrong", bad);
}

/* Operations enqueued on one stream, waited on another (the reference allows it: its completion flags are global). The
 * second stream is synchronized alone; the first one is deliberately kept busy behind a marker event so that a missing
 * order would show. */
static void stall_20ms(void *arg) { (void)arg; struct timespec ts = {0, 20 * 1000 * 1000}; nanosleep(&ts, NULL); }
static void cross_stream_wait(void) {
    if (!t_gpu) return;
    cudaStream_t a, b;
    T_CUDA(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    T_CUDA(cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking));
    const size_t sizes[] = {64, 100000};
    for (unsigned s = 0; s < 2; s++) {
        const size_t n = sizes[s];
        unsigned char *out, *in, *h = (unsigned char *)malloc(n), *h2 = NULL;
        T_CUDA(cudaMallocHost((void **)&h2, n));
        T_OK(MPIX_Alloc_mem(n, MPI_INFO_NULL, &out)); T_OK(MPIX_Alloc_mem(n, MPI_INFO_NULL, &in));
        for (size_t i = 0; i < n; i++) h[i] = pat(i, t_rank + 60 + (int)s);
        T_CUDA(cudaLaunchHostFunc(a, stall_20ms, NULL));                  /* stream a lags: stream b would run ahead of the triggers */
        T_CUDA(cudaMemcpyAsync(out, h, n, cudaMemcpyHostToDevice, a));   /* on stream a: the send must see it */
        T_CUDA(cudaMemsetAsync(in, 0, n, a));
        MPIX_Request rq[2];
        MPI_Status sts[2];
        T_OK(MPIX_Irecv_enqueue(in, (int)n, MPI_BYTE, T_LEFT, 44, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &a));
        T_OK(MPIX_Isend_enqueue(out, (int)n, MPI_BYTE, T_RIGHT, 44, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &a));
        T_OK(MPIX_Waitall_enqueue(2, rq, sts, MPIX_QUEUE_CUDA_STREAM, &b));
        T_CUDA(cudaMemcpyAsync(h2, in, n, cudaMemcpyDeviceToHost, b));
        T_CUDA(cudaStreamSynchronize(b));
        size_t bad = 0;
        for (size_t i = 0; i < n; i++) if (h2[i] != pat(i, T_LEFT + 60 + (int)s)) bad++;
        int cnt = -1;
        MPI_Get_count(&sts[0], MPI_BYTE, &cnt);
        T_CHECK(bad == 0 && cnt == (int)n, "wait on another stream, %zu B: %zu wrong bytes, count %d", n, bad, cnt);
        T_CUDA(cudaStreamSynchronize(a));
        T_OK(MPIX_Free_mem(out)); T_OK(MPIX_Free_mem(in));
        cudaFreeHost(h2); free(h);
    }
    T_CUDA(cudaStreamDestroy(a)); T_CUDA(cudaStreamDestroy(b));
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    truncation(400, 160, 0, 30);          /* eager, stream wait */
    truncation(400, 160, 1, 31);          /* eager, host wait */
    truncation(100000, 50000, 0, 32);     /* rendezvous, stream wait */
    truncation(100000, 50000, 1, 33);     /* rendezvous, host wait */
    truncation(100000, 3000, 0, 34);      /* rendezvous-size message into an eager-size receive */
    truncation(100000, 3000, 1, 35);
    many_requests();
    cross_stream_wait();
    if (t_gpu) tag_collision();
    bad_arguments();
    request_free_completes();
    return t_finish("status_errors");
}
