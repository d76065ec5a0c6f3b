/* Several application threads per rank drive the library at once (the reference demands MPI_THREAD_MULTIPLE but
 * documents its own slot allocator as not thread-safe, src/triggered.cpp:40-43). Each thread owns a stream and a tag
 * (= a channel), runs ring exchanges of eager and rendezvous sizes with stream waits and host waits, and churns the
 * symmetric heap (MPIX_Alloc_mem / MPIX_Free_mem) in between; a last thread runs partitioned rounds with host Pready.
 * Under the software device this program is the ThreadSanitizer target for the host side of the runtime. */
#include "acx_test.h"
#include <pthread.h>

#define NTHREADS 4
#define ITERS 60

static int g_thread_errs[NTHREADS + 1];
static int g_dev = -1;   /* the CUDA device of this rank: every application thread must make it current (CUDA's current device is per thread) */

static unsigned char pat(size_t i, int salt) { return (unsigned char)((i * 37u + (size_t)salt * 11u) & 0xff); }

static void *p2p_thread(void *arg) {
    const int id = (int)(long)arg;
    const int tag = 3 + id;                       /* distinct channels: tickets of different threads never interleave */
    const size_t sizes[3] = {48, 6000, 70000};    /* eager, eager (multi-thread copy), rendezvous */
    cudaStream_t st;
    if (t_gpu) cudaSetDevice(g_dev);
    if (t_gpu) cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking); else st = 0;
    int errs = 0;
    for (int it = 0; it < ITERS && !errs; it++) {
        const size_t n = sizes[(it + id) % 3];
        unsigned char *out, *in, *h = (unsigned char *)malloc(n);
        if (MPIX_Alloc_mem(n, MPI_INFO_NULL, &out) || MPIX_Alloc_mem(n, MPI_INFO_NULL, &in)) { errs++; free(h); break; }
        for (size_t i = 0; i < n; i++) h[i] = pat(i, t_rank * 100 + id * 10 + it);
        if (t_gpu) { cudaMemcpyAsync(out, h, n, cudaMemcpyHostToDevice, st); cudaMemsetAsync(in, 0, n, st); } else { memcpy(out, h, n); memset(in, 0, n); }
        MPIX_Request rq[2];
        MPI_Status sts[2];
        if (MPIX_Irecv_enqueue(in, (int)n, MPI_BYTE, T_LEFT, tag, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &st)) errs++;
        if (MPIX_Isend_enqueue(out, (int)n, MPI_BYTE, T_RIGHT, tag, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &st)) errs++;
        if (it & 1) { if (MPIX_Waitall_enqueue(2, rq, sts, MPIX_QUEUE_CUDA_STREAM, &st)) errs++; }
        else { if (MPIX_Wait(&rq[0], &sts[0]) || MPIX_Wait(&rq[1], &sts[1])) errs++; }
        if (t_gpu) { cudaMemcpyAsync(h, in, n, cudaMemcpyDeviceToHost, st); if (cudaStreamSynchronize(st) != cudaSuccess) errs++; } else memcpy(h, in, n);
        size_t bad = 0;
        for (size_t i = 0; i < n; i++) if (h[i] != pat(i, T_LEFT * 100 + id * 10 + it)) bad++;
        int cnt = -1;
        MPI_Get_count(&sts[0], MPI_BYTE, &cnt);
        if (bad || cnt != (int)n || sts[0].MPI_SOURCE != T_LEFT || sts[0].MPI_TAG != tag) {
            printf("[%d] thread %d iter %d n=%zu: %zu bad bytes, count %d source %d tag %d\n", t_rank, id, it, n, bad, cnt, sts[0].MPI_SOURCE, sts[0].MPI_TAG);
            errs++;
        }
        MPIX_Free_mem(out); MPIX_Free_mem(in); free(h);
    }
    if (t_gpu) cudaStreamDestroy(st);
    g_thread_errs[id] = errs;
    return NULL;
}

static void *partitioned_thread(void *arg) {
    (void)arg;
    enum { NP = 6, PER = 5 };
    static int sendbuf[NP * PER], recvbuf[NP * PER];
    MPIX_Request rq[2];
    int errs = 0;
    if (t_gpu) cudaSetDevice(g_dev);
    if (MPIX_Psend_init(sendbuf, NP, PER, MPI_INT, T_RIGHT, 90, MPI_COMM_WORLD, MPI_INFO_NULL, &rq[0])) errs++;
    if (MPIX_Precv_init(recvbuf, NP, PER, MPI_INT, T_LEFT, 90, MPI_COMM_WORLD, MPI_INFO_NULL, &rq[1])) errs++;
    for (int it = 0; it < 12 && !errs; it++) {
        for (int i = 0; i < NP * PER; i++) { sendbuf[i] = t_rank * 10000 + it * 100 + i; recvbuf[i] = -1; }
        if (MPIX_Startall(2, rq)) errs++;
        for (int p = 0; p < NP; p++) if (MPIX_Pready_host(p, &rq[0])) errs++;
        if (MPIX_Waitall(2, rq, MPI_STATUSES_IGNORE)) errs++;
        for (int i = 0; i < NP * PER; i++) if (recvbuf[i] != T_LEFT * 10000 + it * 100 + i) { errs++; break; }
    }
    if (MPIX_Request_free(&rq[0]) || MPIX_Request_free(&rq[1])) errs++;
    g_thread_errs[NTHREADS] = errs;
    return NULL;
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    if (t_gpu) T_CUDA(cudaGetDevice(&g_dev));
    pthread_t th[NTHREADS + 1];
    for (long i = 0; i < NTHREADS; i++) pthread_create(&th[i], NULL, p2p_thread, (void *)i);
    pthread_create(&th[NTHREADS], NULL, partitioned_thread, NULL);
    for (int i = 0; i <= NTHREADS; i++) pthread_join(th[i], NULL);
    for (int i = 0; i <= NTHREADS; i++) T_CHECK(g_thread_errs[i] == 0, "thread %d reported %d error(s)", i, g_thread_errs[i]);
    MPIX_Acx_stats s;
    MPIX_Get_stats(&s);
    T_CHECK(!t_gpu || s.heap_bytes_in_use <= (NTHREADS + 1) * 2 * 1024 * 1024, "heap not returned: %llu bytes in use", s.heap_bytes_in_use);
    return t_finish("threads");
}
