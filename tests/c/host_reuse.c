/* Pageable host buffers are pinned in place and the registration is cached. A buffer that is unmapped and whose addresses
 * come back from a later mmap must NOT be served from the stale registration (the GPU would keep reading / writing the old,
 * driver-pinned physical pages): the library re-validates cached registrations against the page frames behind them
 * (acx_host_regs_validate). Also: 1-KiB rows of one static array as separate message buffers — rows straddle pages that
 * are pinned one by one (the partial-coverage case behind an intermittent illegal memory access on hardware). */
#include "acx_test.h"
#include <sys/mman.h>
#include <unistd.h>

#define LEN (1u << 20)

static int exchange(unsigned char *sbuf, unsigned char *rbuf, int tag, cudaStream_t st) {
    MPIX_Request q[2];
    T_OK(MPIX_Irecv_enqueue(rbuf, (int)LEN, MPI_BYTE, T_LEFT, tag, MPI_COMM_WORLD, &q[0], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Isend_enqueue(sbuf, (int)LEN, MPI_BYTE, T_RIGHT, tag, MPI_COMM_WORLD, &q[1], MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Waitall_enqueue(2, q, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
    if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    cudaStream_t st = 0;
    if (t_gpu) T_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    unsigned char *rbuf = (unsigned char *)malloc(LEN);
    for (int round = 0; round < 3; round++) {
        /* same address every round (the first mapping picks it, the later ones ask for it back), fresh pages every round */
        static void *want = NULL;
        unsigned char *sbuf = (unsigned char *)mmap(want, LEN, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | (want ? MAP_FIXED_NOREPLACE : 0), -1, 0);
        if (sbuf == MAP_FAILED) { sbuf = (unsigned char *)mmap(NULL, LEN, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0); }
        T_CHECK(sbuf != MAP_FAILED, "mmap failed");
        if (!want) want = sbuf;
        for (size_t i = 0; i < LEN; i++) sbuf[i] = (unsigned char)(i * 7 + 31 * round + t_rank);
        memset(rbuf, 0xee, LEN);
        exchange(sbuf, rbuf, 3, st);
        size_t bad = 0;
        for (size_t i = 0; i < LEN; i++) if (rbuf[i] != (unsigned char)(i * 7 + 31 * round + T_LEFT)) bad++;
        T_CHECK(bad == 0, "round %d (buffer %s at %p): %zu wrong bytes — a stale registration of the unmapped buffer was used", round, sbuf == want ? "back" : "elsewhere", (void *)sbuf, bad);
        munmap(sbuf, LEN);
    }
    /* rows of a static array, each its own message buffer, all in flight at once */
    {
        enum { K = 24, ROW = 257 };
        static int sendv[K][ROW], recvv[K][ROW];
        MPIX_Request rq[2 * K];
        for (int k = 0; k < K; k++) for (int i = 0; i < ROW; i++) { sendv[k][i] = t_rank * 100000 + k * 1000 + i; recvv[k][i] = -1; }
        for (int rep = 0; rep < 4; rep++) {
            for (int k = 0; k < K; k++) T_OK(MPIX_Irecv_enqueue(recvv[k], ROW, MPI_INT, T_LEFT, 20 + (k % 3), MPI_COMM_WORLD, &rq[k], MPIX_QUEUE_CUDA_STREAM, &st));
            for (int k = 0; k < K; k++) T_OK(MPIX_Isend_enqueue(sendv[k], ROW, MPI_INT, T_RIGHT, 20 + (k % 3), MPI_COMM_WORLD, &rq[K + k], MPIX_QUEUE_CUDA_STREAM, &st));
            T_OK(MPIX_Waitall_enqueue(2 * K, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
            if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
            int bad = 0;
            for (int k = 0; k < K; k++) for (int i = 0; i < ROW; i++) if (recvv[k][i] != T_LEFT * 100000 + k * 1000 + i) bad++;
            T_CHECK(bad == 0, "static rows, repetition %d: %d wrong elements", rep, bad);
            for (int k = 0; k < K; k++) for (int i = 0; i < ROW; i++) recvv[k][i] = -1;
        }
    }
    free(rbuf);
    if (t_gpu) T_CUDA(cudaStreamDestroy(st));
    return t_finish("host_reuse");
}
