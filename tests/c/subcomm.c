/* Sub-communicators through the MPIX API. The reference forwards any communicator to its MPI library; here peers are
 * translated to MPI_COMM_WORLD ranks (windows and rings are indexed by world rank) and MPI_Status reports the source in
 * the communicator's own rank space.
 *  - the world is split into an even and an odd half whose ranks run in REVERSE world order (key = -rank);
 *  - ring exchange inside each half: eager and rendezvous sizes, stream waits and host waits, status checked;
 *  - a partitioned ring inside the half driven by the host Pready/Parrived entry points;
 *  - a Cartesian communicator over all ranks: shifts with MPI_PROC_NULL at the non-periodic edges. */
#include "acx_test.h"

static void put(void *dst, const void *src, size_t n) { if (t_gpu) T_CUDA(cudaMemcpy(dst, src, n, cudaMemcpyDefault)); else memcpy(dst, src, n); }

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    cudaStream_t stream = 0;
    MPI_Comm half;
    MPI_Comm_split(MPI_COMM_WORLD, t_rank % 2, -t_rank, &half);
    int hr, hs;
    MPI_Comm_rank(half, &hr); MPI_Comm_size(half, &hs);
    const int right = (hr + 1) % hs, left = (hr + hs - 1) % hs;
    /* world rank of my left neighbour in the half (reverse order: comm rank i is the i-th highest world rank of my parity) */
    int top = t_size - 1; if (top % 2 != t_rank % 2) top--;
    const int left_world = top - 2 * left;
    MPIX_Request rq[2];
    MPI_Status st[2];
    int cnt;

    /* eager, stream waits */
    {
        int v = 10 * t_rank + 3, got = -1;
        memset(st, 0xff, sizeof st);
        T_OK(MPIX_Irecv_enqueue(&got, 1, MPI_INT, left, 21, half, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Isend_enqueue(&v, 1, MPI_INT, right, 21, half, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Waitall_enqueue(2, rq, st, MPIX_QUEUE_CUDA_STREAM, &stream));
        if (t_gpu) T_CUDA(cudaStreamSynchronize(stream));
        T_CHECK(got == 10 * left_world + 3, "eager on half: got %d, expected %d", got, 10 * left_world + 3);
        T_CHECK(st[0].MPI_SOURCE == left && st[0].MPI_TAG == 21, "status source %d (comm rank %d expected) tag %d", st[0].MPI_SOURCE, left, st[0].MPI_TAG);
    }
    /* rendezvous (1 MiB), heap buffers, host waits */
    {
        const int big = 1 << 18;
        int *sb, *rb, *h = (int *)malloc((size_t)big * 4);
        T_OK(MPIX_Alloc_mem((size_t)big * 4, MPI_INFO_NULL, &sb)); T_OK(MPIX_Alloc_mem((size_t)big * 4, MPI_INFO_NULL, &rb));
        for (int i = 0; i < big; i++) h[i] = t_rank * 7 + i;
        put(sb, h, (size_t)big * 4);
        memset(st, 0xff, sizeof st);
        T_OK(MPIX_Irecv_enqueue(rb, big, MPI_INT, left, 22, half, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Isend_enqueue(sb, big, MPI_INT, right, 22, half, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Wait(&rq[1], MPI_STATUS_IGNORE));
        T_OK(MPIX_Wait(&rq[0], &st[0]));
        put(h, rb, (size_t)big * 4);
        int bad = 0;
        for (int i = 0; i < big; i++) if (h[i] != left_world * 7 + i) bad++;
        MPI_Get_count(&st[0], MPI_INT, &cnt);
        T_CHECK(bad == 0 && st[0].MPI_SOURCE == left && cnt == big, "rendezvous on half: bad=%d source=%d count=%d", bad, st[0].MPI_SOURCE, cnt);
        MPI_Barrier(half);
        MPIX_Free_mem(sb); MPIX_Free_mem(rb);
        free(h);
    }
    /* partitioned ring on the half, host-side Pready / Parrived */
    {
        enum { NP = 4, PER = 8 };
        static int ps[NP * PER], pr[NP * PER];
        MPIX_Request pq[2];
        T_OK(MPIX_Psend_init(ps, NP, PER, MPI_INT, right, 23, half, MPI_INFO_NULL, &pq[0]));
        T_OK(MPIX_Precv_init(pr, NP, PER, MPI_INT, left, 23, half, MPI_INFO_NULL, &pq[1]));
        for (int it = 0; it < 3; it++) {
            for (int i = 0; i < NP * PER; i++) { ps[i] = 1000 * t_rank + 10 * it + i; pr[i] = -1; }
            T_OK(MPIX_Startall(2, pq));
            for (int p = 0; p < NP; p++) T_OK(MPIX_Pready_host(p, &pq[0]));
            for (int p = 0; p < NP; p++) { int f = 0; while (!f) T_OK(MPIX_Parrived_host(&pq[1], p, &f)); }
            MPI_Status ps_st[2];
            T_OK(MPIX_Waitall(2, pq, ps_st));
            int bad = 0;
            for (int i = 0; i < NP * PER; i++) if (pr[i] != 1000 * left_world + 10 * it + i) bad++;
            T_CHECK(bad == 0, "partitioned on half, round %d: %d wrong (pr[0]=%d)", it, bad, pr[0]);
        }
        T_OK(MPIX_Request_free(&pq[0])); T_OK(MPIX_Request_free(&pq[1]));
    }
    /* Cartesian communicator: non-periodic line over all ranks; edges talk to MPI_PROC_NULL */
    {
        int dims[1] = {t_size}, periods[1] = {0}, src, dst;
        MPI_Comm line;
        MPI_Cart_create(MPI_COMM_WORLD, 1, dims, periods, 0, &line);
        MPI_Cart_shift(line, 0, 1, &src, &dst);
        int v = 50 + t_rank, got = -5;
        memset(st, 0xff, sizeof st);
        T_OK(MPIX_Irecv_enqueue(&got, 1, MPI_INT, src, 24, line, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Isend_enqueue(&v, 1, MPI_INT, dst, 24, line, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Waitall_enqueue(2, rq, st, MPIX_QUEUE_CUDA_STREAM, &stream));
        if (t_gpu) T_CUDA(cudaStreamSynchronize(stream));
        if (t_rank == 0) T_CHECK(got == -5 && st[0].MPI_SOURCE == MPI_PROC_NULL, "left edge: got %d source %d", got, st[0].MPI_SOURCE);
        else T_CHECK(got == 50 + t_rank - 1 && st[0].MPI_SOURCE == t_rank - 1, "line: got %d source %d", got, st[0].MPI_SOURCE);
        MPI_Comm_free(&line);
    }
    MPI_Comm_free(&half);
    return t_finish("subcomm");
}
