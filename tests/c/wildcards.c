/* MPI_ANY_SOURCE / MPI_ANY_TAG receives. The reference gets them for free from its MPI library; here the host
 * transport passes them to the MPI library too, and the device transport matches them in a resolver kernel at wait
 * time (experimental: ACX_WILDCARDS=1; without it the program checks that they are rejected cleanly and stops).
 *  1. eager, ANY_SOURCE + ANY_TAG, status carries the real source and tag
 *  2. rendezvous-size (1 MiB) message from a known source with ANY_TAG, symmetric-heap buffers (receiver pulls)
 *  3. rendezvous-size message into a PRIVATE buffer with ANY_SOURCE (bounce + copy-out), completed by a host MPIX_Wait
 *  4. fan-in: rank 0 receives one message from every other rank with ANY_SOURCE; each source must show up exactly once */
#include "acx_test.h"

static void put(void *dst, const void *src, size_t n) { if (t_gpu) T_CUDA(cudaMemcpy(dst, src, n, cudaMemcpyDefault)); else memcpy(dst, src, n); }
static void get(void *dst, const void *src, size_t n) { put(dst, src, n); }

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    cudaStream_t stream = 0;
    MPIX_Request rq[2];
    MPI_Status st[2];
    int cnt;
    if (t_gpu && !getenv("ACX_WILDCARDS")) {
        int w;
        MPIX_Request q = MPIX_REQUEST_NULL;
        int rc = MPIX_Irecv_enqueue(&w, 1, MPI_INT, MPI_ANY_SOURCE, 1, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, &stream);
        T_CHECK(rc != MPI_SUCCESS && q == MPIX_REQUEST_NULL, "wildcards must be rejected without ACX_WILDCARDS=1 (rc=%d)", rc);
        return t_finish("wildcards(rejected)");
    }

    /* 1. eager */
    {
        int v = 90 + t_rank, w = -1;
        memset(st, 0xff, sizeof st);
        T_OK(MPIX_Irecv_enqueue(&w, 1, MPI_INT, MPI_ANY_SOURCE, MPI_ANY_TAG, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Isend_enqueue(&v, 1, MPI_INT, T_RIGHT, 40 + t_rank, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Waitall_enqueue(2, rq, st, MPIX_QUEUE_CUDA_STREAM, &stream));
        if (t_gpu) T_CUDA(cudaStreamSynchronize(stream));
        MPI_Get_count(&st[0], MPI_INT, &cnt);
        T_CHECK(w == 90 + T_LEFT && st[0].MPI_SOURCE == T_LEFT && st[0].MPI_TAG == 40 + T_LEFT && cnt == 1 && st[0].MPI_ERROR == MPI_SUCCESS,
                "eager wildcard: got %d from %d tag %d count %d err %d", w, st[0].MPI_SOURCE, st[0].MPI_TAG, cnt, st[0].MPI_ERROR);
    }
    MPI_Barrier(MPI_COMM_WORLD);

    /* 2. rendezvous, heap buffers, ANY_TAG from a known source */
    const int big = 1 << 18;   /* ints = 1 MiB */
    int *h = (int *)malloc((size_t)big * 4);
    {
        int *sb, *rb;
        T_OK(MPIX_Alloc_mem((size_t)big * 4, MPI_INFO_NULL, &sb)); T_OK(MPIX_Alloc_mem((size_t)big * 4, MPI_INFO_NULL, &rb));
        for (int i = 0; i < big; i++) h[i] = 1000 * t_rank + i;
        put(sb, h, (size_t)big * 4);
        memset(h, 0, (size_t)big * 4); put(rb, h, (size_t)big * 4);
        memset(st, 0xff, sizeof st);
        T_OK(MPIX_Isend_enqueue(sb, big, MPI_INT, T_RIGHT, 77, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Irecv_enqueue(rb, big, MPI_INT, T_LEFT, MPI_ANY_TAG, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Waitall_enqueue(2, rq, st, MPIX_QUEUE_CUDA_STREAM, &stream));
        if (t_gpu) T_CUDA(cudaStreamSynchronize(stream));
        get(h, rb, (size_t)big * 4);
        int bad = 0;
        for (int i = 0; i < big; i++) if (h[i] != 1000 * T_LEFT + i) bad++;
        MPI_Get_count(&st[0], MPI_INT, &cnt);
        T_CHECK(bad == 0 && st[0].MPI_SOURCE == T_LEFT && st[0].MPI_TAG == 77 && cnt == big, "rendezvous ANY_TAG: bad=%d src=%d tag=%d cnt=%d", bad, st[0].MPI_SOURCE, st[0].MPI_TAG, cnt);
        MPI_Barrier(MPI_COMM_WORLD);
        MPIX_Free_mem(sb); MPIX_Free_mem(rb);
    }

    /* 3. rendezvous into a private buffer, ANY_SOURCE, host wait */
    {
        int *sb, *rb;
        T_OK(MPIX_Alloc_mem((size_t)big * 4, MPI_INFO_NULL, &sb));
        if (t_gpu) T_CUDA(cudaMalloc((void **)&rb, (size_t)big * 4)); else rb = (int *)malloc((size_t)big * 4);
        for (int i = 0; i < big; i++) h[i] = 7 * t_rank - i;
        put(sb, h, (size_t)big * 4);
        memset(st, 0xff, sizeof st);
        T_OK(MPIX_Irecv_enqueue(rb, big, MPI_INT, MPI_ANY_SOURCE, 78, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Isend_enqueue(sb, big, MPI_INT, T_RIGHT, 78, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Wait(&rq[0], &st[0]));
        T_OK(MPIX_Wait(&rq[1], MPI_STATUS_IGNORE));
        get(h, rb, (size_t)big * 4);
        int bad = 0;
        for (int i = 0; i < big; i++) if (h[i] != 7 * T_LEFT - i) bad++;
        T_CHECK(bad == 0 && st[0].MPI_SOURCE == T_LEFT && st[0].MPI_TAG == 78, "private ANY_SOURCE: bad=%d src=%d tag=%d", bad, st[0].MPI_SOURCE, st[0].MPI_TAG);
        MPI_Barrier(MPI_COMM_WORLD);
        MPIX_Free_mem(sb);
        if (t_gpu) cudaFree(rb); else free(rb);
    }
    free(h);

    /* 4. fan-in to rank 0 */
    if (t_size > 1) {
        if (t_rank == 0) {
            int seen[64] = {0};
            for (int i = 1; i < t_size; i++) {
                int w = -1;
                MPIX_Request q;
                MPI_Status s1;
                T_OK(MPIX_Irecv_enqueue(&w, 1, MPI_INT, MPI_ANY_SOURCE, 5, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, &stream));
                T_OK(MPIX_Wait_enqueue(&q, &s1, MPIX_QUEUE_CUDA_STREAM, &stream));
                if (t_gpu) T_CUDA(cudaStreamSynchronize(stream));
                T_CHECK(s1.MPI_SOURCE >= 1 && s1.MPI_SOURCE < t_size && w == 300 + s1.MPI_SOURCE, "fan-in: value %d from %d", w, s1.MPI_SOURCE);
                if (s1.MPI_SOURCE >= 0 && s1.MPI_SOURCE < 64) seen[s1.MPI_SOURCE]++;
            }
            for (int i = 1; i < t_size && i < 64; i++) T_CHECK(seen[i] == 1, "source %d seen %d times", i, seen[i]);
        } else {
            int v = 300 + t_rank;
            MPIX_Request q;
            T_OK(MPIX_Isend_enqueue(&v, 1, MPI_INT, 0, 5, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, &stream));
            T_OK(MPIX_Wait_enqueue(&q, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
            if (t_gpu) T_CUDA(cudaStreamSynchronize(stream));
        }
    }
    return t_finish("wildcards");
}
