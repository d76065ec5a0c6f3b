/* mini-MPI self-test: p2p (eager + multi-cell), wildcard + unexpected matching,
 * partitioned rounds, collectives. Exit code != 0 on failure. */
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv) {
    int provided, rank, size, errs = 0;
    MPI_Init_thread(&argc, &argv, MPI_THREAD_MULTIPLE, &provided);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    int right = (rank + 1) % size, left = (rank + size - 1) % size;

    /* small ring */
    int s = rank + 1, r = -1;
    MPI_Request q[2];
    MPI_Status st;
    MPI_Isend(&s, 1, MPI_INT, right, 5, MPI_COMM_WORLD, &q[0]);
    MPI_Irecv(&r, 1, MPI_INT, left, 5, MPI_COMM_WORLD, &q[1]);
    MPI_Wait(&q[0], MPI_STATUS_IGNORE);
    MPI_Wait(&q[1], &st);
    if (r != left + 1 || st.MPI_SOURCE != left || st.MPI_TAG != 5) { printf("[%d] small ring bad r=%d\n", rank, r); errs++; }

    /* large ring (many cells), both directions at once */
    size_t n = 3u << 20;
    unsigned char *sb = malloc(n), *rb = malloc(n);
    for (size_t i = 0; i < n; i++) sb[i] = (unsigned char)(i * 7 + rank);
    memset(rb, 0, n);
    MPI_Irecv(rb, (int)n, MPI_BYTE, left, 6, MPI_COMM_WORLD, &q[1]);
    MPI_Isend(sb, (int)n, MPI_BYTE, right, 6, MPI_COMM_WORLD, &q[0]);
    MPI_Waitall(2, q, MPI_STATUSES_IGNORE);
    for (size_t i = 0; i < n; i++) if (rb[i] != (unsigned char)(i * 7 + left)) { printf("[%d] large ring bad at %zu\n", rank, i); errs++; break; }

    /* unexpected + out-of-order tags + wildcard */
    int a = 100 + rank, b = 200 + rank, ra = -1, rb2 = -1;
    MPI_Send(&a, 1, MPI_INT, right, 1, MPI_COMM_WORLD);
    MPI_Send(&b, 1, MPI_INT, right, 2, MPI_COMM_WORLD);
    MPI_Barrier(MPI_COMM_WORLD);
    MPI_Recv(&rb2, 1, MPI_INT, left, 2, MPI_COMM_WORLD, &st);
    MPI_Recv(&ra, 1, MPI_INT, MPI_ANY_SOURCE, MPI_ANY_TAG, MPI_COMM_WORLD, &st);
    int cnt = -1;
    MPI_Get_count(&st, MPI_INT, &cnt);
    if (ra != 100 + left || rb2 != 200 + left || st.MPI_TAG != 1 || st.MPI_SOURCE != left || cnt != 1) { printf("[%d] matching bad %d %d\n", rank, ra, rb2); errs++; }

    /* partitioned */
    enum { P = 8, C = 1000 };
    static double ps[P * C], pr[P * C];
    MPI_Request pq[2];
    MPI_Psend_init(ps, P, C, MPI_DOUBLE, right, 9, MPI_COMM_WORLD, MPI_INFO_NULL, &pq[0]);
    MPI_Precv_init(pr, P, C, MPI_DOUBLE, left, 9, MPI_COMM_WORLD, MPI_INFO_NULL, &pq[1]);
    for (int it = 0; it < 3; it++) {
        for (int i = 0; i < P * C; i++) { ps[i] = rank * 1e6 + it * 1e4 + i; pr[i] = -1; }
        MPI_Startall(2, pq);
        for (int p = P - 1; p >= 0; p--) MPI_Pready(p, pq[0]);
        for (int p = 0; p < P; p++) { int f = 0; while (!f) MPI_Parrived(pq[1], p, &f); if (pr[p * C] != left * 1e6 + it * 1e4 + p * C) { errs++; printf("[%d] part %d bad\n", rank, p); } }
        MPI_Waitall(2, pq, MPI_STATUSES_IGNORE);
        for (int i = 0; i < P * C; i++) if (pr[i] != left * 1e6 + it * 1e4 + i) { errs++; printf("[%d] partitioned payload bad\n", rank); break; }
    }
    MPI_Request_free(&pq[0]);
    MPI_Request_free(&pq[1]);

    /* collectives */
    double v = rank + 1, sum = 0;
    MPI_Allreduce(&v, &sum, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    if (sum != size * (size + 1) / 2.0) { errs++; printf("[%d] allreduce bad %f\n", rank, sum); }
    int root_val = rank == 0 ? 777 : -1;
    MPI_Bcast(&root_val, 1, MPI_INT, 0, MPI_COMM_WORLD);
    if (root_val != 777) errs++;
    int *all = malloc(sizeof(int) * size);
    MPI_Allgather(&rank, 1, MPI_INT, all, 1, MPI_INT, MPI_COMM_WORLD);
    for (int i = 0; i < size; i++) if (all[i] != i) errs++;

    MPI_Allreduce(MPI_IN_PLACE, &errs, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
    if (rank == 0) printf("minimpi_smoke: %s (np=%d)\n", errs ? "FAIL" : "PASS", size);
    MPI_Finalize();
    return errs != 0;
}
