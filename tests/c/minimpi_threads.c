/* MPI_THREAD_MULTIPLE in the mini-MPI: two threads per rank drive independent tag streams through the
 * same pair rings concurrently (this is how the reference's proxy thread and application thread share
 * the MPI library, src/init.cpp:55-154 + src/sendrecv.cu:627). */
#include <mpi.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

static int g_rank, g_size, g_errs;

static void *worker(void *arg) {
    int tid = (int)(long)arg, right = (g_rank + 1) % g_size, left = (g_rank + g_size - 1) % g_size;
    for (int it = 0; it < 300; it++) {
        int n = 1 + (it * 37 + tid * 11) % 5000;
        int *s = malloc(n * sizeof(int)), *r = malloc(n * sizeof(int));
        for (int i = 0; i < n; i++) { s[i] = g_rank * 1000003 + tid * 7919 + it * 31 + i; r[i] = -1; }
        MPI_Request q[2];
        MPI_Irecv(r, n, MPI_INT, left, 100 + tid, MPI_COMM_WORLD, &q[0]);
        MPI_Isend(s, n, MPI_INT, right, 100 + tid, MPI_COMM_WORLD, &q[1]);
        MPI_Waitall(2, q, MPI_STATUSES_IGNORE);
        for (int i = 0; i < n; i++) if (r[i] != left * 1000003 + tid * 7919 + it * 31 + i) { __sync_fetch_and_add(&g_errs, 1); break; }
        free(s); free(r);
    }
    return NULL;
}

int main(int argc, char **argv) {
    int prov;
    MPI_Init_thread(&argc, &argv, MPI_THREAD_MULTIPLE, &prov);
    MPI_Comm_rank(MPI_COMM_WORLD, &g_rank);
    MPI_Comm_size(MPI_COMM_WORLD, &g_size);
    pthread_t th[2];
    for (long t = 0; t < 2; t++) pthread_create(&th[t], NULL, worker, (void *)t);
    for (int t = 0; t < 2; t++) pthread_join(th[t], NULL);
    MPI_Allreduce(MPI_IN_PLACE, &g_errs, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
    if (g_rank == 0) printf("minimpi_threads: %s (np=%d)\n", g_errs ? "FAIL" : "PASS", g_size);
    MPI_Finalize();
    return g_errs != 0;
}
