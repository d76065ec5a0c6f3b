/* Bundled mini-MPI: derived datatypes (contiguous, vector, hvector, subarray and nestings) — the strided layouts of
 * halo faces. Every case sends with one layout and receives with another, so both pack and unpack are checked against
 * index arithmetic done here. */
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int rank, size, errs;
#define CHECK(c, ...) do { if (!(c)) { printf("[%d] %s:%d CHECK(%s): ", rank, __FILE__, __LINE__, #c); printf(__VA_ARGS__); printf("\n"); errs++; } } while (0)

enum { NX = 6, NY = 5, NZ = 4 };   /* a[z][y][x], x fastest */

int main(int argc, char **argv) {
    MPI_Init(&argc, &argv);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    const int right = (rank + 1) % size, left = (rank + size - 1) % size;
    static int a[NZ][NY][NX], b[NZ][NY][NX];
    for (int z = 0; z < NZ; z++) for (int y = 0; y < NY; y++) for (int x = 0; x < NX; x++) a[z][y][x] = 10000 * rank + 100 * z + 10 * y + x;
    MPI_Status st;
    int cnt;

    /* 1. column x=2 of plane z=1 (vector: NY blocks of 1, stride NX) -> contiguous buffer */
    {
        MPI_Datatype col;
        MPI_Type_vector(NY, 1, NX, MPI_INT, &col);
        MPI_Type_commit(&col);
        MPI_Aint lb, ext; int sz;
        MPI_Type_get_extent(col, &lb, &ext); MPI_Type_size(col, &sz);
        CHECK(sz == NY * 4 && ext == ((NY - 1) * NX + 1) * 4 && lb == 0, "vector size %d extent %lld", sz, (long long)ext);
        int got[NY];
        MPI_Sendrecv(&a[1][0][2], 1, col, right, 1, got, NY, MPI_INT, left, 1, MPI_COMM_WORLD, &st);
        MPI_Get_count(&st, MPI_INT, &cnt);
        for (int y = 0; y < NY; y++) CHECK(got[y] == 10000 * left + 100 + 10 * y + 2, "col[%d] = %d", y, got[y]);
        CHECK(cnt == NY, "count %d", cnt);
        /* and back: contiguous -> strided column of b */
        memset(b, 0xff, sizeof b);
        MPI_Sendrecv(got, NY, MPI_INT, left, 2, &b[2][0][3], 1, col, right, 2, MPI_COMM_WORLD, &st);
        MPI_Get_count(&st, col, &cnt);
        for (int y = 0; y < NY; y++) CHECK(b[2][y][3] == 10000 * rank + 100 + 10 * y + 2, "b col[%d] = %d", y, b[2][y][3]);   /* my own column came back */
        CHECK(b[2][0][2] == -1 && b[2][0][4] == -1 && cnt == 1, "neighbours untouched / count %d", cnt);
        MPI_Type_free(&col);
    }
    /* 2. x-face (x = NX-1 for all y,z) as a 3-D subarray -> z-face (z = 0) subarray of b: both NY*NZ vs NX*NY elements differ,
     *    so receive the x-face into a contiguous scratch and check; then y-face via vector-of-vector */
    {
        int sizes[3] = {NZ, NY, NX}, sub[3] = {NZ, NY, 1}, start[3] = {0, 0, NX - 1};
        MPI_Datatype xface;
        MPI_Type_create_subarray(3, sizes, sub, start, MPI_ORDER_C, MPI_INT, &xface);
        MPI_Type_commit(&xface);
        int sz; MPI_Aint ext;
        MPI_Type_size(xface, &sz); MPI_Type_get_extent(xface, NULL, &ext);
        CHECK(sz == NZ * NY * 4 && ext == (MPI_Aint)sizeof a, "subarray size %d extent %lld", sz, (long long)ext);
        int got[NZ * NY];
        MPI_Sendrecv(a, 1, xface, right, 3, got, NZ * NY, MPI_INT, left, 3, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
        for (int z = 0; z < NZ; z++) for (int y = 0; y < NY; y++) CHECK(got[z * NY + y] == 10000 * left + 100 * z + 10 * y + NX - 1, "xface[%d][%d] = %d", z, y, got[z * NY + y]);
        /* receive side subarray: put it into the x=0 ghost face of b */
        int start0[3] = {0, 0, 0};
        MPI_Datatype x0;
        MPI_Type_create_subarray(3, sizes, sub, start0, MPI_ORDER_C, MPI_INT, &x0);
        memset(b, 0xff, sizeof b);
        MPI_Sendrecv(a, 1, xface, right, 4, b, 1, x0, left, 4, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
        int bad = 0;
        for (int z = 0; z < NZ; z++) for (int y = 0; y < NY; y++) for (int x = 0; x < NX; x++) {
            int want = x == 0 ? 10000 * left + 100 * z + 10 * y + NX - 1 : -1;
            if (b[z][y][x] != want) bad++;
        }
        CHECK(bad == 0, "x-face -> x0 ghost face: %d wrong", bad);
        MPI_Type_free(&xface); MPI_Type_free(&x0);
    }
    /* 3. y-face (y = 1, all x, all z): NZ blocks of NX ints, stride NY*NX = one vector; then a 2x2 tile of it via nesting */
    {
        MPI_Datatype yface, row2, tile;
        MPI_Type_vector(NZ, NX, NY * NX, MPI_INT, &yface);
        int got[NZ * NX];
        MPI_Sendrecv(&a[0][1][0], 1, yface, right, 5, got, NZ * NX, MPI_INT, left, 5, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
        for (int z = 0; z < NZ; z++) for (int x = 0; x < NX; x++) CHECK(got[z * NX + x] == 10000 * left + 100 * z + 10 + x, "yface[%d][%d]", z, x);
        /* nested: row2 = 2 ints every 3 (x = 0,3), tile = 2 such rows one plane apart (hvector) */
        MPI_Type_vector(2, 1, 3, MPI_INT, &row2);
        MPI_Type_create_hvector(2, 1, (MPI_Aint)(NY * NX * sizeof(int)), row2, &tile);
        int t4[4];
        MPI_Sendrecv(&a[0][2][0], 1, tile, right, 6, t4, 4, MPI_INT, left, 6, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
        CHECK(t4[0] == 10000 * left + 20 && t4[1] == 10000 * left + 23 && t4[2] == 10000 * left + 120 && t4[3] == 10000 * left + 123, "nested tile %d %d %d %d", t4[0], t4[1], t4[2], t4[3]);
        /* count > 1 of a strided type: 2 columns (extent of row2 is 4 ints), i.e. x = 0,3 and x = 4,7->(wraps into next row) */
        int c2[4];
        MPI_Sendrecv(&a[0][0][0], 2, row2, right, 7, c2, 4, MPI_INT, left, 7, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
        int *flat = &a[0][0][0];
        (void)flat;
        CHECK(c2[0] == 10000 * left + 0 && c2[1] == 10000 * left + 3 && c2[2] == 10000 * left + 4 && c2[3] == 10000 * left + 11, "count=2 of row2: %d %d %d %d", c2[0], c2[1], c2[2], c2[3]);
        MPI_Type_free(&yface); MPI_Type_free(&row2); MPI_Type_free(&tile);
    }
    /* 4. contiguous-of-contiguous folds to plain bytes; vector with stride == blocklen is contiguous */
    {
        MPI_Datatype c3, c6, v;
        MPI_Type_contiguous(3, MPI_INT, &c3);
        MPI_Type_contiguous(2, c3, &c6);
        MPI_Type_vector(4, 2, 2, MPI_INT, &v);
        int s6, sv; MPI_Aint e6, ev;
        MPI_Type_size(c6, &s6); MPI_Type_get_extent(c6, NULL, &e6);
        MPI_Type_size(v, &sv); MPI_Type_get_extent(v, NULL, &ev);
        CHECK(s6 == 24 && e6 == 24 && sv == 32 && ev == 32, "sizes %d %lld %d %lld", s6, (long long)e6, sv, (long long)ev);
        int got[6];
        MPI_Sendrecv(&a[0][0][0], 1, c6, right, 8, got, 6, MPI_INT, left, 8, MPI_COMM_WORLD, &st);
        MPI_Get_count(&st, c3, &cnt);
        CHECK(got[5] == 10000 * left + 5 && cnt == 2, "contiguous: %d count %d", got[5], cnt);
        MPI_Type_free(&c3); MPI_Type_free(&c6); MPI_Type_free(&v);
    }
    MPI_Allreduce(MPI_IN_PLACE, &errs, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
    if (rank == 0) printf("minimpi_types: %s (np=%d)\n", errs ? "FAIL" : "PASS", size);
    MPI_Finalize();
    return errs != 0;
}
