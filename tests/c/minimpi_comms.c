/* Bundled mini-MPI: sub-communicators (MPI_Comm_split), groups, Cartesian topologies.
 * Run with any np >= 1 (more interesting with np >= 4). */
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>

static int rank, size, errs;
#define CHECK(c, ...) do { if (!(c)) { printf("[%d] %s:%d CHECK(%s): ", rank, __FILE__, __LINE__, #c); printf(__VA_ARGS__); printf("\n"); errs++; } } while (0)

int main(int argc, char **argv) {
    MPI_Init(&argc, &argv);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);

    /* --- split into even / odd world ranks, reversed order inside each half (key = -rank) --- */
    MPI_Comm half;
    MPI_Comm_split(MPI_COMM_WORLD, rank % 2, -rank, &half);
    int hr, hs;
    MPI_Comm_rank(half, &hr); MPI_Comm_size(half, &hs);
    int expect_hs = (size + (rank % 2 == 0 ? 1 : 0)) / 2;
    CHECK(hs == expect_hs, "half size %d expected %d", hs, expect_hs);
    /* reversed: the highest world rank of my parity is rank 0 of the half */
    int top = (size - 1) - (((size - 1) % 2) != (rank % 2) ? 1 : 0);
    CHECK(hr == (top - rank) / 2, "half rank %d (world %d, top %d)", hr, rank, top);

    /* group translation back to world ranks */
    MPI_Group gw, gh;
    MPI_Comm_group(MPI_COMM_WORLD, &gw); MPI_Comm_group(half, &gh);
    int me_in_world = -1, gsz = -1, grk = -1;
    MPI_Group_translate_ranks(gh, 1, &hr, gw, &me_in_world);
    MPI_Group_size(gh, &gsz); MPI_Group_rank(gh, &grk);
    CHECK(me_in_world == rank && gsz == hs && grk == hr, "group translate %d size %d rank %d", me_in_world, gsz, grk);
    MPI_Group_free(&gw); MPI_Group_free(&gh);

    /* p2p ring inside the half, comm-relative ranks and status */
    {
        int right = (hr + 1) % hs, left = (hr + hs - 1) % hs, v = 100 * rank + 1, got = -1;
        MPI_Status st;
        MPI_Sendrecv(&v, 1, MPI_INT, right, 9, &got, 1, MPI_INT, MPI_ANY_SOURCE, 9, half, &st);
        /* left neighbour in the reversed order is world rank + 2 (wrapping) */
        int left_world = hr == 0 ? top - 2 * (hs - 1) : rank + 2;
        CHECK(got == 100 * left_world + 1 && st.MPI_SOURCE == left, "half ring: got %d from %d (expected %d from %d)", got, st.MPI_SOURCE, 100 * left_world + 1, left);
    }

    /* collectives on the half */
    {
        int sum = rank, mx = rank;
        MPI_Allreduce(MPI_IN_PLACE, &sum, 1, MPI_INT, MPI_SUM, half);
        MPI_Allreduce(MPI_IN_PLACE, &mx, 1, MPI_INT, MPI_MAX, half);
        int esum = 0;
        for (int r = rank % 2; r < size; r += 2) esum += r;
        CHECK(sum == esum && mx == top, "allreduce on half: sum %d (%d) max %d (%d)", sum, esum, mx, top);
        int b = hr == 0 ? 4711 + rank % 2 : -1;
        MPI_Bcast(&b, 1, MPI_INT, 0, half);
        CHECK(b == 4711 + rank % 2, "bcast on half: %d", b);
        int *all = (int *)malloc(sizeof(int) * hs);
        MPI_Allgather(&rank, 1, MPI_INT, all, 1, MPI_INT, half);
        for (int i = 0; i < hs; i++) CHECK(all[i] == top - 2 * i, "allgather[%d] = %d", i, all[i]);
        free(all);
        MPI_Barrier(half);
    }

    /* handle agreement: only the even half creates extra communicators, then everyone dups WORLD; the new handle
     * must still be the same number everywhere (it is the matching context libacx hashes) */
    {
        MPI_Comm extra1 = MPI_COMM_NULL, extra2 = MPI_COMM_NULL, w2;
        if (rank % 2 == 0) { MPI_Comm_dup(half, &extra1); MPI_Comm_dup(extra1, &extra2); }
        MPI_Comm_dup(MPI_COMM_WORLD, &w2);
        int id = (int)w2, lo = id, hi = id;
        MPI_Allreduce(MPI_IN_PLACE, &lo, 1, MPI_INT, MPI_MIN, MPI_COMM_WORLD);
        MPI_Allreduce(MPI_IN_PLACE, &hi, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
        CHECK(lo == hi, "world dup handle differs across ranks: %d..%d", lo, hi);
        int v = rank, got = -1;
        MPI_Sendrecv(&v, 1, MPI_INT, (rank + 1) % size, 3, &got, 1, MPI_INT, (rank + size - 1) % size, 3, w2, MPI_STATUS_IGNORE);
        CHECK(got == (rank + size - 1) % size, "ring on dup'ed world: %d", got);
        if (extra1 != MPI_COMM_NULL) { MPI_Comm_free(&extra1); MPI_Comm_free(&extra2); }
        MPI_Comm_free(&w2);
    }

    /* MPI_UNDEFINED colour */
    {
        MPI_Comm only0;
        MPI_Comm_split(MPI_COMM_WORLD, rank == 0 ? 0 : MPI_UNDEFINED, 0, &only0);
        CHECK((rank == 0) == (only0 != MPI_COMM_NULL), "undefined colour");
        if (only0 != MPI_COMM_NULL) { int s1; MPI_Comm_size(only0, &s1); CHECK(s1 == 1, "size %d", s1); MPI_Comm_free(&only0); }
    }

    /* Cartesian 2-D periodic grid over all ranks */
    {
        int dims[2] = {0, 0}, periods[2] = {1, 0}, coords[2], r2, src, dst;
        MPI_Dims_create(size, 2, dims);
        CHECK(dims[0] * dims[1] == size && dims[0] >= dims[1], "dims %d x %d", dims[0], dims[1]);
        MPI_Comm cart;
        MPI_Cart_create(MPI_COMM_WORLD, 2, dims, periods, 0, &cart);
        int cr;
        MPI_Comm_rank(cart, &cr);
        MPI_Cart_coords(cart, cr, 2, coords);
        MPI_Cart_rank(cart, coords, &r2);
        CHECK(r2 == cr && coords[0] == cr / dims[1] && coords[1] == cr % dims[1], "coords (%d,%d) of %d", coords[0], coords[1], cr);
        /* periodic dimension 0: shift by one, exchange ranks */
        MPI_Cart_shift(cart, 0, 1, &src, &dst);
        int got = -1;
        MPI_Sendrecv(&cr, 1, MPI_INT, dst, 1, &got, 1, MPI_INT, src, 1, cart, MPI_STATUS_IGNORE);
        CHECK(got == src, "cart shift dim0: got %d expected %d", got, src);
        /* non-periodic dimension 1: edges see MPI_PROC_NULL */
        MPI_Cart_shift(cart, 1, 1, &src, &dst);
        CHECK((coords[1] == 0) == (src == MPI_PROC_NULL) && (coords[1] == dims[1] - 1) == (dst == MPI_PROC_NULL), "cart edges src %d dst %d", src, dst);
        got = -7;
        MPI_Sendrecv(&cr, 1, MPI_INT, dst, 2, &got, 1, MPI_INT, src, 2, cart, MPI_STATUS_IGNORE);
        CHECK(src == MPI_PROC_NULL ? got == -7 : got == src, "cart shift dim1: got %d src %d", got, src);
        MPI_Comm_free(&cart);
    }

    MPI_Comm_free(&half);
    MPI_Allreduce(MPI_IN_PLACE, &errs, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
    if (rank == 0) printf("minimpi_comms: %s (np=%d)\n", errs ? "FAIL" : "PASS", size);
    MPI_Finalize();
    return errs != 0;
}
