// CPU check of the index arithmetic the pack / unpack kernels use for strided datatypes (src/core/internal.h:
// acx_device_layout, acx_layout_units, acx_layout_unit_offset) against the bundled MPI's own pack routine: the packed
// image produced unit by unit must equal what MPI_Sendrecv-to-self delivers into a contiguous buffer. Runs anywhere
// (no GPU needed): the functions under test are __host__ __device__.
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../src/core/internal.h"

static int errs;
#define CHECK(c, ...) do { if (!(c)) { printf("%s:%d CHECK(%s): ", __FILE__, __LINE__, #c); printf(__VA_ARGS__); printf("\n"); errs++; } } while (0)

static void check(const char *name, MPI_Datatype dt, int count, size_t misalign) {
    ACX_Mpi_layout M;
    CHECK(ACX_Mpi_type_layout(dt, &M) == MPI_SUCCESS, "%s: layout", name);
    AcxLayout a;
    memset(&a, 0, sizeof a);
    long long reach = M.disp + M.run;
    for (int d = 0; d < 3; d++) { a.n[d] = M.n[d]; a.stride[d] = M.stride[d]; reach += (M.n[d] - 1) * M.stride[d]; }
    a.run = M.run; a.disp = M.disp; a.extent = M.extent; a.count = count;
    a.span = (long long)(count - 1) * M.extent + reach;
    const size_t packed_bytes = (size_t)M.size * count;
    std::vector<unsigned char> user((size_t)a.span + 64 + misalign), ref(packed_bytes), mine(packed_bytes + 64, 0xEE);
    for (size_t i = 0; i < user.size(); i++) user[i] = (unsigned char)(i * 131 + 7);
    unsigned char *ubase = user.data() + misalign;
    // reference: the MPI library packs (send typed to self, receive as bytes)
    MPI_Sendrecv(ubase, count, dt, 0, 1, ref.data(), (int)packed_bytes, MPI_BYTE, 0, 1, MPI_COMM_SELF, MPI_STATUS_IGNORE);
    unsigned char *pbase = mine.data() + ((16 - ((uintptr_t)mine.data() & 15)) & 15);   // the staging buffer is 256-B aligned in the library
    TypeLayout L = acx_device_layout(a, ubase, pbase);
    CHECK(L.vec >= 1 && L.vec <= 16 && (L.vec & (L.vec - 1)) == 0 && M.run % L.vec == 0, "%s: vec %u run %lld", name, L.vec, M.run);
    if (misalign & 1) CHECK(L.vec == 1, "%s: misaligned base must force byte units (vec %u)", name, L.vec);
    const uint64_t units = acx_layout_units(L, ~0ull);
    CHECK(units * L.vec == packed_bytes, "%s: units %llu x %u != %zu", name, (unsigned long long)units, L.vec, packed_bytes);
    for (uint64_t j = 0; j < units; j++) {
        long long off = acx_layout_unit_offset(L, j);
        if (off < 0 || off + (long long)L.vec > a.span) { CHECK(0, "%s: unit %llu offset %lld outside span %lld", name, (unsigned long long)j, off, a.span); return; }
        memcpy(pbase + j * L.vec, ubase + off, L.vec);
    }
    CHECK(memcmp(pbase, ref.data(), packed_bytes) == 0, "%s (count %d, misalign %zu, vec %u): packed image differs", name, count, misalign, L.vec);
    // truncated message: only whole units below the limit
    const uint64_t lim = packed_bytes / 3;
    CHECK(acx_layout_units(L, lim) == lim / L.vec, "%s: truncated units", name);
}

int main(int argc, char **argv) {
    MPI_Init(&argc, &argv);
    MPI_Datatype col, yface, xface, row2, tile, c6, dbl_col, sub2d;
    MPI_Type_vector(5, 1, 6, MPI_INT, &col);
    MPI_Type_vector(4, 6, 30, MPI_INT, &yface);
    int sizes[3] = {4, 5, 6}, sub[3] = {4, 5, 1}, start[3] = {0, 0, 5};
    MPI_Type_create_subarray(3, sizes, sub, start, MPI_ORDER_C, MPI_INT, &xface);
    MPI_Type_vector(2, 1, 3, MPI_INT, &row2);
    MPI_Type_create_hvector(2, 1, 120, row2, &tile);
    MPI_Type_contiguous(6, MPI_INT, &c6);
    MPI_Type_vector(7, 2, 9, MPI_DOUBLE, &dbl_col);   // 16-byte runs: exercises vec = 16
    int s2[2] = {16, 32}, b2[2] = {8, 8}, o2[2] = {4, 8};
    MPI_Type_create_subarray(2, s2, b2, o2, MPI_ORDER_FORTRAN, MPI_FLOAT, &sub2d);
    struct { const char *name; MPI_Datatype dt; } cases[] = {{"col", col}, {"yface", yface}, {"xface", xface}, {"row2", row2}, {"tile", tile},
                                                           {"c6", c6}, {"dbl_col", dbl_col}, {"sub2d_fortran", sub2d}, {"int", MPI_INT}};
    for (auto &c : cases)
        for (int count = 1; count <= 3; count++)
            for (size_t mis : {(size_t)0, (size_t)1, (size_t)4, (size_t)8}) check(c.name, c.dt, count, mis);
    printf("layout_math: %s\n", errs ? "FAIL" : "PASS");
    MPI_Finalize();
    return errs != 0;
}
