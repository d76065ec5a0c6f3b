/* Derived datatypes through the MPIX API: the strided faces of a 3-D block are sent and received in place, as a
 * halo exchange written against the reference (which hands the datatype to its MPI library) would do.
 * Device transport: pack_kernel gathers the face into a heap staging buffer before the send; the message lands in a
 * staging buffer on the other side and unpack_kernel scatters it after the wait kernel.
 *  small block (8^3 ints):  faces are eager messages, host (static) arrays
 *  large block (64^3 floats): faces are rendezvous messages (63 x 63 x 4 B), symmetric-heap arrays; waits on the stream and on the host */
#include "acx_test.h"

static void put(void *dst, const void *src, size_t n) { if (t_gpu) T_CUDA(cudaMemcpy(dst, src, n, cudaMemcpyDefault)); else memcpy(dst, src, n); }

#define IDX(n, z, y, x) (((size_t)(z) * (n) + (y)) * (n) + (x))

static void faces(int n, int use_heap, int host_wait, MPI_Datatype elem, int tagbase) {
    const size_t cells = (size_t)n * n * n;
    int *h = (int *)malloc(cells * 4), *a;
    if (use_heap) T_OK(MPIX_Alloc_mem(cells * 4, MPI_INFO_NULL, &a));
    else { static int small[8 * 8 * 8]; a = small; }
    int sizes[3] = {n, n, n}, start[3], sub[3];
    MPI_Datatype send_x, recv_x, face_y, row, face_z;
    const int m = n - 1;   /* faces span coordinates 1..n-1 in their two free dimensions: the six buffers are pairwise disjoint
                            * (concurrent operations on overlapping buffers are erroneous MPI, and TSan on the software device says so) */
    /* x faces: column pencils (run = 1 element) as subarrays: send interior x=n-2 to the right, receive into ghost x=0 from the left */
    sub[0] = m; sub[1] = m; sub[2] = 1;
    start[0] = 1; start[1] = 1; start[2] = n - 2; MPI_Type_create_subarray(3, sizes, sub, start, MPI_ORDER_C, elem, &send_x);
    start[2] = 0; MPI_Type_create_subarray(3, sizes, sub, start, MPI_ORDER_C, elem, &recv_x);
    /* y faces as a vector: m blocks (one per z) of m contiguous elements, stride one plane; the buffer argument picks the face */
    MPI_Type_vector(m, m, n * n, elem, &face_y);
    /* z faces: a contiguous row type nested in an hvector (m rows, one block each, a row pitch apart) */
    MPI_Type_contiguous(m, elem, &row);
    MPI_Type_create_hvector(m, 1, (MPI_Aint)n * 4, row, &face_z);
    MPI_Datatype *all[5] = {&send_x, &recv_x, &face_y, &row, &face_z};
    for (int i = 0; i < 5; i++) MPI_Type_commit(all[i]);

    for (size_t i = 0; i < cells; i++) h[i] = (int)(1000000 * (t_rank + 1) + i);
    /* ghost layers start poisoned */
    for (int z = 0; z < n; z++) for (int y = 0; y < n; y++) h[IDX(n, z, y, 0)] = -1;
    for (int z = 0; z < n; z++) for (int x = 0; x < n; x++) h[IDX(n, z, 0, x)] = -2;
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) h[IDX(n, 0, y, x)] = -3;
    put(a, h, cells * 4);

    cudaStream_t stream = 0;
    MPIX_Request rq[6];
    MPI_Status st[6];
    memset(st, 0xff, sizeof st);
    /* receives into ghosts x=0, y=0, z=0 from the left; sends of interior x=n-2, y=n-2, z=n-2 to the right */
    T_OK(MPIX_Irecv_enqueue(a, 1, recv_x, T_LEFT, tagbase + 0, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Irecv_enqueue(a + IDX(n, 1, 0, 1), 1, face_y, T_LEFT, tagbase + 1, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Irecv_enqueue(a + IDX(n, 0, 1, 1), 1, face_z, T_LEFT, tagbase + 2, MPI_COMM_WORLD, &rq[2], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Isend_enqueue(a, 1, send_x, T_RIGHT, tagbase + 0, MPI_COMM_WORLD, &rq[3], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Isend_enqueue(a + IDX(n, 1, n - 2, 1), 1, face_y, T_RIGHT, tagbase + 1, MPI_COMM_WORLD, &rq[4], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Isend_enqueue(a + IDX(n, n - 2, 1, 1), 1, face_z, T_RIGHT, tagbase + 2, MPI_COMM_WORLD, &rq[5], MPIX_QUEUE_CUDA_STREAM, &stream));
    if (host_wait) { for (int i = 0; i < 6; i++) T_OK(MPIX_Wait(&rq[i], &st[i])); }
    else {
        T_OK(MPIX_Waitall_enqueue(6, rq, st, MPIX_QUEUE_CUDA_STREAM, &stream));
        if (t_gpu) T_CUDA(cudaStreamSynchronize(stream));
    }
    put(h, a, cells * 4);
    /* Sender's data is ITS array before the exchange: interior cells carry 1000000*(left+1) + index. */
    const int L = 1000000 * (T_LEFT + 1);
    int bad = 0;
    for (int y = 1; y < n; y++) for (int x = 1; x < n; x++) if (h[IDX(n, 0, y, x)] != L + (int)IDX(n, n - 2, y, x)) bad++;          /* z ghost <- left's z=n-2 plane */
    for (int z = 1; z < n; z++) for (int x = 1; x < n; x++) if (h[IDX(n, z, 0, x)] != L + (int)IDX(n, z, n - 2, x)) bad++;          /* y ghost <- left's y=n-2 plane */
    for (int z = 1; z < n; z++) for (int y = 1; y < n; y++) if (h[IDX(n, z, y, 0)] != L + (int)IDX(n, z, y, n - 2)) bad++;          /* x ghost <- left's x=n-2 plane */
    /* interior untouched */
    for (int z = 1; z < n; z++) for (int y = 1; y < n; y++) for (int x = 1; x < n; x++) if (h[IDX(n, z, y, x)] != 1000000 * (t_rank + 1) + (int)IDX(n, z, y, x)) bad++;
    /* ghost edges (two coordinates zero) belong to no face: still poisoned */
    for (int k = 0; k < n; k++) { if (h[IDX(n, k, 0, 0)] >= 0) bad++; if (h[IDX(n, 0, k, 0)] >= 0) bad++; if (h[IDX(n, 0, 0, k)] >= 0) bad++; }
    T_CHECK(bad == 0, "n=%d heap=%d hostwait=%d: %d wrong cells", n, use_heap, host_wait, bad);
    int cnt = -1;
    MPI_Get_count(&st[0], recv_x, &cnt);
    T_CHECK(cnt == 1 && st[0].MPI_SOURCE == T_LEFT && st[0].MPI_TAG == tagbase && st[0].MPI_ERROR == MPI_SUCCESS, "status: count %d source %d tag %d err %d", cnt, st[0].MPI_SOURCE, st[0].MPI_TAG, st[0].MPI_ERROR);
    MPI_Barrier(MPI_COMM_WORLD);
    for (int i = 0; i < 5; i++) MPI_Type_free(all[i]);
    if (use_heap) MPIX_Free_mem(a);
    free(h);
}

/* Column blocks of a row-major matrix with element sizes 1, 2, 8 and 16 bytes: the pack / unpack kernels move the widest
 * unit that divides the run, the strides and both addresses (1-, 2-, 4-, 8- and 16-byte units). `esize` bytes per element,
 * `w` elements per row segment; strided send into a contiguous receive and back into a strided one. */
static void columns(int esize, int w, int rows, int pitch_elems, int tag) {
    const size_t pitch = (size_t)pitch_elems * esize, seg = (size_t)w * esize, total = pitch * rows;
    unsigned char *h = (unsigned char *)malloc(total), *m, *packed, *back;
    T_OK(MPIX_Alloc_mem(total, MPI_INFO_NULL, &m));
    T_OK(MPIX_Alloc_mem(seg * rows, MPI_INFO_NULL, &packed));
    T_OK(MPIX_Alloc_mem(total, MPI_INFO_NULL, &back));
    MPI_Datatype elem, block;
    MPI_Type_contiguous(esize, MPI_BYTE, &elem);
    MPI_Type_vector(rows, w, pitch_elems, elem, &block);
    MPI_Type_commit(&block);
    for (size_t i = 0; i < total; i++) h[i] = (unsigned char)(i * 5 + t_rank * 3 + esize);
    put(m, h, total);
    memset(h, 0xee, total);
    put(back, h, total);
    const size_t off = (size_t)2 * esize;   /* the block starts at column 2 */
    cudaStream_t stream = 0;
    MPIX_Request rq[4];
    T_OK(MPIX_Irecv_enqueue(packed, (int)(seg * rows), MPI_BYTE, T_LEFT, tag, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Isend_enqueue(m + off, 1, block, T_RIGHT, tag, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
    /* and back: the dense image goes out again and is scattered into a strided block on the other side */
    T_OK(MPIX_Irecv_enqueue(back + off, 1, block, T_RIGHT, tag + 1, MPI_COMM_WORLD, &rq[2], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Isend_enqueue(packed, (int)(seg * rows), MPI_BYTE, T_LEFT, tag + 1, MPI_COMM_WORLD, &rq[3], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Waitall_enqueue(2, rq + 2, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
    if (t_gpu) T_CUDA(cudaStreamSynchronize(stream));
    put(h, back, total);
    /* my block travelled right and came back: inside the block my own data, outside untouched */
    size_t bad = 0;
    for (int r = 0; r < rows; r++) for (size_t b = 0; b < pitch; b++) {
        const size_t i = (size_t)r * pitch + b;
        const int inside = b >= off && b < off + seg;
        const unsigned char want = inside ? (unsigned char)(i * 5 + t_rank * 3 + esize) : 0xee;
        if (h[i] != want) bad++;
    }
    T_CHECK(bad == 0, "column block, %d-byte elements x %d: %zu wrong bytes", esize, w, bad);
    MPI_Type_free(&block); MPI_Type_free(&elem);
    T_OK(MPIX_Free_mem(m)); T_OK(MPIX_Free_mem(packed)); T_OK(MPIX_Free_mem(back));
    free(h);
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    columns(1, 3, 40, 11, 140);        /* 3-byte runs at odd offsets: byte units */
    columns(2, 5, 40, 16, 142);        /* 10-byte runs: 2-byte units */
    columns(8, 3, 30, 9, 144);         /* 24-byte runs: 8-byte units */
    columns(16, 4, 300, 10, 146);      /* 64-byte runs: 16-byte units; 19 KB: rendezvous */
    columns(4, 7, 20, 33, 148);        /* 28-byte runs: 4-byte units */
    faces(8, 0, 0, MPI_INT, 100);      /* eager faces (256 B), host arrays, stream waits */
    faces(8, 0, 1, MPI_INT, 110);      /* same, host waits */
    faces(64, 1, 0, MPI_FLOAT, 120);   /* rendezvous faces (16 KiB), heap arrays, stream waits (MPI_FLOAT: same size as int, payload is opaque) */
    faces(64, 1, 1, MPI_INT, 130);     /* host waits */
    return t_finish("datatypes");
}
