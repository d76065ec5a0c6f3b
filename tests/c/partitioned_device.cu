// The device-side partitioned API beyond MPIX_Pready / MPIX_Parrived (include/acx_device.cuh), which the reference does
// not have: cooperative Pready by a warp or a CTA, zero-copy producers (remote_partition_ptr + pready_signal), blocking
// consumers (wait_arrived + local_partition_ptr), stream-ordered rounds (MPIX_Start_enqueue / MPIX_Pwait_enqueue) and
// the explicit-epoch calls persistent kernels use to run many rounds without the host.
#include "acx_test.h"

#define PARTS 6
#define PER 300        /* ints per partition */
#define ROUNDS 5

__device__ __host__ static int value_of(int rank, int round, int part, int i) { return rank * 1000003 + round * 10007 + part * 101 + i; }

// ---- 1. zero-copy: the producer writes straight into the receiver's landing zone --------------------------------------
__global__ void produce_zero_copy(MPIX_Prequest send, int rank, int round) {   // one CTA per partition
    int *dst = (int *)acx::remote_partition_ptr(send, blockIdx.x);            // blocks until the receiver has started the round
    for (int i = threadIdx.x; i < PER; i += blockDim.x) dst[i] = value_of(rank, round, blockIdx.x, i);
    __syncthreads();
    if (threadIdx.x == 0) acx::pready_signal(blockIdx.x, send);
}
__global__ void consume_blocking(MPIX_Prequest recv, int left, int round, int *bad) {   // one CTA per partition
    if (threadIdx.x == 0) acx::wait_arrived(recv, blockIdx.x);
    __syncthreads();
    const int *src = (const int *)acx::local_partition_ptr(recv, blockIdx.x);
    for (int i = threadIdx.x; i < PER; i += blockDim.x) if (src[i] != value_of(left, round, blockIdx.x, i)) atomicAdd(bad, 1);
}

// ---- 2. cooperative copies from a send buffer ---------------------------------------------------------------------------
__global__ void ready_by_cta(MPIX_Prequest send) { acx::pready_cta(blockIdx.x, send); }                      // one CTA per partition
__global__ void ready_by_grid(MPIX_Prequest send) { acx::pready_grid((int)(blockIdx.x / 3), send, blockIdx.x % 3, 3); }   // three CTAs per partition, the last one to finish signals
__global__ void ready_by_warp(MPIX_Prequest send) { acx::pready_warp((int)(threadIdx.x >> 5), send); }       // one warp per partition
__global__ void poll_all(MPIX_Prequest recv) {                                                              // thread p polls partition p
    int f = 0;
    while (!f) MPIX_Parrived(recv, threadIdx.x, &f);
}

static void fill_send(int *h_send, int round) { for (int p = 0; p < PARTS; p++) for (int i = 0; i < PER; i++) h_send[p * PER + i] = value_of(t_rank, round, p, i); }
static int count_bad(const int *h_recv, int round) {
    int bad = 0;
    for (int p = 0; p < PARTS; p++) for (int i = 0; i < PER; i++) if (h_recv[p * PER + i] != value_of(T_LEFT, round, p, i)) bad++;
    return bad;
}

static void zero_copy_rounds(void) {
    int *inbox, *d_bad, h_bad = 0;
    T_OK(MPIX_Alloc_mem(PARTS * PER * sizeof(int), MPI_INFO_NULL, &inbox));
    T_CUDA(cudaMalloc((void **)&d_bad, sizeof(int)));
    T_CUDA(cudaMemset(d_bad, 0, sizeof(int)));
    MPIX_Request rq[2];
    MPIX_Prequest pq[2];
    T_OK(MPIX_Psend_init(NULL, PARTS, PER, MPI_INT, T_RIGHT, 60, MPI_COMM_WORLD, MPI_INFO_NULL, &rq[0]));   /* no send buffer: the producer owns the data */
    T_OK(MPIX_Precv_init(inbox, PARTS, PER, MPI_INT, T_LEFT, 60, MPI_COMM_WORLD, MPI_INFO_NULL, &rq[1]));
    T_OK(MPIX_Prequest_create(rq[0], &pq[0]));
    T_OK(MPIX_Prequest_create(rq[1], &pq[1]));
    for (int round = 0; round < ROUNDS; round++) {
        T_OK(MPIX_Startall(2, rq));
        T_LAUNCH(produce_zero_copy, PARTS, 64, pq[0], t_rank, round);
        T_LAUNCH(consume_blocking, PARTS, 64, pq[1], T_LEFT, round, d_bad);
        T_CUDA(cudaStreamSynchronize(0));
        T_OK(MPIX_Waitall(2, rq, MPI_STATUSES_IGNORE));
    }
    T_CUDA(cudaMemcpy(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    T_CHECK(h_bad == 0, "zero-copy producer / blocking consumer: %d wrong values", h_bad);
    T_OK(MPIX_Request_free(&rq[0])); T_OK(MPIX_Request_free(&rq[1]));
    T_OK(MPIX_Free_mem(inbox)); T_CUDA(cudaFree(d_bad));
}

static void cooperative_rounds(int use_warp, int heap) {
    static int h_send[PARTS * PER], h_recv[PARTS * PER];
    int *send, *recv;
    if (heap) { T_OK(MPIX_Alloc_mem(sizeof h_send, MPI_INFO_NULL, &send)); T_OK(MPIX_Alloc_mem(sizeof h_recv, MPI_INFO_NULL, &recv)); }
    else { T_CUDA(cudaMalloc((void **)&send, sizeof h_send)); T_CUDA(cudaMalloc((void **)&recv, sizeof h_recv)); }   /* private: landing zone + copy-out by the pollers */
    MPIX_Request rq[2];
    MPIX_Prequest pq[2];
    T_OK(MPIX_Psend_init(send, PARTS, PER, MPI_INT, T_RIGHT, 61 + use_warp, MPI_COMM_WORLD, MPI_INFO_NULL, &rq[0]));
    T_OK(MPIX_Precv_init(recv, PARTS, PER, MPI_INT, T_LEFT, 61 + use_warp, MPI_COMM_WORLD, MPI_INFO_NULL, &rq[1]));
    T_OK(MPIX_Prequest_create(rq[0], &pq[0]));
    T_OK(MPIX_Prequest_create(rq[1], &pq[1]));
    for (int round = 0; round < ROUNDS; round++) {
        fill_send(h_send, round + 10 * use_warp);
        T_CUDA(cudaMemcpy(send, h_send, sizeof h_send, cudaMemcpyHostToDevice));
        T_CUDA(cudaMemset(recv, 0xff, sizeof h_recv));
        T_OK(MPIX_Startall(2, rq));
        if (use_warp == 2) T_LAUNCH(ready_by_grid, PARTS * 3, 64, pq[0]);   /* three CTAs share each partition */
        else if (use_warp) T_LAUNCH(ready_by_warp, 1, 32 * PARTS, pq[0]); else T_LAUNCH(ready_by_cta, PARTS, 64, pq[0]);
        T_LAUNCH(poll_all, 1, PARTS, pq[1]);
        T_CUDA(cudaStreamSynchronize(0));
        T_OK(MPIX_Waitall(2, rq, MPI_STATUSES_IGNORE));
        T_CUDA(cudaMemcpy(h_recv, recv, sizeof h_recv, cudaMemcpyDeviceToHost));
        int bad = count_bad(h_recv, round + 10 * use_warp);
        T_CHECK(bad == 0, "cooperative pready (%s, %s buffers) round %d: %d wrong values", use_warp == 2 ? "grid" : use_warp ? "warp" : "cta", heap ? "heap" : "private", round, bad);
    }
    T_OK(MPIX_Request_free(&rq[0])); T_OK(MPIX_Request_free(&rq[1]));
    if (heap) { T_OK(MPIX_Free_mem(send)); T_OK(MPIX_Free_mem(recv)); } else { T_CUDA(cudaFree(send)); T_CUDA(cudaFree(recv)); }
}

// ---- 3. stream-ordered rounds: nothing returns to the host between rounds ----------------------------------------------
__global__ void stamp(int *buf, int rank, int round) { for (int j = threadIdx.x; j < PARTS * PER; j += blockDim.x) buf[j] = value_of(rank, round, j / PER, j % PER); }
__global__ void check_round(const int *buf, int left, int round, int *bad) {
    for (int j = threadIdx.x; j < PARTS * PER; j += blockDim.x) if (buf[j] != value_of(left, round, j / PER, j % PER)) atomicAdd(bad, 1);
}
static void stream_ordered_rounds(void) {
    int *send, *recv, *d_bad, h_bad = 0;
    T_CUDA(cudaMalloc((void **)&send, PARTS * PER * sizeof(int)));
    T_CUDA(cudaMalloc((void **)&recv, PARTS * PER * sizeof(int)));   /* private destination: MPIX_Pwait_enqueue's kernel copies out */
    T_CUDA(cudaMalloc((void **)&d_bad, sizeof(int)));
    T_CUDA(cudaMemset(d_bad, 0, sizeof(int)));
    cudaStream_t st = 0;
    MPIX_Request rq[2];
    MPIX_Prequest pq[2];
    T_OK(MPIX_Psend_init(send, PARTS, PER, MPI_INT, T_RIGHT, 64, MPI_COMM_WORLD, MPI_INFO_NULL, &rq[0]));
    T_OK(MPIX_Precv_init(recv, PARTS, PER, MPI_INT, T_LEFT, 64, MPI_COMM_WORLD, MPI_INFO_NULL, &rq[1]));
    T_OK(MPIX_Prequest_create(rq[0], &pq[0]));
    T_OK(MPIX_Prequest_create(rq[1], &pq[1]));
    for (int round = 0; round < ROUNDS; round++) {
        T_LAUNCH(stamp, 1, 128, send, t_rank, 100 + round);
        T_OK(MPIX_Start_enqueue(&rq[1], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Start_enqueue(&rq[0], MPIX_QUEUE_CUDA_STREAM, &st));
        T_LAUNCH(ready_by_cta, PARTS, 64, pq[0]);
        T_OK(MPIX_Pwait_enqueue(&rq[0], MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Pwait_enqueue(&rq[1], MPIX_QUEUE_CUDA_STREAM, &st));
        T_LAUNCH(check_round, 1, 128, recv, T_LEFT, 100 + round, d_bad);
    }
    T_CUDA(cudaStreamSynchronize(st));
    T_CUDA(cudaMemcpy(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    T_CHECK(h_bad == 0, "stream-ordered rounds: %d wrong values", h_bad);
    T_OK(MPIX_Request_free(&rq[0])); T_OK(MPIX_Request_free(&rq[1]));
    T_CUDA(cudaFree(send)); T_CUDA(cudaFree(recv)); T_CUDA(cudaFree(d_bad));
}

// ---- 4. explicit epochs: one kernel runs many rounds of a two-sided neighbour exchange on the device ---------------------
// Every rank exchanges with BOTH ring neighbours each round; partitions are double-buffered by round parity. Writing
// round e into a neighbour's set (e & 1) is safe once that neighbour has consumed round e-2 from it, which it did before
// it produced round e-1 — and this rank consumed the neighbour's round e-1 before producing round e. (The data dependence
// is the flow control, as in the fused halo kernel.) One CTA per partition; CTAs are independent of each other.
struct Persist { MPIX_Prequest send[2], recv[2]; int rank, nbr[2], rounds; int *bad; };
__global__ void persistent_exchange(Persist P) {
    const int p = blockIdx.x;
    for (int e = 1; e <= P.rounds; e++) {
        const int slot = (e & 1) * PARTS + p;
        for (int d = 0; d < 2; d++) {
            int *dst = (int *)acx::remote_partition_ptr_nowait(P.send[d], slot);
            for (int i = threadIdx.x; i < PER; i += blockDim.x) dst[i] = value_of(P.rank, e, p, i) + d;
        }
        __syncthreads();
        if (threadIdx.x == 0) for (int d = 0; d < 2; d++) acx::pready_signal_epoch(slot, P.send[d], (uint64_t)e);
        for (int d = 0; d < 2; d++) {
            if (threadIdx.x == 0) acx::wait_arrived_epoch(P.recv[d], slot, (uint64_t)e);
            __syncthreads();
            const int *src = (const int *)acx::local_partition_ptr(P.recv[d], slot);
            for (int i = threadIdx.x; i < PER; i += blockDim.x) if (src[i] != value_of(P.nbr[d], e, p, i) + d) atomicAdd(P.bad, 1);
        }
        __syncthreads();
    }
}
static void explicit_epoch_rounds(void) {
    /* direction 0: send right / receive from left; direction 1: send left / receive from right.
     * What arrives from the left was sent by the left rank in ITS direction 0 (value + 0); from the right, in its direction 1 (value + 1). */
    int *inbox[2], *d_bad, h_bad = 0;
    MPIX_Request sq[2], rq[2];
    Persist P;
    memset(&P, 0, sizeof P);
    T_CUDA(cudaMalloc((void **)&d_bad, sizeof(int)));
    T_CUDA(cudaMemset(d_bad, 0, sizeof(int)));
    for (int d = 0; d < 2; d++) {
        T_OK(MPIX_Alloc_mem(2 * PARTS * PER * sizeof(int), MPI_INFO_NULL, &inbox[d]));
        const int to = d == 0 ? T_RIGHT : T_LEFT, from = d == 0 ? T_LEFT : T_RIGHT;
        T_OK(MPIX_Psend_init(NULL, 2 * PARTS, PER, MPI_INT, to, 70 + d, MPI_COMM_WORLD, MPI_INFO_NULL, &sq[d]));
        T_OK(MPIX_Precv_init(inbox[d], 2 * PARTS, PER, MPI_INT, from, 70 + d, MPI_COMM_WORLD, MPI_INFO_NULL, &rq[d]));
    }
    for (int d = 0; d < 2; d++) { T_OK(MPIX_Prequest_create(sq[d], &P.send[d])); T_OK(MPIX_Prequest_create(rq[d], &P.recv[d])); }
    P.rank = t_rank; P.nbr[0] = T_LEFT; P.nbr[1] = T_RIGHT; P.rounds = 25; P.bad = d_bad;
    MPI_Barrier(MPI_COMM_WORLD);   /* every rank's landing zones exist before anybody writes */
    T_LAUNCH(persistent_exchange, PARTS, 32, P);
    T_CUDA(cudaStreamSynchronize(0));
    T_CUDA(cudaMemcpy(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    T_CHECK(h_bad == 0, "persistent kernel, 25 device-side rounds: %d wrong values", h_bad);
    MPI_Barrier(MPI_COMM_WORLD);
    for (int d = 0; d < 2; d++) { T_OK(MPIX_Request_free(&sq[d])); T_OK(MPIX_Request_free(&rq[d])); T_OK(MPIX_Free_mem(inbox[d])); }
    T_CUDA(cudaFree(d_bad));
}

// ---- 5. peer pointers: a user kernel stores into a neighbour's symmetric-heap object directly (cf. nvshmem_ptr) ----------------
__global__ void put_to_peer(int *peer_buf, int rank, int n) { for (int i = threadIdx.x; i < n; i += blockDim.x) peer_buf[i] = rank * 100000 + i; }
static void peer_pointers(void) {
    enum { N = 5000 };
    static int h[N];
    int *mine, *at_right = NULL, *again = NULL;
    size_t off = 0;
    T_OK(MPIX_Alloc_mem(N * sizeof(int), MPI_INFO_NULL, &mine));   /* same allocation sequence on every rank: same heap offset */
    T_CUDA(cudaMemset(mine, 0, N * sizeof(int)));
    MPI_Barrier(MPI_COMM_WORLD);
    T_OK(MPIX_Peer_ptr(mine, T_RIGHT, (void **)&at_right));
    T_OK(MPIX_Heap_offset(mine, &off));
    T_OK(MPIX_Peer_ptr_at(off, T_RIGHT, (void **)&again));
    T_CHECK(at_right == again && (t_size > 1 || at_right == mine), "peer pointer %p vs %p (mine %p)", (void *)at_right, (void *)again, (void *)mine);
    int dummy;
    T_CHECK(MPIX_Peer_ptr(&dummy, T_RIGHT, (void **)&again) == MPI_ERR_BUFFER, "a pointer outside the heap must be refused");
    T_LAUNCH(put_to_peer, 1, 64, at_right, t_rank, (int)N);
    T_CUDA(cudaStreamSynchronize(0));
    MPI_Barrier(MPI_COMM_WORLD);   /* every rank's stores have been performed (kernel completion + barrier) */
    T_CUDA(cudaMemcpy(h, mine, sizeof h, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < N; i++) if (h[i] != T_LEFT * 100000 + i) bad++;
    T_CHECK(bad == 0, "stores through MPIX_Peer_ptr: %d wrong values", bad);
    MPI_Barrier(MPI_COMM_WORLD);
    T_OK(MPIX_Free_mem(mine));
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 1);
    peer_pointers();
    zero_copy_rounds();
    cooperative_rounds(0, 1);
    cooperative_rounds(1, 1);
    cooperative_rounds(0, 0);
    cooperative_rounds(1, 0);
    cooperative_rounds(2, 1);
    cooperative_rounds(2, 0);
    stream_ordered_rounds();
    explicit_epoch_rounds();
    return t_finish("partitioned_device");
}
