/* MPIX_Send_enqueue / MPIX_Recv_enqueue (extension): blocking on the queue, no request handles. An eager-size receive is
 * one fused post+wait kernel; everything else is the Isend/Irecv + Wait pair underneath.
 *   - token ring on a stream (eager), statuses, many iterations (ring slots wrap);
 *   - larger payloads (still eager; rendezvous) with a deadlock-free ordering of the blocking calls;
 *   - a strided datatype, MPI_PROC_NULL, truncation into the fused receive;
 *   - stream capture relaunched past the ring depth, and explicit-graph mode. */
#include "acx_test.h"
#include <unistd.h>

static int kernels_launched(void) { MPIX_Acx_stats s; MPIX_Get_stats(&s); return (int)s.kernels_launched; }

static void token_ring(void) {
    cudaStream_t st = 0;
    int token = 0, got = -1, bad = 0;
    MPI_Status status;
    const int k0 = kernels_launched();
    for (int it = 0; it < 70; it++) {
        token = t_rank * 1000 + it;
        T_OK(MPIX_Send_enqueue(&token, 1, MPI_INT, T_RIGHT, 5, MPI_COMM_WORLD, MPIX_QUEUE_CUDA_STREAM, &st));
        T_OK(MPIX_Recv_enqueue(&got, 1, MPI_INT, T_LEFT, 5, MPI_COMM_WORLD, &status, MPIX_QUEUE_CUDA_STREAM, &st));
        if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
        int cnt = -1;
        MPI_Get_count(&status, MPI_INT, &cnt);
        if (got != T_LEFT * 1000 + it || cnt != 1 || status.MPI_SOURCE != T_LEFT || status.MPI_TAG != 5 || status.MPI_ERROR != MPI_SUCCESS) bad++;
    }
    T_CHECK(bad == 0, "token ring: %d bad iterations (last got %d)", bad, got);
    if (t_gpu && !getenv("ACX_PEER_BRANCH")) T_CHECK(kernels_launched() - k0 == 2 * 70,   /* per-peer branches keep the general path */ "an eager send + a fused receive should be 2 launches per iteration, measured %d for 70", kernels_launched() - k0);
}

static unsigned char pat(size_t i, int salt) { return (unsigned char)((i * 7u + (size_t)salt * 3u) & 0xff); }

static void payloads(void) {
    const size_t sizes[] = {0, 100, 8192, 8193, 300000};
    cudaStream_t st = 0;
    for (unsigned s = 0; s < sizeof sizes / sizeof sizes[0]; s++) {
        const size_t n = sizes[s];
        unsigned char *out, *in, *h = (unsigned char *)malloc(n + 1);
        T_OK(MPIX_Alloc_mem(n + 1, MPI_INFO_NULL, &out));
        T_OK(MPIX_Alloc_mem(n + 1, MPI_INFO_NULL, &in));
        for (size_t i = 0; i < n; i++) h[i] = pat(i, t_rank + (int)s);
        if (t_gpu) T_CUDA(cudaMemcpy(out, h, n, cudaMemcpyHostToDevice)); else memcpy(out, h, n);
        MPI_Status status;
        if (t_size == 1) {   /* a blocking send to oneself needs the receive posted first */
            MPIX_Request rq;
            T_OK(MPIX_Irecv_enqueue(in, (int)n, MPI_BYTE, 0, 6, MPI_COMM_WORLD, &rq, MPIX_QUEUE_CUDA_STREAM, &st));
            T_OK(MPIX_Send_enqueue(out, (int)n, MPI_BYTE, 0, 6, MPI_COMM_WORLD, MPIX_QUEUE_CUDA_STREAM, &st));
            T_OK(MPIX_Wait_enqueue(&rq, &status, MPIX_QUEUE_CUDA_STREAM, &st));
        } else if (t_rank % 2 == 0) {
            T_OK(MPIX_Send_enqueue(out, (int)n, MPI_BYTE, T_RIGHT, 6, MPI_COMM_WORLD, MPIX_QUEUE_CUDA_STREAM, &st));
            T_OK(MPIX_Recv_enqueue(in, (int)n, MPI_BYTE, T_LEFT, 6, MPI_COMM_WORLD, &status, MPIX_QUEUE_CUDA_STREAM, &st));
        } else {
            T_OK(MPIX_Recv_enqueue(in, (int)n, MPI_BYTE, T_LEFT, 6, MPI_COMM_WORLD, &status, MPIX_QUEUE_CUDA_STREAM, &st));
            T_OK(MPIX_Send_enqueue(out, (int)n, MPI_BYTE, T_RIGHT, 6, MPI_COMM_WORLD, MPIX_QUEUE_CUDA_STREAM, &st));
        }
        if (t_gpu) { T_CUDA(cudaStreamSynchronize(st)); T_CUDA(cudaMemcpy(h, in, n, cudaMemcpyDeviceToHost)); } else memcpy(h, in, n);
        size_t bad = 0;
        for (size_t i = 0; i < n; i++) if (h[i] != pat(i, T_LEFT + (int)s)) bad++;
        int cnt = -1;
        MPI_Get_count(&status, MPI_BYTE, &cnt);
        T_CHECK(bad == 0 && cnt == (int)n && status.MPI_SOURCE == T_LEFT, "payload %zu B: %zu wrong bytes, count %d, source %d", n, bad, cnt, status.MPI_SOURCE);
        T_OK(MPIX_Free_mem(out)); T_OK(MPIX_Free_mem(in));
        free(h);
    }
}

static void odds_and_ends(void) {
    cudaStream_t st = 0;
    MPI_Status status;
    /* strided column of a 16 x 16 matrix (takes the general path), in place */
    static int m[16][16], col[16];
    MPI_Datatype column;
    MPI_Type_vector(16, 1, 16, MPI_INT, &column);
    MPI_Type_commit(&column);
    for (int i = 0; i < 16; i++) for (int j = 0; j < 16; j++) m[i][j] = t_rank * 10000 + i * 16 + j;
    T_OK(MPIX_Send_enqueue(&m[0][3], 1, column, T_RIGHT, 7, MPI_COMM_WORLD, MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Recv_enqueue(col, 16, MPI_INT, T_LEFT, 7, MPI_COMM_WORLD, &status, MPIX_QUEUE_CUDA_STREAM, &st));
    if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
    int bad = 0;
    for (int i = 0; i < 16; i++) if (col[i] != T_LEFT * 10000 + i * 16 + 3) bad++;
    T_CHECK(bad == 0, "strided send into a contiguous fused receive: %d wrong", bad);
    MPI_Type_free(&column);
    /* MPI_PROC_NULL completes at once with the empty status */
    int v = 7;
    T_OK(MPIX_Send_enqueue(&v, 1, MPI_INT, MPI_PROC_NULL, 1, MPI_COMM_WORLD, MPIX_QUEUE_CUDA_STREAM, &st));
    T_OK(MPIX_Recv_enqueue(&v, 1, MPI_INT, MPI_PROC_NULL, 1, MPI_COMM_WORLD, &status, MPIX_QUEUE_CUDA_STREAM, &st));
    T_CHECK(status.MPI_SOURCE == MPI_PROC_NULL && v == 7, "MPI_PROC_NULL: source %d value %d", status.MPI_SOURCE, v);
    /* truncation into the fused receive: 40 ints arrive, room for 10 */
    static int big[40], small_[12];
    for (int i = 0; i < 40; i++) big[i] = t_rank + i;
    for (int i = 0; i < 12; i++) small_[i] = -5;
    T_OK(MPIX_Send_enqueue(big, 40, MPI_INT, T_RIGHT, 8, MPI_COMM_WORLD, MPIX_QUEUE_CUDA_STREAM, &st));
    int rc = MPIX_Recv_enqueue(small_, 10, MPI_INT, T_LEFT, 8, MPI_COMM_WORLD, &status, MPIX_QUEUE_CUDA_STREAM, &st);
    T_CHECK(rc == MPI_SUCCESS || rc == MPI_ERR_TRUNCATE, "truncating receive returned %d", rc);
    if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
    bad = 0;
    for (int i = 0; i < 10; i++) if (small_[i] != T_LEFT + i) bad++;
    T_CHECK(bad == 0 && small_[10] == -5 && small_[11] == -5 && status.MPI_ERROR == MPI_ERR_TRUNCATE, "truncation: %d wrong, guard %d %d, error %d", bad, small_[10], small_[11], status.MPI_ERROR);
}

static void in_graphs(void) {
    int *send, *recv;
    T_CUDA(cudaMallocHost((void **)&send, sizeof(int)));
    T_CUDA(cudaMallocHost((void **)&recv, sizeof(int)));
    cudaStream_t stream;
    T_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    for (int mode = 0; mode < 2; mode++) {
        cudaGraph_t graph;
        cudaGraphExec_t exec;
        *send = 300 + t_rank + 50 * mode; *recv = -1;
        if (mode == 0) {
            T_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeGlobal));
            T_OK(MPIX_Send_enqueue(send, 1, MPI_INT, T_RIGHT, 9, MPI_COMM_WORLD, MPIX_QUEUE_CUDA_STREAM, &stream));
            T_OK(MPIX_Recv_enqueue(recv, 1, MPI_INT, T_LEFT, 9, MPI_COMM_WORLD, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
            T_CUDA(cudaStreamEndCapture(stream, &graph));
        } else {
            cudaGraph_t gs, gr;
            cudaGraphNode_t a, b;
            T_OK(MPIX_Send_enqueue(send, 1, MPI_INT, T_RIGHT, 10, MPI_COMM_WORLD, MPIX_QUEUE_CUDA_GRAPH, &gs));
            T_OK(MPIX_Recv_enqueue(recv, 1, MPI_INT, T_LEFT, 10, MPI_COMM_WORLD, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_GRAPH, &gr));
            T_CUDA(cudaGraphCreate(&graph, 0));
            T_CUDA(cudaGraphAddChildGraphNode(&a, graph, NULL, 0, gs));
            T_CUDA(cudaGraphAddChildGraphNode(&b, graph, &a, 1, gr));
            T_CUDA(cudaGraphDestroy(gs)); T_CUDA(cudaGraphDestroy(gr));
        }
        T_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        const int hops = 40 * t_size;   /* 40 trips round the ring: every rank's token comes home, slots wrap */
        for (int hop = 0; hop < hops; hop++) {
            T_CUDA(cudaGraphLaunch(exec, stream));
            T_CUDA(cudaMemcpyAsync(send, recv, sizeof(int), cudaMemcpyHostToHost, stream));
        }
        T_CUDA(cudaStreamSynchronize(stream));
        T_CHECK(*recv == 300 + t_rank + 50 * mode, "graph mode %d: token came back as %d", mode, *recv);
        T_CUDA(cudaGraphExecDestroy(exec));
        T_CUDA(cudaGraphDestroy(graph));
    }
    T_CUDA(cudaStreamDestroy(stream));
    cudaFreeHost(send); cudaFreeHost(recv);
}

/* MPIX_Sendrecv_enqueue: a whole exchange step in one call — and, for small contiguous messages, one kernel. */
static void sendrecv_steps(void) {
    cudaStream_t st = 0;
    int bad = 0;
    static int out[2048], in[2048];   /* 8 KiB: the eager limit */
    const int k0 = kernels_launched();
    for (int it = 0; it < 70; it++) {   /* > ACX_DEPTH: ring slots wrap */
        const int n = 1 + (it % 7) * 300;   /* 4 B .. 7.2 KB */
        MPI_Status status;
        for (int i = 0; i < n; i++) { out[i] = t_rank * 100000 + it * 1000 + i; in[i] = -1; }
        T_OK(MPIX_Sendrecv_enqueue(out, n, MPI_INT, T_RIGHT, 12, in, 2048, MPI_INT, T_LEFT, 12, MPI_COMM_WORLD, &status, MPIX_QUEUE_CUDA_STREAM, &st));
        if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
        int cnt = -1;
        MPI_Get_count(&status, MPI_INT, &cnt);
        if (cnt != n || status.MPI_SOURCE != T_LEFT || status.MPI_TAG != 12) bad++;
        for (int i = 0; i < n; i++) if (in[i] != T_LEFT * 100000 + it * 1000 + i) { bad++; break; }
    }
    T_CHECK(bad == 0, "sendrecv steps: %d bad iterations", bad);
    if (t_gpu && !getenv("ACX_PEER_BRANCH")) T_CHECK(kernels_launched() - k0 == 70, "an eager exchange step should be ONE launch, measured %d for 70", kernels_launched() - k0);
    /* large halves (rendezvous): the general path underneath; both directions at once cannot deadlock (receive posted first) */
    {
        const size_t n = 150000;
        unsigned char *o, *i_, *h = (unsigned char *)malloc(n);
        T_OK(MPIX_Alloc_mem(n, MPI_INFO_NULL, &o)); T_OK(MPIX_Alloc_mem(n, MPI_INFO_NULL, &i_));
        for (size_t i = 0; i < n; i++) h[i] = pat(i, t_rank + 9);
        if (t_gpu) T_CUDA(cudaMemcpy(o, h, n, cudaMemcpyHostToDevice)); else memcpy(o, h, n);
        T_OK(MPIX_Sendrecv_enqueue(o, (int)n, MPI_BYTE, T_RIGHT, 13, i_, (int)n, MPI_BYTE, T_LEFT, 13, MPI_COMM_WORLD, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &st));
        if (t_gpu) { T_CUDA(cudaStreamSynchronize(st)); T_CUDA(cudaMemcpy(h, i_, n, cudaMemcpyDeviceToHost)); } else memcpy(h, i_, n);
        size_t wrong = 0;
        for (size_t i = 0; i < n; i++) if (h[i] != pat(i, T_LEFT + 9)) wrong++;
        T_CHECK(wrong == 0, "large sendrecv: %zu wrong bytes", wrong);
        T_OK(MPIX_Free_mem(o)); T_OK(MPIX_Free_mem(i_)); free(h);
    }
    /* one half MPI_PROC_NULL (an open chain instead of a ring): rank 0 receives from nobody, the last rank sends to nobody */
    {
        int v = 1000 + t_rank, w = -1;
        MPI_Status status;
        const int to = t_rank + 1 < t_size ? t_rank + 1 : MPI_PROC_NULL, from = t_rank > 0 ? t_rank - 1 : MPI_PROC_NULL;
        T_OK(MPIX_Sendrecv_enqueue(&v, 1, MPI_INT, to, 14, &w, 1, MPI_INT, from, 14, MPI_COMM_WORLD, &status, MPIX_QUEUE_CUDA_STREAM, &st));
        if (t_gpu) T_CUDA(cudaStreamSynchronize(st));
        T_CHECK(t_rank == 0 ? (w == -1 && status.MPI_SOURCE == MPI_PROC_NULL) : (w == 1000 + t_rank - 1 && status.MPI_SOURCE == t_rank - 1), "chain with MPI_PROC_NULL ends: got %d from %d", w, status.MPI_SOURCE);
    }
    if (!t_gpu) return;
    /* in a captured graph, relaunched 40 trips round the ring; and as an explicit graph */
    int *send, *recv;
    T_CUDA(cudaMallocHost((void **)&send, sizeof(int)));
    T_CUDA(cudaMallocHost((void **)&recv, sizeof(int)));
    cudaStream_t stream;
    T_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    for (int mode = 0; mode < 2; mode++) {
        cudaGraph_t graph;
        cudaGraphExec_t exec;
        *send = 800 + t_rank + 50 * mode; *recv = -1;
        if (mode == 0) {
            T_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeGlobal));
            T_OK(MPIX_Sendrecv_enqueue(send, 1, MPI_INT, T_RIGHT, 15, recv, 1, MPI_INT, T_LEFT, 15, MPI_COMM_WORLD, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
            T_CUDA(cudaStreamEndCapture(stream, &graph));
        } else {
            T_OK(MPIX_Sendrecv_enqueue(send, 1, MPI_INT, T_RIGHT, 16, recv, 1, MPI_INT, T_LEFT, 16, MPI_COMM_WORLD, MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_GRAPH, &graph));
        }
        T_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        for (int hop = 0; hop < 40 * t_size; hop++) {
            T_CUDA(cudaGraphLaunch(exec, stream));
            T_CUDA(cudaMemcpyAsync(send, recv, sizeof(int), cudaMemcpyHostToHost, stream));
        }
        T_CUDA(cudaStreamSynchronize(stream));
        T_CHECK(*recv == 800 + t_rank + 50 * mode, "sendrecv in a graph (mode %d): token came back as %d", mode, *recv);
        T_CUDA(cudaGraphExecDestroy(exec));
        T_CUDA(cudaGraphDestroy(graph));
    }
    T_CUDA(cudaStreamDestroy(stream));
    cudaFreeHost(send); cudaFreeHost(recv);
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    token_ring();
    sendrecv_steps();
    payloads();
    odds_and_ends();
    if (t_gpu) in_graphs();
    MPIX_Acx_stats s;
    MPIX_Get_stats(&s);
    for (int i = 0; i < 5000 && t_gpu && s.slots_in_use != 0; i++) { usleep(1000); MPIX_Get_stats(&s); }   /* graph user-object destructors may run after cudaGraphDestroy returns */
    T_CHECK(!t_gpu || s.slots_in_use == 0, "%llu request slots still in use", s.slots_in_use);
    return t_finish("sendrecv_enqueue");
}
