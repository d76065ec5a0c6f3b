/* MPI corner semantics the reference inherits from its MPI library and a switching user may rely on:
 *  - MPI_PROC_NULL as source/destination (boundary ranks of a non-periodic halo exchange): completes at once,
 *    moves nothing, status = {MPI_PROC_NULL, MPI_ANY_TAG, 0 elements}; on a stream, on the host, and (GPU) in a graph;
 *  - duplicated communicators are separate matching contexts even when ranks first use them in different orders;
 *  - MPI_ANY_SOURCE / MPI_ANY_TAG receives (host transport; the device transport reports them as unsupported). */
#include "acx_test.h"

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    cudaStream_t stream = 0;
    MPIX_Request rq[4];
    MPI_Status st[4];
    int cnt;

    /* --- PROC_NULL mixed with a real ring exchange, stream waits --- */
    int token = 500 + t_rank, got = -1, untouched = 4242;
    memset(st, 0xff, sizeof st);
    T_OK(MPIX_Irecv_enqueue(&untouched, 1, MPI_INT, MPI_PROC_NULL, 5, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Irecv_enqueue(&got, 1, MPI_INT, T_LEFT, 5, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Isend_enqueue(&token, 1, MPI_INT, MPI_PROC_NULL, 5, MPI_COMM_WORLD, &rq[2], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Isend_enqueue(&token, 1, MPI_INT, T_RIGHT, 5, MPI_COMM_WORLD, &rq[3], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Waitall_enqueue(4, rq, st, MPIX_QUEUE_CUDA_STREAM, &stream));
    if (t_gpu) T_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i < 4; i++) T_CHECK(rq[i] == MPIX_REQUEST_NULL, "handle %d not nulled", i);
    T_CHECK(got == 500 + T_LEFT && untouched == 4242, "got %d untouched %d", got, untouched);
    T_CHECK(st[0].MPI_SOURCE == MPI_PROC_NULL && st[0].MPI_TAG == MPI_ANY_TAG && st[0].MPI_ERROR == MPI_SUCCESS, "null status %d %d %d", st[0].MPI_SOURCE, st[0].MPI_TAG, st[0].MPI_ERROR);
    MPI_Get_count(&st[0], MPI_INT, &cnt);
    T_CHECK(cnt == 0, "null recv count %d", cnt);
    T_CHECK(st[1].MPI_SOURCE == T_LEFT && st[1].MPI_TAG == 5, "real status %d %d", st[1].MPI_SOURCE, st[1].MPI_TAG);

    /* --- PROC_NULL completed by host waits --- */
    memset(st, 0xff, sizeof st);
    T_OK(MPIX_Irecv_enqueue(&untouched, 1, MPI_INT, MPI_PROC_NULL, 6, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Isend_enqueue(&token, 1, MPI_INT, MPI_PROC_NULL, 6, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Wait(&rq[0], &st[0]));
    T_OK(MPIX_Wait(&rq[1], MPI_STATUS_IGNORE));
    T_CHECK(rq[0] == MPIX_REQUEST_NULL && rq[1] == MPIX_REQUEST_NULL, "handles not nulled");
    T_CHECK(st[0].MPI_SOURCE == MPI_PROC_NULL && untouched == 4242, "host-wait null status %d", st[0].MPI_SOURCE);

    /* --- PROC_NULL in explicit-graph mode: the caller still gets graphs to embed --- */
    if (t_gpu) {
        cudaGraph_t g_op, g_wait, g;
        cudaGraphNode_t n_op, n_wait;
        cudaGraphExec_t ex;
        cudaStream_t s2;
        T_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
        T_OK(MPIX_Isend_enqueue(&token, 1, MPI_INT, MPI_PROC_NULL, 7, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_GRAPH, &g_op));
        T_OK(MPIX_Wait_enqueue(&rq[0], MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_GRAPH, &g_wait));
        T_CUDA(cudaGraphCreate(&g, 0));
        T_CUDA(cudaGraphAddChildGraphNode(&n_op, g, NULL, 0, g_op));
        T_CUDA(cudaGraphAddChildGraphNode(&n_wait, g, &n_op, 1, g_wait));
        T_CUDA(cudaGraphInstantiate(&ex, g, 0));
        T_CUDA(cudaGraphLaunch(ex, s2));
        T_CUDA(cudaGraphLaunch(ex, s2));
        T_CUDA(cudaStreamSynchronize(s2));
        T_CUDA(cudaGraphExecDestroy(ex));
        T_CUDA(cudaGraphDestroy(g)); T_CUDA(cudaGraphDestroy(g_op)); T_CUDA(cudaGraphDestroy(g_wait));
        T_CUDA(cudaStreamDestroy(s2));
    }

    /* --- two duplicated communicators, same tag, first used in opposite orders by even and odd ranks --- */
    {
        MPI_Comm ca, cb;
        MPI_Comm_dup(MPI_COMM_WORLD, &ca);
        MPI_Comm_dup(MPI_COMM_WORLD, &cb);
        int sa = 7000 + t_rank, sb = 8000 + t_rank, ra = -1, rb = -1;
        MPIX_Request q[4];
        if (t_rank % 2 == 0) {
            T_OK(MPIX_Irecv_enqueue(&ra, 1, MPI_INT, T_LEFT, 3, ca, &q[0], MPIX_QUEUE_CUDA_STREAM, &stream));
            T_OK(MPIX_Irecv_enqueue(&rb, 1, MPI_INT, T_LEFT, 3, cb, &q[1], MPIX_QUEUE_CUDA_STREAM, &stream));
            T_OK(MPIX_Isend_enqueue(&sa, 1, MPI_INT, T_RIGHT, 3, ca, &q[2], MPIX_QUEUE_CUDA_STREAM, &stream));
            T_OK(MPIX_Isend_enqueue(&sb, 1, MPI_INT, T_RIGHT, 3, cb, &q[3], MPIX_QUEUE_CUDA_STREAM, &stream));
        } else {
            T_OK(MPIX_Irecv_enqueue(&rb, 1, MPI_INT, T_LEFT, 3, cb, &q[1], MPIX_QUEUE_CUDA_STREAM, &stream));
            T_OK(MPIX_Irecv_enqueue(&ra, 1, MPI_INT, T_LEFT, 3, ca, &q[0], MPIX_QUEUE_CUDA_STREAM, &stream));
            T_OK(MPIX_Isend_enqueue(&sb, 1, MPI_INT, T_RIGHT, 3, cb, &q[3], MPIX_QUEUE_CUDA_STREAM, &stream));
            T_OK(MPIX_Isend_enqueue(&sa, 1, MPI_INT, T_RIGHT, 3, ca, &q[2], MPIX_QUEUE_CUDA_STREAM, &stream));
        }
        T_OK(MPIX_Waitall_enqueue(4, q, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
        if (t_gpu) T_CUDA(cudaStreamSynchronize(stream));
        T_CHECK(ra == 7000 + T_LEFT && rb == 8000 + T_LEFT, "communicator contexts crossed: ra=%d rb=%d", ra, rb);
        MPI_Comm_free(&ca); MPI_Comm_free(&cb);
    }

    /* --- wildcards --- */
    if (!t_gpu) {
        int v = 90 + t_rank, w = -1;
        memset(st, 0xff, sizeof st);
        T_OK(MPIX_Irecv_enqueue(&w, 1, MPI_INT, MPI_ANY_SOURCE, MPI_ANY_TAG, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Isend_enqueue(&v, 1, MPI_INT, T_RIGHT, 40 + t_rank, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Waitall_enqueue(2, rq, st, MPIX_QUEUE_CUDA_STREAM, &stream));
        T_CHECK(w == 90 + T_LEFT && st[0].MPI_SOURCE == T_LEFT && st[0].MPI_TAG == 40 + T_LEFT, "wildcard recv: %d from %d tag %d", w, st[0].MPI_SOURCE, st[0].MPI_TAG);
    } else {
        int w;
        MPIX_Request q = MPIX_REQUEST_NULL;
        int rc = MPIX_Irecv_enqueue(&w, 1, MPI_INT, MPI_ANY_SOURCE, 1, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, &stream);
        T_CHECK(rc == MPI_ERR_RANK && q == MPIX_REQUEST_NULL, "ANY_SOURCE on the device transport must be rejected cleanly, rc=%d", rc);
        rc = MPIX_Irecv_enqueue(&w, 1, MPI_INT, T_LEFT, MPI_ANY_TAG, MPI_COMM_WORLD, &q, MPIX_QUEUE_CUDA_STREAM, &stream);
        T_CHECK(rc == MPI_ERR_TAG && q == MPIX_REQUEST_NULL, "ANY_TAG on the device transport must be rejected cleanly, rc=%d", rc);
    }
    return t_finish("procnull_comms");
}
