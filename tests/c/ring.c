/* On-stream send/recv around a ring with host (stack) buffers.
 * Covers the behaviours of the reference's test/src/ring.c: two MPIX_Wait_enqueue with a
 * real MPI_Status, then the same exchange completed by host-side MPIX_Wait. Extra: a
 * third round through a user-created non-blocking stream. */
#include "acx_test.h"

static void check_status(MPI_Status *st, int tag) {
    T_CHECK(st->MPI_SOURCE == T_LEFT, "source %d", st->MPI_SOURCE);
    T_CHECK(st->MPI_TAG == tag, "tag %d", st->MPI_TAG);
    T_CHECK(st->MPI_ERROR == MPI_SUCCESS, "error %d", st->MPI_ERROR);
    int cnt = -1;
    MPI_Get_count(st, MPI_INT, &cnt);
    T_CHECK(cnt == 1, "count %d", cnt);
}

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    int token = 1000 + t_rank, got = -1;
    MPIX_Request rq[2];
    MPI_Status st;
    cudaStream_t stream = 0;

    /* round 1: waits enqueued on the stream */
    memset(&st, 0xff, sizeof st);
    T_OK(MPIX_Isend_enqueue(&token, 1, MPI_INT, T_RIGHT, 0, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Irecv_enqueue(&got, 1, MPI_INT, T_LEFT, 0, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Wait_enqueue(&rq[0], MPI_STATUS_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Wait_enqueue(&rq[1], &st, MPIX_QUEUE_CUDA_STREAM, &stream));
    T_CHECK(rq[0] == MPIX_REQUEST_NULL && rq[1] == MPIX_REQUEST_NULL, "handles not nulled");
    if (t_gpu) T_CUDA(cudaStreamSynchronize(stream));
    T_CHECK(got == 1000 + T_LEFT, "got %d", got);
    check_status(&st, 0);

    /* round 2: host waits */
    got = -1;
    memset(&st, 0xff, sizeof st);
    token += 7;
    T_OK(MPIX_Isend_enqueue(&token, 1, MPI_INT, T_RIGHT, 1, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Irecv_enqueue(&got, 1, MPI_INT, T_LEFT, 1, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
    T_OK(MPIX_Wait(&rq[0], MPI_STATUS_IGNORE));
    T_OK(MPIX_Wait(&rq[1], &st));
    T_CHECK(rq[1] == MPIX_REQUEST_NULL, "handle not nulled");
    T_CHECK(got == 1007 + T_LEFT, "got %d", got);
    check_status(&st, 1);

    /* round 3: user stream, recv posted before send */
    if (t_gpu) {
        cudaStream_t s2;
        T_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
        got = -1;
        token += 7;
        T_OK(MPIX_Irecv_enqueue(&got, 1, MPI_INT, T_LEFT, 2, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &s2));
        T_OK(MPIX_Isend_enqueue(&token, 1, MPI_INT, T_RIGHT, 2, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &s2));
        T_OK(MPIX_Waitall_enqueue(2, rq, MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &s2));
        T_CUDA(cudaStreamSynchronize(s2));
        T_CHECK(got == 1014 + T_LEFT, "got %d", got);
        T_CUDA(cudaStreamDestroy(s2));
    }
    return t_finish("ring");
}
