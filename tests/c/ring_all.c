/* MPIX_Waitall_enqueue with MPI_STATUSES_IGNORE and with a status array (reference
 * test/src/ring-all.c covers the former), several rounds back to back on one stream. */
#include "acx_test.h"

int main(int argc, char **argv) {
    t_init(&argc, &argv, 0);
    cudaStream_t stream = 0;
    for (int round = 0; round < 5; round++) {
        double out[3] = {t_rank + 0.5, round, 42.0}, in[3] = {-1, -1, -1};
        MPIX_Request rq[2];
        MPI_Status sts[2];
        T_OK(MPIX_Isend_enqueue(out, 3, MPI_DOUBLE, T_RIGHT, round, MPI_COMM_WORLD, &rq[0], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Irecv_enqueue(in, 3, MPI_DOUBLE, T_LEFT, round, MPI_COMM_WORLD, &rq[1], MPIX_QUEUE_CUDA_STREAM, &stream));
        T_OK(MPIX_Waitall_enqueue(2, rq, (round & 1) ? sts : MPI_STATUSES_IGNORE, MPIX_QUEUE_CUDA_STREAM, &stream));
        if (t_gpu) T_CUDA(cudaStreamSynchronize(stream));
        T_CHECK(in[0] == T_LEFT + 0.5 && in[1] == round && in[2] == 42.0, "round %d got %f %f %f", round, in[0], in[1], in[2]);
        if (round & 1) {
            int cnt = -1;
            MPI_Get_count(&sts[1], MPI_DOUBLE, &cnt);
            T_CHECK(sts[1].MPI_SOURCE == T_LEFT && sts[1].MPI_TAG == round && sts[1].MPI_ERROR == MPI_SUCCESS && cnt == 3, "status src %d tag %d cnt %d",
                    sts[1].MPI_SOURCE, sts[1].MPI_TAG, cnt);
        }
    }
    return t_finish("ring_all");
}
