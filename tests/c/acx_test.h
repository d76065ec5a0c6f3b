/* Shared scaffolding for the acx-b200 conformance programs (run under acxrun). */
#ifndef ACX_TEST_H
#define ACX_TEST_H
#include <mpi.h>
#include <mpi-acx.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int t_rank, t_size, t_errs, t_gpu;

#define T_CHECK(cond, ...)                                                        \
    do {                                                                          \
        if (!(cond)) {                                                            \
            printf("[%d] %s:%d CHECK(%s) failed: ", t_rank, __FILE__, __LINE__, #cond); \
            printf(__VA_ARGS__);                                                  \
            printf("\n");                                                         \
            t_errs++;                                                             \
        }                                                                         \
    } while (0)

#define T_CUDA(stmt)                                                              \
    do {                                                                          \
        cudaError_t e_ = (stmt);                                                  \
        if (e_ != cudaSuccess) {                                                  \
            printf("[%d] %s:%d %s -> %s\n", t_rank, __FILE__, __LINE__, #stmt, cudaGetErrorString(e_)); \
            MPI_Abort(MPI_COMM_WORLD, 3);                                         \
        }                                                                         \
    } while (0)

#define T_OK(stmt)                                                                \
    do {                                                                          \
        int r_ = (stmt);                                                          \
        if (r_ != 0) {                                                            \
            printf("[%d] %s:%d %s returned %d\n", t_rank, __FILE__, __LINE__, #stmt, r_); \
            MPI_Abort(MPI_COMM_WORLD, 4);                                         \
        }                                                                         \
    } while (0)

/* Init MPI + device selection (rank % ndevices, like the reference tests) + ACX. */
static void t_init(int *argc, char ***argv, int need_gpu) {
    int provided, ndev = 0;
    setvbuf(stdout, NULL, _IONBF, 0);
    MPI_Init_thread(argc, argv, MPI_THREAD_MULTIPLE, &provided);
    if (provided < MPI_THREAD_MULTIPLE) MPI_Abort(MPI_COMM_WORLD, 1);
    MPI_Comm_rank(MPI_COMM_WORLD, &t_rank);
    MPI_Comm_size(MPI_COMM_WORLD, &t_size);
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 && !getenv("ACX_FORCE_HOST")) {
        T_CUDA(cudaSetDevice(t_rank % ndev));
    }
    if (MPIX_Init()) { printf("MPIX_Init failed\n"); MPI_Abort(MPI_COMM_WORLD, 2); }
    t_gpu = MPIX_Is_device_transport();
    if (need_gpu && !t_gpu) {
        if (t_rank == 0) printf("SKIP: this test needs a GPU\n");
        MPIX_Finalize(); MPI_Finalize(); exit(77);
    }
}

static int t_finish(const char *name) {
    if (t_gpu) {   /* the library must not leave a CUDA "last error" behind (cudaErrorNotReady from an event query, a failed probe, ...): applications check it after their own launches */
        cudaError_t le = cudaGetLastError();
        T_CHECK(le == cudaSuccess, "stale CUDA last-error after the test: %s", cudaGetErrorString(le));
    }
    MPI_Allreduce(MPI_IN_PLACE, &t_errs, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
    if (MPIX_Finalize()) t_errs++;
    if (t_rank == 0) printf("%s: %s (np=%d, %s transport)\n", name, t_errs ? "FAIL" : "PASS", t_size, t_gpu ? "device" : "host");
    MPI_Finalize();
    return t_errs != 0;
}
/* kern<<<grid, block>>>(args) on the default stream; under the software device (tests/emu) kernels are plain functions */
#ifdef ACX_EMU
#define T_LAUNCH(kern, grid, block, ...) T_CUDA(emu::launch_user(kern, grid, block, (cudaStream_t)0, __VA_ARGS__))
#elif defined(__CUDACC__)
#define T_LAUNCH(kern, grid, block, ...) kern<<<grid, block>>>(__VA_ARGS__)
#endif
#define T_RIGHT ((t_rank + 1) % t_size)
#define T_LEFT ((t_rank + t_size - 1) % t_size)
#endif
