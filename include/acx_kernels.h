/* acx-b200 fused compute+communication kernels (C entry points; implemented in
 * src/kernels, shipped inside lib/libacx.so). */
#ifndef ACX_KERNELS_H
#define ACX_KERNELS_H
#include <cuda_runtime.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- halo exchange + 7-point stencil, bf16, periodic px*py*pz decomposition ------------ */
typedef struct acx_halo acx_halo;
/* Collective. Block nx*ny*nz per rank (nx 256 or 512, ny a multiple of 8, nz of 32). u_new = c0*u + c1*sum(6 nbrs).
 * The two fields live in the symmetric heap WITH their ghost shell (neighbours store face cells into it in place), so
 * the heap (ACX_HEAP_BYTES) must hold 2 x (nx+16)(ny+2)(nz+2) x 2 bytes. */
int acx_halo_create(int nx, int ny, int nz, int px, int py, int pz, float c0, float c1, acx_halo **out);
/* Device pointer (bf16, x fastest) of interior cell (0,0,0) of the current field; the field is PITCHED: cell (x,y,z) sits
 * at x + y*row_pitch + z*plane_pitch (acx_halo_field_pitch). Fill it before the first step — in place, or from / to a
 * dense [nz][ny][nx] array with acx_halo_import_dense / acx_halo_export_dense. */
void *acx_halo_field(acx_halo *h);
void acx_halo_field_pitch(acx_halo *h, long long *row_pitch_cells, long long *plane_pitch_cells);
int acx_halo_import_dense(acx_halo *h, const void *dense, cudaStream_t stream);
int acx_halo_export_dense(acx_halo *h, void *dense, cudaStream_t stream);
unsigned long long acx_halo_steps_done(acx_halo *h);
int acx_halo_exchange_initial(acx_halo *h, cudaStream_t stream);
/* Enqueue nsteps fused steps on `stream` (one kernel each; capturable in a CUDA graph). */
int acx_halo_step(acx_halo *h, int nsteps, cudaStream_t stream);
int acx_halo_destroy(acx_halo *h);

/* ---- fp8 (e4m3) GEMM, tcgen05/TMEM: Y[M][N] = scale * X[M][K] * W[N][K]^T (M%128, N%256, K%128 == 0) ---- */
int acx_gemm_fp8(const void *X, const void *W, void *Y, int M, int N, int K, float scale, cudaStream_t stream);

/* ---- pipeline-parallel fp8 GEMM chain (tcgen05/TMEM), one stage per rank ------------------- */
typedef struct acx_ppgemm acx_ppgemm;
/* Collective. Stage r computes Y = scale * X * W_r^T with X [M][D], W_r [D][D] (e4m3), and its
 * epilogue stores Y straight into stage r+1's input buffer, one release flag per row-panel
 * (`partitions` panels); stage r+1's TMA producer is gated on those flags. Double-buffered. */
int acx_ppgemm_create(int M, int D, int partitions, float scale, acx_ppgemm **out);
void *acx_ppgemm_weights(acx_ppgemm *h);            /* [D][D] e4m3, this stage */
void *acx_ppgemm_input(acx_ppgemm *h, int buf);     /* [M][D] e4m3 input buffer `buf` (stage 0: fill buf 0) */
void *acx_ppgemm_output(acx_ppgemm *h, int buf);    /* where this stage's output for buffer `buf` lands (local view; peer-mapped unless last stage) */
int acx_ppgemm_step(acx_ppgemm *h, cudaStream_t stream);   /* enqueue the next microbatch on this stage */
unsigned long long acx_ppgemm_microbatches(acx_ppgemm *h);
int acx_ppgemm_destroy(acx_ppgemm *h);

#ifdef __cplusplus
}
#endif
#endif
