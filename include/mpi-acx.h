/* acx-b200 public API.
 *
 * Same entry points and signatures as NVIDIA/mpi-acx (reference include/mpi-acx.h:42-104):
 * stream/graph-triggered MPIX_Isend/Irecv/Wait/Waitall_enqueue and kernel-triggered
 * partitioned MPIX_Psend_init/Precv_init/Start/Pready/Parrived — but implemented
 * GPU-natively: payload and completion flags move GPU->GPU over NVLink from inside
 * CUDA kernels; there is no CPU proxy thread and no MPI call on any data path.
 *
 * Build: include <mpi.h> (a real MPI, or the bundled mini-MPI in include/mpi.h) first.
 * Device functions are header-inline (no relocatable-device-code link step needed,
 * unlike the reference, Makefile:8).
 */
#ifndef MPI_ACX_H
#define MPI_ACX_H

#include <mpi.h>
#include <cuda.h>
#include <cuda_runtime.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPIX_ACX_B200 1
#define MPIX_ACX_VERSION_MAJOR 0
#define MPIX_ACX_VERSION_MINOR 1

/* Opaque handles. MPIX_Prequest is a device pointer; kernels receive it by value. */
typedef void *MPIX_Request;
typedef void *MPIX_Prequest;
#define MPIX_REQUEST_NULL  NULL
#define MPIX_PREQUEST_NULL NULL

/* Kind of `queue_handle` in the *_enqueue calls. */
enum {
    MPIX_QUEUE_CUDA_STREAM,  /* cudaStream_t*  (in; the stream may be capturing)        */
    MPIX_QUEUE_CUDA_GRAPH    /* cudaGraph_t*   (out: a new graph that contains the op)  */
};

/* ---- lifecycle (after MPI_Init_thread and cudaSetDevice; collective over MPI_COMM_WORLD) ---- */
int MPIX_Init(void);
int MPIX_Finalize(void);

/* ---- stream / graph triggered two-sided operations -------------------------------------------
 * The transfer starts no earlier than the point where the call sits in the stream or graph; the
 * buffer must stay valid until a wait on the returned request has completed. MPI_PROC_NULL peers are
 * accepted (the operation completes at its wait, moving nothing); MPI_ANY_SOURCE / MPI_ANY_TAG receives
 * are honoured on the host transport and rejected (MPI_ERR_RANK / MPI_ERR_TAG) on the device transport.
 * `comm` may be any communicator made of MPI_COMM_WORLD members (dup, split, Cartesian); ranks and
 * MPI_Status.MPI_SOURCE are in that communicator's rank space. */
int MPIX_Isend_enqueue(const void *sendbuf, int nelems, MPI_Datatype dtype, int dst_rank, int tag, MPI_Comm comm, MPIX_Request *req_out, int queue_kind, void *queue_handle);
int MPIX_Irecv_enqueue(void *recvbuf, int nelems, MPI_Datatype dtype, int src_rank, int tag, MPI_Comm comm, MPIX_Request *req_out, int queue_kind, void *queue_handle);

/* Work queued behind these runs only after the operation(s) completed; the handles are consumed
 * (set to MPIX_REQUEST_NULL) at enqueue time. Not for partitioned requests. */
int MPIX_Wait_enqueue(MPIX_Request *req, MPI_Status *status, int queue_kind, void *queue_handle);
int MPIX_Waitall_enqueue(int nreq, MPIX_Request *reqs, MPI_Status *statuses, int queue_kind, void *queue_handle);

/* Host-blocking completion: basic requests that were enqueued on a stream (frees them), or one round
 * of a partitioned request (which stays allocated, inactive, for the next MPIX_Start). */
int MPIX_Wait(MPIX_Request *req, MPI_Status *status);
int MPIX_Waitall(int nreq, MPIX_Request *reqs, MPI_Status *statuses);

/* ---- persistent partitioned operations (MPI-4 semantics: `nelems` is PER PARTITION) ----------- */
int MPIX_Psend_init(const void *sendbuf, int partitions, MPI_Count nelems, MPI_Datatype dtype, int dst_rank, int tag, MPI_Comm comm, MPI_Info info, MPIX_Request *req_out);
int MPIX_Precv_init(void *recvbuf, int partitions, MPI_Count nelems, MPI_Datatype dtype, int src_rank, int tag, MPI_Comm comm, MPI_Info info, MPIX_Request *req_out);
int MPIX_Start(MPIX_Request *req);
int MPIX_Startall(int nreq, MPIX_Request *reqs);
int MPIX_Request_free(MPIX_Request *req);

/* Device-side handle of a partitioned request (several may exist per request). */
int MPIX_Prequest_create(MPIX_Request req, MPIX_Prequest *preq_out);
int MPIX_Prequest_free(MPIX_Prequest *preq);

/* EXTENSIONS (not in the reference) ******************************************/

/* Symmetric communication heap: device memory that every peer GPU has mapped at init.
 * Buffers allocated here are sent/received zero-copy (the sending kernel stores
 * straight into the peer's HBM). Any other buffer (cudaMalloc, host) still works
 * but is staged through a bounce buffer in this heap. */
int MPIX_Alloc_mem(size_t bytes, MPI_Info info, void *baseptr /* void** */);
int MPIX_Free_mem(void *base);

/* Address of a symmetric-heap object as mapped from `peer`'s copy of the heap, i.e. the pointer
 * this GPU dereferences to load/store the peer's instance directly over NVLink (cf. nvshmem_ptr).
 * `local` must come from MPIX_Alloc_mem at the same allocation sequence point on both ranks. */
int MPIX_Peer_ptr(const void *local, int peer, void **peer_ptr);
/* Pageable host buffers are pinned in place on first use and the registration is cached (and re-validated against the page
 * frames behind it where /proc/self/pagemap shows them). Call this before freeing a host buffer the library has seen if you
 * want its pages unpinned right away; it is never required for correctness where the validation works. */
int MPIX_Buffer_forget(const void *buf, size_t bytes);
/* Asymmetric allocations: exchange MPIX_Heap_offset() values, then map the peer's offset. */
int MPIX_Heap_offset(const void *local, size_t *offset);
int MPIX_Peer_ptr_at(size_t offset, int peer, void **peer_ptr);

/* Host-side versions of the __host__ __device__ pair below (usable from C). */
int MPIX_Pready_host(int partition, MPIX_Request *request);
int MPIX_Parrived_host(MPIX_Request *request, int partition, int *flag);

/* Blocking-on-the-queue send and receive: work enqueued after the call runs once the operation is complete (= the Isend /
 * Irecv + Wait pair, without a request handle). An eager-size MPIX_Recv_enqueue on a stream is ONE kernel (post + wait
 * fused) and an eager MPIX_Send_enqueue is just its trigger, so a small-message exchange costs two launches instead of three. */
int MPIX_Send_enqueue(const void *sendbuf, int nelems, MPI_Datatype dtype, int dst_rank, int tag, MPI_Comm comm, int queue_kind, void *queue_handle);
int MPIX_Recv_enqueue(void *recvbuf, int nelems, MPI_Datatype dtype, int src_rank, int tag, MPI_Comm comm, MPI_Status *status, int queue_kind, void *queue_handle);

/* MPI_Sendrecv on the queue: later work runs once the receive has completed and the send buffer may be reused. With both
 * halves contiguous and of eager size the whole exchange step is ONE kernel (post the receive, deposit the send, complete
 * the receive); otherwise it is the Irecv + Isend + Waitall triple. */
int MPIX_Sendrecv_enqueue(const void *sendbuf, int sendcount, MPI_Datatype sendtype, int dst_rank, int sendtag, void *recvbuf, int recvcount, MPI_Datatype recvtype,
                          int src_rank, int recvtag, MPI_Comm comm, MPI_Status *status, int queue_kind, void *queue_handle);

/* Stream-ordered partitioned round control (the reference only has host Start/Wait):
 * enqueue the Start / the completion wait of a partitioned request on a stream. */
int MPIX_Start_enqueue(MPIX_Request *request, int qtype, void *queue);
int MPIX_Pwait_enqueue(MPIX_Request *request, int qtype, void *queue);

/* Runtime introspection / observability. */
typedef struct MPIX_Acx_stats {
    unsigned long long sends_enqueued, recvs_enqueued, waits_enqueued;
    unsigned long long bytes_sent, bytes_received;
    unsigned long long eager_msgs, rendezvous_msgs, bounce_copies;
    unsigned long long kernels_launched, graph_nodes_added;
    unsigned long long slots_in_use, slots_high_water;
    unsigned long long heap_bytes_in_use, heap_bytes_high_water;
    unsigned long long copy_engine_msgs;   /* rendezvous messages moved by cudaMemcpyAsync (ACX_CE_MIN) */
    unsigned long long overflow_sends;     /* sends whose trigger waited for its ring slot on a pool stream (more than ACX_DEPTH un-waited sends on one channel) */
} MPIX_Acx_stats;
int MPIX_Get_stats(MPIX_Acx_stats *stats);
/* 1 when the GPU transport is active, 0 when running on the host (no-GPU) transport. */
int MPIX_Is_device_transport(void);
/* Dump the device-side trace ring (ACX_TRACE=1) as JSON lines to `path`. */
int MPIX_Trace_dump(const char *path);

#ifdef __cplusplus
}
#endif

/* DEVICE FUNCTIONS FOR PARTITIONED OPERATIONS ********************************/

#ifdef __CUDACC__
#include "acx_device.cuh"
/* Request argument is void* so host and device share one signature (as in the
 * reference, include/mpi-acx.h:96-103):
 *   __host__   -- MPIX_Request *request
 *   __device__ -- MPIX_Prequest  request (the handle value itself)
 */
static __host__ __device__ __forceinline__ int MPIX_Pready(int partition, void *request) {
#ifdef __CUDA_ARCH__
    return acx::pready(partition, request);
#else
    return MPIX_Pready_host(partition, (MPIX_Request *)request);
#endif
}
static __host__ __device__ __forceinline__ int MPIX_Parrived(void *request, int partition, int *flag) {
#ifdef __CUDA_ARCH__
    return acx::parrived(request, partition, flag);
#else
    return MPIX_Parrived_host((MPIX_Request *)request, partition, flag);
#endif
}
#elif defined(ACX_EMU)
/* Software device (tests/emu): kernels are plain functions run on fibers; one body serves host and "device" callers. */
#include "acx_device.cuh"
static inline int MPIX_Pready(int partition, void *request) {
    return emu::cta ? acx::pready(partition, request) : MPIX_Pready_host(partition, (MPIX_Request *)request);
}
static inline int MPIX_Parrived(void *request, int partition, int *flag) {
    if (!emu::cta) return MPIX_Parrived_host((MPIX_Request *)request, partition, flag);
    const int rc = acx::parrived(request, partition, flag);
    if (!*flag) emu::yield();   /* a polling loop must let the CTA's other fibers run */
    return rc;
}
#endif /* __CUDACC__ */

#endif
