/* acx mini-MPI: a small single-node MPI subset shipped with acx-b200.
 *
 * Why it exists: the build/GPU image has no MPI at all (no mpi.h, no mpiexec), and
 * the public ACX API is expressed in MPI types (MPI_Comm, MPI_Datatype, MPI_Status).
 * This header + libacxmpi give programs written against <mpi.h> + <mpi-acx.h>
 * (e.g. the six reference tests, the .c files under /root/reference/test/src) everything they use:
 * init/finalize, rank/size, barrier/bcast/allreduce/allgather, two-sided
 * nonblocking p2p with tag matching, and the MPI-4 partitioned API.
 *
 * It is ALSO the "MPI library" underneath the unmodified reference build used as
 * the measured baseline (baseline/): host buffers move through POSIX shm rings,
 * device buffers through CUDA IPC + cudaMemcpyAsync, i.e. the intra-node data path
 * of a CUDA-aware MPI. acx-b200's own GPU data path never calls into it; it only
 * uses the bootstrap (rank/size/shm segment) and, with no GPU present, the host
 * transport (BASELINE.json config #1: "host buffers, MPI only, no GPU").
 *
 * Launch: `acxrun -np N prog` (launcher/acxrun.cpp) or torchrun (RANK/WORLD_SIZE).
 */
#ifndef ACX_MINI_MPI_H
#define ACX_MINI_MPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACX_MINI_MPI 1
#define MPI_VERSION 4
#define MPI_SUBVERSION 0

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef int MPI_Errhandler;
typedef int MPI_Group;
typedef long long MPI_Count;
typedef long long MPI_Aint;
typedef struct acx_mpi_request_s *MPI_Request;

typedef struct MPI_Status {
    int MPI_SOURCE;
    int MPI_TAG;
    int MPI_ERROR;
    long long _bytes; /* received payload size, for MPI_Get_count */
} MPI_Status;

#define MPI_COMM_NULL  ((MPI_Comm)0)
#define MPI_COMM_WORLD ((MPI_Comm)1)
#define MPI_COMM_SELF  ((MPI_Comm)2)
#define MPI_GROUP_NULL  ((MPI_Group)0)
#define MPI_GROUP_EMPTY ((MPI_Group)1)

/* datatype handle = 0x100 | log-free encoding: low byte is the element size */
#define ACX_DT(id, size) ((MPI_Datatype)(((id) << 8) | (size)))
#define MPI_DATATYPE_NULL ((MPI_Datatype)0)
#define MPI_CHAR          ACX_DT(1, 1)
#define MPI_SIGNED_CHAR   ACX_DT(2, 1)
#define MPI_UNSIGNED_CHAR ACX_DT(3, 1)
#define MPI_BYTE          ACX_DT(4, 1)
#define MPI_SHORT         ACX_DT(5, 2)
#define MPI_UNSIGNED_SHORT ACX_DT(6, 2)
#define MPI_INT           ACX_DT(7, 4)
#define MPI_UNSIGNED      ACX_DT(8, 4)
#define MPI_LONG          ACX_DT(9, 8)
#define MPI_UNSIGNED_LONG ACX_DT(10, 8)
#define MPI_LONG_LONG     ACX_DT(11, 8)
#define MPI_LONG_LONG_INT MPI_LONG_LONG
#define MPI_UNSIGNED_LONG_LONG ACX_DT(12, 8)
#define MPI_FLOAT         ACX_DT(13, 4)
#define MPI_DOUBLE        ACX_DT(14, 8)
#define MPI_INT8_T        ACX_DT(15, 1)
#define MPI_UINT8_T       ACX_DT(16, 1)
#define MPI_INT16_T       ACX_DT(17, 2)
#define MPI_UINT16_T      ACX_DT(18, 2)
#define MPI_INT32_T       ACX_DT(19, 4)
#define MPI_UINT32_T      ACX_DT(20, 4)
#define MPI_INT64_T       ACX_DT(21, 8)
#define MPI_UINT64_T      ACX_DT(22, 8)
#define MPI_C_BOOL        ACX_DT(23, 1)
#define MPI_COUNT         ACX_DT(24, 8)
#define MPI_AINT          ACX_DT(25, 8)
/* acx extensions: 16-bit / 8-bit float payload types (reductions unsupported) */
#define MPIX_BFLOAT16     ACX_DT(26, 2)
#define MPIX_FLOAT16      ACX_DT(27, 2)
#define MPIX_FLOAT8_E4M3  ACX_DT(28, 1)

/* derived datatypes: handles >= ACX_DT_DERIVED_BASE index a per-process table */
#define ACX_DT_DERIVED_BASE 0x100000
#define MPI_ORDER_C       56
#define MPI_ORDER_FORTRAN 57

#define MPI_OP_NULL ((MPI_Op)0)
#define MPI_SUM  ((MPI_Op)1)
#define MPI_MAX  ((MPI_Op)2)
#define MPI_MIN  ((MPI_Op)3)
#define MPI_PROD ((MPI_Op)4)
#define MPI_LAND ((MPI_Op)5)
#define MPI_LOR  ((MPI_Op)6)
#define MPI_BAND ((MPI_Op)7)
#define MPI_BOR  ((MPI_Op)8)

#define MPI_INFO_NULL ((MPI_Info)0)
#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_STATUS_IGNORE   ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_IN_PLACE ((void *)-1)
#define MPI_ANY_SOURCE (-2)
#define MPI_ANY_TAG    (-1)
#define MPI_PROC_NULL  (-3)
#define MPI_UNDEFINED  (-32766)
#define MPI_TAG_UB_VALUE 0x3fffffff

#define MPI_THREAD_SINGLE     0
#define MPI_THREAD_FUNNELED   1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE   3

#define MPI_SUCCESS      0
#define MPI_ERR_BUFFER   1
#define MPI_ERR_COUNT    2
#define MPI_ERR_TYPE     3
#define MPI_ERR_TAG      4
#define MPI_ERR_COMM     5
#define MPI_ERR_RANK     6
#define MPI_ERR_REQUEST  7
#define MPI_ERR_TRUNCATE 15
#define MPI_ERR_OTHER    16
#define MPI_ERR_INTERN   17
#define MPI_ERR_ARG      13
#define MPI_ERR_LASTCODE 64
#define MPI_MAX_ERROR_STRING 256
#define MPI_MAX_PROCESSOR_NAME 256

#define MPI_ERRORS_ARE_FATAL ((MPI_Errhandler)1)
#define MPI_ERRORS_RETURN    ((MPI_Errhandler)2)

int MPI_Init(int *argc, char ***argv);
int MPI_Init_thread(int *argc, char ***argv, int required, int *provided);
int MPI_Finalize(void);
int MPI_Initialized(int *flag);
int MPI_Finalized(int *flag);
int MPI_Query_thread(int *provided);
int MPI_Abort(MPI_Comm comm, int errorcode);
int MPI_Error_string(int errorcode, char *string, int *resultlen);
int MPI_Comm_set_errhandler(MPI_Comm comm, MPI_Errhandler eh);
double MPI_Wtime(void);
int MPI_Get_processor_name(char *name, int *resultlen);

int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm);   /* color MPI_UNDEFINED -> MPI_COMM_NULL */
int MPI_Comm_group(MPI_Comm comm, MPI_Group *group);
int MPI_Group_size(MPI_Group group, int *size);
int MPI_Group_rank(MPI_Group group, int *rank);
int MPI_Group_translate_ranks(MPI_Group group1, int n, const int ranks1[], MPI_Group group2, int ranks2[]);
int MPI_Group_free(MPI_Group *group);
/* Cartesian topologies (reorder is accepted and ignored: ranks keep their order) */
int MPI_Dims_create(int nnodes, int ndims, int dims[]);
int MPI_Cart_create(MPI_Comm comm_old, int ndims, const int dims[], const int periods[], int reorder, MPI_Comm *comm_cart);
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int coords[]);
int MPI_Cart_rank(MPI_Comm comm, const int coords[], int *rank);
int MPI_Cart_shift(MPI_Comm comm, int direction, int disp, int *rank_source, int *rank_dest);
int MPI_Type_size(MPI_Datatype datatype, int *size);
int MPI_Get_count(const MPI_Status *status, MPI_Datatype datatype, int *count);
/* Derived datatypes (the regular strided layouts halo exchanges use). Any of these may be nested as long as the result
 * is expressible as <= 3 strided dimensions around one contiguous run. */
int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_create_hvector(int count, int blocklength, MPI_Aint stride_bytes, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_create_subarray(int ndims, const int sizes[], const int subsizes[], const int starts[], int order,
                             MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *datatype);
int MPI_Type_free(MPI_Datatype *datatype);
int MPI_Type_get_extent(MPI_Datatype datatype, MPI_Aint *lb, MPI_Aint *extent);
int MPI_Pack_size(int incount, MPI_Datatype datatype, MPI_Comm comm, int *size);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sdt, void *recvbuf, int recvcount,
                  MPI_Datatype rdt, MPI_Comm comm);

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Status *status);
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Sendrecv(const void *sendbuf, int sendcount, MPI_Datatype sdt, int dest, int sendtag, void *recvbuf,
                 int recvcount, MPI_Datatype rdt, int source, int recvtag, MPI_Comm comm, MPI_Status *status);
int MPI_Test(MPI_Request *req, int *flag, MPI_Status *status);
int MPI_Wait(MPI_Request *req, MPI_Status *status);
int MPI_Waitall(int count, MPI_Request *reqs, MPI_Status *statuses);
int MPI_Testall(int count, MPI_Request *reqs, int *flag, MPI_Status *statuses);
int MPI_Request_free(MPI_Request *req);

/* MPI-4 partitioned communication */
int MPI_Psend_init(const void *buf, int partitions, MPI_Count count, MPI_Datatype dt, int dest, int tag,
                   MPI_Comm comm, MPI_Info info, MPI_Request *req);
int MPI_Precv_init(void *buf, int partitions, MPI_Count count, MPI_Datatype dt, int source, int tag,
                   MPI_Comm comm, MPI_Info info, MPI_Request *req);
int MPI_Start(MPI_Request *req);
int MPI_Startall(int count, MPI_Request *reqs);
int MPI_Pready(int partition, MPI_Request req);
int MPI_Parrived(MPI_Request req, int partition, int *flag);

/* ---- acx bootstrap extensions (not MPI): used by libacx and the launcher ---- */
int  ACX_Mpi_local_rank(void);           /* rank within the node (== world rank here) */
const char *ACX_Mpi_jobid(void);         /* unique id of this job's shm segment */
/* Symmetric exchange of a fixed-size blob among all ranks of COMM_WORLD (allgather
 * over the shm control block; usable before any CUDA state exists). */
int  ACX_Mpi_exchange(const void *mine, void *all, size_t bytes_per_rank);
/* Raw pointer + size of a per-job shared scratch area libacx may carve up
 * (channel directories, registration tables). Zero-initialised at job start. */
void *ACX_Mpi_shared_scratch(size_t *bytes);
/* Memory layout of a datatype in canonical form, for layers that move the bytes themselves (libacx's pack/unpack
 * kernels): element (i2,i1,i0) of one item starts at  disp + i2*stride[2] + i1*stride[1] + i0*stride[0]  and is `run`
 * contiguous bytes; consecutive items of a count>1 transfer are `extent` bytes apart. Unused dimensions have n = 1. */
typedef struct ACX_Mpi_layout {
    long long n[3], stride[3];
    long long run, disp, extent, size;   /* size: packed bytes of one item = n[0]*n[1]*n[2]*run */
    int contiguous;                      /* 1: one item is a single run of `size` bytes starting at `disp` and extent == size */
} ACX_Mpi_layout;
int ACX_Mpi_type_layout(MPI_Datatype datatype, ACX_Mpi_layout *layout);
/* Called once at the start of MPI_Abort (diagnostics of layered libraries). */
void ACX_Mpi_set_abort_hook(void (*fn)(void));

#ifdef __cplusplus
}
#endif
#endif /* ACX_MINI_MPI_H */
