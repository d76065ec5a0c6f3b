/* NOT an MPI implementation: a declarations-only stand-in shaped like an MPI whose handles are POINTERS to opaque
 * structs (Open MPI style), used for one thing — checking at compile time that libacx does not assume the bundled
 * mini-MPI's integer handles anywhere outside its #ifdef ACX_MINI_MPI sections (`make extmpi-check`). */
#ifndef EXTMPI_STUB_H
#define EXTMPI_STUB_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct stub_comm_s *MPI_Comm;
typedef struct stub_dt_s *MPI_Datatype;
typedef struct stub_op_s *MPI_Op;
typedef struct stub_info_s *MPI_Info;
typedef struct stub_group_s *MPI_Group;
typedef struct stub_req_s *MPI_Request;
typedef long long MPI_Count;
typedef ptrdiff_t MPI_Aint;
typedef struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR, _cancelled; size_t _ucount; } MPI_Status;
extern struct stub_comm_s stub_comm_world;
extern struct stub_dt_s stub_byte, stub_int, stub_ull;
extern struct stub_op_s stub_min;
#define MPI_COMM_WORLD (&stub_comm_world)
#define MPI_BYTE (&stub_byte)
#define MPI_INT (&stub_int)
#define MPI_UNSIGNED_LONG_LONG (&stub_ull)
#define MPI_MIN (&stub_min)
#define MPI_INFO_NULL ((MPI_Info)0)
#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_PROC_NULL (-2)
#define MPI_THREAD_MULTIPLE 3
enum { MPI_SUCCESS = 0, MPI_ERR_BUFFER = 1, MPI_ERR_COUNT = 2, MPI_ERR_TYPE = 3, MPI_ERR_TAG = 4, MPI_ERR_COMM = 5, MPI_ERR_RANK = 6,
       MPI_ERR_REQUEST = 7, MPI_ERR_ARG = 13, MPI_ERR_TRUNCATE = 15, MPI_ERR_OTHER = 16 };
int MPI_Initialized(int *), MPI_Finalized(int *), MPI_Query_thread(int *), MPI_Init_thread(int *, char ***, int, int *);
int MPI_Comm_rank(MPI_Comm, int *), MPI_Comm_size(MPI_Comm, int *), MPI_Comm_dup(MPI_Comm, MPI_Comm *), MPI_Comm_free(MPI_Comm *);
int MPI_Comm_group(MPI_Comm, MPI_Group *), MPI_Group_free(MPI_Group *), MPI_Group_translate_ranks(MPI_Group, int, const int *, MPI_Group, int *);
int MPI_Type_size(MPI_Datatype, int *), MPI_Type_get_extent(MPI_Datatype, MPI_Aint *, MPI_Aint *);
int MPI_Barrier(MPI_Comm), MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *), MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Wait(MPI_Request *, MPI_Status *), MPI_Request_free(MPI_Request *), MPI_Start(MPI_Request *);
int MPI_Psend_init(const void *, int, MPI_Count, MPI_Datatype, int, int, MPI_Comm, MPI_Info, MPI_Request *);
int MPI_Precv_init(void *, int, MPI_Count, MPI_Datatype, int, int, MPI_Comm, MPI_Info, MPI_Request *);
int MPI_Pready(int, MPI_Request), MPI_Parrived(MPI_Request, int, int *);
int MPI_Abort(MPI_Comm, int);
#ifdef __cplusplus
}
#endif
#endif
