/* oracle/mpi_mp/mpi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A small MULTI-PROCESS MPI replacement, so that the unmodified reference SLATE sources under /root/reference can run on a
 * p x q process grid in an image that has no MPI: ranks are processes started by oracle/mprun.py, every pair of ranks is
 * connected by a Unix socket pair, a progress thread per rank drains its sockets into a matching queue (so sends never
 * block on the receiver), collectives are linear algorithms over point-to-point messages with a fixed combination order.
 * Only the MPI names the reference uses are provided (the list of oracle/mpi_stub/mpi.h, the one-rank stub).
 * MPI_THREAD_MULTIPLE: every entry point may be called from several OpenMP task threads at once, as the reference does.
 * The implementation is oracle/mpi_mp/mpi_mp.cc.
 */
#ifndef SB200_ORACLE_MPI_MP_H
#define SB200_ORACLE_MPI_MP_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int  MPI_Comm;
typedef int  MPI_Group;
typedef int  MPI_Request;
typedef int  MPI_Op;
typedef long MPI_Datatype;
typedef long MPI_Aint;
typedef struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; } MPI_Status;
typedef void (MPI_User_function)(void* in, void* inout, int* len, MPI_Datatype* type);

enum { MPI_COMM_NULL = 0, MPI_COMM_WORLD = 1, MPI_COMM_SELF = 2 };
enum { MPI_GROUP_NULL = 0, MPI_REQUEST_NULL = 0, MPI_SUCCESS = 0, MPI_UNDEFINED = -32766 };
enum { MPI_THREAD_SINGLE = 0, MPI_THREAD_FUNNELED = 1, MPI_THREAD_SERIALIZED = 2, MPI_THREAD_MULTIPLE = 3 };
enum { MPI_SUM = 1, MPI_MAX, MPI_MIN, MPI_PROD, MPI_MAXLOC, MPI_MINLOC, MPI_LAND, MPI_LOR };
enum { MPI_MAX_ERROR_STRING = 128, MPI_TAG_UB = 64 };

#define MPI_STATUS_IGNORE    ((MPI_Status*) 0)
#define MPI_STATUSES_IGNORE  ((MPI_Status*) 0)
#define MPI_IN_PLACE         ((void*) -1)

/* basic datatypes: small codes (sizes in mpi_mp.cc); derived datatypes are handles >= 1000 */
#define MPI_BYTE              ((MPI_Datatype) 1)
#define MPI_CHAR              ((MPI_Datatype) 1)
#define MPI_CXX_BOOL          ((MPI_Datatype) 1)
#define MPI_INT               ((MPI_Datatype) 2)
#define MPI_UNSIGNED          ((MPI_Datatype) 3)
#define MPI_LONG              ((MPI_Datatype) 4)
#define MPI_INT64_T           ((MPI_Datatype) 4)
#define MPI_FLOAT             ((MPI_Datatype) 5)
#define MPI_DOUBLE            ((MPI_Datatype) 6)
#define MPI_C_COMPLEX         ((MPI_Datatype) 7)
#define MPI_C_FLOAT_COMPLEX   ((MPI_Datatype) 7)
#define MPI_C_DOUBLE_COMPLEX  ((MPI_Datatype) 8)
#define MPI_2INT              ((MPI_Datatype) 9)
#define MPI_FLOAT_INT         ((MPI_Datatype) 10)
#define MPI_DOUBLE_INT        ((MPI_Datatype) 11)

int MPI_Init(int* argc, char*** argv);
int MPI_Init_thread(int* argc, char*** argv, int required, int* provided);
int MPI_Initialized(int* flag);
int MPI_Finalize(void);
double MPI_Wtime(void);
int MPI_Error_string(int code, char* str, int* len);

int MPI_Comm_rank(MPI_Comm c, int* rank);
int MPI_Comm_size(MPI_Comm c, int* size);
int MPI_Comm_group(MPI_Comm c, MPI_Group* g);
int MPI_Comm_free(MPI_Comm* c);
int MPI_Comm_create_group(MPI_Comm c, MPI_Group g, int tag, MPI_Comm* out);
int MPI_Group_incl(MPI_Group g, int n, const int* ranks, MPI_Group* out);
int MPI_Group_free(MPI_Group* g);
int MPI_Group_translate_ranks(MPI_Group a, int n, const int* in, MPI_Group b, int* out);

int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype* t);
int MPI_Type_vector(int count, int blocklen, int stride, MPI_Datatype old, MPI_Datatype* t);
int MPI_Type_commit(MPI_Datatype* t);
int MPI_Type_free(MPI_Datatype* t);
int MPI_Op_create(MPI_User_function* f, int commute, MPI_Op* op);
int MPI_Op_free(MPI_Op* op);

int MPI_Barrier(MPI_Comm c);
int MPI_Bcast(void* buf, int n, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);
int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* counts, const int* displs, MPI_Datatype rt, MPI_Comm c);

int MPI_Wait(MPI_Request* r, MPI_Status* s);
int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s);
int MPI_Request_free(MPI_Request* r);
int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c);
int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* s);
int MPI_Isend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* r);
int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* r);
int MPI_Sendrecv(const void* sb, int sn, MPI_Datatype st, int dst, int stag, void* rb, int rn, MPI_Datatype rt, int src, int rtag,
                 MPI_Comm c, MPI_Status* s);

#ifdef __cplusplus
}
#endif
#endif /* SB200_ORACLE_MPI_MP_H */
