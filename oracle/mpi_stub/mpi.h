/* oracle/mpi_stub/mpi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * One-rank MPI replacement so the reference SLATE sources under /root/reference
 * can be compiled in an image that has no MPI.  Only the ~60 MPI names the
 * reference actually uses are provided.  With a single rank every collective is
 * the identity on the local buffer and every point-to-point call is unreachable
 * (the reference short-circuits them when the rank set has one member), so
 * those abort loudly if they are ever hit.
 *
 * A datatype handle is simply the byte size of one element, which lets the
 * reduction stubs copy the right amount of data without a type table.
 */
#ifndef SB200_ORACLE_MPI_STUB_H
#define SB200_ORACLE_MPI_STUB_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

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
enum { MPI_GROUP_NULL = 0, MPI_REQUEST_NULL = 0, MPI_SUCCESS = 0 };
enum { MPI_THREAD_SINGLE = 0, MPI_THREAD_FUNNELED = 1,
       MPI_THREAD_SERIALIZED = 2, MPI_THREAD_MULTIPLE = 3 };
enum { MPI_SUM = 1, MPI_MAX, MPI_MIN, MPI_PROD, MPI_MAXLOC, MPI_MINLOC, MPI_LAND, MPI_LOR };
enum { MPI_MAX_ERROR_STRING = 128, MPI_TAG_UB = 64 };

#define MPI_STATUS_IGNORE    ((MPI_Status*) 0)
#define MPI_STATUSES_IGNORE  ((MPI_Status*) 0)
#define MPI_IN_PLACE         ((void*) -1)

#define MPI_BYTE              ((MPI_Datatype) 1)
#define MPI_CHAR              ((MPI_Datatype) 1)
#define MPI_CXX_BOOL          ((MPI_Datatype) 1)
#define MPI_INT               ((MPI_Datatype) sizeof(int))
#define MPI_UNSIGNED          ((MPI_Datatype) sizeof(unsigned))
#define MPI_LONG              ((MPI_Datatype) sizeof(long))
#define MPI_INT64_T           ((MPI_Datatype) 8)
#define MPI_FLOAT             ((MPI_Datatype) 4)
#define MPI_DOUBLE            ((MPI_Datatype) 8)
#define MPI_C_COMPLEX         ((MPI_Datatype) 8)
#define MPI_C_FLOAT_COMPLEX   ((MPI_Datatype) 8)
#define MPI_C_DOUBLE_COMPLEX  ((MPI_Datatype) 16)
#define MPI_2INT              ((MPI_Datatype) 8)
#define MPI_FLOAT_INT         ((MPI_Datatype) 8)
#define MPI_DOUBLE_INT        ((MPI_Datatype) 16)

#define SB200_MPI_UNREACHABLE(name) \
    do { fprintf(stderr, "mpi stub: %s called with a single rank\n", name); abort(); } while (0)

static inline int sb200_mpi_local_copy(const void* src, void* dst, long bytes)
{
    if (src != MPI_IN_PLACE && src != dst && bytes > 0)
        memcpy(dst, src, (size_t) bytes);
    return MPI_SUCCESS;
}

/* environment */
static inline int MPI_Init(int* argc, char*** argv) { (void) argc; (void) argv; return MPI_SUCCESS; }
static inline int MPI_Init_thread(int* argc, char*** argv, int required, int* provided)
{ (void) argc; (void) argv; *provided = required; return MPI_SUCCESS; }
static inline int MPI_Initialized(int* flag) { *flag = 1; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline double MPI_Wtime(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec;
}
static inline int MPI_Error_string(int code, char* str, int* len)
{ *len = snprintf(str, MPI_MAX_ERROR_STRING, "mpi stub error %d", code); return MPI_SUCCESS; }

/* communicators and groups: there is exactly one member everywhere */
static inline int MPI_Comm_rank(MPI_Comm c, int* rank) { (void) c; *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int* size) { (void) c; *size = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_group(MPI_Comm c, MPI_Group* g) { (void) c; *g = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_free(MPI_Comm* c) { *c = MPI_COMM_NULL; return MPI_SUCCESS; }
static inline int MPI_Comm_create_group(MPI_Comm c, MPI_Group g, int tag, MPI_Comm* out)
{ (void) c; (void) g; (void) tag; *out = MPI_COMM_SELF; return MPI_SUCCESS; }
static inline int MPI_Group_incl(MPI_Group g, int n, const int* ranks, MPI_Group* out)
{ (void) g; (void) n; (void) ranks; *out = 1; return MPI_SUCCESS; }
static inline int MPI_Group_free(MPI_Group* g) { *g = MPI_GROUP_NULL; return MPI_SUCCESS; }
static inline int MPI_Group_translate_ranks(MPI_Group a, int n, const int* in, MPI_Group b, int* out)
{ (void) a; (void) b; for (int i = 0; i < n; ++i) out[i] = in[i]; return MPI_SUCCESS; }

/* derived datatypes: keep only the payload size */
static inline int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype* t)
{ *t = (MPI_Datatype) count * old; return MPI_SUCCESS; }
static inline int MPI_Type_vector(int count, int blocklen, int stride, MPI_Datatype old, MPI_Datatype* t)
{ (void) stride; *t = (MPI_Datatype) count * blocklen * old; return MPI_SUCCESS; }
static inline int MPI_Type_commit(MPI_Datatype* t) { (void) t; return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype* t) { (void) t; return MPI_SUCCESS; }
static inline int MPI_Op_create(MPI_User_function* f, int commute, MPI_Op* op)
{ (void) f; (void) commute; *op = 100; return MPI_SUCCESS; }
static inline int MPI_Op_free(MPI_Op* op) { (void) op; return MPI_SUCCESS; }

/* collectives over one rank */
static inline int MPI_Barrier(MPI_Comm c) { (void) c; return MPI_SUCCESS; }
static inline int MPI_Bcast(void* buf, int n, MPI_Datatype t, int root, MPI_Comm c)
{ (void) buf; (void) n; (void) t; (void) root; (void) c; return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{ (void) op; (void) c; return sb200_mpi_local_copy(s, r, (long) n * t); }
static inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{ (void) op; (void) root; (void) c; return sb200_mpi_local_copy(s, r, (long) n * t); }
static inline int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* counts,
                                 const int* displs, MPI_Datatype rt, MPI_Comm c)
{ (void) counts; (void) c; return sb200_mpi_local_copy(s, (char*) r + (long) displs[0] * rt, (long) sn * st); }

/* requests complete immediately */
static inline int MPI_Wait(MPI_Request* r, MPI_Status* s) { (void) r; (void) s; return MPI_SUCCESS; }
static inline int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s) { (void) n; (void) r; (void) s; return MPI_SUCCESS; }
static inline int MPI_Request_free(MPI_Request* r) { *r = MPI_REQUEST_NULL; return MPI_SUCCESS; }

/* point to point: never legal with one rank */
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c)
{ (void) b; (void) n; (void) t; (void) dst; (void) tag; (void) c; SB200_MPI_UNREACHABLE("MPI_Send"); return 1; }
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* s)
{ (void) b; (void) n; (void) t; (void) src; (void) tag; (void) c; (void) s; SB200_MPI_UNREACHABLE("MPI_Recv"); return 1; }
static inline int MPI_Isend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* r)
{ (void) b; (void) n; (void) t; (void) dst; (void) tag; (void) c; (void) r; SB200_MPI_UNREACHABLE("MPI_Isend"); return 1; }
static inline int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* r)
{ (void) b; (void) n; (void) t; (void) src; (void) tag; (void) c; (void) r; SB200_MPI_UNREACHABLE("MPI_Irecv"); return 1; }
static inline int MPI_Sendrecv(const void* sb, int sn, MPI_Datatype st, int dst, int stag,
                               void* rb, int rn, MPI_Datatype rt, int src, int rtag,
                               MPI_Comm c, MPI_Status* s)
{ (void) sb; (void) sn; (void) st; (void) dst; (void) stag; (void) rb; (void) rn; (void) rt;
  (void) src; (void) rtag; (void) c; (void) s; SB200_MPI_UNREACHABLE("MPI_Sendrecv"); return 1; }

#ifdef __cplusplus
}
#endif
#endif /* SB200_ORACLE_MPI_STUB_H */
