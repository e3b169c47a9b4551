/* Single-rank MPI stand-in used ONLY to compile the reference's C sources into
 * oracle/_ref (test infrastructure, never shipped in the product path).
 * Semantics: one rank, collectives are copies, p2p is an error. */
#ifndef GOMA_B200_ORACLE_MPI_STUB_H
#define GOMA_B200_ORACLE_MPI_STUB_H
#include <stddef.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef long MPI_Aint;
typedef long long MPI_Count;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_CHAR 1
#define MPI_BYTE 2
#define MPI_INT 3
#define MPI_UINT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 5
#define MPI_DOUBLE 6
#define MPI_DOUBLE_INT 7
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_LONG 9
#define MPI_2INT 10
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_LOR 4
#define MPI_BOR 5
#define MPI_MAXLOC 6
#define MPI_MINLOC 7
#define MPI_LAND 8
#define MPI_IN_PLACE ((void *)1)
#define MPI_BOTTOM ((void *)0)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_MAX_ERROR_STRING 256
int MPI_Init(int *, char ***);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm, int);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Barrier(MPI_Comm);
double MPI_Wtime(void);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Iallreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm, MPI_Request *);
int MPI_Reduce(const void *, void *, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Scan(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Gather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Scatterv(const void *, const int *, const int *, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *);
int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Wait(MPI_Request *, MPI_Status *);
int MPI_Waitall(int, MPI_Request *, MPI_Status *);
int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *);
int MPI_Get_address(const void *, MPI_Aint *);
int MPI_Type_create_struct(int, const int *, const MPI_Aint *, const MPI_Datatype *, MPI_Datatype *);
int MPI_Type_commit(MPI_Datatype *);
int MPI_Type_free(MPI_Datatype *);
int MPI_Type_size(MPI_Datatype, int *);
int MPI_Type_get_extent(MPI_Datatype, MPI_Aint *, MPI_Aint *);
int MPI_Error_string(int, char *, int *);
#endif
