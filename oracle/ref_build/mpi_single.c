/* Single-rank MPI semantics for the oracle build of the reference's C sources.
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md): never linked into the product. */
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static size_t dt_size(MPI_Datatype t) {
  switch (t) {
  case MPI_CHAR: case MPI_BYTE: return 1;
  case MPI_INT: case MPI_UINT: case MPI_FLOAT: return 4;
  case MPI_DOUBLE: case MPI_UNSIGNED_LONG_LONG: case MPI_LONG: case MPI_2INT: return 8;
  case MPI_DOUBLE_INT: return 16;
  default: return 0; /* derived types: nothing to move on one rank */
  }
}
static int copy_if(const void *s, void *r, int n, MPI_Datatype t) {
  if (s != MPI_IN_PLACE && s != r && s && r) memcpy(r, s, (size_t)n * dt_size(t));
  return MPI_SUCCESS;
}
static void no_p2p(const char *w) {
  fprintf(stderr, "oracle mpi_single: %s called on a single rank\n", w);
  abort();
}
int MPI_Init(int *a, char ***b) { (void)a; (void)b; return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm c, int e) { (void)c; fprintf(stderr, "MPI_Abort(%d)\n", e); fflush(NULL); _Exit(e ? e : 1); }
int MPI_Comm_size(MPI_Comm c, int *n) { (void)c; *n = 1; return 0; }
int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
double MPI_Wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c) { (void)o; (void)c; return copy_if(s, r, n, t); }
int MPI_Iallreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c, MPI_Request *q) { (void)o; (void)c; if (q) *q = 0; return copy_if(s, r, n, t); }
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, int root, MPI_Comm c) { (void)o; (void)c; (void)root; return copy_if(s, r, n, t); }
int MPI_Scan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c) { (void)o; (void)c; return copy_if(s, r, n, t); }
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
int MPI_Gather(const void *s, int n, MPI_Datatype t, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { (void)rn; (void)rt; (void)root; (void)c; return copy_if(s, r, n, t); }
int MPI_Allgather(const void *s, int n, MPI_Datatype t, void *r, int rn, MPI_Datatype rt, MPI_Comm c) { (void)rn; (void)rt; (void)c; return copy_if(s, r, n, t); }
int MPI_Scatterv(const void *s, const int *cnt, const int *dis, MPI_Datatype t, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  (void)rn; (void)rt; (void)root; (void)c;
  if (s && r) memcpy(r, (const char *)s + (size_t)dis[0] * dt_size(t), (size_t)cnt[0] * dt_size(t));
  return 0;
}
int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; no_p2p("MPI_Send"); return 1; }
int MPI_Recv(void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Status *s) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)s; no_p2p("MPI_Recv"); return 1; }
int MPI_Isend(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *q) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)q; no_p2p("MPI_Isend"); return 1; }
int MPI_Irecv(void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *q) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)q; no_p2p("MPI_Irecv"); return 1; }
int MPI_Wait(MPI_Request *q, MPI_Status *s) { (void)q; (void)s; return 0; }
int MPI_Waitall(int n, MPI_Request *q, MPI_Status *s) { (void)n; (void)q; (void)s; return 0; }
int MPI_Get_count(const MPI_Status *s, MPI_Datatype t, int *n) { (void)s; (void)t; *n = 0; return 0; }
int MPI_Get_address(const void *p, MPI_Aint *a) { *a = (MPI_Aint)p; return 0; }
int MPI_Type_create_struct(int n, const int *b, const MPI_Aint *d, const MPI_Datatype *t, MPI_Datatype *nt) { (void)n; (void)b; (void)d; (void)t; *nt = 1000; return 0; }
int MPI_Type_commit(MPI_Datatype *t) { (void)t; return 0; }
int MPI_Type_free(MPI_Datatype *t) { (void)t; return 0; }
int MPI_Type_size(MPI_Datatype t, int *s) { *s = (int)dt_size(t); return 0; }
int MPI_Type_get_extent(MPI_Datatype t, MPI_Aint *lb, MPI_Aint *e) { *lb = 0; *e = (MPI_Aint)dt_size(t); return 0; }
int MPI_Error_string(int e, char *s, int *l) { *l = sprintf(s, "mpi_single error %d", e); return 0; }
