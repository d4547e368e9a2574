/*
 * Single-rank MPI stand-in used ONLY to compile the unmodified GIRIH reference
 * (from /root/reference) in a container that has no MPI installation.
 *
 * TEST INFRASTRUCTURE -- not product code.  Nothing under girih_b200/ may
 * include this header.
 *
 * Semantics: one process, rank 0 of 1.  Every neighbour is MPI_PROC_NULL, so
 * the reference never enters an exchange branch (each one is guarded by
 * p->t.shape[d] > 1).  Point-to-point, datatype and collective calls succeed
 * without doing anything, except MPI_Reduce/MPI_Bcast which behave as the
 * one-rank case requires (Reduce copies send -> recv).
 *
 * The reference's loop templates call omp_get_thread_num() without including
 * <omp.h>, and use memcpy/time without their headers, so they are pulled in here.
 */
#ifndef GIRIH_ORACLE_MPI_SHIM_H
#define GIRIH_ORACLE_MPI_SHIM_H

#include <string.h>
#include <time.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD        0
#define MPI_SUCCESS           0
#define MPI_ANY_TAG          (-1)
#define MPI_PROC_NULL        (-2)
#define MPI_ORDER_FORTRAN     1
#define MPI_DOUBLE            8
#define MPI_FLOAT             4
#define MPI_INT               5   /* sizeof == 4; distinguished from FLOAT by value */
#define MPI_SUM               1
#define MPI_MIN               2
#define MPI_MAX               3
#define MPI_THREAD_MULTIPLE   3
#define MPI_MAX_ERROR_STRING  256
#define MPI_ERRORS_RETURN     0

static inline size_t girih_shim_sizeof(MPI_Datatype t)
{ return t == MPI_DOUBLE ? 8u : 4u; }

static inline int MPI_Init(int *c, char ***v) { (void)c; (void)v; return 0; }
static inline int MPI_Init_thread(int *c, char ***v, int req, int *prov)
{ (void)c; (void)v; *prov = req; return 0; }
static inline int MPI_Query_thread(int *prov) { *prov = MPI_THREAD_MULTIPLE; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline double MPI_Wtime(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline int MPI_Cart_create(MPI_Comm o, int nd, int *dims, int *per, int reorder, MPI_Comm *n)
{ (void)o; (void)nd; (void)dims; (void)per; (void)reorder; *n = 0; return 0; }
static inline int MPI_Cart_coords(MPI_Comm c, int rank, int nd, int *coords)
{ int i; (void)c; (void)rank; for (i = 0; i < nd; i++) coords[i] = 0; return 0; }
static inline int MPI_Cart_shift(MPI_Comm c, int dir, int disp, int *src, int *dst)
{ (void)c; (void)dir; (void)disp; *src = MPI_PROC_NULL; *dst = MPI_PROC_NULL; return 0; }
static inline int MPI_Type_create_subarray(int nd, int *sizes, int *sub, int *starts, int order,
                                           MPI_Datatype old, MPI_Datatype *newt)
{ (void)nd; (void)sizes; (void)sub; (void)starts; (void)order; *newt = old; return 0; }
static inline int MPI_Type_contiguous(int n, MPI_Datatype old, MPI_Datatype *newt)
{ (void)n; *newt = old; return 0; }
static inline int MPI_Type_commit(MPI_Datatype *t) { (void)t; return 0; }
static inline int MPI_Type_free(MPI_Datatype *t) { (void)t; return 0; }
static inline int MPI_Isend(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *r)
{ (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; *r = 0; return 0; }
static inline int MPI_Irecv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *r)
{ (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; *r = 0; return 0; }
static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; return 0; }
static inline int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *s)
{ (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)s; return 0; }
static inline int MPI_Wait(MPI_Request *r, MPI_Status *s) { (void)r; (void)s; return 0; }
static inline int MPI_Waitall(int n, MPI_Request *r, MPI_Status *s) { (void)n; (void)r; (void)s; return 0; }
static inline int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{ (void)op; (void)root; (void)c; memcpy(r, s, (size_t)n * girih_shim_sizeof(t)); return 0; }

#endif /* GIRIH_ORACLE_MPI_SHIM_H */
