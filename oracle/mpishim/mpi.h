/*
 * TEST INFRASTRUCTURE ONLY -- a single-node stand-in for <mpi.h>.
 *
 * This image has no MPI (no mpicxx, mpirun, mpi.h or libmpi), so the reference's
 * USE_MPI=1 path cannot be built with stock tools ("next" row N2).  This
 * header + mpishim.c implement exactly the 14 MPI-1 entry points lulesh.cc,
 * lulesh-comm.cc, lulesh-init.cc and lulesh-util.cc use, over fork()ed processes and one
 * POSIX shared-memory segment (launcher: oracle/_ref/mpirun_shim -np N prog args...).
 * With it the UNMODIFIED reference compiles with -DUSE_MPI=1, which gives a true
 * multi-rank oracle (its own CommSBN / CommSyncPosVel / CommMonoQ code paths) and the
 * "MPI+OpenMP -np 8" CPU baseline.  It is not a general MPI.
 */
#ifndef LULESH_B200_MPISHIM_H
#define LULESH_B200_MPISHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_REQUEST_NULL 0
#define MPI_FLOAT 4          /* the value is the element size in bytes */
#define MPI_DOUBLE 8
#define MPI_MIN 1
#define MPI_MAX 2
#define MPI_SUM 3
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)

int MPI_Init(int *argc, char ***argv);
int MPI_Init_thread(int *argc, char ***argv, int required, int *provided);
int MPI_Finalize(void);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Wait(MPI_Request *req, MPI_Status *status);
int MPI_Waitall(int count, MPI_Request *reqs, MPI_Status *statuses);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);

#ifdef __cplusplus
}
#endif
#endif
