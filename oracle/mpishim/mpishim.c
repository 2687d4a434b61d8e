/* TEST INFRASTRUCTURE ONLY -- see mpi.h.  Eager point-to-point over per-(source,
 * destination, tag) slots in one shared-memory segment; collectives over a slot table and a
 * sense-reversing barrier.  LULESH has at most one message in flight per (source,
 * destination, tag), which is all the slots support. */
#define _GNU_SOURCE
#include "mpi.h"

#include <fcntl.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#define SHIM_MAX_RANKS 64
#define SHIM_TAGS 4            /* tag / 1024 in 1..3 (MSG_COMM_SBN, MSG_SYNC_POS_VEL, MSG_MONOQ) */
#define SHIM_MAX_REDUCE 8
#define SHIM_MAX_REQ 256

typedef struct {
   int nranks;
   long slot_bytes;
   volatile int barrier_count, barrier_gen;
   volatile int abort_code;
   double reduce[SHIM_MAX_RANKS][SHIM_MAX_REDUCE];
} shim_header;

typedef struct { volatile int full; int bytes; } slot_header;

static shim_header *g_hdr;
static char *g_slots;
static int g_rank = 0, g_size = 1;
static struct { void *buf; long bytes; int src, tag, active; } g_req[SHIM_MAX_REQ];

static size_t slot_stride(void) { return sizeof(slot_header) + (size_t)g_hdr->slot_bytes; }

static slot_header *slot(int src, int dst, int tag)
{
   const int t = tag / 1024;
   if (t < 1 || t >= SHIM_TAGS) { fprintf(stderr, "mpishim: unsupported tag %d\n", tag); abort(); }
   const size_t idx = ((size_t)src * g_size + dst) * SHIM_TAGS + t;
   return (slot_header *)(g_slots + idx * slot_stride());
}

static void check_abort(void)
{
   if (g_hdr->abort_code) exit(g_hdr->abort_code & 0xff ? g_hdr->abort_code : 1);
}

size_t mpishim_segment_bytes(int nranks, long slot_bytes)
{
   return sizeof(shim_header) + (size_t)nranks * nranks * SHIM_TAGS * (sizeof(slot_header) + (size_t)slot_bytes);
}

int MPI_Init(int *argc, char ***argv)
{
   (void)argc; (void)argv;
   const char *name = getenv("MPISHIM_SHM");
   if (!name) {   /* not under mpirun_shim: a private one-rank world */
      static shim_header solo;
      solo.nranks = 1;
      g_hdr = &solo; g_rank = 0; g_size = 1;
      return MPI_SUCCESS;
   }
   g_rank = atoi(getenv("MPISHIM_RANK"));
   g_size = atoi(getenv("MPISHIM_SIZE"));
   const long slot_bytes = atol(getenv("MPISHIM_SLOT_BYTES"));
   int fd = shm_open(name, O_RDWR, 0600);
   if (fd < 0) { perror("mpishim: shm_open"); exit(1); }
   const size_t bytes = mpishim_segment_bytes(g_size, slot_bytes);
   void *p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
   if (p == MAP_FAILED) { perror("mpishim: mmap"); exit(1); }
   close(fd);
   g_hdr = (shim_header *)p;
   g_slots = (char *)p + sizeof(shim_header);
   return MPI_SUCCESS;
}

int MPI_Init_thread(int *argc, char ***argv, int required, int *provided)
{
   (void)required;
   if (provided) *provided = MPI_THREAD_FUNNELED;
   return MPI_Init(argc, argv);
}

int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = g_size; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = g_rank; return MPI_SUCCESS; }

int MPI_Barrier(MPI_Comm c)
{
   (void)c;
   if (g_size == 1) return MPI_SUCCESS;
   const int gen = g_hdr->barrier_gen;
   if (__sync_add_and_fetch(&g_hdr->barrier_count, 1) == g_size) {
      g_hdr->barrier_count = 0;
      __sync_synchronize();
      g_hdr->barrier_gen = gen + 1;
   } else {
      while (g_hdr->barrier_gen == gen) { check_abort(); sched_yield(); }
   }
   __sync_synchronize();
   return MPI_SUCCESS;
}

int MPI_Finalize(void) { return MPI_Barrier(MPI_COMM_WORLD); }

int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm c, MPI_Request *req)
{
   (void)c;
   slot_header *s = slot(g_rank, dest, tag);
   const long bytes = (long)count * type;
   if (bytes > g_hdr->slot_bytes) { fprintf(stderr, "mpishim: message of %ld bytes exceeds the slot (raise MPISHIM_SLOT_MB)\n", bytes); abort(); }
   while (s->full) { check_abort(); sched_yield(); }
   memcpy((char *)(s + 1), buf, (size_t)bytes);
   s->bytes = (int)bytes;
   __sync_synchronize();
   s->full = 1;
   if (req) *req = MPI_REQUEST_NULL;   /* eager: complete on return */
   return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm c, MPI_Request *req)
{
   (void)c;
   for (int i = 1; i < SHIM_MAX_REQ; ++i)
      if (!g_req[i].active) {
         g_req[i].buf = buf; g_req[i].bytes = (long)count * type;
         g_req[i].src = source; g_req[i].tag = tag; g_req[i].active = 1;
         *req = i;
         return MPI_SUCCESS;
      }
   fprintf(stderr, "mpishim: out of request slots\n");
   abort();
}

int MPI_Wait(MPI_Request *req, MPI_Status *status)
{
   const int i = *req;
   if (i == MPI_REQUEST_NULL) return MPI_SUCCESS;
   slot_header *s = slot(g_req[i].src, g_rank, g_req[i].tag);
   while (!s->full) { check_abort(); sched_yield(); }
   __sync_synchronize();
   if (s->bytes > g_req[i].bytes) { fprintf(stderr, "mpishim: message truncated\n"); abort(); }
   memcpy(g_req[i].buf, (char *)(s + 1), (size_t)s->bytes);
   if (status) { status->MPI_SOURCE = g_req[i].src; status->MPI_TAG = g_req[i].tag; status->MPI_ERROR = 0; }
   __sync_synchronize();
   s->full = 0;
   g_req[i].active = 0;
   *req = MPI_REQUEST_NULL;
   return MPI_SUCCESS;
}

int MPI_Waitall(int count, MPI_Request *reqs, MPI_Status *statuses)
{
   for (int i = 0; i < count; ++i) MPI_Wait(&reqs[i], statuses ? &statuses[i] : NULL);
   return MPI_SUCCESS;
}

static void reduce_all(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int store)
{
   if (count > SHIM_MAX_REDUCE) { fprintf(stderr, "mpishim: reduce count %d too large\n", count); abort(); }
   double v[SHIM_MAX_REDUCE];
   for (int k = 0; k < count; ++k) v[k] = (type == MPI_DOUBLE) ? ((const double *)sendbuf)[k] : ((const float *)sendbuf)[k];
   if (g_size > 1) {
      for (int k = 0; k < count; ++k) g_hdr->reduce[g_rank][k] = v[k];
      MPI_Barrier(MPI_COMM_WORLD);
      for (int k = 0; k < count; ++k) {
         double r = g_hdr->reduce[0][k];
         for (int p = 1; p < g_size; ++p) {
            const double x = g_hdr->reduce[p][k];
            if (op == MPI_MIN) r = x < r ? x : r;
            else if (op == MPI_MAX) r = x > r ? x : r;
            else r += x;
         }
         v[k] = r;
      }
      MPI_Barrier(MPI_COMM_WORLD);   /* nobody overwrites the table before everybody has read it */
   }
   if (store)
      for (int k = 0; k < count; ++k) {
         if (type == MPI_DOUBLE) ((double *)recvbuf)[k] = v[k];
         else ((float *)recvbuf)[k] = (float)v[k];
      }
}

int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype type, MPI_Op op, MPI_Comm c)
{
   (void)c;
   reduce_all(s, r, count, type, op, 1);
   return MPI_SUCCESS;
}

int MPI_Reduce(const void *s, void *r, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm c)
{
   (void)c;
   reduce_all(s, r, count, type, op, g_rank == root);
   return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm c, int errorcode)
{
   (void)c;
   if (g_hdr && g_size > 1) g_hdr->abort_code = errorcode ? errorcode : 1;
   exit(errorcode);
}

double MPI_Wtime(void)
{
   struct timespec t;
   clock_gettime(CLOCK_MONOTONIC, &t);
   return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
