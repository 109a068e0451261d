/* hd_mpi.c -- implementation of the control-plane MPI shim (include/mpi.h): rank and size come
 * from the NCCL communicator when one exists, else from the launcher's RANK / WORLD_SIZE. */
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "hdk.h"

static int g_mpi_init = 0;

static int env_int(const char *name, int dflt)
{
   const char *v = getenv(name);
   return v ? atoi(v) : dflt;
}

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; g_mpi_init = 1; return MPI_SUCCESS; }
int MPI_Initialized(int *flag) { *flag = g_mpi_init; return MPI_SUCCESS; }
int MPI_Finalize(void) { g_mpi_init = 0; return MPI_SUCCESS; }

int MPI_Comm_rank(MPI_Comm comm, int *rank)
{
   if (comm == MPI_COMM_SELF) { *rank = 0; return MPI_SUCCESS; }
   *rank = hdk_comm_size() > 1 ? hdk_comm_rank() : env_int("RANK", 0);
   return MPI_SUCCESS;
}

int MPI_Comm_size(MPI_Comm comm, int *size)
{
   if (comm == MPI_COMM_SELF) { *size = 1; return MPI_SUCCESS; }
   *size = hdk_comm_size() > 1 ? hdk_comm_size() : env_int("WORLD_SIZE", 1);
   return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm comm)
{
   /* all ranks meet in a collective of the NCCL communicator when there is one */
   int64_t sum = 0;
   hdk_sync();
   if (comm != MPI_COMM_SELF && hdk_comm_size() > 1) hdk_comm_sum_i64(1, &sum);
   return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int errorcode)
{
   (void)comm;
   fprintf(stderr, "Abort(%d)\n", errorcode);
   exit(errorcode ? errorcode : EXIT_FAILURE);
}

double MPI_Wtime(void)
{
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---- Cartesian communicators: handles 16.. index a small table of (dims, size) ---- */
#define HD_MAX_CART 16
static struct { int used, ndims, dims[3]; } g_cart[HD_MAX_CART];

static int cart_slot(MPI_Comm comm) { return (comm >= 16 && comm < 16 + HD_MAX_CART && g_cart[comm - 16].used) ? comm - 16 : -1; }

int MPI_Cart_create(MPI_Comm comm, int ndims, const int dims[], const int periods[], int reorder, MPI_Comm *cart)
{
   (void)periods; (void)reorder;
   int size = 1, want = 1;
   MPI_Comm_size(comm, &size);
   if (ndims < 1 || ndims > 3) return 1;
   for (int d = 0; d < ndims; d++) want *= dims[d];
   if (want != size) { fprintf(stderr, "MPI shim: Cartesian grid of %d processes on a communicator of %d\n", want, size); return 1; }
   for (int s = 0; s < HD_MAX_CART; s++)
      if (!g_cart[s].used)
      {
         g_cart[s].used = 1; g_cart[s].ndims = ndims;
         for (int d = 0; d < 3; d++) g_cart[s].dims[d] = d < ndims ? dims[d] : 1;
         *cart = 16 + s;
         return MPI_SUCCESS;
      }
   return 1;
}

int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int coords[])
{
   int s = cart_slot(comm);
   if (s < 0) return 1;
   /* row-major like MPI: the last dimension varies fastest */
   for (int d = g_cart[s].ndims - 1; d >= 0; d--)
   {
      if (d < maxdims) coords[d] = rank % g_cart[s].dims[d];
      rank /= g_cart[s].dims[d];
   }
   return MPI_SUCCESS;
}

int MPI_Cart_rank(MPI_Comm comm, const int coords[], int *rank)
{
   int s = cart_slot(comm), r = 0;
   if (s < 0) return 1;
   for (int d = 0; d < g_cart[s].ndims; d++)
   {
      if (coords[d] < 0 || coords[d] >= g_cart[s].dims[d]) { *rank = MPI_PROC_NULL; return 1; }
      r = r * g_cart[s].dims[d] + coords[d];
   }
   *rank = r;
   return MPI_SUCCESS;
}

int MPI_Cart_shift(MPI_Comm comm, int direction, int disp, int *source, int *dest)
{
   int s = cart_slot(comm), me = 0, c[3] = {0, 0, 0}, t[3];
   if (s < 0 || direction < 0 || direction >= g_cart[s].ndims) return 1;
   MPI_Comm_rank(comm, &me);
   MPI_Cart_coords(comm, me, 3, c);
   memcpy(t, c, sizeof(t)); t[direction] = c[direction] - disp;
   if (t[direction] < 0 || t[direction] >= g_cart[s].dims[direction]) *source = MPI_PROC_NULL; else MPI_Cart_rank(comm, t, source);
   memcpy(t, c, sizeof(t)); t[direction] = c[direction] + disp;
   if (t[direction] < 0 || t[direction] >= g_cart[s].dims[direction]) *dest = MPI_PROC_NULL; else MPI_Cart_rank(comm, t, dest);
   return MPI_SUCCESS;
}

int MPI_Comm_free(MPI_Comm *comm)
{
   int s = comm ? cart_slot(*comm) : -1;
   if (s >= 0) g_cart[s].used = 0;
   if (comm) *comm = MPI_COMM_NULL;
   return MPI_SUCCESS;
}

static int needs_peer(const char *what, int peer)
{
   if (peer == MPI_PROC_NULL) return 0;
   fprintf(stderr, "MPI shim: %s to/from process %d is not available without an MPI library "
                   "(the data path of hypredrive_b200 uses NCCL; host point-to-point is caller-side only)\n", what, peer);
   return 1;
}
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{
   (void)buf; (void)count; (void)type; (void)tag; (void)comm;
   if (req) *req = MPI_REQUEST_NULL;
   return needs_peer("MPI_Isend", dest);
}
int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request *req)
{
   (void)buf; (void)count; (void)type; (void)tag; (void)comm;
   if (req) *req = MPI_REQUEST_NULL;
   return needs_peer("MPI_Irecv", source);
}
int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status statuses[]) { (void)count; (void)reqs; (void)statuses; return MPI_SUCCESS; }

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm)
{
   int size = 1;
   MPI_Comm_size(comm, &size);
   size_t w = (type == MPI_DOUBLE || type == MPI_LONG_LONG_INT) ? 8 : (type == MPI_CHAR ? 1 : 4);
   if (sendbuf != recvbuf && sendbuf) memcpy(recvbuf, sendbuf, w * (size_t)count);
   if (size <= 1) return MPI_SUCCESS;
   /* several ranks: 64-bit integers and doubles go through the NCCL communicator (control plane) */
   if (type == MPI_DOUBLE && op == MPI_SUM && count == 1 && sizeof(double) == 8)
   {
      /* exact for the integer-valued quantities callers reduce this way (counts, timings are informative) */
      long long v = 0, gsum = 0;
      double    d = *(double *)recvbuf;
      if (d == (double)(long long)d) { v = (long long)d; if (hdk_comm_sum_i64(v, (int64_t *)&gsum) == 0) { *(double *)recvbuf = (double)gsum; return MPI_SUCCESS; } }
   }
   if (type == MPI_LONG_LONG_INT && count == 1)
   {
      int64_t v = *(long long *)recvbuf, gv = 0;
      int rc = (op == MPI_MAX) ? hdk_comm_max_i64(v, &gv) : (op == MPI_SUM ? hdk_comm_sum_i64(v, &gv) : 1);
      if (rc == 0) { *(long long *)recvbuf = gv; return MPI_SUCCESS; }
   }
   fprintf(stderr, "MPI shim: MPI_Allreduce(type %d, op %d, count %d) over %d processes is not available without an MPI library\n",
           type, op, count, size);
   return 1;
}
