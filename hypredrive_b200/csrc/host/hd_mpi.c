/* hd_mpi.c -- implementation of the control-plane MPI shim (include/mpi.h): rank and size come
 * from the NCCL communicator when one exists, else from the launcher's RANK / WORLD_SIZE. */
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
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

int MPI_Barrier(MPI_Comm comm) { (void)comm; hdk_sync(); return MPI_SUCCESS; }

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
