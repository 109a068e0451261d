/* mpi.h -- single-header MPI shim for hypredrive_b200.
 *
 * The reference API carries an MPI_Comm in HYPREDRV_Create (include/HYPREDRV.h:257 of the
 * reference).  This image has no MPI; ranks are one process per GPU launched by torchrun or any
 * launcher that exports RANK / WORLD_SIZE, and all data-path communication goes through NCCL
 * (hdk_comm_*).  This shim keeps caller code compiling unchanged: communicators are opaque
 * ints and the few control-plane calls map onto the process environment.
 * If a real MPI is installed, put its include directory first and this file is never seen.
 */
#ifndef HYPREDRV_B200_MPI_SHIM_H
#define HYPREDRV_B200_MPI_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_SUCCESS 0
#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
int    MPI_Init(int *argc, char ***argv);
int    MPI_Initialized(int *flag);
int    MPI_Finalize(void);
int    MPI_Comm_rank(MPI_Comm comm, int *rank);
int    MPI_Comm_size(MPI_Comm comm, int *size);
int    MPI_Barrier(MPI_Comm comm);
int    MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);
#ifdef __cplusplus
}
#endif
#endif
