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
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_PROC_NULL (-2)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_REQUEST_NULL 0
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_SUCCESS 0
#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_LONG_LONG_INT 3
#define MPI_LONG_LONG 3
#define MPI_CHAR 4
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
/* Cartesian topologies and point-to-point calls as the reference examples use them
 * (examples/src/C_laplacian/laplacian.c:547-600, 1666-1790).  With one process they are exact
 * (every neighbour is MPI_PROC_NULL); with several, ranks are one process per GPU and the example
 * drivers run one rank per process through the launcher -- the calls that would move data between
 * processes report an error instead of returning wrong values. */
int    MPI_Cart_create(MPI_Comm comm, int ndims, const int dims[], const int periods[], int reorder, MPI_Comm *cart);
int    MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int coords[]);
int    MPI_Cart_rank(MPI_Comm comm, const int coords[], int *rank);
int    MPI_Cart_shift(MPI_Comm comm, int direction, int disp, int *source, int *dest);
int    MPI_Comm_free(MPI_Comm *comm);
int    MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int    MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request *req);
int    MPI_Waitall(int count, MPI_Request reqs[], MPI_Status statuses[]);
int    MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
#ifdef __cplusplus
}
#endif
#endif
