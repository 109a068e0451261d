/* HYPRE.h -- the subset of hypre's public types and IJ interface that hypredrive callers use to
 * hand matrices and vectors to the HYPREDRV API (reference call sites:
 * examples/src/C_laplacian/laplacian.c:734-747, 895, 906, 913-914;
 * tests/test_setmatrix_from_csr.c:143-149).  Objects built through this interface are plain
 * host-side containers; HYPREDRV_LinearSystemSetMatrix / SetRHS upload them to the GPU.
 * Widths follow hypre's "mixedint" build: HYPRE_Int 32 bit, HYPRE_BigInt 64 bit, fp64 values.
 */
#ifndef HYPREDRV_B200_HYPRE_SHIM_H
#define HYPREDRV_B200_HYPRE_SHIM_H
#include <mpi.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int       HYPRE_Int;
typedef long long HYPRE_BigInt;
typedef double    HYPRE_Real;
typedef double    HYPRE_Complex;
typedef int       HYPRE_MemoryLocation;
typedef int       HYPRE_ExecutionPolicy;
#define HYPRE_MEMORY_HOST 0
#define HYPRE_MEMORY_DEVICE 1
#define HYPRE_EXEC_HOST 0
#define HYPRE_EXEC_DEVICE 1
#define HYPRE_PARCSR 5555
#define HYPRE_USING_GPU 1
#define HYPRE_USING_CUDA 1
#define HYPRE_MIXEDINT 1
#define HYPRE_RELEASE_NUMBER 30100

struct hypre_IJMatrix_struct;
struct hypre_IJVector_struct;
typedef struct hypre_IJMatrix_struct *HYPRE_IJMatrix;
typedef struct hypre_IJVector_struct *HYPRE_IJVector;
typedef struct hypre_Matrix_struct   *HYPRE_Matrix;
typedef struct hypre_Vector_struct   *HYPRE_Vector;
typedef struct hypre_IJMatrix_struct *HYPRE_ParCSRMatrix;
typedef struct hypre_IJVector_struct *HYPRE_ParVector;
typedef struct hypre_Solver_struct   *HYPRE_Solver;

HYPRE_Int HYPRE_Initialize(void);
HYPRE_Int HYPRE_Finalize(void);
/* accepted and recorded; the solve path of this library always executes on the device */
HYPRE_Int HYPRE_SetMemoryLocation(HYPRE_MemoryLocation loc);
HYPRE_Int HYPRE_SetExecutionPolicy(HYPRE_ExecutionPolicy policy);

HYPRE_Int HYPRE_IJMatrixCreate(MPI_Comm comm, HYPRE_BigInt ilower, HYPRE_BigInt iupper, HYPRE_BigInt jlower,
                               HYPRE_BigInt jupper, HYPRE_IJMatrix *matrix);
HYPRE_Int HYPRE_IJMatrixDestroy(HYPRE_IJMatrix matrix);
HYPRE_Int HYPRE_IJMatrixSetObjectType(HYPRE_IJMatrix matrix, HYPRE_Int type);
HYPRE_Int HYPRE_IJMatrixSetRowSizes(HYPRE_IJMatrix matrix, const HYPRE_Int *sizes);
HYPRE_Int HYPRE_IJMatrixSetDiagOffdSizes(HYPRE_IJMatrix matrix, const HYPRE_Int *diag, const HYPRE_Int *offd);
HYPRE_Int HYPRE_IJMatrixInitialize(HYPRE_IJMatrix matrix);
HYPRE_Int HYPRE_IJMatrixInitialize_v2(HYPRE_IJMatrix matrix, HYPRE_MemoryLocation loc);
HYPRE_Int HYPRE_IJMatrixSetValues(HYPRE_IJMatrix matrix, HYPRE_Int nrows, HYPRE_Int *ncols, const HYPRE_BigInt *rows,
                                  const HYPRE_BigInt *cols, const HYPRE_Complex *values);
HYPRE_Int HYPRE_IJMatrixAddToValues(HYPRE_IJMatrix matrix, HYPRE_Int nrows, HYPRE_Int *ncols, const HYPRE_BigInt *rows,
                                    const HYPRE_BigInt *cols, const HYPRE_Complex *values);
HYPRE_Int HYPRE_IJMatrixAssemble(HYPRE_IJMatrix matrix);
HYPRE_Int HYPRE_IJMatrixGetLocalRange(HYPRE_IJMatrix matrix, HYPRE_BigInt *ilower, HYPRE_BigInt *iupper,
                                      HYPRE_BigInt *jlower, HYPRE_BigInt *jupper);
HYPRE_Int HYPRE_IJMatrixGetObject(HYPRE_IJMatrix matrix, void **object);
HYPRE_Int HYPRE_IJMatrixMigrate(HYPRE_IJMatrix matrix, HYPRE_MemoryLocation loc);

HYPRE_Int HYPRE_IJVectorCreate(MPI_Comm comm, HYPRE_BigInt jlower, HYPRE_BigInt jupper, HYPRE_IJVector *vector);
HYPRE_Int HYPRE_IJVectorDestroy(HYPRE_IJVector vector);
HYPRE_Int HYPRE_IJVectorSetObjectType(HYPRE_IJVector vector, HYPRE_Int type);
HYPRE_Int HYPRE_IJVectorInitialize(HYPRE_IJVector vector);
HYPRE_Int HYPRE_IJVectorInitialize_v2(HYPRE_IJVector vector, HYPRE_MemoryLocation loc);
HYPRE_Int HYPRE_IJVectorSetValues(HYPRE_IJVector vector, HYPRE_Int nvalues, const HYPRE_BigInt *indices,
                                  const HYPRE_Complex *values);
HYPRE_Int HYPRE_IJVectorAddToValues(HYPRE_IJVector vector, HYPRE_Int nvalues, const HYPRE_BigInt *indices,
                                    const HYPRE_Complex *values);
HYPRE_Int HYPRE_IJVectorGetValues(HYPRE_IJVector vector, HYPRE_Int nvalues, const HYPRE_BigInt *indices,
                                  HYPRE_Complex *values);
HYPRE_Int HYPRE_IJVectorAssemble(HYPRE_IJVector vector);
HYPRE_Int HYPRE_IJVectorGetLocalRange(HYPRE_IJVector vector, HYPRE_BigInt *jlower, HYPRE_BigInt *jupper);
HYPRE_Int HYPRE_IJVectorGetObject(HYPRE_IJVector vector, void **object);
HYPRE_Int HYPRE_IJVectorMigrate(HYPRE_IJVector vector, HYPRE_MemoryLocation loc);
#ifdef __cplusplus
}
#endif
#endif
