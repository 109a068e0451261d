/* hd_ij.c -- host-side IJ matrix / vector containers behind the hypre interface shim
 * (include/HYPRE.h).  They only collect what the caller sets; HYPREDRV_LinearSystemSetMatrix /
 * SetRHS flatten them to CSR and upload to the GPU (hdk_csr_from_host), where the ParCSR
 * diag/offd split and the diagonal-first reorder of hypre's IJ assembly happen.
 * Reference call sites: examples/src/C_laplacian/laplacian.c:734-747, 895, 906, 913-914. */
#include "hd_internal.h"
#include <stdlib.h>
#include <string.h>

HYPRE_Int HYPRE_Initialize(void) { return hdk_init(-1); }
HYPRE_Int HYPRE_Finalize(void) { return 0; }
/* hypre's host/device switches: accepted (callers such as the reference's tests request host
 * execution for their own assembly); the solve path of this library always runs on the device */
HYPRE_Int HYPRE_SetMemoryLocation(HYPRE_MemoryLocation loc) { (void)loc; return 0; }
HYPRE_Int HYPRE_SetExecutionPolicy(HYPRE_ExecutionPolicy policy) { (void)policy; return 0; }

HYPRE_Int HYPRE_IJMatrixCreate(MPI_Comm comm, HYPRE_BigInt ilower, HYPRE_BigInt iupper, HYPRE_BigInt jlower,
                               HYPRE_BigInt jupper, HYPRE_IJMatrix *matrix)
{
   (void)comm;
   struct hypre_IJMatrix_struct *A = calloc(1, sizeof(*A));
   if (!A) return 1;
   A->magic = HD_IJMAT_MAGIC;
   A->ilower = ilower; A->iupper = iupper; A->jlower = jlower; A->jupper = jupper;
   A->nrows = iupper >= ilower ? (int64_t)(iupper - ilower + 1) : 0;
   *matrix = A;
   return 0;
}

static void ij_free_rows(struct hypre_IJMatrix_struct *A)
{
   if (A->row_cols)
      for (int64_t i = 0; i < A->nrows; i++) { free(A->row_cols[i]); free(A->row_vals[i]); }
   free(A->row_cols); free(A->row_vals); free(A->row_len); free(A->row_cap);
   A->row_cols = NULL; A->row_vals = NULL; A->row_len = NULL; A->row_cap = NULL;
}

HYPRE_Int HYPRE_IJMatrixDestroy(HYPRE_IJMatrix A)
{
   if (!A || A->magic != HD_IJMAT_MAGIC) return 1;
   ij_free_rows(A);
   free(A->indptr); free(A->cols); free(A->vals);
   A->magic = 0;
   free(A);
   return 0;
}

HYPRE_Int HYPRE_IJMatrixSetObjectType(HYPRE_IJMatrix A, HYPRE_Int type) { (void)A; return type == HYPRE_PARCSR ? 0 : 1; }

static int ij_alloc_rows(struct hypre_IJMatrix_struct *A)
{
   if (A->row_len) return 0;
   size_t n = (size_t)(A->nrows > 0 ? A->nrows : 1);
   A->row_len  = calloc(n, sizeof(int64_t));
   A->row_cap  = calloc(n, sizeof(int64_t));
   A->row_cols = calloc(n, sizeof(HYPRE_BigInt *));
   A->row_vals = calloc(n, sizeof(double *));
   return (A->row_len && A->row_cap && A->row_cols && A->row_vals) ? 0 : 1;
}

HYPRE_Int HYPRE_IJMatrixSetRowSizes(HYPRE_IJMatrix A, const HYPRE_Int *sizes)
{
   if (!A || ij_alloc_rows(A)) return 1;
   for (int64_t i = 0; i < A->nrows; i++)
      if (sizes[i] > A->row_cap[i])
      {
         A->row_cols[i] = realloc(A->row_cols[i], sizeof(HYPRE_BigInt) * (size_t)sizes[i]);
         A->row_vals[i] = realloc(A->row_vals[i], sizeof(double) * (size_t)sizes[i]);
         A->row_cap[i]  = sizes[i];
      }
   return 0;
}

HYPRE_Int HYPRE_IJMatrixSetDiagOffdSizes(HYPRE_IJMatrix A, const HYPRE_Int *diag, const HYPRE_Int *offd)
{
   if (!A || ij_alloc_rows(A)) return 1;
   for (int64_t i = 0; i < A->nrows; i++)
   {
      int64_t s = (int64_t)diag[i] + (offd ? offd[i] : 0);
      if (s > A->row_cap[i])
      {
         A->row_cols[i] = realloc(A->row_cols[i], sizeof(HYPRE_BigInt) * (size_t)s);
         A->row_vals[i] = realloc(A->row_vals[i], sizeof(double) * (size_t)s);
         A->row_cap[i]  = s;
      }
   }
   return 0;
}

HYPRE_Int HYPRE_IJMatrixInitialize(HYPRE_IJMatrix A)
{
   if (!A) return 1;
   A->assembled = 0;
   if (ij_alloc_rows(A)) return 1;
   for (int64_t i = 0; i < A->nrows; i++) A->row_len[i] = 0;
   return 0;
}
HYPRE_Int HYPRE_IJMatrixInitialize_v2(HYPRE_IJMatrix A, HYPRE_MemoryLocation loc) { (void)loc; return HYPRE_IJMatrixInitialize(A); }

static int ij_set(struct hypre_IJMatrix_struct *A, HYPRE_Int nrows, HYPRE_Int *ncols, const HYPRE_BigInt *rows,
                  const HYPRE_BigInt *cols, const double *values, int add)
{
   if (!A || ij_alloc_rows(A)) return 1;
   int64_t off = 0;
   A->assembled = 0;
   for (HYPRE_Int r = 0; r < nrows; r++)
   {
      int64_t i = (int64_t)(rows[r] - A->ilower);
      int     nc = ncols[r];
      if (i < 0 || i >= A->nrows)
      {
         /* hypre ignores SetValues to rows of other ranks, but stashes AddToValues contributions and
          * ships them in Assemble (the FEM assembly pattern).  This container has no such exchange:
          * an off-rank AddTo is an error, never a silently wrong operator */
         if (add)
         {
            hd_err_set(HYPREDRV_ERROR_HYPRE_INTERNAL);
            hd_err_msg("HYPRE_IJMatrixAddToValues: row %lld belongs to another rank; off-process contributions are not supported "
                       "(assemble each row on its owner)", (long long)rows[r]);
            return 1;
         }
         off += nc;
         continue;
      }
      for (int c = 0; c < nc; c++)
      {
         HYPRE_BigInt col = cols[off + c];
         double       v   = values[off + c];
         int64_t      k;
         for (k = 0; k < A->row_len[i]; k++)
            if (A->row_cols[i][k] == col) break;
         if (k < A->row_len[i]) { if (add) A->row_vals[i][k] += v; else A->row_vals[i][k] = v; continue; }
         if (A->row_len[i] == A->row_cap[i])
         {
            int64_t nc2    = A->row_cap[i] ? 2 * A->row_cap[i] : 8;
            A->row_cols[i] = realloc(A->row_cols[i], sizeof(HYPRE_BigInt) * (size_t)nc2);
            A->row_vals[i] = realloc(A->row_vals[i], sizeof(double) * (size_t)nc2);
            if (!A->row_cols[i] || !A->row_vals[i]) return 1;
            A->row_cap[i] = nc2;
         }
         A->row_cols[i][A->row_len[i]] = col;
         A->row_vals[i][A->row_len[i]] = v;
         A->row_len[i]++;
      }
      off += nc;
   }
   return 0;
}

HYPRE_Int HYPRE_IJMatrixSetValues(HYPRE_IJMatrix A, HYPRE_Int nrows, HYPRE_Int *ncols, const HYPRE_BigInt *rows,
                                  const HYPRE_BigInt *cols, const HYPRE_Complex *values)
{
   return ij_set(A, nrows, ncols, rows, cols, values, 0);
}
HYPRE_Int HYPRE_IJMatrixAddToValues(HYPRE_IJMatrix A, HYPRE_Int nrows, HYPRE_Int *ncols, const HYPRE_BigInt *rows,
                                    const HYPRE_BigInt *cols, const HYPRE_Complex *values)
{
   return ij_set(A, nrows, ncols, rows, cols, values, 1);
}

int hd_ij_matrix_flatten(struct hypre_IJMatrix_struct *A)
{
   if (A->assembled && A->indptr) return 0;
   if (ij_alloc_rows(A)) return 1;
   free(A->indptr); free(A->cols); free(A->vals);
   A->indptr = malloc(sizeof(int64_t) * ((size_t)A->nrows + 1));
   int64_t nnz = 0;
   for (int64_t i = 0; i < A->nrows; i++) { A->indptr[i] = nnz; nnz += A->row_len[i]; }
   A->indptr[A->nrows] = nnz;
   A->cols = malloc(sizeof(HYPRE_BigInt) * (size_t)(nnz > 0 ? nnz : 1));
   A->vals = malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
   if (!A->indptr || !A->cols || !A->vals) return 1;
   for (int64_t i = 0; i < A->nrows; i++)
   {
      memcpy(A->cols + A->indptr[i], A->row_cols[i], sizeof(HYPRE_BigInt) * (size_t)A->row_len[i]);
      memcpy(A->vals + A->indptr[i], A->row_vals[i], sizeof(double) * (size_t)A->row_len[i]);
   }
   A->assembled = 1;
   return 0;
}

HYPRE_Int HYPRE_IJMatrixAssemble(HYPRE_IJMatrix A) { return A ? hd_ij_matrix_flatten(A) : 1; }

HYPRE_Int HYPRE_IJMatrixGetLocalRange(HYPRE_IJMatrix A, HYPRE_BigInt *ilower, HYPRE_BigInt *iupper,
                                      HYPRE_BigInt *jlower, HYPRE_BigInt *jupper)
{
   if (!A) return 1;
   if (ilower) *ilower = A->ilower;
   if (iupper) *iupper = A->iupper;
   if (jlower) *jlower = A->jlower;
   if (jupper) *jupper = A->jupper;
   return 0;
}
HYPRE_Int HYPRE_IJMatrixGetObject(HYPRE_IJMatrix A, void **object) { *object = A; return 0; }
HYPRE_Int HYPRE_IJMatrixMigrate(HYPRE_IJMatrix A, HYPRE_MemoryLocation loc) { (void)A; (void)loc; return 0; }

HYPRE_Int HYPRE_IJVectorCreate(MPI_Comm comm, HYPRE_BigInt jlower, HYPRE_BigInt jupper, HYPRE_IJVector *vector)
{
   (void)comm;
   struct hypre_IJVector_struct *v = calloc(1, sizeof(*v));
   if (!v) return 1;
   v->magic = HD_IJVEC_MAGIC;
   v->jlower = jlower; v->jupper = jupper;
   v->n = jupper >= jlower ? (int64_t)(jupper - jlower + 1) : 0;
   *vector = v;
   return 0;
}
HYPRE_Int HYPRE_IJVectorDestroy(HYPRE_IJVector v)
{
   if (!v || v->magic != HD_IJVEC_MAGIC) return 1;
   free(v->data);
   v->magic = 0;
   free(v);
   return 0;
}
HYPRE_Int HYPRE_IJVectorSetObjectType(HYPRE_IJVector v, HYPRE_Int type) { (void)v; return type == HYPRE_PARCSR ? 0 : 1; }
HYPRE_Int HYPRE_IJVectorInitialize(HYPRE_IJVector v)
{
   if (!v) return 1;
   if (!v->data) v->data = calloc((size_t)(v->n > 0 ? v->n : 1), sizeof(double));
   else memset(v->data, 0, sizeof(double) * (size_t)v->n);
   return v->data ? 0 : 1;
}
HYPRE_Int HYPRE_IJVectorInitialize_v2(HYPRE_IJVector v, HYPRE_MemoryLocation loc) { (void)loc; return HYPRE_IJVectorInitialize(v); }

static int ijv_set(struct hypre_IJVector_struct *v, HYPRE_Int n, const HYPRE_BigInt *idx, const double *vals, int add)
{
   if (!v) return 1;
   if (!v->data && HYPRE_IJVectorInitialize(v)) return 1;
   for (HYPRE_Int k = 0; k < n; k++)
   {
      int64_t i = idx ? (int64_t)(idx[k] - v->jlower) : k;
      if (i < 0 || i >= v->n) continue;
      if (add) v->data[i] += vals[k]; else v->data[i] = vals[k];
   }
   return 0;
}
HYPRE_Int HYPRE_IJVectorSetValues(HYPRE_IJVector v, HYPRE_Int n, const HYPRE_BigInt *idx, const HYPRE_Complex *vals) { return ijv_set(v, n, idx, vals, 0); }
HYPRE_Int HYPRE_IJVectorAddToValues(HYPRE_IJVector v, HYPRE_Int n, const HYPRE_BigInt *idx, const HYPRE_Complex *vals) { return ijv_set(v, n, idx, vals, 1); }
HYPRE_Int HYPRE_IJVectorGetValues(HYPRE_IJVector v, HYPRE_Int n, const HYPRE_BigInt *idx, HYPRE_Complex *vals)
{
   if (!v || !v->data) return 1;
   for (HYPRE_Int k = 0; k < n; k++)
   {
      int64_t i = idx ? (int64_t)(idx[k] - v->jlower) : k;
      vals[k]   = (i >= 0 && i < v->n) ? v->data[i] : 0.0;
   }
   return 0;
}
HYPRE_Int HYPRE_IJVectorAssemble(HYPRE_IJVector v) { return v ? 0 : 1; }
HYPRE_Int HYPRE_IJVectorGetLocalRange(HYPRE_IJVector v, HYPRE_BigInt *jlower, HYPRE_BigInt *jupper)
{
   if (!v) return 1;
   if (jlower) *jlower = v->jlower;
   if (jupper) *jupper = v->jupper;
   return 0;
}
HYPRE_Int HYPRE_IJVectorGetObject(HYPRE_IJVector v, void **object) { *object = v; return 0; }
HYPRE_Int HYPRE_IJVectorMigrate(HYPRE_IJVector v, HYPRE_MemoryLocation loc) { (void)v; (void)loc; return 0; }
