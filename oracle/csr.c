/* csr.c -- CSR containers, matvec and vector kernels of the oracle (TEST INFRASTRUCTURE).
 * Restates hypre's seq_mv/csr_matvec.c, seq_mv/csr_matop.c (transpose, reorder) and
 * utilities/random.c as used through src/internal/linsys.c:3030-3032, 2875 of the reference.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_set_threads(int n)
{
#ifdef _OPENMP
   if (n > 0) omp_set_num_threads(n);
   return omp_get_max_threads();
#else
   (void)n;
   return 1;
#endif
}

/* HYPRE_ParVectorSetRandomValues on one rank (hypre seq_mv/vector.c: hypre_SeqVectorSetRandomValues):
 * hypre_SeedRand(seed); x[i] = 2 hypre_Rand() - 1.  Used by rhs_mode random / randsol and
 * init_guess_mode random with seed 2023 (reference src/internal/linsys.c:1810-1838, 2034-2040). */
void ovec_set_random(int seed, int n, double *x)
{
   oracle_rand_stream(seed, n, x);
   for (int i = 0; i < n; i++) x[i] = 2.0 * x[i] - 1.0;
}

ocsr *ocsr_alloc(int nrows, int ncols, int64_t nnz, int with_values)
{
   ocsr *A  = (ocsr *)calloc(1, sizeof(ocsr));
   A->nrows = nrows;
   A->ncols = ncols;
   A->ia    = (int *)calloc((size_t)nrows + 1, sizeof(int));
   A->ja    = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
   A->a     = with_values ? (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)) : NULL;
   return A;
}

void ocsr_free(ocsr *A)
{
   if (!A) return;
   free(A->ia);
   free(A->ja);
   free(A->a);
   free(A);
}

ocsr *ocsr_from_arrays(int nrows, int ncols, const int *ia, const int *ja, const double *a)
{
   int64_t nnz = ia[nrows] - ia[0];
   ocsr   *A   = ocsr_alloc(nrows, ncols, nnz, 1);
   for (int i = 0; i <= nrows; i++) A->ia[i] = ia[i] - ia[0];
   memcpy(A->ja, ja + ia[0], sizeof(int) * (size_t)nnz);
   memcpy(A->a, a + ia[0], sizeof(double) * (size_t)nnz);
   return A;
}

/* hypre IJMatrixAssemble / hypre_CSRMatrixReorder: the diagonal entry is SWAPPED into the
 * first slot of its row (the displaced entry takes the diagonal's old slot). */
void ocsr_diag_first(ocsr *A)
{
   if (A->nrows != A->ncols) return;
   for (int i = 0; i < A->nrows; i++)
   {
      int j0 = A->ia[i];
      for (int j = j0; j < A->ia[i + 1]; j++)
      {
         if (A->ja[j] == i)
         {
            if (j != j0)
            {
               int    tj = A->ja[j0];
               double tv = A->a[j0];
               A->ja[j0] = i;
               A->a[j0]  = A->a[j];
               A->ja[j]  = tj;
               A->a[j]   = tv;
            }
            break;
         }
      }
   }
}

/* hypre_CSRMatrixTranspose: counting transpose; entries of each row of A^T appear in
 * increasing original-row order. */
ocsr *ocsr_transpose(const ocsr *A)
{
   int64_t nnz = A->ia[A->nrows];
   ocsr   *T   = ocsr_alloc(A->ncols, A->nrows, nnz, A->a != NULL);
   for (int64_t k = 0; k < nnz; k++) T->ia[A->ja[k] + 1]++;
   for (int i = 0; i < A->ncols; i++) T->ia[i + 1] += T->ia[i];
   int *pos = (int *)malloc(sizeof(int) * ((size_t)A->ncols + 1));
   memcpy(pos, T->ia, sizeof(int) * ((size_t)A->ncols + 1));
   for (int i = 0; i < A->nrows; i++)
      for (int k = A->ia[i]; k < A->ia[i + 1]; k++)
      {
         int p    = pos[A->ja[k]]++;
         T->ja[p] = i;
         if (A->a) T->a[p] = A->a[k];
      }
   free(pos);
   return T;
}

/* y = alpha*A*x + beta*y ; row sums are accumulated sequentially in CSR order with separately
 * rounded multiply and add (compile with -ffp-contract=off) -- the device SpMV reproduces
 * this order so alpha=1,beta=0 results are bit-identical. */
void ocsr_matvec(double alpha, const ocsr *A, const double *x, double beta, double *y)
{
#pragma omp parallel for schedule(static)
   for (int i = 0; i < A->nrows; i++)
   {
      double s = 0.0;
      for (int k = A->ia[i]; k < A->ia[i + 1]; k++) s += A->a[k] * x[A->ja[k]];
      if (beta == 0.0)
         y[i] = alpha * s;
      else
         y[i] = alpha * s + beta * y[i];
   }
}

/* r = b - A x */
void ocsr_residual(const ocsr *A, const double *x, const double *b, double *r)
{
#pragma omp parallel for schedule(static)
   for (int i = 0; i < A->nrows; i++)
   {
      double s = 0.0;
      for (int k = A->ia[i]; k < A->ia[i + 1]; k++) s += A->a[k] * x[A->ja[k]];
      r[i] = b[i] - s;
   }
}

/* Deterministic blocked inner product (independent of the OpenMP thread count). */
double ovec_dot(int n, const double *x, const double *y)
{
   const int BS   = 4096;
   int       nb   = (n + BS - 1) / BS;
   double   *part = (double *)malloc(sizeof(double) * (size_t)(nb > 0 ? nb : 1));
#pragma omp parallel for schedule(static)
   for (int b = 0; b < nb; b++)
   {
      int    lo = b * BS, hi = lo + BS < n ? lo + BS : n;
      double s = 0.0;
      for (int i = lo; i < hi; i++) s += x[i] * y[i];
      part[b] = s;
   }
   double s = 0.0;
   for (int b = 0; b < nb; b++) s += part[b];
   free(part);
   return s;
}

/* hypre utilities/random.c: Park-Miller minimal standard generator via Schrage's method.
 * hypre_SeedRand(seed); out[i] = hypre_Rand() for i = 0..n-1. */
void oracle_rand_stream(int seed, int n, double *out)
{
   const int a = 16807, m = 2147483647, q = 127773, r = 2836;
   int       s = seed;
   if (s < 1) s = 1;
   else if (s >= m) s = m - 1;
   for (int i = 0; i < n; i++)
   {
      int high = s / q, low = s % q;
      int test = a * low - r * high;
      s        = test > 0 ? test : test + m;
      out[i]   = (double)s / m;
   }
}
