/* amg_solve.c -- CPU restatement of the BoomerAMG solve phase (TEST INFRASTRUCTURE ONLY).
 * Reference trigger: HYPRE_BoomerAMGSolve through PreconSolveDispatch
 * (src/internal/solver.c:314-329) and HYPREDRV_PreconApply (src/HYPREDRV.c:3345).
 * Restates hypre parcsr_ls/par_relax.c (types 0,3,4,6,7,9,11,12,13,14,18),
 * par_relax_more.c (two-stage GS), par_gauss_elim.c / hypre_gselim and par_cycle.c (V-cycle).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* hypre_gselim: Gaussian elimination without pivoting on a dense row-major copy. */
static void gselim(double *A, double *x, int n)
{
   if (n == 1) { if (A[0] != 0.0) x[0] /= A[0]; return; }
   for (int k = 0; k < n - 1; k++)
      if (A[(size_t)k * n + k] != 0.0)
         for (int j = k + 1; j < n; j++)
            if (A[(size_t)j * n + k] != 0.0)
            {
               double factor = A[(size_t)j * n + k] / A[(size_t)k * n + k];
               for (int m = k + 1; m < n; m++) A[(size_t)j * n + m] -= factor * A[(size_t)k * n + m];
               x[j] -= factor * x[k];
            }
   for (int k = n - 1; k > 0; --k)
   {
      if (A[(size_t)k * n + k] != 0.0)
      {
         x[k] /= A[(size_t)k * n + k];
         for (int j = 0; j < k; j++)
            if (A[(size_t)j * n + k] != 0.0) x[j] -= x[k] * A[(size_t)j * n + k];
      }
   }
   if (A[0] != 0.0) x[0] /= A[0];
}

/* One relaxation sweep on all points (relax_points = 0), one rank.
 *  18: l1-Jacobi  u += w (f - A u_old)/l1          (hypre_BoomerAMGRelax18WeightedL1Jacobi)
 *   7: Jacobi     u += w (f - A u_old)/a_ii        (hypre_BoomerAMGRelax7Jacobi)
 *   0: weighted Jacobi (same update as 7)
 *  13/14, 3/4: forward / backward Gauss-Seidel (hybrid variants reduce to GS on one rank;
 *             13/14 divide by the truncated l1 norm which equals a_ii on one rank)
 *   6: symmetric GS (forward then backward)
 *  11/12: two-stage Gauss-Seidel with 1 / 2 inner Jacobi iterations
 *   9: direct solve by Gaussian elimination
 * Residual rows are accumulated as res = f_i; res -= a_ij*u_j in CSR order. */
void oamg_relax(const ocsr *A, const double *f, double *u, int type, double weight,
                const double *l1, double *tmp, double *tmp2)
{
   int n = A->nrows;
   if (type == 18 || type == 7 || type == 0)
   {
      memcpy(tmp, u, sizeof(double) * (size_t)n);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; i++)
      {
         double d = (type == 18) ? l1[i] : A->a[A->ia[i]];
         if (d != 0.0)
         {
            double res = f[i];
            for (int k = A->ia[i]; k < A->ia[i + 1]; k++) res -= A->a[k] * tmp[A->ja[k]];
            u[i] += (weight * res) / d;
         }
      }
   }
   else if (type == 13 || type == 3 || type == 14 || type == 4 || type == 6)
   {
      int fwd = (type == 13 || type == 3 || type == 6);
      int bwd = (type == 14 || type == 4 || type == 6);
      if (fwd)
         for (int i = 0; i < n; i++)
         {
            double d = (type == 13) ? l1[i] : A->a[A->ia[i]];
            if (d != 0.0)
            {
               double res = f[i];
               for (int k = A->ia[i]; k < A->ia[i + 1]; k++) res -= A->a[k] * u[A->ja[k]];
               u[i] += res / d;
            }
         }
      if (bwd)
         for (int i = n - 1; i >= 0; i--)
         {
            double d = (type == 14) ? l1[i] : A->a[A->ia[i]];
            if (d != 0.0)
            {
               double res = f[i];
               for (int k = A->ia[i]; k < A->ia[i + 1]; k++) res -= A->a[k] * u[A->ja[k]];
               u[i] += res / d;
            }
         }
   }
   else if (type == 11 || type == 12)
   {
      /* hypre_BoomerAMGRelaxTwoStageGaussSeidelHost: r = w(f - A u); r <- D^{-1} r; u += r;
       * then inner: r <- D^{-1} L r (in place, bottom to top), u += (-1)^k r */
      int     inner = (type == 11) ? 1 : 2;
      double *r     = tmp;
      double  mult  = 1.0;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; i++)
      {
         double res = f[i];
         for (int k = A->ia[i]; k < A->ia[i + 1]; k++) res -= A->a[k] * u[A->ja[k]];
         r[i] = (weight * res) / A->a[A->ia[i]];
      }
      for (int i = 0; i < n; i++) u[i] += r[i];
      for (int it = 0; it < inner; it++)
      {
         for (int i = n - 1; i >= 0; i--)
         {
            double res = 0.0;
            for (int k = A->ia[i]; k < A->ia[i + 1]; k++)
               if (A->ja[k] < i) res += A->a[k] * r[A->ja[k]];
            r[i] = res / A->a[A->ia[i]];
         }
         mult = -mult;
         for (int i = 0; i < n; i++) u[i] += mult * r[i];
      }
   }
   (void)tmp2;
}

static void coarse_solve(oamg *h, int l)
{
   int n = h->A[l]->nrows;
   if (h->ge && h->ge_n == n && (h->prm.relax_coarse == 9 || h->prm.relax_coarse == 99))
   {
      double *M = (double *)malloc(sizeof(double) * ((size_t)n * n + 1));
      memcpy(M, h->ge, sizeof(double) * (size_t)n * n);
      memcpy(h->u[l], h->f[l], sizeof(double) * (size_t)n);
      gselim(M, h->u[l], n);
      free(M);
   }
   else
   {
      for (int s = 0; s < h->prm.sweeps_coarse; s++)
         oamg_relax(h->A[l], h->f[l], h->u[l], h->prm.relax_coarse, h->prm.relax_weight,
                    h->l1_down[l], h->t[l], h->t2[l]);
   }
}

/* hypre_BoomerAMGCycle, cycle_type 1 (V), relax_order 0. */
void oamg_vcycle(oamg *h, const double *f, double *u)
{
   int n0 = h->A[0]->nrows, L = h->nlev - 1;
   memcpy(h->f[0], f, sizeof(double) * (size_t)n0);
   memcpy(h->u[0], u, sizeof(double) * (size_t)n0);
   for (int l = 0; l < L; l++)
   {
      const ocsr *A = h->A[l];
      for (int s = 0; s < h->prm.sweeps_down; s++)
         oamg_relax(A, h->f[l], h->u[l], h->prm.relax_down, h->prm.relax_weight, h->l1_down[l],
                    h->t[l], h->t2[l]);
      /* Vtemp = f - A u ; F_c = P^T Vtemp ; U_c = 0 */
      ocsr_residual(A, h->u[l], h->f[l], h->t[l]);
      ocsr_matvec(1.0, h->R[l], h->t[l], 0.0, h->f[l + 1]);
      memset(h->u[l + 1], 0, sizeof(double) * (size_t)h->A[l + 1]->nrows);
   }
   coarse_solve(h, L);
   for (int l = L - 1; l >= 0; l--)
   {
      const ocsr *P = h->P[l];
      double     *uf = h->u[l], *uc = h->u[l + 1];
#pragma omp parallel for schedule(static)
      for (int i = 0; i < P->nrows; i++)
      {
         double s = uf[i];
         for (int k = P->ia[i]; k < P->ia[i + 1]; k++) s += P->a[k] * uc[P->ja[k]];
         uf[i] = s;
      }
      for (int s = 0; s < h->prm.sweeps_up; s++)
         oamg_relax(h->A[l], h->f[l], h->u[l], h->prm.relax_up, h->prm.relax_weight, h->l1_up[l],
                    h->t[l], h->t2[l]);
   }
   memcpy(u, h->u[0], sizeof(double) * (size_t)n0);
}

/* z = M^{-1} r : one V-cycle from a zero initial guess (max_iter 1, tol 0;
 * src/internal/amg.c:224-226; hypre's PCG clears z before calling the preconditioner). */
void oamg_precond(oamg *h, const double *r, double *z)
{
   memset(z, 0, sizeof(double) * (size_t)h->A[0]->nrows);
   oamg_vcycle(h, r, z);
}
