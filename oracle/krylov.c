/* krylov.c -- CPU restatement of hypre's PCG and GMRES (TEST INFRASTRUCTURE ONLY).
 * Reference trigger: HYPRE_ParCSRPCGSolve / HYPRE_ParCSRGMRESSolve (src/internal/solver.c:211,
 * 223, 614) with the option sets of src/internal/pcg.c:15-25 (two_norm 1, rel_change 0,
 * stop_crit 0, recompute_res 0) and src/internal/gmres.c:16-27 (rel_change 0, min_iter 0).
 * Restates hypre krylov/pcg.c (hypre_PCGSolve) and krylov/gmres.c (hypre_GMRESSolve).
 * M == NULL means no preconditioner (identity copy).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void precond(oamg *M, int n, const double *r, double *z)
{
   if (M) oamg_precond(M, r, z);
   else memcpy(z, r, sizeof(double) * (size_t)n);
}

static void axpy(int n, double a, const double *x, double *y)
{
#pragma omp parallel for schedule(static)
   for (int i = 0; i < n; i++) y[i] += a * x[i];
}

int opcg(const ocsr *A, oamg *M, const double *b, double *x, okrylov *k)
{
   int     n = A->nrows;
   double *r = (double *)calloc((size_t)n + 1, sizeof(double));
   double *p = (double *)calloc((size_t)n + 1, sizeof(double));
   double *s = (double *)calloc((size_t)n + 1, sizeof(double));
   k->iters = 0; k->converged = 0; k->rel_res_norm = 0.0;

   double bi_prod = ovec_dot(n, b, b);
   double eps     = k->rel_tol * k->rel_tol;
   if (bi_prod > 0.0)
   {
      if (k->abs_tol > 0 && k->abs_tol * k->abs_tol / bi_prod > eps) eps = k->abs_tol * k->abs_tol / bi_prod;
   }
   else
   {
      /* zero right-hand side: x = 0 is the solution */
      memset(x, 0, sizeof(double) * (size_t)n);
      k->converged = 1;
      free(r); free(p); free(s);
      return 0;
   }
   ocsr_residual(A, x, b, r);
   precond(M, n, r, p);
   double gamma  = ovec_dot(n, r, p);
   double i_prod = ovec_dot(n, r, r);
   if (k->hist) k->hist[0] = sqrt(i_prod);
   int i = 0;
   while (i + 1 <= k->max_iter)
   {
      i++;
      ocsr_matvec(1.0, A, p, 0.0, s);
      double sdotp = ovec_dot(n, s, p);
      if (sdotp == 0.0) { i--; break; }
      double alpha     = gamma / sdotp;
      double gamma_old = gamma;
      axpy(n, alpha, p, x);
      axpy(n, -alpha, s, r);
      precond(M, n, r, s);
      gamma  = ovec_dot(n, r, s);
      i_prod = ovec_dot(n, r, r);
      if (k->hist) k->hist[i] = sqrt(i_prod);
      if (i_prod / bi_prod < eps) { k->converged = 1; break; }
      if (gamma < 1.0e-292 && -gamma < 1.0e-292) break;
      double beta = gamma / gamma_old;
#pragma omp parallel for schedule(static)
      for (int j = 0; j < n; j++) p[j] = s[j] + beta * p[j];
   }
   k->iters        = i;
   k->rel_res_norm = sqrt(i_prod / bi_prod);
   free(r); free(p); free(s);
   return 0;
}

/* flexible != 0: hypre krylov/flexgmres.c -- the preconditioned directions z_j = M^{-1} p_j are
 * stored and the correction is x += sum_j y_j z_j (no extra preconditioner application). */
static int gmres_impl(const ocsr *A, oamg *M, const double *b, double *x, okrylov *k, int flexible)
{
   int      n = A->nrows, kd = k->krylov_dim > 0 ? k->krylov_dim : 30;
   double  *r = (double *)calloc((size_t)n + 1, sizeof(double));
   double  *w = (double *)calloc((size_t)n + 1, sizeof(double));
   double **p = (double **)malloc(sizeof(double *) * (size_t)(kd + 1));
   for (int j = 0; j <= kd; j++) p[j] = (double *)calloc((size_t)n + 1, sizeof(double));
   double **z = NULL;
   if (flexible)
   {
      z = (double **)malloc(sizeof(double *) * (size_t)(kd + 1));
      for (int j = 0; j <= kd; j++) z[j] = (double *)calloc((size_t)n + 1, sizeof(double));
   }
   double  *c = (double *)calloc((size_t)kd + 1, sizeof(double));
   double  *s = (double *)calloc((size_t)kd + 1, sizeof(double));
   double  *rs = (double *)calloc((size_t)kd + 2, sizeof(double));
   double  *hh = (double *)calloc((size_t)(kd + 2) * (kd + 1), sizeof(double));
#define HH(a, b_) hh[(size_t)(a) * (kd + 1) + (b_)]
   const double epsmac = 1.e-16;
   k->iters = 0; k->converged = 0;

   double b_norm = sqrt(ovec_dot(n, b, b));
   ocsr_residual(A, x, b, p[0]);
   double r_norm = sqrt(ovec_dot(n, p[0], p[0]));
   double den    = b_norm > 0.0 ? b_norm : r_norm;
   double eps    = k->rel_tol * den;
   if (k->abs_tol > eps) eps = k->abs_tol;
   double real_r_norm_old = r_norm, real_r_norm_new;
   if (k->hist) k->hist[0] = r_norm;
   int iter = 0;
   while (iter < k->max_iter)
   {
      rs[0] = r_norm;
      if (r_norm == 0.0) { k->converged = 1; break; }
      if (r_norm <= eps)
      {
         ocsr_residual(A, x, b, r);
         r_norm = sqrt(ovec_dot(n, r, r));
         if (r_norm <= eps) { k->converged = 1; break; }
      }
      double t = 1.0 / r_norm;
      for (int j = 0; j < n; j++) p[0][j] *= t;
      int i = 0;
      while (i < kd && iter < k->max_iter)
      {
         i++; iter++;
         double *zz = flexible ? z[i - 1] : r;
         precond(M, n, p[i - 1], zz);
         ocsr_matvec(1.0, A, zz, 0.0, p[i]);
         for (int j = 0; j < i; j++)
         {
            HH(j, i - 1) = ovec_dot(n, p[j], p[i]);
            axpy(n, -HH(j, i - 1), p[j], p[i]);
         }
         t            = sqrt(ovec_dot(n, p[i], p[i]));
         HH(i, i - 1) = t;
         if (t != 0.0) { t = 1.0 / t; for (int j = 0; j < n; j++) p[i][j] *= t; }
         for (int j = 1; j < i; j++)
         {
            t                = HH(j - 1, i - 1);
            HH(j - 1, i - 1) = s[j - 1] * HH(j, i - 1) + c[j - 1] * t;
            HH(j, i - 1)     = -s[j - 1] * t + c[j - 1] * HH(j, i - 1);
         }
         t = HH(i, i - 1) * HH(i, i - 1);
         t += HH(i - 1, i - 1) * HH(i - 1, i - 1);
         double gamma = sqrt(t);
         if (gamma == 0.0) gamma = epsmac;
         c[i - 1] = HH(i - 1, i - 1) / gamma;
         s[i - 1] = HH(i, i - 1) / gamma;
         rs[i]    = -HH(i, i - 1) * rs[i - 1];
         rs[i] /= gamma;
         rs[i - 1]        = c[i - 1] * rs[i - 1];
         HH(i - 1, i - 1) = s[i - 1] * HH(i, i - 1) + c[i - 1] * HH(i - 1, i - 1);
         r_norm           = fabs(rs[i]);
         if (k->hist) k->hist[iter] = r_norm;
         if (r_norm <= eps) break;
      }
      /* solve the upper triangular system, form the correction */
      rs[i - 1] = rs[i - 1] / HH(i - 1, i - 1);
      for (int kk = i - 2; kk >= 0; kk--)
      {
         t = 0.0;
         for (int j = kk + 1; j < i; j++) t -= HH(kk, j) * rs[j];
         t += rs[kk];
         rs[kk] = t / HH(kk, kk);
      }
      if (flexible)
      {
         for (int j = i - 1; j >= 0; j--) axpy(n, rs[j], z[j], x);
      }
      else
      {
         for (int j = 0; j < n; j++) w[j] = rs[i - 1] * p[i - 1][j];
         for (int j = i - 2; j >= 0; j--) axpy(n, rs[j], p[j], w);
         precond(M, n, w, r);
         axpy(n, 1.0, r, x);
      }
      if (r_norm <= eps)
      {
         if (k->skip_real_res_check) { k->converged = 1; break; }
         ocsr_residual(A, x, b, r);
         real_r_norm_new = r_norm = sqrt(ovec_dot(n, r, r));
         if (r_norm <= eps) { k->converged = 1; break; }
         /* false convergence: restart from the true residual unless it stagnates */
         if (real_r_norm_new >= real_r_norm_old) { k->converged = 1; break; }
         memcpy(p[0], r, sizeof(double) * (size_t)n);
         i               = 0;
         real_r_norm_old = real_r_norm_new;
      }
      /* residual vector for the restart from the rotations */
      for (int j = i; j > 0; j--)
      {
         rs[j - 1] = -s[j - 1] * rs[j];
         rs[j]     = c[j - 1] * rs[j];
      }
      if (i) axpy(n, rs[i] - 1.0, p[i], p[i]);
      for (int j = i - 1; j > 0; j--) axpy(n, rs[j], p[j], p[i]);
      if (i)
      {
         axpy(n, rs[0] - 1.0, p[0], p[0]);
         axpy(n, 1.0, p[i], p[0]);
      }
   }
   k->iters        = iter;
   k->rel_res_norm = b_norm > 0.0 ? r_norm / b_norm : r_norm;
   for (int j = 0; j <= kd; j++) free(p[j]);
   if (z) { for (int j = 0; j <= kd; j++) free(z[j]); free(z); }
   free(p); free(c); free(s); free(rs); free(hh); free(r); free(w);
#undef HH
   return 0;
}

int ogmres(const ocsr *A, oamg *M, const double *b, double *x, okrylov *k) { return gmres_impl(A, M, b, x, k, 0); }
int ofgmres(const ocsr *A, oamg *M, const double *b, double *x, okrylov *k) { return gmres_impl(A, M, b, x, k, 1); }

/* hypre krylov/bicgstab.c (hypre_BiCGSTABSolve), right-preconditioned, stop_crit 0. */
int obicgstab(const ocsr *A, oamg *M, const double *b, double *x, okrylov *k)
{
   int     n = A->nrows;
   double *r0 = (double *)calloc((size_t)n + 1, sizeof(double)), *r = (double *)calloc((size_t)n + 1, sizeof(double));
   double *p = (double *)calloc((size_t)n + 1, sizeof(double)), *v = (double *)calloc((size_t)n + 1, sizeof(double));
   double *q = (double *)calloc((size_t)n + 1, sizeof(double)), *s = (double *)calloc((size_t)n + 1, sizeof(double));
   const double epsmac = 1.e-128;
   k->iters = 0; k->converged = 0;
   double b_norm = sqrt(ovec_dot(n, b, b));
   ocsr_residual(A, x, b, r0);
   memcpy(r, r0, sizeof(double) * (size_t)n);
   memcpy(p, r0, sizeof(double) * (size_t)n);
   double r_norm = sqrt(ovec_dot(n, r, r));
   double den = b_norm > 0.0 ? b_norm : r_norm;
   double eps = k->rel_tol * den;
   if (k->abs_tol > eps) eps = k->abs_tol;
   if (k->hist) k->hist[0] = r_norm;
   int    iter = 0;
   double res = ovec_dot(n, r0, r);
   while (iter < k->max_iter && res != 0.0)
   {
      if (r_norm == 0.0) { k->converged = 1; break; }
      iter++;
      precond(M, n, p, v);
      ocsr_matvec(1.0, A, v, 0.0, q);
      double temp = ovec_dot(n, r0, q);
      if (fabs(temp) < epsmac) break;
      double alpha = res / temp;
      axpy(n, alpha, v, x);
      axpy(n, -alpha, q, r);
      precond(M, n, r, v);
      ocsr_matvec(1.0, A, v, 0.0, s);
      double gn = ovec_dot(n, r, s), gd = ovec_dot(n, s, s);
      double gamma = (gn == 0.0 && gd == 0.0) ? 0.0 : gn / gd;
      axpy(n, gamma, v, x);
      axpy(n, -gamma, s, r);
      r_norm = sqrt(ovec_dot(n, r, r));
      if (k->hist) k->hist[iter] = r_norm;
      if (r_norm <= eps)
      {
         ocsr_residual(A, x, b, r);
         r_norm = sqrt(ovec_dot(n, r, r));
         if (r_norm <= eps) { k->converged = 1; break; }
      }
      if (fabs(res) < epsmac) break;
      double beta = 1.0 / res;
      res = ovec_dot(n, r0, r);
      beta *= res;
      axpy(n, -gamma, q, p);
      if (fabs(gamma) < epsmac) break;
      double sc = beta * alpha / gamma;
      for (int j = 0; j < n; j++) p[j] *= sc;
      axpy(n, 1.0, r, p);
   }
   k->iters = iter;
   k->rel_res_norm = b_norm > 0.0 ? r_norm / b_norm : r_norm;
   free(r0); free(r); free(p); free(v); free(q); free(s);
   return 0;
}
