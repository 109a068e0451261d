// hdk_krylov.cu -- PCG and right-preconditioned GMRES(k) drivers on the device.
// Stands in for HYPRE_ParCSRPCGSolve / HYPRE_ParCSRGMRESSolve (hypre krylov/pcg.c,
// krylov/gmres.c) as called from src/internal/solver.c:211, 223, 614 of the reference, with the
// option sets of src/internal/pcg.c:15-25 (two_norm 1, rel_change 0, stop_crit 0) and
// src/internal/gmres.c:16-27.  All vector work and every scalar recurrence stay on the GPU:
// the host only reads one small scalar block per iteration to take the stopping decision, and
// that read overlaps the preconditioner application that hypre performs before its test.
#include "hdk_internal.cuh"
#include "hdk_amg.cuh"
#include <math.h>

namespace hdk {

// with one rank the last block of the producing kernel applies `fin`; with several ranks the
// local partial sits in S_TMP0 and one mailbox kernel sums it over NVLink (rank order, identical on
// every rank) and applies the recurrence
static int finish_dot(int fin, double *out)
{
   if (g.nranks <= 1) return HDK_OK;
   return allreduce_fin_dev(g.dscal + S_TMP0, 1, fin, out);
}
static inline int local_fin(int fin) { return g.nranks <= 1 ? fin : FIN_STORE; }
static inline double *local_out(double *out) { return g.nranks <= 1 ? out : g.dscal + S_TMP0; }

// z = M^{-1} r, fused <r,z> -> fin
static int precond_dot(hdk_amg_s *M, const double *r, double *z, int64_t n, int fin, double *out)
{
   if (M)
   {
      HDK_TRY(amg_precond(M, r, z, local_fin(fin), local_out(out)));
   }
   else
   {
      HDK_TRY(vec_copy_dot(z, r, n, local_fin(fin), local_out(out)));
   }
   return finish_dot(fin, out);
}

static int read_scalars(int count)
{
   HDK_CUDA(cudaMemcpyAsync(g.hscal, g.dscal, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaEventRecord(g.ev_scal, g.stream));
   return HDK_OK;
}

// Modified Gram-Schmidt step in one pass: y -= h x (h on the device) and <z, y_new> for the next
// coefficient (z = the next basis vector, or y itself for the final norm): 4 vector accesses per
// basis vector instead of the 5 of a separate dot + axpy
__global__ void __launch_bounds__(256) k_axpy_dot(const double *coef, const double *x, double *y, const double *z, int64_t n,
                                                  double *partials, unsigned *ticket, double *out, double *scal)
{
   __shared__ double sm[256 / 32];
   __shared__ int    flag;
   const double      a = -coef[0];
   const bool        self = (z == y);
   double            acc = 0.0;
   for (int64_t i = blockIdx.x * (int64_t)256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
   {
      const double yn = __dadd_rn(y[i], __dmul_rn(a, x[i]));
      y[i]            = yn;
      acc += (self ? yn : z[i]) * yn;
   }
   double bs = block_sum<256>(acc, sm);
   __syncthreads();
   grid_finish<256>(bs, partials, ticket, FIN_STORE, out, scal, sm, &flag);
}
static int axpy_dot_dev(const double *coef, const double *x, double *y, const double *z, int64_t n, double *out)
{
   int64_t want = (n + 1023) / 1024, cap = (int64_t)g.sm_count * 8;
   int     grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
   k_axpy_dot<<<grid, 256, 0, g.stream>>>(coef, x, y, z, n, g.partials, g.counters, out, g.dscal);
   HDK_LAUNCH_CHECK();
   return HDK_OK;
}

} // namespace hdk

using namespace hdk;

extern "C" {

int hdk_pcg(const hdk_csr *A, hdk_amg *M, const double *b, double *x, hdk_krylov *k)
{
   HDK_TRY(require_init());
   if (!A || !b || !x || !k) return set_error(HDK_ERR_INVALID, "hdk_pcg: null argument");
   const int64_t n = A->diag.nrows;
   k->iters = 0; k->converged = 0; k->rel_res_norm = 0.0; k->solve_ms = 0.0;
   double *r, *p, *s;
   HDK_TRY(dalloc(&r, (size_t)n + 8));
   HDK_TRY(dalloc(&p, (size_t)n + 8));
   HDK_TRY(dalloc(&s, (size_t)n + 8));
   double *S = g.dscal;
   HDK_CUDA(cudaEventRecord(g.ev_a, g.stream));

   double bi_prod;
   HDK_TRY(vec_dot_host(b, b, n, &bi_prod));
   double eps = k->rel_tol * k->rel_tol;
   int    rc  = HDK_OK;
   int    i   = 0;
   double i_prod = 0.0;
   if (bi_prod > 0.0)
   {
      if (k->abs_tol > 0 && k->abs_tol * k->abs_tol / bi_prod > eps) eps = k->abs_tol * k->abs_tol / bi_prod;
   }
   else
   {
      // zero right-hand side: x = 0 (hypre_PCGSolve)
      rc = vec_fill(x, 0.0, n);
      k->converged = 1;
      goto done;
   }
   {
      // r = b - A x ; p = M^{-1} r ; gamma = <r,p> ; i_prod = <r,r>
      SpmvArgs a;
      a.x = x; a.y = r; a.b = b;
      if ((rc = parcsr_matvec(*A, SPMV_RESIDUAL, a))) goto done;
      if ((rc = precond_dot(M, r, p, n, FIN_STORE, S + S_GAMMA))) goto done;
   }
   while (i + 1 <= k->max_iter)
   {
      i++;
      // s = A p ; sdotp = <s,p> ; alpha = gamma / sdotp   (one kernel)
      SpmvArgs a;
      a.x = p; a.y = s; a.dotv = p; a.fin = local_fin(FIN_SDOTP); a.fin_out = local_out(nullptr);
      tl_mark(0, 7);
      if ((rc = parcsr_matvec(*A, SPMV_SET, a))) goto done;
      if ((rc = finish_dot(FIN_SDOTP, nullptr))) goto done;
      // x += alpha p ; r -= alpha s ; i_prod = <r,r>       (one kernel)
      //   + the first smoothing sweep of the coming V-cycle, z0 = (w r)/l1, written where the cycle expects it
      {
         double       *zb = nullptr;
         const double *zd = nullptr;
         double        zw = 1.0;
         const hdk_csr_s *zreader = nullptr;
         const bool    pf = M && amg_prefill_target(M, s, &zb, &zd, &zw, &zreader);
         tl_mark(0, 8);
         if ((rc = pcg_update_xr(x, r, p, s, n, S, local_fin(FIN_IPROD), local_out(nullptr), pf ? zb : nullptr, zd, zw, pf ? zreader : nullptr))) goto done;
         if (pf) M->prefilled_at = 0;
      }
      if ((rc = finish_dot(FIN_IPROD, nullptr))) goto done;
      if ((rc = read_scalars(8))) goto done;
      // s = M^{-1} r ; gamma_new = <r,s> ; beta = gamma_new / gamma  (V-cycle; the host
      // decision below overlaps it -- hypre also preconditions before testing)
      if ((rc = precond_dot(M, r, s, n, FIN_GAMMA, nullptr))) goto done;
      cudaError_t e = cudaEventSynchronize(g.ev_scal);
      if (e != cudaSuccess) { rc = set_error(HDK_ERR_CUDA, "event sync: %s", cudaGetErrorString(e)); goto done; }
      i_prod = g.hscal[S_IPROD];
      if (g.hscal[S_SDOTP] == 0.0 || i_prod != i_prod) { break; }
      if (i_prod / bi_prod < eps) { k->converged = 1; break; }
      // p = s + beta p   (the next iteration's A p reads it: its halo is filled here)
      tl_mark(0, 9);
      if ((rc = pcg_update_p(p, s, n, S, i + 1 <= k->max_iter ? A : nullptr))) goto done;
   }
   k->iters        = i;
   k->rel_res_norm = sqrt(i_prod / bi_prod);
done:
   cudaEventRecord(g.ev_b, g.stream);
   cudaEventSynchronize(g.ev_b);
   float ms = 0.f;
   cudaEventElapsedTime(&ms, g.ev_a, g.ev_b);
   k->solve_ms = ms;
   dfree(r); dfree(p); dfree(s);
   if (rc == HDK_OK) rc = comm_check_error();
   return rc;
}

// flexible: hypre krylov/flexgmres.c -- keep z_j = M^{-1} p_j and form x += sum_j y_j z_j
static int gmres_impl(const hdk_csr *A, hdk_amg *M, const double *b, double *x, hdk_krylov *k, bool flexible)
{
   HDK_TRY(require_init());
   if (!A || !b || !x || !k) return set_error(HDK_ERR_INVALID, "hdk_gmres: null argument");
   const int64_t n  = A->diag.nrows;
   const int     kd = k->krylov_dim > 0 ? k->krylov_dim : 30;
   k->iters = 0; k->converged = 0; k->rel_res_norm = 0.0; k->solve_ms = 0.0;
   std::vector<double *> p((size_t)kd + 1, nullptr);
   double *r = nullptr, *w = nullptr;
   int     rc = HDK_OK;
   for (int j = 0; j <= kd; j++) if ((rc = dalloc(&p[(size_t)j], (size_t)n + 8))) return rc;
   std::vector<double *> z(flexible ? (size_t)kd + 1 : 0, nullptr);
   for (size_t j = 0; j < z.size(); j++) if ((rc = dalloc(&z[j], (size_t)n + 8))) return rc;
   HDK_TRY(dalloc(&r, (size_t)n + 8));
   HDK_TRY(dalloc(&w, (size_t)n + 8));
   std::vector<double> c((size_t)kd + 1, 0.0), s((size_t)kd + 1, 0.0), rs((size_t)kd + 2, 0.0);
   std::vector<double> hh((size_t)(kd + 2) * (kd + 1), 0.0);
#define HH(a_, b_) hh[(size_t)(a_) * (kd + 1) + (b_)]
   const double epsmac = 1.e-16;
   // Hessenberg column of the current inner iteration: kd + 2 coefficients on the device with a host
   // mirror (sized by krylov_dim, so any restart length works)
   double *H = nullptr;
   HDK_TRY(dalloc(&H, (size_t)kd + 2));
   std::vector<double> hcol((size_t)kd + 2, 0.0);
   HDK_CUDA(cudaEventRecord(g.ev_a, g.stream));
   double b_norm = 0, r_norm = 0, t = 0;
   int    iter = 0;
   {
      double d;
      if ((rc = vec_dot_host(b, b, n, &d))) goto done;
      b_norm = sqrt(d);
      SpmvArgs a;
      a.x = x; a.y = p[0]; a.b = b;
      if ((rc = parcsr_matvec(*A, SPMV_RESIDUAL, a))) goto done;
      if ((rc = vec_dot_host(p[0], p[0], n, &d))) goto done;
      r_norm = sqrt(d);
   }
   {
      double den = b_norm > 0.0 ? b_norm : r_norm;
      double eps = k->rel_tol * den;
      if (k->abs_tol > eps) eps = k->abs_tol;
      double real_old = r_norm;
      while (iter < k->max_iter)
      {
         rs[0] = r_norm;
         if (r_norm == 0.0) { k->converged = 1; break; }
         if (r_norm <= eps && iter >= k->min_iter)
         {
            SpmvArgs a;
            a.x = x; a.y = r; a.b = b;
            if ((rc = parcsr_matvec(*A, SPMV_RESIDUAL, a))) goto done;
            double d;
            if ((rc = vec_dot_host(r, r, n, &d))) goto done;
            r_norm = sqrt(d);
            if (r_norm <= eps) { k->converged = 1; break; }
         }
         if ((rc = vec_scale(1.0 / r_norm, p[0], n))) goto done;
         int i = 0;
         while (i < kd && iter < k->max_iter)
         {
            i++; iter++;
            // zz = M^{-1} p[i-1] ; p[i] = A zz
            double *zz = flexible ? z[(size_t)i - 1] : r;
            if (M) { if ((rc = amg_precond(M, p[(size_t)i - 1], zz, FIN_NONE, nullptr))) goto done; }
            else if ((rc = vec_copy(zz, p[(size_t)i - 1], n))) goto done;
            // modified Gram-Schmidt, coefficients on the device (S_H0 + j): h_0 = <p_0, A zz> comes out
            // of the SpMV kernel; step j subtracts h_j p_j and produces h_{j+1} (or <p_i,p_i>) in one pass
            SpmvArgs a;
            a.x = zz; a.y = p[(size_t)i];
            a.dotv = p[0]; a.fin = FIN_STORE; a.fin_out = H;
            if ((rc = parcsr_matvec(*A, SPMV_SET, a))) goto done;
            if ((rc = allreduce_dev(H, 1))) goto done;
            for (int j = 0; j < i; j++)
            {
               const double *znext = (j + 1 < i) ? p[(size_t)j + 1] : p[(size_t)i];
               if ((rc = axpy_dot_dev(H + j, p[(size_t)j], p[(size_t)i], znext, n, H + j + 1))) goto done;
               if ((rc = allreduce_dev(H + j + 1, 1))) goto done;
            }
            HDK_CUDA(cudaMemcpyAsync(hcol.data(), H, sizeof(double) * (size_t)(i + 1), cudaMemcpyDeviceToHost, g.stream));
            HDK_CUDA(cudaStreamSynchronize(g.stream));
            for (int j = 0; j < i; j++) HH(j, i - 1) = hcol[(size_t)j];
            t            = sqrt(hcol[(size_t)i]);
            HH(i, i - 1) = t;
            if (t != 0.0) { if ((rc = vec_scale(1.0 / t, p[(size_t)i], n))) goto done; }
            for (int j = 1; j < i; j++)
            {
               t                = HH(j - 1, i - 1);
               HH(j - 1, i - 1) = s[(size_t)j - 1] * HH(j, i - 1) + c[(size_t)j - 1] * t;
               HH(j, i - 1)     = -s[(size_t)j - 1] * t + c[(size_t)j - 1] * HH(j, i - 1);
            }
            t = HH(i, i - 1) * HH(i, i - 1);
            t += HH(i - 1, i - 1) * HH(i - 1, i - 1);
            double gamma = sqrt(t);
            if (gamma == 0.0) gamma = epsmac;
            c[(size_t)i - 1] = HH(i - 1, i - 1) / gamma;
            s[(size_t)i - 1] = HH(i, i - 1) / gamma;
            rs[(size_t)i]    = -HH(i, i - 1) * rs[(size_t)i - 1];
            rs[(size_t)i] /= gamma;
            rs[(size_t)i - 1] = c[(size_t)i - 1] * rs[(size_t)i - 1];
            HH(i - 1, i - 1)  = s[(size_t)i - 1] * HH(i, i - 1) + c[(size_t)i - 1] * HH(i - 1, i - 1);
            r_norm            = fabs(rs[(size_t)i]);
            if (r_norm <= eps && iter >= k->min_iter) break;
         }
         // solve the triangular system, w = sum rs_j p_j, x += M^{-1} w
         rs[(size_t)i - 1] = rs[(size_t)i - 1] / HH(i - 1, i - 1);
         for (int kk = i - 2; kk >= 0; kk--)
         {
            t = 0.0;
            for (int j = kk + 1; j < i; j++) t -= HH(kk, j) * rs[(size_t)j];
            t += rs[(size_t)kk];
            rs[(size_t)kk] = t / HH(kk, kk);
         }
         if (flexible)
         {
            for (int j = i - 1; j >= 0; j--) if ((rc = vec_axpy(rs[(size_t)j], z[(size_t)j], x, n))) goto done;
         }
         else
         {
            if ((rc = vec_copy(w, p[(size_t)i - 1], n))) goto done;
            if ((rc = vec_scale(rs[(size_t)i - 1], w, n))) goto done;
            for (int j = i - 2; j >= 0; j--) if ((rc = vec_axpy(rs[(size_t)j], p[(size_t)j], w, n))) goto done;
            if (M) { if ((rc = amg_precond(M, w, r, FIN_NONE, nullptr))) goto done; }
            else if ((rc = vec_copy(r, w, n))) goto done;
            if ((rc = vec_axpy(1.0, r, x, n))) goto done;
         }
         if (r_norm <= eps && iter >= k->min_iter)
         {
            if (k->skip_real_res_check) { k->converged = 1; break; }
            SpmvArgs a;
            a.x = x; a.y = r; a.b = b;
            if ((rc = parcsr_matvec(*A, SPMV_RESIDUAL, a))) goto done;
            double d;
            if ((rc = vec_dot_host(r, r, n, &d))) goto done;
            double real_new = r_norm = sqrt(d);
            if (r_norm <= eps) { k->converged = 1; break; }
            if (real_new >= real_old) { k->converged = 1; break; }
            if ((rc = vec_copy(p[0], r, n))) goto done;
            i        = 0;
            real_old = real_new;
         }
         // residual vector for the restart from the rotations
         for (int j = i; j > 0; j--)
         {
            rs[(size_t)j - 1] = -s[(size_t)j - 1] * rs[(size_t)j];
            rs[(size_t)j]     = c[(size_t)j - 1] * rs[(size_t)j];
         }
         if (i) if ((rc = vec_scale(rs[(size_t)i], p[(size_t)i], n))) goto done; // p_i += (rs_i - 1) p_i
         for (int j = i - 1; j > 0; j--) if ((rc = vec_axpy(rs[(size_t)j], p[(size_t)j], p[(size_t)i], n))) goto done;
         if (i)
         {
            if ((rc = vec_scale(rs[0], p[0], n))) goto done;
            if ((rc = vec_axpy(1.0, p[(size_t)i], p[0], n))) goto done;
         }
      }
      k->iters        = iter;
      k->rel_res_norm = b_norm > 0.0 ? r_norm / b_norm : r_norm;
   }
#undef HH
done:
   cudaEventRecord(g.ev_b, g.stream);
   cudaEventSynchronize(g.ev_b);
   float ms = 0.f;
   cudaEventElapsedTime(&ms, g.ev_a, g.ev_b);
   k->solve_ms = ms;
   for (auto q : p) dfree(q);
   for (auto q : z) dfree(q);
   dfree(r); dfree(w); dfree(H);
   if (rc == HDK_OK) rc = comm_check_error();
   return rc;
}

int hdk_gmres(const hdk_csr *A, hdk_amg *M, const double *b, double *x, hdk_krylov *k) { return gmres_impl(A, M, b, x, k, false); }
int hdk_fgmres(const hdk_csr *A, hdk_amg *M, const double *b, double *x, hdk_krylov *k) { return gmres_impl(A, M, b, x, k, true); }

// hypre krylov/bicgstab.c (hypre_BiCGSTABSolve), right-preconditioned (reference
// src/internal/solver.c:241-252, options src/internal/bicgstab.c)
int hdk_bicgstab(const hdk_csr *A, hdk_amg *M, const double *b, double *x, hdk_krylov *k)
{
   HDK_TRY(require_init());
   if (!A || !b || !x || !k) return set_error(HDK_ERR_INVALID, "hdk_bicgstab: null argument");
   const int64_t n = A->diag.nrows;
   const double  epsmac = 1.e-128;
   k->iters = 0; k->converged = 0; k->rel_res_norm = 0.0; k->solve_ms = 0.0;
   double *v[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   for (int j = 0; j < 6; j++) HDK_TRY(dalloc(&v[j], (size_t)n + 8));
   double *r0 = v[0], *r = v[1], *p = v[2], *vv = v[3], *q = v[4], *s = v[5];
   int     rc = HDK_OK, iter = 0;
   double  d = 0, b_norm = 0, r_norm = 0, res = 0;
   HDK_CUDA(cudaEventRecord(g.ev_a, g.stream));
   auto prec = [&](const double *in, double *out) -> int {
      if (M) return amg_precond(M, in, out, FIN_NONE, nullptr);
      return vec_copy(out, in, n);
   };
   auto mv = [&](const double *in, double *out) -> int { SpmvArgs a; a.x = in; a.y = out; return parcsr_matvec(*A, SPMV_SET, a); };
   do
   {
      if ((rc = vec_dot_host(b, b, n, &d))) break;
      b_norm = sqrt(d);
      SpmvArgs a;
      a.x = x; a.y = r0; a.b = b;
      if ((rc = parcsr_matvec(*A, SPMV_RESIDUAL, a))) break;
      if ((rc = vec_copy(r, r0, n)) || (rc = vec_copy(p, r0, n))) break;
      if ((rc = vec_dot_host(r, r, n, &d))) break;
      r_norm = sqrt(d);
      double den = b_norm > 0.0 ? b_norm : r_norm, eps = k->rel_tol * den;
      if (k->abs_tol > eps) eps = k->abs_tol;
      res = d; // <r0, r>
      while (iter < k->max_iter && res != 0.0)
      {
         if (r_norm == 0.0) { k->converged = 1; break; }
         iter++;
         double temp, gn, gd;
         if ((rc = prec(p, vv)) || (rc = mv(vv, q)) || (rc = vec_dot_host(r0, q, n, &temp))) break;
         if (fabs(temp) < epsmac) break;
         double alpha = res / temp;
         if ((rc = vec_axpy(alpha, vv, x, n)) || (rc = vec_axpy(-alpha, q, r, n))) break;
         if ((rc = prec(r, vv)) || (rc = mv(vv, s))) break;
         if ((rc = vec_dot_host(r, s, n, &gn)) || (rc = vec_dot_host(s, s, n, &gd))) break;
         double gamma = (gn == 0.0 && gd == 0.0) ? 0.0 : gn / gd;
         if ((rc = vec_axpy(gamma, vv, x, n)) || (rc = vec_axpy(-gamma, s, r, n))) break;
         if ((rc = vec_dot_host(r, r, n, &d))) break;
         r_norm = sqrt(d);
         if (r_norm <= eps && iter >= k->min_iter)
         {
            SpmvArgs t;
            t.x = x; t.y = r; t.b = b;
            if ((rc = parcsr_matvec(*A, SPMV_RESIDUAL, t)) || (rc = vec_dot_host(r, r, n, &d))) break;
            r_norm = sqrt(d);
            if (r_norm <= eps) { k->converged = 1; break; }
         }
         if (fabs(res) < epsmac) break;
         double beta = 1.0 / res;
         if ((rc = vec_dot_host(r0, r, n, &res))) break;
         beta *= res;
         if ((rc = vec_axpy(-gamma, q, p, n))) break;
         if (fabs(gamma) < epsmac) break;
         if ((rc = vec_scale(beta * alpha / gamma, p, n)) || (rc = vec_axpy(1.0, r, p, n))) break;
      }
   } while (0);
   k->iters        = iter;
   k->rel_res_norm = b_norm > 0.0 ? r_norm / b_norm : r_norm;
   cudaEventRecord(g.ev_b, g.stream);
   cudaEventSynchronize(g.ev_b);
   float ms = 0.f;
   cudaEventElapsedTime(&ms, g.ev_a, g.ev_b);
   k->solve_ms = ms;
   for (int j = 0; j < 6; j++) dfree(v[j]);
   if (rc == HDK_OK) rc = comm_check_error();
   return rc;
}

} // extern "C"
