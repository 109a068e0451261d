// hdk_vec.cu -- Krylov vector layer: fused axpy / dot / norm single-pass kernels with
// warp-shuffle block reductions (north-star item 3).  Stands in for hypre_ParVectorAxpy,
// hypre_ParVectorInnerProd, hypre_ParVectorScale, hypre_ParVectorCopy as driven from
// hypre_PCGSolve / hypre_GMRESSolve (reference call sites src/internal/solver.c:614,
// src/internal/linsys.c:2875).
#include "hdk_internal.cuh"
#include <math.h>

namespace hdk {

constexpr int VT = 256;

static inline int vec_grid(int64_t n)
{
   int64_t want = (n + (int64_t)VT * 4 - 1) / ((int64_t)VT * 4);
   int64_t cap  = (int64_t)g.sm_count * 8;
   if (want < 1) want = 1;
   return (int)(want < cap ? want : cap);
}

__global__ void __launch_bounds__(VT) k_fill(double *x, double v, int64_t n)
{
   for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) x[i] = v;
}

__global__ void __launch_bounds__(VT) k_copy(double *__restrict__ d, const double *__restrict__ s, int64_t n)
{
   for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) d[i] = s[i];
}

__global__ void __launch_bounds__(VT) k_axpy(double a, const double *__restrict__ x, double *__restrict__ y, int64_t n)
{
   for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
      y[i] = __dadd_rn(y[i], __dmul_rn(a, x[i]));
}

__global__ void __launch_bounds__(VT) k_scale(double a, double *x, int64_t n)
{
   for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) x[i] *= a;
}

// u = (w*f)/d : an l1-Jacobi sweep from a zero initial guess (bit-identical to the general
// sweep with u_old = 0)
__global__ void __launch_bounds__(VT) k_scaled_div(double *__restrict__ u, const double *__restrict__ f,
                                                   const double *__restrict__ d, double w, int64_t n, const HaloExport ex)
{
   int nexp = 0;
   for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
   {
      double dd = d[i];
      double v  = (dd != 0.0) ? __ddiv_rn(__dmul_rn(w, f[i]), dd) : 0.0;
      u[i]      = v;
      if (ex.seq) nexp += export_row(ex, (int)i, v);
   }
   export_finish(ex, nexp);
}

// kind: 0 dot(x,y), 1 sum|x|, 2 max|x| (max uses the same tree with fmax)
template <int KIND>
__global__ void __launch_bounds__(VT) k_reduce(const double *__restrict__ x, const double *__restrict__ y,
                                               int64_t n, double *partials, unsigned *ticket, int fin,
                                               double *out, double *scal)
{
   __shared__ double sm[VT / 32];
   __shared__ int    flag;
   double            acc = 0.0;
   for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
   {
      if (KIND == 0) acc += x[i] * y[i];
      else if (KIND == 1) acc += fabs(x[i]);
      else acc = fmax(acc, fabs(x[i]));
   }
   if (KIND == 2)
   {
      // max-reduction: reuse the sum tree on a per-block basis through shared memory
      int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
      for (int o = 16; o > 0; o >>= 1) acc = fmax(acc, __shfl_down_sync(0xffffffffu, acc, o));
      if (lane == 0) sm[w] = acc;
      __syncthreads();
      if (w == 0)
      {
         acc = lane < VT / 32 ? sm[lane] : 0.0;
         for (int o = 16; o > 0; o >>= 1) acc = fmax(acc, __shfl_down_sync(0xffffffffu, acc, o));
      }
      if (threadIdx.x == 0)
      {
         partials[blockIdx.x] = acc;
         __threadfence();
         unsigned t = atomicInc(ticket, gridDim.x - 1);
         if (t == gridDim.x - 1)
         {
            __threadfence();
            double m = 0.0;
            for (unsigned b = 0; b < gridDim.x; b++) m = fmax(m, __ldcg(partials + b));
            out[0] = m;
         }
      }
      return;
   }
   double bs = block_sum<VT>(acc, sm);
   __syncthreads();
   grid_finish<VT>(bs, partials, ticket, fin, out, scal, sm, &flag);
}

// PCG: x += alpha p ; r -= alpha s ; i_prod = <r,r>   (one pass, 48 bytes per row)
// PREFILL: also z0 = (w r)/d, the first l1-Jacobi sweep of the coming V-cycle from a zero guess
// (the expression of k_scaled_div), 64 bytes per row.  Two rows per iteration, loads first.
template <bool PREFILL>
__device__ __forceinline__ double pcg_xr_row(double alpha, double xi, double pi, double ri, double si, double di, double zw,
                                             double *x, double *r, double *z0, int64_t i)
{
   x[i]            = __dadd_rn(xi, __dmul_rn(alpha, pi));
   const double rn = __dadd_rn(ri, -__dmul_rn(alpha, si));
   r[i]            = rn;
   if (PREFILL) z0[i] = (di != 0.0) ? __ddiv_rn(__dmul_rn(zw, rn), di) : 0.0;
   return rn * rn;
}

// 128-bit accesses: each thread owns two consecutive rows per trip (ld/st.global.v2.f64) and runs
// two trips per loop iteration with all loads issued before the first store, i.e. 4 rows x 5 input
// streams = 20 independent 16-byte loads in flight per thread.  `n2` = number of row pairs; an odd
// last row is handled by the scalar tail of block 0.
template <bool PREFILL>
__global__ void __launch_bounds__(VT) k_pcg_xr(double *x, double *r, const double *p, const double *s, int64_t n,
                                               double *partials, unsigned *ticket, double *scal, int fin, double *fin_out,
                                               double *z0, const double *zd, double zw, const HaloExport ex)
{
   __shared__ double sm[VT / 32];
   __shared__ int    flag;
   const double      alpha = scal[S_ALPHA];
   double            acc   = 0.0;
   const int64_t     n2 = n >> 1, stride = (int64_t)gridDim.x * VT;
   int nexp = 0;
   double2          *x2 = reinterpret_cast<double2 *>(x), *r2 = reinterpret_cast<double2 *>(r);
   const double2    *p2 = reinterpret_cast<const double2 *>(p), *s2 = reinterpret_cast<const double2 *>(s);
   double2          *z2 = reinterpret_cast<double2 *>(z0);
   const double2    *d2 = reinterpret_cast<const double2 *>(zd);
   int64_t           i = blockIdx.x * (int64_t)VT + threadIdx.x;
   for (; i + stride < n2; i += 2 * stride)
   {
      const int64_t j = i + stride;
      const double2 xi = x2[i], pi = p2[i], ri = r2[i], si = s2[i];
      const double2 xj = x2[j], pj = p2[j], rj = r2[j], sj = s2[j];
      double2       di = make_double2(0.0, 0.0), dj = di;
      if (PREFILL) { di = d2[i]; dj = d2[j]; }
      double2 xo, ro, zo;
      xo.x = __dadd_rn(xi.x, __dmul_rn(alpha, pi.x)); xo.y = __dadd_rn(xi.y, __dmul_rn(alpha, pi.y));
      ro.x = __dadd_rn(ri.x, -__dmul_rn(alpha, si.x)); ro.y = __dadd_rn(ri.y, -__dmul_rn(alpha, si.y));
      x2[i] = xo; r2[i] = ro;
      if (PREFILL)
      {
         zo.x = (di.x != 0.0) ? __ddiv_rn(__dmul_rn(zw, ro.x), di.x) : 0.0;
         zo.y = (di.y != 0.0) ? __ddiv_rn(__dmul_rn(zw, ro.y), di.y) : 0.0;
         z2[i] = zo;
         if (ex.seq) { nexp += export_row(ex, (int)(2 * i), zo.x); nexp += export_row(ex, (int)(2 * i + 1), zo.y); }
      }
      acc += ro.x * ro.x; acc += ro.y * ro.y;
      xo.x = __dadd_rn(xj.x, __dmul_rn(alpha, pj.x)); xo.y = __dadd_rn(xj.y, __dmul_rn(alpha, pj.y));
      ro.x = __dadd_rn(rj.x, -__dmul_rn(alpha, sj.x)); ro.y = __dadd_rn(rj.y, -__dmul_rn(alpha, sj.y));
      x2[j] = xo; r2[j] = ro;
      if (PREFILL)
      {
         zo.x = (dj.x != 0.0) ? __ddiv_rn(__dmul_rn(zw, ro.x), dj.x) : 0.0;
         zo.y = (dj.y != 0.0) ? __ddiv_rn(__dmul_rn(zw, ro.y), dj.y) : 0.0;
         z2[j] = zo;
         if (ex.seq) { nexp += export_row(ex, (int)(2 * j), zo.x); nexp += export_row(ex, (int)(2 * j + 1), zo.y); }
      }
      acc += ro.x * ro.x; acc += ro.y * ro.y;
   }
   if (i < n2)
   {
      const double2 xi = x2[i], pi = p2[i], ri = r2[i], si = s2[i];
      double2       di = make_double2(0.0, 0.0);
      if (PREFILL) di = d2[i];
      double2 xo, ro, zo;
      xo.x = __dadd_rn(xi.x, __dmul_rn(alpha, pi.x)); xo.y = __dadd_rn(xi.y, __dmul_rn(alpha, pi.y));
      ro.x = __dadd_rn(ri.x, -__dmul_rn(alpha, si.x)); ro.y = __dadd_rn(ri.y, -__dmul_rn(alpha, si.y));
      x2[i] = xo; r2[i] = ro;
      if (PREFILL)
      {
         zo.x = (di.x != 0.0) ? __ddiv_rn(__dmul_rn(zw, ro.x), di.x) : 0.0;
         zo.y = (di.y != 0.0) ? __ddiv_rn(__dmul_rn(zw, ro.y), di.y) : 0.0;
         z2[i] = zo;
         if (ex.seq) { nexp += export_row(ex, (int)(2 * i), zo.x); nexp += export_row(ex, (int)(2 * i + 1), zo.y); }
      }
      acc += ro.x * ro.x; acc += ro.y * ro.y;
   }
   if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
   {
      const int64_t k  = n - 1;
      const double  dk = PREFILL ? zd[k] : 0.0;
      acc += pcg_xr_row<PREFILL>(alpha, x[k], p[k], r[k], s[k], dk, zw, x, r, z0, k);
      if (PREFILL && ex.seq) nexp += export_row(ex, (int)k, z0[k]);
   }
   if (PREFILL) export_finish(ex, nexp);
   double bs = block_sum<VT>(acc, sm);
   __syncthreads();
   grid_finish<VT>(bs, partials, ticket, fin, fin_out, scal, sm, &flag);
}

// PCG: p = z + beta p   (128-bit accesses, two row pairs per trip)
__global__ void __launch_bounds__(VT) k_pcg_p(double *__restrict__ p, const double *__restrict__ z,
                                              int64_t n, const double *__restrict__ scal, const HaloExport ex)
{
   const double   beta = scal[S_BETA];
   const int64_t  n2 = n >> 1, stride = (int64_t)gridDim.x * VT;
   int nexp = 0;
   double2       *p2 = reinterpret_cast<double2 *>(p);
   const double2 *z2 = reinterpret_cast<const double2 *>(z);
   int64_t        i = blockIdx.x * (int64_t)VT + threadIdx.x;
   for (; i + stride < n2; i += 2 * stride)
   {
      const int64_t j = i + stride;
      const double2 pi = p2[i], zi = z2[i], pj = p2[j], zj = z2[j];
      const double2 qi = make_double2(__dadd_rn(zi.x, __dmul_rn(beta, pi.x)), __dadd_rn(zi.y, __dmul_rn(beta, pi.y)));
      const double2 qj = make_double2(__dadd_rn(zj.x, __dmul_rn(beta, pj.x)), __dadd_rn(zj.y, __dmul_rn(beta, pj.y)));
      p2[i] = qi; p2[j] = qj;
      if (ex.seq)
      {
         nexp += export_row(ex, (int)(2 * i), qi.x); nexp += export_row(ex, (int)(2 * i + 1), qi.y);
         nexp += export_row(ex, (int)(2 * j), qj.x); nexp += export_row(ex, (int)(2 * j + 1), qj.y);
      }
   }
   if (i < n2)
   {
      const double2 pi = p2[i], zi = z2[i];
      const double2 qi = make_double2(__dadd_rn(zi.x, __dmul_rn(beta, pi.x)), __dadd_rn(zi.y, __dmul_rn(beta, pi.y)));
      p2[i] = qi;
      if (ex.seq) { nexp += export_row(ex, (int)(2 * i), qi.x); nexp += export_row(ex, (int)(2 * i + 1), qi.y); }
   }
   if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
   {
      const double q = __dadd_rn(z[n - 1], __dmul_rn(beta, p[n - 1]));
      p[n - 1] = q;
      if (ex.seq) nexp += export_row(ex, (int)(n - 1), q);
   }
   export_finish(ex, nexp);
}

// scalar forms for operands that are not 16-byte aligned (sub-vectors at odd offsets)
template <bool PREFILL>
__global__ void __launch_bounds__(VT) k_pcg_xr_scalar(double *x, double *r, const double *p, const double *s, int64_t n,
                                                      double *partials, unsigned *ticket, double *scal, int fin, double *fin_out,
                                                      double *z0, const double *zd, double zw)
{
   __shared__ double sm[VT / 32];
   __shared__ int    flag;
   const double      alpha = scal[S_ALPHA];
   double            acc   = 0.0;
   for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
   {
      const double di = PREFILL ? zd[i] : 0.0;
      acc += pcg_xr_row<PREFILL>(alpha, x[i], p[i], r[i], s[i], di, zw, x, r, z0, i);
   }
   double bs = block_sum<VT>(acc, sm);
   __syncthreads();
   grid_finish<VT>(bs, partials, ticket, fin, fin_out, scal, sm, &flag);
}
__global__ void __launch_bounds__(VT) k_pcg_p_scalar(double *__restrict__ p, const double *__restrict__ z,
                                                     int64_t n, const double *__restrict__ scal)
{
   const double beta = scal[S_BETA];
   for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
      p[i] = __dadd_rn(z[i], __dmul_rn(beta, p[i]));
}

// z = r and <r,r> (identity preconditioner)
__global__ void __launch_bounds__(VT) k_copy_dot(double *__restrict__ z, const double *__restrict__ r,
                                                 int64_t n, double *partials, unsigned *ticket, int fin,
                                                 double *out, double *scal)
{
   __shared__ double sm[VT / 32];
   __shared__ int    flag;
   double            acc = 0.0;
   for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
   {
      double v = r[i];
      z[i]     = v;
      acc += v * v;
   }
   double bs = block_sum<VT>(acc, sm);
   __syncthreads();
   grid_finish<VT>(bs, partials, ticket, fin, out, scal, sm, &flag);
}

// HYPRE_ParVectorSetRandomValues(v, seed) on one rank: hypre_SeedRand(seed), then x_i = 2 hypre_Rand() - 1
// in index order (hypre seq_mv/vector.c, utilities/random.c -- the Park-Miller generator).  Element i is
// reached by skip-ahead, s_i = seed * 16807^(i+1) mod (2^31 - 1), with i the GLOBAL index, so the vector
// does not depend on the partition (hypre itself re-seeds every rank with seed * (rank + 1)).
__device__ __forceinline__ uint32_t pm_mulmod31(uint32_t a, uint32_t b)
{
   const uint64_t M = 2147483647ull;
   uint64_t p = (uint64_t)a * b;
   uint64_t r = (p & M) + (p >> 31);
   r          = (r & M) + (r >> 31);
   if (r >= M) r -= M;
   return (uint32_t)r;
}
__global__ void __launch_bounds__(VT) k_random(double *x, int64_t n, int64_t off, uint32_t seed0)
{
   for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
   {
      uint64_t e = (uint64_t)(off + i) + 1;
      uint32_t base = 16807u, acc = seed0;
      while (e)
      {
         if (e & 1) acc = pm_mulmod31(acc, base);
         base = pm_mulmod31(base, base);
         e >>= 1;
      }
      x[i] = __dadd_rn(__dmul_rn(2.0, __ddiv_rn((double)acc, 2147483647.0)), -1.0);
   }
}

int vec_fill(double *x, double v, int64_t n)
{
   if (n <= 0) return HDK_OK;
   k_fill<<<vec_grid(n), VT, 0, g.stream>>>(x, v, n);
   HDK_LAUNCH_CHECK();
   return HDK_OK;
}
int vec_copy(double *d, const double *s, int64_t n)
{
   if (n <= 0 || d == s) return HDK_OK;
   k_copy<<<vec_grid(n), VT, 0, g.stream>>>(d, s, n);
   HDK_LAUNCH_CHECK();
   return HDK_OK;
}
int vec_axpy(double a, const double *x, double *y, int64_t n)
{
   if (n <= 0) return HDK_OK;
   k_axpy<<<vec_grid(n), VT, 0, g.stream>>>(a, x, y, n);
   HDK_LAUNCH_CHECK();
   return HDK_OK;
}
int vec_scale(double a, double *x, int64_t n)
{
   if (n <= 0) return HDK_OK;
   k_scale<<<vec_grid(n), VT, 0, g.stream>>>(a, x, n);
   HDK_LAUNCH_CHECK();
   return HDK_OK;
}
int vec_scaled_div(double *u, const double *f, const double *d, double w, int64_t n, const hdk_csr_s *export_to)
{
   if (n <= 0) return HDK_OK;
   HaloExport ex;
   if (export_to) halo_export_begin(*export_to, &ex);
   k_scaled_div<<<vec_grid(n), VT, 0, g.stream>>>(u, f, d, w, n, ex);
   HDK_LAUNCH_CHECK();
   return HDK_OK;
}
int vec_dot_dev(const double *x, const double *y, int64_t n, int fin, double *out_d)
{
   k_reduce<0><<<vec_grid(n), VT, 0, g.stream>>>(x, y, n, g.partials, g.counters, fin, out_d, g.dscal);
   HDK_LAUNCH_CHECK();
   return HDK_OK;
}
int vec_copy_dot(double *z, const double *r, int64_t n, int fin, double *out_d)
{
   k_copy_dot<<<vec_grid(n), VT, 0, g.stream>>>(z, r, n, g.partials, g.counters, fin, out_d, g.dscal);
   HDK_LAUNCH_CHECK();
   return HDK_OK;
}
static inline bool aligned16(const void *a, const void *b = nullptr, const void *c = nullptr, const void *d = nullptr,
                             const void *e = nullptr, const void *f = nullptr)
{
   return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d | (uintptr_t)e | (uintptr_t)f) & 15u) == 0;
}

int pcg_update_xr(double *x, double *r, const double *p, const double *s, int64_t n, double *scal, int fin, double *fin_out,
                  double *z0, const double *zd, double zw, const hdk_csr_s *export_z0_to)
{
   HaloExport ex;
   if (z0 && export_z0_to && aligned16(x, r, p, s, z0, zd)) halo_export_begin(*export_z0_to, &ex);
   if (!aligned16(x, r, p, s, z0, zd))
   {
      if (z0) k_pcg_xr_scalar<true><<<vec_grid(n), VT, 0, g.stream>>>(x, r, p, s, n, g.partials, g.counters, scal, fin, fin_out, z0, zd, zw);
      else k_pcg_xr_scalar<false><<<vec_grid(n), VT, 0, g.stream>>>(x, r, p, s, n, g.partials, g.counters, scal, fin, fin_out, z0, zd, zw);
      HDK_LAUNCH_CHECK();
      return HDK_OK;
   }
   if (z0) k_pcg_xr<true><<<vec_grid(n), VT, 0, g.stream>>>(x, r, p, s, n, g.partials, g.counters, scal, fin, fin_out, z0, zd, zw, ex);
   else k_pcg_xr<false><<<vec_grid(n), VT, 0, g.stream>>>(x, r, p, s, n, g.partials, g.counters, scal, fin, fin_out, z0, zd, zw, ex);
   HDK_LAUNCH_CHECK();
   return HDK_OK;
}
int pcg_update_p(double *p, const double *z, int64_t n, const double *scal, const hdk_csr_s *export_to)
{
   if (n <= 0) return HDK_OK;
   HaloExport ex;
   if (export_to && aligned16(p, z)) halo_export_begin(*export_to, &ex);
   if (!aligned16(p, z)) k_pcg_p_scalar<<<vec_grid(n), VT, 0, g.stream>>>(p, z, n, scal);
   else k_pcg_p<<<vec_grid(n), VT, 0, g.stream>>>(p, z, n, scal, ex);
   HDK_LAUNCH_CHECK();
   return HDK_OK;
}

// device-resident partial result -> global result on the host (sum over ranks through NCCL)
int vec_dot_host(const double *x, const double *y, int64_t n, double *out_h)
{
   HDK_TRY(vec_dot_dev(x, y, n, FIN_STORE, g.dscal + S_TMP0));
   HDK_TRY(allreduce_dev(g.dscal + S_TMP0, 1));
   HDK_CUDA(cudaMemcpyAsync(g.hscal + S_TMP0, g.dscal + S_TMP0, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   *out_h = g.hscal[S_TMP0];
   return HDK_OK;
}

// wait budget / error flag of this translation unit's copy of the flag-wait globals
int wait_globals_vec(long long tmo, int *err)
{
   return wait_globals_set(tmo, err) == cudaSuccess ? HDK_OK : set_error(HDK_ERR_CUDA, "cannot set the wait budget");
}

} // namespace hdk

using namespace hdk;

extern "C" {

int hdk_vec_alloc(int64_t n, double **x_d)
{
   HDK_TRY(require_init());
   HDK_TRY(dalloc(x_d, (size_t)(n > 0 ? n : 1) + 8));
   return vec_fill(*x_d, 0.0, n + 8);
}
int hdk_vec_free(double *x_d) { return hdk_free_device(x_d); }
int hdk_host_alloc(size_t bytes, void **p_h)
{
   HDK_TRY(require_init());
   if (!p_h) return set_error(HDK_ERR_INVALID, "hdk_host_alloc: null output");
   HDK_CUDA(cudaHostAlloc(p_h, bytes ? bytes : 8, cudaHostAllocDefault));
   return HDK_OK;
}
int hdk_host_free(void *p_h)
{
   if (p_h) HDK_CUDA(cudaFreeHost(p_h));
   return HDK_OK;
}
int hdk_vec_h2d(double *x_d, const double *x_h, int64_t n) { return hdk_copy_h2d(x_d, x_h, sizeof(double) * (size_t)n); }
int hdk_vec_d2h(double *x_h, const double *x_d, int64_t n) { return hdk_copy_d2h(x_h, x_d, sizeof(double) * (size_t)n); }
int hdk_vec_fill(double *x_d, double v, int64_t n) { HDK_TRY(require_init()); return vec_fill(x_d, v, n); }
int hdk_vec_copy(double *d, const double *s, int64_t n) { HDK_TRY(require_init()); return vec_copy(d, s, n); }
int hdk_vec_axpy(double a, const double *x, double *y, int64_t n) { HDK_TRY(require_init()); return vec_axpy(a, x, y, n); }
int hdk_vec_scale(double a, double *x, int64_t n) { HDK_TRY(require_init()); return vec_scale(a, x, n); }
int hdk_vec_dot(const double *x, const double *y, int64_t n, double *r) { HDK_TRY(require_init()); return vec_dot_host(x, y, n, r); }

int hdk_vec_norm(const double *x_d, int64_t n, int kind, double *result_h)
{
   HDK_TRY(require_init());
   if (kind == 1)
   {
      double d;
      HDK_TRY(vec_dot_host(x_d, x_d, n, &d));
      *result_h = sqrt(d);
      return HDK_OK;
   }
   if (kind == 0)
      k_reduce<1><<<vec_grid(n), VT, 0, g.stream>>>(x_d, x_d, n, g.partials, g.counters, FIN_STORE, g.dscal + S_TMP0, g.dscal);
   else
      k_reduce<2><<<vec_grid(n), VT, 0, g.stream>>>(x_d, x_d, n, g.partials, g.counters, FIN_STORE, g.dscal + S_TMP0, g.dscal);
   HDK_LAUNCH_CHECK();
   if (kind == 0) HDK_TRY(allreduce_dev(g.dscal + S_TMP0, 1));   // L1: sum over ranks
   else HDK_TRY(allreduce_max_dev(g.dscal + S_TMP0, 1));        // Linf: max over ranks
   HDK_CUDA(cudaMemcpyAsync(g.hscal + S_TMP0, g.dscal + S_TMP0, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   *result_h = g.hscal[S_TMP0];
   return HDK_OK;
}

int hdk_vec_random(double *x_d, int64_t n, int64_t off, int seed)
{
   HDK_TRY(require_init());
   if (n <= 0) return HDK_OK;
   const uint32_t s0 = seed < 1 ? 1u : (seed >= 2147483647 ? 2147483646u : (uint32_t)seed); // hypre_SeedRand's clamp
   k_random<<<vec_grid(n), VT, 0, g.stream>>>(x_d, n, off, s0);
   HDK_LAUNCH_CHECK();
   return HDK_OK;
}

} // extern "C"
