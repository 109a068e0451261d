// hdk_spmv.cu -- the fused CSR SpMV family (north-star items 1 and 2):
//   y = A x | r = b - A x | l1-Jacobi sweep fused with its residual | y += A x | axpby
// each optionally fused with a dot product of the result (PCG <Ap,p>, <r,z>).
// Stands in for hypre_ParCSRMatrixMatvec / hypre_ParCSRMatrixMatvecOutOfPlace and
// hypre_BoomerAMGRelax (types 7, 18) -- reference call sites src/internal/linsys.c:1835, 3031,
// src/internal/solver.c:614 (every Krylov iteration) and the V-cycle behind solver.c:314-329.
//
// Kernel choice by measured row-length statistics (csr_analyze):
//  * STREAM (short rows, max_row <= 1024): each CTA owns ~3072 consecutive non-zeros.  All
//    256 threads stream col/val with 128-bit loads (perfectly coalesced regardless of row
//    boundaries), gather x, and park the products in shared memory; then one thread per row
//    adds its products sequentially in CSR order.  The per-row order equals the oracle's, so
//    results are bit-identical to the CPU restatement.
//  * VECTOR (long rows): one warp per row, 128-bit loads along the row, shuffle reduction.
#include "hdk_internal.cuh"

namespace hdk {

constexpr int ST     = 256;        // threads per CTA
constexpr int S_TGT  = 3072;       // target non-zeros per CTA
constexpr int S_CAP  = 4096;       // product slots in shared memory (32 KB)
constexpr int S_MAXR = S_CAP - S_TGT; // longest row the stream kernel accepts (1024)

struct SpmvDev
{
   const int    *rowptr, *col, *blk_row;
   const double *val;
   const double *x, *b, *d, *dotv;
   double       *y;
   double        w, alpha, beta;
   int           nrows, fin;
   double       *fin_out, *scal, *partials;
   unsigned     *ticket;
};

__device__ __forceinline__ int4 ld_int4(const int *p)
{
   return __ldg(reinterpret_cast<const int4 *>(p));
}
__device__ __forceinline__ double2 ld_double2(const double *p)
{
   return __ldg(reinterpret_cast<const double2 *>(p));
}

// per-row operands of the epilogue; the first PF rounds are prefetched before the CTA barrier
// so that phase 2 does not pay a second full DRAM round trip
struct RowOps
{
   int    s, e;
   double b, d, xo, yo, dv;
};

template <int MODE, bool DOT>
__device__ __forceinline__ void row_load(const SpmvDev &a, int r, int ka, RowOps &o)
{
   o.s = __ldg(a.rowptr + r) - ka;
   o.e = __ldg(a.rowptr + r + 1) - ka;
   if (MODE == SPMV_RESIDUAL || MODE == SPMV_JACOBI || MODE == SPMV_JACOBI_R) o.b = a.b[r];
   if (MODE == SPMV_JACOBI || MODE == SPMV_JACOBI_R) o.d = a.d[r];
   if (MODE == SPMV_JACOBI) o.xo = a.x[r];
   if (MODE == SPMV_ADD || MODE == SPMV_AXPBY) o.yo = a.y[r];
   if (DOT) o.dv = a.dotv[r];
}

template <int MODE>
__device__ __forceinline__ double row_epilogue(const SpmvDev &a, const RowOps &o, const double *prod)
{
   double acc;
   if (MODE == SPMV_SET || MODE == SPMV_AXPBY)
   {
      acc = 0.0;
      for (int k = o.s; k < o.e; k++) acc = __dadd_rn(acc, prod[k]);
      if (MODE == SPMV_AXPBY)
         acc = (a.beta == 0.0) ? __dmul_rn(a.alpha, acc)
                               : __dadd_rn(__dmul_rn(a.alpha, acc), __dmul_rn(a.beta, o.yo));
      return acc;
   }
   if (MODE == SPMV_ADD)
   {
      acc = o.yo;
      for (int k = o.s; k < o.e; k++) acc = __dadd_rn(acc, prod[k]);
      return acc;
   }
   // residual-type modes: res = b; res -= a_ij x_j in CSR order
   acc = o.b;
   for (int k = o.s; k < o.e; k++) acc = __dadd_rn(acc, -prod[k]);
   if (MODE == SPMV_RESIDUAL) return acc;
   if (MODE == SPMV_JACOBI)
      return (o.d != 0.0) ? __dadd_rn(o.xo, __ddiv_rn(__dmul_rn(a.w, acc), o.d)) : o.xo;
   /* SPMV_JACOBI_R */
   return (o.d != 0.0) ? __ddiv_rn(__dmul_rn(a.w, acc), o.d) : 0.0;
}

template <int MODE, bool DOT>
__global__ void __launch_bounds__(ST) k_spmv_stream(SpmvDev a)
{
   constexpr int NIT = S_CAP / (4 * ST); // phase-1 steps per thread
   constexpr int PF  = 2;                // prefetched phase-2 rounds
   __shared__ __align__(16) double prod[S_CAP];
   __shared__ double red[ST / 32];
   __shared__ int    flag;
   const int tid = threadIdx.x;
   const int r0 = a.blk_row[blockIdx.x], r1 = a.blk_row[blockIdx.x + 1];
   double    dacc = 0.0;
   if (r0 < r1)
   {
      const int k0 = __ldg(a.rowptr + r0), k1 = __ldg(a.rowptr + r1);
      const int ka = k0 & ~3;
      // phase 1: stream 4 non-zeros per thread per step (128-bit loads) ...
      int4    c[NIT];
      double2 v0[NIT], v1[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it)
      {
         int k = ka + (it * ST + tid) * 4;
         if (k < k1)
         {
            c[it]  = ld_int4(a.col + k);
            v0[it] = ld_double2(a.val + k);
            v1[it] = ld_double2(a.val + k + 2);
         }
      }
      // ... issue the per-row operand loads of phase 2 while those are in flight ...
      RowOps ro[PF];
#pragma unroll
      for (int j = 0; j < PF; ++j)
      {
         int r = r0 + tid + j * ST;
         if (r < r1) row_load<MODE, DOT>(a, r, ka, ro[j]);
      }
      // ... gather x and park the products in shared memory
#pragma unroll
      for (int it = 0; it < NIT; ++it)
      {
         int k = ka + (it * ST + tid) * 4;
         if (k < k1)
         {
            double x0 = __ldg(a.x + c[it].x), x1 = __ldg(a.x + c[it].y);
            double x2 = __ldg(a.x + c[it].z), x3 = __ldg(a.x + c[it].w);
            double2 p0, p1;
            p0.x = __dmul_rn(v0[it].x, x0); p0.y = __dmul_rn(v0[it].y, x1);
            p1.x = __dmul_rn(v1[it].x, x2); p1.y = __dmul_rn(v1[it].y, x3);
            *reinterpret_cast<double2 *>(prod + (k - ka))     = p0;
            *reinterpret_cast<double2 *>(prod + (k - ka) + 2) = p1;
         }
      }
      __syncthreads();
      // phase 2: one thread per row, sequential sum in CSR order
#pragma unroll
      for (int j = 0; j < PF; ++j)
      {
         int r = r0 + tid + j * ST;
         if (r < r1)
         {
            double yn = row_epilogue<MODE>(a, ro[j], prod);
            a.y[r]    = yn;
            if (DOT) dacc += ro[j].dv * yn;
         }
      }
      for (int r = r0 + tid + PF * ST; r < r1; r += ST)
      {
         RowOps o;
         row_load<MODE, DOT>(a, r, ka, o);
         double yn = row_epilogue<MODE>(a, o, prod);
         a.y[r]    = yn;
         if (DOT) dacc += o.dv * yn;
      }
   }
   if (DOT)
   {
      double bs = block_sum<ST>(dacc, red);
      __syncthreads();
      grid_finish<ST>(bs, a.partials, a.ticket, a.fin, a.fin_out, a.scal, red, &flag);
   }
}

// one warp per row; lanes stride the row with scalar loads (rows here are long, so each warp
// reads whole 128-byte lines).  Summation order differs from the oracle (tolerance parity).
template <int MODE, bool DOT>
__global__ void __launch_bounds__(ST) k_spmv_vector(SpmvDev a)
{
   __shared__ double red[ST / 32];
   __shared__ int    flag;
   const int lane = threadIdx.x & 31;
   const int wpb  = ST / 32;
   double    dacc = 0.0;
   for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < a.nrows; r += gridDim.x * wpb)
   {
      int    s = a.rowptr[r], e = a.rowptr[r + 1];
      double acc = 0.0;
      for (int k = s + lane; k < e; k += 32) acc += __ldg(a.val + k) * __ldg(a.x + __ldg(a.col + k));
      acc = warp_sum(acc);
      if (lane == 0)
      {
         double yn;
         if (MODE == SPMV_SET) yn = acc;
         else if (MODE == SPMV_AXPBY) yn = (a.beta == 0.0) ? a.alpha * acc : a.alpha * acc + a.beta * a.y[r];
         else if (MODE == SPMV_ADD) yn = a.y[r] + acc;
         else
         {
            double res = a.b[r] - acc;
            if (MODE == SPMV_RESIDUAL) yn = res;
            else
            {
               double dd = a.d[r];
               if (MODE == SPMV_JACOBI) yn = (dd != 0.0) ? a.x[r] + (a.w * res) / dd : a.x[r];
               else yn = (dd != 0.0) ? (a.w * res) / dd : 0.0;
            }
         }
         a.y[r] = yn;
         if (DOT) dacc += a.dotv[r] * yn;
      }
   }
   if (DOT)
   {
      double bs = block_sum<ST>(dacc, red);
      __syncthreads();
      grid_finish<ST>(bs, a.partials, a.ticket, a.fin, a.fin_out, a.scal, red, &flag);
   }
}

template <int MODE>
static int launch_mode(const DevCSR &A, const SpmvDev &d, bool dot)
{
   if (A.kind == 0)
   {
      if (dot) k_spmv_stream<MODE, true><<<A.nblk, ST, 0, g.stream>>>(d);
      else k_spmv_stream<MODE, false><<<A.nblk, ST, 0, g.stream>>>(d);
   }
   else
   {
      int grid = cdiv(A.nrows, ST / 32);
      int cap  = g.sm_count * 16;
      if (grid > cap) grid = cap;
      if (grid < 1) grid = 1;
      if (dot) k_spmv_vector<MODE, true><<<grid, ST, 0, g.stream>>>(d);
      else k_spmv_vector<MODE, false><<<grid, ST, 0, g.stream>>>(d);
   }
   HDK_LAUNCH_CHECK();
   return HDK_OK;
}

int spmv_launch(const DevCSR &A, int mode, const SpmvArgs &s)
{
   if (A.nrows <= 0) return HDK_OK;
   SpmvDev d;
   d.rowptr = A.rowptr; d.col = A.col; d.val = A.val; d.blk_row = A.blk_row;
   d.x = s.x; d.b = s.b; d.d = s.d; d.dotv = s.dotv; d.y = s.y;
   d.w = s.w; d.alpha = s.alpha; d.beta = s.beta;
   d.nrows = A.nrows; d.fin = s.fin; d.fin_out = s.fin_out;
   d.scal = g.dscal; d.partials = g.partials; d.ticket = g.counters;
   bool dot = (s.fin != FIN_NONE && s.dotv != nullptr);
   if (A.kind == 0 && A.nblk > PARTIALS_CAP) return set_error(HDK_ERR_UNSUPPORTED, "matrix too large for reduction scratch");
   switch (mode)
   {
      case SPMV_SET: return launch_mode<SPMV_SET>(A, d, dot);
      case SPMV_RESIDUAL: return launch_mode<SPMV_RESIDUAL>(A, d, dot);
      case SPMV_JACOBI: return launch_mode<SPMV_JACOBI>(A, d, dot);
      case SPMV_ADD: return launch_mode<SPMV_ADD>(A, d, dot);
      case SPMV_AXPBY: return launch_mode<SPMV_AXPBY>(A, d, dot);
      case SPMV_JACOBI_R: return launch_mode<SPMV_JACOBI_R>(A, d, dot);
   }
   return set_error(HDK_ERR_INVALID, "unknown spmv mode %d", mode);
}

// ---------------------------------------------------------------------------------------
// analysis: row-length statistics and the nnz-balanced row blocks of the stream kernel
// ---------------------------------------------------------------------------------------
__global__ void k_row_stats(const int *rowptr, int nrows, int *max_row)
{
   int m = 0;
   for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x)
   {
      int l = rowptr[r + 1] - rowptr[r];
      m     = l > m ? l : m;
   }
   for (int o = 16; o > 0; o >>= 1) { int t = __shfl_down_sync(0xffffffffu, m, o); m = t > m ? t : m; }
   if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(max_row, m);
}

__global__ void k_blk_rows(const int *rowptr, int nrows, int nblk, int *blk_row)
{
   int b = blockIdx.x * blockDim.x + threadIdx.x;
   if (b > nblk) return;
   if (b == nblk) { blk_row[b] = nrows; return; }
   int target = b * S_TGT;
   int lo = 0, hi = nrows; // first i in [0,nrows] with rowptr[i] >= target
   while (lo < hi)
   {
      int mid = (lo + hi) >> 1;
      if (rowptr[mid] >= target) hi = mid; else lo = mid + 1;
   }
   blk_row[b] = lo;
}

int csr_analyze(DevCSR &A)
{
   if (A.blk_row) { dfree(A.blk_row); A.blk_row = nullptr; }
   A.nblk = 0; A.kind = 0; A.max_row = 0; A.avg_row = 0.0;
   if (A.nrows <= 0) return HDK_OK;
   int *dmax = reinterpret_cast<int *>(g.dscal + S_TMP3);
   HDK_CUDA(cudaMemsetAsync(dmax, 0, sizeof(int), g.stream));
   int grid = cdiv(A.nrows, 256);
   if (grid > g.sm_count * 8) grid = g.sm_count * 8;
   k_row_stats<<<grid, 256, 0, g.stream>>>(A.rowptr, A.nrows, dmax);
   HDK_LAUNCH_CHECK();
   int hmax = 0;
   HDK_CUDA(cudaMemcpyAsync(&hmax, dmax, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   A.max_row = hmax;
   A.avg_row = (double)A.nnz / (double)A.nrows;
   A.kind    = (hmax <= S_MAXR) ? 0 : 1;
   if (A.kind == 0)
   {
      A.nblk = A.nnz / S_TGT + 1;
      HDK_TRY(dalloc(&A.blk_row, (size_t)A.nblk + 1));
      k_blk_rows<<<cdiv(A.nblk + 1, 256), 256, 0, g.stream>>>(A.rowptr, A.nrows, A.nblk, A.blk_row);
      HDK_LAUNCH_CHECK();
   }
   return HDK_OK;
}

int csr_alloc(DevCSR &A, int nrows, int ncols, int nnz, bool values)
{
   A.nrows = nrows; A.ncols = ncols; A.nnz = nnz; A.owns = true;
   HDK_TRY(dalloc(&A.rowptr, (size_t)nrows + 1));
   HDK_TRY(dalloc(&A.col, (size_t)nnz + 8));
   HDK_CUDA(cudaMemsetAsync(A.col + nnz, 0, sizeof(int) * 8, g.stream));
   if (values)
   {
      HDK_TRY(dalloc(&A.val, (size_t)nnz + 8));
      HDK_CUDA(cudaMemsetAsync(A.val + nnz, 0, sizeof(double) * 8, g.stream));
   }
   else A.val = nullptr;
   return HDK_OK;
}

void csr_free(DevCSR &A)
{
   if (A.owns) { dfree(A.rowptr); dfree(A.col); dfree(A.val); }
   dfree(A.blk_row);
   A = DevCSR();
}

} // namespace hdk
