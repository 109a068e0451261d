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
#include <stdlib.h>

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

// ---------------------------------------------------------------------------------------
// TMA variant of the stream kernel (default): persistent CTAs, two shared-memory stages.
// One elected thread issues 1-D bulk copies (cp.async.bulk -> UBLKCP) of the block's
// contiguous val / col ranges, completion is tracked by an mbarrier; the copies of block
// i+1 are in flight while block i gathers x, multiplies in place and reduces its rows.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
   asm volatile("{\n"
                ".reg .pred P1;\n"
                "LAB_WAIT:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                "@P1 bra DONE;\n"
                "bra LAB_WAIT;\n"
                "DONE:\n"
                "}" ::"r"(bar), "r"(parity) : "memory");
}

struct BlkMeta { int r0, r1, k0, k1; };
__device__ __forceinline__ BlkMeta blk_meta(const SpmvDev &a, int b)
{
   BlkMeta m;
   m.r0 = __ldg(a.blk_row + b);
   m.r1 = __ldg(a.blk_row + b + 1);
   m.k0 = __ldg(a.rowptr + m.r0);
   m.k1 = __ldg(a.rowptr + m.r1);
   return m;
}

template <int MODE, bool DOT>
__global__ void __launch_bounds__(ST, 4) k_spmv_tma(SpmvDev a, int nblk, int cap)
{
   constexpr int PF = 2;
   extern __shared__ __align__(128) unsigned char smem_raw[];
   __shared__ double red[ST / 32];
   __shared__ int    flag;
   double *vbuf = reinterpret_cast<double *>(smem_raw + 128);
   int    *cbuf = reinterpret_cast<int *>(smem_raw + 128 + (size_t)2 * cap * sizeof(double));
   const uint32_t bar0 = smem_u32(smem_raw), bar1 = bar0 + 8;
   const int tid = threadIdx.x, G = gridDim.x;
   if (tid == 0)
   {
      mbar_init(bar0, 1);
      mbar_init(bar1, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncthreads();
   int      b = blockIdx.x, nb = b + G;
   int      stage = 0;
   uint32_t ph[2] = {0u, 0u};
   double   dacc = 0.0;
   BlkMeta  cur, nxt;
   cur = nxt = BlkMeta{0, 0, 0, 0};
   if (b < nblk)
   {
      cur = blk_meta(a, b);
      int ka = cur.k0 & ~3, len4 = (cur.k1 - ka + 3) & ~3;
      if (tid == 0 && len4 > 0)
      {
         mbar_expect_tx(bar0, (uint32_t)len4 * 12u);
         tma_load_1d(smem_u32(vbuf), a.val + ka, (uint32_t)len4 * 8u, bar0);
         tma_load_1d(smem_u32(cbuf), a.col + ka, (uint32_t)len4 * 4u, bar0);
      }
   }
   if (nb < nblk) nxt = blk_meta(a, nb);
   while (b < nblk)
   {
      // 1. bulk copies of the next block into the other stage (its readers finished before the
      //    barrier that closed the previous iteration)
      if (nb < nblk && tid == 0)
      {
         int ka = nxt.k0 & ~3, len4 = (nxt.k1 - ka + 3) & ~3;
         if (len4 > 0)
         {
            const uint32_t bar = stage ? bar0 : bar1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, (uint32_t)len4 * 12u);
            tma_load_1d(smem_u32(vbuf + (size_t)(stage ^ 1) * cap), a.val + ka, (uint32_t)len4 * 8u, bar);
            tma_load_1d(smem_u32(cbuf + (size_t)(stage ^ 1) * cap), a.col + ka, (uint32_t)len4 * 4u, bar);
         }
      }
      // 2. metadata two blocks ahead, 3. per-row operands of this block (plain global loads)
      const int nnb = nb + G;
      BlkMeta   nn  = BlkMeta{0, 0, 0, 0};
      if (nnb < nblk) nn = blk_meta(a, nnb);
      const int ka = cur.k0 & ~3, len = cur.k1 - ka;
      RowOps    ro[PF];
#pragma unroll
      for (int j = 0; j < PF; ++j)
      {
         int r = cur.r0 + tid + j * ST;
         if (r < cur.r1) row_load<MODE, DOT>(a, r, ka, ro[j]);
      }
      // 4. wait for this block's val / col
      double *prod = vbuf + (size_t)stage * cap;
      const int *cs = cbuf + (size_t)stage * cap;
      if (len > 0)
      {
         mbar_wait(stage ? bar1 : bar0, ph[stage]);
         ph[stage] ^= 1u;
      }
      // 5. gather x and multiply in place
      for (int k = tid * 4; k < len; k += ST * 4)
      {
         int4    c  = *reinterpret_cast<const int4 *>(cs + k);
         double2 v0 = *reinterpret_cast<const double2 *>(prod + k);
         double2 v1 = *reinterpret_cast<const double2 *>(prod + k + 2);
         double  x0 = __ldg(a.x + c.x), x1 = __ldg(a.x + c.y), x2 = __ldg(a.x + c.z), x3 = __ldg(a.x + c.w);
         v0.x = __dmul_rn(v0.x, x0); v0.y = __dmul_rn(v0.y, x1);
         v1.x = __dmul_rn(v1.x, x2); v1.y = __dmul_rn(v1.y, x3);
         *reinterpret_cast<double2 *>(prod + k)     = v0;
         *reinterpret_cast<double2 *>(prod + k + 2) = v1;
      }
      __syncthreads();
      // 6. one thread per row, sequential sum in CSR order, fused epilogue
#pragma unroll
      for (int j = 0; j < PF; ++j)
      {
         int r = cur.r0 + tid + j * ST;
         if (r < cur.r1)
         {
            double yn = row_epilogue<MODE>(a, ro[j], prod);
            a.y[r]    = yn;
            if (DOT) dacc += ro[j].dv * yn;
         }
      }
      for (int r = cur.r0 + tid + PF * ST; r < cur.r1; r += ST)
      {
         RowOps o;
         row_load<MODE, DOT>(a, r, ka, o);
         double yn = row_epilogue<MODE>(a, o, prod);
         a.y[r]    = yn;
         if (DOT) dacc += o.dv * yn;
      }
      __syncthreads();
      b = nb; nb = nnb; cur = nxt; nxt = nn; stage ^= 1;
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

static int g_spmv_impl = -1; // 0 = LDG stream kernel, 1 = TMA stream kernel (default)

template <int MODE, bool DOT>
static int launch_tma(const DevCSR &A, const SpmvDev &d)
{
   static int    occ = 0;
   static size_t occ_smem = 0;
   size_t        smem = 128 + (size_t)2 * A.cap * 12;
   if (smem != occ_smem)
   {
      HDK_CUDA(cudaFuncSetAttribute(k_spmv_tma<MODE, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      HDK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv_tma<MODE, DOT>, ST, smem));
      if (occ < 1) occ = 1;
      occ_smem = smem;
   }
   int grid = g.sm_count * occ;
   if (grid > A.nblk) grid = A.nblk;
   k_spmv_tma<MODE, DOT><<<grid, ST, smem, g.stream>>>(d, A.nblk, A.cap);
   return HDK_OK;
}

template <int MODE>
static int launch_mode(const DevCSR &A, const SpmvDev &d, bool dot)
{
   if (g_spmv_impl < 0)
   {
      const char *e = getenv("HDK_SPMV_IMPL");
      g_spmv_impl   = (e && !strcmp(e, "ldg")) ? 0 : 1;
   }
   if (A.kind == 0 && g_spmv_impl == 1)
   {
      if (dot) HDK_TRY((launch_tma<MODE, true>(A, d)));
      else HDK_TRY((launch_tma<MODE, false>(A, d)));
   }
   else if (A.kind == 0)
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

__global__ void k_blk_rows(const int *rowptr, int nrows, int nblk, int tgt, int *blk_row)
{
   int b = blockIdx.x * blockDim.x + threadIdx.x;
   if (b > nblk) return;
   if (b == nblk) { blk_row[b] = nrows; return; }
   int target = b * tgt;
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
      // non-zeros per CTA: about one row per thread (256 rows), 64-aligned, within [1024, 3072]
      int tgt = ((int)(A.avg_row * 256.0) / 64) * 64;
      if (tgt < 1024) tgt = 1024;
      if (tgt > S_TGT) tgt = S_TGT;
      A.tgt  = tgt;
      A.cap  = (tgt + hmax + 8 + 3) & ~3;   // shared-memory entries per stage (<= S_CAP)
      A.nblk = A.nnz / tgt + 1;
      HDK_TRY(dalloc(&A.blk_row, (size_t)A.nblk + 1));
      k_blk_rows<<<cdiv(A.nblk + 1, 256), 256, 0, g.stream>>>(A.rowptr, A.nrows, A.nblk, tgt, A.blk_row);
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
