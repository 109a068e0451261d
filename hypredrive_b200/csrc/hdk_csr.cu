// hdk_csr.cu -- ParCSR assembly on the device and the matvec entry points.
// Stands in for HYPRE_IJMatrixCreate/SetValues/Assemble and the ParCSR diag/offd split
// (reference: src/internal/linsys.c:1190-1405, src/HYPREDRV.c:2141-2191) and for the example
// generators (examples/src/C_laplacian/laplacian.c:720-921, 1138-1356;
// examples/src/C_convdif/convdif.c:782-990).
#include "hdk_internal.cuh"
#include <cub/cub.cuh>

namespace hdk {

// ---- split global-column CSR rows into diag (local int32 cols) and offd (global cols) ----
__global__ void k_split_count(const int64_t *indptr, const int64_t *cols, int nrows, int64_t rs,
                              int64_t re, int *cnt_d, int *cnt_o)
{
   int r = blockIdx.x * blockDim.x + threadIdx.x;
   if (r > nrows) return;
   if (r == nrows) { cnt_d[r] = 0; cnt_o[r] = 0; return; }
   int cd = 0, co = 0;
   for (int64_t k = indptr[r]; k < indptr[r + 1]; k++)
   {
      int64_t c = cols[k];
      if (c >= rs && c <= re) cd++; else co++;
   }
   cnt_d[r] = cd; cnt_o[r] = co;
}

// fills both blocks in input order, then swaps the diagonal entry to the front of the diag row
// (hypre IJ assembly / hypre_CSRMatrixReorder semantics: a swap, not a rotation)
__global__ void k_split_fill(const int64_t *indptr, const int64_t *cols, const double *vals, int nrows,
                             int64_t rs, int64_t re, const int *rp_d, int *col_d, double *val_d,
                             const int *rp_o, int64_t *gcol_o, double *val_o, int square)
{
   int r = blockIdx.x * blockDim.x + threadIdx.x;
   if (r >= nrows) return;
   int pd = rp_d[r], po = rp_o[r];
   int dpos = -1;
   const int p0 = pd;
   for (int64_t k = indptr[r]; k < indptr[r + 1]; k++)
   {
      int64_t c = cols[k];
      if (c >= rs && c <= re)
      {
         int lc = (int)(c - rs);
         if (square && lc == r && dpos < 0) dpos = pd;
         col_d[pd] = lc; val_d[pd] = vals[k]; pd++;
      }
      else { gcol_o[po] = c; val_o[po] = vals[k]; po++; }
   }
   if (dpos > p0)
   {
      int tc = col_d[p0]; double tv = val_d[p0];
      col_d[p0] = col_d[dpos]; val_d[p0] = val_d[dpos];
      col_d[dpos] = tc; val_d[dpos] = tv;
   }
}

__global__ void k_map_offd(const int64_t *gcol, int nnz, const int64_t *map, int nmap, int *col)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= nnz) return;
   int64_t c = gcol[k];
   int lo = 0, hi = nmap - 1;
   while (lo < hi) { int mid = (lo + hi) >> 1; if (map[mid] < c) lo = mid + 1; else hi = mid; }
   col[k] = lo;
}

// ---- exclusive prefix sums (hand-written reduce-then-scan: block sums -> scan of the block sums
// (recursive) -> block-local scan + offset).  1024 items per block, 4 per thread; in == out allowed.
constexpr int SCAN_T = 256, SCAN_ITEMS = 4, SCAN_B = SCAN_T * SCAN_ITEMS;

template <class TI, class TO>
__global__ void __launch_bounds__(SCAN_T) k_scan_sums(const TI *in, int64_t n, TO *bsum)
{
   __shared__ TO sm[SCAN_T / 32];
   const int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_ITEMS;
   TO            s = 0;
#pragma unroll
   for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) s += (TO)in[base + k];
   for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
   if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
   __syncthreads();
   if (threadIdx.x == 0)
   {
      TO t = 0;
      for (int w = 0; w < SCAN_T / 32; w++) t += sm[w];
      bsum[blockIdx.x] = t;
   }
}

template <class TI, class TO>
__global__ void __launch_bounds__(SCAN_T) k_scan_apply(const TI *in, TO *out, int64_t n, const TO *boff)
{
   __shared__ TO sm[SCAN_T / 32];
   const int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_ITEMS;
   const int     lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   TO            v[SCAN_ITEMS], s = 0;
#pragma unroll
   for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? (TO)in[base + k] : (TO)0; s += v[k]; }
   TO incl = s;                                  // inclusive scan of the thread sums inside the warp
#pragma unroll
   for (int o = 1; o < 32; o <<= 1) { TO u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
   if (lane == 31) sm[w] = incl;
   __syncthreads();
   TO woff = 0;
   for (int q = 0; q < w; q++) woff += sm[q];
   TO run = (boff ? boff[blockIdx.x] : (TO)0) + woff + incl - s;
#pragma unroll
   for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = run; run += v[k]; }
}

template <class TI, class TO>
static int exclusive_scan_t(const TI *in, TO *out, int64_t n)
{
   if (n <= 0) return HDK_OK;
   const int64_t nb = (n + SCAN_B - 1) / SCAN_B;
   if (nb == 1)
   {
      k_scan_apply<TI, TO><<<1, SCAN_T, 0, g.stream>>>(in, out, n, nullptr);
      HDK_LAUNCH_CHECK();
      return HDK_OK;
   }
   TO *bsum;
   HDK_TRY(dalloc(&bsum, (size_t)nb));
   k_scan_sums<TI, TO><<<(unsigned)nb, SCAN_T, 0, g.stream>>>(in, n, bsum);
   HDK_LAUNCH_CHECK();
   HDK_TRY((exclusive_scan_t<TO, TO>(bsum, bsum, nb)));
   k_scan_apply<TI, TO><<<(unsigned)nb, SCAN_T, 0, g.stream>>>(in, out, n, bsum);
   HDK_LAUNCH_CHECK();
   dfree(bsum);
   return HDK_OK;
}

int exclusive_scan_int(const int *in, int *out, int n) { return exclusive_scan_t<int, int>(in, out, n); }
int exclusive_scan_i64(const int *in, int64_t *out, int n) { return exclusive_scan_t<int, int64_t>(in, out, n); }
int exclusive_scan_i64_i64(const int64_t *in, int64_t *out, int64_t n) { return exclusive_scan_t<int64_t, int64_t>(in, out, n); }

int build_halo_plan(hdk_csr_s &A, int64_t *gcol_sorted_unique, int n_halo); // hdk_comm.cu

// rows with at least one off-rank entry (hypre keeps the same list as offd "rownnz")
__global__ void k_offd_rows(const int *rowptr, int nrows, int *list, int *count)
{
   int r = blockIdx.x * blockDim.x + threadIdx.x;
   if (r < nrows && rowptr[r + 1] > rowptr[r]) list[atomicAdd(count, 1)] = r;
}

// copy of the caller's rows with global columns, diagonal entry swapped to the front: kept at
// N > 1 so that the multi-rank setup can reassemble the global operator in its original order
__global__ void k_orig_diag_first(const int64_t *indptr, int64_t *cols, double *vals, int nrows, int64_t rs)
{
   int r = blockIdx.x * blockDim.x + threadIdx.x;
   if (r >= nrows) return;
   int64_t b = indptr[r], e = indptr[r + 1], me = rs + r;
   for (int64_t k = b; k < e; k++)
      if (cols[k] == me)
      {
         if (k != b)
         {
            int64_t tc = cols[b]; double tv = vals[b];
            cols[b] = cols[k]; vals[b] = vals[k];
            cols[k] = tc; vals[k] = tv;
         }
         break;
      }
}

// Build the rank-local ParCSR block pair from rows [rs,re] with global columns.  Columns in
// [cs,ce] (this rank's share of the column space) go to the diag block.  `distributed`: the
// matrix is one slab of a matrix spread over all ranks (collective: every rank must call).
int parcsr_build(int64_t rs, int64_t re, int64_t cs, int64_t ce, int64_t grows, int64_t gcols, bool square,
                 bool distributed, bool keep_orig, const int64_t *indptr, const int64_t *cols, const double *vals,
                 hdk_csr_s **out, bool analyze)
{
   // a slab of a distributed matrix may be empty (re == rs - 1): a rank that owns no coarse point
   if (re < rs - 1 || (re < rs && !distributed))
      return set_error(HDK_ERR_INVALID, "empty local row range [%lld,%lld]", (long long)rs, (long long)re);
   int64_t n64 = re - rs + 1;
   if (n64 > 2000000000LL) return set_error(HDK_ERR_UNSUPPORTED, "local rows exceed int32");
   int        n = (int)n64;
   hdk_csr_s *A = new hdk_csr_s();
   A->row_start = rs; A->row_end = re; A->global_rows = grows;
   A->col_start = cs; A->col_end = ce; A->global_cols = gcols;
   int *cnt_d, *cnt_o, *rp_d, *rp_o;
   HDK_TRY(dalloc(&cnt_d, (size_t)n + 1));
   HDK_TRY(dalloc(&cnt_o, (size_t)n + 1));
   HDK_TRY(dalloc(&rp_d, (size_t)n + 1));
   HDK_TRY(dalloc(&rp_o, (size_t)n + 1));
   k_split_count<<<cdiv(n + 1, 256), 256, 0, g.stream>>>(indptr, cols, n, cs, ce, cnt_d, cnt_o);
   HDK_LAUNCH_CHECK();
   HDK_TRY(exclusive_scan_int(cnt_d, rp_d, n + 1));
   HDK_TRY(exclusive_scan_int(cnt_o, rp_o, n + 1));
   int tot[2];
   HDK_CUDA(cudaMemcpyAsync(&tot[0], rp_d + n, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaMemcpyAsync(&tot[1], rp_o + n, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   dfree(cnt_d); dfree(cnt_o);
   int nh_guess = tot[1];
   HDK_TRY(csr_alloc(A->diag, n, (int)(ce - cs + 1), tot[0]));
   HDK_TRY(csr_alloc(A->offd, n, 0, nh_guess));
   dfree(A->diag.rowptr); A->diag.rowptr = rp_d;
   dfree(A->offd.rowptr); A->offd.rowptr = rp_o;
   int64_t *gcol_o;
   HDK_TRY(dalloc(&gcol_o, (size_t)tot[1] + 1));
   if (n > 0)
   {
      k_split_fill<<<cdiv(n, 256), 256, 0, g.stream>>>(indptr, cols, vals, n, cs, ce, rp_d, A->diag.col,
                                                       A->diag.val, rp_o, gcol_o, A->offd.val, square ? 1 : 0);
      HDK_LAUNCH_CHECK();
   }
   int n_halo = 0;
   int64_t *uniq = nullptr;
   if (tot[1] > 0)
   {
      // sorted unique global ids of the off-rank columns -> col_map_offd
      int64_t *sorted;
      int     *nsel;
      HDK_TRY(dalloc(&sorted, (size_t)tot[1]));
      HDK_TRY(dalloc(&uniq, (size_t)tot[1]));
      HDK_TRY(dalloc(&nsel, 1));
      size_t b1 = 0, b2 = 0;
      HDK_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, b1, gcol_o, sorted, tot[1], 0, 64, g.stream));
      HDK_CUDA(cub::DeviceSelect::Unique(nullptr, b2, sorted, uniq, nsel, tot[1], g.stream));
      char *tmp;
      HDK_TRY(dalloc(&tmp, b1 > b2 ? b1 : b2));
      HDK_CUDA(cub::DeviceRadixSort::SortKeys(tmp, b1, gcol_o, sorted, tot[1], 0, 64, g.stream));
      HDK_CUDA(cub::DeviceSelect::Unique(tmp, b2, sorted, uniq, nsel, tot[1], g.stream));
      HDK_CUDA(cudaMemcpyAsync(&n_halo, nsel, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
      HDK_CUDA(cudaStreamSynchronize(g.stream));
      k_map_offd<<<cdiv(tot[1], 256), 256, 0, g.stream>>>(gcol_o, tot[1], uniq, n_halo, A->offd.col);
      HDK_LAUNCH_CHECK();
      dfree(tmp); dfree(sorted); dfree(nsel);
      A->offd.ncols = n_halo;
   }
   if (distributed && g.nranks > 1)
   {
      int rc = build_halo_plan(*A, uniq, n_halo); // collective, also with n_halo == 0
      if (rc != HDK_OK) return rc;
   }
   else if (tot[1] > 0)
      return set_error(HDK_ERR_INVALID, "matrix has %d off-rank columns but is not distributed", n_halo);
   dfree(gcol_o);
   A->diag.offd_rowptr = (tot[1] > 0) ? A->offd.rowptr : nullptr;
   if (analyze) HDK_TRY(csr_analyze(A->diag));
   if (A->offd.nnz > 0)
   {
      int *cnt = reinterpret_cast<int *>(g.dscal + S_TMP2);
      HDK_TRY(dalloc(&A->offd_rows, (size_t)(A->offd.nnz < n ? A->offd.nnz : n) + 1));
      HDK_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int), g.stream));
      k_offd_rows<<<cdiv(n, 256), 256, 0, g.stream>>>(A->offd.rowptr, n, A->offd_rows, cnt);
      HDK_LAUNCH_CHECK();
      HDK_CUDA(cudaMemcpyAsync(&A->n_offd_rows, cnt, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
      HDK_CUDA(cudaStreamSynchronize(g.stream));
   }
   // global nnz
   double loc = (double)tot[0] + (double)tot[1];
   if (distributed && g.nranks > 1)
   {
      HDK_CUDA(cudaMemcpyAsync(g.dscal + S_TMP0, &loc, sizeof(double), cudaMemcpyHostToDevice, g.stream));
      HDK_TRY(allreduce_dev(g.dscal + S_TMP0, 1));
      HDK_CUDA(cudaMemcpyAsync(&loc, g.dscal + S_TMP0, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
      HDK_CUDA(cudaStreamSynchronize(g.stream));
   }
   A->global_nnz = (int64_t)loc;
   static const bool force_dist = getenv("HDK_SETUP_DIST_FORCE") && atoi(getenv("HDK_SETUP_DIST_FORCE")) == 1;
   if (keep_orig && distributed && (g.nranks > 1 || force_dist))
   {
      int64_t nnz = (int64_t)tot[0] + tot[1];
      HDK_TRY(dalloc(&A->orig_indptr, (size_t)n + 1));
      HDK_TRY(dalloc(&A->orig_cols, (size_t)nnz + 1));
      HDK_TRY(dalloc(&A->orig_vals, (size_t)nnz + 1));
      HDK_CUDA(cudaMemcpyAsync(A->orig_indptr, indptr, sizeof(int64_t) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, g.stream));
      HDK_CUDA(cudaMemcpyAsync(A->orig_cols, cols, sizeof(int64_t) * (size_t)nnz, cudaMemcpyDeviceToDevice, g.stream));
      HDK_CUDA(cudaMemcpyAsync(A->orig_vals, vals, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToDevice, g.stream));
      if (n > 0)
      {
         k_orig_diag_first<<<cdiv(n, 256), 256, 0, g.stream>>>(A->orig_indptr, A->orig_cols, A->orig_vals, n, rs);
         HDK_LAUNCH_CHECK();
      }
      A->orig_nnz = nnz;
   }
   *out = A;
   return HDK_OK;
}

static int parcsr_from_device(int64_t rs, int64_t re, int64_t grows, const int64_t *indptr, const int64_t *cols,
                              const double *vals, hdk_csr_s **out)
{
   return parcsr_build(rs, re, rs, re, grows, grows, true, true, true, indptr, cols, vals, out);
}

// ---- y = op(A) x with halo exchange: diag kernel overlaps the exchange, offd kernel follows
__global__ void k_offd_correct(const int *rows, const int *rowptr, const int *col, const double *val, int nrows,
                               const double *xh, double *y, const double *d, double w, int mode,
                               double alpha, const double *dotv, const double *diag_part,
                               double *dot_out, double *partials, unsigned *ticket);

// true when parcsr_matvec runs as ONE kernel (no off-rank block, or the block fused into the
// sliced-ELL kernel): then every fused epilogue sees the complete row sum
bool parcsr_single_kernel(const hdk_csr_s &A)
{
   bool exch = g.nranks > 1 && (A.halo.n_send > 0 || A.halo.n_halo > 0);
   if (!exch || A.offd.nnz == 0) return true;
   return A.diag.kind == 2 && A.diag.sl_offd_flags;
}

// (hdk_time_kernel 8: time the product without its halo exchange)
static bool skip_exchange = false;
void dbg_skip_exchange(bool on) { skip_exchange = on; }

int parcsr_matvec(const hdk_csr_s &A, int mode, SpmvArgs a)
{
   // p2p exchange partners must match: a rank takes part when it sends OR receives
   bool exch = g.nranks > 1 && (A.halo.n_send > 0 || A.halo.n_halo > 0);
   if (!exch) return spmv_launch(A.diag, mode, a);
   const bool fused_offd = A.diag.kind == 2 && A.diag.sl_offd_flags && A.offd.nnz > 0;
   if (mode == SPMV_SET_DIV && !fused_offd && A.offd.nnz > 0)
   {
      // the second output needs the final y: plain product first, then the scaled division
      SpmvArgs a1 = a;
      a1.y2 = nullptr;
      HDK_TRY(parcsr_matvec(A, SPMV_SET, a1));
      return vec_scaled_div(a.y2, a.y, a.d, a.w, A.diag.nrows);
   }
   if (fused_offd)
   {
      // sliced-ELL with flagged boundary rows: ONE kernel does the local block, adds the off-rank entries of
      // the flagged rows and runs the fused epilogue / dot exactly as on a single rank.  The exchange is complete before the kernel starts (peer-memory
      // path: inside the pack / exporting kernel; NCCL path: an event).
      if (!skip_exchange)
      {
         HDK_TRY(halo_exchange_begin(A, a.x));
         HDK_TRY(halo_exchange_end(A));
      }
      OffdFuse of;
      of.orp = A.offd.rowptr; of.ocol = A.offd.col; of.oval = A.offd.val;
      of.xh = halo_buffer(A);
      return spmv_launch(A.diag, mode, a, &of);
   }
   // start the exchange, run the diag block.  A fused dot <dotv, y> is linear in the off-diagonal
   // correction: the diag kernel leaves <dotv, y_diag> in a scratch scalar and the correction
   // kernel adds <dotv, delta y> over its rows -- no extra pass over the vectors
   HDK_TRY(halo_exchange_begin(A, a.x));
   const bool want_dot  = (a.fin != FIN_NONE && a.dotv);
   const bool fused_dot = want_dot && a.fin == FIN_STORE && a.fin_out != nullptr;
   SpmvArgs ad = a;
   if (fused_dot && A.offd.nnz > 0) ad.fin_out = g.dscal + S_TMP1;
   else if (!fused_dot) { ad.fin = FIN_NONE; ad.dotv = nullptr; }
   HDK_TRY(spmv_launch(A.diag, mode, ad));
   HDK_TRY(halo_exchange_end(A));
   // offd contribution: y_i += sign * sum_o (linear correction of the diag-only epilogue)
   if (A.offd.nnz > 0)
   {
      int           grid = cdiv(A.n_offd_rows, 128);
      k_offd_correct<<<grid, 128, 0, g.stream>>>(A.offd_rows, A.offd.rowptr, A.offd.col, A.offd.val, A.n_offd_rows,
                                                 halo_buffer(A), a.y, a.d, a.w, mode, a.alpha, fused_dot ? a.dotv : nullptr,
                                                 g.dscal + S_TMP1, a.fin_out, g.partials, g.counters + 1);
      HDK_LAUNCH_CHECK();
   }
   if (want_dot && !fused_dot)
   {
      HDK_TRY(vec_dot_dev(a.dotv, a.y, A.diag.nrows, a.fin, a.fin_out));
   }
   return HDK_OK;
}

__global__ void k_offd_correct(const int *rows, const int *rowptr, const int *col, const double *val, int nrows,
                               const double *xh, double *y, const double *d, double w, int mode,
                               double alpha, const double *dotv, const double *diag_part,
                               double *dot_out, double *partials, unsigned *ticket)
{
   __shared__ double red[128 / 32];
   __shared__ int    lastflag;
   double            dcorr = 0.0; // dotv[r] * (change of y[r])
   // (peer-memory exchange: complete before this kernel starts, see halo_exchange_begin)
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < nrows)
   {
      int r = rows[i];
      int s = rowptr[r], e = rowptr[r + 1];
      double acc = 0.0;
      for (int k = s; k < e; k++) acc += val[k] * __ldcg(xh + col[k]); // L2: written by a peer
      double delta = 0.0;
      switch (mode)
      {
         case SPMV_SET:
         case SPMV_ADD: delta = acc; break;
         case SPMV_AXPBY: delta = alpha * acc; break;
         case SPMV_RESIDUAL: delta = -acc; break;
         case SPMV_JACOBI:
         case SPMV_JACOBI_R: { double dd = d[r]; if (dd != 0.0) delta = -(w * acc) / dd; break; }
      }
      y[r] += delta;
      if (dotv) dcorr = dotv[r] * delta;
   }
   if (dotv)
   {
      // <dotv, y> = <dotv, y_diag> (left in *diag_part by the diag kernel) + sum of the corrections
      double bs = block_sum<128>(dcorr, red);
      if (threadIdx.x == 0)
      {
         partials[blockIdx.x] = bs;
         __threadfence();
         unsigned t = atomicInc(ticket, gridDim.x - 1);
         lastflag   = (t == gridDim.x - 1);
      }
      __syncthreads();
      if (lastflag)
      {
         __threadfence();
         double sacc = 0.0;
         for (unsigned b = threadIdx.x; b < gridDim.x; b += 128) sacc += __ldcg(partials + b);
         __syncthreads();
         sacc = block_sum<128>(sacc, red);
         if (threadIdx.x == 0) *dot_out = *diag_part + sacc;
      }
   }
}

// ---- synthetic stencils generated on the device ------------------------------------------
__device__ __forceinline__ void grid_xyz(int64_t row, int nx, int ny, int &gx, int &gy, int &gz)
{
   gx = (int)(row % nx);
   int64_t t = row / nx;
   gy = (int)(t % ny);
   gz = (int)(t / ny);
}

__device__ __forceinline__ double axial_velocity(double y, double z, double Hy, double Hz, double umax)
{
   return 16.0 * umax * (y / Hy) * (1.0 - y / Hy) * (z / Hz) * (1.0 - z / Hz);
}

__global__ void k_stencil_count(int kind, int nx, int ny, int nz, int64_t rs, int nrows, int64_t *cnt)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i > nrows) return;
   if (i == nrows) { cnt[i] = 0; return; }
   int gx, gy, gz;
   grid_xyz(rs + i, nx, ny, gx, gy, gz);
   int c;
   if (kind == 27)
   {
      int ax = 1 + (gx > 0) + (gx < nx - 1), ay = 1 + (gy > 0) + (gy < ny - 1), az = 1 + (gz > 0) + (gz < nz - 1);
      c = ax * ay * az;
   }
   else c = 1 + (gx > 0) + (gx < nx - 1) + (gy > 0) + (gy < ny - 1) + (gz > 0) + (gz < nz - 1);
   cnt[i] = c;
}

__global__ void k_stencil_fill(int kind, int nx, int ny, int nz, double c0, double c1, double c2,
                               int64_t rs, int nrows, const int64_t *indptr, int64_t *cols, double *vals,
                               double *b)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= nrows) return;
   int64_t row = rs + i;
   int gx, gy, gz;
   grid_xyz(row, nx, ny, gx, gy, gz);
   int64_t p = indptr[i];
   const int64_t sx = 1, sy = nx, sz = (int64_t)nx * ny;
   if (kind == 7)
   {
      cols[p] = row; vals[p++] = 2.0 * (c0 + c1 + c2);
      if (gz > 0) { cols[p] = row - sz; vals[p++] = -c2; }
      if (gy > 0) { cols[p] = row - sy; vals[p++] = -c1; }
      if (gx > 0) { cols[p] = row - sx; vals[p++] = -c0; }
      if (gx < nx - 1) { cols[p] = row + sx; vals[p++] = -c0; }
      if (gy < ny - 1) { cols[p] = row + sy; vals[p++] = -c1; }
      if (gz < nz - 1) { cols[p] = row + sz; vals[p++] = -c2; }
      if (b) b[i] = (gy == 0) ? 1.0 : 0.0;
   }
   else if (kind == 27)
   {
      double center = 0.0;
      for (int dz = -1; dz <= 1; dz++)
         for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++)
            {
               if (!dx && !dy && !dz) continue;
               int    ndiff = (dx != 0) + (dy != 0) + (dz != 0);
               double adj   = 0.0;
               if (dx) adj += c0 / ndiff;
               if (dy) adj += c1 / ndiff;
               if (dz) adj += c2 / ndiff;
               int x = gx + dx, y = gy + dy, z = gz + dz;
               if (x >= 0 && x < nx && y >= 0 && y < ny && z >= 0 && z < nz)
               {
                  cols[p] = row + dx * sx + dy * sy + dz * sz; vals[p++] = -adj;
               }
               center += adj;
            }
      cols[p] = row; vals[p++] = center; // centre last: assembly swaps it to the front
      if (b) b[i] = (gy == 0) ? 1.0 : 0.0;
   }
   else
   {
      // upwind FV convection-diffusion, 4 x 1 x 1 duct, c = {kappa, umax, dt}, wmax = 0
      const double Lx = 4.0, Ly = 1.0, Lz = 1.0;
      const double hx = Lx / nx, hy = Ly / ny, hz = Lz / nz;
      const double vol = hx * hy * hz;
      const double Dx = c0 * hy * hz / hx, Dy = c0 * hx * hz / hy, Dz = c0 * hx * hy / hz;
      const double Ax = hy * hz;
      double z = ((double)gz + 0.5) * hz, y = ((double)gy + 0.5) * hy;
      double Cf = axial_velocity(y, z, Ly, Lz, c1) * Ax;
      double Cf_pos = Cf > 0.0 ? Cf : 0.0, Cf_neg = Cf < 0.0 ? -Cf : 0.0;
      int64_t dpos = p++;
      double  diag = vol / c2, rhs = 0.0;
      cols[dpos] = row;
      if (gx > 0) { diag += Dx + Cf_neg; cols[p] = row - sx; vals[p++] = -(Dx + Cf_pos); }
      else { diag += 2.0 * Dx + Cf_neg; rhs = (2.0 * Dx + Cf_pos) * 1.0; }
      if (gx < nx - 1) { diag += Dx + Cf_pos; cols[p] = row + sx; vals[p++] = -(Dx + Cf_neg); }
      else diag += Cf_pos;
      if (gy > 0) { diag += Dy; cols[p] = row - sy; vals[p++] = -Dy; }
      if (gy < ny - 1) { diag += Dy; cols[p] = row + sy; vals[p++] = -Dy; }
      if (gz > 0) { diag += Dz; cols[p] = row - sz; vals[p++] = -Dz; }
      if (gz < nz - 1) { diag += Dz; cols[p] = row + sz; vals[p++] = -Dz; }
      vals[dpos] = diag;
      if (b) b[i] = rhs;
   }
}

// wait budget / error flag of this translation unit's copy of the flag-wait globals
int wait_globals_csr(long long tmo, int *err)
{
   return wait_globals_set(tmo, err) == cudaSuccess ? HDK_OK : set_error(HDK_ERR_CUDA, "cannot set the wait budget");
}

} // namespace hdk

using namespace hdk;

extern "C" {

int hdk_csr_from_device(int64_t row_start, int64_t row_end, int64_t global_rows, const int64_t *indptr_d,
                        const int64_t *cols_d, const double *vals_d, hdk_csr **A)
{
   HDK_TRY(require_init());
   return parcsr_from_device(row_start, row_end, global_rows, indptr_d, cols_d, vals_d, A);
}

int hdk_csr_from_host(int64_t row_start, int64_t row_end, int64_t global_rows, const int64_t *indptr_h,
                      const int64_t *cols_h, const double *vals_h, hdk_csr **A)
{
   HDK_TRY(require_init());
   if (!indptr_h) return set_error(HDK_ERR_INVALID, "indptr is NULL");
   if (row_end < row_start) return set_error(HDK_ERR_INVALID, "invalid row range");
   int64_t n = row_end - row_start + 1;
   int64_t k0 = indptr_h[0], k1 = indptr_h[n];
   if (k0 < 0 || k1 < k0) return set_error(HDK_ERR_INVALID, "indptr must be nonnegative and nondecreasing");
   int64_t nnz = k1 - k0;
   if (nnz > 0 && (!cols_h || !vals_h)) return set_error(HDK_ERR_INVALID, "col_indices/data are NULL");
   if (nnz > 2000000000LL) return set_error(HDK_ERR_UNSUPPORTED, "local nnz exceeds int32");
   int64_t *ip, *cj;
   double  *va;
   HDK_TRY(dalloc(&ip, (size_t)n + 1));
   HDK_TRY(dalloc(&cj, (size_t)nnz + 1));
   HDK_TRY(dalloc(&va, (size_t)nnz + 1));
   // shift indptr on the host into a small staging buffer (offset slabs: indptr[0] > 0)
   std::vector<int64_t> shifted((size_t)n + 1);
   for (int64_t i = 0; i <= n; i++)
   {
      shifted[(size_t)i] = indptr_h[i] - k0;
      if (i > 0 && indptr_h[i] < indptr_h[i - 1])
      {
         dfree(ip); dfree(cj); dfree(va);
         return set_error(HDK_ERR_INVALID, "indptr must be nondecreasing");
      }
   }
   HDK_CUDA(cudaMemcpyAsync(ip, shifted.data(), sizeof(int64_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, g.stream));
   if (nnz > 0)
   {
      HDK_CUDA(cudaMemcpyAsync(cj, cols_h + k0, sizeof(int64_t) * (size_t)nnz, cudaMemcpyHostToDevice, g.stream));
      HDK_CUDA(cudaMemcpyAsync(va, vals_h + k0, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, g.stream));
   }
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   int rc = parcsr_from_device(row_start, row_end, global_rows, ip, cj, va, A);
   dfree(ip); dfree(cj); dfree(va);
   return rc;
}

int hdk_csr_stencil(int kind, int nx, int ny, int nz, const double c[3], int64_t row_start, int64_t row_end,
                    hdk_csr **A, double *b_d)
{
   HDK_TRY(require_init());
   if (kind != 7 && kind != 27 && kind != 107) return set_error(HDK_ERR_INVALID, "unknown stencil kind %d", kind);
   int64_t N = (int64_t)nx * ny * nz;
   if (row_start < 0 || row_end >= N || row_end < row_start) return set_error(HDK_ERR_INVALID, "row range outside the grid");
   int      n = (int)(row_end - row_start + 1);
   int64_t *cnt, *ip;
   HDK_TRY(dalloc(&cnt, (size_t)n + 1));
   HDK_TRY(dalloc(&ip, (size_t)n + 1));
   k_stencil_count<<<cdiv(n + 1, 256), 256, 0, g.stream>>>(kind, nx, ny, nz, row_start, n, cnt);
   HDK_LAUNCH_CHECK();
   HDK_TRY(exclusive_scan_i64_i64(cnt, ip, (int64_t)n + 1));
   int64_t nnz = 0;
   HDK_CUDA(cudaMemcpyAsync(&nnz, ip + n, sizeof(int64_t), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   dfree(cnt);
   int64_t *cols;
   double  *vals;
   HDK_TRY(dalloc(&cols, (size_t)nnz + 1));
   HDK_TRY(dalloc(&vals, (size_t)nnz + 1));
   k_stencil_fill<<<cdiv(n, 256), 256, 0, g.stream>>>(kind, nx, ny, nz, c[0], c[1], c[2], row_start, n, ip, cols, vals, b_d);
   HDK_LAUNCH_CHECK();
   int rc = parcsr_from_device(row_start, row_end, N, ip, cols, vals, A);
   dfree(ip); dfree(cols); dfree(vals);
   return rc;
}

int hdk_csr_destroy(hdk_csr *A)
{
   if (!A) return HDK_OK;
   if (g.inited)
   {
      csr_free(A->diag); csr_free(A->offd);
      halo_plan_free(A->halo);
      dfree(A->orig_indptr); dfree(A->orig_cols); dfree(A->orig_vals); dfree(A->offd_rows);
   }
   delete A;
   return HDK_OK;
}

int hdk_csr_info(const hdk_csr *A, int64_t *local_rows, int64_t *global_rows, int64_t *local_nnz, int64_t *global_nnz)
{
   if (!A) return set_error(HDK_ERR_INVALID, "null matrix");
   if (local_rows) *local_rows = A->diag.nrows;
   if (global_rows) *global_rows = A->global_rows;
   if (local_nnz) *local_nnz = (int64_t)A->diag.nnz + A->offd.nnz;
   if (global_nnz) *global_nnz = A->global_nnz;
   return HDK_OK;
}

int hdk_csr_get_diag(const hdk_csr *A, int32_t *rowptr_h, int32_t *col_h, double *val_h)
{
   HDK_TRY(require_init());
   if (!A) return set_error(HDK_ERR_INVALID, "null matrix");
   if (rowptr_h) HDK_CUDA(cudaMemcpyAsync(rowptr_h, A->diag.rowptr, sizeof(int) * ((size_t)A->diag.nrows + 1), cudaMemcpyDeviceToHost, g.stream));
   if (col_h && A->diag.nnz) HDK_CUDA(cudaMemcpyAsync(col_h, A->diag.col, sizeof(int) * (size_t)A->diag.nnz, cudaMemcpyDeviceToHost, g.stream));
   if (val_h && A->diag.nnz) HDK_CUDA(cudaMemcpyAsync(val_h, A->diag.val, sizeof(double) * (size_t)A->diag.nnz, cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   return HDK_OK;
}

int hdk_csr_matvec(const hdk_csr *A, double alpha, const double *x_d, double beta, double *y_d)
{
   HDK_TRY(require_init());
   if (!A) return set_error(HDK_ERR_INVALID, "null matrix");
   SpmvArgs a;
   a.x = x_d; a.y = y_d; a.alpha = alpha; a.beta = beta;
   if (alpha == 1.0 && beta == 0.0) return parcsr_matvec(*A, SPMV_SET, a);
   return parcsr_matvec(*A, SPMV_AXPBY, a);
}

int hdk_csr_residual(const hdk_csr *A, const double *x_d, const double *b_d, double *r_d)
{
   HDK_TRY(require_init());
   if (!A) return set_error(HDK_ERR_INVALID, "null matrix");
   SpmvArgs a;
   a.x = x_d; a.y = r_d; a.b = b_d;
   return parcsr_matvec(*A, SPMV_RESIDUAL, a);
}

int hdk_csr_spmv_kind(const hdk_csr *A, int *kind, double *avg_row, int *max_row)
{
   if (!A) return set_error(HDK_ERR_INVALID, "null matrix");
   if (kind) *kind = A->diag.kind;
   if (avg_row) *avg_row = A->diag.avg_row;
   if (max_row) *max_row = A->diag.max_row;
   return HDK_OK;
}

} // extern "C"
