// hdk_amg_dist.cu -- BoomerAMG setup on a ROW-PARTITIONED matrix (one slab of rows per GPU).
// Stands in for the distributed side of HYPRE_BoomerAMGSetup (reference trigger
// src/internal/solver.c:296, src/internal/precon.c:107; partition contract include/HYPREDRV.h:836-839).
//
// Nothing global is ever assembled.  Every rank keeps its rows of A_l, P_l, R_l and builds each
// level from "extended blocks": its own rows plus the few off-rank rows a stage needs, fetched from
// their owners and renumbered into a compact local index space that keeps the global order.  The
// serial device kernels (strength, PMIS sweeps, extended+i, Galerkin product -- hdk_amg_setup.cu)
// run unchanged on those blocks over the owned row range, so every owned row goes through exactly
// the operation sequence of the one-rank setup and the hierarchy is BIT-IDENTICAL to the
// single-GPU (and oracle) hierarchy for any partition:
//   strength     row-local: owned + ring-1 rows (the rows of my off-rank columns).
//   PMIS         measures = |S^T_i| + hypre_Rand stream indexed by the GLOBAL row; column counts
//                are completed by a reverse halo sum; each sweep exchanges the marks of the halo
//                (reverse: "cleared by a neighbour", forward: the owners' decisions).
//   ext+i        needs A and S rows of ring 1 and the C/F marks and coarse numbers of rings 1-2;
//                the global coarse numbering is an exclusive scan of the per-rank C counts.
//   R = P^T      every rank routes its (coarse, fine, weight) triples to the owner of the coarse
//                row, which orders each row by fine index -- hypre's transpose order.
//   RAP          owner-computes: the A rows of the off-rank fine points of my R rows and the P rows
//                of their columns are fetched, then the serial warp kernel runs on the owned coarse rows.
// Levels with at most `replicate_rows` global rows are gathered once and continued by the serial
// setup on every rank (replicated tail, as in the solve phase).  All exchanges are sized
// surface-of-slab; the global row and non-zero counts never meet an int32.
#include "hdk_amg.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <stdlib.h>

namespace hdk {

#define SF_PT -3

// rows [row0, row0 + n) of a distributed matrix with GLOBAL columns in serial storage order
struct GRows
{
   int      n = 0;
   int64_t  row0 = 0, nnz = 0;
   int64_t *ip = nullptr, *col = nullptr; // ip: n + 1 offsets starting at 0
   double  *val = nullptr;
   bool     owns = true;
};
static void grows_free(GRows &G)
{
   if (G.owns) { dfree(G.ip); dfree(G.col); dfree(G.val); }
   G = GRows();
}

// ---------------------------------------------------------------------------------------------
// id lists
// ---------------------------------------------------------------------------------------------
static int sorted_unique_i64(const int64_t *in, int64_t n, int64_t **out, int *nout)
{
   *nout = 0;
   HDK_TRY(dalloc(out, (size_t)(n > 0 ? n : 1)));
   if (n <= 0) return HDK_OK;
   int64_t *sorted;
   int     *nsel;
   HDK_TRY(dalloc(&sorted, (size_t)n));
   HDK_TRY(dalloc(&nsel, 1));
   size_t b1 = 0, b2 = 0;
   HDK_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, b1, in, sorted, n, 0, 64, g.stream));
   HDK_CUDA(cub::DeviceSelect::Unique(nullptr, b2, sorted, *out, nsel, n, g.stream));
   char *tmp;
   HDK_TRY(dalloc(&tmp, b1 > b2 ? b1 : b2));
   HDK_CUDA(cub::DeviceRadixSort::SortKeys(tmp, b1, in, sorted, n, 0, 64, g.stream));
   HDK_CUDA(cub::DeviceSelect::Unique(tmp, b2, sorted, *out, nsel, n, g.stream));
   g.launches += 2;
   HDK_CUDA(cudaMemcpyAsync(nout, nsel, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   dfree(tmp); dfree(sorted); dfree(nsel);
   return HDK_OK;
}

__device__ __forceinline__ int lower_bound_i64(const int64_t *v, int n, int64_t key)
{
   int lo = 0, hi = n;
   while (lo < hi) { int mid = (lo + hi) >> 1; if (v[mid] < key) lo = mid + 1; else hi = mid; }
   return lo;
}
__device__ __forceinline__ bool contains_i64(const int64_t *v, int n, int64_t key)
{
   int p = lower_bound_i64(v, n, key);
   return p < n && v[p] == key;
}

// flag[k] = 1 when cols[k] lies outside [lo, hi) and in neither sorted exclusion list
__global__ void k_flag_foreign(const int64_t *cols, int64_t n, int64_t lo, int64_t hi, const int64_t *ex1, int n1,
                               const int64_t *ex2, int n2, int *flag)
{
   int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
   if (k > n) return;
   int f = 0;
   if (k < n)
   {
      int64_t c = cols[k];
      f = !(c >= lo && c < hi) && !(n1 > 0 && contains_i64(ex1, n1, c)) && !(n2 > 0 && contains_i64(ex2, n2, c));
   }
   flag[k] = f;
}
__global__ void k_scatter_flagged(const int64_t *cols, int64_t n, const int *flag, const int *pos, int64_t *out)
{
   int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
   if (k < n && flag[k]) out[pos[k]] = cols[k];
}

// sorted unique ids among cols[0..n) outside [lo,hi) and outside the exclusion lists
static int collect_foreign(const int64_t *cols, int64_t n, int64_t lo, int64_t hi, const int64_t *ex1, int n1,
                           const int64_t *ex2, int n2, int64_t **out, int *nout)
{
   *nout = 0;
   if (n <= 0) { HDK_TRY(dalloc(out, 1)); return HDK_OK; }
   if (n >= 2000000000LL) return set_error(HDK_ERR_UNSUPPORTED, "collect_foreign: list too long");
   int *flag, *pos;
   HDK_TRY(dalloc(&flag, (size_t)n + 1));
   HDK_TRY(dalloc(&pos, (size_t)n + 1));
   k_flag_foreign<<<cdiv(n + 1, 256), 256, 0, g.stream>>>(cols, n, lo, hi, ex1, n1, ex2, n2, flag);
   HDK_LAUNCH_CHECK();
   HDK_TRY(exclusive_scan_int(flag, pos, (int)n + 1));
   int cnt = 0;
   HDK_CUDA(cudaMemcpyAsync(&cnt, pos + n, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   int64_t *sel;
   HDK_TRY(dalloc(&sel, (size_t)(cnt > 0 ? cnt : 1)));
   if (cnt > 0)
   {
      k_scatter_flagged<<<cdiv(n, 256), 256, 0, g.stream>>>(cols, n, flag, pos, sel);
      HDK_LAUNCH_CHECK();
   }
   dfree(flag); dfree(pos);
   int rc = sorted_unique_i64(sel, cnt, out, nout);
   dfree(sel);
   return rc;
}

// sorted unique union of two sorted unique lists
static int union_sorted(const int64_t *a, int na, const int64_t *b, int nb, int64_t **out, int *nout)
{
   int64_t *cat;
   HDK_TRY(dalloc(&cat, (size_t)na + nb + 1));
   if (na) HDK_CUDA(cudaMemcpyAsync(cat, a, sizeof(int64_t) * (size_t)na, cudaMemcpyDeviceToDevice, g.stream));
   if (nb) HDK_CUDA(cudaMemcpyAsync(cat + na, b, sizeof(int64_t) * (size_t)nb, cudaMemcpyDeviceToDevice, g.stream));
   int rc = sorted_unique_i64(cat, (int64_t)na + nb, out, nout);
   dfree(cat);
   return rc;
}

// ---------------------------------------------------------------------------------------------
// extended index space: [ghosts below | owned rows lo..hi | ghosts above], order-preserving
// ---------------------------------------------------------------------------------------------
struct ExtMap
{
   int64_t        lo = 0, hi = 0; // owned global range [lo, hi)
   int            n_own = 0, nb = 0, m = 0; // owned count, ghosts below lo, ghost count
   const int64_t *ids = nullptr;  // device, sorted unique ghost ids (m)
   int            n_ext() const { return n_own + m; }
};
__device__ __forceinline__ int ext_index(const ExtMap &e, int64_t gid)
{
   if (gid >= e.lo && gid < e.hi) return e.nb + (int)(gid - e.lo);
   int p = lower_bound_i64(e.ids, e.m, gid);
   return p < e.nb ? p : p + e.n_own;
}
__global__ void k_lower_bound_one(const int64_t *v, int n, int64_t key, int *out) { *out = lower_bound_i64(v, n, key); }

static int extmap_make(ExtMap &E, int64_t lo, int64_t hi, const int64_t *ids, int m)
{
   E.lo = lo; E.hi = hi; E.n_own = (int)(hi - lo); E.ids = ids; E.m = m; E.nb = 0;
   if (m > 0)
   {
      int *d = reinterpret_cast<int *>(g.dscal + S_TMP3);
      k_lower_bound_one<<<1, 1, 0, g.stream>>>(ids, m, lo, d);
      HDK_LAUNCH_CHECK();
      HDK_CUDA(cudaMemcpyAsync(&E.nb, d, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
      HDK_CUDA(cudaStreamSynchronize(g.stream));
   }
   return HDK_OK;
}

__global__ void k_len_own(const int64_t *ip, int n, int *len)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) len[i] = (int)(ip[i + 1] - ip[i]);
}
__global__ void k_len_ghost_ext(const int64_t *ip, const int64_t *gids, int n, ExtMap rm, int *len_ext)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k < n) len_ext[ext_index(rm, gids[k])] = (int)(ip[k + 1] - ip[k]);
}
// owned rows are contiguous and in the same order in both layouts: entry-parallel copy + column map
__global__ void k_fill_own_ext(const int64_t *col, const double *val, int64_t nnz, int64_t base, ExtMap cm, int *ecol, double *eval)
{
   int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
   if (k >= nnz) return;
   ecol[base + k] = ext_index(cm, col[k]);
   if (eval) eval[base + k] = val[k];
}
__global__ void k_fill_ghost_ext(const int64_t *ip, const int64_t *col, const double *val, const int64_t *gids, int n,
                                 ExtMap rm, ExtMap cm, const int *erp, int *ecol, double *eval)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= n) return;
   int p = erp[ext_index(rm, gids[k])];
   for (int64_t j = ip[k]; j < ip[k + 1]; j++, p++)
   {
      ecol[p] = ext_index(cm, col[j]);
      if (eval) eval[p] = val[j];
   }
}

// E: rm.n_ext() rows; owned rows from `own`, ghost rows `gh` (ids `gh_ids`, a subset of rm.ids), all
// other rows empty; columns renumbered through cm (every column must be inside cm's range or list)
static int ext_assemble(const ExtMap &rm, const ExtMap &cm, const GRows &own, const GRows *gh, const int64_t *gh_ids, DevCSR &E)
{
   const int ne = rm.n_ext();
   int      *len;
   HDK_TRY(dalloc(&len, (size_t)ne + 1));
   HDK_CUDA(cudaMemsetAsync(len, 0, sizeof(int) * ((size_t)ne + 1), g.stream));
   if (own.n > 0) { k_len_own<<<cdiv(own.n, 256), 256, 0, g.stream>>>(own.ip, own.n, len + rm.nb); HDK_LAUNCH_CHECK(); }
   if (gh && gh->n > 0) { k_len_ghost_ext<<<cdiv(gh->n, 256), 256, 0, g.stream>>>(gh->ip, gh_ids, gh->n, rm, len); HDK_LAUNCH_CHECK(); }
   E = DevCSR();
   E.nrows = ne; E.ncols = cm.n_ext(); E.owns = true;
   HDK_TRY(dalloc(&E.rowptr, (size_t)ne + 1));
   HDK_TRY(exclusive_scan_int(len, E.rowptr, ne + 1));
   int tot[2] = {0, 0};
   HDK_CUDA(cudaMemcpyAsync(&tot[0], E.rowptr + ne, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaMemcpyAsync(&tot[1], E.rowptr + rm.nb, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   dfree(len);
   if ((int64_t)own.nnz + (gh ? gh->nnz : 0) != (int64_t)tot[0])
      return set_error(HDK_ERR_INVALID, "extended block: %lld entries expected, %d placed", (long long)(own.nnz + (gh ? gh->nnz : 0)), tot[0]);
   E.nnz = tot[0];
   HDK_TRY(dalloc(&E.col, (size_t)E.nnz + 8));
   HDK_TRY(dalloc(&E.val, (size_t)E.nnz + 8));
   HDK_CUDA(cudaMemsetAsync(E.col + E.nnz, 0, sizeof(int) * 8, g.stream));
   HDK_CUDA(cudaMemsetAsync(E.val + E.nnz, 0, sizeof(double) * 8, g.stream));
   if (own.nnz > 0)
   {
      k_fill_own_ext<<<cdiv(own.nnz, 256), 256, 0, g.stream>>>(own.col, own.val, own.nnz, (int64_t)tot[1], cm, E.col, E.val);
      HDK_LAUNCH_CHECK();
   }
   if (gh && gh->n > 0)
   {
      k_fill_ghost_ext<<<cdiv(gh->n, 128), 128, 0, g.stream>>>(gh->ip, gh->col, gh->val, gh_ids, gh->n, rm, cm, E.rowptr, E.col, E.val);
      HDK_LAUNCH_CHECK();
   }
   return HDK_OK;
}

// rows [row_lo, row_lo + n) of an extended result block -> GRows with global columns (cmap: local column -> global id)
__global__ void k_ext_rows_ip(const int *erp, int row_lo, int n, int64_t *ip)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i <= n) ip[i] = (int64_t)erp[row_lo + i] - (int64_t)erp[row_lo];
}
__global__ void k_ext_rows_entries(const int *ecol, const double *eval, int64_t base, int64_t nnz, const int64_t *cmap, int64_t *col, double *val)
{
   int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
   if (k >= nnz) return;
   col[k] = cmap[ecol[base + k]];
   val[k] = eval[base + k];
}
static int ext_rows_to_grows(const DevCSR &E, int row_lo, int n, int64_t row0, const int64_t *cmap, GRows &G)
{
   int b[2] = {0, 0};
   HDK_CUDA(cudaMemcpyAsync(&b[0], E.rowptr + row_lo, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaMemcpyAsync(&b[1], E.rowptr + row_lo + n, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   G = GRows();
   G.n = n; G.row0 = row0; G.nnz = (int64_t)b[1] - b[0];
   HDK_TRY(dalloc(&G.ip, (size_t)n + 1));
   HDK_TRY(dalloc(&G.col, (size_t)G.nnz + 1));
   HDK_TRY(dalloc(&G.val, (size_t)G.nnz + 1));
   k_ext_rows_ip<<<cdiv(n + 1, 256), 256, 0, g.stream>>>(E.rowptr, row_lo, n, G.ip);
   HDK_LAUNCH_CHECK();
   if (G.nnz > 0)
   {
      k_ext_rows_entries<<<cdiv(G.nnz, 256), 256, 0, g.stream>>>(E.col, E.val, (int64_t)b[0], G.nnz, cmap, G.col, G.val);
      HDK_LAUNCH_CHECK();
   }
   return HDK_OK;
}
// identity-with-list column map of an ExtMap: cmap[ext index] = global id
__global__ void k_extmap_ids(ExtMap e, int64_t *cmap)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= e.n_own + e.m) return;
   int64_t gid;
   if (i < e.nb) gid = e.ids[i];
   else if (i < e.nb + e.n_own) gid = e.lo + (i - e.nb);
   else gid = e.ids[i - e.n_own];
   cmap[i] = gid;
}

// ---------------------------------------------------------------------------------------------
// request plans: "send me what you own for these global ids" and the reverse direction
// ---------------------------------------------------------------------------------------------
struct IdPlan
{
   int                  m = 0;          // ids I ask for (sorted, so grouped by owner)
   std::vector<int64_t> woff, soff;     // element offsets per rank (nranks + 1): what I receive / what I serve
   int                  nserve = 0;
   int                 *serve_idx = nullptr; // device: local index of every id another rank asked me for
};
static void idplan_free(IdPlan &P) { dfree(P.serve_idx); P = IdPlan(); }

__global__ void k_ids_local(const int64_t *ids, int n, int64_t lo, int64_t hi, int *idx, int *bad)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   int64_t v = ids[i];
   if (v < lo || v >= hi) { atomicAdd(bad, 1); idx[i] = 0; }
   else idx[i] = (int)(v - lo);
}

static void scaled(const std::vector<int64_t> &off, int64_t elem, std::vector<int64_t> &out)
{
   out.resize(off.size());
   for (size_t i = 0; i < off.size(); i++) out[i] = off[i] * elem;
}

// collective.  ids: device, sorted unique, none owned by me.  starts: nranks + 1 partition boundaries.
static int idplan_build(IdPlan &P, const int64_t *ids_d, int m, const std::vector<int64_t> &starts)
{
   const int R = g.nranks, me = g.rank;
   P = IdPlan();
   P.m = m;
   std::vector<int64_t> ids((size_t)m);
   if (m > 0)
   {
      HDK_CUDA(cudaMemcpyAsync(ids.data(), ids_d, sizeof(int64_t) * (size_t)m, cudaMemcpyDeviceToHost, g.stream));
      HDK_CUDA(cudaStreamSynchronize(g.stream));
   }
   std::vector<int> want((size_t)R, 0);
   for (int i = 0; i < m; i++)
   {
      int r = (int)(std::upper_bound(starts.begin(), starts.end(), ids[(size_t)i]) - starts.begin()) - 1;
      if (r < 0 || r >= R || r == me)
         return set_error(HDK_ERR_INVALID, "requested id %lld is outside the partition or owned by the requester", (long long)ids[(size_t)i]);
      want[(size_t)r]++;
   }
   std::vector<int> all;
   HDK_TRY(allgather_i32_host(want.data(), R, all));
   P.woff.assign((size_t)R + 1, 0);
   P.soff.assign((size_t)R + 1, 0);
   for (int r = 0; r < R; r++)
   {
      P.woff[(size_t)r + 1] = P.woff[(size_t)r] + want[(size_t)r];
      P.soff[(size_t)r + 1] = P.soff[(size_t)r] + all[(size_t)r * R + me];
   }
   P.nserve = (int)P.soff[(size_t)R];
   int64_t *req;
   HDK_TRY(dalloc(&req, (size_t)P.nserve + 1));
   HDK_TRY(dalloc(&P.serve_idx, (size_t)P.nserve + 1));
   std::vector<int64_t> sb, rb;
   scaled(P.woff, 8, sb); scaled(P.soff, 8, rb);
   HDK_TRY(alltoallv_bytes(ids_d, sb.data(), req, rb.data()));
   if (P.nserve > 0)
   {
      int *bad = reinterpret_cast<int *>(g.dscal + S_TMP3), hbad = 0;
      HDK_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), g.stream));
      k_ids_local<<<cdiv(P.nserve, 256), 256, 0, g.stream>>>(req, P.nserve, starts[(size_t)me], starts[(size_t)me + 1], P.serve_idx, bad);
      HDK_LAUNCH_CHECK();
      HDK_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
      HDK_CUDA(cudaStreamSynchronize(g.stream));
      if (hbad) return set_error(HDK_ERR_COMM, "%d requested ids do not belong to this rank", hbad);
   }
   dfree(req);
   return HDK_OK;
}

__global__ void k_gather4(const int *src, const int *idx, int n, int *dst)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_gather8(const int64_t *src, const int *idx, int n, int64_t *dst)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) dst[i] = src[idx[i]];
}

// ghost[k] = owner's owned[id_k]   (elem = 4 or 8 bytes)
static int idplan_fetch(const IdPlan &P, const void *owned, void *ghost, int elem)
{
   char *sbuf;
   HDK_TRY(dalloc(&sbuf, (size_t)(P.nserve + 1) * 8));
   if (P.nserve > 0)
   {
      if (elem == 4) k_gather4<<<cdiv(P.nserve, 256), 256, 0, g.stream>>>((const int *)owned, P.serve_idx, P.nserve, (int *)sbuf);
      else k_gather8<<<cdiv(P.nserve, 256), 256, 0, g.stream>>>((const int64_t *)owned, P.serve_idx, P.nserve, (int64_t *)sbuf);
      HDK_LAUNCH_CHECK();
   }
   std::vector<int64_t> sb, rb;
   scaled(P.soff, elem, sb); scaled(P.woff, elem, rb);
   int rc = alltoallv_bytes(sbuf, sb.data(), ghost, rb.data());
   dfree(sbuf);
   return rc;
}

__global__ void k_reverse_add(const int *recv, const int *idx, int n, int *owned)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n && recv[i]) atomicAdd(owned + idx[i], recv[i]);
}
// PMIS: a neighbour rank cleared the mark of one of my candidates (its copy went from 1 to 0)
__global__ void k_reverse_clear(const int *recv, const int *idx, int n, int *cf, const double *measure)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   int j = idx[i];
   if (recv[i] == 0 && measure[j] > 1.0) cf[j] = 0;
}
// the ghosts' values travel back to their owners; op 0: owned += value, op 1: PMIS clear
static int idplan_reverse_i32(const IdPlan &P, const int *ghost, int *owned, int op, const double *measure)
{
   int *rbuf;
   HDK_TRY(dalloc(&rbuf, (size_t)P.nserve + 1));
   std::vector<int64_t> sb, rb;
   scaled(P.woff, 4, sb); scaled(P.soff, 4, rb);
   HDK_TRY(alltoallv_bytes(ghost, sb.data(), rbuf, rb.data()));
   if (P.nserve > 0)
   {
      if (op == 0) k_reverse_add<<<cdiv(P.nserve, 256), 256, 0, g.stream>>>(rbuf, P.serve_idx, P.nserve, owned);
      else k_reverse_clear<<<cdiv(P.nserve, 256), 256, 0, g.stream>>>(rbuf, P.serve_idx, P.nserve, owned, measure);
      HDK_LAUNCH_CHECK();
   }
   dfree(rbuf);
   return HDK_OK;
}

__global__ void k_serve_len(const int64_t *ip, const int *idx, int n, int *len)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k <= n) len[k] = (k < n) ? (int)(ip[idx[k] + 1] - ip[idx[k]]) : 0;
}
__global__ void k_serve_pack(const int64_t *ip, const int64_t *col, const double *val, const int *idx, int n,
                             const int64_t *spos, int64_t *scol, double *sval)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= n) return;
   int64_t p = spos[k];
   for (int64_t j = ip[idx[k]]; j < ip[idx[k] + 1]; j++, p++) { scol[p] = col[j]; sval[p] = val[j]; }
}
__global__ void k_pick_i64(const int64_t *v, const int64_t *where, int n, int64_t *out)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) out[i] = v[where[i]];
}

// the rows of `own` for the plan's ids, complete (global columns, owner's storage order)
static int idplan_fetch_rows(const IdPlan &P, const GRows &own, GRows &gh)
{
   const int R = g.nranks;
   if (R + 1 > 64) return set_error(HDK_ERR_UNSUPPORTED, "more than 63 ranks");
   gh = GRows();
   gh.n = P.m;
   // 1. row lengths
   int *slen, *glen;
   int64_t *spos;
   HDK_TRY(dalloc(&slen, (size_t)P.nserve + 1));
   HDK_TRY(dalloc(&glen, (size_t)P.m + 1));
   HDK_TRY(dalloc(&spos, (size_t)P.nserve + 1));
   HDK_TRY(dalloc(&gh.ip, (size_t)P.m + 1));
   k_serve_len<<<cdiv(P.nserve + 1, 256), 256, 0, g.stream>>>(own.ip, P.serve_idx, P.nserve, slen);
   HDK_LAUNCH_CHECK();
   {
      std::vector<int64_t> sb, rb;
      scaled(P.soff, 4, sb); scaled(P.woff, 4, rb);
      HDK_TRY(alltoallv_bytes(slen, sb.data(), glen, rb.data()));
   }
   HDK_CUDA(cudaMemsetAsync(glen + P.m, 0, sizeof(int), g.stream));
   HDK_TRY(exclusive_scan_i64(slen, spos, P.nserve + 1));
   HDK_TRY(exclusive_scan_i64(glen, gh.ip, P.m + 1));
   // 2. entry offsets at the rank boundaries of both sides
   std::vector<int64_t> bnd((size_t)2 * (R + 1)), hb((size_t)2 * (R + 1));
   for (int r = 0; r <= R; r++) { bnd[(size_t)r] = P.soff[(size_t)r]; bnd[(size_t)(R + 1 + r)] = P.woff[(size_t)r]; }
   int64_t *dw, *dv;
   HDK_TRY(dalloc(&dw, (size_t)2 * (R + 1)));
   HDK_TRY(dalloc(&dv, (size_t)2 * (R + 1)));
   HDK_CUDA(cudaMemcpyAsync(dw, bnd.data(), sizeof(int64_t) * bnd.size(), cudaMemcpyHostToDevice, g.stream));
   k_pick_i64<<<1, 64, 0, g.stream>>>(spos, dw, R + 1, dv);
   HDK_LAUNCH_CHECK();
   k_pick_i64<<<1, 64, 0, g.stream>>>(gh.ip, dw + (R + 1), R + 1, dv + (R + 1));
   HDK_LAUNCH_CHECK();
   HDK_CUDA(cudaMemcpyAsync(hb.data(), dv, sizeof(int64_t) * hb.size(), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   dfree(dw); dfree(dv);
   std::vector<int64_t> se(hb.begin(), hb.begin() + (R + 1)), ge(hb.begin() + (R + 1), hb.end());
   const int64_t ns = se[(size_t)R];
   gh.nnz = ge[(size_t)R];
   // 3. pack and exchange the entries
   int64_t *scol;
   double  *sval;
   HDK_TRY(dalloc(&scol, (size_t)ns + 1));
   HDK_TRY(dalloc(&sval, (size_t)ns + 1));
   HDK_TRY(dalloc(&gh.col, (size_t)gh.nnz + 1));
   HDK_TRY(dalloc(&gh.val, (size_t)gh.nnz + 1));
   if (P.nserve > 0)
   {
      k_serve_pack<<<cdiv(P.nserve, 128), 128, 0, g.stream>>>(own.ip, own.col, own.val, P.serve_idx, P.nserve, spos, scol, sval);
      HDK_LAUNCH_CHECK();
   }
   std::vector<int64_t> sb, rb;
   scaled(se, 8, sb); scaled(ge, 8, rb);
   HDK_TRY(alltoallv_bytes(scol, sb.data(), gh.col, rb.data()));
   HDK_TRY(alltoallv_bytes(sval, sb.data(), gh.val, rb.data()));
   dfree(scol); dfree(sval); dfree(slen); dfree(glen); dfree(spos);
   return HDK_OK;
}

// ---------------------------------------------------------------------------------------------
// distributed transpose: R rows = my coarse points, columns = global fine ids, ascending
// ---------------------------------------------------------------------------------------------
struct Trip { int64_t c, i; double v; };

__device__ __forceinline__ int owner_of(const int64_t *starts, int R, int64_t id)
{
   int lo = 0, hi = R; // last r with starts[r] <= id
   while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (starts[mid] <= id) lo = mid; else hi = mid; }
   return lo;
}
// only entries whose coarse column lives on another rank travel as triples; the local ones feed the
// counting transpose directly
__global__ void k_trip_count(const int64_t *ip, const int64_t *col, int n, const int64_t *cstarts, int R, int64_t cs, int64_t ce, int *cnt)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   for (int64_t k = ip[i]; k < ip[i + 1]; k++)
   {
      int64_t c = col[k];
      if (c < cs || c >= ce) atomicAdd(cnt + owner_of(cstarts, R, c), 1);
   }
}
__global__ void k_trip_scatter(const int64_t *ip, const int64_t *col, const double *val, int n, int64_t row0,
                               const int64_t *cstarts, int R, int64_t cs, int64_t ce, const int64_t *soff, int *cursor, Trip *out)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   for (int64_t k = ip[i]; k < ip[i + 1]; k++)
   {
      int64_t c = col[k];
      if (c >= cs && c < ce) continue;
      int  d = owner_of(cstarts, R, c);
      Trip t;
      t.c = c; t.i = row0 + i; t.v = val[k];
      out[soff[d] + atomicAdd(cursor + d, 1)] = t;
   }
}
__global__ void k_trip_rowcount(const Trip *t, int64_t n, int64_t cs, int *cnt)
{
   int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
   if (k < n) atomicAdd(cnt + (int)(t[k].c - cs), 1);
}
__global__ void k_local_rowcount(const int64_t *col, int64_t nnz, int64_t cs, int64_t ce, int *cnt)
{
   int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
   if (k >= nnz) return;
   int64_t c = col[k];
   if (c >= cs && c < ce) atomicAdd(cnt + (int)(c - cs), 1);
}
__global__ void k_trip_place(const Trip *t, int64_t n, int64_t cs, const int64_t *rip, int *cursor, int64_t *col, double *val)
{
   int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
   if (k >= n) return;
   int     r = (int)(t[k].c - cs);
   int64_t p = rip[r] + atomicAdd(cursor + r, 1);
   col[p] = t[k].i; val[p] = t[k].v;
}
__global__ void k_local_place(const int64_t *ip, const int64_t *col, const double *val, int n, int64_t row0, int64_t cs, int64_t ce,
                              const int64_t *rip, int *cursor, int64_t *rcol, double *rval)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   for (int64_t k = ip[i]; k < ip[i + 1]; k++)
   {
      int64_t c = col[k];
      if (c < cs || c >= ce) continue;
      int     r = (int)(c - cs);
      int64_t p = rip[r] + atomicAdd(cursor + r, 1);
      rcol[p] = row0 + i; rval[p] = val[k];
   }
}
// insertion sort of every row by (unique) fine index: hypre_CSRMatrixTranspose order
__global__ void k_sort_rows_i64(const int64_t *ip, int n, int64_t *col, double *val)
{
   int r = blockIdx.x * blockDim.x + threadIdx.x;
   if (r >= n) return;
   int64_t b = ip[r], e = ip[r + 1];
   for (int64_t a = b + 1; a < e; a++)
   {
      int64_t c = col[a]; double v = val[a];
      int64_t j = a - 1;
      while (j >= b && col[j] > c) { col[j + 1] = col[j]; val[j + 1] = val[j]; j--; }
      col[j + 1] = c; val[j + 1] = v;
   }
}

// P: my fine rows with global coarse columns.  Rt: rows [cs, ce) of P^T with global fine columns.
static int dist_transpose(const GRows &P, const std::vector<int64_t> &cstarts, GRows &Rt)
{
   const int     R = g.nranks, me = g.rank;
   const int64_t cs = cstarts[(size_t)me], ce = cstarts[(size_t)me + 1];
   const int     nc = (int)(ce - cs);
   int64_t *dstarts;
   int     *cnt;
   HDK_TRY(dalloc(&dstarts, (size_t)R + 1));
   HDK_TRY(dalloc(&cnt, (size_t)2 * R + 2));
   HDK_CUDA(cudaMemcpyAsync(dstarts, cstarts.data(), sizeof(int64_t) * ((size_t)R + 1), cudaMemcpyHostToDevice, g.stream));
   HDK_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)2 * R + 2), g.stream));
   if (P.n > 0 && R > 1) { k_trip_count<<<cdiv(P.n, 256), 256, 0, g.stream>>>(P.ip, P.col, P.n, dstarts, R, cs, ce, cnt); HDK_LAUNCH_CHECK(); }
   std::vector<int> hc((size_t)R, 0), all;
   HDK_CUDA(cudaMemcpyAsync(hc.data(), cnt, sizeof(int) * (size_t)R, cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   HDK_TRY(allgather_i32_host(hc.data(), R, all));
   std::vector<int64_t> so((size_t)R + 1, 0), ro((size_t)R + 1, 0);
   for (int r = 0; r < R; r++)
   {
      so[(size_t)r + 1] = so[(size_t)r] + hc[(size_t)r];
      ro[(size_t)r + 1] = ro[(size_t)r] + all[(size_t)r * R + me];
   }
   const int64_t nsend = so[(size_t)R], nrecv = ro[(size_t)R];
   Trip    *sbuf, *rbuf;
   int64_t *dso;
   HDK_TRY(dalloc(&sbuf, (size_t)nsend + 1));
   HDK_TRY(dalloc(&rbuf, (size_t)nrecv + 1));
   HDK_TRY(dalloc(&dso, (size_t)R + 1));
   HDK_CUDA(cudaMemcpyAsync(dso, so.data(), sizeof(int64_t) * ((size_t)R + 1), cudaMemcpyHostToDevice, g.stream));
   if (nsend > 0)
   {
      k_trip_scatter<<<cdiv(P.n, 256), 256, 0, g.stream>>>(P.ip, P.col, P.val, P.n, P.row0, dstarts, R, cs, ce, dso, cnt + R + 1, sbuf);
      HDK_LAUNCH_CHECK();
   }
   std::vector<int64_t> sb, rb;
   scaled(so, (int64_t)sizeof(Trip), sb); scaled(ro, (int64_t)sizeof(Trip), rb);
   HDK_TRY(alltoallv_bytes(sbuf, sb.data(), rbuf, rb.data()));
   // counting transpose of my local entries and of the triples I received
   int *rcnt, *cursor;
   HDK_TRY(dalloc(&rcnt, (size_t)nc + 1));
   HDK_TRY(dalloc(&cursor, (size_t)nc + 1));
   HDK_CUDA(cudaMemsetAsync(rcnt, 0, sizeof(int) * ((size_t)nc + 1), g.stream));
   HDK_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * ((size_t)nc + 1), g.stream));
   if (P.nnz > 0) { k_local_rowcount<<<cdiv(P.nnz, 256), 256, 0, g.stream>>>(P.col, P.nnz, cs, ce, rcnt); HDK_LAUNCH_CHECK(); }
   if (nrecv > 0) { k_trip_rowcount<<<cdiv(nrecv, 256), 256, 0, g.stream>>>(rbuf, nrecv, cs, rcnt); HDK_LAUNCH_CHECK(); }
   Rt = GRows();
   Rt.n = nc; Rt.row0 = cs;
   HDK_TRY(dalloc(&Rt.ip, (size_t)nc + 1));
   HDK_TRY(exclusive_scan_i64(rcnt, Rt.ip, nc + 1));
   HDK_CUDA(cudaMemcpyAsync(&Rt.nnz, Rt.ip + nc, sizeof(int64_t), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream)); // also: so / dstarts host buffers are no longer read
   if (Rt.nnz != P.nnz - nsend + nrecv)
      return set_error(HDK_ERR_INVALID, "transpose: %lld entries placed, %lld expected", (long long)Rt.nnz, (long long)(P.nnz - nsend + nrecv));
   HDK_TRY(dalloc(&Rt.col, (size_t)Rt.nnz + 1));
   HDK_TRY(dalloc(&Rt.val, (size_t)Rt.nnz + 1));
   if (P.n > 0 && P.nnz > 0)
   {
      k_local_place<<<cdiv(P.n, 256), 256, 0, g.stream>>>(P.ip, P.col, P.val, P.n, P.row0, cs, ce, Rt.ip, cursor, Rt.col, Rt.val);
      HDK_LAUNCH_CHECK();
   }
   if (nrecv > 0) { k_trip_place<<<cdiv(nrecv, 256), 256, 0, g.stream>>>(rbuf, nrecv, cs, Rt.ip, cursor, Rt.col, Rt.val); HDK_LAUNCH_CHECK(); }
   if (nc > 0 && Rt.nnz > 0) { k_sort_rows_i64<<<cdiv(nc, 128), 128, 0, g.stream>>>(Rt.ip, nc, Rt.col, Rt.val); HDK_LAUNCH_CHECK(); }
   dfree(sbuf); dfree(rbuf); dfree(dso); dfree(dstarts); dfree(cnt); dfree(rcnt); dfree(cursor);
   return HDK_OK;
}

// ---------------------------------------------------------------------------------------------
// small kernels of the level driver
// ---------------------------------------------------------------------------------------------
__global__ void k_cf_pos(const int *cf, int n, int *flag)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i <= n) flag[i] = (i < n && cf[i] > 0) ? 1 : 0;
}
// global coarse id of my C points (cs + local rank among my C points), -1 for F points
__global__ void k_gcid_own(const int *cf_ext, const int *f2c_ext, int nb, int n, int64_t cs, int64_t *gcid)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   gcid[i] = cf_ext[nb + i] > 0 ? cs + (int64_t)(f2c_ext[nb + i] - f2c_ext[nb]) : -1;
}
// cmap[local coarse index] = global coarse id for every C point of the extended set
__global__ void k_cmap_fill(const int *cf_ext, const int *f2c_ext, const int64_t *gcid_own, const int64_t *gcid_gh,
                            int nb, int n_own, int n_ext, int64_t *cmap)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n_ext || cf_ext[i] <= 0) return;
   int64_t v;
   if (i < nb) v = gcid_gh[i];
   else if (i < nb + n_own) v = gcid_own[i - nb];
   else v = gcid_gh[i - n_own];
   cmap[f2c_ext[i]] = v;
}
// ghost <-> extended layout of per-node arrays: ghosts live at [0, nb) and [nb + n_own, n_ext)
template <class T>
__global__ void k_ghost_scatter(const T *gh, int m, int nb, int n_own, T *ext)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k < m) ext[k < nb ? k : k + n_own] = gh[k];
}
template <class T>
__global__ void k_ghost_gather(const T *ext, int m, int nb, int n_own, T *gh)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k < m) gh[k] = ext[k < nb ? k : k + n_own];
}
__global__ void k_ghost_measure_zero(const int *cf_ext, int m, int nb, int n_own, double *measure_ext)
{
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= m) return;
   int e = k < nb ? k : k + n_own;
   if (cf_ext[e] != 0) measure_ext[e] = 0.0;
}
__global__ void k_copy_i64(const int64_t *in, int64_t *out, int64_t n, int64_t shift)
{
   int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
   if (i < n) out[i] = in[i] + shift;
}

// forward exchange of a per-node array of the extended layout: ghosts <- owners
template <class T>
static int ext_forward(const IdPlan &P, const ExtMap &E, T *ext, T *gh_tmp)
{
   HDK_TRY(idplan_fetch(P, ext + E.nb, gh_tmp, (int)sizeof(T)));
   if (E.m > 0) { k_ghost_scatter<T><<<cdiv(E.m, 256), 256, 0, g.stream>>>(gh_tmp, E.m, E.nb, E.n_own, ext); HDK_LAUNCH_CHECK(); }
   return HDK_OK;
}

static int grows_to_parcsr(const GRows &G, int64_t cs, int64_t ce, int64_t grows, int64_t gcols, bool square, bool distributed,
                           hdk_csr_s **out)
{
   return parcsr_build(G.row0, G.row0 + G.n - 1, cs, ce - 1, grows, gcols, square, distributed, false, G.ip, G.col, G.val, out, true);
}

static int keep_debug_copy(AmgLevel &L, int which, const GRows &G)
{
   HDK_TRY(dalloc(&L.dbg_ip[which], (size_t)G.n + 1));
   HDK_TRY(dalloc(&L.dbg_col[which], (size_t)G.nnz + 1));
   HDK_TRY(dalloc(&L.dbg_val[which], (size_t)G.nnz + 1));
   HDK_CUDA(cudaMemcpyAsync(L.dbg_ip[which], G.ip, sizeof(int64_t) * ((size_t)G.n + 1), cudaMemcpyDeviceToDevice, g.stream));
   if (G.nnz) HDK_CUDA(cudaMemcpyAsync(L.dbg_col[which], G.col, sizeof(int64_t) * (size_t)G.nnz, cudaMemcpyDeviceToDevice, g.stream));
   if (G.nnz) HDK_CUDA(cudaMemcpyAsync(L.dbg_val[which], G.val, sizeof(double) * (size_t)G.nnz, cudaMemcpyDeviceToDevice, g.stream));
   L.dbg_nnz[which] = G.nnz; L.row0 = G.row0;
   return HDK_OK;
}

// gather the rows of every rank (small level) into one serial ParCSR on every rank
static int gather_level(const GRows &Gm, const std::vector<int64_t> &starts, hdk_csr_s **out)
{
   const int R = g.nranks, me = g.rank;
   const int64_t N = starts[(size_t)R];
   std::vector<int64_t> nnz_all;
   HDK_TRY(allgather_i64_host(Gm.nnz, nnz_all));
   std::vector<int64_t> koff((size_t)R + 1, 0);
   for (int r = 0; r < R; r++) koff[(size_t)r + 1] = koff[(size_t)r] + nnz_all[(size_t)r];
   const int64_t NNZ = koff[(size_t)R];
   if (N > 2000000000LL || NNZ > 2000000000LL)
      return set_error(HDK_ERR_UNSUPPORTED, "replicated level too large (%lld rows, %lld nnz): lower replicate_rows", (long long)N, (long long)NNZ);
   int64_t *gip, *gcj;
   double  *gva;
   HDK_TRY(dalloc(&gip, (size_t)N + 1));
   HDK_TRY(dalloc(&gcj, (size_t)NNZ + 1));
   HDK_TRY(dalloc(&gva, (size_t)NNZ + 1));
   if (Gm.n > 0)
   {
      k_copy_i64<<<cdiv(Gm.n, 256), 256, 0, g.stream>>>(Gm.ip, gip + starts[(size_t)me], Gm.n, koff[(size_t)me]);
      HDK_LAUNCH_CHECK();
   }
   if (Gm.nnz > 0)
   {
      HDK_CUDA(cudaMemcpyAsync(gcj + koff[(size_t)me], Gm.col, sizeof(int64_t) * (size_t)Gm.nnz, cudaMemcpyDeviceToDevice, g.stream));
      HDK_CUDA(cudaMemcpyAsync(gva + koff[(size_t)me], Gm.val, sizeof(double) * (size_t)Gm.nnz, cudaMemcpyDeviceToDevice, g.stream));
   }
   HDK_CUDA(cudaMemcpyAsync(gip + N, &NNZ, sizeof(int64_t), cudaMemcpyHostToDevice, g.stream));
   std::vector<int64_t> ob((size_t)R + 1);
   for (int r = 0; r <= R; r++) ob[(size_t)r] = starts[(size_t)r] * 8;
   HDK_TRY(allgatherv_bytes(gip, ob.data()));
   for (int r = 0; r <= R; r++) ob[(size_t)r] = koff[(size_t)r] * 8;
   HDK_TRY(allgatherv_bytes(gcj, ob.data()));
   HDK_TRY(allgatherv_bytes(gva, ob.data()));
   HDK_CUDA(cudaStreamSynchronize(g.stream)); // NNZ (host) is read by the copy above
   int rc = parcsr_build(0, N - 1, 0, N - 1, N, N, true, false, false, gip, gcj, gva, out, true);
   dfree(gip); dfree(gcj); dfree(gva);
   return rc;
}

} // namespace hdk

using namespace hdk;

// One level of the distributed setup: from my rows of A_l to my rows of P_l, R_l and A_{l+1}.
static int dist_level(const hdk_amg_params *prm, const GRows &A, const hdk_csr_s *Apar, const std::vector<int64_t> &fstarts,
                      int level, bool keep_debug, int *cf_out /* n */, std::vector<int64_t> &cstarts, GRows &P, GRows &Rt,
                      GRows &C, bool &stop)
{
   const int     R = g.nranks, me = g.rank;
   const int64_t fs = fstarts[(size_t)me], fe = fstarts[(size_t)me + 1];
   const int     n = (int)(fe - fs);
   stop = false;
   setup_stage_mark(nullptr, level);
   // ---- ring 1: rows of my off-rank columns -------------------------------------------------
   const int64_t *g1 = Apar->halo.col_map;
   const int      n1 = Apar->halo.n_halo;
   IdPlan         F1;
   GRows          A1;
   HDK_TRY(idplan_build(F1, g1, n1, fstarts));
   HDK_TRY(idplan_fetch_rows(F1, A, A1));
   // ---- ring 2: their off-rank columns; U = ring 1 + ring 2 ------------------------------------
   int64_t *g2 = nullptr, *U = nullptr;
   int      n2 = 0, m = 0;
   HDK_TRY(collect_foreign(A1.col, A1.nnz, fs, fe, g1, n1, nullptr, 0, &g2, &n2));
   HDK_TRY(union_sorted(g1, n1, g2, n2, &U, &m));
   dfree(g2);
   ExtMap EU;
   HDK_TRY(extmap_make(EU, fs, fe, U, m));
   const int ne = EU.n_ext(), nb = EU.nb;
   IdPlan    FU;
   HDK_TRY(idplan_build(FU, U, m, fstarts));
   DevCSR EA, ES;
   HDK_TRY(ext_assemble(EU, EU, A, &A1, g1, EA));
   grows_free(A1);
   idplan_free(F1);
   setup_stage_mark("halo rows", level);
   // ---- strength (row-local, exact for owned and ring-1 rows) ------------------------------------
   HDK_TRY(build_strength_csr(EA, nullptr, nullptr, prm->strong_th, prm->max_row_sum, ES));
   setup_stage_mark("strength", level);
   // ---- PMIS ---------------------------------------------------------------------------------
   int    *cf, *cnt, *gi;
   double *measure, *gd;
   HDK_TRY(dalloc(&cf, (size_t)ne + 1));
   HDK_TRY(dalloc(&cnt, (size_t)ne + 1));
   HDK_TRY(dalloc(&measure, (size_t)ne + 1));
   HDK_TRY(dalloc(&gi, (size_t)m + 1));
   HDK_TRY(dalloc(&gd, (size_t)m + 1));
   HDK_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)ne + 1), g.stream));
   HDK_CUDA(cudaMemsetAsync(cf, 0, sizeof(int) * ((size_t)ne + 1), g.stream));
   HDK_CUDA(cudaMemsetAsync(measure, 0, sizeof(double) * ((size_t)ne + 1), g.stream));
   HDK_TRY(pmis_count_cols(ES, nb, n, cnt)); // my rows' strong columns; counts of ghost columns go back to their owners
   if (m > 0) { k_ghost_gather<int><<<cdiv(m, 256), 256, 0, g.stream>>>(cnt, m, nb, n, gi); HDK_LAUNCH_CHECK(); }
   HDK_TRY(idplan_reverse_i32(FU, gi, cnt + nb, 0, nullptr));
   HDK_TRY(pmis_measure(cnt + nb, n, prm->rand_seed, fs, measure + nb)); // random part indexed by the global row
   HDK_TRY(pmis_init(ES, nb, n, cf, measure));
   HDK_TRY(ext_forward<int>(FU, EU, cf, gi));
   HDK_TRY(ext_forward<double>(FU, EU, measure, gd));
   {
      int *rem = reinterpret_cast<int *>(g.dscal + S_TMP2);
      int  hrem = 1, iters = 0;
      while (hrem > 0 && iters < 1000)
      {
         HDK_CUDA(cudaMemsetAsync(rem, 0, sizeof(int), g.stream));
         HDK_TRY(pmis_mark(ne, cf, measure));
         HDK_TRY(pmis_elim(ES, nb, n, cf, measure));
         if (m > 0) { k_ghost_gather<int><<<cdiv(m, 256), 256, 0, g.stream>>>(cf, m, nb, n, gi); HDK_LAUNCH_CHECK(); }
         HDK_TRY(idplan_reverse_i32(FU, gi, cf + nb, 1, measure + nb));
         HDK_TRY(ext_forward<int>(FU, EU, cf, gi));
         HDK_TRY(pmis_set(ES, nb, n, cf, measure, rem));
         HDK_TRY(ext_forward<int>(FU, EU, cf, gi));
         if (m > 0) { k_ghost_measure_zero<<<cdiv(m, 256), 256, 0, g.stream>>>(cf, m, nb, n, measure); HDK_LAUNCH_CHECK(); }
         HDK_TRY(allreduce_i32_dev(rem, 1));
         HDK_CUDA(cudaMemcpyAsync(&hrem, rem, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
         HDK_CUDA(cudaStreamSynchronize(g.stream));
         iters++;
      }
   }
   dfree(cnt); dfree(measure); dfree(gd);
   setup_stage_mark("pmis", level);
   // ---- coarse numbering --------------------------------------------------------------------------
   int *flag, *f2c;
   HDK_TRY(dalloc(&flag, (size_t)ne + 1));
   HDK_TRY(dalloc(&f2c, (size_t)ne + 1));
   k_cf_pos<<<cdiv(ne + 1, 256), 256, 0, g.stream>>>(cf, ne, flag);
   HDK_LAUNCH_CHECK();
   HDK_TRY(exclusive_scan_int(flag, f2c, ne + 1));
   int h3[3] = {0, 0, 0};
   HDK_CUDA(cudaMemcpyAsync(&h3[0], f2c + nb, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaMemcpyAsync(&h3[1], f2c + nb + n, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaMemcpyAsync(&h3[2], f2c + ne, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   dfree(flag);
   const int nc_own = h3[1] - h3[0], nc_ext = h3[2];
   std::vector<int64_t> call;
   HDK_TRY(allgather_i64_host(nc_own, call));
   cstarts.assign((size_t)R + 1, 0);
   for (int r = 0; r < R; r++) cstarts[(size_t)r + 1] = cstarts[(size_t)r] + call[(size_t)r];
   const int64_t ncg = cstarts[(size_t)R], nfg = fstarts[(size_t)R];
   const int64_t cs = cstarts[(size_t)me], ce = cstarts[(size_t)me + 1];
   if (cf_out) HDK_CUDA(cudaMemcpyAsync(cf_out, cf + nb, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, g.stream));
   if (ncg == 0 || ncg == nfg || ncg < prm->min_coarse_size)
   {
      // no coarser level (same rule as the serial setup); the marks stay as computed
      stop = true;
      csr_free(EA); csr_free(ES); dfree(cf); dfree(f2c); dfree(gi); dfree(U);
      idplan_free(FU);
      return HDK_OK;
   }
   int64_t *gcid, *gcid_gh, *cmap;
   HDK_TRY(dalloc(&gcid, (size_t)n + 1));
   HDK_TRY(dalloc(&gcid_gh, (size_t)m + 1));
   HDK_TRY(dalloc(&cmap, (size_t)nc_ext + 1));
   if (n > 0) { k_gcid_own<<<cdiv(n, 256), 256, 0, g.stream>>>(cf, f2c, nb, n, cs, gcid); HDK_LAUNCH_CHECK(); }
   HDK_TRY(idplan_fetch(FU, gcid, gcid_gh, 8));
   k_cmap_fill<<<cdiv(ne, 256), 256, 0, g.stream>>>(cf, f2c, gcid, gcid_gh, nb, n, ne, cmap);
   HDK_LAUNCH_CHECK();
   dfree(gcid); dfree(gcid_gh); dfree(gi);
   idplan_free(FU);
   // ---- extended+i interpolation on my rows ----------------------------------------------------
   {
      DevCSR EP;
      HDK_TRY(build_interp(EA, ES, cf, f2c, nc_ext, prm->max_nnz_row, prm->trunc_factor, EP, nb, nb + n));
      if (cf_out) HDK_CUDA(cudaMemcpyAsync(cf_out, cf + nb, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, g.stream)); // SF -> F
      HDK_TRY(ext_rows_to_grows(EP, nb, n, fs, cmap, P));
      csr_free(EP);
   }
   csr_free(EA); csr_free(ES); dfree(cf); dfree(f2c); dfree(cmap); dfree(U);
   setup_stage_mark("interp", level);
   // ---- R = P^T ------------------------------------------------------------------------------------
   HDK_TRY(dist_transpose(P, cstarts, Rt));
   setup_stage_mark("transpose", level);
   // ---- Galerkin product on my coarse rows --------------------------------------------------------
   {
      // fine points: XA = off-rank columns of my R rows (their A rows are needed),
      // V = XA + off-rank columns of my A rows and of the XA rows (their P rows are needed)
      int64_t *XA = nullptr, *x2 = nullptr, *V1 = nullptr, *V = nullptr;
      int      nxa = 0, nx2 = 0, nv1 = 0, nv = 0;
      HDK_TRY(collect_foreign(Rt.col, Rt.nnz, fs, fe, nullptr, 0, nullptr, 0, &XA, &nxa));
      IdPlan FA;
      GRows  AX;
      HDK_TRY(idplan_build(FA, XA, nxa, fstarts));
      HDK_TRY(idplan_fetch_rows(FA, A, AX));
      idplan_free(FA);
      HDK_TRY(collect_foreign(AX.col, AX.nnz, fs, fe, nullptr, 0, nullptr, 0, &x2, &nx2));
      HDK_TRY(union_sorted(XA, nxa, g1, n1, &V1, &nv1));
      HDK_TRY(union_sorted(V1, nv1, x2, nx2, &V, &nv));
      dfree(x2); dfree(V1);
      IdPlan FP;
      GRows  PX;
      HDK_TRY(idplan_build(FP, V, nv, fstarts));
      HDK_TRY(idplan_fetch_rows(FP, P, PX));
      idplan_free(FP);
      // coarse points: W = off-rank coarse columns of my P rows and of the fetched P rows
      int64_t *w1 = nullptr, *w2 = nullptr, *W = nullptr;
      int      nw1 = 0, nw2 = 0, nw = 0;
      HDK_TRY(collect_foreign(P.col, P.nnz, cs, ce, nullptr, 0, nullptr, 0, &w1, &nw1));
      HDK_TRY(collect_foreign(PX.col, PX.nnz, cs, ce, nullptr, 0, nullptr, 0, &w2, &nw2));
      HDK_TRY(union_sorted(w1, nw1, w2, nw2, &W, &nw));
      dfree(w1); dfree(w2);
      ExtMap EV, EW;
      HDK_TRY(extmap_make(EV, fs, fe, V, nv));
      HDK_TRY(extmap_make(EW, cs, ce, W, nw));
      DevCSR ER, EA2, EP2, EC;
      HDK_TRY(ext_assemble(EW, EV, Rt, nullptr, nullptr, ER));
      HDK_TRY(ext_assemble(EV, EV, A, &AX, XA, EA2));
      HDK_TRY(ext_assemble(EV, EW, P, &PX, V, EP2));
      grows_free(AX); grows_free(PX);
      HDK_TRY(build_rap(ER, EA2, EP2, EC, EW.nb, EW.nb + EW.n_own));
      csr_free(ER); csr_free(EA2); csr_free(EP2);
      int64_t *wmap;
      HDK_TRY(dalloc(&wmap, (size_t)EW.n_ext() + 1));
      k_extmap_ids<<<cdiv(EW.n_ext(), 256), 256, 0, g.stream>>>(EW, wmap);
      HDK_LAUNCH_CHECK();
      HDK_TRY(ext_rows_to_grows(EC, EW.nb, EW.n_own, cs, wmap, C));
      csr_free(EC);
      dfree(wmap); dfree(XA); dfree(V); dfree(W);
   }
   setup_stage_mark("rap", level);
   (void)keep_debug;
   return HDK_OK;
}

extern "C" int setup_distributed_rows(const hdk_csr_s *A0, const hdk_amg_params *prm, hdk_amg_s **out)
{
   const int R = g.nranks, me = g.rank;
   if (!A0->orig_indptr) return set_error(HDK_ERR_INVALID, "distributed setup needs the rows with global columns (matrix built at N > 1)");
   setup_stage_mark(nullptr, -2);
   std::vector<int64_t> rows_all;
   HDK_TRY(allgather_i64_host(A0->diag.nrows, rows_all));
   std::vector<int64_t> fstarts((size_t)R + 1, 0);
   for (int r = 0; r < R; r++) fstarts[(size_t)r + 1] = fstarts[(size_t)r] + rows_all[(size_t)r];
   if (fstarts[(size_t)me] != A0->row_start) return set_error(HDK_ERR_INVALID, "row partition is not contiguous in rank order");
   const int64_t rep_rows = tune_replicate_rows();
   const bool    keep_debug = tune_amg_keep_debug();

   hdk_amg_s *M = new hdk_amg_s();
   M->prm = *prm;
   M->keep_debug = keep_debug;
   GRows A;
   A.n = A0->diag.nrows; A.row0 = A0->row_start; A.nnz = A0->orig_nnz;
   A.ip = A0->orig_indptr; A.col = A0->orig_cols; A.val = A0->orig_vals; A.owns = false;
   const hdk_csr_s *Apar = A0;
   int    rc = HDK_OK, level = 0;
   double nnz_sum = 0.0, nnz0 = (double)A0->global_nnz;
   bool   tail_final = false; // the gathered level is the coarsest one (no further coarsening)
   {
      bool empty = false;
      for (int r = 0; r < R; r++) if (fstarts[(size_t)r + 1] <= fstarts[(size_t)r]) empty = true;
      // a small (or degenerate) fine level: the whole hierarchy is replicated
      bool distributed = !(fstarts[(size_t)R] <= rep_rows || empty || prm->max_levels <= 1);
      while (distributed && rc == HDK_OK)
      {
         const int64_t nfg = fstarts[(size_t)R];
         M->lev.emplace_back();
         {
            AmgLevel &L = M->lev.back();
            L.A = const_cast<hdk_csr_s *>(Apar); L.owns_A = (level > 0); L.n = A.n; L.row0 = A.row0;
            if ((rc = dalloc(&L.cf, (size_t)A.n + 1))) break;
         }
         std::vector<int64_t> cstarts;
         GRows P, Rt, C;
         bool  stop = false;
         rc = dist_level(prm, A, Apar, fstarts, level, keep_debug, M->lev.back().cf, cstarts, P, Rt, C, stop);
         if (rc) break;
         if (stop)
         {
            // coarsening produced nothing (or everything): this level is the coarsest one and becomes
            // the single level of the tail (same rule as the serial setup)
            AmgLevel &L = M->lev.back();
            dfree(L.cf); L.cf = nullptr;
            if (L.owns_A) destroy_local(L.A);
            M->lev.pop_back();
            tail_final = true;
            break;
         }
         nnz_sum += (double)Apar->global_nnz;
         AmgLevel &L = M->lev.back();
         const int64_t ncg = cstarts[(size_t)R], cs = cstarts[(size_t)me], ce = cstarts[(size_t)me + 1];
         bool next_empty = false;
         for (int r = 0; r < R; r++) if (cstarts[(size_t)r + 1] <= cstarts[(size_t)r]) next_empty = true;
         // hypre's stopping rule: the level just produced is the last one when it is small enough or
         // the level budget is used up; such a level (and every small one) is replicated
         const bool final_next = (ncg <= prm->max_coarse_size) || (level + 1 >= prm->max_levels - 1);
         const bool next_replicated = final_next || ncg <= rep_rows || next_empty;
         if (keep_debug) { if ((rc = keep_debug_copy(L, 0, A)) || (rc = keep_debug_copy(L, 1, P))) break; }
         if (next_replicated) rc = grows_to_parcsr(P, 0, ncg, nfg, ncg, false, false, &L.P); // reads the complete coarse vector
         else rc = grows_to_parcsr(P, cs, ce, nfg, ncg, false, true, &L.P);
         if (rc) break;
         if ((rc = grows_to_parcsr(Rt, fstarts[(size_t)me], fstarts[(size_t)me + 1], ncg, nfg, false, true, &L.R))) break;
         grows_free(P); grows_free(Rt);
         setup_stage_mark("parcsr P R", level);
         grows_free(A);
         A = C;
         fstarts = cstarts;
         level++;
         if (next_replicated) { tail_final = final_next; break; }
         hdk_csr_s *An = nullptr;
         if ((rc = grows_to_parcsr(A, cs, ce, ncg, ncg, true, true, &An))) break;
         Apar = An;
         setup_stage_mark("parcsr A", level);
      }
   }
   if (rc == HDK_OK)
   {
      // replicated tail: gather the level and continue with the serial setup on every rank
      hdk_csr_s *G = nullptr;
      rc = gather_level(A, fstarts, &G);
      if (rc == HDK_OK)
      {
         hdk_amg_params tp = *prm;
         tp.max_levels = tail_final ? 1 : prm->max_levels - level;
         if (tp.max_levels < 1) tp.max_levels = 1;
         hdk_amg_s *T = nullptr;
         rc = setup_serial(G, &tp, &T, false, -1);
         if (rc == HDK_OK)
         {
            T->lev[0].owns_A = true;
            M->tail = T; M->tail_level = 0;
            M->tail_n   = T->lev[0].n;
            M->tail_off = fstarts[(size_t)me];
            M->tail_cnt = fstarts[(size_t)me + 1] - fstarts[(size_t)me];
            for (int l = 0; l < T->nlev; l++) nnz_sum += (double)T->lev[(size_t)l].A->diag.nnz;
         }
         else destroy_local(G);
      }
      setup_stage_mark("tail", -2);
   }
   grows_free(A);
   if (rc == HDK_OK)
   {
      M->nlev = (int)M->lev.size();
      M->op_complexity = nnz_sum / (nnz0 > 0 ? nnz0 : 1.0);
      rc = finalize_levels(M, prm, -1);
      M->vcycle_bytes += M->tail ? M->tail->vcycle_bytes : 0.0;
   }
   if (rc == HDK_OK) rc = dalloc(&M->full_f, (size_t)M->tail_n + 8);
   if (rc == HDK_OK) rc = dalloc(&M->full_u, (size_t)M->tail_n + 8);
   if (rc == HDK_OK && g.nranks > 1 && M->nlev > 0) rc = ipc_gather_alloc(M->gather, M->tail_n);
   if (rc == HDK_OK && cudaStreamSynchronize(g.stream) != cudaSuccess)
      rc = set_error(HDK_ERR_CUDA, "distributed setup sync failed: %s", cudaGetErrorString(cudaGetLastError()));
   setup_stage_mark("finalize", -2);
   if (rc != HDK_OK) { hdk_amg_destroy(M); return rc; }
   *out = M;
   return HDK_OK;
}

// my rows of A_l (which = 0) or P_l (which = 1) of a distributed level with global columns in the
// serial storage order (kept when the tunable amg_keep_debug is set); sizes first, then the arrays
extern "C" int hdk_amg_get_rows(const hdk_amg *M, int level, int which, int64_t *row0, int64_t *nrows, int64_t *nnz,
                                int64_t *indptr_h, int64_t *cols_h, double *vals_h)
{
   HDK_TRY(require_init());
   if (!M || level < 0 || level >= M->nlev || which < 0 || which > 1) return set_error(HDK_ERR_INVALID, "level / matrix out of range");
   const AmgLevel &L = M->lev[(size_t)level];
   if (!L.dbg_ip[which]) return set_error(HDK_ERR_INVALID, "rows of level %d were not kept (set the tunable amg_keep_debug before the setup)", level);
   if (row0) *row0 = L.row0;
   if (nrows) *nrows = L.n;
   if (nnz) *nnz = L.dbg_nnz[which];
   if (indptr_h) HDK_CUDA(cudaMemcpyAsync(indptr_h, L.dbg_ip[which], sizeof(int64_t) * ((size_t)L.n + 1), cudaMemcpyDeviceToHost, g.stream));
   if (cols_h && L.dbg_nnz[which]) HDK_CUDA(cudaMemcpyAsync(cols_h, L.dbg_col[which], sizeof(int64_t) * (size_t)L.dbg_nnz[which], cudaMemcpyDeviceToHost, g.stream));
   if (vals_h && L.dbg_nnz[which]) HDK_CUDA(cudaMemcpyAsync(vals_h, L.dbg_val[which], sizeof(double) * (size_t)L.dbg_nnz[which], cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   return HDK_OK;
}
