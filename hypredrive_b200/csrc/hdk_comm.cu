// hdk_comm.cu -- one process per GPU: NCCL communicator, halo plans and exchanges.
// Replaces the MPI layer hypre uses underneath hypredrive (hypre_ParCSRCommPkg /
// hypre_ParCSRCommHandle persistent Isend/Irecv and MPI_Allreduce; reference plumbing:
// src/HYPREDRV.c:1014-1043, src/internal/runtime.c:118-133).  NCCL is bound at run time with
// dlopen so that a process that already carries torch's libnccl shares that copy.
#include "hdk_internal.cuh"
#include <dlfcn.h>
#include <algorithm>

namespace hdk {

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void *ncclComm_p;
enum { NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_MAX = 2 };

struct NcclApi
{
   void *h = nullptr;
   int (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
   int (*CommInitRank)(ncclComm_p *, int, ncclUniqueId_t, int) = nullptr;
   int (*CommDestroy)(ncclComm_p) = nullptr;
   int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
   int (*AllGather)(const void *, void *, size_t, int, ncclComm_p, cudaStream_t) = nullptr;
   int (*Send)(const void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
   int (*Recv)(void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
   int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
   int (*GroupStart)() = nullptr;
   int (*GroupEnd)() = nullptr;
   const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi nccl;

static int nccl_load()
{
   if (nccl.h) return HDK_OK;
   const char *cands[] = {getenv("HDK_NCCL_LIB"), "libnccl.so.2", "libnccl.so",
                          "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
   for (const char *c : cands)
   {
      if (!c) continue;
      nccl.h = dlopen(c, RTLD_NOW | RTLD_LOCAL);
      if (nccl.h) break;
   }
   if (!nccl.h) return set_error(HDK_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
#define LD(field, name)                                                           \
   *(void **)(&nccl.field) = dlsym(nccl.h, name);                                 \
   if (!nccl.field) return set_error(HDK_ERR_COMM, "NCCL symbol %s missing", name)
   LD(GetUniqueId, "ncclGetUniqueId"); LD(CommInitRank, "ncclCommInitRank");
   LD(CommDestroy, "ncclCommDestroy"); LD(AllReduce, "ncclAllReduce"); LD(AllGather, "ncclAllGather");
   LD(Broadcast, "ncclBroadcast"); LD(Send, "ncclSend"); LD(Recv, "ncclRecv"); LD(GroupStart, "ncclGroupStart");
   LD(GroupEnd, "ncclGroupEnd"); LD(GetErrorString, "ncclGetErrorString");
#undef LD
   return HDK_OK;
}

#define HDK_NCCL(call)                                                                         \
   do {                                                                                        \
      int r__ = (call);                                                                        \
      if (r__ != 0)                                                                            \
         return set_error(HDK_ERR_COMM, "%s:%d %s -> %s", __FILE__, __LINE__, #call,           \
                          nccl.GetErrorString ? nccl.GetErrorString(r__) : "nccl error");      \
   } while (0)

int allreduce_dev(double *buf_d, int count)
{
   if (g.nranks <= 1) return HDK_OK;
   HDK_NCCL(nccl.AllReduce(buf_d, buf_d, (size_t)count, NCCL_FLOAT64, NCCL_SUM, (ncclComm_p)g.nccl, g.stream));
   return HDK_OK;
}

int bcast_bytes(void *buf_d, size_t bytes, int root)
{
   if (g.nranks <= 1 || bytes == 0) return HDK_OK;
   HDK_NCCL(nccl.Broadcast(buf_d, buf_d, bytes, 0 /* ncclInt8 */, root, (ncclComm_p)g.nccl, g.stream));
   return HDK_OK;
}

// in-place all-gather of variable-size segments: rank r owns bytes [offs[r], offs[r+1]) of base
int allgatherv_bytes(void *base_d, const int64_t *offs)
{
   if (g.nranks <= 1) return HDK_OK;
   char *b = static_cast<char *>(base_d);
   HDK_NCCL(nccl.GroupStart());
   for (int r = 0; r < g.nranks; r++)
   {
      size_t bytes = (size_t)(offs[r + 1] - offs[r]);
      if (bytes == 0) continue;
      HDK_NCCL(nccl.Broadcast(b + offs[r], b + offs[r], bytes, 0 /* ncclInt8 */, r, (ncclComm_p)g.nccl, g.stream));
   }
   HDK_NCCL(nccl.GroupEnd());
   return HDK_OK;
}

int allreduce_max_dev(double *buf_d, int count)
{
   if (g.nranks <= 1) return HDK_OK;
   HDK_NCCL(nccl.AllReduce(buf_d, buf_d, (size_t)count, NCCL_FLOAT64, NCCL_MAX, (ncclComm_p)g.nccl, g.stream));
   return HDK_OK;
}

// gather one int64 from every rank onto the host
int allgather_i64_host(int64_t mine, std::vector<int64_t> &all)
{
   all.assign((size_t)g.nranks, mine);
   if (g.nranks <= 1) return HDK_OK;
   int64_t *d;
   HDK_TRY(dalloc(&d, (size_t)g.nranks + 1));
   HDK_CUDA(cudaMemcpyAsync(d + g.nranks, &mine, sizeof(int64_t), cudaMemcpyHostToDevice, g.stream));
   HDK_NCCL(nccl.AllGather(d + g.nranks, d, 1, NCCL_INT64, (ncclComm_p)g.nccl, g.stream));
   HDK_CUDA(cudaMemcpyAsync(all.data(), d, sizeof(int64_t) * (size_t)g.nranks, cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   dfree(d);
   return HDK_OK;
}

// gather `cnt` ints from every rank onto the host (all[r*cnt + i])
int allgather_i32_host(const int *mine, int cnt, std::vector<int> &all)
{
   all.assign((size_t)g.nranks * cnt, 0);
   if (g.nranks <= 1) { std::copy(mine, mine + cnt, all.begin()); return HDK_OK; }
   int *d;
   HDK_TRY(dalloc(&d, (size_t)(g.nranks + 1) * cnt));
   HDK_CUDA(cudaMemcpyAsync(d + (size_t)g.nranks * cnt, mine, sizeof(int) * (size_t)cnt, cudaMemcpyHostToDevice, g.stream));
   HDK_NCCL(nccl.AllGather(d + (size_t)g.nranks * cnt, d, (size_t)cnt, NCCL_INT32, (ncclComm_p)g.nccl, g.stream));
   HDK_CUDA(cudaMemcpyAsync(all.data(), d, sizeof(int) * (size_t)g.nranks * cnt, cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   dfree(d);
   return HDK_OK;
}

__global__ void k_ids_to_local(const int64_t *ids, int n, int64_t row_start, int *idx)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) idx[i] = (int)(ids[i] - row_start);
}

// Build the comm package of one ParCSR matrix: who owns my halo columns, who needs my rows.
// `uniq` (device, sorted unique global ids, n_halo) is adopted as halo.col_map.
int build_halo_plan(hdk_csr_s &A, int64_t *uniq, int n_halo)
{
   HaloPlan &H = A.halo;
   H.n_halo  = n_halo;
   H.col_map = uniq;
   if (g.nranks <= 1)
      return set_error(HDK_ERR_INVALID, "matrix has %d off-rank columns but the communicator has one rank", n_halo);
   std::vector<int64_t> starts;
   HDK_TRY(allgather_i64_host(A.col_start, starts)); // partition of the COLUMN space
   starts.push_back(A.global_cols);
   A.row_starts = starts;
   std::vector<int64_t> ids((size_t)n_halo);
   if (n_halo > 0)
   {
      HDK_CUDA(cudaMemcpyAsync(ids.data(), uniq, sizeof(int64_t) * (size_t)n_halo, cudaMemcpyDeviceToHost, g.stream));
      HDK_CUDA(cudaStreamSynchronize(g.stream));
   }
   std::vector<int> want((size_t)g.nranks, 0);
   for (int i = 0; i < n_halo; i++)
   {
      int r = (int)(std::upper_bound(starts.begin(), starts.end(), ids[(size_t)i]) - starts.begin()) - 1;
      if (r < 0 || r >= g.nranks || r == g.rank)
         return set_error(HDK_ERR_INVALID, "column %lld outside the global partition", (long long)ids[(size_t)i]);
      want[(size_t)r]++;
   }
   std::vector<int> all;
   HDK_TRY(allgather_i32_host(want.data(), g.nranks, all));
   int off = 0;
   for (int r = 0; r < g.nranks; r++)
      if (want[(size_t)r] > 0) { H.recv_rank.push_back(r); H.recv_off.push_back(off); H.recv_cnt.push_back(want[(size_t)r]); off += want[(size_t)r]; }
   int soff = 0;
   for (int r = 0; r < g.nranks; r++)
   {
      int c = all[(size_t)r * g.nranks + g.rank];
      if (c > 0) { H.send_rank.push_back(r); H.send_off.push_back(soff); H.send_cnt.push_back(c); soff += c; }
   }
   H.n_send = soff;
   int64_t *req;
   HDK_TRY(dalloc(&req, (size_t)soff + 1));
   HDK_TRY(dalloc(&H.send_idx, (size_t)soff + 1));
   HDK_TRY(dalloc(&H.send_buf, (size_t)soff + 1));
   HDK_TRY(dalloc(&H.x_halo, (size_t)n_halo + 8));
   HDK_NCCL(nccl.GroupStart());
   for (size_t i = 0; i < H.recv_rank.size(); i++)
      HDK_NCCL(nccl.Send(uniq + H.recv_off[i], (size_t)H.recv_cnt[i], NCCL_INT64, H.recv_rank[i], (ncclComm_p)g.nccl, g.stream));
   for (size_t i = 0; i < H.send_rank.size(); i++)
      HDK_NCCL(nccl.Recv(req + H.send_off[i], (size_t)H.send_cnt[i], NCCL_INT64, H.send_rank[i], (ncclComm_p)g.nccl, g.stream));
   HDK_NCCL(nccl.GroupEnd());
   if (soff > 0)
   {
      k_ids_to_local<<<cdiv(soff, 256), 256, 0, g.stream>>>(req, soff, A.col_start, H.send_idx);
      HDK_LAUNCH_CHECK();
   }
   dfree(req);
   return HDK_OK;
}

__global__ void k_pack(const double *x, const int *idx, int n, double *buf)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) buf[i] = x[idx[i]];
}

// pack on the compute stream, exchange on the comm stream (overlaps the diag-block SpMV)
int halo_exchange_begin(const hdk_csr_s &A, const double *x)
{
   const HaloPlan &H = A.halo;
   if (g.nranks <= 1) return HDK_OK;
   if (H.n_send > 0)
   {
      k_pack<<<cdiv(H.n_send, 256), 256, 0, g.stream>>>(x, H.send_idx, H.n_send, H.send_buf);
      HDK_LAUNCH_CHECK();
   }
   HDK_CUDA(cudaEventRecord(g.ev_pack, g.stream));
   HDK_CUDA(cudaStreamWaitEvent(g.comm_stream, g.ev_pack, 0));
   HDK_NCCL(nccl.GroupStart());
   for (size_t i = 0; i < H.send_rank.size(); i++)
      HDK_NCCL(nccl.Send(H.send_buf + H.send_off[i], (size_t)H.send_cnt[i], NCCL_FLOAT64, H.send_rank[i], (ncclComm_p)g.nccl, g.comm_stream));
   for (size_t i = 0; i < H.recv_rank.size(); i++)
      HDK_NCCL(nccl.Recv(H.x_halo + H.recv_off[i], (size_t)H.recv_cnt[i], NCCL_FLOAT64, H.recv_rank[i], (ncclComm_p)g.nccl, g.comm_stream));
   HDK_NCCL(nccl.GroupEnd());
   HDK_CUDA(cudaEventRecord(g.ev_halo, g.comm_stream));
   return HDK_OK;
}

int halo_exchange_end(const hdk_csr_s &A)
{
   if (g.nranks <= 1) return HDK_OK;
   HDK_CUDA(cudaStreamWaitEvent(g.stream, g.ev_halo, 0));
   return HDK_OK;
}

} // namespace hdk

using namespace hdk;

extern "C" {

int hdk_comm_unique_id(void *id128_h)
{
   HDK_TRY(nccl_load());
   ncclUniqueId_t id;
   HDK_NCCL(nccl.GetUniqueId(&id));
   memcpy(id128_h, &id, 128);
   return HDK_OK;
}

int hdk_comm_init(int rank, int nranks, const void *id128_h)
{
   HDK_TRY(require_init());
   if (nranks <= 1) { g.rank = 0; g.nranks = 1; return HDK_OK; }
   if (g.nccl) return HDK_OK;
   HDK_TRY(nccl_load());
   ncclUniqueId_t id;
   memcpy(&id, id128_h, 128);
   ncclComm_p comm = nullptr;
   HDK_NCCL(nccl.CommInitRank(&comm, nranks, id, rank));
   g.nccl = comm; g.rank = rank; g.nranks = nranks;
   return HDK_OK;
}

int hdk_comm_max_i64(int64_t local, int64_t *global)
{
   *global = local;
   if (g.nranks <= 1) return HDK_OK;
   std::vector<int64_t> all;
   HDK_TRY(allgather_i64_host(local, all));
   for (int64_t v : all) if (v > *global) *global = v;
   return HDK_OK;
}

int hdk_comm_sum_i64(int64_t local, int64_t *global)
{
   *global = local;
   if (g.nranks <= 1) return HDK_OK;
   std::vector<int64_t> all;
   HDK_TRY(allgather_i64_host(local, all));
   *global = 0;
   for (int64_t v : all) *global += v;
   return HDK_OK;
}

int hdk_comm_rank(void) { return g.rank; }
int hdk_comm_size(void) { return g.nranks; }

int hdk_comm_finalize(void)
{
   if (g.nccl && nccl.CommDestroy) nccl.CommDestroy((ncclComm_p)g.nccl);
   g.nccl = nullptr; g.rank = 0; g.nranks = 1;
   return HDK_OK;
}

} // extern "C"
