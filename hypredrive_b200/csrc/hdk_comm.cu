// hdk_comm.cu -- one process per GPU: NCCL communicator, halo plans and exchanges.
// Replaces the MPI layer hypre uses underneath hypredrive (hypre_ParCSRCommPkg /
// hypre_ParCSRCommHandle persistent Isend/Irecv and MPI_Allreduce; reference plumbing:
// src/HYPREDRV.c:1014-1043, src/internal/runtime.c:118-133).  NCCL is bound at run time with
// dlopen so that a process that already carries torch's libnccl shares that copy.
#include "hdk_internal.cuh"
#include <sys/stat.h>
#include <time.h>
#include <dlfcn.h>
#include <time.h>
#include <algorithm>
#include <map>

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

// ------------------------------------------------------------------------------------------
// Peer-memory arena: one cudaMalloc'ed block per rank, opened by every other rank through CUDA
// IPC at communicator creation.  Halo plans carve their receive buffers and flag slots out of
// it, so a neighbour can store into them directly over NVLink.
// ------------------------------------------------------------------------------------------
// Reduction mailbox (the first MAIL_BYTES of every arena): slot[parity][source rank][MB_VALS] doubles
// followed by flag[parity][source rank] sequence numbers.  An all-reduce of a few scalars is ONE
// small kernel: every rank stores its partial into every peer's mailbox over NVLink, raises the
// flag, waits for the peers' flags in its own mailbox and adds the R partials in rank order --
// the same order on every rank, so all ranks hold bit-identical sums (and take identical
// decisions), and the scalar recurrence of the Krylov driver (apply_fin) runs in the same kernel.
constexpr int    MB_MAXR = 16, MB_VALS = 4;
constexpr size_t MAIL_SLOT_BYTES = (size_t)2 * MB_MAXR * MB_VALS * sizeof(double);
constexpr size_t MAIL_BYTES = 8192; // slots (1 KB) + flags (256 B), padded
struct MailArgs
{
   double             *peer_slot[MB_MAXR];
   unsigned long long *peer_flag[MB_MAXR];
   double             *my_slot;
   unsigned long long *my_flag;
   int                 R, me;
   unsigned long long  seq;
   long long           tmo;
   int                *err;
};

struct IpcState
{
   bool                on = false;
   bool                mail_on = false;
   unsigned long long  mail_seq = 0;
   long long           tmo = 0;              // wait budget of the flag waits (clock64 ticks)
   int                *err_d = nullptr;      // device word raised by a kernel whose wait ran out (polled from
                                             // L2 by the waiting threads, read by the host at the end of a solve)
   char               *base = nullptr;
   size_t              size = 0;
   std::vector<char *> peer;                 // peer[r] = rank r's arena in my address space
   std::map<size_t, size_t>               free_; // offset -> bytes
   std::vector<std::pair<size_t, size_t>> pending; // released, reusable after the next collective
};
static IpcState ipc;

static void arena_release_now(size_t off, size_t bytes)
{
   auto it = ipc.free_.emplace(off, bytes).first;
   auto nx = std::next(it);
   if (nx != ipc.free_.end() && it->first + it->second == nx->first) { it->second += nx->second; ipc.free_.erase(nx); }
   if (it != ipc.free_.begin())
   {
      auto pv = std::prev(it);
      if (pv->first + pv->second == it->first) { pv->second += it->second; ipc.free_.erase(it); }
   }
}
static void arena_collect()
{
   for (auto &p : ipc.pending) arena_release_now(p.first, p.second);
   ipc.pending.clear();
}
static int64_t arena_alloc(size_t bytes)
{
   bytes = (bytes + 255) & ~(size_t)255;
   for (auto it = ipc.free_.begin(); it != ipc.free_.end(); ++it)
      if (it->second >= bytes)
      {
         size_t off = it->first, rest = it->second - bytes;
         ipc.free_.erase(it);
         if (rest) ipc.free_.emplace(off + bytes, rest);
         return (int64_t)off;
      }
   return -1;
}

static int ipc_setup()
{
   const char *e = getenv("HDK_HALO_IPC");
   if (e && atoi(e) == 0) return HDK_OK;
   size_t mb = 256;
   if (getenv("HDK_IPC_ARENA_MB")) mb = (size_t)atoll(getenv("HDK_IPC_ARENA_MB"));
   int ok = 1;
   cudaIpcMemHandle_t mine;
   memset(&mine, 0, sizeof(mine));
   if (cudaMalloc((void **)&ipc.base, mb << 20) != cudaSuccess) { ok = 0; ipc.base = nullptr; cudaGetLastError(); }
   // zeroed on the stream the flags are used on (the allgather below orders it before any peer's store)
   if (ok && cudaMemsetAsync(ipc.base, 0, mb << 20, g.stream) != cudaSuccess) ok = 0;
   if (ok && !ipc.err_d)
   {
      if (cudaMalloc((void **)&ipc.err_d, sizeof(int)) != cudaSuccess) { ok = 0; ipc.err_d = nullptr; cudaGetLastError(); }
      else if (cudaMemsetAsync(ipc.err_d, 0, sizeof(int), g.stream) != cudaSuccess) ok = 0;
      double secs = 300.0;
      if (getenv("HDK_IPC_TIMEOUT_S")) secs = atof(getenv("HDK_IPC_TIMEOUT_S"));
      int khz = 0;
      cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, g.device);
      ipc.tmo = secs > 0.0 ? (long long)(secs * 1000.0 * (double)(khz > 0 ? khz : 1900000)) : 0;
      if (ok && (wait_globals_set(ipc.tmo, ipc.err_d) != cudaSuccess || wait_globals_spmv(ipc.tmo, ipc.err_d) ||
                 wait_globals_csr(ipc.tmo, ipc.err_d) || wait_globals_vec(ipc.tmo, ipc.err_d))) ok = 0;
   }
   if (ok && cudaIpcGetMemHandle(&mine, ipc.base) != cudaSuccess) { ok = 0; cudaGetLastError(); }
   // exchange the handles (and whether every rank got this far)
   const size_t   hb = sizeof(cudaIpcMemHandle_t) + 8;
   char          *d;
   std::vector<char> host((size_t)g.nranks * hb, 0), my(hb, 0);
   memcpy(my.data(), &mine, sizeof(mine));
   my[sizeof(mine)] = (char)ok;
   HDK_TRY(dalloc(&d, (size_t)(g.nranks + 1) * hb));
   HDK_CUDA(cudaMemcpyAsync(d + (size_t)g.nranks * hb, my.data(), hb, cudaMemcpyHostToDevice, g.stream));
   HDK_NCCL(nccl.AllGather(d + (size_t)g.nranks * hb, d, hb, 0 /* ncclInt8 */, (ncclComm_p)g.nccl, g.stream));
   HDK_CUDA(cudaMemcpyAsync(host.data(), d, (size_t)g.nranks * hb, cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   dfree(d);
   for (int r = 0; r < g.nranks; r++) if (!host[(size_t)r * hb + sizeof(mine)]) ok = 0;
   ipc.peer.assign((size_t)g.nranks, nullptr);
   if (ok)
   {
      for (int r = 0; r < g.nranks && ok; r++)
      {
         if (r == g.rank) { ipc.peer[(size_t)r] = ipc.base; continue; }
         cudaIpcMemHandle_t h;
         memcpy(&h, host.data() + (size_t)r * hb, sizeof(h));
         void *p = nullptr;
         if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
         ipc.peer[(size_t)r] = static_cast<char *>(p);
      }
   }
   // every rank must agree
   std::vector<int64_t> oks;
   HDK_TRY(allgather_i64_host(ok, oks));
   for (int64_t v : oks) if (!v) ok = 0;
   if (!ok)
   {
      for (int r = 0; r < g.nranks; r++)
         if (r != g.rank && ipc.peer[(size_t)r]) cudaIpcCloseMemHandle(ipc.peer[(size_t)r]);
      if (ipc.base) cudaFree(ipc.base);
      if (ipc.err_d) cudaFree(ipc.err_d);
      ipc = IpcState();
      cudaGetLastError();
      return HDK_OK; // NCCL send/recv path stays in use
   }
   ipc.size = mb << 20;
   ipc.free_.clear();
   ipc.free_.emplace(MAIL_BYTES, ipc.size - MAIL_BYTES); // the first bytes of every arena are its reduction mailbox
   ipc.on = true;
   ipc.mail_seq = 0;
   ipc.mail_on = g.nranks <= MB_MAXR && !(getenv("HDK_MAILBOX") && atoi(getenv("HDK_MAILBOX")) == 0);
   return HDK_OK;
}

static void ipc_teardown()
{
   if (!ipc.base) return;
   cudaDeviceSynchronize();
   for (int r = 0; r < (int)ipc.peer.size(); r++)
      if (r != g.rank && ipc.peer[(size_t)r]) cudaIpcCloseMemHandle(ipc.peer[(size_t)r]);
   cudaFree(ipc.base);
   if (ipc.err_d) cudaFree(ipc.err_d);
   ipc = IpcState();
}

int comm_check_error()
{
   if (!ipc.err_d) return HDK_OK;
   int flag = 0;
   HDK_CUDA(cudaMemcpyAsync(&flag, ipc.err_d, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   if (!flag) return HDK_OK;
   HDK_CUDA(cudaMemsetAsync(ipc.err_d, 0, sizeof(int), g.stream));
   ipc.on = false; // plans built from now on use NCCL send/recv
   return set_error(HDK_ERR_COMM, "a peer-memory halo wait ran out of its budget (HDK_IPC_TIMEOUT_S): a neighbour rank is late or lost; "
                                  "results of this operation are invalid.  Use HDK_HALO_IPC=0 under profilers and sanitizers");
}

__global__ void k_apply_fin_comm(int fin, const double *v, double *out, double *scal) { apply_fin(fin, v[0], out, scal); }

__global__ void __launch_bounds__(32) k_mailbox_allreduce(double *vals, int count, MailArgs a, int fin, double *fin_out, double *scal)
{
   const int lane = threadIdx.x, par = (int)(a.seq & 1ull);
   if (lane < a.R)
   {
      double *dst = a.peer_slot[lane] + ((size_t)par * MB_MAXR + a.me) * MB_VALS;
      for (int c = 0; c < count; c++) dst[c] = vals[c];
      __threadfence_system();
      st_release_sys_u64(a.peer_flag[lane] + par * MB_MAXR + a.me, a.seq);
      wait_seq_sys(a.my_flag + par * MB_MAXR + lane, a.seq, a.tmo, a.err);
   }
   __syncwarp();
   if (lane == 0)
   {
      for (int c = 0; c < count; c++)
      {
         double s = 0.0;
         for (int p = 0; p < a.R; p++) s += __ldcg(a.my_slot + ((size_t)par * MB_MAXR + p) * MB_VALS + c); // rank order
         if (fin != FIN_NONE && c == 0) apply_fin(fin, s, fin_out, scal);
         else vals[c] = s;
      }
   }
}

static MailArgs mail_args()
{
   MailArgs a;
   memset(&a, 0, sizeof(a));
   a.R = g.nranks; a.me = g.rank; a.seq = ++ipc.mail_seq; a.tmo = ipc.tmo; a.err = ipc.err_d;
   for (int p = 0; p < g.nranks; p++)
   {
      a.peer_slot[p] = reinterpret_cast<double *>(ipc.peer[(size_t)p]);
      a.peer_flag[p] = reinterpret_cast<unsigned long long *>(ipc.peer[(size_t)p] + MAIL_SLOT_BYTES);
   }
   a.my_slot = a.peer_slot[g.rank]; a.my_flag = a.peer_flag[g.rank];
   return a;
}

// sum over ranks of buf_d[0..count) in place; with fin != FIN_NONE the sum of buf_d[0] is handed to
// the Krylov scalar recurrence (apply_fin) instead of being stored
int allreduce_fin_dev(double *buf_d, int count, int fin, double *fin_out)
{
   if (g.nranks <= 1)
   {
      if (fin != FIN_NONE) { k_apply_fin_comm<<<1, 1, 0, g.stream>>>(fin, buf_d, fin_out, g.dscal); HDK_LAUNCH_CHECK(); }
      return HDK_OK;
   }
   if (ipc.on && ipc.mail_on && count <= MB_VALS)
   {
      k_mailbox_allreduce<<<1, 32, 0, g.stream>>>(buf_d, count, mail_args(), fin, fin_out, g.dscal);
      HDK_LAUNCH_CHECK();
      return HDK_OK;
   }
   HDK_NCCL(nccl.AllReduce(buf_d, buf_d, (size_t)count, NCCL_FLOAT64, NCCL_SUM, (ncclComm_p)g.nccl, g.stream));
   if (fin != FIN_NONE) { k_apply_fin_comm<<<1, 1, 0, g.stream>>>(fin, buf_d, fin_out, g.dscal); HDK_LAUNCH_CHECK(); }
   return HDK_OK;
}

int allreduce_dev(double *buf_d, int count)
{
   if (g.nranks <= 1) return HDK_OK;
   return allreduce_fin_dev(buf_d, count, FIN_NONE, nullptr);
}

// ---- all-gather of the ranks' slices of a replicated vector through peer stores ---------------
// (the restricted right-hand side of the first replicated level: every rank owns a disjoint slice,
// so the sum over ranks is an all-gather).  buf lives in the arena, two halves by sequence parity.
struct GatherArgs
{
   double             *peer_buf[MB_MAXR];   // peer's vector (current half)
   unsigned long long *peer_flag[MB_MAXR];  // peer's flag slot for me
   const unsigned long long *my_flag;       // my flag slots, one per source rank
   int                 R, me;
   unsigned long long  seq;
   unsigned           *ticket;
   long long           tmo;
   int                *err;
};
__global__ void k_ipc_gather_send(const double *mine, int64_t off, int64_t cnt, GatherArgs a)
{
   const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
   if (i < cnt)
   {
      const double v = mine[off + i];
      for (int p = 0; p < a.R; p++) if (p != a.me) a.peer_buf[p][off + i] = v;
   }
   __syncthreads();
   if (threadIdx.x == 0)
   {
      __threadfence_system();
      unsigned t = atomicInc(a.ticket, gridDim.x - 1);
      if (t == gridDim.x - 1)
      {
         __threadfence_system();
         for (int p = 0; p < a.R; p++) if (p != a.me) st_release_sys_u64(a.peer_flag[p], a.seq);
      }
   }
}
__global__ void __launch_bounds__(32) k_ipc_gather_wait(GatherArgs a)
{
   const int lane = threadIdx.x;
   if (lane < a.R && lane != a.me) wait_seq_sys(a.my_flag + lane, a.seq, a.tmo, a.err);
}

int ipc_gather_alloc(IpcGather &G, int64_t n)
{
   G = IpcGather();
   if (!ipc.on || !ipc.mail_on || g.nranks > MB_MAXR) return HDK_OK;
   arena_collect();
   const size_t half = (((size_t)n + 15) & ~(size_t)15) * sizeof(double);
   const size_t bytes = 256 + 2 * half;
   int64_t off = arena_alloc(bytes);
   if (off >= 0) HDK_CUDA(cudaMemsetAsync(ipc.base + off, 0, 256, g.stream));
   std::vector<int64_t> offs;
   HDK_TRY(allgather_i64_host(off, offs)); // also orders the memset before any peer's first store
   bool all_ok = true;
   for (int64_t v : offs) if (v < 0) all_ok = false;
   if (!all_ok)
   {
      if (off >= 0) arena_release_now((size_t)off, (bytes + 255) & ~(size_t)255);
      return HDK_OK;
   }
   G.on = true; G.n = n; G.region_off = off; G.region_bytes = (bytes + 255) & ~(size_t)255; G.seq = 0;
   G.flags = reinterpret_cast<unsigned long long *>(ipc.base + off);
   G.buf[0] = reinterpret_cast<double *>(ipc.base + off + 256);
   G.buf[1] = reinterpret_cast<double *>(ipc.base + off + 256 + half);
   for (int p = 0; p < g.nranks; p++)
   {
      char *rb = ipc.peer[(size_t)p] + offs[(size_t)p];
      G.peer_buf[0][p] = reinterpret_cast<double *>(rb + 256);
      G.peer_buf[1][p] = reinterpret_cast<double *>(rb + 256 + half);
      G.peer_flag[p]   = reinterpret_cast<unsigned long long *>(rb) + g.rank;
   }
   HDK_TRY(dalloc(&G.ticket, 1));
   HDK_CUDA(cudaMemsetAsync(G.ticket, 0, sizeof(unsigned), g.stream));
   return HDK_OK;
}
void ipc_gather_free(IpcGather &G)
{
   if (!G.on) return;
   dfree(G.ticket);
   if (ipc.on) ipc.pending.emplace_back((size_t)G.region_off, G.region_bytes);
   G = IpcGather();
}
// the half the next gather fills (the owner writes its own slice there before calling ipc_gather)
double *ipc_gather_buffer(IpcGather &G) { return G.buf[(G.seq + 1) & 1]; }
// my slice [off, off + cnt) of the current buffer goes to every peer; returns when all slices are in
int ipc_gather(IpcGather &G, int64_t off, int64_t cnt)
{
   G.seq++;
   GatherArgs a;
   memset(&a, 0, sizeof(a));
   a.R = g.nranks; a.me = g.rank; a.seq = G.seq; a.ticket = G.ticket; a.tmo = ipc.tmo; a.err = ipc.err_d;
   a.my_flag = G.flags;
   for (int p = 0; p < g.nranks; p++) { a.peer_buf[p] = G.peer_buf[G.seq & 1][p]; a.peer_flag[p] = G.peer_flag[p]; }
   const int grid = cnt > 0 ? cdiv(cnt, 256) : 1;
   k_ipc_gather_send<<<grid, 256, 0, g.stream>>>(G.buf[G.seq & 1], off, cnt, a);
   HDK_LAUNCH_CHECK();
   k_ipc_gather_wait<<<1, 32, 0, g.stream>>>(a);
   HDK_LAUNCH_CHECK();
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

int allreduce_i32_dev(int *buf_d, int count)
{
   if (g.nranks <= 1) return HDK_OK;
   HDK_NCCL(nccl.AllReduce(buf_d, buf_d, (size_t)count, NCCL_INT32, NCCL_SUM, (ncclComm_p)g.nccl, g.stream));
   return HDK_OK;
}

// personalised all-to-all of byte segments on the compute stream: segment [soff[r], soff[r+1]) of
// `send` goes to rank r and lands in [roff[q], roff[q+1]) of `recv` on rank q's side for sender q.
// Both sides know every size (count matrices are exchanged first), empty pairs are skipped, the
// segment to oneself is a device copy.
int alltoallv_bytes(const void *send_d, const int64_t *soff, void *recv_d, const int64_t *roff)
{
   const char *sb = static_cast<const char *>(send_d);
   char       *rb = static_cast<char *>(recv_d);
   const int   me = g.rank;
   if (soff[me + 1] - soff[me] != roff[me + 1] - roff[me])
      return set_error(HDK_ERR_COMM, "alltoallv: self segment sizes differ");
   if (soff[me + 1] > soff[me])
      HDK_CUDA(cudaMemcpyAsync(rb + roff[me], sb + soff[me], (size_t)(soff[me + 1] - soff[me]), cudaMemcpyDeviceToDevice, g.stream));
   if (g.nranks <= 1) return HDK_OK;
   HDK_NCCL(nccl.GroupStart());
   for (int r = 0; r < g.nranks; r++)
   {
      if (r == me) continue;
      size_t sbytes = (size_t)(soff[r + 1] - soff[r]), rbytes = (size_t)(roff[r + 1] - roff[r]);
      if (sbytes) HDK_NCCL(nccl.Send(sb + soff[r], sbytes, 0 /* ncclInt8 */, r, (ncclComm_p)g.nccl, g.stream));
      if (rbytes) HDK_NCCL(nccl.Recv(rb + roff[r], rbytes, 0 /* ncclInt8 */, r, (ncclComm_p)g.nccl, g.stream));
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

// ---- peer-memory halo plans -------------------------------------------------------------
// Gather the boundary values and store them into the neighbours' halo buffers (NVLink peer
// stores); the last block to finish raises this rank's sequence flag at every neighbour and waits
// for theirs, so the exchange is complete when the kernel ends (a rank that sends nothing runs one
// block that only does the flags).  Why no acknowledgement is needed before a buffer half is
// reused: see export_row (hdk_internal.cuh).
__global__ void k_pack_ipc(const double *x, const int *idx, int n, IpcSendArgs a)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
   {
      int p = 0;
      while (p + 1 < a.npeer && i >= a.off[p + 1]) p++;
      a.dst[p][i - a.off[p]] = x[idx[i]];
   }
   __syncthreads(); // then ONE system fence per CTA (thread 0) orders all its peer stores before the flag
   if (threadIdx.x == 0)
   {
      __threadfence_system();
      unsigned t = atomicInc(a.ticket, gridDim.x - 1);
      if (t == gridDim.x - 1) exchange_complete(a.flag, a.rflag, a.nflag, a.seq);
   }
}

// one region per plan in my arena: [16 sequence flags | x_halo half 0 | x_halo half 1]
static size_t ipc_half_doubles(int n_halo) { return ((size_t)n_halo + 15) & ~(size_t)15; }

// neighbours of rank r in the exchange: every rank it sends to or receives from (a symmetric relation)
static void ipc_neighbours(const std::vector<int> &all, int R, int r, std::vector<int> &out)
{
   out.clear();
   for (int q = 0; q < R; q++)
      if (q != r && (all[(size_t)r * R + q] > 0 || all[(size_t)q * R + r] > 0)) out.push_back(q);
}

static int ipc_plan_build(HaloPlan &H, const std::vector<int> &all /* all[r*R+q]: columns r wants from q */)
{
   IpcHalo &I = H.ipc;
   I = IpcHalo();
   if (!ipc.on) return HDK_OK;
   arena_collect(); // the host-synchronised collectives earlier in the plan build make released regions safe
   const int R = g.nranks, me = g.rank;
   std::vector<int> nbr, theirs;
   ipc_neighbours(all, R, me, nbr);
   int ok = nbr.size() <= (size_t)IPC_MAXP ? 1 : 0;
   const size_t bytes = 128 + 2 * ipc_half_doubles(H.n_halo) * sizeof(double);
   int64_t off = ok ? arena_alloc(bytes) : -1;
   if (off >= 0) HDK_CUDA(cudaMemsetAsync(ipc.base + off, 0, 128, g.stream)); // flags start at sequence 0
   std::vector<int64_t> offs;
   HDK_TRY(allgather_i64_host(off, offs)); // also orders my memset before any neighbour's first store
   bool all_ok = true;
   for (int64_t v : offs) if (v < 0) all_ok = false;
   if (!all_ok)
   {
      if (off >= 0) arena_release_now((size_t)off, (bytes + 255) & ~(size_t)255);
      return HDK_OK; // this plan stays on NCCL send/recv (same decision on every rank)
   }
   I.region_off = off; I.region_bytes = (bytes + 255) & ~(size_t)255;
   I.data_flag = reinterpret_cast<unsigned long long *>(ipc.base + off);
   I.xh[0]     = reinterpret_cast<double *>(ipc.base + off + 128);
   I.xh[1]     = I.xh[0] + ipc_half_doubles(H.n_halo);
   // where my values start in every send neighbour's halo, from the global want matrix
   for (size_t i = 0; i < H.send_rank.size(); i++)
   {
      const int D = H.send_rank[i];
      int       roff = 0, nh = 0;
      for (int q = 0; q < R; q++)
      {
         int c = all[(size_t)D * R + q];
         if (q < me) roff += c;
         nh += c;
      }
      char *rb = ipc.peer[(size_t)D] + offs[(size_t)D];
      I.dst[0][i] = reinterpret_cast<double *>(rb + 128) + roff;
      I.dst[1][i] = I.dst[0][i] + ipc_half_doubles(nh);
   }
   // my flag slot at every neighbour: my position in ITS neighbour list
   I.nnbr = (int)nbr.size();
   for (size_t i = 0; i < nbr.size(); i++)
   {
      const int D = nbr[i];
      ipc_neighbours(all, R, D, theirs);
      const int pos = (int)(std::find(theirs.begin(), theirs.end(), me) - theirs.begin());
      I.nbr_flag[i] = reinterpret_cast<unsigned long long *>(ipc.peer[(size_t)D] + offs[(size_t)D]) + pos;
   }
   HDK_TRY(dalloc(&I.tickets, 2));
   HDK_CUDA(cudaMemsetAsync(I.tickets, 0, 2 * sizeof(unsigned), g.stream));
   I.seq = 0;
   I.on  = true;
   return HDK_OK;
}

void halo_plan_free(HaloPlan &H)
{
   dfree(H.col_map); dfree(H.send_idx); dfree(H.send_buf); dfree(H.x_halo);
   H.col_map = nullptr; H.send_idx = nullptr; H.send_buf = nullptr; H.x_halo = nullptr;
   if (H.ipc.on)
   {
      dfree(H.ipc.tickets); dfree(H.ipc.exp_dir); dfree(H.ipc.exp_ptr); dfree(H.ipc.exp_slot);
      if (ipc.on) ipc.pending.emplace_back((size_t)H.ipc.region_off, H.ipc.region_bytes);
      H.ipc = IpcHalo();
   }
}

// where the consumer of the halo finds it (the current buffer half)
const double *halo_buffer(const hdk_csr_s &A)
{
   const HaloPlan &H = A.halo;
   return H.ipc.on ? H.ipc.xh[H.ipc.seq & 1] : H.x_halo;
}

// inverse of the send list (row -> its positions in the concatenated send buffer) for exports folded
// into the kernel that produces the vector; built once per plan on the host
static int ipc_export_build(HaloPlan &H, int A_rows /* rows of the vector that is exported */)
{
   IpcHalo &I = H.ipc;
   if (!I.on || H.n_send <= 0) return HDK_OK;
   std::vector<int> idx((size_t)H.n_send);
   HDK_CUDA(cudaMemcpyAsync(idx.data(), H.send_idx, sizeof(int) * (size_t)H.n_send, cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   std::vector<std::pair<int, int>> rs((size_t)H.n_send); // (row, slot)
   for (int s = 0; s < H.n_send; s++) rs[(size_t)s] = std::make_pair(idx[(size_t)s], s);
   std::sort(rs.begin(), rs.end());
   std::vector<int> rows, ptr, slot((size_t)H.n_send);
   for (int k = 0; k < H.n_send; k++)
   {
      if (k == 0 || rs[(size_t)k].first != rs[(size_t)k - 1].first) { rows.push_back(rs[(size_t)k].first); ptr.push_back(k); }
      slot[(size_t)k] = rs[(size_t)k].second;
   }
   ptr.push_back(H.n_send);
   const int m = (int)rows.size();
   // widest run of rows without exports (for slabs: everything between the two boundary planes)
   int lo_end = 0, hi_begin = 0, best = -1;
   for (int k = 0; k + 1 < m; k++)
   {
      const int gap = rows[(size_t)k + 1] - rows[(size_t)k] - 1;
      if (gap > best) { best = gap; lo_end = rows[(size_t)k] + 1; hi_begin = rows[(size_t)k + 1]; }
   }
   const int nloc = A_rows;
   if (rows[0] > best) { best = rows[0]; lo_end = 0; hi_begin = rows[0]; }
   if (nloc - 1 - rows[(size_t)m - 1] > best) { best = nloc - 1 - rows[(size_t)m - 1]; lo_end = rows[(size_t)m - 1] + 1; hi_begin = nloc; }
   if (best <= 0) { lo_end = hi_begin = 0; }
   // direct table: rows [0, lo_end) then rows [hi_begin, n) -> index of the exported row or -1
   std::vector<int> dir((size_t)lo_end + (size_t)(nloc - hi_begin) + 1, -1);
   for (int k = 0; k < m; k++)
   {
      const int r = rows[(size_t)k];
      dir[(size_t)(r < lo_end ? r : lo_end + (r - hi_begin))] = k;
   }
   HDK_TRY(dalloc(&I.exp_dir, dir.size()));
   HDK_TRY(dalloc(&I.exp_ptr, (size_t)m + 2));
   HDK_TRY(dalloc(&I.exp_slot, (size_t)H.n_send + 1));
   HDK_CUDA(cudaMemcpyAsync(I.exp_dir, dir.data(), sizeof(int) * dir.size(), cudaMemcpyHostToDevice, g.stream));
   HDK_CUDA(cudaMemcpyAsync(I.exp_ptr, ptr.data(), sizeof(int) * ((size_t)m + 1), cudaMemcpyHostToDevice, g.stream));
   HDK_CUDA(cudaMemcpyAsync(I.exp_slot, slot.data(), sizeof(int) * (size_t)H.n_send, cudaMemcpyHostToDevice, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream)); // the host vectors go out of scope
   I.exp_lo_end = lo_end; I.exp_hi_begin = hi_begin;
   return HDK_OK;
}

static bool export_enabled()
{
   static int on = -1;
   if (on < 0) { const char *e = getenv("HDK_HALO_EXPORT"); on = (e && atoi(e) == 0) ? 0 : 1; }
   return on == 1;
}

static void fill_send_args(const HaloPlan &H, IpcSendArgs &a)
{
   const IpcHalo &I = H.ipc;
   memset(&a, 0, sizeof(a));
   a.npeer = (int)H.send_rank.size();
   for (int p = 0; p < a.npeer; p++)
   {
      a.dst[p] = I.dst[I.seq & 1][p];
      a.off[p] = H.send_off[(size_t)p];
   }
   a.off[a.npeer] = H.n_send;
   a.nflag = I.nnbr;
   for (int p = 0; p < I.nnbr; p++) a.flag[p] = I.nbr_flag[p];
   a.rflag = I.data_flag; a.seq = I.seq; a.ticket = I.tickets;
}

bool halo_export_begin(const hdk_csr_s &A, HaloExport *e)
{
   *e = HaloExport();
   const HaloPlan &H = A.halo;
   if (g.nranks <= 1 || !H.ipc.on || !export_enabled() || H.ipc.preposted) return false;
   if (H.ipc.nnbr == 0 || H.n_send <= 0) return false;     // no part in the exchange / receives only: the flags are
   IpcHalo &I = const_cast<IpcHalo &>(H.ipc);              // raised by the (empty) pack kernel before the consumer
   I.seq++;
   I.preposted = true;
   e->dir = I.exp_dir; e->ptr = I.exp_ptr; e->slot = I.exp_slot; e->total = H.n_send;
   e->lo_end = I.exp_lo_end; e->hi_begin = I.exp_hi_begin;
   e->npeer = (int)H.send_rank.size();
   for (int p = 0; p < e->npeer; p++)
   {
      e->dst[p] = I.dst[I.seq & 1][p];
      e->off[p] = H.send_off[(size_t)p];
   }
   e->off[e->npeer] = H.n_send;
   e->nflag = I.nnbr;
   for (int p = 0; p < I.nnbr; p++) e->flag[p] = I.nbr_flag[p];
   e->rflag = I.data_flag; e->seq = I.seq; e->ticket = I.tickets + 1;
   return true;
}

void halo_export_cancel(const hdk_csr_s &A)
{
   // the data already sits in the neighbours' buffers under its sequence number; the next exchange
   // simply uses the following one
   const_cast<IpcHalo &>(A.halo.ipc).preposted = false;
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
   HDK_TRY(ipc_plan_build(H, all));
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
   return ipc_export_build(H, (int)(A.col_end - A.col_start + 1)); // the exported vector lives in my share of the column space
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
   if (H.ipc.on)
   {
      // peer-memory path: one kernel packs, stores over NVLink and signals; the consumer waits
      IpcHalo &I = const_cast<IpcHalo &>(H.ipc);
      if (I.preposted) { I.preposted = false; return HDK_OK; } // filled by the kernel that produced x (halo_export_begin)
      I.seq++;
      IpcSendArgs a;
      fill_send_args(H, a);
      k_pack_ipc<<<H.n_send > 0 ? cdiv(H.n_send, 256) : 1, H.n_send > 0 ? 256 : 32, 0, g.stream>>>(x, H.send_idx, H.n_send, a);
      HDK_LAUNCH_CHECK();
      return HDK_OK;
   }
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
   if (g.nranks <= 1 || A.halo.ipc.on) return HDK_OK;
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
   return ipc_setup();
}

// Communicator bootstrap for callers that have no channel of their own for the NCCL unique id (the C
// API, hypredrive-cli): the launcher's RANK / WORLD_SIZE / LOCAL_RANK say who we are, rank 0 publishes
// the id in a file every rank can see (HDK_NCCL_ID_FILE, default /tmp/hdk_nccl_id.<MASTER_ADDR>.<MASTER_PORT>),
// the others wait for it.  No-op with one process or when a communicator already exists.
int hdk_comm_init_from_env(void)
{
   HDK_TRY(require_init());
   const char *ws = getenv("WORLD_SIZE"), *rk = getenv("RANK");
   const int   world = ws ? atoi(ws) : 1, rank = rk ? atoi(rk) : 0;
   if (world <= 1 || g.nccl) return HDK_OK;
   char path[1024];
   const char *pf = getenv("HDK_NCCL_ID_FILE");
   if (pf && *pf) snprintf(path, sizeof(path), "%s", pf);
   else snprintf(path, sizeof(path), "/tmp/hdk_nccl_id.%s.%s", getenv("MASTER_ADDR") ? getenv("MASTER_ADDR") : "local",
                 getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0");
   unsigned char id[128];
   if (rank == 0)
   {
      HDK_TRY(hdk_comm_unique_id(id));
      char tmp[1100];
      snprintf(tmp, sizeof(tmp), "%s.tmp", path);
      FILE *fp = fopen(tmp, "wb");
      if (!fp || fwrite(id, 1, 128, fp) != 128) { if (fp) fclose(fp); return set_error(HDK_ERR_COMM, "cannot write the NCCL id file %s", tmp); }
      fclose(fp);
      if (rename(tmp, path)) return set_error(HDK_ERR_COMM, "cannot publish the NCCL id file %s", path);
   }
   else
   {
      // a file left behind by a run that died before rank 0 removed it must not be taken for this run's:
      // only files written at most two minutes before this rank got here count
      const time_t t_entry = time(nullptr);
      bool         got = false;
      for (int tries = 0; tries < 1200 && !got; tries++) // up to 2 minutes
      {
         struct stat st;
         FILE       *fp = (stat(path, &st) == 0 && st.st_mtime + 120 >= t_entry) ? fopen(path, "rb") : nullptr;
         if (fp) { got = fread(id, 1, 128, fp) == 128; fclose(fp); }
         if (!got) { struct timespec ts = {0, 100000000L}; nanosleep(&ts, nullptr); }
      }
      if (!got) return set_error(HDK_ERR_COMM, "timed out waiting for the NCCL id file %s (WORLD_SIZE=%d but rank 0 did not publish it)", path, world);
   }
   int rc = hdk_comm_init(rank, world, id);
   if (rc == HDK_OK)
   {
      int64_t s = 0;
      rc = hdk_comm_sum_i64(1, &s); // every rank has read the file
      if (rank == 0) remove(path);
   }
   return rc;
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

int hdk_comm_halo_mode(void) { return (g.nranks > 1) ? (ipc.on ? 2 : 1) : 0; }
int hdk_comm_rank(void) { return g.rank; }
int hdk_comm_size(void) { return g.nranks; }

int hdk_comm_finalize(void)
{
   ipc_teardown();
   if (g.nccl && nccl.CommDestroy) nccl.CommDestroy((ncclComm_p)g.nccl);
   g.nccl = nullptr; g.rank = 0; g.nranks = 1;
   return HDK_OK;
}

} // extern "C"
