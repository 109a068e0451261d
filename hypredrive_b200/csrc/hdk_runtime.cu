// hdk_runtime.cu -- context, error reporting, raw memory entry points of the hdk C-ABI.
// Stands in for HYPRE_Initialize / HYPRE_Finalize and hypre's device memory pool
// (reference: src/internal/runtime.c:101-133, src/HYPREDRV.c:308-349).
#include "hdk_internal.cuh"
#include <stdarg.h>
#include <stdlib.h>

namespace hdk {

Ctx g;

int set_error(int code, const char *fmt, ...)
{
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(g.err, sizeof(g.err), fmt, ap);
   va_end(ap);
   return code;
}

int require_init()
{
   if (!g.inited) return hdk_init(-1);
   return HDK_OK;
}

} // namespace hdk

using namespace hdk;

extern "C" {

const char *hdk_last_error(void) { return g.err; }

int hdk_device_count(void)
{
   int         n = 0;
   cudaError_t e = cudaGetDeviceCount(&n);
   if (e != cudaSuccess) { cudaGetLastError(); return 0; }
   return n;
}

int hdk_init(int device)
{
   if (g.inited) return HDK_OK;
   int n = hdk_device_count();
   if (n <= 0)
      return set_error(HDK_ERR_NO_DEVICE,
                       "no CUDA device visible: hypredrive_b200 has no CPU fallback");
   if (device < 0)
   {
      const char *lr = getenv("LOCAL_RANK");
      const char *nd = getenv("HYPREDRV_DEVICE");
      if (nd) device = atoi(nd);
      else if (lr) device = atoi(lr) % n;
      else device = 0;
   }
   if (device >= n) device = device % n;
   HDK_CUDA(cudaSetDevice(device));
   g.device = device;
   cudaDeviceProp prop;
   HDK_CUDA(cudaGetDeviceProperties(&prop, device));
   g.sm_count = prop.multiProcessorCount;
   HDK_CUDA(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
   // halo exchange stream: highest priority so NCCL's few CTAs are placed before the persistent
   // SpMV grid that becomes runnable at the same moment (both wait for the pack kernel)
   int prio_lo = 0, prio_hi = 0;
   HDK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
   HDK_CUDA(cudaStreamCreateWithPriority(&g.comm_stream, cudaStreamNonBlocking, prio_hi));
   HDK_CUDA(cudaEventCreate(&g.ev_a));
   HDK_CUDA(cudaEventCreate(&g.ev_b));
   HDK_CUDA(cudaEventCreateWithFlags(&g.ev_scal, cudaEventDisableTiming));
   HDK_CUDA(cudaEventCreateWithFlags(&g.ev_halo, cudaEventDisableTiming));
   HDK_CUDA(cudaEventCreateWithFlags(&g.ev_pack, cudaEventDisableTiming));
   // keep freed blocks cached in the pool: setup allocates and frees large scratch per level
   cudaMemPool_t pool;
   HDK_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
   uint64_t thresh = UINT64_MAX;
   HDK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
   HDK_CUDA(cudaMalloc((void **)&g.partials, sizeof(double) * PARTIALS_CAP));
   HDK_CUDA(cudaMalloc((void **)&g.counters, sizeof(unsigned) * 64));
   HDK_CUDA(cudaMemset(g.counters, 0, sizeof(unsigned) * 64));
   HDK_CUDA(cudaMalloc((void **)&g.dscal, sizeof(double) * 64));
   HDK_CUDA(cudaMemset(g.dscal, 0, sizeof(double) * 64));
   HDK_CUDA(cudaMallocHost((void **)&g.hscal, sizeof(double) * 64));
   g.inited = true;
   g.err[0] = 0;
   return HDK_OK;
}

int hdk_finalize(void)
{
   if (!g.inited) return HDK_OK;
   cudaStreamSynchronize(g.stream);
   cudaStreamSynchronize(g.comm_stream);
   hdk_comm_finalize();
   cudaFree(g.partials); cudaFree(g.counters); cudaFree(g.dscal); cudaFreeHost(g.hscal);
   cudaEventDestroy(g.ev_a); cudaEventDestroy(g.ev_b); cudaEventDestroy(g.ev_scal);
   cudaEventDestroy(g.ev_halo); cudaEventDestroy(g.ev_pack);
   cudaStreamDestroy(g.stream); cudaStreamDestroy(g.comm_stream);
   g = Ctx();
   return HDK_OK;
}

int hdk_sync(void)
{
   HDK_TRY(require_init());
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   return HDK_OK;
}

void *hdk_stream(void) { return (void *)g.stream; }

int64_t hdk_launch_count_reset(void)
{
   int64_t n = g.launches;
   g.launches = 0;
   return n;
}

int hdk_malloc_device(void **p_d, size_t bytes)
{
   HDK_TRY(require_init());
   char *p;
   HDK_TRY(dalloc(&p, bytes));
   *p_d = p;
   return HDK_OK;
}

int hdk_free_device(void *p_d)
{
   if (!g.inited) return HDK_OK;
   dfree(p_d);
   return HDK_OK;
}

int hdk_copy_d2h(void *dst_h, const void *src_d, size_t bytes)
{
   HDK_TRY(require_init());
   HDK_CUDA(cudaMemcpyAsync(dst_h, src_d, bytes, cudaMemcpyDeviceToHost, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   return HDK_OK;
}

int hdk_copy_h2d(void *dst_d, const void *src_h, size_t bytes)
{
   HDK_TRY(require_init());
   HDK_CUDA(cudaMemcpyAsync(dst_d, src_h, bytes, cudaMemcpyHostToDevice, g.stream));
   HDK_CUDA(cudaStreamSynchronize(g.stream));
   return HDK_OK;
}

} // extern "C"

// ---- timeline (diagnostics): hdk_tune("timeline", 1) starts recording one CUDA event before every
// operation of the V-cycle / PCG loop, hdk_tune("timeline", 0) prints the average time of each
// (level, operation) segment -- launch gaps and waits included -- to stderr and stops.
#include <map>
#include <cstring>
namespace hdk {
struct TlMark { cudaEvent_t ev; int code; };
static std::vector<TlMark> tl_marks;
static bool                tl_on = false;
void tl_mark(int level, int op)
{
   if (!tl_on || tl_marks.size() >= 1000000) return; // (a forgotten timeline does not grow without bound)
   TlMark m;
   m.code = level * 16 + op;
   if (cudaEventCreate(&m.ev) != cudaSuccess) return;
   cudaEventRecord(m.ev, g.stream);
   tl_marks.push_back(m);
}
static void tl_dump()
{
   static const char *names[16] = {"cycle-end", "residual", "restrict", "prolong", "post-sweep", "first-sweep", "tail", "pcg-spmv",
                                   "pcg-xr", "pcg-p", "pre-sweep", "gather", "?", "?", "?", "?"};
   if (tl_marks.empty()) return;
   cudaEventSynchronize(tl_marks.back().ev);
   std::map<int, std::pair<double, int>> acc;
   for (size_t i = 0; i + 1 < tl_marks.size(); i++)
   {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, tl_marks[i].ev, tl_marks[i + 1].ev);
      auto &a = acc[tl_marks[i].code];
      a.first += ms; a.second++;
   }
   double tot = 0.0;
   for (auto &kv : acc) tot += kv.second.first;
   fprintf(stderr, "[hdk timeline rank %d] %zu marks, %.3f ms\n", g.rank, tl_marks.size(), tot);
   for (auto &kv : acc)
         fprintf(stderr, "[hdk timeline rank %d] level %d %-11s n=%3d avg %8.2f us total %8.3f ms\n", g.rank, kv.first >> 4,
                 names[kv.first & 15], kv.second.second, 1e3 * kv.second.first / kv.second.second, kv.second.first);
   for (auto &m : tl_marks) cudaEventDestroy(m.ev);
   tl_marks.clear();
}
} // namespace hdk

extern "C" int hdk_tune(const char *key, double value)
{
   if (!key) return hdk::set_error(HDK_ERR_INVALID, "hdk_tune: null key");
   if (!strcmp(key, "timeline"))
   {
      if (value != 0.0) hdk::tl_on = true;
      else { hdk::tl_on = false; hdk::tl_dump(); }
      return HDK_OK;
   }
   return hdk::tune_set(key, value);
}

// cudaProfilerStart/Stop bracket for `ncu --profile-from-start off` (bench.py HDK_PROFILE_RANGE=1)
#include <cuda_profiler_api.h>
extern "C" int hdk_profiler_range(int start)
{
   HDK_TRY(hdk::require_init());
   cudaStreamSynchronize(hdk::g.stream);
   if (start) cudaProfilerStart(); else cudaProfilerStop();
   return HDK_OK;
}
