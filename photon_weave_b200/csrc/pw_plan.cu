// Host-side index planning: validates (dims, targets) and builds the fiber/target maps.
#include "pw_internal.cuh"

#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <string>

namespace pw {

thread_local PlanCacheCtx* g_plan_ctx = nullptr;
std::atomic<uint64_t> g_options_epoch{0};

namespace {
std::mutex g_plan_mu;
std::set<const void*> g_pinned;
std::map<std::string, PlanCacheEntry*> g_plans;   // key bytes start with the operator pointer

void free_entry(PlanCacheEntry* e) {
  for (int i = 0; i < e->nbufs; ++i) cudaFree(e->bufs[i]);
  if (e->ready) cudaEventDestroy(e->ready);
  delete e;
}
}  // namespace

PlanScope::PlanScope(const void* op, int tag, int dtype, int K, const int64_t* dims, int n, const int32_t* t1,
                     int nt1, const int32_t* t2, int nt2, bool aliased, cudaStream_t stream) {
  stream_ = stream;
  if (!op || !dims || n < 1 || n > kMaxAxes || nt1 < 0 || nt1 > kMaxTargets || nt2 < 0 || nt2 > kMaxTargets) return;
  std::lock_guard<std::mutex> lock(g_plan_mu);
  if (g_pinned.find(op) == g_pinned.end()) return;
  int dev = 0;
  cudaGetDevice(&dev);
  std::string key((const char*)&op, sizeof(op));
  const int64_t head[8] = {tag, dtype, K, n, nt1, nt2, aliased ? 1 : 0, dev};
  key.append((const char*)head, sizeof(head));
  const uint64_t epoch = g_options_epoch.load();
  key.append((const char*)&epoch, sizeof(epoch));
  key.append((const char*)dims, sizeof(int64_t) * n);
  if (nt1) key.append((const char*)t1, sizeof(int32_t) * nt1);
  if (nt2) key.append((const char*)t2, sizeof(int32_t) * nt2);
  auto it = g_plans.find(key);
  PlanCacheEntry* e;
  if (it == g_plans.end()) {
    e = new PlanCacheEntry();
    e->fill_stream = stream;
    g_plans[key] = e;
  } else {
    e = it->second;
    if (!e->filled) return;   // another thread is recording it right now: run uncached
  }
  ctx.e = e;
  ctx.replay = e->filled;
  if (ctx.replay && stream != e->fill_stream && e->ready) {
    // inside a CUDA-graph capture the recording finished long ago (the caller warmed up and
    // synchronised before capturing): no cross-stream edge to an event from outside the capture
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap == cudaStreamCaptureStatusNone)
      cudaStreamWaitEvent(stream, e->ready, 0);
  }
  g_plan_ctx = &ctx;
  active = true;
}

int PlanScope::finish(int rc) {
  if (!active) return rc;
  g_plan_ctx = nullptr;
  active = false;
  std::lock_guard<std::mutex> lock(g_plan_mu);
  PlanCacheEntry* e = ctx.e;
  if (!ctx.replay) {
    if (rc == PW_OK && !ctx.broken) {
      if (cudaEventCreateWithFlags(&e->ready, cudaEventDisableTiming) == cudaSuccess) cudaEventRecord(e->ready, stream_);
      e->filled = true;
    } else {
      // a failed or unsupported recording is dropped (its buffers may still be in use by kernels
      // already enqueued: cudaFree waits for them)
      for (auto it = g_plans.begin(); it != g_plans.end(); ++it)
        if (it->second == e) { g_plans.erase(it); break; }
      free_entry(e);
    }
  }
  return rc;
}

PlanScope::~PlanScope() {
  if (active) finish(PW_ERR_INVALID);
}

std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_paths[5];
thread_local int g_last_cuda_error = 0;

void keep_pool_memory() {
  static thread_local int done_for = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return; }
  if (done_for == dev) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = ~0ull;
    (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  (void)cudaGetLastError();
  done_for = dev;
}

static int build_generic(const int64_t* dims, int n, const int* targets, int nt, Plan* plan) {
  if (n < 1 || n > kMaxAxes || nt < 0 || nt > kMaxTargets || !dims || (nt && !targets))
    return PW_ERR_INVALID;
  int64_t stride[kMaxAxes];
  uint64_t N = 1;
  for (int i = n - 1; i >= 0; --i) {
    if (dims[i] < 1) return PW_ERR_INVALID;
    stride[i] = (int64_t)N;
    if (N > (1ull << 62) / (uint64_t)dims[i]) return PW_ERR_INVALID;
    N *= (uint64_t)dims[i];
  }
  bool is_target[kMaxAxes] = {false};
  TargetMap& tm = plan->tm;
  tm.nt = nt;
  uint64_t t = 1;
  for (int k = 0; k < nt; ++k) {
    int a = targets[k];
    if (a < 0 || a >= n || is_target[a]) return PW_ERR_INVALID;
    is_target[a] = true;
    tm.tdim[k] = (uint32_t)dims[a];
    tm.tstride[k] = stride[a];
    t *= (uint64_t)dims[a];
    if (t > 0xffffffffull) return PW_ERR_UNSUPPORTED;
  }
  tm.t = (uint32_t)t;
  FiberMap& fm = plan->fm;
  fm.ng = 0;
  fm.nfib = 1;
  int i = 0;
  while (i < n) {
    if (is_target[i]) { ++i; continue; }
    uint64_t size = 1;
    int last = i;
    while (i < n && !is_target[i]) { size *= (uint64_t)dims[i]; last = i; ++i; }
    if (size == 1) continue;  // singleton axes contribute nothing
    fm.gsize[fm.ng] = size;
    fm.gstride[fm.ng] = stride[last];
    fm.nfib *= size;
    ++fm.ng;
  }
  plan->N = N;
  return PW_OK;
}

int build_plan_axes(const int64_t* axes, int naxes, const int* targets, int nt, Plan* plan) {
  return build_generic(axes, naxes, targets, nt, plan);
}

int build_plan(const int64_t* dims, int n, const int32_t* targets, int nt, Plan* plan) {
  if (n > PW_MAX_SUBSYSTEMS || nt > PW_MAX_TARGETS) return PW_ERR_INVALID;
  return build_generic(dims, n, targets, nt, plan);
}

int build_plan_matrix(const int64_t* dims, int n, const int32_t* targets, int nt, bool rows,
                      bool cols, Plan* plan) {
  if (n < 1 || n > PW_MAX_SUBSYSTEMS || nt < 0 || nt > PW_MAX_TARGETS || !dims) return PW_ERR_INVALID;
  int64_t d2[kMaxAxes];
  int t2[kMaxTargets];
  for (int i = 0; i < n; ++i) d2[i] = d2[n + i] = dims[i];
  int k2 = 0;
  for (int k = 0; k < nt; ++k) {
    if (targets[k] < 0 || targets[k] >= n) return PW_ERR_INVALID;
  }
  if (rows) for (int k = 0; k < nt; ++k) t2[k2++] = targets[k];
  if (cols) for (int k = 0; k < nt; ++k) t2[k2++] = n + targets[k];
  return build_generic(d2, 2 * n, t2, k2, plan);
}

}  // namespace pw

using namespace pw;

extern "C" int pw_op_pin(const void* op) {
  if (!op) return PW_ERR_INVALID;
  std::lock_guard<std::mutex> lock(g_plan_mu);
  g_pinned.insert(op);
  return PW_OK;
}

extern "C" int pw_op_unpin(const void* op) {
  if (!op) return PW_ERR_INVALID;
  std::lock_guard<std::mutex> lock(g_plan_mu);
  g_pinned.erase(op);
  for (auto it = g_plans.begin(); it != g_plans.end();) {
    const void* k;
    std::memcpy(&k, it->first.data(), sizeof(k));
    if (k == op) { free_entry(it->second); it = g_plans.erase(it); } else ++it;
  }
  return PW_OK;
}

extern "C" int pw_plan_cache_size(void) {
  std::lock_guard<std::mutex> lock(g_plan_mu);
  return (int)g_plans.size();
}
