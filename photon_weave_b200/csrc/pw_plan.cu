// Host-side index planning: validates (dims, targets) and builds the fiber/target maps.
#include "pw_internal.cuh"

namespace pw {

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
