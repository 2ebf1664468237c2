// Internal shared definitions for libpwb200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "pw_b200.h"

namespace pw {

constexpr int kMaxAxes = 2 * PW_MAX_SUBSYSTEMS;  // density matrices double the axes
constexpr int kMaxTargets = 2 * PW_MAX_TARGETS;  // row + column targets of a matrix
constexpr int kMaxGroups = kMaxTargets + 1;      // runs of non-target axes
constexpr int kNumSMs = 148;                     // B200

// ---- complex helpers ------------------------------------------------------------
template <typename R> struct CplxOf;
template <> struct CplxOf<float> { using type = float2; };
template <> struct CplxOf<double> { using type = double2; };
template <typename V> struct RealOf;
template <> struct RealOf<float2> { using type = float; };
template <> struct RealOf<double2> { using type = double; };

template <typename V>
__host__ __device__ __forceinline__ V czero() {
  V z; z.x = 0; z.y = 0; return z;
}
// acc += a * b
template <typename V>
__device__ __forceinline__ void cfma(V& acc, const V a, const V b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a) * b
template <typename V>
__device__ __forceinline__ void cfma_conja(V& acc, const V a, const V b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
// two adjacent complex64 values moved as ONE 128-bit access (fiber pairs of the complex64
// register kernels: the same bytes per thread and per instruction as complex128)
struct __align__(16) C64x2 { float2 a, b; };
template <>
__host__ __device__ __forceinline__ C64x2 czero<C64x2>() {
  C64x2 z; z.a = czero<float2>(); z.b = czero<float2>(); return z;
}
__device__ __forceinline__ void cfma(C64x2& acc, const float2 o, const C64x2 x) {
  cfma(acc.a, o, x.a);
  cfma(acc.b, o, x.b);
}
__device__ __forceinline__ C64x2 ld_c64x2(const float2* p) {
  const float4 v = __ldcg(reinterpret_cast<const float4*>(p));
  C64x2 r; r.a = make_float2(v.x, v.y); r.b = make_float2(v.z, v.w); return r;
}
__device__ __forceinline__ void st_c64x2(float2* p, const C64x2 v) {
  __stcg(reinterpret_cast<float4*>(p), make_float4(v.a.x, v.a.y, v.b.x, v.b.y));
}

template <typename V>
__device__ __forceinline__ typename RealOf<V>::type cabs2(const V a) {
  return fma(a.x, a.x, a.y * a.y);
}

// ---- index maps -------------------------------------------------------------------
// The state is a row-major tensor.  Target axes (in operator order) carry the
// operator index j = mixed-radix(tdim) -> element offset sum digit*tstride.  The
// remaining axes are merged into runs ("groups") of consecutive axes; a fiber id
// f in [0, nfib) is mixed-radix over the groups in tensor order (last fastest).
// `estride` multiplies every offset (N+1 walks the diagonal of a density matrix).
struct FiberMap {
  int ng;
  uint64_t gsize[kMaxGroups];
  int64_t gstride[kMaxGroups];
  uint64_t nfib;

  __host__ __device__ __forceinline__ int64_t base(uint64_t f) const {
    int64_t off = 0;
    if (nfib <= 0xffffffffull) {
      uint32_t g = (uint32_t)f;
#pragma unroll 1
      for (int i = ng - 1; i > 0; --i) {
        uint32_t gs = (uint32_t)gsize[i];
        uint32_t q = g / gs;
        off += (int64_t)(g - q * gs) * gstride[i];
        g = q;
      }
      if (ng > 0) off += (int64_t)g * gstride[0];
    } else {
#pragma unroll 1
      for (int i = ng - 1; i > 0; --i) {
        uint64_t q = f / gsize[i];
        off += (int64_t)(f - q * gsize[i]) * gstride[i];
        f = q;
      }
      if (ng > 0) off += (int64_t)f * gstride[0];
    }
    return off;
  }
};

struct TargetMap {
  int nt;
  uint32_t tdim[kMaxTargets];
  int64_t tstride[kMaxTargets];
  uint32_t t;  // prod(tdim)

  __host__ __device__ __forceinline__ int64_t offset(uint32_t j) const {
    int64_t off = 0;
#pragma unroll 1
    for (int i = nt - 1; i >= 0; --i) {
      uint32_t q = j / tdim[i];
      off += (int64_t)(j - q * tdim[i]) * tstride[i];
      j = q;
    }
    return off;
  }
};

// fiber-pair map of `fm`, or false: the innermost group must be the contiguous direction with an
// even extent (then every other stride and every target offset is even: 16-byte aligned pairs)
inline bool pair_fiber_map(const FiberMap& fm, const void* in, const void* out, FiberMap* pm) {
  if (fm.ng < 1 || fm.gstride[fm.ng - 1] != 1 || (fm.gsize[fm.ng - 1] & 1) != 0) return false;
  if ((((uintptr_t)in | (uintptr_t)out) & 15) != 0) return false;
  *pm = fm;
  pm->gsize[fm.ng - 1] = fm.gsize[fm.ng - 1] / 2;
  pm->gstride[fm.ng - 1] = 2;
  pm->nfib = fm.nfib / 2;
  return true;
}


struct Plan {
  uint64_t N;  // total elements of the tensor the maps index
  FiberMap fm;
  TargetMap tm;
};

// Build maps for a tensor with `n` axes `dims` and target axes `targets` (given
// order).  Returns PW_OK / PW_ERR_INVALID.
int build_plan(const int64_t* dims, int n, const int32_t* targets, int nt, Plan* plan);
// Doubled (density-matrix) tensor: axes dims ++ dims, targets rows then columns.
int build_plan_matrix(const int64_t* dims, int n, const int32_t* targets, int nt,
                      bool rows, bool cols, Plan* plan);
// Arbitrary axis list (up to kMaxAxes axes / kMaxTargets targets), e.g. dims ++ dims.
int build_plan_axes(const int64_t* axes, int naxes, const int* targets, int nt, Plan* plan);

// ---- bookkeeping -----------------------------------------------------------------
extern std::atomic<uint64_t> g_launches;
extern std::atomic<uint64_t> g_paths[5];  // small, generic, tiled, tiled-rejected, blocks
extern thread_local int g_last_cuda_error;

inline int cuda_status(cudaError_t e) {
  if (e == cudaSuccess) return PW_OK;
  g_last_cuda_error = (int)e;
  return PW_ERR_CUDA;
}
#define PW_CUDA(call)                                   \
  do {                                                  \
    int _s = ::pw::cuda_status((call));                 \
    if (_s != PW_OK) return _s;                         \
  } while (0)
#define PW_TRY(call)                \
  do {                              \
    int _s = (call);                \
    if (_s != PW_OK) return _s;     \
  } while (0)
#define PW_LAUNCHED()                                   \
  do {                                                  \
    ::pw::g_launches.fetch_add(1, std::memory_order_relaxed); \
    PW_CUDA(cudaGetLastError());                        \
  } while (0)

inline int grid_for(uint64_t work_items, int block, int max_ctas_per_sm) {
  uint64_t g = (work_items + block - 1) / block;
  uint64_t cap = (uint64_t)kNumSMs * max_ctas_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Keep freed scratch memory cached in the device's default pool (the default release
// threshold of 0 hands multi-GB scratch back to the driver at every synchronisation).
void keep_pool_memory();

// ---- analysed-plan cache ------------------------------------------------------------------
// An apply / Kraus call analyses its operator on the device (block structure, super-blocks,
// superoperator, CSR, diagonal test) into scratch buffers before the kernel that does the work.
// For an operator the caller has PINNED (pw_op_pin: "this device memory stays alive and
// unchanged") those buffers are kept: the first call with a given (operator, dims, targets,
// dtype, aliasing, options) records them, later calls REPLAY -- Scratch::alloc hands the recorded
// buffers back in the same order and the analysis kernels (guarded by plan_replay()) are not
// launched again.  Descriptors are read-only once built, so replays on other streams only wait
// for the event recorded after the build.
struct PlanCacheEntry {
  void* bufs[24];
  size_t sizes[24];
  int nbufs = 0;
  cudaEvent_t ready = nullptr;
  cudaStream_t fill_stream = nullptr;
  bool filled = false;
};
struct PlanCacheCtx {
  PlanCacheEntry* e = nullptr;
  int cursor = 0;
  bool replay = false;
  bool broken = false;   // replay did not match the recording (never expected): the call fails
};
extern thread_local PlanCacheCtx* g_plan_ctx;
inline bool plan_replay() { return g_plan_ctx != nullptr && g_plan_ctx->replay; }

// Stream-ordered scratch buffer (freed on the same stream when it goes out of scope); inside a
// cached call: a persistent buffer owned by the plan cache.
struct Scratch {
  void* p = nullptr;
  cudaStream_t s = nullptr;
  bool cached = false;
  int alloc(size_t bytes, cudaStream_t stream) {
    s = stream;
    if (bytes == 0) bytes = 16;
    if (PlanCacheCtx* c = g_plan_ctx) {
      PlanCacheEntry* e = c->e;
      cached = true;
      if (c->replay) {
        if (c->cursor >= e->nbufs || e->sizes[c->cursor] != bytes) { c->broken = true; p = nullptr; return PW_ERR_WORKSPACE; }
        p = e->bufs[c->cursor++];
        return PW_OK;
      }
      if (e->nbufs >= 24 || cudaMalloc(&p, bytes) != cudaSuccess) {
        p = nullptr;
        (void)cudaGetLastError();
        c->broken = true;
        return PW_ERR_WORKSPACE;
      }
      e->bufs[e->nbufs] = p;
      e->sizes[e->nbufs++] = bytes;
      return PW_OK;
    }
    return alloc_temp(bytes, stream);
  }
  // state-sized temporaries: never cached
  int alloc_temp(size_t bytes, cudaStream_t stream) {
    s = stream;
    cached = false;
    keep_pool_memory();
    if (cudaMallocAsync(&p, bytes ? bytes : 16, stream) != cudaSuccess) {
      p = nullptr;
      (void)cudaGetLastError();
      return PW_ERR_WORKSPACE;
    }
    return PW_OK;
  }
  ~Scratch() {
    if (p && !cached) cudaFreeAsync(p, s);
  }
};

// RAII around one C-ABI apply call: activates the plan cache when `op` is pinned.
struct PlanScope {
  PlanCacheCtx ctx;
  bool active = false;
  PlanScope(const void* op, int tag, int dtype, int K, const int64_t* dims, int n, const int32_t* t1, int nt1,
            const int32_t* t2, int nt2, bool aliased, cudaStream_t stream);
  int finish(int rc);   // call with the dispatch's status; returns the status to hand to the caller
  ~PlanScope();
  cudaStream_t stream_ = nullptr;
};

// ---- tuning knobs (pw_set_option / PW_* environment) -------------------------
extern std::atomic<uint64_t> g_options_epoch;  // bumped by pw_set_option: cached plans of older epochs are not reused
struct Options {
  std::atomic<int> vec_path{0}, mat_mode{0}, tile_elems{0}, ctas_per_sm{0};
  std::atomic<int> small_r{16};                // inner untouched run below this -> tiled staging
  std::atomic<int> mat_super_single{0};        // single operators on rho via superoperator tiles
  std::atomic<int> contig_kernel{1};           // warp-transposing kernel for contiguous fibers
  std::atomic<int> diag_kernel{1};             // diagonal t <= 8 operators on large vectors: one streaming pass
  std::atomic<int> contig_mma{1};              // ... on the FP64 tensor cores (complex128, one 5..8-level target)
  std::atomic<int> fuse_pairs{1};              // circuit entry points fuse adjacent single-axis ops
  std::atomic<int> host_overlap{256};          // pw_circuit_host: states of at least this many MiB upload slab-wise, overlapped with the operations that allow it (0 = off)
  std::atomic<int> c64_pairs{1};              // complex64 register / block kernels: two adjacent fibers per thread (128-bit accesses) when the contiguous untouched run is even
  std::atomic<int> multi_f{0};                 // resident-tile kernel: fibers per thread (0 = auto)
  std::atomic<int> fuse_group{3};              // ... up to this many per HBM pass (resident-tile kernel, 2..4)
  std::atomic<int> mat_pair_max_d{4};          // single-pass U rho U^dagger register tile up to this d
  std::atomic<int> seg_kernel{1};              // warp-staged block kernel for innermost-target layouts
  std::atomic<int> nstage{0};                  // tiled kernel pipeline depth (0 = auto)
  std::atomic<int> blocks_min_elems{1 << 18};  // below this, skip structure analysis
  std::atomic<int> blocks_min_elems_small_t{1 << 22};  // same, single t <= 8 operator on a density matrix
  std::atomic<int> two_sided{1};               // single-pass U rho U^dagger through shared-memory tiles
  std::atomic<int> ts_ctas{0};                 // CTAs per SM of that kernel (0 = default 2)
  std::atomic<int> seg_mma{1};                 // staged block kernel on the FP64 tensor cores (complex128, t <= 64)
  std::atomic<int> seg_fpw{0};                 // staged block kernel: fibers per warp stage (0 = auto; 32, 16, 8)
  std::atomic<int> dense_mma{1};               // dense complex128 operators, 8 < t <= 64: FP64 tensor-core kernel
  std::atomic<int> ts_order{0};                // tensor-core kernel, several super-blocks: 0 chunk-fastest, 1 pair-fastest unit order
  std::atomic<int> draw_exact{1};              // pw_draw: re-run the sequential cumsum when the draw is within rounding distance of a bin edge
  std::atomic<int> ts_mma_c64{1};              // complex64: the same FP64 tensor-core kernels with float2 storage
  std::atomic<int> ts_mma{1};                  // complex128: FP64 tensor cores (mma.sync m8n8k4) in that kernel
};
Options& options();

// ---- internal launchers (one per .cu) ------------------------------------------
// out = [out +] (O on plan.tm) in;  conj_op: use conj(O).  in may alias out unless
// accumulate.  `op` is a device pointer, row-major (t x t).
// `skip`: optional device flag; when *skip != 0 the kernels exit at once (another path
// already produced the result).
int launch_apply(const void* in, void* out, const void* op, const Plan& plan, int dtype,
                 bool conj_op, bool accumulate, cudaStream_t stream, const int* skip = nullptr);

// row-parallel dense apply for small states (pw_evolve.cu); PW_ERR_UNSUPPORTED if in == out
int launch_apply_rows(const void* in, void* out, const void* op, const Plan& plan, int dtype,
                      bool conj_op, cudaStream_t stream);

}  // namespace pw
