// apply_operation / Kraus kernels (rows a1-a6 of SURVEY.md section 8).
//
// The state is never transposed: a fiber (all operator-index positions for one
// setting of the untouched subsystems) is addressed through FiberMap/TargetMap,
// multiplied by the t x t operator and written back to the same positions.
#include <cstdlib>

#include "pw_internal.cuh"
#include "pw_tiled.cuh"

namespace pw {

// ---------------------------------------------------------------------------------
// Register path: t <= 8.  One thread owns one fiber: T vector loads (16 B for
// complex128), T*T complex FMAs against the operator broadcast from shared
// memory, T vector stores.  Consecutive threads walk the fastest non-target run,
// so when that run is the trailing one every load/store instruction of a warp
// covers a contiguous 512 B (c128) segment.
// ---------------------------------------------------------------------------------
// Register budget: the operator must stay in shared memory (broadcast LDS), not be
// hoisted into registers -- occupancy (bytes in flight), not FLOPs, bounds this kernel.
template <typename V, int T>
constexpr int small_min_blocks() {
  return (sizeof(V) == 16) ? (T <= 4 ? 4 : (T <= 6 ? 3 : 2)) : (T <= 6 ? 4 : 3);
}

template <typename V, int T, bool CONJ, bool ACC>
__global__ void __launch_bounds__(256, small_min_blocks<V, T>())
k_apply_small(const V* in, V* out, const V* __restrict__ op, const FiberMap fm,
              const TargetMap tm) {
  __shared__ V s_op[T * T];
  __shared__ int64_t s_off[T];
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    V o = op[i];
    if (CONJ) o.y = -o.y;
    s_op[i] = o;
  }
  if (threadIdx.x < T) s_off[threadIdx.x] = tm.offset(threadIdx.x);
  __syncthreads();

  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib; f += stride) {
    const int64_t base = fm.base(f);
    // opaque zero: keeps the operator reads inside the loop (broadcast LDS) instead of
    // letting the compiler hoist T*T complex values into registers and spill
    int zero = 0;
    asm volatile("" : "+r"(zero));
    const V* sop = s_op + zero;
    V x[T];
#pragma unroll
    for (int j = 0; j < T; ++j) x[j] = in[base + s_off[j]];
#pragma unroll
    for (int a = 0; a < T; ++a) {
      V acc = czero<V>();
      if (ACC) acc = out[base + s_off[a]];
#pragma unroll
      for (int b = 0; b < T; ++b) cfma(acc, sop[a * T + b], x[b]);
      out[base + s_off[a]] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------
// Generic path: any t.  One thread owns one fiber and re-reads it once per output
// row (L1/L2 hits); operator rows stream from global/L2.  Requires in != out.
// Correctness fallback for shapes the tiled kernels do not cover.
// ---------------------------------------------------------------------------------
template <typename V, bool CONJ, bool ACC>
__global__ void __launch_bounds__(128)
k_apply_generic(const V* __restrict__ in, V* __restrict__ out, const V* __restrict__ op,
                const FiberMap fm, const TargetMap tm, const int64_t* __restrict__ toff) {
  const uint32_t t = tm.t;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib; f += stride) {
    const int64_t base = fm.base(f);
    for (uint32_t a = 0; a < t; ++a) {
      V acc = czero<V>();
      if (ACC) acc = out[base + toff[a]];
      const V* row = op + (size_t)a * t;
      for (uint32_t b = 0; b < t; ++b) {
        V o = row[b];
        if (CONJ) o.y = -o.y;
        cfma(acc, o, in[base + toff[b]]);
      }
      out[base + toff[a]] = acc;
    }
  }
}

__global__ void k_target_offsets(const TargetMap tm, int64_t* toff) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < tm.t) toff[j] = tm.offset(j);
}

template <typename V, int T>
static int launch_small(const V* in, V* out, const V* op, const Plan& p, bool conj_op,
                        bool acc, cudaStream_t s) {
  const int block = 256;
  const int grid = grid_for(p.fm.nfib, block, 8);
  g_paths[0].fetch_add(1, std::memory_order_relaxed);
  if (conj_op) {
    if (acc) k_apply_small<V, T, true, true><<<grid, block, 0, s>>>(in, out, op, p.fm, p.tm);
    else     k_apply_small<V, T, true, false><<<grid, block, 0, s>>>(in, out, op, p.fm, p.tm);
  } else {
    if (acc) k_apply_small<V, T, false, true><<<grid, block, 0, s>>>(in, out, op, p.fm, p.tm);
    else     k_apply_small<V, T, false, false><<<grid, block, 0, s>>>(in, out, op, p.fm, p.tm);
  }
  PW_LAUNCHED();
  return PW_OK;
}

template <typename V>
static int launch_apply_t(const V* in, V* out, const V* op, const Plan& p, bool conj_op,
                          bool acc, cudaStream_t s) {
  switch (p.tm.t) {
    case 1: return launch_small<V, 1>(in, out, op, p, conj_op, acc, s);
    case 2: return launch_small<V, 2>(in, out, op, p, conj_op, acc, s);
    case 3: return launch_small<V, 3>(in, out, op, p, conj_op, acc, s);
    case 4: return launch_small<V, 4>(in, out, op, p, conj_op, acc, s);
    case 5: return launch_small<V, 5>(in, out, op, p, conj_op, acc, s);
    case 6: return launch_small<V, 6>(in, out, op, p, conj_op, acc, s);
    case 7: return launch_small<V, 7>(in, out, op, p, conj_op, acc, s);
    case 8: return launch_small<V, 8>(in, out, op, p, conj_op, acc, s);
    default: break;
  }
  g_paths[1].fetch_add(1, std::memory_order_relaxed);
  // generic: needs a non-aliased source and the offset table in global memory
  Scratch toff, tmp;
  PW_TRY(toff.alloc(sizeof(int64_t) * p.tm.t, s));
  k_target_offsets<<<(p.tm.t + 127) / 128, 128, 0, s>>>(p.tm, (int64_t*)toff.p);
  PW_LAUNCHED();
  const V* src = in;
  if ((const void*)in == (const void*)out) {
    PW_TRY(tmp.alloc(sizeof(V) * p.N, s));
    PW_CUDA(cudaMemcpyAsync(tmp.p, in, sizeof(V) * p.N, cudaMemcpyDeviceToDevice, s));
    src = (const V*)tmp.p;
  }
  const int block = 128;
  const int grid = grid_for(p.fm.nfib, block, 8);
  if (conj_op) {
    if (acc) k_apply_generic<V, true, true><<<grid, block, 0, s>>>(src, out, op, p.fm, p.tm, (int64_t*)toff.p);
    else     k_apply_generic<V, true, false><<<grid, block, 0, s>>>(src, out, op, p.fm, p.tm, (int64_t*)toff.p);
  } else {
    if (acc) k_apply_generic<V, false, true><<<grid, block, 0, s>>>(src, out, op, p.fm, p.tm, (int64_t*)toff.p);
    else     k_apply_generic<V, false, false><<<grid, block, 0, s>>>(src, out, op, p.fm, p.tm, (int64_t*)toff.p);
  }
  PW_LAUNCHED();
  return PW_OK;
}

int launch_apply(const void* in, void* out, const void* op, const Plan& plan, int dtype,
                 bool conj_op, bool accumulate, cudaStream_t stream) {
  if (dtype == PW_C128)
    return launch_apply_t<double2>((const double2*)in, (double2*)out, (const double2*)op, plan,
                                   conj_op, accumulate, stream);
  if (dtype == PW_C64)
    return launch_apply_t<float2>((const float2*)in, (float2*)out, (const float2*)op, plan,
                                  conj_op, accumulate, stream);
  return PW_ERR_INVALID;
}

static size_t elem_size(int dtype) { return dtype == PW_C128 ? 16 : 8; }

}  // namespace pw

using namespace pw;

static int vec_path() { return options().vec_path.load(std::memory_order_relaxed); }
static int mat_mode() { return options().mat_mode.load(std::memory_order_relaxed); }
static uint32_t tile_elems(int dtype) {
  const int v = options().tile_elems.load(std::memory_order_relaxed);
  if (v > 0) return (uint32_t)v;
  return dtype == PW_C128 ? 1024u : 2048u;
}

extern "C" int pw_apply_vec(const void* psi, void* out, const void* op, const int64_t* dims,
                            int n, const int32_t* targets, int nt, int dtype, void* stream) {
  if (!psi || !out || !op || nt < 1) return PW_ERR_INVALID;
  if (dtype != PW_C64 && dtype != PW_C128) return PW_ERR_INVALID;
  Plan p;
  PW_TRY(build_plan(dims, n, targets, nt, &p));
  const bool want_tiled = vec_path() == 2 || (vec_path() == 0 && p.tm.t > 8);
  if (want_tiled) {
    TiledGeom g;
    if (plan_tiled(dims, n, targets, nt, nullptr, 0, -1, elem_size(dtype), tile_elems(dtype), &g)) {
      const int rc = launch_tiled(psi, out, op, 1, g, kModeVector, dtype, (cudaStream_t)stream);
      if (rc != PW_ERR_UNSUPPORTED) return rc;
    }
    g_paths[3].fetch_add(1, std::memory_order_relaxed);
  }
  return launch_apply(psi, out, op, p, dtype, false, false, (cudaStream_t)stream);
}

// One HBM pass over rho when the (row target, column target) tile fits shared memory.
static int try_tiled_matrix(const void* rho, void* out, const void* ops, int K, const int64_t* dims,
                            int n, const int32_t* targets, int nt, int dtype, cudaStream_t s) {
  if (mat_mode() == 1 || n > PW_MAX_SUBSYSTEMS || nt > PW_MAX_TARGETS) return PW_ERR_UNSUPPORTED;
  int64_t axes[kMaxAxes];
  int rows[kMaxTargets], cols[kMaxTargets], both[kMaxTargets];
  uint64_t t = 1;
  for (int i = 0; i < n; ++i) axes[i] = axes[n + i] = dims[i];
  for (int k = 0; k < nt; ++k) {
    if (targets[k] < 0 || targets[k] >= n) return PW_ERR_INVALID;
    rows[k] = both[k] = targets[k];
    cols[k] = both[nt + k] = n + targets[k];
    t *= (uint64_t)dims[targets[k]];
  }
  bool super = K > 1 && t * t <= 256 && (uint64_t)K * 2 >= t;
  if (mat_mode() == 2) super = false;
  if (mat_mode() == 3) super = t * t <= 1024;
  TiledGeom g;
  if (super) {
    if (!plan_tiled(axes, 2 * n, both, 2 * nt, nullptr, 0, -1, elem_size(dtype), tile_elems(dtype), &g))
      return PW_ERR_UNSUPPORTED;
    return launch_tiled(rho, out, ops, K, g, kModeMatrixSuper, dtype, s);
  }
  if (!plan_tiled(axes, 2 * n, rows, nt, cols, nt, -1, elem_size(dtype), tile_elems(dtype), &g))
    return PW_ERR_UNSUPPORTED;
  return launch_tiled(rho, out, ops, K, g, kModeMatrixTwoSided, dtype, s);
}

// one side of a density-matrix update as a (possibly tiled) strided apply
static int apply_side(const void* in, void* out, const void* op, const int64_t* dims, int n,
                      const int32_t* targets, int nt, bool cols, bool accumulate, int dtype,
                      cudaStream_t s) {
  Plan p;
  PW_TRY(build_plan_matrix(dims, n, targets, nt, !cols, cols, &p));
  return launch_apply(in, out, op, p, dtype, cols, accumulate, s);
}

extern "C" int pw_apply_mat(const void* rho, void* out, const void* op, const int64_t* dims,
                            int n, const int32_t* targets, int nt, int dtype, void* stream) {
  if (!rho || !out || !op || nt < 1) return PW_ERR_INVALID;
  if (dtype != PW_C64 && dtype != PW_C128) return PW_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  const int rc = try_tiled_matrix(rho, out, op, 1, dims, n, targets, nt, dtype, s);
  if (rc != PW_ERR_UNSUPPORTED) return rc;
  g_paths[3].fetch_add(1, std::memory_order_relaxed);
  // fallback, two HBM passes: left multiply on the row axes, conj(O) on the column axes
  PW_TRY(apply_side(rho, out, op, dims, n, targets, nt, false, false, dtype, s));
  return apply_side(out, out, op, dims, n, targets, nt, true, false, dtype, s);
}

extern "C" int pw_kraus_mat(const void* rho, void* out, const void* ops, int K,
                            const int64_t* dims, int n, const int32_t* targets, int nt,
                            int dtype, void* stream) {
  if (!rho || !out || !ops || nt < 1 || K < 1) return PW_ERR_INVALID;
  if (dtype != PW_C64 && dtype != PW_C128) return PW_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  const int rc = try_tiled_matrix(rho, out, ops, K, dims, n, targets, nt, dtype, s);
  if (rc != PW_ERR_UNSUPPORTED) return rc;
  g_paths[3].fetch_add(1, std::memory_order_relaxed);
  Plan rows;
  PW_TRY(build_plan_matrix(dims, n, targets, nt, true, false, &rows));
  const size_t es = elem_size(dtype);
  const size_t op_bytes = (size_t)rows.tm.t * rows.tm.t * es;
  if (K == 1) {
    PW_TRY(apply_side(rho, out, ops, dims, n, targets, nt, false, false, dtype, s));
    return apply_side(out, out, ops, dims, n, targets, nt, true, false, dtype, s);
  }
  // fallback: out = sum_k (K_k rho) K_k^dagger; tmp holds K_k rho, second pass accumulates
  Scratch tmp, src;
  PW_TRY(tmp.alloc(rows.N * es, s));
  const void* in = rho;
  if (rho == out) {
    PW_TRY(src.alloc(rows.N * es, s));
    PW_CUDA(cudaMemcpyAsync(src.p, rho, rows.N * es, cudaMemcpyDeviceToDevice, s));
    in = src.p;
  }
  for (int k = 0; k < K; ++k) {
    const char* opk = (const char*)ops + (size_t)k * op_bytes;
    PW_TRY(apply_side(in, tmp.p, opk, dims, n, targets, nt, false, false, dtype, s));
    PW_TRY(apply_side(tmp.p, out, opk, dims, n, targets, nt, true, k > 0, dtype, s));
  }
  return PW_OK;
}
