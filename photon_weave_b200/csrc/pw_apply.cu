// apply_operation / Kraus kernels (rows a1-a6 of SURVEY.md section 8).
//
// The state is never transposed: a fiber (all operator-index positions for one
// setting of the untouched subsystems) is addressed through FiberMap/TargetMap,
// multiplied by the t x t operator and written back to the same positions.
#include "pw_internal.cuh"

namespace pw {

// ---------------------------------------------------------------------------------
// Register path: t <= 8.  One thread owns one fiber: T vector loads (16 B for
// complex128), T*T complex FMAs against the operator broadcast from shared
// memory, T vector stores.  Consecutive threads walk the fastest non-target run,
// so when that run is the trailing one every load/store instruction of a warp
// covers a contiguous 512 B (c128) segment.
// ---------------------------------------------------------------------------------
template <typename V, int T, bool CONJ, bool ACC>
__global__ void __launch_bounds__(256)
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
    V x[T];
#pragma unroll
    for (int j = 0; j < T; ++j) x[j] = in[base + s_off[j]];
#pragma unroll
    for (int a = 0; a < T; ++a) {
      V acc = czero<V>();
      if (ACC) acc = out[base + s_off[a]];
#pragma unroll
      for (int b = 0; b < T; ++b) cfma(acc, s_op[a * T + b], x[b]);
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

extern "C" int pw_apply_vec(const void* psi, void* out, const void* op, const int64_t* dims,
                            int n, const int32_t* targets, int nt, int dtype, void* stream) {
  if (!psi || !out || !op || nt < 1) return PW_ERR_INVALID;
  Plan p;
  PW_TRY(build_plan(dims, n, targets, nt, &p));
  return launch_apply(psi, out, op, p, dtype, false, false, (cudaStream_t)stream);
}

extern "C" int pw_apply_mat(const void* rho, void* out, const void* op, const int64_t* dims,
                            int n, const int32_t* targets, int nt, int dtype, void* stream) {
  if (!rho || !out || !op || nt < 1) return PW_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  Plan rows, cols;
  PW_TRY(build_plan_matrix(dims, n, targets, nt, true, false, &rows));
  PW_TRY(build_plan_matrix(dims, n, targets, nt, false, true, &cols));
  // two passes: left multiply on the row axes, conj(O) on the column axes
  PW_TRY(launch_apply(rho, out, op, rows, dtype, false, false, s));
  return launch_apply(out, out, op, cols, dtype, true, false, s);
}

extern "C" int pw_kraus_mat(const void* rho, void* out, const void* ops, int K,
                            const int64_t* dims, int n, const int32_t* targets, int nt,
                            int dtype, void* stream) {
  if (!rho || !out || !ops || nt < 1 || K < 1) return PW_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  Plan rows, cols;
  PW_TRY(build_plan_matrix(dims, n, targets, nt, true, false, &rows));
  PW_TRY(build_plan_matrix(dims, n, targets, nt, false, true, &cols));
  const size_t es = elem_size(dtype);
  const size_t op_bytes = (size_t)rows.tm.t * rows.tm.t * es;
  if (K == 1) {
    PW_TRY(launch_apply(rho, out, ops, rows, dtype, false, false, s));
    return launch_apply(out, out, ops, cols, dtype, true, false, s);
  }
  // out = sum_k (K_k rho) K_k^dagger : tmp holds K_k rho, second pass accumulates.
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
    PW_TRY(launch_apply(in, tmp.p, opk, rows, dtype, false, false, s));
    // column pass must not alias when accumulating into out
    PW_TRY(launch_apply(tmp.p, out, opk, cols, dtype, true, k > 0, s));
  }
  return PW_OK;
}
