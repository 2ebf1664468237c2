// apply_operation / Kraus kernels (rows a1-a6 of SURVEY.md section 8).
//
// The state is never transposed: a fiber (all operator-index positions for one
// setting of the untouched subsystems) is addressed through FiberMap/TargetMap,
// multiplied by the t x t operator and written back to the same positions.
#include <cstdlib>

#include "pw_internal.cuh"
#include "pw_tiled.cuh"
#include "pw_blocks.cuh"
#include "pw_dense.cuh"

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
              const TargetMap tm, const int* __restrict__ skip) {
  if (skip && *skip) return;
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
    // streaming data, no reuse: L2-only loads avoid L1's sector over-fetch on rows that are
    // not 128 B aligned (measured: 12 % extra DRAM reads with default caching at R = 36)
#pragma unroll
    for (int j = 0; j < T; ++j) x[j] = __ldcg(in + base + s_off[j]);
#pragma unroll
    for (int a = 0; a < T; ++a) {
      V acc = czero<V>();
      if (ACC) acc = out[base + s_off[a]];
#pragma unroll
      for (int b = 0; b < T; ++b) cfma(acc, sop[a * T + b], x[b]);
      __stcg(out + base + s_off[a], acc);
    }
  }
}

// complex64 with an even contiguous untouched run: two adjacent fibers per thread, every access
// 128 bits wide (the complex128 kernel's bytes per thread and per instruction).  `fm` counts
// fiber PAIRS (pair_fiber_map).
template <int T, bool CONJ>
__global__ void __launch_bounds__(256, T <= 4 ? 4 : (T <= 6 ? 3 : 2))
k_apply_small_c64x2(const float2* in, float2* out, const float2* __restrict__ op, const FiberMap fm,
                    const TargetMap tm, const int* __restrict__ skip) {
  if (skip && *skip) return;
  __shared__ float2 s_op[T * T];
  __shared__ int64_t s_off[T];
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    float2 o = op[i];
    if (CONJ) o.y = -o.y;
    s_op[i] = o;
  }
  if (threadIdx.x < T) s_off[threadIdx.x] = tm.offset(threadIdx.x);
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib; f += stride) {
    const int64_t base = fm.base(f);
    int zero = 0;
    asm volatile("" : "+r"(zero));
    const float2* sop = s_op + zero;
    C64x2 x[T];
#pragma unroll
    for (int j = 0; j < T; ++j) x[j] = ld_c64x2(in + base + s_off[j]);
#pragma unroll
    for (int a = 0; a < T; ++a) {
      C64x2 acc = czero<C64x2>();
#pragma unroll
      for (int b = 0; b < T; ++b) cfma(acc, sop[a * T + b], x[b]);
      st_c64x2(out + base + s_off[a], acc);
    }
  }
}

// ---------------------------------------------------------------------------------
// Register path, CONTIGUOUS fibers (the targets are the innermost axes, so the tensor is
// an array of t-element fibers).  Thread-per-fiber loads would touch 32 different lines
// per instruction; instead a warp loads 32 fibers as T fully coalesced 512 B rows,
// transposes them through a warp-private padded shared-memory buffer (no block barrier),
// computes one fiber per lane in registers, and stores back the same way.
// ---------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ int pad_idx(int e) { return e + (e >> 3); }

template <typename V, int T, bool CONJ>
__global__ void __launch_bounds__(256, (sizeof(V) == 16 && T >= 7) ? 2 : 3)
k_apply_small_contig(const V* in, V* out, const V* __restrict__ op, const uint64_t nfib,
                     const TargetMap tm, const int* __restrict__ skip) {
  if (skip && *skip) return;
  constexpr int kWarps = 8;
  constexpr int kBuf = 32 * T + (32 * T) / 8 + 1;
  __shared__ V s_op[T * T];
  __shared__ int s_perm[T];      // operator index j -> position inside the contiguous fiber
  __shared__ V s_buf[kWarps][kBuf];
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    V o = op[i];
    if (CONJ) o.y = -o.y;
    s_op[i] = o;
  }
  if (threadIdx.x < T) s_perm[threadIdx.x] = (int)tm.offset(threadIdx.x);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  V* buf = s_buf[warp];
  const uint64_t ngroups = (nfib + 31) / 32;
  const uint64_t total = nfib * T;
  const uint64_t gstride = (uint64_t)gridDim.x * kWarps;
  for (uint64_t g = (uint64_t)blockIdx.x * kWarps + warp; g < ngroups; g += gstride) {
    const uint64_t base = g * 32 * T;
    V r[T];
#pragma unroll
    for (int i = 0; i < T; ++i) {
      const uint64_t e = base + i * 32 + lane;
      if (e < total) r[i] = __ldcg(in + e);
    }
#pragma unroll
    for (int i = 0; i < T; ++i) buf[pad_idx<T>(i * 32 + lane)] = r[i];
    __syncwarp();
    int zero = 0;
    asm volatile("" : "+r"(zero));
    const V* sop = s_op + zero;
    V x[T];
#pragma unroll
    for (int j = 0; j < T; ++j) x[j] = buf[pad_idx<T>(lane * T + s_perm[j])];
    __syncwarp();
#pragma unroll
    for (int a = 0; a < T; ++a) {
      V acc = czero<V>();
#pragma unroll
      for (int b = 0; b < T; ++b) cfma(acc, sop[a * T + b], x[b]);
      buf[pad_idx<T>(lane * T + s_perm[a])] = acc;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < T; ++i) {
      const uint64_t e = base + i * 32 + lane;
      if (e < total) __stcg(out + e, buf[pad_idx<T>(i * 32 + lane)]);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------
// Fused pair: two single-subsystem operators on DIFFERENT axes in ONE HBM pass.  A thread
// holds the D x D sub-tensor spanned by the two axes in registers, applies U_a along the
// first index and U_b along the second (2*D^3 complex FMAs), and writes it back: half the
// memory traffic of two separate applications, identical arithmetic.
// ---------------------------------------------------------------------------------
template <typename V, int D>
__global__ void __launch_bounds__(256, 1)
k_apply_pair(const V* in, V* out, const V* __restrict__ opa, const V* __restrict__ opb,
             const FiberMap fm, const TargetMap tm) {
  __shared__ V s_a[D * D], s_b[D * D];
  __shared__ int64_t s_off[D * D];
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    s_a[i] = opa[i];
    s_b[i] = opb[i];
    s_off[i] = tm.offset(i);
  }
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib; f += stride) {
    const int64_t base = fm.base(f);
    int zero = 0;
    asm volatile("" : "+r"(zero));
    const V* sa = s_a + zero;
    const V* sb = s_b + zero;
    V x[D][D];
    // first index, column by column: load column j, x[:, j] <- U_a column (keeps the live
    // register set at one D x D tile plus one column)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      V c[D];
#pragma unroll
      for (int b = 0; b < D; ++b) c[b] = __ldcg(in + base + s_off[b * D + j]);
#pragma unroll
      for (int a = 0; a < D; ++a) {
        V acc = czero<V>();
#pragma unroll
        for (int b = 0; b < D; ++b) cfma(acc, sa[a * D + b], c[b]);
        x[a][j] = acc;
      }
    }
    // second index: x[i, :] <- U_b x[i, :], stored as each row completes
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
      for (int a = 0; a < D; ++a) {
        V acc = czero<V>();
#pragma unroll
        for (int b = 0; b < D; ++b) cfma(acc, sb[a * D + b], x[i][b]);
        __stcg(out + base + s_off[i * D + a], acc);
      }
    }
  }
}

// Split variant: Q adjacent lanes share one D x D sub-tensor, each owning D/Q columns of
// the SECOND index.  The first-index product is lane-local; for the second index every row
// is assembled from the Q lanes with warp shuffles.  Keeps the register tile at D*D/Q
// elements (D = 8 complex128 no longer fits one thread) and the loads coalesced.
template <typename V, int D, int Q, bool CONJB>
__global__ void __launch_bounds__(256, 1)
k_apply_pair_split(const V* in, V* out, const V* __restrict__ opa, const V* __restrict__ opb,
                   const FiberMap fm, const TargetMap tm, const int* __restrict__ skip) {
  if (skip && *skip) return;
  constexpr int C = D / Q;
  static_assert(D % Q == 0 && (Q & (Q - 1)) == 0 && Q <= 32, "Q must be a power of two dividing D");
  __shared__ V s_a[D * D], s_b[D * D];
  __shared__ int64_t s_off[D * D];
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    s_a[i] = opa[i];
    V b = opb[i];
    if (CONJB) b.y = -b.y;
    s_b[i] = b;
    s_off[i] = tm.offset(i);
  }
  __syncthreads();
  const int q = threadIdx.x % Q;
  const int lane = threadIdx.x & 31, lane0 = lane - q;
  const uint64_t nwork = fm.nfib * Q;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  // all lanes of a warp must take part in the shuffles: iterate on a warp-uniform bound
  const uint64_t nwork_pad = (nwork + 31) / 32 * 32;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwork_pad; w += stride) {
    const bool live = w < nwork;
    const uint64_t f = live ? w / Q : 0;
    const int64_t base = fm.base(f);
    int zero = 0;
    asm volatile("" : "+r"(zero));
    const V* sa = s_a + zero;
    const V* sb = s_b + zero;
    V x[D][C];
#pragma unroll
    for (int jl = 0; jl < C; ++jl) {
      V c[D];
#pragma unroll
      for (int b = 0; b < D; ++b) c[b] = live ? __ldcg(in + base + s_off[b * D + q * C + jl]) : czero<V>();
#pragma unroll
      for (int a = 0; a < D; ++a) {
        V acc = czero<V>();
#pragma unroll
        for (int b = 0; b < D; ++b) cfma(acc, sa[a * D + b], c[b]);
        x[a][jl] = acc;
      }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      V y[D];
#pragma unroll
      for (int qq = 0; qq < Q; ++qq)
#pragma unroll
        for (int jl = 0; jl < C; ++jl) {
          y[qq * C + jl].x = __shfl_sync(0xffffffffu, x[i][jl].x, lane0 + qq);
          y[qq * C + jl].y = __shfl_sync(0xffffffffu, x[i][jl].y, lane0 + qq);
        }
#pragma unroll
      for (int al = 0; al < C; ++al) {
        V acc = czero<V>();
#pragma unroll
        for (int b = 0; b < D; ++b) cfma(acc, sb[(q * C + al) * D + b], y[b]);
        if (live) __stcg(out + base + s_off[i * D + q * C + al], acc);
      }
    }
  }
}

template <typename V, int D, int Q>
static void launch_pair_split(const V* in, V* out, const V* opa, const V* opb, const Plan& p,
                              bool conj_b, cudaStream_t s, const int* skip) {
  const int grid = grid_for(p.fm.nfib * Q, 256, 1);
  if (conj_b) k_apply_pair_split<V, D, Q, true><<<grid, 256, 0, s>>>(in, out, opa, opb, p.fm, p.tm, skip);
  else        k_apply_pair_split<V, D, Q, false><<<grid, 256, 0, s>>>(in, out, opa, opb, p.fm, p.tm, skip);
}

template <typename V>
static int launch_pair_t(const V* in, V* out, const V* opa, const V* opb, const Plan& p, int d,
                         bool conj_b, cudaStream_t s, const int* skip = nullptr) {
  const int grid = grid_for(p.fm.nfib, 256, 1);
  const bool wide = sizeof(V) == 16;  // complex128: register tiles are twice as large
  switch (d) {
    case 2: launch_pair_split<V, 2, 1>(in, out, opa, opb, p, conj_b, s, skip); break;
    case 3: launch_pair_split<V, 3, 1>(in, out, opa, opb, p, conj_b, s, skip); break;
    case 4: if (wide) launch_pair_split<V, 4, 2>(in, out, opa, opb, p, conj_b, s, skip);
            else      launch_pair_split<V, 4, 1>(in, out, opa, opb, p, conj_b, s, skip);
            break;
    case 5: if (conj_b || skip) return PW_ERR_UNSUPPORTED;
            k_apply_pair<V, 5><<<grid, 256, 0, s>>>(in, out, opa, opb, p.fm, p.tm); break;
    case 6: if (wide) launch_pair_split<V, 6, 2>(in, out, opa, opb, p, conj_b, s, skip);
            else      launch_pair_split<V, 6, 1>(in, out, opa, opb, p, conj_b, s, skip);
            break;
    case 7: if (conj_b || wide) return PW_ERR_UNSUPPORTED;
            return PW_ERR_UNSUPPORTED;
    case 8: if (wide) launch_pair_split<V, 8, 8>(in, out, opa, opb, p, conj_b, s, skip);
            else      launch_pair_split<V, 8, 4>(in, out, opa, opb, p, conj_b, s, skip);
            break;
    default: return PW_ERR_UNSUPPORTED;
  }
  g_paths[0].fetch_add(1, std::memory_order_relaxed);
  PW_LAUNCHED();
  return PW_OK;
}

// ---------------------------------------------------------------------------------
// Generic path: any t.  One thread owns one fiber and re-reads it once per output
// row (L1/L2 hits); operator rows stream from global/L2.  Requires in != out.
// Correctness fallback for shapes the tiled kernels do not cover.
// ---------------------------------------------------------------------------------
template <typename V, bool CONJ, bool ACC>
__global__ void __launch_bounds__(128)
k_apply_generic(const V* __restrict__ in, V* __restrict__ out, const V* __restrict__ op,
                const FiberMap fm, const TargetMap tm, const int64_t* __restrict__ toff,
                const int* __restrict__ skip) {
  if (skip && *skip) return;
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

// ---------------------------------------------------------------------------------
// CONTIGUOUS fibers on the FP64 tensor cores (complex128, one 5..8-level target on the
// innermost axis).  The transposed product Y^T = X^T W^T makes the FIBERS the M dimension of
// mma.sync.m8n8k4: lane (gid, tig) of an A fragment needs X[fiber gid][input tig] -- a plain
// 128-bit global load, the eight lanes of a quarter warp reading two 64-byte pieces -- and a lane
// of the C fragment holds two ADJACENT outputs of one fiber: one 256-bit store.  No shared
// memory, no barrier, 8 loads + 32 DMMAs + 4 stores per lane and 32 fibers; the operator sits in
// two registers per lane.  (The warp-transposing register kernel needs 36 broadcast shared
// loads per fiber and is sensitive to the SM clock: 11.7 ms alone, 14.1 ms inside the
// power-capped bench layer on the 34.8 GB vector.)
// ---------------------------------------------------------------------------------
// S = float2 (complex64): same FP64 tensor-core product, values widened as they are loaded and
// rounded as they are stored (half the bytes per DMMA: the kernel that is HBM-bound at 49 % DMMA
// utilisation in complex128 becomes DMMA-bound near the complex64 HBM roof)
__device__ __forceinline__ double2 cm_wide(const double2 v) { return v; }
__device__ __forceinline__ double2 cm_wide(const float2 v) { return make_double2((double)v.x, (double)v.y); }

template <bool CONJ, typename S>
__global__ void __launch_bounds__(256)
k_apply_contig_mma(const S* __restrict__ in, S* __restrict__ out, const S* __restrict__ op,
                   const int t, const uint64_t nfib, const int* __restrict__ skip) {
  using V = double2;
  if (skip && *skip) return;
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const V zero = make_double2(0.0, 0.0);
  // B = W^T fragments: lane (k = tig, n = gid) holds W[gid][tig] and W[gid][4 + tig]
  V w0 = zero, w1 = zero;
  if (gid < t) {
    if (tig < t) w0 = cm_wide(op[gid * t + tig]);
    if (4 + tig < t) w1 = cm_wide(op[gid * t + 4 + tig]);
    if (CONJ) { w0.y = -w0.y; w1.y = -w1.y; }
  }
  const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  const uint64_t ngroups = (nfib + 31) / 32;
  for (uint64_t g = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < ngroups; g += nwarps) {
    V x0[4], x1[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const uint64_t f = g * 32 + 8 * m + gid;
      const S* src = in + f * t;
      x0[m] = (f < nfib && tig < t) ? cm_wide(__ldcg(src + tig)) : zero;
      x1[m] = (f < nfib && 4 + tig < t) ? cm_wide(__ldcg(src + 4 + tig)) : zero;
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      double re0 = 0.0, re1 = 0.0, im0 = 0.0, im1 = 0.0;
      asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(re0), "+d"(re1) : "d"(x0[m].x), "d"(w0.x));
      asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(im0), "+d"(im1) : "d"(x0[m].x), "d"(w0.y));
      asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(re0), "+d"(re1) : "d"(-x0[m].y), "d"(w0.y));
      asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(im0), "+d"(im1) : "d"(x0[m].y), "d"(w0.x));
      asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(re0), "+d"(re1) : "d"(x1[m].x), "d"(w1.x));
      asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(im0), "+d"(im1) : "d"(x1[m].x), "d"(w1.y));
      asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(re0), "+d"(re1) : "d"(-x1[m].y), "d"(w1.y));
      asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(im0), "+d"(im1) : "d"(x1[m].y), "d"(w1.x));
      const uint64_t f = g * 32 + 8 * m + gid;
      S* dst = out + f * t + 2 * tig;
      const bool ok0 = f < nfib && 2 * tig < t, ok1 = f < nfib && 2 * tig + 1 < t;
      if constexpr (sizeof(S) == 16) {
        if (ok0 && ok1 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
          asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "d"(re0), "d"(im0), "d"(re1), "d"(im1) : "memory");
        } else {
          if (ok0) __stcg((double2*)dst, make_double2(re0, im0));
          if (ok1) __stcg((double2*)dst + 1, make_double2(re1, im1));
        }
      } else {
        if (ok0 && ok1 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
          __stcg(reinterpret_cast<float4*>(dst), make_float4((float)re0, (float)im0, (float)re1, (float)im1));
        } else {
          if (ok0) __stcg((float2*)dst, make_float2((float)re0, (float)im0));
          if (ok1) __stcg((float2*)dst + 1, make_float2((float)re1, (float)im1));
        }
      }
    }
  }
}

template <typename V, int T>
static int launch_small(const V* in, V* out, const V* op, const Plan& p, bool conj_op,
                        bool acc, cudaStream_t s, const int* skip) {
  const int block = 256;
  const int grid = grid_for(p.fm.nfib, block, 8);
  g_paths[0].fetch_add(1, std::memory_order_relaxed);
  // array-of-contiguous-fibers layout: the warp-transposing variant keeps loads coalesced
  if (!acc && T >= 2 && p.fm.ng == 1 && p.fm.gstride[0] == (int64_t)T && p.N == p.fm.nfib * T &&
      p.fm.nfib >= 4096 && options().contig_kernel.load(std::memory_order_relaxed)) {
    // (complex64 through this kernel measured SLOWER than the warp-transposing register kernel:
    //  8.46 against 7.85 ms on the 17.4 GB vector; it stays available as contig_mma = 2)
    if (T >= 5 && p.tm.nt == 1 && in != out && options().contig_mma.load(std::memory_order_relaxed) &&
        (sizeof(V) == 16 || options().contig_mma.load(std::memory_order_relaxed) >= 2)) {
      const int g3 = grid_for((p.fm.nfib + 31) / 32 * 32, block, 8);
      if (conj_op) k_apply_contig_mma<true, V><<<g3, block, 0, s>>>(in, out, op, T, p.fm.nfib, skip);
      else         k_apply_contig_mma<false, V><<<g3, block, 0, s>>>(in, out, op, T, p.fm.nfib, skip);
      PW_LAUNCHED();
      return PW_OK;
    }
    const int g2 = grid_for((p.fm.nfib + 31) / 32 * 32, block, 3);
    if (conj_op) k_apply_small_contig<V, T, true><<<g2, block, 0, s>>>(in, out, op, p.fm.nfib, p.tm, skip);
    else         k_apply_small_contig<V, T, false><<<g2, block, 0, s>>>(in, out, op, p.fm.nfib, p.tm, skip);
    PW_LAUNCHED();
    return PW_OK;
  }
  FiberMap pm;
  if (sizeof(V) == 8 && !acc && T >= 2 && options().c64_pairs.load(std::memory_order_relaxed) &&
      pair_fiber_map(p.fm, in, out, &pm)) {
    const int gp = grid_for(pm.nfib, block, 8);
    if (conj_op) k_apply_small_c64x2<T, true><<<gp, block, 0, s>>>((const float2*)in, (float2*)out, (const float2*)op, pm, p.tm, skip);
    else         k_apply_small_c64x2<T, false><<<gp, block, 0, s>>>((const float2*)in, (float2*)out, (const float2*)op, pm, p.tm, skip);
    PW_LAUNCHED();
    return PW_OK;
  }
  if (conj_op) {
    if (acc) k_apply_small<V, T, true, true><<<grid, block, 0, s>>>(in, out, op, p.fm, p.tm, skip);
    else     k_apply_small<V, T, true, false><<<grid, block, 0, s>>>(in, out, op, p.fm, p.tm, skip);
  } else {
    if (acc) k_apply_small<V, T, false, true><<<grid, block, 0, s>>>(in, out, op, p.fm, p.tm, skip);
    else     k_apply_small<V, T, false, false><<<grid, block, 0, s>>>(in, out, op, p.fm, p.tm, skip);
  }
  PW_LAUNCHED();
  return PW_OK;
}

template <typename V>
static int launch_apply_t(const V* in, V* out, const V* op, const Plan& p, bool conj_op,
                          bool acc, cudaStream_t s, const int* skip) {
  switch (p.tm.t) {
    case 1: return launch_small<V, 1>(in, out, op, p, conj_op, acc, s, skip);
    case 2: return launch_small<V, 2>(in, out, op, p, conj_op, acc, s, skip);
    case 3: return launch_small<V, 3>(in, out, op, p, conj_op, acc, s, skip);
    case 4: return launch_small<V, 4>(in, out, op, p, conj_op, acc, s, skip);
    case 5: return launch_small<V, 5>(in, out, op, p, conj_op, acc, s, skip);
    case 6: return launch_small<V, 6>(in, out, op, p, conj_op, acc, s, skip);
    case 7: return launch_small<V, 7>(in, out, op, p, conj_op, acc, s, skip);
    case 8: return launch_small<V, 8>(in, out, op, p, conj_op, acc, s, skip);
    default: break;
  }
  g_paths[1].fetch_add(1, std::memory_order_relaxed);
  // generic: needs a non-aliased source and the offset table in global memory
  Scratch toff, tmp;
  PW_TRY(toff.alloc(sizeof(int64_t) * p.tm.t, s));
  if (!plan_replay()) {
    k_target_offsets<<<(p.tm.t + 127) / 128, 128, 0, s>>>(p.tm, (int64_t*)toff.p);
    PW_LAUNCHED();
  }
  const V* src = in;
  if ((const void*)in == (const void*)out) {
    PW_TRY(tmp.alloc_temp(sizeof(V) * p.N, s));
    PW_CUDA(cudaMemcpyAsync(tmp.p, in, sizeof(V) * p.N, cudaMemcpyDeviceToDevice, s));
    src = (const V*)tmp.p;
  }
  const int block = 128;
  const int grid = grid_for(p.fm.nfib, block, 8);
  if (conj_op) {
    if (acc) k_apply_generic<V, true, true><<<grid, block, 0, s>>>(src, out, op, p.fm, p.tm, (int64_t*)toff.p, skip);
    else     k_apply_generic<V, true, false><<<grid, block, 0, s>>>(src, out, op, p.fm, p.tm, (int64_t*)toff.p, skip);
  } else {
    if (acc) k_apply_generic<V, false, true><<<grid, block, 0, s>>>(src, out, op, p.fm, p.tm, (int64_t*)toff.p, skip);
    else     k_apply_generic<V, false, false><<<grid, block, 0, s>>>(src, out, op, p.fm, p.tm, (int64_t*)toff.p, skip);
  }
  PW_LAUNCHED();
  return PW_OK;
}

int launch_apply(const void* in, void* out, const void* op, const Plan& plan, int dtype,
                 bool conj_op, bool accumulate, cudaStream_t stream, const int* skip) {
  if (dtype == PW_C128)
    return launch_apply_t<double2>((const double2*)in, (double2*)out, (const double2*)op, plan,
                                   conj_op, accumulate, stream, skip);
  if (dtype == PW_C64)
    return launch_apply_t<float2>((const float2*)in, (float2*)out, (const float2*)op, plan,
                                  conj_op, accumulate, stream, skip);
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
// states below this many elements are launch-latency bound: skip structure analysis
static uint64_t blocks_min_elems() {
  return (uint64_t)options().blocks_min_elems.load(std::memory_order_relaxed);
}
// A SINGLE operator with t <= 8 on a density matrix: the structured one-pass paths cost ~50 us of
// analysis kernels, the fallback is two register-kernel passes -- the analysis pays off only
// when a pass takes longer than that (measured on the 6.25 MB rho of the lossy-circuit config:
// dense 5 x 5 69 us with analysis, 14 us without).
static uint64_t matrix_min_elems(int K, uint64_t t) {
  const uint64_t base = blocks_min_elems();
  if (K != 1 || t > 8 || base == 0) return base;
  const uint64_t small_t = (uint64_t)options().blocks_min_elems_small_t.load(std::memory_order_relaxed);
  return small_t > base ? small_t : base;
}

// One strided apply on an arbitrary axis list: out = (O or conj(O) on targets) in.
//   t <= 8                      register kernel
//   t <= 128, large, in != out  block-structured attempt, then a fallback that exits on
//                               the device when the block kernel took the operator
//   fallback                    TMA-tiled CSR kernel if eligible, else the generic kernel
// ---------------------------------------------------------------------------------
// Diagonal operators (phase shifts, Kerr phases, ...): out[i] = O[j, j] in[i] with j read off the
// digits of i -- a pure streaming pass, fully coalesced WHATEVER the target axes are (the
// register kernel walks fibers and loses 5-20 % on short untouched inner runs).  The operator
// lives on the device, so a one-warp kernel tests it and the streaming kernel exits when it is
// not diagonal; the regular path then runs with the same flag as its skip flag.
// ---------------------------------------------------------------------------------
template <typename V>
__global__ void k_diag_check(const V* __restrict__ op, const uint32_t t, int* __restrict__ flag) {
  int ok = 1;
  for (uint32_t e = threadIdx.x; e < t * t; e += 32) {
    const V v = op[e];
    if (e / t != e % t && (v.x != 0 || v.y != 0)) ok = 0;
  }
  ok = __all_sync(0xffffffffu, ok);
  if (threadIdx.x == 0) *flag = ok;
}

// two adjacent elements with ONE access: 256-bit for complex128 (sm_100 LDG/STG.E.256), 128-bit
// for complex64 (two separate 16-byte accesses would fetch every sector twice from L2)
__device__ __forceinline__ void ld_pair(const double2* p, double2& a, double2& b) {
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}
__device__ __forceinline__ void st_pair2(double2* p, const double2 a, const double2 b) {
  asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}
__device__ __forceinline__ void ld_pair(const float2* p, float2& a, float2& b) {
  const float4 v = __ldcg(reinterpret_cast<const float4*>(p));
  a = make_float2(v.x, v.y);
  b = make_float2(v.z, v.w);
}
__device__ __forceinline__ void st_pair2(float2* p, const float2 a, const float2 b) {
  __stcg(reinterpret_cast<float4*>(p), make_float4(a.x, a.y, b.x, b.y));
}

// 32-bit indices (N < 2^32) keep the digit arithmetic cheap; PAIR: every target stride is even, so
// elements 2 i and 2 i + 1 share their digits and a thread moves them with ONE 2 x 16-byte access
// pair (complex128: 256-bit).
template <typename V, bool CONJ, bool PAIR>
__global__ void __launch_bounds__(256)
k_apply_diag(const V* __restrict__ in, V* __restrict__ out, const V* __restrict__ op, const uint32_t N,
             const TargetMap tm, const int* __restrict__ flag) {
  if (!*flag) return;
  __shared__ V s_d[8];
  if (threadIdx.x < tm.t) {
    V v = op[threadIdx.x * (tm.t + 1)];
    if (CONJ) v.y = -v.y;
    s_d[threadIdx.x] = v;
  }
  __syncthreads();
  const int nt = tm.nt;
  const uint32_t ts0 = (uint32_t)tm.tstride[0], td0 = tm.tdim[0];  // (one target: registers only)
  const uint32_t step = PAIR ? 2 : 1;
  const uint32_t stride = gridDim.x * blockDim.x * step;
  // (i + stride may wrap for N close to 2^32: the loop condition checks the wrap)
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * step; i < N; ) {
    uint32_t j = (i / ts0) % td0;
#pragma unroll 1
    for (int k = 1; k < nt; ++k) j = j * tm.tdim[k] + (i / (uint32_t)tm.tstride[k]) % tm.tdim[k];
    const V dgl = s_d[j];
    if (PAIR) {
      V x0, x1, y0, y1;
      ld_pair(in + i, x0, x1);
      y0.x = dgl.x * x0.x - dgl.y * x0.y; y0.y = dgl.x * x0.y + dgl.y * x0.x;
      y1.x = dgl.x * x1.x - dgl.y * x1.y; y1.y = dgl.x * x1.y + dgl.y * x1.x;
      st_pair2(out + i, y0, y1);
    } else {
      const V x = __ldcg(in + i);
      V y;
      y.x = dgl.x * x.x - dgl.y * x.y; y.y = dgl.x * x.y + dgl.y * x.x;
      __stcg(out + i, y);
    }
    const uint32_t nxt = i + stride;
    if (nxt < i) break;
    i = nxt;
  }
}

template <typename V>
static int try_diag(const V* in, V* out, const V* op, const Plan& p, bool conj_op, cudaStream_t s,
                    Scratch* flag_mem, const int** flag) {
  *flag = nullptr;
  if (p.N >= (1ull << 32) || p.N < (1ull << 22) || p.tm.t > 8) return PW_OK;
  TargetMap tm = p.tm;
  bool even = (p.N & 1) == 0 && (((uintptr_t)in | (uintptr_t)out) & (2 * sizeof(V) - 1)) == 0;
  for (int k = 0; k < tm.nt; ++k) {
    if (tm.tdim[k] == 1) { tm.tstride[k] = 1; continue; }  // (singleton targets: digit 0)
    if (tm.tstride[k] <= 0) return PW_OK;
    if (tm.tstride[k] & 1) even = false;
  }
  PW_TRY(flag_mem->alloc(sizeof(int), s));
  if (!plan_replay()) {
    k_diag_check<V><<<1, 32, 0, s>>>(op, p.tm.t, (int*)flag_mem->p);
    PW_LAUNCHED();
  }
  const int* fl = (const int*)flag_mem->p;
  const int grid = grid_for(even ? p.N / 2 : p.N, 256, 8);
  if (even) {
    if (conj_op) k_apply_diag<V, true, true><<<grid, 256, 0, s>>>(in, out, op, (uint32_t)p.N, tm, fl);
    else         k_apply_diag<V, false, true><<<grid, 256, 0, s>>>(in, out, op, (uint32_t)p.N, tm, fl);
  } else {
    if (conj_op) k_apply_diag<V, true, false><<<grid, 256, 0, s>>>(in, out, op, (uint32_t)p.N, tm, fl);
    else         k_apply_diag<V, false, false><<<grid, 256, 0, s>>>(in, out, op, (uint32_t)p.N, tm, fl);
  }
  PW_LAUNCHED();
  *flag = fl;
  return PW_OK;
}

// size of the innermost run of untouched axes if it is the contiguous (stride-1) one,
// else 0 (the innermost axis of the tensor is a target)
static uint64_t inner_rest_run(const Plan& p) {
  if (p.fm.ng == 0) return 0;
  return p.fm.gstride[p.fm.ng - 1] == 1 ? p.fm.gsize[p.fm.ng - 1] : 0;
}

static int apply_axes(const void* in, void* out, const void* op, const int64_t* axes, int naxes,
                      const int* targets, int nt, bool conj_op, int dtype, cudaStream_t s) {
  Plan p;
  PW_TRY(build_plan_axes(axes, naxes, targets, nt, &p));
  const int vp = vec_path();
  const uint64_t r_in = inner_rest_run(p);
  const uint64_t small_r = (uint64_t)options().small_r.load(std::memory_order_relaxed);
  const bool large = p.N >= blocks_min_elems();
  // Measured on B200 (profiles/r1_kbench_*): the register kernel wins for every t <= 8
  // layout; for t > 8 the direct block kernel wins when the contiguous direction is an
  // untouched run (>= small_r elements), TMA-staged tiles win when it belongs to the fibers.
  const bool prefer_tiled = vp == 2 || (vp == 0 && large && r_in < small_r && p.tm.t > 8);
  if (p.tm.t <= 8 && vp != 3 && !prefer_tiled) {
    // 5..8-level complex128 operator on the innermost axis (contiguous fibers): the staged
    // tensor-core kernel (HBM-bound) first, the warp-transposing register kernel as fallback
    BlockDesc bd8;
    const int* skip8 = nullptr;
    // diagonal operators on large states: one streaming pass (decided on the device)
    Scratch diag_flag;
    // (only where the register kernel is weak: a short untouched inner run -- measured 11.3 ms
    //  for every axis of the 34.8 GB vector against 11.4 / 13.1 / 13.4 ms at R >= 1296 / 216 / 36;
    //  contiguous fibers, R = 0, have the tensor-core kernel at 10.7 ms)
    if (vp == 0 && large && p.tm.t >= 2 && r_in > 0 && r_in < 1024 &&
        options().diag_kernel.load(std::memory_order_relaxed)) {
      if (dtype == PW_C128)
        PW_TRY(try_diag<double2>((const double2*)in, (double2*)out, (const double2*)op, p, conj_op, s, &diag_flag, &skip8));
      else
        PW_TRY(try_diag<float2>((const float2*)in, (float2*)out, (const float2*)op, p, conj_op, s, &diag_flag, &skip8));
      if (skip8) return launch_apply(in, out, op, p, dtype, conj_op, false, s, skip8);
    }
    if (vp == 0 && dtype == PW_C128 && p.tm.t >= 5 && large && in != out && r_in == 0 &&
        options().seg_mma.load(std::memory_order_relaxed) > 1) {
      const int rc = launch_blocks_seg(in, out, op, axes, naxes, targets, nt, dtype, conj_op, &bd8, s, true);
      if (rc == PW_OK) skip8 = bd8.skip_flag();
      else if (rc != PW_ERR_UNSUPPORTED) return rc;
    }
    return launch_apply(in, out, op, p, dtype, conj_op, false, s, skip8);
  }
  // launch-latency-bound states: one row-parallel dense kernel, no analysis, no tiles
  if (vp == 0 && !large && in != out) {
    const int rc = launch_apply_rows(in, out, op, p, dtype, conj_op, s);
    if (rc != PW_ERR_UNSUPPORTED) return rc;
  }
  BlockDesc bd;
  const int* skip = nullptr;
  if ((vp == 0 || vp == 3) && !prefer_tiled && in != out && p.tm.t <= kBlockMaxT && large) {
    const int rc = launch_blocks(in, out, op, p, dtype, conj_op, &bd, s);
    if (rc == PW_OK) skip = bd.skip_flag();
    else if (rc != PW_ERR_UNSUPPORTED) return rc;
  } else if (vp == 0 && prefer_tiled && in != out && r_in == 0 &&
             options().seg_kernel.load(std::memory_order_relaxed)) {
    // the contiguous direction belongs to the fibers: warp-staged block kernel first
    const int rc = launch_blocks_seg(in, out, op, axes, naxes, targets, nt, dtype, conj_op, &bd, s);
    if (rc == PW_OK) skip = bd.skip_flag();
    else if (rc != PW_ERR_UNSUPPORTED) return rc;
  }
  // dense (or not usefully structured) complex128 operators up to 64 x 64 with adjacent fibers:
  // FP64 tensor cores
  if (vp == 0 && large && options().dense_mma.load(std::memory_order_relaxed)) {
    const int rc = launch_dense_mma(in, out, op, p, dtype, conj_op, s, skip);
    if (rc != PW_ERR_UNSUPPORTED) return rc;
  }
  if (vp != 1) {
    TiledGeom g;
    if (plan_tiled(axes, naxes, targets, nt, nullptr, 0, -1, elem_size(dtype), tile_elems(dtype), &g)) {
      const int rc = launch_tiled(in, out, op, 1, g, conj_op ? kModeVectorConj : kModeVector, dtype, s, skip);
      if (rc != PW_ERR_UNSUPPORTED) return rc;
    }
    g_paths[3].fetch_add(1, std::memory_order_relaxed);
  }
  return launch_apply(in, out, op, p, dtype, conj_op, false, s, skip);
}

extern "C" int pw_apply_vec(const void* psi, void* out, const void* op, const int64_t* dims,
                            int n, const int32_t* targets, int nt, int dtype, void* stream) {
  if (!psi || !out || !op || nt < 1 || !dims || !targets) return PW_ERR_INVALID;
  if (dtype != PW_C64 && dtype != PW_C128) return PW_ERR_INVALID;
  if (n > PW_MAX_SUBSYSTEMS || nt > PW_MAX_TARGETS) return PW_ERR_INVALID;
  PlanScope scope(op, 0, dtype, 1, dims, n, targets, nt, nullptr, 0, psi == out, (cudaStream_t)stream);
  return scope.finish(apply_axes(psi, out, op, dims, n, targets, nt, false, dtype, (cudaStream_t)stream));
}

// A (possibly rectangular) block of a density matrix as a tensor over `naxes` axes with the
// operator acting on `rows` (as O) and on `cols` (as conj(O)).  pw_apply_mat / pw_kraus_mat
// use axes = dims ++ dims; the multi-GPU layer passes a rank's row slab (pw_kraus_axes).
struct MatrixReq {
  int64_t axes[kMaxAxes];
  int rows[kMaxTargets], cols[kMaxTargets], both[kMaxTargets];
  int naxes, ntr, ntc;
  uint64_t t, total;  // operator dimension; elements of the block
};

static int matrix_req_axes(const int64_t* axes, int naxes, const int32_t* rows, int ntr,
                           const int32_t* cols, int ntc, MatrixReq* r) {
  if (!axes || !rows || !cols || naxes < 2 || naxes > kMaxAxes || ntr < 1 || ntc < 1 ||
      ntr + ntc > kMaxTargets)
    return PW_ERR_INVALID;
  r->naxes = naxes; r->ntr = ntr; r->ntc = ntc; r->t = 1; r->total = 1;
  for (int i = 0; i < naxes; ++i) {
    if (axes[i] < 1) return PW_ERR_INVALID;
    r->axes[i] = axes[i];
    if (r->total > (1ull << 62) / (uint64_t)axes[i]) return PW_ERR_INVALID;
    r->total *= (uint64_t)axes[i];
  }
  uint64_t tc = 1;
  for (int k = 0; k < ntr; ++k) {
    if (rows[k] < 0 || rows[k] >= naxes) return PW_ERR_INVALID;
    r->rows[k] = r->both[k] = rows[k];
    r->t *= (uint64_t)axes[rows[k]];
  }
  for (int k = 0; k < ntc; ++k) {
    if (cols[k] < 0 || cols[k] >= naxes) return PW_ERR_INVALID;
    r->cols[k] = r->both[ntr + k] = cols[k];
    tc *= (uint64_t)axes[cols[k]];
  }
  if (tc != r->t) return PW_ERR_INVALID;  // the same operator acts on both sides
  return PW_OK;
}

static int matrix_req(const int64_t* dims, int n, const int32_t* targets, int nt, MatrixReq* r) {
  if (!dims || !targets || n < 1 || n > PW_MAX_SUBSYSTEMS || nt < 1 || nt > PW_MAX_TARGETS)
    return PW_ERR_INVALID;
  int64_t axes[kMaxAxes];
  int32_t rows[PW_MAX_TARGETS], cols[PW_MAX_TARGETS];
  for (int i = 0; i < n; ++i) axes[i] = axes[n + i] = dims[i];
  for (int k = 0; k < nt; ++k) {
    if (targets[k] < 0 || targets[k] >= n) return PW_ERR_INVALID;
    rows[k] = targets[k];
    cols[k] = n + targets[k];
  }
  return matrix_req_axes(axes, 2 * n, rows, nt, cols, nt, r);
}

// S = sum_k K_k (x) conj(K_k) on the device, then the block-structured kernel on the
// (row targets, column targets) of the doubled tensor: ONE HBM pass whatever K is.
static int try_super_blocks(const void* rho, void* out, const void* ops, int K, const MatrixReq& r,
                            int dtype, cudaStream_t s, Scratch* super, BlockDesc* bd) {
  if (rho == out || r.t * r.t > kBlockMaxT || r.total < matrix_min_elems(K, r.t)) return PW_ERR_UNSUPPORTED;
  Plan p;
  PW_TRY(build_plan_axes(r.axes, r.naxes, r.both, r.ntr + r.ntc, &p));
  const uint64_t r_in = inner_rest_run(p);
  const bool staged = mat_mode() == 0 && r_in < (uint64_t)options().small_r.load(std::memory_order_relaxed);
  if (staged && (r_in != 0 || !options().seg_kernel.load(std::memory_order_relaxed)))
    return PW_ERR_UNSUPPORTED;  // short untouched inner run: the tiled kernel stages it
  const size_t es = elem_size(dtype);
  const uint64_t tt = r.t * r.t;
  PW_TRY(super->alloc(es * tt * tt, s));
  PW_TRY(build_superoperator(ops, K, (uint32_t)r.t, super->p, dtype, s));
  if (staged)  // the column target is the innermost axis: warp-staged block kernel
    return launch_blocks_seg(rho, out, super->p, r.axes, r.naxes, r.both, r.ntr + r.ntc, dtype, false, bd, s);
  return launch_blocks(rho, out, super->p, p, dtype, false, bd, s);
}

// one side of rho' = O rho O^dagger as a strided apply on the doubled tensor
static int apply_side(const void* in, void* out, const void* op, const MatrixReq& r, bool cols,
                      bool accumulate, int dtype, cudaStream_t s, const int* skip) {
  const int* tg = cols ? r.cols : r.rows;
  if (r.t <= 8 || skip || accumulate) {
    Plan p;
    PW_TRY(build_plan_axes(r.axes, r.naxes, tg, cols ? r.ntc : r.ntr, &p));
    return launch_apply(in, out, op, p, dtype, cols, accumulate, s, skip);
  }
  return apply_axes(in, out, op, r.axes, r.naxes, tg, cols ? r.ntc : r.ntr, cols, dtype, s);
}

// fallback: two HBM passes per operator (row side, then conj on the column side)
static int two_pass(const void* rho, void* out, const void* ops, int K, const MatrixReq& r,
                    int dtype, cudaStream_t s, const int* skip) {
  const size_t es = elem_size(dtype);
  if (K == 1) {
    if (r.t > 8 && rho != out) {
      // large operators: keep BOTH sides out of place (rho -> tmp -> out) so each side can
      // take the block-structured kernels (a beam splitter on rho: 2 x ~90 % passes)
      Scratch tmp2;
      if (tmp2.alloc_temp((size_t)r.total * es, s) == PW_OK) {
        // (with a skip flag the sides run the skippable kernels; in place would cost an
        // unconditional copy of rho)
        PW_TRY(apply_side(rho, tmp2.p, ops, r, false, false, dtype, s, skip));
        return apply_side(tmp2.p, out, ops, r, true, false, dtype, s, skip);
      }
    }
    PW_TRY(apply_side(rho, out, ops, r, false, false, dtype, s, skip));
    return apply_side(out, out, ops, r, true, false, dtype, s, skip);
  }
  // out = sum_k (K_k rho) K_k^dagger; tmp holds K_k rho, the column pass accumulates
  const size_t bytes = (size_t)r.total * es, op_bytes = (size_t)(r.t * r.t) * es;
  Scratch tmp, src;
  PW_TRY(tmp.alloc_temp(bytes, s));
  const void* in = rho;
  if (rho == out) {
    PW_TRY(src.alloc_temp(bytes, s));
    PW_CUDA(cudaMemcpyAsync(src.p, rho, bytes, cudaMemcpyDeviceToDevice, s));
    in = src.p;
  }
  for (int k = 0; k < K; ++k) {
    const char* opk = (const char*)ops + (size_t)k * op_bytes;
    PW_TRY(apply_side(in, tmp.p, opk, r, false, false, dtype, s, skip));
    PW_TRY(apply_side(tmp.p, out, opk, r, true, k > 0, dtype, s, skip));
  }
  return PW_OK;
}

static int matrix_dispatch(const void* rho, void* out, const void* ops, int K, const MatrixReq& r,
                           int dtype, cudaStream_t s) {
  const int mm = mat_mode();
  const size_t es = elem_size(dtype);
  Scratch super;
  BlockDesc bd;
  const int* skip = nullptr;
  if (mm == 0 || mm == 4) {
    const int rc = try_super_blocks(rho, out, ops, K, r, dtype, s, &super, &bd);
    if (rc == PW_OK) skip = bd.skip_flag();
    else if (rc != PW_ERR_UNSUPPORTED) return rc;
    // a Kraus sum whose superoperator is dense (no usable block structure): the t^2 x t^2
    // superoperator is ONE dense one-sided operator on the doubled tensor -> FP64 tensor cores
    // (one pass whatever K is; K row + column passes otherwise)
    if (rc == PW_OK && K > 1 && mm == 0 && !bd.staged && r.t * r.t <= 64 &&
        options().dense_mma.load(std::memory_order_relaxed)) {
      Plan pb;
      PW_TRY(build_plan_axes(r.axes, r.naxes, r.both, r.ntr + r.ntc, &pb));
      const int rc2 = launch_dense_mma(rho, out, super.p, pb, dtype, false, s, skip);
      if (rc2 != PW_ERR_UNSUPPORTED) return rc2;
    }
  }
  // single operator, untouched innermost axis: ONE pass through shared-memory tiles
  // (dense d <= 8 operators, beam-splitter-like block structure); decided on the device
  TwoSided ts;
  if (K == 1 && mm == 0 && rho != out && r.t >= 2 && r.t <= kBlockMaxT && r.total >= matrix_min_elems(K, r.t) &&
      options().two_sided.load(std::memory_order_relaxed)) {
    Plan pb, pr, pc;
    PW_TRY(build_plan_axes(r.axes, r.naxes, r.both, r.ntr + r.ntc, &pb));
    // (innermost axis a target: launch_two_sided takes the geometries it has a kernel for)
    const uint64_t r_in = inner_rest_run(pb);
    if (r_in >= 4 || r_in == 0) {
      PW_TRY(build_plan_axes(r.axes, r.naxes, r.rows, r.ntr, &pr));
      PW_TRY(build_plan_axes(r.axes, r.naxes, r.cols, r.ntc, &pc));
      const int rc = launch_two_sided(rho, out, ops, pb.fm, pr.tm, pc.tm, dtype, skip, &ts, s);
      if (rc == PW_OK) skip = ts.skip;
      else if (rc != PW_ERR_UNSUPPORTED) return rc;
    }
  }
  // single-pass tiled kernel (operator pairs applied on the resident tile)
  // single operators: two register-kernel passes beat the tiled two-sided kernel (measured)
  const bool tiled_ok = mm == 2 || mm == 3 || (mm == 0 && (K > 1 || r.t > 8));
  if (tiled_ok) {
    // superoperator pass when it is cheap enough: Kraus sums always (t^2 <= 256), single
    // operators when the structure analysis may find small blocks (t^2 <= 128)
    bool super_mode = r.t * r.t <= 256 && ((K > 1 && (uint64_t)K * 2 >= r.t) || (K == 1 && skip == nullptr && r.t * r.t <= kBlockMaxT && mm == 0 && options().mat_super_single.load(std::memory_order_relaxed)));
    if (mm == 2) super_mode = false;
    if (mm == 3) super_mode = r.t * r.t <= 1024;
    TiledGeom g;
    int rc = PW_ERR_UNSUPPORTED;
    if (super_mode) {
      if (plan_tiled(r.axes, r.naxes, r.both, r.ntr + r.ntc, nullptr, 0, -1, es, tile_elems(dtype), &g))
        rc = launch_tiled(rho, out, ops, K, g, kModeMatrixSuper, dtype, s, skip);
    } else if (plan_tiled(r.axes, r.naxes, r.rows, r.ntr, r.cols, r.ntc, -1, es, tile_elems(dtype), &g)) {
      rc = launch_tiled(rho, out, ops, K, g, kModeMatrixTwoSided, dtype, s, skip);
    }
    if (rc != PW_ERR_UNSUPPORTED) return rc;
    g_paths[3].fetch_add(1, std::memory_order_relaxed);
  }
  // single operator on ONE subsystem: rows and columns are two commuting single-axis
  // operators (O, conj(O)) on the doubled tensor -> one fused register pass
  // (measured: wins for d <= 4; for larger d the register tile limits occupancy and two
  //  register-kernel passes are faster -- profiles/r1_kbench_mat_v6.txt)
  const uint64_t pair_max_d = (uint64_t)options().mat_pair_max_d.load(std::memory_order_relaxed);
  if (K == 1 && r.ntr == 1 && r.ntc == 1 && rho != out && mm != 1 && r.t <= pair_max_d) {
    const int tg[2] = {r.rows[0], r.cols[0]};
    Plan p;
    PW_TRY(build_plan_axes(r.axes, r.naxes, tg, 2, &p));
    int rc;
    if (dtype == PW_C128)
      rc = launch_pair_t<double2>((const double2*)rho, (double2*)out, (const double2*)ops,
                                  (const double2*)ops, p, (int)r.t, true, s, skip);
    else
      rc = launch_pair_t<float2>((const float2*)rho, (float2*)out, (const float2*)ops,
                                 (const float2*)ops, p, (int)r.t, true, s, skip);
    if (rc != PW_ERR_UNSUPPORTED) return rc;
  }
  return two_pass(rho, out, ops, K, r, dtype, s, skip);
}

extern "C" int pw_apply_mat(const void* rho, void* out, const void* op, const int64_t* dims,
                            int n, const int32_t* targets, int nt, int dtype, void* stream) {
  if (!rho || !out || !op) return PW_ERR_INVALID;
  if (dtype != PW_C64 && dtype != PW_C128) return PW_ERR_INVALID;
  MatrixReq r;
  PW_TRY(matrix_req(dims, n, targets, nt, &r));
  PlanScope scope(op, 1, dtype, 1, dims, n, targets, nt, nullptr, 0, rho == out, (cudaStream_t)stream);
  return scope.finish(matrix_dispatch(rho, out, op, 1, r, dtype, (cudaStream_t)stream));
}

extern "C" int pw_kraus_mat(const void* rho, void* out, const void* ops, int K,
                            const int64_t* dims, int n, const int32_t* targets, int nt,
                            int dtype, void* stream) {
  if (!rho || !out || !ops || K < 1) return PW_ERR_INVALID;
  if (dtype != PW_C64 && dtype != PW_C128) return PW_ERR_INVALID;
  MatrixReq r;
  PW_TRY(matrix_req(dims, n, targets, nt, &r));
  PlanScope scope(ops, 2, dtype, K, dims, n, targets, nt, nullptr, 0, rho == out, (cudaStream_t)stream);
  return scope.finish(matrix_dispatch(rho, out, ops, K, r, dtype, (cudaStream_t)stream));
}

extern "C" int pw_kraus_axes(const void* x, void* out, const void* ops, int K, const int64_t* axes,
                             int naxes, const int32_t* row_targets, int ntr,
                             const int32_t* col_targets, int ntc, int dtype, void* stream) {
  if (!x || !out || !ops || K < 1) return PW_ERR_INVALID;
  if (dtype != PW_C64 && dtype != PW_C128) return PW_ERR_INVALID;
  MatrixReq r;
  PW_TRY(matrix_req_axes(axes, naxes, row_targets, ntr, col_targets, ntc, &r));
  PlanScope scope(ops, 3, dtype, K, axes, naxes, row_targets, ntr, col_targets, ntc, x == out, (cudaStream_t)stream);
  return scope.finish(matrix_dispatch(x, out, ops, K, r, dtype, (cudaStream_t)stream));
}

extern "C" int pw_apply_vec_pair(const void* psi, void* out, const void* op_a, const void* op_b,
                                 const int64_t* dims, int n, int axis_a, int axis_b, int dtype,
                                 void* stream) {
  if (!psi || !out || !op_a || !op_b || !dims || n < 2 || n > PW_MAX_SUBSYSTEMS) return PW_ERR_INVALID;
  if (axis_a < 0 || axis_b < 0 || axis_a >= n || axis_b >= n || axis_a == axis_b) return PW_ERR_INVALID;
  if (dims[axis_a] != dims[axis_b] || dims[axis_a] < 2 || dims[axis_a] > 8) return PW_ERR_UNSUPPORTED;
  const int tg[2] = {axis_a, axis_b};
  Plan p;
  PW_TRY(build_plan(dims, n, tg, 2, &p));
  if (dtype == PW_C128)
    return launch_pair_t<double2>((const double2*)psi, (double2*)out, (const double2*)op_a,
                                  (const double2*)op_b, p, (int)dims[axis_a], false, (cudaStream_t)stream);
  if (dtype == PW_C64)
    return launch_pair_t<float2>((const float2*)psi, (float2*)out, (const float2*)op_a,
                                 (const float2*)op_b, p, (int)dims[axis_a], false, (cudaStream_t)stream);
  return PW_ERR_INVALID;
}
