// Reductions that follow an apply (rows a7-a12 of SURVEY.md section 8):
// measurement probabilities, partial traces, collapse gathers, POVM weights.
//
// All floating-point reductions are two-stage and ORDER-DETERMINISTIC (per-thread
// private accumulators -> fixed tree inside the block -> fixed-order sum over
// blocks): the same input always yields bit-identical probabilities, so the
// sampled outcome under a fixed key is reproducible run to run.
#include "pw_internal.cuh"
#include "pw_threefry.h"

namespace pw {

constexpr int kRedBlock = 128;
constexpr int kRedMaxBlocks = kNumSMs * 4;
constexpr size_t kRedSmemBudget = 160 * 1024;

// deterministic in-block tree sum of one double per thread; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* s_scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_scratch[warp] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) r += s_scratch[w];
  return r;
}

// ---- measurement probabilities -------------------------------------------------------
// partial[b][j] = sum over the fibers block b owns of |x[fiber, j]|^2
template <typename V>
__global__ void __launch_bounds__(kRedBlock)
k_probs_partial(const V* __restrict__ x, const FiberMap fm, const TargetMap tm,
                const int64_t mult, double* __restrict__ partial) {
  extern __shared__ double s_acc[];  // [t][blockDim] then scratch[4] then off[t]
  const uint32_t t = tm.t;
  double* s_scratch = s_acc + (size_t)t * blockDim.x;
  int64_t* s_off = (int64_t*)(s_scratch + 4);
  for (uint32_t j = threadIdx.x; j < t; j += blockDim.x) s_off[j] = tm.offset(j) * mult;
  for (uint32_t j = 0; j < t; ++j) s_acc[(size_t)j * blockDim.x + threadIdx.x] = 0.0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib; f += stride) {
    const int64_t base = fm.base(f) * mult;
    for (uint32_t j = 0; j < t; ++j) {
      const V v = x[base + s_off[j]];
      s_acc[(size_t)j * blockDim.x + threadIdx.x] += (double)v.x * v.x + (double)v.y * v.y;
    }
  }
  for (uint32_t j = 0; j < t; ++j) {
    double r = block_sum(s_acc[(size_t)j * blockDim.x + threadIdx.x], s_scratch);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * t + j] = r;
  }
}

// density-matrix diagonal: only the real part is summed (core/kernels.py:57)
template <typename V>
__global__ void __launch_bounds__(kRedBlock)
k_probs_partial_diag(const V* __restrict__ x, const FiberMap fm, const TargetMap tm,
                     const int64_t mult, double* __restrict__ partial) {
  extern __shared__ double s_acc[];
  const uint32_t t = tm.t;
  double* s_scratch = s_acc + (size_t)t * blockDim.x;
  int64_t* s_off = (int64_t*)(s_scratch + 4);
  for (uint32_t j = threadIdx.x; j < t; j += blockDim.x) s_off[j] = tm.offset(j) * mult;
  for (uint32_t j = 0; j < t; ++j) s_acc[(size_t)j * blockDim.x + threadIdx.x] = 0.0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib; f += stride) {
    const int64_t base = fm.base(f) * mult;
    for (uint32_t j = 0; j < t; ++j)
      s_acc[(size_t)j * blockDim.x + threadIdx.x] += (double)x[base + s_off[j]].x;
  }
  for (uint32_t j = 0; j < t; ++j) {
    double r = block_sum(s_acc[(size_t)j * blockDim.x + threadIdx.x], s_scratch);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * t + j] = r;
  }
}

// large-t fallback: block (j, chunk) reduces one operator index over a slice of fibers
template <typename V, bool DIAG>
__global__ void __launch_bounds__(kRedBlock)
k_probs_partial_wide(const V* __restrict__ x, const FiberMap fm, const TargetMap tm,
                     const int64_t mult, const int nchunks, double* __restrict__ partial) {
  __shared__ double s_scratch[4];
  const uint32_t j = blockIdx.x / nchunks, c = blockIdx.x % nchunks;
  const int64_t off = tm.offset(j) * mult;
  double acc = 0;
  for (uint64_t f = (uint64_t)c * blockDim.x + threadIdx.x; f < fm.nfib;
       f += (uint64_t)nchunks * blockDim.x) {
    const V v = x[fm.base(f) * mult + off];
    acc += DIAG ? (double)v.x : (double)v.x * v.x + (double)v.y * v.y;
  }
  double r = block_sum(acc, s_scratch);
  if (threadIdx.x == 0) partial[(size_t)c * tm.t + j] = r;
}

__global__ void __launch_bounds__(256)
k_sums_finalize(const double* __restrict__ partial, const int nblocks, const uint32_t t,
                double* __restrict__ sums) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= t) return;
  double p = 0;
  for (int b = 0; b < nblocks; ++b) p += partial[(size_t)b * t + j];
  sums[j] = p;
}

// probs[j] = sums[j] / sum(sums), divided in the state's real dtype like k_probs_finalize
template <typename R>
__global__ void __launch_bounds__(256)
k_probs_from_sums(const double* __restrict__ sums, const uint32_t t, R* __restrict__ probs) {
  __shared__ double s_scratch[8];
  __shared__ double s_total;
  double local = 0;
  for (uint32_t j = threadIdx.x; j < t; j += blockDim.x) local += sums[j];
  const double tot = block_sum(local, s_scratch);
  if (threadIdx.x == 0) s_total = tot;
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < t; j += blockDim.x) probs[j] = (R)((R)sums[j] / (R)s_total);
}

// ---- strided views (multi-GPU layer) ----------------------------------------------------
// out[j * tc + j'] = sum_f x[base(f) + row(j) + col(j')]: the partial trace of a block of rho
// whose rows / columns / traced pairs are arbitrary strided axes
template <typename V>
__global__ void __launch_bounds__(kRedBlock)
k_trace_strided(const V* __restrict__ x, const FiberMap fm, const TargetMap rows, const TargetMap cols,
                V* __restrict__ out) {
  __shared__ double s_scratch[4];
  const uint64_t nout = (uint64_t)rows.t * cols.t;
  for (uint64_t o = blockIdx.x; o < nout; o += gridDim.x) {
    const int64_t off = rows.offset((uint32_t)(o / cols.t)) + cols.offset((uint32_t)(o % cols.t));
    double re = 0, im = 0;
    for (uint64_t f = threadIdx.x; f < fm.nfib; f += blockDim.x) {
      const V v = x[off + fm.base(f)];
      re += v.x; im += v.y;
    }
    re = block_sum(re, s_scratch);
    im = block_sum(im, s_scratch);
    if (threadIdx.x == 0) { V v; v.x = re; v.y = im; out[o] = v; }
  }
}

// out[a, b] = sum_f x[f, a] conj(y[f, b]): the operator cotangent of a linear apply (the VJP of
// y = O psi with respect to O pairs the output cotangent with the input state)
template <typename V>
__global__ void __launch_bounds__(kRedBlock)
k_cross_reduced(const V* __restrict__ x, const V* __restrict__ y, const FiberMap fm, const TargetMap tm,
                V* __restrict__ out) {
  __shared__ double s_scratch[4];
  const uint64_t t = tm.t, nout = t * t;
  for (uint64_t o = blockIdx.x; o < nout; o += gridDim.x) {
    const int64_t oa = tm.offset((uint32_t)(o / t)), ob = tm.offset((uint32_t)(o % t));
    double re = 0, im = 0;
    for (uint64_t f = threadIdx.x; f < fm.nfib; f += blockDim.x) {
      const int64_t base = fm.base(f);
      const V u = x[base + oa], v = y[base + ob];
      re += (double)u.x * v.x + (double)u.y * v.y;   // u * conj(v)
      im += (double)u.y * v.x - (double)u.x * v.y;
    }
    re = block_sum(re, s_scratch);
    im = block_sum(im, s_scratch);
    if (threadIdx.x == 0) { V r; r.x = re; r.y = im; out[o] = r; }
  }
}

// out[f] = x[base(f)] / den,  den = 1 | sqrt(probs[o] + 1e-18) | probs[o] + 1e-18 (the collapse
// rescaling of core/kernels.py:205,225 on a strided slice)
template <typename V>
__global__ void k_gather_strided(const V* __restrict__ x, V* __restrict__ out, const FiberMap fm,
                                 const typename RealOf<V>::type* __restrict__ probs,
                                 const int64_t* __restrict__ outcome, const int mode) {
  using R = typename RealOf<V>::type;
  R den = (R)1;
  if (mode && probs) {
    const R p = probs[*outcome] + (R)1e-18;
    den = mode == 1 ? sqrt(p) : p;
  }
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib;
       f += (uint64_t)gridDim.x * blockDim.x) {
    V v = x[fm.base(f)];
    v.x = v.x / den;
    v.y = v.y / den;
    out[f] = v;
  }
}

template <typename R>
__global__ void __launch_bounds__(256)
k_probs_finalize(const double* __restrict__ partial, const int nblocks, const uint32_t t,
                 R* __restrict__ probs) {
  __shared__ double s_scratch[8];
  __shared__ double s_total;
  double local = 0;
  for (uint32_t j = threadIdx.x; j < t; j += blockDim.x) {
    double p = 0;
    for (int b = 0; b < nblocks; ++b) p += partial[(size_t)b * t + j];
    local += p;
  }
  double tot = block_sum(local, s_scratch);
  if (threadIdx.x == 0) s_total = tot;
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < t; j += blockDim.x) {
    double p = 0;
    for (int b = 0; b < nblocks; ++b) p += partial[(size_t)b * t + j];
    // the reference divides in the state's real dtype (probs / jnp.sum(probs))
    probs[j] = (R)((R)p / (R)s_total);
  }
}

// t <= 8: private register accumulators, one fiber per thread per iteration
template <typename V, int T, bool DIAG>
__global__ void __launch_bounds__(256)
k_probs_small(const V* __restrict__ x, const FiberMap fm, const TargetMap tm, const int64_t mult,
              double* __restrict__ partial) {
  __shared__ double s_scratch[8];
  __shared__ int64_t s_off[T];
  if (threadIdx.x < T) s_off[threadIdx.x] = tm.offset(threadIdx.x) * mult;
  __syncthreads();
  double acc[T];
#pragma unroll
  for (int j = 0; j < T; ++j) acc[j] = 0.0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib; f += stride) {
    const int64_t base = fm.base(f) * mult;
    V v[T];
#pragma unroll
    for (int j = 0; j < T; ++j) v[j] = __ldcg(x + base + s_off[j]);
#pragma unroll
    for (int j = 0; j < T; ++j)
      acc[j] += DIAG ? (double)v[j].x : (double)v[j].x * v[j].x + (double)v[j].y * v[j].y;
  }
#pragma unroll
  for (int j = 0; j < T; ++j) {
    const double r = block_sum(acc[j], s_scratch);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * T + j] = r;
  }
}

template <typename V, int T>
static void launch_probs_small(const V* x, const Plan& p, int64_t mult, bool diag, int nblocks,
                               double* partial, cudaStream_t s) {
  if (diag) k_probs_small<V, T, true><<<nblocks, 256, 0, s>>>(x, p.fm, p.tm, mult, partial);
  else      k_probs_small<V, T, false><<<nblocks, 256, 0, s>>>(x, p.fm, p.tm, mult, partial);
}

template <typename V>
static int probs_impl(const V* x, void* probs, const Plan& p, int64_t mult, bool diag,
                      cudaStream_t s, bool raw = false) {
  using R = typename RealOf<V>::type;
  const uint32_t t = p.tm.t;
  const size_t smem = ((size_t)t * kRedBlock + 4) * sizeof(double) + (size_t)t * sizeof(int64_t);
  Scratch partial;
  int nblocks;
  if (t <= 8) {
    nblocks = grid_for(p.fm.nfib, 256, 6);
    PW_TRY(partial.alloc(sizeof(double) * (size_t)nblocks * t, s));
    double* pp = (double*)partial.p;
    switch (t) {
      case 1: launch_probs_small<V, 1>(x, p, mult, diag, nblocks, pp, s); break;
      case 2: launch_probs_small<V, 2>(x, p, mult, diag, nblocks, pp, s); break;
      case 3: launch_probs_small<V, 3>(x, p, mult, diag, nblocks, pp, s); break;
      case 4: launch_probs_small<V, 4>(x, p, mult, diag, nblocks, pp, s); break;
      case 5: launch_probs_small<V, 5>(x, p, mult, diag, nblocks, pp, s); break;
      case 6: launch_probs_small<V, 6>(x, p, mult, diag, nblocks, pp, s); break;
      case 7: launch_probs_small<V, 7>(x, p, mult, diag, nblocks, pp, s); break;
      default: launch_probs_small<V, 8>(x, p, mult, diag, nblocks, pp, s); break;
    }
    PW_LAUNCHED();
  } else if (smem <= kRedSmemBudget) {
    nblocks = grid_for(p.fm.nfib, kRedBlock, 4);
    if (nblocks > kRedMaxBlocks) nblocks = kRedMaxBlocks;
    PW_TRY(partial.alloc(sizeof(double) * (size_t)nblocks * t, s));
    auto kern = diag ? k_probs_partial_diag<V> : k_probs_partial<V>;
    if (smem > 48 * 1024)
      PW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nblocks, kRedBlock, smem, s>>>(x, p.fm, p.tm, mult, (double*)partial.p);
    PW_LAUNCHED();
  } else {
    uint64_t want = (p.fm.nfib + kRedBlock - 1) / kRedBlock;
    uint64_t cap = (uint64_t)kRedMaxBlocks * 8 / t;
    nblocks = (int)(want < cap ? want : cap);
    if (nblocks < 1) nblocks = 1;
    if ((uint64_t)nblocks * t > 0x7fffffffull) return PW_ERR_UNSUPPORTED;
    PW_TRY(partial.alloc(sizeof(double) * (size_t)nblocks * t, s));
    if (diag)
      k_probs_partial_wide<V, true><<<nblocks * t, kRedBlock, 0, s>>>(x, p.fm, p.tm, mult, nblocks, (double*)partial.p);
    else
      k_probs_partial_wide<V, false><<<nblocks * t, kRedBlock, 0, s>>>(x, p.fm, p.tm, mult, nblocks, (double*)partial.p);
    PW_LAUNCHED();
  }
  if (raw)  // un-normalised double sums: the multi-GPU layer adds them across ranks first
    k_sums_finalize<<<(t + 255) / 256, 256, 0, s>>>((const double*)partial.p, nblocks, t, (double*)probs);
  else
    k_probs_finalize<R><<<1, 256, 0, s>>>((const double*)partial.p, nblocks, t, (R*)probs);
  PW_LAUNCHED();
  return PW_OK;
}

// ---- reduced density matrix of kept subsystems from a state vector ------------------
// partial[b][a*t+a'] (a <= a') = sum_f x[f,a] conj(x[f,a'])
template <typename V>
__global__ void __launch_bounds__(kRedBlock)
k_reduced_partial(const V* __restrict__ x, const FiberMap fm, const TargetMap tm,
                  double2* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const uint32_t t = tm.t, npair = t * (t + 1) / 2;
  double2* s_acc = (double2*)s_raw;                       // [npair][blockDim]
  double2* s_x = s_acc + (size_t)npair * blockDim.x;      // [t][blockDim]
  double* s_scratch = (double*)(s_x + (size_t)t * blockDim.x);
  int64_t* s_off = (int64_t*)(s_scratch + 4);
  for (uint32_t j = threadIdx.x; j < t; j += blockDim.x) s_off[j] = tm.offset(j);
  for (uint32_t q = 0; q < npair; ++q) s_acc[(size_t)q * blockDim.x + threadIdx.x] = make_double2(0, 0);
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib; f += stride) {
    const int64_t base = fm.base(f);
    for (uint32_t j = 0; j < t; ++j) {
      const V v = x[base + s_off[j]];
      s_x[(size_t)j * blockDim.x + threadIdx.x] = make_double2(v.x, v.y);
    }
    uint32_t q = 0;
    for (uint32_t a = 0; a < t; ++a) {
      const double2 xa = s_x[(size_t)a * blockDim.x + threadIdx.x];
      for (uint32_t b = a; b < t; ++b, ++q) {
        const double2 xb = s_x[(size_t)b * blockDim.x + threadIdx.x];
        double2 acc = s_acc[(size_t)q * blockDim.x + threadIdx.x];
        acc.x += xa.x * xb.x + xa.y * xb.y;   // xa * conj(xb)
        acc.y += xa.y * xb.x - xa.x * xb.y;
        s_acc[(size_t)q * blockDim.x + threadIdx.x] = acc;
      }
    }
  }
  for (uint32_t q = 0; q < npair; ++q) {
    const double2 v = s_acc[(size_t)q * blockDim.x + threadIdx.x];
    double re = block_sum(v.x, s_scratch);
    double im = block_sum(v.y, s_scratch);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * npair + q] = make_double2(re, im);
  }
}

template <typename V>
__global__ void k_reduced_finalize(const double2* __restrict__ partial, const int nblocks,
                                   const uint32_t t, V* __restrict__ out) {
  const uint32_t npair = t * (t + 1) / 2;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < npair; q += gridDim.x * blockDim.x) {
    double re = 0, im = 0;
    for (int b = 0; b < nblocks; ++b) {
      const double2 v = partial[(size_t)b * npair + q];
      re += v.x; im += v.y;
    }
    // invert q -> (a, b), a <= b, row-major upper triangle
    uint32_t a = 0, rem = q;
    while (rem >= t - a) { rem -= t - a; ++a; }
    const uint32_t b2 = a + rem;
    V v; v.x = re; v.y = im;
    out[(size_t)a * t + b2] = v;
    if (b2 != a) { v.y = -im; out[(size_t)b2 * t + a] = v; }
  }
}

// wide fallback: one block per output element (a, a'), all fibers
template <typename V>
__global__ void __launch_bounds__(kRedBlock)
k_reduced_wide(const V* __restrict__ x, const FiberMap fm, const TargetMap tm, V* __restrict__ out) {
  __shared__ double s_scratch[4];
  const uint32_t t = tm.t, a = blockIdx.x / t, b = blockIdx.x % t;
  const int64_t oa = tm.offset(a), ob = tm.offset(b);
  double re = 0, im = 0;
  for (uint64_t f = threadIdx.x; f < fm.nfib; f += blockDim.x) {
    const int64_t base = fm.base(f);
    const V xa = x[base + oa], xb = x[base + ob];
    re += (double)xa.x * xb.x + (double)xa.y * xb.y;
    im += (double)xa.y * xb.x - (double)xa.x * xb.y;
  }
  re = block_sum(re, s_scratch);
  im = block_sum(im, s_scratch);
  if (threadIdx.x == 0) { V v; v.x = re; v.y = im; out[blockIdx.x] = v; }
}

// t <= 8: Hermitian accumulators in registers.  PAIR (complex64, even contiguous untouched run,
// `fm` = fiber pairs): two adjacent fibers per thread from one 128-bit load each -- with 8-byte
// loads the kernel sat at 52 % of the complex64 roof (latency-bound at ~100 registers).
template <typename V, int T, bool PAIR>
__global__ void __launch_bounds__(256)
k_reduced_small(const V* __restrict__ x, const FiberMap fm, const TargetMap tm,
                double2* __restrict__ partial) {
  constexpr int NP = T * (T + 1) / 2;
  __shared__ double s_scratch[8];
  __shared__ int64_t s_off[T];
  if (threadIdx.x < T) s_off[threadIdx.x] = tm.offset(threadIdx.x);
  __syncthreads();
  double re[NP], im[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) { re[q] = 0.0; im[q] = 0.0; }
  auto accumulate = [&](const double (&vx)[T], const double (&vy)[T]) {
    int q = 0;
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
      for (int b = a; b < T; ++b, ++q) {
        re[q] = fma(vx[a], vx[b], fma(vy[a], vy[b], re[q]));   // xa * conj(xb)
        if (b != a) im[q] = fma(vy[a], vx[b], fma(-vx[a], vy[b], im[q]));
      }
  };
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib; f += stride) {
    const int64_t base = fm.base(f);
    double vx[T], vy[T];
    if constexpr (PAIR) {
      double wx[T], wy[T];
#pragma unroll
      for (int j = 0; j < T; ++j) {
        const C64x2 v = ld_c64x2((const float2*)x + base + s_off[j]);
        vx[j] = v.a.x; vy[j] = v.a.y; wx[j] = v.b.x; wy[j] = v.b.y;
      }
      accumulate(vx, vy);
      accumulate(wx, wy);
    } else {
#pragma unroll
      for (int j = 0; j < T; ++j) { const V v = __ldcg(x + base + s_off[j]); vx[j] = v.x; vy[j] = v.y; }
      accumulate(vx, vy);
    }
  }
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const double r = block_sum(re[q], s_scratch);
    const double i = block_sum(im[q], s_scratch);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * NP + q] = make_double2(r, i);
  }
}

template <typename V, bool PAIR>
static bool launch_reduced_small_p(const V* x, const FiberMap& fm, const TargetMap& tm, int nblocks,
                                   double2* partial, cudaStream_t s) {
  switch (tm.t) {
    case 1: k_reduced_small<V, 1, PAIR><<<nblocks, 256, 0, s>>>(x, fm, tm, partial); return true;
    case 2: k_reduced_small<V, 2, PAIR><<<nblocks, 256, 0, s>>>(x, fm, tm, partial); return true;
    case 3: k_reduced_small<V, 3, PAIR><<<nblocks, 256, 0, s>>>(x, fm, tm, partial); return true;
    case 4: k_reduced_small<V, 4, PAIR><<<nblocks, 256, 0, s>>>(x, fm, tm, partial); return true;
    case 5: k_reduced_small<V, 5, PAIR><<<nblocks, 256, 0, s>>>(x, fm, tm, partial); return true;
    case 6: k_reduced_small<V, 6, PAIR><<<nblocks, 256, 0, s>>>(x, fm, tm, partial); return true;
    case 7: k_reduced_small<V, 7, PAIR><<<nblocks, 256, 0, s>>>(x, fm, tm, partial); return true;
    case 8: k_reduced_small<V, 8, PAIR><<<nblocks, 256, 0, s>>>(x, fm, tm, partial); return true;
    default: return false;
  }
}

template <typename V>
static bool launch_reduced_small(const V* x, const Plan& p, int nblocks, double2* partial,
                                 cudaStream_t s) {
  if constexpr (sizeof(V) == 8) {
    FiberMap pm;
    if (options().c64_pairs.load(std::memory_order_relaxed) && pair_fiber_map(p.fm, x, x, &pm))
      return launch_reduced_small_p<V, true>(x, pm, p.tm, nblocks, partial, s);
  }
  return launch_reduced_small_p<V, false>(x, p.fm, p.tm, nblocks, partial, s);
}

template <typename V>
static int reduced_impl(const V* x, V* out, const Plan& p, cudaStream_t s) {
  const uint32_t t = p.tm.t;
  const size_t npair = (size_t)t * (t + 1) / 2;
  const size_t smem = npair * kRedBlock * sizeof(double2) + 4 * sizeof(double) +
                      t * sizeof(int64_t) + (size_t)t * kRedBlock * sizeof(double2);
  // launch-bound states (the Mach-Zehnder / Jaynes-Cummings configs): ONE launch, one block per
  // output entry, no scratch allocation, instead of partial sums + finalize
  if (t <= 8 && (uint64_t)p.fm.nfib * t <= 16384) {
    k_reduced_wide<V><<<t * t, kRedBlock, 0, s>>>(x, p.fm, p.tm, out);
    PW_LAUNCHED();
    return PW_OK;
  }
  if (t <= 8) {
    const int nblocks = grid_for(p.fm.nfib, 256, 2);
    Scratch partial;
    PW_TRY(partial.alloc(sizeof(double2) * (size_t)nblocks * npair, s));
    launch_reduced_small<V>(x, p, nblocks, (double2*)partial.p, s);
    PW_LAUNCHED();
    k_reduced_finalize<V><<<(int)((npair + 127) / 128), 128, 0, s>>>((const double2*)partial.p, nblocks, t, out);
    PW_LAUNCHED();
    return PW_OK;
  }
  if (smem <= kRedSmemBudget) {
    int nblocks = grid_for(p.fm.nfib, kRedBlock, 1);
    Scratch partial;
    PW_TRY(partial.alloc(sizeof(double2) * (size_t)nblocks * npair, s));
    if (smem > 48 * 1024)
      PW_CUDA(cudaFuncSetAttribute(k_reduced_partial<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_reduced_partial<V><<<nblocks, kRedBlock, smem, s>>>(x, p.fm, p.tm, (double2*)partial.p);
    PW_LAUNCHED();
    k_reduced_finalize<V><<<(int)((npair + 127) / 128), 128, 0, s>>>((const double2*)partial.p, nblocks, t, out);
    PW_LAUNCHED();
    return PW_OK;
  }
  if ((uint64_t)t * t > 0x7fffffffull) return PW_ERR_UNSUPPORTED;
  k_reduced_wide<V><<<t * t, kRedBlock, 0, s>>>(x, p.fm, p.tm, out);
  PW_LAUNCHED();
  return PW_OK;
}

// ---- partial trace of a density matrix ----------------------------------------------
// out[k, k'] = sum_f rho[row(k) + rest(f), col(k') + rest(f)]
template <typename V, bool BLOCK_PER_OUT>
__global__ void __launch_bounds__(kRedBlock)
k_trace_out(const V* __restrict__ rho, const FiberMap fm, const TargetMap tm, const int64_t N,
            V* __restrict__ out) {
  __shared__ double s_scratch[4];
  const uint64_t t = tm.t, nout = t * t;
  if (BLOCK_PER_OUT) {
    for (uint64_t o = blockIdx.x; o < nout; o += gridDim.x) {
      const int64_t off = tm.offset((uint32_t)(o / t)) * N + tm.offset((uint32_t)(o % t));
      double re = 0, im = 0;
      for (uint64_t f = threadIdx.x; f < fm.nfib; f += blockDim.x) {
        const V v = rho[off + fm.base(f) * (N + 1)];
        re += v.x; im += v.y;
      }
      re = block_sum(re, s_scratch);
      im = block_sum(im, s_scratch);
      if (threadIdx.x == 0) { V v; v.x = re; v.y = im; out[o] = v; }
    }
  } else {
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nout;
         o += (uint64_t)gridDim.x * blockDim.x) {
      const int64_t off = tm.offset((uint32_t)(o / t)) * N + tm.offset((uint32_t)(o % t));
      double re = 0, im = 0;
      for (uint64_t f = 0; f < fm.nfib; ++f) {
        const V v = rho[off + fm.base(f) * (N + 1)];
        re += v.x; im += v.y;
      }
      V v; v.x = re; v.y = im; out[o] = v;
    }
  }
}

// ---- collapse (gather the selected slice and rescale) ---------------------------------
template <typename V>
__global__ void k_collapse_vec(const V* __restrict__ psi, V* __restrict__ post,
                               const int64_t* __restrict__ outcome,
                               const typename RealOf<V>::type* __restrict__ probs,
                               const FiberMap fm, const TargetMap tm) {
  using R = typename RealOf<V>::type;
  const uint32_t o = (uint32_t)*outcome;
  const int64_t off = tm.offset(o);
  // the reference divides (core/kernels.py:205); dividing keeps the last bit equal.
  // probs == nullptr: plain gather (jnp.take of the legacy path, measurements.py:135)
  const R den = probs ? sqrt(probs[o] + (R)1e-18) : (R)1;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib;
       f += (uint64_t)gridDim.x * blockDim.x) {
    V v = psi[fm.base(f) + off];
    v.x = v.x / den;
    v.y = v.y / den;
    post[f] = v;
  }
}

template <typename V>
__global__ void k_collapse_mat(const V* __restrict__ rho, V* __restrict__ post,
                               const int64_t* __restrict__ outcome,
                               const typename RealOf<V>::type* __restrict__ probs,
                               const FiberMap fm, const TargetMap tm, const int64_t N) {
  using R = typename RealOf<V>::type;
  const uint32_t o = (uint32_t)*outcome;
  const int64_t off = tm.offset(o);
  const R den = probs ? probs[o] + (R)1e-18 : (R)1;
  const uint64_t r = fm.nfib, total = r * r;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t fr = i / r, fc = i - fr * r;
    V v = rho[(fm.base(fr) + off) * N + fm.base(fc) + off];
    v.x = v.x / den;
    v.y = v.y / den;
    post[i] = v;
  }
}

// ---- sampling ---------------------------------------------------------------------
// jax.random.choice(key, a=t, p=probs): c = cumsum(p); r = c[-1]*(1-u); searchsorted-left.
// The oracle's cumsum is the SEQUENTIAL one (np.cumsum), which no parallel sum reproduces
// bit for bit, so the draw runs in two tiers:
//  * parallel locate: chunk sums in double (any association) give every cumulative value
//    with an error <= margin = 4 t eps (sum p); when the located bin is clear of its two
//    edges by more than the margin, EVERY summation order -- the oracle's sequential one and
//    whatever scan XLA lowers jnp.cumsum to -- returns the same index, and we are done;
//  * otherwise (the draw landed within rounding distance of a bin edge: probability
//    ~ 8 t^2 eps for t comparable bins) one warp re-runs the sequential cumsum in the
//    probabilities' own precision, with coalesced loads and register shuffles
//    (~6 cycles per element instead of one dependent global load per element).
// An index that is never reached (NaN probabilities, zero state) is clamped to t-1, the
// value an out-of-range gather index takes in JAX.
constexpr int kDrawChunk = 4096;
constexpr int kDrawBlock = 256;

template <typename R>
__device__ __forceinline__ R draw_uniform(const uint32_t k0, const uint32_t k1) {
  uint32_t b0, b1;
  threefry2x32(k0, k1, 0u, 0u, &b0, &b1);
  if (sizeof(R) == 8) {
    const unsigned long long bits = ((((unsigned long long)b0 << 32) | b1) >> 12) | 0x3FF0000000000000ull;
    return (R)(__longlong_as_double((long long)bits) - 1.0);
  }
  const uint32_t bits = ((b0 ^ b1) >> 9) | 0x3F800000u;
  return (R)(__uint_as_float(bits) - 1.0f);
}

// one warp, all lanes carry the same running sum: the exact sequential cumsum semantics.
// STOP = false: returns the total; STOP = true: returns in *idx the first i with !(c < r).
template <typename R, bool STOP>
__device__ __forceinline__ R warp_sequential_cumsum(const R* __restrict__ probs, const int64_t t,
                                                    const R r, int64_t* idx) {
  const int lane = threadIdx.x & 31;
  R c = 0;
  constexpr int U = 8;
  for (int64_t base = 0; base < t; base += 32 * U) {
    R v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int64_t i = base + k * 32 + lane;
      v[k] = i < t ? probs[i] : (R)0;
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int64_t left = t - (base + k * 32);
      if (left <= 0) break;
      const int m = left < 32 ? (int)left : 32;
      if (m == 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          c += __shfl_sync(0xffffffffu, v[k], j);
          if (STOP && !(c < r)) { *idx = base + k * 32 + j; return c; }
        }
      } else {
        for (int j = 0; j < m; ++j) {
          c += __shfl_sync(0xffffffffu, v[k], j);
          if (STOP && !(c < r)) { *idx = base + k * 32 + j; return c; }
        }
      }
    }
  }
  return c;
}

// stage 1 (t > kDrawChunk): chunk_sum[b] = sum of chunk b in double, fixed order
template <typename R>
__global__ void __launch_bounds__(kDrawBlock)
k_draw_chunks(const R* __restrict__ probs, const int64_t t, double* __restrict__ chunk_sum) {
  __shared__ double s_scratch[kDrawBlock / 32];
  const int64_t lo = (int64_t)blockIdx.x * kDrawChunk;
  double acc = 0;
  for (int i = threadIdx.x; i < kDrawChunk; i += kDrawBlock)
    if (lo + i < t) acc += (double)probs[lo + i];
  acc = block_sum(acc, s_scratch);
  if (threadIdx.x == 0) chunk_sum[blockIdx.x] = acc;
}

// stage 2: one block.  chunk_sum == nullptr: t <= kDrawChunk, the single chunk is summed here.
template <typename R>
__global__ void __launch_bounds__(kDrawBlock)
k_draw(const R* __restrict__ probs, const int64_t t, const double* __restrict__ chunk_sum,
       const int64_t nchunks, uint32_t k0, uint32_t k1, const uint32_t* __restrict__ key_dev,
       const int exact, int64_t* __restrict__ outcome) {
  __shared__ double s_part[kDrawBlock];
  __shared__ double s_total, s_before;
  __shared__ int64_t s_chunk;
  __shared__ int s_clear;
  const int tid = threadIdx.x;
  if (key_dev) { k0 = key_dev[0]; k1 = key_dev[1]; }  // key produced on the device (XLA FFI path)
  const R u = draw_uniform<R>(k0, k1);
  const double eps = sizeof(R) == 8 ? 1.1102230246251565e-16 : 5.9604644775390625e-8;

  // ---- approximate prefix over the chunks: thread-contiguous slices, then a 256-entry scan
  const int64_t per = (nchunks + kDrawBlock - 1) / kDrawBlock;
  const int64_t c_lo = (int64_t)tid * per, c_hi = min(c_lo + per, nchunks);
  double slice = 0;
  if (chunk_sum) {
    for (int64_t b = c_lo; b < c_hi; ++b) slice += chunk_sum[b];
  } else {
    // single chunk: thread slices of the elements themselves (16 consecutive per thread)
    for (int i = tid * (kDrawChunk / kDrawBlock); i < (tid + 1) * (kDrawChunk / kDrawBlock); ++i)
      if (i < t) slice += (double)probs[i];
  }
  s_part[tid] = slice;
  __syncthreads();
  if (tid == 0) {
    double tot = 0;
    for (int i = 0; i < kDrawBlock; ++i) tot += s_part[i];
    s_total = tot;
  }
  __syncthreads();
  const double total = s_total;
  const double r_a = total * (1.0 - (double)u);
  const double margin = 4.0 * (double)t * eps * fabs(total);

  if (tid == 0) {
    // locate the first element whose approximate cumulative value reaches r_a
    double before = 0;
    int sl = 0;
    while (sl < kDrawBlock - 1 && before + s_part[sl] < r_a) before += s_part[sl++];
    int64_t chunk = 0;
    if (chunk_sum) {
      chunk = min((int64_t)sl * per, nchunks - 1);
      const int64_t last = min((int64_t)(sl + 1) * per, nchunks) - 1;
      while (chunk < last && before + chunk_sum[chunk] < r_a) before += chunk_sum[chunk++];
    }
    s_before = chunk_sum ? before : 0.0;  // single chunk: the element slices are re-walked below
    s_chunk = chunk;
    s_clear = 0;
  }
  __syncthreads();
  // inside the chunk: 16 consecutive elements per thread
  const int64_t e_lo = s_chunk * kDrawChunk;
  constexpr int E = kDrawChunk / kDrawBlock;
  double mine = 0;
  for (int i = 0; i < E; ++i) {
    const int64_t g = e_lo + tid * E + i;
    if (g < t) mine += (double)probs[g];
  }
  __syncthreads();
  s_part[tid] = mine;
  __syncthreads();
  if (tid == 0) {
    double c = s_before;
    int th = 0;
    while (th < kDrawBlock - 1 && c + s_part[th] < r_a) c += s_part[th++];
    int64_t idx = -1;
    double prev = c, at = c;
    for (int i = 0; i < E; ++i) {
      const int64_t g = e_lo + th * E + i;
      if (g >= t) break;
      prev = at;
      at += (double)probs[g];
      if (!(at < r_a)) { idx = g; break; }
    }
    // clear of both edges by more than the summation error => order independent
    const bool ok = idx >= 0 && (at - r_a > margin) && (idx == 0 || r_a - prev > margin) &&
                    total > 0 && total < 1.0e300;
    if ((ok && exact != 2) || !exact) {  // exact == 2 (tests): always take the sequential tier
      *outcome = idx >= 0 ? idx : t - 1;
      s_clear = 1;
    }
  }
  __syncthreads();
  if (s_clear || tid >= 32) return;
  // ---- exact tier: the sequential cumsum of the oracle, one warp
  int64_t idx = t - 1;
  const R tot = warp_sequential_cumsum<R, false>(probs, t, (R)0, nullptr);
  const R r = tot * ((R)1 - u);
  warp_sequential_cumsum<R, true>(probs, t, r, &idx);
  if (tid == 0) *outcome = idx;
}

// ---- POVM weights on the reduced state ----------------------------------------------
template <typename V>
__global__ void __launch_bounds__(kRedBlock)
k_povm_traces(const V* __restrict__ rho_t, const V* __restrict__ ops, const int64_t t,
              double* __restrict__ traces) {
  __shared__ double s_scratch[4];
  const V* E = ops + (size_t)blockIdx.x * t * t;
  double acc = 0;
  // Tr(E rho E^dagger) = sum_{a,b,c} E[a,b] rho[b,c] conj(E[a,c]); threads over (a,b)
  for (int64_t ab = threadIdx.x; ab < t * t; ab += blockDim.x) {
    const int64_t a = ab / t, b = ab % t;
    const V e = E[ab];
    double2 inner = make_double2(0, 0);
    for (int64_t c = 0; c < t; ++c) {
      const V r = rho_t[b * t + c], ec = E[a * t + c];
      inner.x += (double)r.x * ec.x + (double)r.y * ec.y;   // r * conj(ec)
      inner.y += (double)r.y * ec.x - (double)r.x * ec.y;
    }
    acc += (double)e.x * inner.x - (double)e.y * inner.y;    // real part only
  }
  acc = block_sum(acc, s_scratch);
  if (threadIdx.x == 0) traces[blockIdx.x] = acc;
}

template <typename R>
__global__ void k_povm_normalise(const double* __restrict__ traces_d, const int K,
                                 R* __restrict__ probs, R* __restrict__ traces) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  R total = 0;
  for (int k = 0; k < K; ++k) total += (R)traces_d[k];
  for (int k = 0; k < K; ++k) {
    traces[k] = (R)traces_d[k];
    probs[k] = (R)traces_d[k] / total;
  }
}

template <typename V>
__global__ void k_povm_select(const V* __restrict__ ops, const int64_t t,
                              const int64_t* __restrict__ outcome,
                              const typename RealOf<V>::type* __restrict__ traces,
                              V* __restrict__ op_out) {
  using R = typename RealOf<V>::type;
  const int64_t o = *outcome;
  const R s = sqrt(traces[o] + (R)1e-18);
  for (int64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < t * t; i += gridDim.x * blockDim.x) {
    V v = ops[o * t * t + i];
    v.x /= s; v.y /= s;
    op_out[i] = v;
  }
}

// ---- norm / max / scale ---------------------------------------------------------------
template <typename V>
__global__ void __launch_bounds__(256)
k_norm_partial(const V* __restrict__ x, const int64_t count, double* __restrict__ partial) {
  __shared__ double s_scratch[8];
  __shared__ double s_max[8];
  double sum = 0, mx = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const V v = x[i];
    const double a = (double)v.x * v.x + (double)v.y * v.y;
    sum += a;
    mx = a > mx ? a : mx;
  }
  sum = block_sum(sum, s_scratch);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)blockDim.x / 32; ++w) mx = fmax(mx, s_max[w]);
    partial[2 * blockIdx.x] = sum;
    partial[2 * blockIdx.x + 1] = mx;
  }
}

__global__ void k_norm_finalize(const double* __restrict__ partial, const int nblocks,
                                double* __restrict__ stats) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double sum = 0, mx = 0;
  for (int b = 0; b < nblocks; ++b) {
    sum += partial[2 * b];
    mx = fmax(mx, partial[2 * b + 1]);
  }
  stats[0] = sum;
  stats[1] = mx;
}

// DIV: x /= s with a true division, so that x / ||x|| rounds exactly like the reference's
// `state / jnp.linalg.norm(state)` (a basis state comes out as exactly 1.0, which the
// Vector -> Label contraction tests for).
template <typename V, bool DIV>
__global__ void k_scale(V* __restrict__ x, const int64_t count, const double* __restrict__ scale) {
  using R = typename RealOf<V>::type;
  const R s = (R)*scale;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    V v = x[i];
    if (DIV) { v.x /= s; v.y /= s; } else { v.x *= s; v.y *= s; }
    x[i] = v;
  }
}


// ---- fused post-apply bookkeeping ---------------------------
// One pass over a state vector (or the diagonal of a density matrix) that yields, for
// EVERY subsystem at once, the highest basis index carrying a non-zero amplitude, plus
// the squared norm (trace) and the largest |x|^2.  The reference obtains the same
// integers with one trace_out -> psi psi^dagger -> nonzero() per Fock mode after every
// composite apply (state/composite_envelope.py:688-696, _math/ops.py:472-512).
//
// Index model: the trailing axes whose product Q fits one tile are "lo"; the axis before
// them is cut into chunks of `ch` rows so that a tile (ch*Q contiguous elements) stays
// within kOccCap; everything above is "hi".  A thread always owns the same in-tile
// positions p = tid + 256 u, so per element it only sets bit u of a 64-bit mask; the
// mixed-radix decode of p happens once per thread at the end.  Hi digits are uniform over
// a tile and advance by an odometer.
constexpr int kOccCap = 16384;   // 256 threads x 64 mask bits
constexpr int kOccMaxAxes = 16;

struct OccParams {
  int n, j;                 // axes; chunked axis (-1: the whole state is one tile)
  uint32_t Q, ch, nch, dj;  // lo product, chunk rows, chunks per axis j, dims[j]
  uint64_t ntiles, tiles_per_cta, stride;
  uint32_t dims[kOccMaxAxes];
};

template <typename V, bool DIAG>
__global__ void __launch_bounds__(256, 4)
k_occupancy(const V* __restrict__ x, const OccParams P, int* __restrict__ levels,
            double* __restrict__ partial) {
  __shared__ int s_lvl[kOccMaxAxes];
  __shared__ double s_scratch[8];
  __shared__ double s_max[8];
  const int tid = threadIdx.x;
  if (tid < kOccMaxAxes) s_lvl[tid] = -1;
  __syncthreads();

  const uint64_t t0 = (uint64_t)blockIdx.x * P.tiles_per_cta;
  uint64_t t1 = t0 + P.tiles_per_cta;
  if (t1 > P.ntiles) t1 = P.ntiles;

  // odometer over (hi digits..., chunk c)
  uint32_t dg[kOccMaxAxes];
  int lvl[kOccMaxAxes];
#pragma unroll
  for (int i = 0; i < kOccMaxAxes; ++i) { dg[i] = 0; lvl[i] = -1; }
  uint32_t c = (uint32_t)(t0 % P.nch);
  {
    uint64_t h = t0 / P.nch;
#pragma unroll
    for (int i = kOccMaxAxes - 1; i >= 0; --i)
      if (i < P.j) { dg[i] = (uint32_t)(h % P.dims[i]); h /= P.dims[i]; }
  }
  uint64_t hi_flat = t0 / P.nch;

  uint64_t mask_lo = 0, mask_c = 0;
  int cmax = -1;
  double sum = 0, mx = 0;
  const uint32_t rowQ = P.dj * P.Q;

  for (uint64_t t = t0; t < t1; ++t) {
    const uint64_t base = hi_flat * rowQ + (uint64_t)c * P.ch * P.Q;
    const uint32_t rows = (c + 1 == P.nch) ? (P.dj - c * P.ch) : P.ch;
    const uint32_t len = rows * P.Q;
    uint64_t m = 0;
    for (uint32_t u0 = 0; u0 * 256u < len; u0 += 8) {
      V v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t p = tid + 256u * (u0 + k);
        v[k] = p < len ? __ldcg(x + (base + p) * P.stride) : V{0, 0};
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double a2 = (double)v[k].x * v[k].x + (double)v[k].y * v[k].y;
        sum += DIAG ? (double)v[k].x : a2;
        mx = a2 > mx ? a2 : mx;
        m |= (uint64_t)(v[k].x != 0 || v[k].y != 0) << (u0 + k);
      }
    }
    if (m) {
      mask_lo |= m;
      if ((int)c > cmax) { cmax = (int)c; mask_c = m; }
      else if ((int)c == cmax) mask_c |= m;
#pragma unroll
      for (int i = 0; i < kOccMaxAxes; ++i)
        if (i < P.j) lvl[i] = max(lvl[i], (int)dg[i]);
    }
    // advance
    if (++c == P.nch) {
      c = 0;
      ++hi_flat;
      bool carry = true;
#pragma unroll
      for (int i = kOccMaxAxes - 1; i >= 0; --i)
        if (i < P.j && carry) {
          if (++dg[i] == P.dims[i]) dg[i] = 0; else carry = false;
        }
    }
  }

  // decode this thread's positions once
  if (mask_lo) {
#pragma unroll
    for (int i = 0; i < kOccMaxAxes; ++i)
      if (i < P.j && lvl[i] >= 0) atomicMax(&s_lvl[i], lvl[i]);
    int lo_lvl[kOccMaxAxes];
#pragma unroll
    for (int i = 0; i < kOccMaxAxes; ++i) lo_lvl[i] = -1;
    int row_best = -1;
    for (int u = 0; u < 64; ++u) {
      if (!((mask_lo >> u) & 1)) continue;
      const uint32_t p = tid + 256u * u;
      uint32_t q = p % P.Q;
      if ((mask_c >> u) & 1) row_best = max(row_best, (int)(p / P.Q));
#pragma unroll
      for (int i = kOccMaxAxes - 1; i >= 0; --i)
        if (i > P.j && i < P.n) {
          lo_lvl[i] = max(lo_lvl[i], (int)(q % P.dims[i]));
          q /= P.dims[i];
        }
    }
#pragma unroll
    for (int i = 0; i < kOccMaxAxes; ++i)
      if (i > P.j && i < P.n && lo_lvl[i] >= 0) atomicMax(&s_lvl[i], lo_lvl[i]);
    if (P.j >= 0 && row_best >= 0) {
      // chunk axis: the maximum over blocks combines (cmax, row) lexicographically
      atomicMax(&s_lvl[P.j], cmax * (int)P.ch + row_best);
    }
  }
  sum = block_sum(sum, s_scratch);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) s_max[tid >> 5] = mx;
  __syncthreads();
  if (tid < P.n && s_lvl[tid] >= 0) atomicMax(&levels[tid], s_lvl[tid]);
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) mx = fmax(mx, s_max[w]);
    partial[2 * blockIdx.x] = sum;
    partial[2 * blockIdx.x + 1] = mx;
  }
}

}  // namespace pw

using namespace pw;

#define PW_DISPATCH(dtype, FN, ...)                                   \
  ((dtype) == PW_C128 ? FN<double2>(__VA_ARGS__)                      \
                      : ((dtype) == PW_C64 ? FN<float2>(__VA_ARGS__) : PW_ERR_INVALID))

template <typename V>
static int probs_vec_t(const void* psi, void* probs, const Plan& p, cudaStream_t s) {
  return probs_impl<V>((const V*)psi, probs, p, 1, false, s);
}
template <typename V>
static int probs_mat_t(const void* rho, void* probs, const Plan& p, cudaStream_t s) {
  return probs_impl<V>((const V*)rho, probs, p, (int64_t)p.N + 1, true, s);
}

extern "C" int pw_probs_vec(const void* psi, void* probs, const int64_t* dims, int n,
                            const int32_t* targets, int nt, int dtype, void* stream) {
  if (!psi || !probs) return PW_ERR_INVALID;
  Plan p;
  PW_TRY(build_plan(dims, n, targets, nt, &p));
  return PW_DISPATCH(dtype, probs_vec_t, psi, probs, p, (cudaStream_t)stream);
}

extern "C" int pw_probs_mat(const void* rho, void* probs, const int64_t* dims, int n,
                            const int32_t* targets, int nt, int dtype, void* stream) {
  if (!rho || !probs) return PW_ERR_INVALID;
  Plan p;  // plan over the n-axis index; offsets are scaled by N+1 to walk the diagonal
  PW_TRY(build_plan(dims, n, targets, nt, &p));
  return PW_DISPATCH(dtype, probs_mat_t, rho, probs, p, (cudaStream_t)stream);
}

// ---- strided views: maps straight from (extent, stride) lists -----------------------------
static int strided_fibers(const int64_t* extent, const int64_t* stride, int n, FiberMap* fm) {
  if (n < 0 || (n && (!extent || !stride))) return PW_ERR_INVALID;
  fm->ng = 0;
  fm->nfib = 1;
  for (int i = 0; i < n; ++i) {
    if (extent[i] < 1) return PW_ERR_INVALID;
    if (extent[i] == 1) continue;
    if (fm->nfib > (1ull << 62) / (uint64_t)extent[i]) return PW_ERR_INVALID;
    fm->nfib *= (uint64_t)extent[i];
    // merge with the previous group when the two walk one contiguous run
    if (fm->ng > 0 && fm->gstride[fm->ng - 1] == stride[i] * extent[i]) {
      fm->gsize[fm->ng - 1] *= (uint64_t)extent[i];
      fm->gstride[fm->ng - 1] = stride[i];
      continue;
    }
    if (fm->ng >= kMaxGroups) return PW_ERR_UNSUPPORTED;
    fm->gsize[fm->ng] = (uint64_t)extent[i];
    fm->gstride[fm->ng] = stride[i];
    ++fm->ng;
  }
  return PW_OK;
}

static int strided_targets(const int64_t* extent, const int64_t* stride, int n, TargetMap* tm) {
  if (n < 0 || n > kMaxTargets || (n && (!extent || !stride))) return PW_ERR_INVALID;
  tm->nt = n;
  uint64_t t = 1;
  for (int i = 0; i < n; ++i) {
    if (extent[i] < 1) return PW_ERR_INVALID;
    tm->tdim[i] = (uint32_t)extent[i];
    tm->tstride[i] = stride[i];
    t *= (uint64_t)extent[i];
    if (t > 0xffffffffull) return PW_ERR_UNSUPPORTED;
  }
  tm->t = (uint32_t)t;
  return PW_OK;
}

template <typename V>
static int sums_strided_t(const void* x, double* sums, const Plan& p, int real_part, cudaStream_t s) {
  return probs_impl<V>((const V*)x, sums, p, 1, real_part != 0, s, true);
}

extern "C" int pw_sums_strided(const void* x, const int64_t* rest_extent, const int64_t* rest_stride,
                               int nrest, const int64_t* tgt_extent, const int64_t* tgt_stride,
                               int ntgt, int real_part, double* sums, int dtype, void* stream) {
  if (!x || !sums) return PW_ERR_INVALID;
  Plan p;
  PW_TRY(strided_fibers(rest_extent, rest_stride, nrest, &p.fm));
  PW_TRY(strided_targets(tgt_extent, tgt_stride, ntgt, &p.tm));
  p.N = p.fm.nfib * p.tm.t;
  return PW_DISPATCH(dtype, sums_strided_t, x, sums, p, real_part, (cudaStream_t)stream);
}

extern "C" int pw_probs_from_sums(const double* sums, int64_t t, void* probs, int dtype, void* stream) {
  if (!sums || !probs || t < 1 || t > 0xffffffffll) return PW_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PW_C128) k_probs_from_sums<double><<<1, 256, 0, s>>>(sums, (uint32_t)t, (double*)probs);
  else if (dtype == PW_C64) k_probs_from_sums<float><<<1, 256, 0, s>>>(sums, (uint32_t)t, (float*)probs);
  else return PW_ERR_INVALID;
  PW_LAUNCHED();
  return PW_OK;
}

template <typename V>
static int trace_strided_t(const void* x, void* out, const FiberMap& fm, const TargetMap& rows,
                           const TargetMap& cols, cudaStream_t s) {
  const uint64_t nout = (uint64_t)rows.t * cols.t;
  const int grid = (int)(nout < (uint64_t)kNumSMs * 16 ? nout : (uint64_t)kNumSMs * 16);
  k_trace_strided<V><<<grid, kRedBlock, 0, s>>>((const V*)x, fm, rows, cols, (V*)out);
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_trace_strided(const void* x, const int64_t* rest_extent, const int64_t* rest_stride,
                                int nrest, const int64_t* row_extent, const int64_t* row_stride, int nrow,
                                const int64_t* col_extent, const int64_t* col_stride, int ncol, void* out,
                                int dtype, void* stream) {
  if (!x || !out) return PW_ERR_INVALID;
  FiberMap fm;
  TargetMap rows, cols;
  PW_TRY(strided_fibers(rest_extent, rest_stride, nrest, &fm));
  PW_TRY(strided_targets(row_extent, row_stride, nrow, &rows));
  PW_TRY(strided_targets(col_extent, col_stride, ncol, &cols));
  if ((uint64_t)rows.t * cols.t > 0xffffffffull) return PW_ERR_UNSUPPORTED;
  return PW_DISPATCH(dtype, trace_strided_t, x, out, fm, rows, cols, (cudaStream_t)stream);
}

template <typename V>
static int gather_strided_t(const void* x, void* out, const FiberMap& fm, const void* probs,
                            const int64_t* outcome, int mode, cudaStream_t s) {
  using R = typename RealOf<V>::type;
  k_gather_strided<V><<<grid_for(fm.nfib, 256, 8), 256, 0, s>>>((const V*)x, (V*)out, fm, (const R*)probs,
                                                              outcome, mode);
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_gather_strided(const void* x, const int64_t* extent, const int64_t* stride, int naxes,
                                 void* out, const void* probs, const int64_t* outcome, int mode,
                                 int dtype, void* stream) {
  if (!x || !out || mode < 0 || mode > 2 || (mode && (!probs || !outcome))) return PW_ERR_INVALID;
  FiberMap fm;
  PW_TRY(strided_fibers(extent, stride, naxes, &fm));
  return PW_DISPATCH(dtype, gather_strided_t, x, out, fm, probs, outcome, mode, (cudaStream_t)stream);
}

template <typename V>
static int cross_reduced_t(const void* x, const void* y, void* out, const Plan& p, cudaStream_t s) {
  const uint64_t nout = (uint64_t)p.tm.t * p.tm.t;
  const int grid = (int)(nout < (uint64_t)kNumSMs * 16 ? nout : (uint64_t)kNumSMs * 16);
  k_cross_reduced<V><<<grid, kRedBlock, 0, s>>>((const V*)x, (const V*)y, p.fm, p.tm, (V*)out);
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_cross_reduced(const void* x, const void* y, void* out, const int64_t* axes, int naxes,
                                const int32_t* keep, int nkeep, int dtype, void* stream) {
  if (!x || !y || !out) return PW_ERR_INVALID;
  Plan p;
  int tg[kMaxTargets];
  if (nkeep < 1 || nkeep > kMaxTargets || !keep) return PW_ERR_INVALID;
  for (int i = 0; i < nkeep; ++i) tg[i] = keep[i];
  PW_TRY(build_plan_axes(axes, naxes, tg, nkeep, &p));
  if ((uint64_t)p.tm.t * p.tm.t > 0xffffffffull) return PW_ERR_UNSUPPORTED;
  return PW_DISPATCH(dtype, cross_reduced_t, x, y, out, p, (cudaStream_t)stream);
}

template <typename V>
static int reduced_t(const void* psi, void* out, const Plan& p, cudaStream_t s) {
  return reduced_impl<V>((const V*)psi, (V*)out, p, s);
}

extern "C" int pw_reduced_from_vec(const void* psi, void* out, const int64_t* dims, int n,
                                   const int32_t* keep, int nkeep, int dtype, void* stream) {
  if (!psi || !out) return PW_ERR_INVALID;
  Plan p;
  PW_TRY(build_plan(dims, n, keep, nkeep, &p));
  return PW_DISPATCH(dtype, reduced_t, psi, out, p, (cudaStream_t)stream);
}

template <typename V>
static int trace_out_t(const void* rho, void* out, const Plan& p, cudaStream_t s) {
  const uint64_t nout = (uint64_t)p.tm.t * p.tm.t;
  if (p.fm.nfib >= 64) {
    int grid = (int)(nout < (uint64_t)kNumSMs * 16 ? nout : (uint64_t)kNumSMs * 16);
    k_trace_out<V, true><<<grid, kRedBlock, 0, s>>>((const V*)rho, p.fm, p.tm, (int64_t)p.N, (V*)out);
  } else {
    int grid = grid_for(nout, kRedBlock, 16);
    k_trace_out<V, false><<<grid, kRedBlock, 0, s>>>((const V*)rho, p.fm, p.tm, (int64_t)p.N, (V*)out);
  }
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_trace_out_mat(const void* rho, void* out, const int64_t* dims, int n,
                                const int32_t* keep, int nkeep, int dtype, void* stream) {
  if (!rho || !out) return PW_ERR_INVALID;
  Plan p;
  PW_TRY(build_plan(dims, n, keep, nkeep, &p));
  if ((uint64_t)p.tm.t * p.tm.t > 0xffffffffull) return PW_ERR_UNSUPPORTED;
  return PW_DISPATCH(dtype, trace_out_t, rho, out, p, (cudaStream_t)stream);
}

template <typename V>
static int draw_t(const void* probs, int64_t t, uint32_t k0, uint32_t k1, const uint32_t* key_dev,
                  int64_t* outcome, cudaStream_t s) {
  using R = typename RealOf<V>::type;
  const int exact = options().draw_exact.load();
  if (t <= kDrawChunk) {
    k_draw<R><<<1, kDrawBlock, 0, s>>>((const R*)probs, t, nullptr, 1, k0, k1, key_dev, exact, outcome);
    PW_LAUNCHED();
    return PW_OK;
  }
  const int64_t nchunks = (t + kDrawChunk - 1) / kDrawChunk;
  Scratch cs;
  PW_TRY(cs.alloc(sizeof(double) * nchunks, s));
  k_draw_chunks<R><<<(unsigned)nchunks, kDrawBlock, 0, s>>>((const R*)probs, t, (double*)cs.p);
  PW_LAUNCHED();
  k_draw<R><<<1, kDrawBlock, 0, s>>>((const R*)probs, t, (const double*)cs.p, nchunks, k0, k1, key_dev,
                                    exact, outcome);
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_draw(const void* probs, int64_t t, uint32_t key0, uint32_t key1, int dtype,
                       int64_t* outcome, void* stream) {
  if (!probs || !outcome || t < 1) return PW_ERR_INVALID;
  return PW_DISPATCH(dtype, draw_t, probs, t, key0, key1, nullptr, outcome, (cudaStream_t)stream);
}

extern "C" int pw_draw_devkey(const void* probs, int64_t t, const uint32_t* key_dev, int dtype,
                              int64_t* outcome, void* stream) {
  if (!probs || !outcome || !key_dev || t < 1) return PW_ERR_INVALID;
  return PW_DISPATCH(dtype, draw_t, probs, t, 0u, 0u, key_dev, outcome, (cudaStream_t)stream);
}

template <typename V>
static int collapse_vec_t(const void* psi, void* post, const int64_t* outcome, const void* probs,
                          const Plan& p, cudaStream_t s) {
  using R = typename RealOf<V>::type;
  k_collapse_vec<V><<<grid_for(p.fm.nfib, 256, 8), 256, 0, s>>>((const V*)psi, (V*)post, outcome,
                                                               (const R*)probs, p.fm, p.tm);
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_collapse_vec(const void* psi, void* post, const int64_t* outcome,
                               const void* probs, const int64_t* dims, int n,
                               const int32_t* targets, int nt, int dtype, void* stream) {
  if (!psi || !post || !outcome) return PW_ERR_INVALID;
  Plan p;
  PW_TRY(build_plan(dims, n, targets, nt, &p));
  return PW_DISPATCH(dtype, collapse_vec_t, psi, post, outcome, probs, p, (cudaStream_t)stream);
}

template <typename V>
static int collapse_mat_t(const void* rho, void* post, const int64_t* outcome, const void* probs,
                          const Plan& p, cudaStream_t s) {
  using R = typename RealOf<V>::type;
  k_collapse_mat<V><<<grid_for(p.fm.nfib * p.fm.nfib, 256, 8), 256, 0, s>>>(
      (const V*)rho, (V*)post, outcome, (const R*)probs, p.fm, p.tm, (int64_t)p.N);
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_collapse_mat(const void* rho, void* post, const int64_t* outcome,
                               const void* probs, const int64_t* dims, int n,
                               const int32_t* targets, int nt, int dtype, void* stream) {
  if (!rho || !post || !outcome) return PW_ERR_INVALID;
  Plan p;
  PW_TRY(build_plan(dims, n, targets, nt, &p));
  return PW_DISPATCH(dtype, collapse_mat_t, rho, post, outcome, probs, p, (cudaStream_t)stream);
}

template <typename V>
static int povm_probs_t(const void* rho_t, const void* ops, int K, int64_t t, void* probs,
                        void* traces, cudaStream_t s) {
  using R = typename RealOf<V>::type;
  Scratch tr;
  PW_TRY(tr.alloc(sizeof(double) * K, s));
  k_povm_traces<V><<<K, kRedBlock, 0, s>>>((const V*)rho_t, (const V*)ops, t, (double*)tr.p);
  PW_LAUNCHED();
  k_povm_normalise<R><<<1, 32, 0, s>>>((const double*)tr.p, K, (R*)probs, (R*)traces);
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_povm_probs(const void* rho_t, const void* ops, int K, int64_t t, void* probs,
                             void* traces, int dtype, void* stream) {
  if (!rho_t || !ops || !probs || !traces || K < 1 || t < 1) return PW_ERR_INVALID;
  return PW_DISPATCH(dtype, povm_probs_t, rho_t, ops, K, t, probs, traces, (cudaStream_t)stream);
}

template <typename V>
static int povm_select_t(const void* ops, int64_t t, const int64_t* outcome, const void* traces,
                         void* op_out, cudaStream_t s) {
  using R = typename RealOf<V>::type;
  k_povm_select<V><<<grid_for((uint64_t)t * t, 128, 1), 128, 0, s>>>((const V*)ops, t, outcome,
                                                                    (const R*)traces, (V*)op_out);
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_povm_select(const void* ops, int64_t t, const int64_t* outcome,
                              const void* traces, void* op_out, int dtype, void* stream) {
  if (!ops || !outcome || !traces || !op_out || t < 1) return PW_ERR_INVALID;
  return PW_DISPATCH(dtype, povm_select_t, ops, t, outcome, traces, op_out, (cudaStream_t)stream);
}

template <typename V>
static int norm_stats_t(const void* x, int64_t count, double* stats, cudaStream_t s) {
  const int nblocks = grid_for((uint64_t)count, 256, 4);
  Scratch partial;
  PW_TRY(partial.alloc(sizeof(double) * 2 * nblocks, s));
  k_norm_partial<V><<<nblocks, 256, 0, s>>>((const V*)x, count, (double*)partial.p);
  PW_LAUNCHED();
  k_norm_finalize<<<1, 32, 0, s>>>((const double*)partial.p, nblocks, stats);
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_norm_stats(const void* x, int64_t count, int dtype, double* stats, void* stream) {
  if (!x || !stats || count < 0) return PW_ERR_INVALID;
  return PW_DISPATCH(dtype, norm_stats_t, x, count, stats, (cudaStream_t)stream);
}

template <typename V>
static int occupancy_t(const void* x, const OccParams& P, int nblocks, int is_matrix,
                       int32_t* levels, double* stats, cudaStream_t s) {
  Scratch partial;
  PW_TRY(partial.alloc(sizeof(double) * 2 * nblocks, s));
  PW_CUDA(cudaMemsetAsync(levels, 0xFF, sizeof(int32_t) * P.n, s));  // all -1
  if (is_matrix)
    k_occupancy<V, true><<<nblocks, 256, 0, s>>>((const V*)x, P, levels, (double*)partial.p);
  else
    k_occupancy<V, false><<<nblocks, 256, 0, s>>>((const V*)x, P, levels, (double*)partial.p);
  PW_LAUNCHED();
  k_norm_finalize<<<1, 32, 0, s>>>((const double*)partial.p, nblocks, stats);
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_occupancy(const void* state, const int64_t* dims, int n, int is_matrix,
                            int dtype, int32_t* levels, double* stats, void* stream) {
  if (!state || !dims || !levels || !stats || n < 1) return PW_ERR_INVALID;
  if (n > kOccMaxAxes) return PW_ERR_UNSUPPORTED;
  OccParams P{};
  uint64_t N = 1;
  for (int i = 0; i < n; ++i) {
    if (dims[i] < 1 || dims[i] > 0x7fffffff) return PW_ERR_INVALID;
    P.dims[i] = (uint32_t)dims[i];
    N *= (uint64_t)dims[i];
  }
  P.n = n;
  uint64_t Q = 1;
  int k = n;  // axes [k, n) are lo
  while (k > 0 && Q * (uint64_t)dims[k - 1] <= (uint64_t)kOccCap) Q *= (uint64_t)dims[--k];
  P.j = k - 1;
  P.Q = (uint32_t)Q;
  if (P.j >= 0) {
    P.dj = P.dims[P.j];
    P.ch = (uint32_t)std::min<uint64_t>(P.dj, (uint64_t)kOccCap / Q);
    P.nch = (P.dj + P.ch - 1) / P.ch;
  } else {
    P.dj = P.ch = P.nch = 1;
  }
  P.ntiles = N / ((uint64_t)P.dj * Q) * P.nch;
  P.stride = is_matrix ? N + 1 : 1;
  const uint64_t want = std::min<uint64_t>(P.ntiles, (uint64_t)kNumSMs * 4);  // one resident wave
  P.tiles_per_cta = (P.ntiles + want - 1) / want;
  const int nblocks = (int)((P.ntiles + P.tiles_per_cta - 1) / P.tiles_per_cta);
  return PW_DISPATCH(dtype, occupancy_t, state, P, nblocks, is_matrix, levels, stats,
                     (cudaStream_t)stream);
}

template <typename V>
static int scale_t(void* x, int64_t count, const double* scale, cudaStream_t s) {
  k_scale<V, false><<<grid_for((uint64_t)count, 256, 8), 256, 0, s>>>((V*)x, count, scale);
  PW_LAUNCHED();
  return PW_OK;
}

template <typename V>
static int divide_t(void* x, int64_t count, const double* divisor, cudaStream_t s) {
  k_scale<V, true><<<grid_for((uint64_t)count, 256, 8), 256, 0, s>>>((V*)x, count, divisor);
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_scale(void* x, int64_t count, const double* scale_dev, int dtype, void* stream) {
  if (!x || !scale_dev || count < 0) return PW_ERR_INVALID;
  return PW_DISPATCH(dtype, scale_t, x, count, scale_dev, (cudaStream_t)stream);
}

extern "C" int pw_divide(void* x, int64_t count, const double* divisor_dev, int dtype, void* stream) {
  if (!x || !divisor_dev || count < 0) return PW_ERR_INVALID;
  return PW_DISPATCH(dtype, divide_t, x, count, divisor_dev, (cudaStream_t)stream);
}
