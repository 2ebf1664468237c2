// Fused multi-step evolution of a SMALL state whose subsystems are all targeted
// (the Jaynes-Cummings loop, examples/jaynes_cummings_model.py:130-141: one
// apply_operation of the same dense propagator plus one trace_out per step, 10^4 steps).
// Through the per-call API every step costs host launches; here ONE kernel keeps the
// propagator distributed over the registers of a 1024-thread CTA, the state in shared
// memory, and emits the reduced density matrix of the kept subsystems after every step.
//
// Also: a row-parallel dense apply for launch-latency-bound states (t > 8, few fibers).
#include <cooperative_groups.h>

#include "pw_internal.cuh"

namespace cg = cooperative_groups;

namespace pw {

constexpr int kEvoMaxN = 128;   // state dimension handled by the fused kernel
constexpr int kEvoRows = 32;    // output rows (= warps) per CTA
constexpr int kEvoSeg = kEvoMaxN / 32;

// A thread-block CLUSTER of ceil(N/32) CTAs shares the work: CTA c owns rows
// [32c, 32c+32), one warp per row, lane l holds U[row, l + 32 j] in registers.  Every CTA
// keeps the full state in its shared memory; after each step the owners write their new
// amplitudes into EVERY CTA's buffer through distributed shared memory and the cluster
// synchronises (one barrier per step, no global memory traffic inside the loop).
template <typename V>
__global__ void __launch_bounds__(1024, 1)
k_evolve(const V* __restrict__ psi0, V* __restrict__ psi_out, const V* __restrict__ U, const int N,
         const int nsteps, const FiberMap fm, const TargetMap tm, V* __restrict__ reduced_out) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = cluster.block_rank(), csize = cluster.num_blocks();
  __shared__ V xs[2][kEvoMaxN];
  __shared__ int s_koff[64];
  __shared__ int s_fbase[kEvoMaxN];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a = (int)crank * kEvoRows + warp;  // my output row
  V u[kEvoSeg];
#pragma unroll
  for (int j = 0; j < kEvoSeg; ++j) {
    const int b = lane + 32 * j;
    u[j] = (a < N && b < N) ? U[(size_t)a * N + b] : czero<V>();
  }
  const int tk = (int)tm.t, nfib = (int)fm.nfib;
  const bool observer = reduced_out != nullptr && crank == 0;
  if (observer) {
    for (int i = tid; i < tk; i += blockDim.x) s_koff[i] = (int)tm.offset(i);
    for (int f = tid; f < nfib; f += blockDim.x) s_fbase[f] = (int)fm.base(f);
  }
  for (int i = tid; i < N; i += blockDim.x) xs[0][i] = psi0[i];
  cluster.sync();
  int cur = 0;
  for (int step = 0; step < nsteps; ++step) {
    V acc = czero<V>();
#pragma unroll
    for (int j = 0; j < kEvoSeg; ++j) {
      const int b = lane + 32 * j;
      if (b < N) cfma(acc, u[j], xs[cur][b]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc.x += __shfl_down_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_down_sync(0xffffffffu, acc.y, o);
    }
    acc.x = __shfl_sync(0xffffffffu, acc.x, 0);
    acc.y = __shfl_sync(0xffffffffu, acc.y, 0);
    // lane r < csize publishes the row's new amplitude into CTA r's next-state buffer
    if (a < N && lane < (int)csize) {
      V* remote = cluster.map_shared_rank(&xs[cur ^ 1][0], lane);
      remote[a] = acc;
    }
    cluster.sync();
    cur ^= 1;
    if (observer) {
      // rho_T[i, j] = sum_f x[f, i] conj(x[f, j]) of the NEW state; one warp per entry
      const int nwarps = blockDim.x >> 5;
      for (int e = warp; e < tk * tk; e += nwarps) {
        const int i = e / tk, j = e - i * tk;
        double re = 0, im = 0;
        for (int f = lane; f < nfib; f += 32) {
          const V xi = xs[cur][s_fbase[f] + s_koff[i]], xj = xs[cur][s_fbase[f] + s_koff[j]];
          re += (double)xi.x * xj.x + (double)xi.y * xj.y;
          im += (double)xi.y * xj.x - (double)xi.x * xj.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          re += __shfl_down_sync(0xffffffffu, re, o);
          im += __shfl_down_sync(0xffffffffu, im, o);
        }
        if (lane == 0) {
          V v; v.x = re; v.y = im;
          reduced_out[(size_t)step * tk * tk + e] = v;
        }
      }
    }
  }
  if (crank == 0)
    for (int i = tid; i < N; i += blockDim.x) psi_out[i] = xs[cur][i];
}

// Row-parallel dense apply for small states: block = one fiber (staged in shared memory),
// threads = output rows.  A handful of microseconds for a 128 x 128 operator.
template <typename V>
__global__ void __launch_bounds__(256)
k_apply_rows(const V* __restrict__ in, V* __restrict__ out, const V* __restrict__ op,
             const FiberMap fm, const TargetMap tm, const bool conj_op) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  V* sx = (V*)s_raw;
  const uint32_t t = tm.t;
  for (uint64_t f = blockIdx.x; f < fm.nfib; f += gridDim.x) {
    const int64_t base = fm.base(f);
    for (uint32_t b = threadIdx.x; b < t; b += blockDim.x) sx[b] = in[base + tm.offset(b)];
    __syncthreads();
    for (uint32_t a = threadIdx.x; a < t; a += blockDim.x) {
      V acc = czero<V>();
      const V* row = op + (size_t)a * t;
      for (uint32_t b = 0; b < t; ++b) {
        V o = row[b];
        if (conj_op) o.y = -o.y;
        cfma(acc, o, sx[b]);
      }
      out[base + tm.offset(a)] = acc;
    }
    __syncthreads();
  }
}

// The same for a HANDFUL of fibers (one Jaynes-Cummings state: a single fiber): one WARP per output
// row -- coalesced operator rows, a shuffle reduction -- and 8 rows per block, so a 128 x 128
// operator on one state runs on 16 blocks instead of on 128 threads of one (16 -> ~3 us).
template <typename V>
__global__ void __launch_bounds__(256)
k_apply_rows_warp(const V* __restrict__ in, V* __restrict__ out, const V* __restrict__ op,
                  const FiberMap fm, const TargetMap tm, const bool conj_op) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  V* sx = (V*)s_raw;
  const uint32_t t = tm.t;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t base = fm.base(blockIdx.y);
  for (uint32_t b = threadIdx.x; b < t; b += blockDim.x) sx[b] = in[base + tm.offset(b)];
  __syncthreads();
  const uint32_t a = blockIdx.x * 8 + warp;
  if (a >= t) return;
  V acc = czero<V>();
  const V* row = op + (size_t)a * t;
  for (uint32_t b = lane; b < t; b += 32) {
    V o = row[b];
    if (conj_op) o.y = -o.y;
    cfma(acc, o, sx[b]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc.x += __shfl_down_sync(0xffffffffu, acc.x, o);
    acc.y += __shfl_down_sync(0xffffffffu, acc.y, o);
  }
  if (lane == 0) out[base + tm.offset(a)] = acc;
}

int launch_apply_rows(const void* in, void* out, const void* op, const Plan& p, int dtype,
                      bool conj_op, cudaStream_t s) {
  const size_t es = dtype == PW_C128 ? 16 : 8;
  const size_t smem = (size_t)p.tm.t * es;
  if (smem > 48 * 1024 || in == out) return PW_ERR_UNSUPPORTED;
  if (p.fm.nfib <= 64 && p.tm.t >= 32) {
    const dim3 grid2((p.tm.t + 7) / 8, (unsigned)p.fm.nfib);
    if (dtype == PW_C128)
      k_apply_rows_warp<double2><<<grid2, 256, smem, s>>>((const double2*)in, (double2*)out, (const double2*)op, p.fm, p.tm, conj_op);
    else
      k_apply_rows_warp<float2><<<grid2, 256, smem, s>>>((const float2*)in, (float2*)out, (const float2*)op, p.fm, p.tm, conj_op);
    g_paths[1].fetch_add(1, std::memory_order_relaxed);
    PW_LAUNCHED();
    return PW_OK;
  }
  const int grid = (int)(p.fm.nfib < (uint64_t)kNumSMs * 8 ? p.fm.nfib : (uint64_t)kNumSMs * 8);
  if (dtype == PW_C128)
    k_apply_rows<double2><<<grid, 256, smem, s>>>((const double2*)in, (double2*)out, (const double2*)op, p.fm, p.tm, conj_op);
  else
    k_apply_rows<float2><<<grid, 256, smem, s>>>((const float2*)in, (float2*)out, (const float2*)op, p.fm, p.tm, conj_op);
  g_paths[1].fetch_add(1, std::memory_order_relaxed);
  PW_LAUNCHED();
  return PW_OK;
}

}  // namespace pw

using namespace pw;

extern "C" int pw_evolve_vec(const void* psi, void* out, const void* op, const int64_t* dims, int n,
                             const int32_t* keep, int nkeep, int nsteps, void* reduced_out,
                             int dtype, void* stream) {
  if (!psi || !out || !op || !dims || n < 1 || nsteps < 0) return PW_ERR_INVALID;
  Plan p;
  PW_TRY(build_plan(dims, n, keep, reduced_out ? nkeep : 0, &p));
  if (p.N > (uint64_t)kEvoMaxN) return PW_ERR_UNSUPPORTED;
  if (reduced_out && p.tm.t > 8) return PW_ERR_UNSUPPORTED;
  const int N = (int)p.N;
  const unsigned csize = (unsigned)((N + kEvoRows - 1) / kEvoRows);  // 1..4 CTAs in one cluster
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize);
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (dtype == PW_C128) {
    PW_CUDA(cudaLaunchKernelEx(&cfg, k_evolve<double2>, (const double2*)psi, (double2*)out,
                               (const double2*)op, N, nsteps, p.fm, p.tm, (double2*)reduced_out));
  } else if (dtype == PW_C64) {
    PW_CUDA(cudaLaunchKernelEx(&cfg, k_evolve<float2>, (const float2*)psi, (float2*)out,
                               (const float2*)op, N, nsteps, p.fm, p.tm, (float2*)reduced_out));
  } else {
    return PW_ERR_INVALID;
  }
  PW_LAUNCHED();
  return PW_OK;
}
