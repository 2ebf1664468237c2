// Several single-subsystem operators on DIFFERENT axes in ONE HBM pass (they commute).
//
// The bench layer applies a single-mode operator to every mode: twelve passes over 34.8 GB where
// the arithmetic (6 complex FMAs per amplitude) hides completely behind the memory system.  A
// tile kernel keeps the sub-tensor spanned by k target axes resident in shared memory and applies
// the k operators one after the other, so the state crosses HBM once per k operators:
//
//   tile[j][c]: j = joint index of the k target axes (t = d^k), c = C consecutive fibers of the
//   innermost untouched run (so a row of the tile is a contiguous piece of >= 64 bytes of memory;
//   when the innermost axis of the tensor is itself a target, j's last digit is the contiguous
//   direction instead and the tile is loaded j-fastest).
//   load (cp.async; tiles are single-buffered, several small CTAs per SM overlap each other's
//   phases) -> k in-place passes, one operator each, a thread owns F whole fibers of that operator
//   (D loads each, D*D complex FMAs against the operator broadcast from shared memory, D stores;
//   __syncthreads between passes) -> coalesced store.
//
// Per amplitude and operator: 2 shared-memory accesses and D complex FMAs; the k = 3, d = 6 case
// needs 144 real flops per 32 bytes of HBM traffic, i.e. it sits near the corner of the HBM and
// FP64 roofs (8.6 ms of FP64 work against a 10.7 ms pass for the 34.8 GB vector).
// Operators must have one common dimension D (2..8); anything else returns PW_ERR_UNSUPPORTED
// and the caller applies them one by one.
#include "pw_internal.cuh"

namespace pw {

constexpr int kMuMaxOps = 4;

struct MultiGeom {
  int nops;
  int t;             // D^nops
  int C;             // fibers per tile
  int tile;          // t * C elements
  int st[kMuMaxOps]; // in-tile element stride of operator i's digit
  int64_t cstride;   // element stride between consecutive fibers of a chunk
  int j_fastest;     // load / store order: 1 = joint index fastest (innermost axis is a target)
  int logC;          // C == 1 << logC, or -1
  uint32_t inv_c, inv_t, inv_st[kMuMaxOps];  // ceil(2^32 / x): x / y for x < 2^16 is umulhi(x, inv)
  uint64_t units;    // nfib / C
  const void* op[kMuMaxOps];  // device operator of digit i (D x D, row-major)
};

__device__ __forceinline__ void mu_cp16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mu_cp8(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mu_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void mu_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// dynamic smem: ops [nops][D*D] | offs [t] (int64) | tile [tile]
// F fibers of a pass per thread: every operator element read from shared memory (a broadcast
// LDS.128 costs the load/store unit four cycles per warp whatever the data) is used for F
// fibers, which is what keeps the kernel off the shared-memory instruction roof: with one fiber
// per thread the 48 LDS/STS per 144 DFMAs bound it at 25 ms for three operators (measured);
// F = 3 needs 24.  Tiles are single-buffered: several small CTAs per SM overlap each other's
// load / compute / store phases (a second buffer per CTA with the next tile's cp.async in flight
// during the passes measured SLOWER: 28.3 against 23.6 ms for three operators on the same box).
// CTA size / resident CTAs per fibers-per-thread: F fibers of D complex values live in registers
constexpr int mu_threads(int F) { return F == 1 ? 384 : (F == 2 ? 192 : (F == 3 ? 128 : 96)); }
constexpr int mu_ctas(int F) { return F == 1 ? 2 : (F == 2 ? 3 : 4); }  // (F = 3 at 5 CTAs spills: 28 ms against 22.9)

template <typename V, int D, int F>
__global__ void __launch_bounds__(mu_threads(F), mu_ctas(F))
k_apply_multi(const V* in, V* out, const FiberMap fm, const TargetMap tm, const MultiGeom g) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  V* s_ops = (V*)s_raw;
  int64_t* s_off = (int64_t*)(s_ops + ((g.nops * D * D + 1) & ~1));
  V* tile = (V*)(s_off + ((g.t + 1) & ~1));
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < g.nops * D * D; i += nthr) s_ops[i] = ((const V*)g.op[i / (D * D)])[i % (D * D)];
  for (int j = tid; j < g.t; j += nthr) s_off[j] = tm.offset((uint32_t)j);
  __syncthreads();
  // tile slot e -> (joint index j, fiber c) without integer division
  auto split = [&](int e, int& j, int& c) {
    if (g.j_fastest) { c = (int)__umulhi((uint32_t)e, g.inv_t); j = e - c * g.t; }
    else if (g.logC >= 0) { j = e >> g.logC; c = e & (g.C - 1); }
    else { j = (int)__umulhi((uint32_t)e, g.inv_c); c = e - j * g.C; }
  };
  const int nf = g.tile / D;  // fibers per pass
  for (uint64_t u = blockIdx.x; u < g.units; u += gridDim.x) {
    const int64_t base = fm.base(u * (uint64_t)g.C);
    const V* src = in + base;
    for (int e = tid; e < g.tile; e += nthr) {
      int j, c;
      split(e, j, c);
      const V* p = src + s_off[j] + (int64_t)c * g.cstride;
      if (sizeof(V) == 16) mu_cp16(tile + j * g.C + c, p); else mu_cp8(tile + j * g.C + c, p);
    }
    mu_commit();
    mu_wait<0>();
    __syncthreads();
    for (int i = 0; i < g.nops; ++i) {
      const int st = g.st[i];
      int zero = 0;
      asm volatile("" : "+r"(zero));  // keep the operator in shared memory (broadcast loads)
      const V* sop = s_ops + i * D * D + zero;
      for (int q0 = tid * F; q0 < nf; q0 += nthr * F) {
        V* x0[F];
        V x[F][D];
        bool live[F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
          live[f] = q0 + f < nf;                 // ragged tail: idle slots touch nothing
          const int q = live[f] ? q0 + f : q0;
          const int hi = st == 1 ? q : (int)__umulhi((uint32_t)q, g.inv_st[i]), lo = q - hi * st;  // (2^32 / 1 overflows)
          x0[f] = tile + hi * st * D + lo;
#pragma unroll
          for (int b = 0; b < D; ++b) x[f][b] = x0[f][b * st];
        }
#pragma unroll
        for (int a = 0; a < D; ++a) {
          V acc[F];
#pragma unroll
          for (int f = 0; f < F; ++f) acc[f] = czero<V>();
#pragma unroll
          for (int b = 0; b < D; ++b) {
            const V o = sop[a * D + b];
#pragma unroll
            for (int f = 0; f < F; ++f) cfma(acc[f], o, x[f][b]);
          }
#pragma unroll
          for (int f = 0; f < F; ++f)
            if (f == 0 || live[f]) x0[f][a * st] = acc[f];
        }
      }
      __syncthreads();
    }
    V* dst = out + base;
    for (int e = tid; e < g.tile; e += nthr) {
      int j, c;
      split(e, j, c);
      __stcg(dst + s_off[j] + (int64_t)c * g.cstride, tile[j * g.C + c]);
    }
    __syncthreads();
  }
}

template <typename V, int D, int F>
static int launch_multi_f(const V* in, V* out, const Plan& p, const MultiGeom& g, cudaStream_t s) {
  const size_t smem = sizeof(V) * (size_t)((g.nops * D * D + 1) & ~1) + sizeof(int64_t) * (size_t)((g.t + 1) & ~1) +
                      sizeof(V) * (size_t)g.tile;
  if (smem > 100 * 1024) return PW_ERR_UNSUPPORTED;
  static bool attr = false;
  if (!attr) {
    PW_CUDA(cudaFuncSetAttribute(k_apply_multi<V, D, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  const int nf = g.tile / D;
  int threads = (((nf + F - 1) / F + 31) / 32) * 32;
  const int tmax = mu_threads(F);
  if (threads > tmax) threads = tmax;
  if (threads < 32) threads = 32;
  int ctas = (int)((220 * 1024) / (smem + 1024));   // resident CTAs per SM by shared memory
  const int cmax = mu_ctas(F);
  if (ctas > cmax) ctas = cmax;
  if (ctas < 1) ctas = 1;
  uint64_t grid = (uint64_t)kNumSMs * ctas;
  if (grid > g.units) grid = g.units;
  k_apply_multi<V, D, F><<<(unsigned)grid, threads, smem, s>>>(in, out, p.fm, p.tm, g);
  PW_LAUNCHED();
  return PW_OK;
}

template <typename V, int D>
static int launch_multi_d(const V* in, V* out, const Plan& p, const MultiGeom& g, cudaStream_t s) {
  // fibers per thread: 3 (24 regs of state per fiber in complex128) unless the operators are
  // large; knob "multi_f" overrides (1..4)
  int F = options().multi_f.load(std::memory_order_relaxed);
  if (F <= 0) F = D <= 6 ? 3 : 2;
  switch (F) {
    case 1: return launch_multi_f<V, D, 1>(in, out, p, g, s);
    case 2: return launch_multi_f<V, D, 2>(in, out, p, g, s);
    case 3: return launch_multi_f<V, D, 3>(in, out, p, g, s);
    default: return launch_multi_f<V, D, 4>(in, out, p, g, s);
  }
}

template <typename V>
static int launch_multi_t(const V* in, V* out, const Plan& p, const MultiGeom& g, int d, cudaStream_t s) {
  switch (d) {
    case 2: return launch_multi_d<V, 2>(in, out, p, g, s);
    case 3: return launch_multi_d<V, 3>(in, out, p, g, s);
    case 4: return launch_multi_d<V, 4>(in, out, p, g, s);
    case 5: return launch_multi_d<V, 5>(in, out, p, g, s);
    case 6: return launch_multi_d<V, 6>(in, out, p, g, s);
    case 7: return launch_multi_d<V, 7>(in, out, p, g, s);
    case 8: return launch_multi_d<V, 8>(in, out, p, g, s);
    default: return PW_ERR_UNSUPPORTED;
  }
}

}  // namespace pw

using namespace pw;

// `ops`: HOST array of nops DEVICE pointers (D x D operators, row-major), ops[i] acts on axes[i].
// The operators commute (different axes), so they are applied in ascending axis order: the last
// digit of the tile's joint index is then the target with the smallest stride.  `out` may alias
// `psi` (tiles are disjoint).
extern "C" int pw_apply_vec_multi(const void* psi, void* out, const void* const* ops, const int64_t* dims,
                                  int n, const int32_t* axes, int nops, int dtype, void* stream) {
  if (!psi || !out || !ops || !dims || !axes || n < 1 || n > PW_MAX_SUBSYSTEMS) return PW_ERR_INVALID;
  if (dtype != PW_C64 && dtype != PW_C128) return PW_ERR_INVALID;
  if (nops < 2 || nops > kMuMaxOps) return PW_ERR_UNSUPPORTED;
  int order[kMuMaxOps];
  for (int i = 0; i < nops; ++i) {
    if (axes[i] < 0 || axes[i] >= n || !ops[i]) return PW_ERR_INVALID;
    order[i] = i;
  }
  for (int i = 1; i < nops; ++i)  // insertion sort by axis
    for (int k = i; k > 0 && axes[order[k]] < axes[order[k - 1]]; --k) { int t = order[k]; order[k] = order[k - 1]; order[k - 1] = t; }
  int32_t sorted[kMuMaxOps];
  for (int i = 0; i < nops; ++i) sorted[i] = axes[order[i]];
  const int64_t d = dims[sorted[0]];
  if (d < 2 || d > 8) return PW_ERR_UNSUPPORTED;
  for (int i = 1; i < nops; ++i)
    if (dims[sorted[i]] != d) return PW_ERR_UNSUPPORTED;
  Plan p;
  PW_TRY(build_plan(dims, n, sorted, nops, &p));  // rejects duplicate axes
  MultiGeom g;
  g.nops = nops;
  g.t = (int)p.tm.t;
  for (int i = 0; i < nops; ++i) g.op[i] = ops[order[i]];
  // fibers per tile: consecutive fibers of the innermost untouched run; a tile of ~28 KB (c128)
  const size_t es = dtype == PW_C128 ? 16 : 8;
  const int cmax = (int)((28 * 1024) / (es * (size_t)g.t));
  if (cmax < 1) return PW_ERR_UNSUPPORTED;
  int C = 1;
  int64_t cstride = 1;
  const bool inner_is_target = sorted[nops - 1] == n - 1;
  if (p.fm.ng > 0) {
    const uint64_t run = p.fm.gsize[p.fm.ng - 1];
    for (int c = cmax; c >= 1; --c)
      if (run % (uint64_t)c == 0) { C = c; break; }
    cstride = p.fm.gstride[p.fm.ng - 1];
  }
  // coalescing: a row of the tile must be a contiguous piece of >= 64 bytes (>= 32 when the
  // contiguous direction is the innermost target itself)
  if (!inner_is_target && (cstride != 1 || (size_t)C * es < 64)) return PW_ERR_UNSUPPORTED;
  if (inner_is_target && (size_t)d * es < 32) return PW_ERR_UNSUPPORTED;
  g.C = C;
  g.tile = g.t * C;
  g.cstride = cstride;
  g.j_fastest = inner_is_target ? 1 : 0;
  g.units = p.fm.nfib / (uint64_t)C;
  int st = C;
  for (int i = nops - 1; i >= 0; --i) { g.st[i] = st; st *= (int)d; }
  auto recip = [](int x) { return (uint32_t)((0x100000000ull + (uint64_t)x - 1) / (uint64_t)x); };
  if (g.tile >= (1 << 16)) return PW_ERR_UNSUPPORTED;
  g.inv_c = recip(C);
  g.inv_t = recip(g.t);
  for (int i = 0; i < nops; ++i) g.inv_st[i] = recip(g.st[i]);
  g.logC = -1;
  for (int l = 0; l < 16; ++l)
    if ((1 << l) == C) g.logC = l;
  if (dtype == PW_C128)
    return launch_multi_t<double2>((const double2*)psi, (double2*)out, p, g, (int)d, (cudaStream_t)stream);
  return launch_multi_t<float2>((const float2*)psi, (float2*)out, p, g, (int)d, (cudaStream_t)stream);
}
