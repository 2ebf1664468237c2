// Dense one-sided apply on the FP64 tensor cores (complex128, 8 < t <= 64).
//
// A dense t x t operator on a complex128 state costs 8 t flop per 32 bytes moved; beyond t ~ 16
// the FP64 pipe, not HBM, is the roof (DESIGN.md "c128 FP64 second roof").  The FMA pipe with its
// broadcast shared-memory operand loads reaches ~6 TFLOP/s there (k_tiled: 10 % of the HBM roof
// for a dense 36 x 36 two-mode operator); mma.sync.m8n8k4.f64 keeps the operands in fragments
// and performs 256 FMAs per instruction.
//
// A CTA stages tiles of (all t operator indices) x (64 adjacent fibers) in shared memory with
// cp.async (double buffered, XOR swizzled).  A warp item is (8 output rows) x (32 fibers): four
// complex 8 x 8 accumulator tiles; per step of four complex inputs it loads ONE operator fragment
// (reused by the four fiber octets) and four state fragments -- each a single 128-bit shared
// load that yields the real AND the imaginary part -- and issues 16 DMMAs:
//   [Yr; Yi] = [Wr -Wi; Wi Wr] [Xr; Xi].
// The accumulators are streamed to HBM straight from the fragments (a lane owns two adjacent
// fibers of one output row: one full 32-byte sector).
//
// INNER: the targets ARE the innermost axes, i.e. a fiber is a contiguous run of t elements and
// 64 adjacent fibers are one contiguous piece of memory.  The tile keeps that order ([fiber][t],
// pitch = 4 mod 8 elements so that the state fragments of a quarter warp -- two fibers x four
// inputs -- fall into eight different 16-byte bank groups), the loads are one coalesced copy, and
// a store instruction writes four full 128-byte lines (eight adjacent output rows per fiber).
#include "pw_internal.cuh"
#include "pw_dense.cuh"
#include "pw_tf32.cuh"

namespace pw {

namespace {

constexpr int kDmMaxT = 128;     // 64 < t <= 128: 32-fiber tiles, operator fragments from global memory

__device__ __forceinline__ void dmma(double& d0, double& d1, const double a, const double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
// two values that are usually adjacent in memory: one 256-bit store when they are
__device__ __forceinline__ void st_pair(double2* p0, double2* p1, const double2 v0, const double2 v1,
                                        const bool ok0, const bool ok1) {
  if (ok0 && ok1 && p1 == p0 + 1 && (reinterpret_cast<uintptr_t>(p0) & 31) == 0) {
    asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p0), "d"(v0.x), "d"(v0.y), "d"(v1.x), "d"(v1.y)
                 : "memory");
  } else {
    if (ok0) __stcg(p0, v0);
    if (ok1) __stcg(p1, v1);
  }
}
__device__ __forceinline__ void cp8(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
// storage type S (double2 | float2 for complex64): the tile keeps S, fragments are widened to FP64
// when read, accumulators rounded when stored -- same DMMA count, half the bytes
__device__ __forceinline__ void cp_elem(double2* d, const double2* g) { cp16(d, g); }
__device__ __forceinline__ void cp_elem(float2* d, const float2* g) { cp8(d, g); }
__device__ __forceinline__ double2 wide(const double2 v) { return v; }
__device__ __forceinline__ double2 wide(const float2 v) { return make_double2((double)v.x, (double)v.y); }
__device__ __forceinline__ void st_pair(float2* p0, float2* p1, const double2 v0, const double2 v1,
                                        const bool ok0, const bool ok1) {
  const float2 f0 = make_float2((float)v0.x, (float)v0.y), f1 = make_float2((float)v1.x, (float)v1.y);
  if (ok0 && ok1 && p1 == p0 + 1 && (reinterpret_cast<uintptr_t>(p0) & 15) == 0) {
    __stcg(reinterpret_cast<float4*>(p0), make_float4(f0.x, f0.y, f1.x, f1.y));
  } else {
    if (ok0) __stcg(p0, f0);
    if (ok1) __stcg(p1, f1);
  }
}
__device__ __forceinline__ void st_one(double2* p, const double2 v) { __stcg(p, v); }
__device__ __forceinline__ void st_one(float2* p, const double2 v) { __stcg(p, make_float2((float)v.x, (float)v.y)); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// blockDim = 32 * nwarps; dynamic smem: W [8 nm][wp] | offs [t] | base [2][64] | tiles [2][t][64]
// (INNER: tiles [2][64][tp]; base[buf][0..1] = first element of the tile, valid fibers)
// FIB: fibers per tile (64; 32 for 64 < t <= 128 so that two tiles still fit).  WG: the operator
// fragments come from global memory (L1 / L2 resident) instead of a shared-memory copy -- a
// 128 x 128 complex128 operator (the Jaynes-Cummings propagator, cutoff 64 (x) qubit, applied to a
// BATCH of trajectories) is 256 KB and does not fit next to the tiles.
// TF (complex64 only): the products run on the TF32 tensor cores, error-compensated (3 x TF32,
// pw_tf32.cuh) -- the FP64 path is DMMA-bound for complex64 (same DMMA count, half the bytes).
template <bool INNER, typename S, int FIB, bool WG, bool TF>
__global__ void __launch_bounds__(512, 1)
k_apply_dense_mma(const S* __restrict__ in, S* __restrict__ out,
                  const S* __restrict__ op, const int conj_op, const FiberMap fm,
                  const TargetMap tm, const int* __restrict__ skip) {
  using V = double2;
  if (skip != nullptr && *skip != 0) return;
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int t = (int)tm.t;
  const int nm = (t + 7) >> 3;             // groups of 8 output rows
  const int ks = (t + 3) >> 2;             // steps of 4 complex inputs
  const int wp = (4 * ks + 7) & ~7;        // operand row pitch (multiple of 8: bank analysis)
  V* s_w = (V*)s_raw;                                        // [8 nm][wp], column c at (c ^ 4 (row & 1))
  int64_t* s_off = (int64_t*)(s_w + (WG ? 0 : 8 * nm * wp)); // [t]
  int64_t* s_base = s_off + ((t + 1) & ~1);                  // [2][FIB]
  S* tile0 = (S*)(s_base + 2 * FIB);                      // [2][t][FIB], fiber f of row j at (f ^ 2 (j & 3))
  int tp = t;                                                // INNER: fiber pitch, 4 mod 8
  while ((tp & 7) != 4) ++tp;
  const size_t tile_elems = INNER ? (size_t)FIB * tp : (size_t)t * FIB;
  const uint32_t inv_t = 0xffffffffu / (uint32_t)t + 1;      // e / t for e < 2^16
  const uint64_t gl = INNER ? fm.gsize[fm.ng - 1] : 1;       // INNER: fibers per contiguous run
  const uint64_t cpr = (gl + FIB - 1) / FIB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  for (int i = threadIdx.x; i < (WG ? 0 : 8 * nm * wp); i += blockDim.x) {
    const int r = i / wp, c = i - r * wp;
    V v = make_double2(0.0, 0.0);
    if (r < t && c < t) {
      v = wide(op[(size_t)r * t + c]);
      if (conj_op) v.y = -v.y;
    }
    s_w[r * wp + (c ^ ((r & 1) << 2))] = v;
  }
  for (int j = threadIdx.x; j < t; j += blockDim.x) s_off[j] = tm.offset((uint32_t)j);
  __syncthreads();
  const uint64_t units = INNER ? (fm.nfib / gl) * cpr : (fm.nfib + FIB - 1) / FIB;

  auto issue = [&](uint64_t u, int buf) {
    if (INNER) {
      if (u < units) {
        S* tile = tile0 + (size_t)buf * tile_elems;
        const uint64_t o = u / cpr, w = u - o * cpr;
        const uint64_t f0 = w * FIB;
        const int nvalid = (gl - f0) < (uint64_t)FIB ? (int)(gl - f0) : FIB;
        const int64_t base0 = fm.base(o * gl + f0);
        if (threadIdx.x == 0) { s_base[buf * FIB] = base0; s_base[buf * FIB + 1] = nvalid; }
        const S* src = in + base0;
        for (int e = threadIdx.x; e < nvalid * t; e += blockDim.x) {
          const int f = (int)__umulhi((uint32_t)e, inv_t);
          cp_elem(tile + f * tp + (e - f * t), src + e);
        }
      }
      cp_commit();
      return;
    }
    if (u < units) {
      S* tile = tile0 + (size_t)buf * tile_elems;
      const uint64_t fa = u * FIB + lane, fb = fa + 32;
      const int64_t ba = fa < fm.nfib ? fm.base(fa) : -1, bb = (FIB == 64 && fb < fm.nfib) ? fm.base(fb) : -1;
      if (warp == 0) {
        s_base[buf * FIB + lane] = ba;
        if (FIB == 64) s_base[buf * FIB + 32 + lane] = bb;
      }
      for (int j = warp; j < t; j += nwarps) {
        const int sw = 2 * (j & 3);
        const int64_t o = s_off[j];
        if (ba >= 0) cp_elem(tile + j * FIB + (lane ^ sw), in + ba + o);
        if (FIB == 64 && bb >= 0) cp_elem(tile + j * FIB + 32 + (lane ^ sw), in + bb + o);
      }
    }
    cp_commit();
  };

  const int nitems = nm * (FIB / 32);
  // WG: fragment W[row][j] straight from global memory
  auto wglob = [&](int row, int j) {
    V v = make_double2(0.0, 0.0);
    if (row < t && j < t) {
      v = wide(__ldg(op + (size_t)row * t + j));
      if (conj_op) v.y = -v.y;
    }
    return v;
  };
  const int fsw = (gid & 1) << 2;
  uint64_t u = blockIdx.x;
  int cur = 0;
  issue(u, 0);
  for (; u < units; u += gridDim.x) {
    const S* tile = tile0 + (size_t)cur * tile_elems;
    issue(u + gridDim.x, cur ^ 1);
    cp_wait<1>();
    __syncthreads();
    const int64_t* sb = s_base + cur * FIB;
    for (int it = warp; it < nitems; it += nwarps) {
      const int m = FIB == 64 ? it >> 1 : it, h = FIB == 64 ? it & 1 : 0;
      const int row = 8 * m + gid;
      double acc[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[n][q] = 0.0;
      const V* wrow = s_w + row * wp;
      const S* xcol = INNER ? tile + (32 * h + gid) * tp                 // + 8 n tp + position of j
                            : tile + 32 * h + (gid ^ (2 * tig));        // + j * 64 + 8 n   (j & 3 == tig)
      if constexpr (TF) {
        // two k4 steps per m16n8k8 MMA: inputs 8 k2 + tig and 8 k2 + 4 + tig
        float c[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int q = 0; q < 4; ++q) c[n][q] = 0.f;
        const int ks2 = (ks + 1) >> 1;
#pragma unroll 2
        for (int k2 = 0; k2 < ks2; ++k2) {
          const int ja = 8 * k2 + tig, jb = ja + 4;
          const V wa = WG ? wglob(row, ja) : wrow[(8 * k2 + tig) ^ fsw];
          const V wb = WG ? wglob(row, jb) : ((8 * k2 + 4) < wp ? wrow[(8 * k2 + 4 + tig) ^ fsw] : make_double2(0.0, 0.0));
          const TfW w = tf_w(make_float2((float)wa.x, (float)wa.y), make_float2((float)wb.x, (float)wb.y));
          float2 xa[4], xb[4];
          if (INNER) {
            const int pa = ja < t ? (int)s_off[ja] : 0, pb = jb < t ? (int)s_off[jb] : 0;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              xa[n] = ja < t ? xcol[8 * n * tp + pa] : make_float2(0.f, 0.f);
              xb[n] = jb < t ? xcol[8 * n * tp + pb] : make_float2(0.f, 0.f);
            }
          } else {
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              xa[n] = ja < t ? xcol[ja * FIB + 8 * n] : make_float2(0.f, 0.f);
              xb[n] = jb < t ? xcol[jb * FIB + 8 * n] : make_float2(0.f, 0.f);
            }
          }
          cprod_tf32_batch<4, true>(w, xa, xb, c);
        }
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[n][q] = (double)c[n][q];
      }
      V wnext = (WG && !TF) ? wglob(row, tig) : make_double2(0.0, 0.0);
#pragma unroll 2
      for (int k = 0; k < (TF ? 0 : ks); ++k) {
        const int j = 4 * k + tig;
        const V w = WG ? wnext : wrow[(4 * k + tig) ^ fsw];
        if (WG && k + 1 < ks) wnext = wglob(row, j + 4);   // next step's fragment: hides the L1 / L2 latency
        V x[4];
        if (INNER) {
          const int pj = j < t ? (int)s_off[j] : 0;
#pragma unroll
          for (int n = 0; n < 4; ++n) x[n] = j < t ? wide(xcol[8 * n * tp + pj]) : make_double2(0.0, 0.0);
        } else {
#pragma unroll
          for (int n = 0; n < 4; ++n) x[n] = j < t ? wide(xcol[j * FIB + 8 * n]) : make_double2(0.0, 0.0);
        }
        // (Gauss' three-product complex multiply -- 12 DMMAs per step instead of 16 -- measured SLOWER
        //  here: 28.6-30.9 against 25.1 ms, also with the sums software-pipelined, although it gains
        //  3-7 % in the two-sided and row-tile kernels: see cmma_batch in pw_blocks.cu)
        // step-major issue order: the eight accumulator chains advance together (a warp issues
        // in order; consecutive DMMAs must be independent to keep the tensor pipe busy)
#pragma unroll
        for (int n = 0; n < 4; ++n) dmma(acc[n][0], acc[n][1], w.x, x[n].x);
#pragma unroll
        for (int n = 0; n < 4; ++n) dmma(acc[n][2], acc[n][3], w.y, x[n].x);
#pragma unroll
        for (int n = 0; n < 4; ++n) dmma(acc[n][0], acc[n][1], -w.y, x[n].y);
#pragma unroll
        for (int n = 0; n < 4; ++n) dmma(acc[n][2], acc[n][3], w.x, x[n].y);
      }
      if (INNER) {
        if (row < t) {
          const int nvalid = (int)sb[1];
          S* dst = out + sb[0] + s_off[row];
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const int f = 32 * h + 8 * n + 2 * tig;
            if (f < nvalid) st_one(dst + (int64_t)f * t, make_double2(acc[n][0], acc[n][2]));
            if (f + 1 < nvalid) st_one(dst + (int64_t)(f + 1) * t, make_double2(acc[n][1], acc[n][3]));
          }
        }
      } else if (row < t) {
        const int64_t ro = s_off[row];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int f = 32 * h + 8 * n + 2 * tig;
          const int64_t b0 = sb[f], b1 = sb[f + 1];
          st_pair(out + b0 + ro, out + b1 + ro, make_double2(acc[n][0], acc[n][2]),
                  make_double2(acc[n][1], acc[n][3]), b0 >= 0, b1 >= 0);
        }
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  cp_wait<0>();
}

}  // namespace

template <typename S, int FIB, bool WG, bool TF>
static int launch_dense_mma_s(const S* in, S* out, const S* op, const Plan& p, bool conj_op, bool inner,
                              cudaStream_t s, const int* skip) {
  const int t = (int)p.tm.t;
  const int nm = (t + 7) >> 3, ks = (t + 3) >> 2, wp = (4 * ks + 7) & ~7;
  int tp = t;
  while ((tp & 7) != 4) ++tp;
  const size_t smem = (WG ? 0 : sizeof(double2) * (size_t)(8 * nm * wp)) +
                      sizeof(int64_t) * (size_t)(((t + 1) & ~1) + 2 * FIB) +
                      sizeof(S) * (size_t)2 * (inner ? tp : t) * FIB;
  if (smem > 220 * 1024) return PW_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    PW_CUDA(cudaFuncSetAttribute(k_apply_dense_mma<false, S, FIB, WG, TF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    PW_CUDA(cudaFuncSetAttribute(k_apply_dense_mma<true, S, FIB, WG, TF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    attr_set = true;
  }
  // one warp per item (8 rows x 32 fibers); two CTAs per SM when the tiles allow it
  int nwarps = nm * (FIB / 32);
  if (nwarps > 16) nwarps = 16;
  const int ctas = (smem + 1024) * 2 <= 227 * 1024 ? 2 : 1;
  const uint64_t gl = inner ? p.fm.gsize[p.fm.ng - 1] : 1;
  const uint64_t units = inner ? (p.fm.nfib / gl) * ((gl + FIB - 1) / FIB) : (p.fm.nfib + FIB - 1) / FIB;
  uint64_t grid = (uint64_t)kNumSMs * ctas;
  if (grid > units) grid = units;
  if (inner)
    k_apply_dense_mma<true, S, FIB, WG, TF><<<(unsigned)grid, nwarps * 32, smem, s>>>(in, out, op, conj_op ? 1 : 0, p.fm, p.tm, skip);
  else
    k_apply_dense_mma<false, S, FIB, WG, TF><<<(unsigned)grid, nwarps * 32, smem, s>>>(in, out, op, conj_op ? 1 : 0, p.fm, p.tm, skip);
  PW_LAUNCHED();
  return PW_OK;
}

int launch_dense_mma(const void* in, void* out, const void* op, const Plan& p, int dtype, bool conj_op,
                     cudaStream_t s, const int* skip) {
  const int t = (int)p.tm.t;
  if (in == out || t <= 8 || t > kDmMaxT || p.fm.ng == 0) return PW_ERR_UNSUPPORTED;
  if (dtype == PW_C64 && !options().ts_mma_c64.load(std::memory_order_relaxed)) return PW_ERR_UNSUPPORTED;
  // fibers adjacent in memory (untouched innermost run), or the targets are the innermost run
  bool inner = false;
  if (p.fm.gstride[p.fm.ng - 1] != 1) {
    inner = p.fm.gstride[p.fm.ng - 1] == (int64_t)t;
    for (int k = 0; k < p.tm.nt && inner; ++k) inner = p.tm.tdim[k] == 1 || p.tm.tstride[k] < (int64_t)t;
    if (!inner) return PW_ERR_UNSUPPORTED;
  } else if (p.fm.gsize[p.fm.ng - 1] < 16) {
    return PW_ERR_UNSUPPORTED;
  }
  const bool tf = options().ts_mma_c64.load(std::memory_order_relaxed) == 1;  // complex64: 3 x TF32 (2 = widened FP64)
  if (t > 64) {
    if (dtype == PW_C128)
      return launch_dense_mma_s<double2, 32, true, false>((const double2*)in, (double2*)out, (const double2*)op, p, conj_op, inner, s, skip);
    if (dtype == PW_C64 && tf)
      return launch_dense_mma_s<float2, 32, true, true>((const float2*)in, (float2*)out, (const float2*)op, p, conj_op, inner, s, skip);
    if (dtype == PW_C64)
      return launch_dense_mma_s<float2, 32, true, false>((const float2*)in, (float2*)out, (const float2*)op, p, conj_op, inner, s, skip);
    return PW_ERR_UNSUPPORTED;
  }
  if (dtype == PW_C128)
    return launch_dense_mma_s<double2, 64, false, false>((const double2*)in, (double2*)out, (const double2*)op, p, conj_op, inner, s, skip);
  if (dtype == PW_C64 && tf)
    return launch_dense_mma_s<float2, 64, false, true>((const float2*)in, (float2*)out, (const float2*)op, p, conj_op, inner, s, skip);
  if (dtype == PW_C64)
    return launch_dense_mma_s<float2, 64, false, false>((const float2*)in, (float2*)out, (const float2*)op, p, conj_op, inner, s, skip);
  return PW_ERR_UNSUPPORTED;
}

}  // namespace pw
