// Tiled apply path (see pw_tiled.cuh): host tile planner, CSR / superoperator
// preparation kernels, and the TMA + mbarrier pipelined tile kernel.
#include "pw_tiled.cuh"

#include <mutex>

namespace pw {

// =====================================================================================
// host: tile planner
// =====================================================================================

static uint32_t pick_box(uint64_t size, uint64_t cap) {
  if (cap < 1) cap = 1;
  if (size <= cap) return (uint32_t)size;
  uint64_t best = 1;
  for (uint64_t d = cap; d >= 1; --d) {
    if (size % d == 0) { best = d; break; }
  }
  if (best * 2 >= cap) return (uint32_t)best;
  return (uint32_t)cap;  // partial edge tiles: TMA zero-fills loads and clips stores
}

bool plan_tiled(const int64_t* axes, int naxes, const int* kind0, int nt0, const int* kind1,
                int nt1, int boundary, size_t elem_bytes, uint32_t want_elems, TiledGeom* g) {
  if (naxes < 1 || naxes > kMaxAxes || nt0 < 1 || nt0 > 2 * PW_MAX_TARGETS ||
      nt1 > 2 * PW_MAX_TARGETS)
    return false;
  int64_t stride[kMaxAxes];
  int cls[kMaxAxes];
  uint64_t N = 1;
  for (int i = naxes - 1; i >= 0; --i) {
    stride[i] = (int64_t)N;
    N *= (uint64_t)axes[i];
    cls[i] = 0;
  }
  for (int k = 0; k < nt0; ++k) cls[kind0[k]] = 1;
  for (int k = 0; k < nt1; ++k) cls[kind1[k]] = 2;

  // merge consecutive axes of the same class into runs (size-1 axes vanish)
  struct Run { uint64_t size; int64_t stride; int cls; int first, last; };
  Run runs[kMaxAxes];
  int run_of_axis[kMaxAxes];
  int nr = 0;
  for (int i = 0; i < naxes; ++i) {
    run_of_axis[i] = -1;
    if (axes[i] == 1) continue;
    const bool brk = (i == boundary);
    if (nr > 0 && runs[nr - 1].cls == cls[i] && !brk && runs[nr - 1].last >= 0) {
      // mergeable only if no axis of another class (size > 1) sits in between: true by
      // construction since such an axis would have opened a new run
      runs[nr - 1].size *= (uint64_t)axes[i];
      runs[nr - 1].stride = stride[i];
      runs[nr - 1].last = i;
    } else {
      runs[nr].size = (uint64_t)axes[i];
      runs[nr].stride = stride[i];
      runs[nr].cls = cls[i];
      runs[nr].first = runs[nr].last = i;
      ++nr;
    }
    run_of_axis[i] = nr - 1;
  }
  if (nr < 1 || nr > kMaxRuns) return false;
  // every target axis of size 1 is irrelevant; targets of size > 1 must map to runs
  uint64_t t_all = 1;
  for (int r = 0; r < nr; ++r)
    if (runs[r].cls != 0) t_all *= runs[r].size;
  if (t_all > want_elems * 4ull) return false;  // a single fiber would not fit a tile

  // box sizes (outermost-first arrays indexed like runs)
  uint32_t box[kMaxRuns];
  uint64_t fwant = want_elems / t_all;
  if (fwant < 1) fwant = 1;
  for (int r = nr - 1; r >= 0; --r) {
    const bool innermost = (r == nr - 1);
    if (runs[r].cls != 0) {
      if (runs[r].size > (innermost ? 128u : 256u)) return false;
      box[r] = (uint32_t)runs[r].size;
    } else {
      uint64_t cap = innermost ? 128 : 256;
      if (cap > fwant) cap = fwant;
      box[r] = pick_box(runs[r].size, cap);
      fwant = fwant / box[r];
      if (fwant < 1) fwant = 1;
    }
  }
  uint32_t tstride[kMaxRuns];
  uint64_t te = 1;
  for (int r = nr - 1; r >= 0; --r) { tstride[r] = (uint32_t)te; te *= box[r]; }
  if (te > 0x7fffffffull / 4) return false;

  g->rank = nr;
  for (int r = 0; r < nr; ++r) {
    const int d = nr - 1 - r;  // tensor-map dim (0 = innermost)
    g->gdim[d] = runs[r].size * (d == 0 ? 2 : 1);
    g->gstride[d] = (uint64_t)runs[r].stride * elem_bytes;
    g->box[d] = box[r] * (d == 0 ? 2 : 1);
    if (g->gdim[d] > 0xffffffffull) return false;
    if (d > 0 && (g->gstride[d] % 16 != 0 || g->gstride[d] >= (1ull << 40))) return false;
    if (g->box[d] > 256) return false;
  }
  if (runs[nr - 1].stride != 1) return false;
  if (((uint64_t)g->box[0] * (elem_bytes / 2)) % 16 != 0) return false;

  g->nrest = 0;
  g->ntiles = 1;
  for (int r = 0; r < nr; ++r) {
    if (runs[r].cls != 0) continue;
    const int q = g->nrest++;
    g->rest_box[q] = box[r];
    g->rest_ntiles[q] = (uint32_t)((runs[r].size + box[r] - 1) / box[r]);
    g->rest_tmdim[q] = nr - 1 - r;
    g->ntiles *= g->rest_ntiles[q];
  }
  g->tile_elems = (uint32_t)te;

  g->nkinds = nt1 > 0 ? 2 : 1;
  for (int q = 0; q < g->nkinds; ++q) {
    PassKind& pk = g->kind[q];
    const int* tg = q == 0 ? kind0 : kind1;
    const int nt = q == 0 ? nt0 : nt1;
    pk.nt = 0;
    uint64_t t = 1;
    for (int k = 0; k < nt; ++k) {
      const int a = tg[k];
      // size-1 target axes still consume an operator digit of extent 1
      uint32_t ts = 0;
      if (axes[a] > 1) {
        const int r = run_of_axis[a];
        uint64_t inner = 1;
        for (int j = a + 1; j <= runs[r].last; ++j) inner *= (uint64_t)axes[j];
        ts = (uint32_t)(tstride[r] * inner);
      }
      pk.tdim[pk.nt] = (uint32_t)axes[a];
      pk.tstride[pk.nt] = ts;
      ++pk.nt;
      t *= (uint64_t)axes[a];
    }
    pk.t = (uint32_t)t;
    pk.nfr = 0;
    uint64_t nf = 1;
    for (int r = 0; r < nr; ++r) {
      if (runs[r].cls == q + 1) continue;
      pk.fr_size[pk.nfr] = box[r];
      pk.fr_stride[pk.nfr] = tstride[r];
      ++pk.nfr;
      nf *= box[r];
    }
    pk.nfib = (uint32_t)nf;
    pk.nfib_pad = (uint32_t)((nf + 31) / 32 * 32);
  }
  return true;
}

// =====================================================================================
// device: operator preparation
// =====================================================================================

__device__ __forceinline__ int32_t kind_offset(const PassKind& pk, uint32_t j) {
  int32_t off = 0;
  for (int i = pk.nt - 1; i >= 0; --i) {
    const uint32_t q = j / pk.tdim[i];
    off += (int32_t)((j - q * pk.tdim[i]) * pk.tstride[i]);
    j = q;
  }
  return off;
}

// Dense (K, t, t) -> CSR per operator; exact zeros are dropped (adding 0*x is a no-op).
template <typename V>
__global__ void __launch_bounds__(256)
k_build_csr(const V* __restrict__ ops, const uint32_t t, const uint32_t cap, const PassKind k0,
            const PassKind k1, const int nkinds, int32_t* __restrict__ rowptr,
            Entry<V>* __restrict__ ents, const int* __restrict__ skip) {
  extern __shared__ int32_t s_cnt[];  // [t + 1]
  if (skip != nullptr && *skip != 0) return;  // the consumer exits on the same flag
  const V* op = ops + (size_t)blockIdx.x * t * t;
  int32_t* rp = rowptr + (size_t)blockIdx.x * (t + 1);
  Entry<V>* en = ents + (size_t)blockIdx.x * cap;
  for (uint32_t a = threadIdx.x; a < t; a += blockDim.x) {
    int c = 0;
    for (uint32_t b = 0; b < t; ++b) {
      const V v = op[(size_t)a * t + b];
      c += (v.x != 0 || v.y != 0);
    }
    s_cnt[a] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t run = 0;
    for (uint32_t a = 0; a < t; ++a) { const int32_t c = s_cnt[a]; s_cnt[a] = run; run += c; }
    s_cnt[t] = run;
  }
  __syncthreads();
  for (uint32_t a = threadIdx.x; a <= t; a += blockDim.x) rp[a] = s_cnt[a];
  for (uint32_t a = threadIdx.x; a < t; a += blockDim.x) {
    int32_t w = s_cnt[a];
    for (uint32_t b = 0; b < t; ++b) {
      const V v = op[(size_t)a * t + b];
      if (v.x != 0 || v.y != 0) {
        Entry<V> e;
        e.val = v;
        e.off[0] = kind_offset(k0, b);
        e.off[1] = nkinds > 1 ? kind_offset(k1, b) : 0;
        en[w++] = e;
      }
    }
  }
}

// S[(a,a'),(b,b')] = sum_k K_k[a,b] conj(K_k[a',b'])   (t^2 x t^2, row-major)
template <typename V>
__global__ void __launch_bounds__(256)
k_build_super(const V* __restrict__ ops, const int K, const uint32_t t, V* __restrict__ S) {
  const uint64_t t2 = (uint64_t)t * t, total = t2 * t2;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t row = (uint32_t)(i / t2), col = (uint32_t)(i % t2);
    const uint32_t a = row / t, ap = row % t, b = col / t, bp = col % t;
    V acc = czero<V>();
    for (int k = 0; k < K; ++k) {
      const V x = ops[((size_t)k * t + a) * t + b];
      const V y = ops[((size_t)k * t + ap) * t + bp];
      // x * conj(y)
      acc.x += x.x * y.x + x.y * y.y;
      acc.y += x.y * y.x - x.x * y.y;
    }
    S[i] = acc;
  }
}

// =====================================================================================
// device: TMA / mbarrier primitives (inline PTX, sm_100a)
// =====================================================================================

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t a = smem_u32(bar);
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void tma_load(const CUtensorMap* tm, int rank, void* dst, uint64_t* bar,
                                         const int32_t* c) {
  const uint64_t d = (uint64_t)tm;
  const uint32_t s = smem_u32(dst), b = smem_u32(bar);
  switch (rank) {
    case 1:
      asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];"
                   ::"r"(s), "l"(d), "r"(b), "r"(c[0]) : "memory");
      break;
    case 2:
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(s), "l"(d), "r"(b), "r"(c[0]), "r"(c[1]) : "memory");
      break;
    case 3:
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(s), "l"(d), "r"(b), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
      break;
    case 4:
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(s), "l"(d), "r"(b), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
      break;
    default:
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                   ::"r"(s), "l"(d), "r"(b), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
      break;
  }
}

__device__ __forceinline__ void tma_store(const CUtensorMap* tm, int rank, const void* src,
                                          const int32_t* c) {
  const uint64_t d = (uint64_t)tm;
  const uint32_t s = smem_u32(src);
  switch (rank) {
    case 1:
      asm volatile("cp.async.bulk.tensor.1d.global.shared::cta.tile.bulk_group [%0, {%2}], [%1];"
                   ::"l"(d), "r"(s), "r"(c[0]) : "memory");
      break;
    case 2:
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                   ::"l"(d), "r"(s), "r"(c[0]), "r"(c[1]) : "memory");
      break;
    case 3:
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(d), "r"(s), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
      break;
    case 4:
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                   ::"l"(d), "r"(s), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
      break;
    default:
      asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                   ::"l"(d), "r"(s), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
      break;
  }
}

// =====================================================================================
// device: the tile kernel
// =====================================================================================

// tile id -> tensor-map coordinates (innermost first); dim 0 counts reals
__device__ __forceinline__ void tile_coords(const TiledGeom& g, uint64_t tile, int32_t* c) {
#pragma unroll
  for (int d = 0; d < kMaxRuns; ++d) c[d] = 0;
  for (int q = g.nrest - 1; q >= 0; --q) {
    const uint64_t n = g.rest_ntiles[q];
    const uint64_t ti = tile % n;
    tile /= n;
    const int d = g.rest_tmdim[q];
    c[d] = (int32_t)(ti * g.rest_box[q] * (d == 0 ? 2 : 1));
  }
}

// One sparse pass over the resident tile: dst[fiber, a] (+)= sum_b O[a,b] src[fiber, b].
// Work unit = (operator row a, chunk of 32 fibers): the row is warp-uniform, so the
// CSR entry loads are broadcast (one L1 wavefront) and the state reads are one
// conflict-free LDS.128 per lane.
template <typename V, bool CONJ>
__device__ __forceinline__ void run_pass(const V* __restrict__ src, V* __restrict__ dst,
                                         const PassKind& pk, const int kidx,
                                         const int32_t* __restrict__ s_foff,
                                         const int32_t* __restrict__ s_toff,
                                         const int32_t* __restrict__ rowptr,
                                         const Entry<V>* __restrict__ ents, const bool accumulate) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t nchunk = pk.nfib_pad >> 5;
  const uint32_t units = pk.t * nchunk;
  for (uint32_t u = warp; u < units; u += nw) {
    const uint32_t a = u / nchunk;
    const uint32_t fid = (u - a * nchunk) * 32 + lane;
    const int32_t beg = __ldg(rowptr + a), end = __ldg(rowptr + a + 1);
    if (fid < pk.nfib) {
      const int32_t fo = s_foff[fid];
      V acc0 = czero<V>(), acc1 = czero<V>();
      int32_t e = beg;
      for (; e + 1 < end; e += 2) {
        const Entry<V> e0 = ents[e], e1 = ents[e + 1];
        const V x0 = src[fo + e0.off[kidx]], x1 = src[fo + e1.off[kidx]];
        if (CONJ) { cfma_conja(acc0, e0.val, x0); cfma_conja(acc1, e1.val, x1); }
        else      { cfma(acc0, e0.val, x0);       cfma(acc1, e1.val, x1); }
      }
      if (e < end) {
        const Entry<V> e0 = ents[e];
        const V x0 = src[fo + e0.off[kidx]];
        if (CONJ) cfma_conja(acc0, e0.val, x0); else cfma(acc0, e0.val, x0);
      }
      acc0.x += acc1.x; acc0.y += acc1.y;
      const int32_t o = fo + s_toff[a];
      if (accumulate) { const V d = dst[o]; acc0.x += d.x; acc0.y += d.y; }
      dst[o] = acc0;
    }
  }
}

// Block-structured pass: the operator is a set of small dense blocks (rows R_b, columns
// C_b).  Work item = (block, fiber), fibers fastest, so a warp mostly shares one block:
// descriptor loads are broadcast, the <= MAXN inputs of the block are gathered from the
// tile into registers, the dense product runs against broadcast values.
template <typename V>
__device__ __noinline__ void run_pass_blocks(const V* __restrict__ src, V* __restrict__ dst,
                                             const uint32_t nfib,
                                             const int32_t* __restrict__ s_foff,
                                             const BlockArrays A) {
  const int nblocks = A.hdr->nblocks, nzero = A.hdr->nzero;
  const V* vals = (const V*)A.vals;
  const uint32_t items = (uint32_t)nblocks * nfib;
  for (uint32_t it = threadIdx.x; it < items; it += blockDim.x) {
    const uint32_t b = it / nfib;
    const int32_t fo = s_foff[it - b * nfib];
    const int4 h = __ldg(A.blocks + b);
    const int64_t* bo = A.offs + h.z;
    block_product<V>(vals + h.w, h.x, h.y,
                     [&](int j) { return src[fo + (int32_t)__ldg(bo + j)]; },
                     [&](int a, V v) { dst[fo + (int32_t)__ldg(bo + h.y + a)] = v; });
  }
  const uint32_t zitems = (uint32_t)nzero * nfib;
  for (uint32_t it = threadIdx.x; it < zitems; it += blockDim.x) {
    const uint32_t z = it / nfib;
    dst[s_foff[it - z * nfib] + (int32_t)__ldg(A.zero_offs + z)] = czero<V>();
  }
}

// Two instantiations share the pipeline: BLOCKS = block-structured register compute (runs
// only when the device-side analysis marked the operator usable), !BLOCKS = CSR compute
// (runs otherwise).  Both are enqueued; exactly one does the work.
template <typename V, bool BLOCKS>
__global__ void __launch_bounds__(kTileThreads, 3)
k_tiled(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
        const __grid_constant__ TiledParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (P.skip && *P.skip) return;
  {
    // block-structured compute when the device-side analysis found only small blocks
    bool use_blocks = P.blk[0].hdr != nullptr && P.blk[0].hdr->usable != 0;
    if (P.mode == kModeMatrixTwoSided)
      use_blocks = use_blocks && P.K == 1 && P.blk[1].hdr != nullptr && P.blk[1].hdr->usable != 0;
    if (use_blocks != BLOCKS) return;
  }
  const TiledGeom& g = P.g;
  const uint32_t tile_bytes = (g.tile_elems * (uint32_t)sizeof(V) + 127u) & ~127u;
  const int NS = P.nstage;
  const int nwork = P.mode == kModeMatrixTwoSided ? 2 : 1;
  V* w0 = (V*)(smem + (size_t)NS * tile_bytes);
  V* w1 = (V*)(smem + (size_t)(NS + 1) * tile_bytes);
  unsigned char* tail = smem + (size_t)(NS + nwork) * tile_bytes;
  uint64_t* full = (uint64_t*)tail;  // [NS] (room for 8)
  int32_t* s_foff0 = (int32_t*)(tail + 64);
  int32_t* s_toff0 = s_foff0 + g.kind[0].nfib;
  int32_t* s_foff1 = s_toff0 + g.kind[0].t;
  int32_t* s_toff1 = s_foff1 + (g.nkinds > 1 ? g.kind[1].nfib : 0);

  // ---- one-time setup: offset tables, barriers ----------------------------------
  for (int q = 0; q < g.nkinds; ++q) {
    const PassKind& pk = g.kind[q];
    int32_t* foff = q == 0 ? s_foff0 : s_foff1;
    int32_t* toff = q == 0 ? s_toff0 : s_toff1;
    for (uint32_t f = threadIdx.x; f < pk.nfib; f += blockDim.x) {
      uint32_t r = f;
      int32_t off = 0;
      for (int i = pk.nfr - 1; i >= 0; --i) {
        const uint32_t qq = r / pk.fr_size[i];
        off += (int32_t)((r - qq * pk.fr_size[i]) * pk.fr_stride[i]);
        r = qq;
      }
      foff[f] = off;
    }
    for (uint32_t a = threadIdx.x; a < pk.t; a += blockDim.x) toff[a] = kind_offset(pk, a);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();

  const uint64_t G = gridDim.x;
  const uint32_t tx_bytes = g.tile_elems * (uint32_t)sizeof(V);
  int32_t c[kMaxRuns];
  if (threadIdx.x == 0) {
    for (int b = 0; b < NS; ++b) {
      const uint64_t tile = blockIdx.x + (uint64_t)b * G;
      if (tile < g.ntiles) {
        tile_coords(g, tile, c);
        mbar_expect_tx(&full[b], tx_bytes);
        tma_load(&tm_in, g.rank, smem + (size_t)b * tile_bytes, &full[b], c);
      }
    }
  }

  const Entry<V>* ents = (const Entry<V>*)P.ents;
  uint32_t it = 0, stage = 0, phase = 0;
  for (uint64_t tile = blockIdx.x; tile < g.ntiles; tile += G, ++it) {
    V* in_tile = (V*)(smem + (size_t)stage * tile_bytes);
    mbar_wait(&full[stage], phase);
    // the result buffer of the previous tile must have been read out by its TMA store
    if (threadIdx.x == 0) tma_store_wait_read0();
    __syncthreads();

    V* res;
    if (P.mode == kModeMatrixTwoSided) {
      if (BLOCKS) {
        run_pass_blocks<V>(in_tile, w0, g.kind[0].nfib, s_foff0, P.blk[0]);
        __syncthreads();
        run_pass_blocks<V>(w0, w1, g.kind[1].nfib, s_foff1, P.blk[1]);
      } else {
        for (int k = 0; k < P.K; ++k) {
          const int32_t* rp = P.rowptr + (size_t)k * (g.kind[0].t + 1);
          const Entry<V>* en = ents + (size_t)k * P.csr_cap;
          run_pass<V, false>(in_tile, w0, g.kind[0], 0, s_foff0, s_toff0, rp, en, false);
          __syncthreads();
          run_pass<V, true>(w0, w1, g.kind[1], 1, s_foff1, s_toff1, rp, en, k > 0);
          if (k + 1 < P.K) __syncthreads();
        }
      }
      res = w1;
    } else {
      if (BLOCKS)
        run_pass_blocks<V>(in_tile, w0, g.kind[0].nfib, s_foff0, P.blk[0]);
      else if (P.mode == kModeVectorConj)
        run_pass<V, true>(in_tile, w0, g.kind[0], 0, s_foff0, s_toff0, P.rowptr, ents, false);
      else
        run_pass<V, false>(in_tile, w0, g.kind[0], 0, s_foff0, s_toff0, P.rowptr, ents, false);
      res = w0;
    }
    fence_proxy_async();  // generic-proxy writes of `res` -> visible to the TMA store
    __syncthreads();
    if (threadIdx.x == 0) {
      tile_coords(g, tile, c);
      tma_store(&tm_out, g.rank, res, c);
      tma_store_commit();
      const uint64_t next = tile + (uint64_t)NS * G;  // this stage is free again: prefetch
      if (next < g.ntiles) {
        tile_coords(g, next, c);
        mbar_expect_tx(&full[stage], tx_bytes);
        tma_load(&tm_in, g.rank, in_tile, &full[stage], c);
      }
    }
    if (++stage == (uint32_t)NS) { stage = 0; phase ^= 1; }
  }
  if (threadIdx.x == 0) tma_store_wait_all();
}

// =====================================================================================
// host: launch
// =====================================================================================

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      (void)cudaGetLastError();
  });
  return fn;
}

static bool encode_map(CUtensorMap* tm, const void* base, const TiledGeom& g, bool f64) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[kMaxRuns], gstr[kMaxRuns];
  cuuint32_t box[kMaxRuns], estr[kMaxRuns];
  for (int d = 0; d < g.rank; ++d) {
    gdim[d] = g.gdim[d];
    box[d] = g.box[d];
    estr[d] = 1;
    if (d > 0) gstr[d - 1] = g.gstride[d];
  }
  const CUresult r = enc(tm, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                         (cuuint32_t)g.rank, const_cast<void*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static TargetMap kind_target_map(const PassKind& pk) {
  TargetMap tm;
  tm.nt = pk.nt;
  tm.t = pk.t;
  for (int i = 0; i < pk.nt; ++i) { tm.tdim[i] = pk.tdim[i]; tm.tstride[i] = (int64_t)pk.tstride[i]; }
  return tm;
}

template <typename V>
static int launch_tiled_t(const V* in, V* out, const V* ops, int K, const TiledGeom& g, int mode,
                          cudaStream_t s, const int* skip) {
  if (((uintptr_t)in | (uintptr_t)out) & 15) return PW_ERR_UNSUPPORTED;
  const int dtype = sizeof(V) == 16 ? PW_C128 : PW_C64;
  const PassKind& k0 = g.kind[0];
  const V* csr_src = ops;
  int csrK = K;
  uint32_t t = k0.t;
  Scratch super;
  if (mode == kModeMatrixSuper) {
    // kind 0 acts on (row targets, col targets) with the t^2 x t^2 superoperator
    uint32_t ts = 1;
    while ((uint64_t)ts * ts < t) ++ts;
    PW_TRY(super.alloc(sizeof(V) * (size_t)t * t, s));
    PW_TRY(build_superoperator(ops, K, ts, super.p, dtype, s));
    csr_src = (const V*)super.p;
    csrK = 1;
  }
  const uint32_t cap = t * t;
  Scratch rowptr, ents, bmem0, bmem1;
  PW_TRY(rowptr.alloc(sizeof(int32_t) * (size_t)csrK * (t + 1), s));
  PW_TRY(ents.alloc(sizeof(Entry<V>) * (size_t)csrK * cap, s));
  if (!plan_replay()) {
    k_build_csr<V><<<csrK, 256, sizeof(int32_t) * (t + 1), s>>>(
        csr_src, t, cap, g.kind[0], g.kind[g.nkinds > 1 ? 1 : 0], g.nkinds, (int32_t*)rowptr.p,
        (Entry<V>*)ents.p, skip);
    PW_LAUNCHED();
  }

  TiledParams P;
  P.g = g;
  P.mode = mode;
  P.K = csrK;
  P.skip = skip;
  P.csr_cap = cap;
  P.rowptr = (const int32_t*)rowptr.p;
  P.ents = ents.p;
  // block-structured descriptors with TILE offsets (single operator only)
  if (csrK == 1 && t <= kBlockMaxT && options().vec_path.load(std::memory_order_relaxed) != 2) {
    PW_TRY(build_block_desc(csr_src, t, mode == kModeVectorConj, kind_target_map(g.kind[0]), dtype,
                            &bmem0, &P.blk[0], s, skip));
    if (mode == kModeMatrixTwoSided)
      PW_TRY(build_block_desc(csr_src, t, true, kind_target_map(g.kind[1]), dtype, &bmem1, &P.blk[1], s, skip));
  }

  CUtensorMap tm_in, tm_out;
  if (!encode_map(&tm_in, in, g, sizeof(V) == 16) || !encode_map(&tm_out, out, g, sizeof(V) == 16))
    return PW_ERR_UNSUPPORTED;

  const uint32_t tile_bytes = (g.tile_elems * (uint32_t)sizeof(V) + 127u) & ~127u;
  const int nwork = mode == kModeMatrixTwoSided ? 2 : 1;
  size_t tables = 64;
  for (int q = 0; q < g.nkinds; ++q) tables += sizeof(int32_t) * ((size_t)g.kind[q].nfib + g.kind[q].t);
  // stages: as many as keep >= 2 CTAs per SM, between 2 and 4
  int nstage = 4;
  while (nstage > 2 && (size_t)(nstage + nwork) * tile_bytes + tables > 100 * 1024) --nstage;
  const int forced_ns = options().nstage.load(std::memory_order_relaxed);
  if (forced_ns >= 2 && forced_ns <= 8) nstage = forced_ns;
  P.nstage = nstage;
  const size_t smem = (size_t)(nstage + nwork) * tile_bytes + tables;
  if (smem > 220 * 1024) return PW_ERR_UNSUPPORTED;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    PW_CUDA(cudaFuncSetAttribute(k_tiled<V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    PW_CUDA(cudaFuncSetAttribute(k_tiled<V, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    attr_set = true;
  }
  int ctas_per_sm = (int)((227 * 1024) / (smem + 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 4) ctas_per_sm = 4;
  const int forced = options().ctas_per_sm.load(std::memory_order_relaxed);
  if (forced > 0) ctas_per_sm = forced;
  uint64_t grid = (uint64_t)kNumSMs * ctas_per_sm;
  if (grid > g.ntiles) grid = g.ntiles;
  g_paths[2].fetch_add(1, std::memory_order_relaxed);
  if (P.blk[0].hdr != nullptr) {
    k_tiled<V, true><<<(unsigned)grid, kTileThreads, smem, s>>>(tm_in, tm_out, P);
    PW_LAUNCHED();
  }
  k_tiled<V, false><<<(unsigned)grid, kTileThreads, smem, s>>>(tm_in, tm_out, P);
  PW_LAUNCHED();
  return PW_OK;
}

int launch_tiled(const void* in, void* out, const void* ops, int K, const TiledGeom& g, int mode,
                 int dtype, cudaStream_t stream, const int* skip) {
  if (dtype == PW_C128)
    return launch_tiled_t<double2>((const double2*)in, (double2*)out, (const double2*)ops, K, g, mode, stream, skip);
  if (dtype == PW_C64)
    return launch_tiled_t<float2>((const float2*)in, (float2*)out, (const float2*)ops, K, g, mode, stream, skip);
  return PW_ERR_INVALID;
}

int build_superoperator(const void* ops, int K, uint32_t t, void* S, int dtype, cudaStream_t s) {
  if (plan_replay()) return PW_OK;   // cached plan: S is already built
  const uint64_t total = (uint64_t)t * t * t * t;
  if (dtype == PW_C128)
    k_build_super<double2><<<grid_for(total, 256, 8), 256, 0, s>>>((const double2*)ops, K, t, (double2*)S);
  else if (dtype == PW_C64)
    k_build_super<float2><<<grid_for(total, 256, 8), 256, 0, s>>>((const float2*)ops, K, t, (float2*)S);
  else
    return PW_ERR_INVALID;
  PW_LAUNCHED();
  return PW_OK;
}

}  // namespace pw
