// Block-structured apply path.
//
// Most photonic operators are sparse with structure: phase shifts are diagonal,
// creation/annihilation and loss Kraus operators are single off-diagonals, beam
// splitters conserve total photon number (block diagonal after a permutation), and the
// superoperator sum_k K_k (x) conj(K_k) of a loss channel couples only entries with
// equal row-column difference.  A preparation kernel finds the connected components of
// the operator's bipartite row/column graph on the device; each component is a small
// dense block (rows R_b, columns C_b).  Because the blocks partition rows and columns,
// every fiber element is read once and written once, so the apply kernel needs no
// shared-memory staging at all: one thread owns one fiber, gathers a block's <= MAXN
// inputs with coalesced 128-bit loads (threads walk the innermost untouched run), does
// the dense |R_b| x |C_b| product against broadcast operator values and writes the
// block's outputs -- the same access pattern as the t <= 8 register kernel, for any t.
//
// The host never learns the structure (no sync): the apply kernel reads `usable` from
// the descriptor and exits when some block exceeds MAXN; the caller then enqueues the
// fallback path, whose kernels exit when `usable` is set.
#include "pw_internal.cuh"
#include "pw_blocks.cuh"

namespace pw {

// ---- preparation ---------------------------------------------------------------------
// Single CTA.  Label propagation on the bipartite graph (row a -- column b iff O[a,b] != 0)
// converges to label = smallest column index of the component.
template <typename V>
__global__ void __launch_bounds__(256)
k_build_blocks(const V* __restrict__ op, const uint32_t t, const bool conj_op, const TargetMap tm,
               BlockDescHeader* __restrict__ hdr, int4* __restrict__ blocks,
               int64_t* __restrict__ offs, V* __restrict__ vals, int64_t* __restrict__ zero_offs) {
  extern __shared__ int32_t s_i[];
  int32_t* lr = s_i;           // [t] row labels
  int32_t* lc = lr + t;        // [t] column labels
  int32_t* bid = lc + t;       // [t] label -> block id
  __shared__ int s_changed;
  __shared__ int s_nblocks, s_maxn, s_nzero;
  const int32_t INF = 0x7fffffff;
  for (uint32_t i = threadIdx.x; i < t; i += blockDim.x) { lr[i] = INF; lc[i] = (int32_t)i; }
  __syncthreads();
  for (uint32_t iter = 0; iter < 2 * t + 2; ++iter) {
    if (threadIdx.x == 0) s_changed = 0;
    __syncthreads();
    for (uint32_t a = threadIdx.x; a < t; a += blockDim.x) {
      int32_t m = lr[a];
      for (uint32_t b = 0; b < t; ++b) {
        const V v = op[(size_t)a * t + b];
        if (v.x != 0 || v.y != 0) m = min(m, lc[b]);
      }
      if (m != lr[a]) { lr[a] = m; s_changed = 1; }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < t; b += blockDim.x) {
      int32_t m = lc[b];
      for (uint32_t a = 0; a < t; ++a) {
        const V v = op[(size_t)a * t + b];
        if (v.x != 0 || v.y != 0) m = min(m, lr[a]);
      }
      if (m != lc[b]) { lc[b] = m; s_changed = 1; }
    }
    __syncthreads();
    if (!s_changed) break;
    __syncthreads();
  }
  // columns whose label is their own index AND that touch at least one row open a block
  if (threadIdx.x == 0) {
    int nb = 0;
    for (uint32_t b = 0; b < t; ++b) {
      bid[b] = -1;
      if (lc[b] == (int32_t)b) {
        bool used = false;
        for (uint32_t a = 0; a < t && !used; ++a) used = (lr[a] == (int32_t)b);
        if (used) bid[b] = nb++;
      }
    }
    s_nblocks = nb;
    s_maxn = 0;
    int nz = 0;
    for (uint32_t a = 0; a < t; ++a)
      if (lr[a] == INF) zero_offs[nz++] = tm.offset(a);
    s_nzero = nz;
  }
  __syncthreads();
  // one thread per block gathers its members; offsets/values are laid out per block in
  // fixed-size slots (a block can never exceed t members)
  for (uint32_t root = threadIdx.x; root < t; root += blockDim.x) {
    const int b = bid[root];
    if (b < 0) continue;
    int nc = 0, nr = 0;
    for (uint32_t c = 0; c < t; ++c) nc += (lc[c] == (int32_t)root);
    for (uint32_t a = 0; a < t; ++a) nr += (lr[a] == (int32_t)root);
    atomicMax(&s_maxn, max(nr, nc));
    blocks[b] = make_int4(nr, nc, (int)root, 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // exclusive scans for the packed layouts
    int io = 0, vo = 0;
    for (int b = 0; b < s_nblocks; ++b) {
      int4 h = blocks[b];
      const int nr = h.x, nc = h.y;
      h.w = vo;            // value offset
      blocks[b] = make_int4(nr, nc, h.z, vo);
      // reuse bid[root] to carry the index offset of this block
      bid[h.z] = io;
      io += nr + nc;
      vo += nr * nc;
    }
    hdr->nblocks = s_nblocks;
    hdr->n_offs = io;
    hdr->n_vals = (s_maxn <= kBlockMaxN) ? vo : 0;
    hdr->maxn = s_maxn;
    hdr->nzero = s_nzero;
    hdr->usable = (s_maxn <= kBlockMaxN) ? 1 : 0;
    // position a is written by the block of ROW a and read by the block of COLUMN a:
    // in-place is safe iff those coincide wherever both exist (zero rows write too)
    int safe = 1;
    for (uint32_t a = 0; a < t; ++a) {
      bool col_used = false;
      for (uint32_t r = 0; r < t && !col_used; ++r) col_used = (lr[r] == lc[a]);
      if (!col_used) continue;                 // nobody reads position a
      if (lr[a] != lc[a]) safe = 0;            // written by another block (or cleared)
    }
    hdr->inplace_safe = safe;
    hdr->seg_handled = (hdr->usable && safe) ? 1 : 0;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < s_nblocks; b += blockDim.x) {
    const int4 h = blocks[b];
    const int nr = h.x, nc = h.y, root = h.z, vo = h.w;
    const int io = bid[root];
    int cols[kBlockMaxN];
    // columns first, then rows
    int k = 0;
    for (uint32_t c = 0; c < t; ++c)
      if (lc[c] == root) {
        if (k < kBlockMaxN) cols[k] = (int)c;
        offs[io + k] = tm.offset(c);
        ++k;
      }
    int r = 0;
    for (uint32_t a = 0; a < t; ++a)
      if (lr[a] == root) {
        offs[io + nc + r] = tm.offset(a);
        if (nc <= kBlockMaxN)
          for (int j = 0; j < nc; ++j) {
            V v = op[(size_t)a * t + cols[j]];
            if (conj_op) v.y = -v.y;
            vals[vo + r * nc + j] = v;
          }
        ++r;
      }
    blocks[b] = make_int4(nr, nc, io, vo);
  }
}

// ---- apply -----------------------------------------------------------------------------
constexpr uint32_t kDescSmemCap = 24 * 1024;

template <typename V>
__global__ void __launch_bounds__(256, 3)
k_apply_blocks(const V* __restrict__ in, V* __restrict__ out, const FiberMap fm, const BlockArrays A) {
  extern __shared__ __align__(16) unsigned char s_desc[];
  if (!A.hdr->usable) return;
  const DescView<V> dv = desc_view<V>(A, s_desc, kDescSmemCap);
  const bool scalar_blocks = A.hdr->maxn == 1;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib; f += stride) {
    const int64_t base = fm.base(f);
    const V* src = in + base;
    V* dst = out + base;
    if (scalar_blocks) {
      // every block is 1 x 1 (diagonal / shift / generalized permutation): out[r_b] = v_b in[c_b];
      // blocks are laid out back to back (offs = c0 r0 c1 r1 ..., vals = v0 v1 ...), and the
      // loads are independent, so issue four at a time
      int b = 0;
      for (; b + 4 <= dv.nblocks; b += 4) {
        V x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = __ldcg(src + dv.offs[2 * (b + u)]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          V acc = czero<V>();
          cfma(acc, dv.vals[b + u], x[u]);
          __stcg(dst + dv.offs[2 * (b + u) + 1], acc);
        }
      }
      for (; b < dv.nblocks; ++b) {
        V acc = czero<V>();
        cfma(acc, dv.vals[b], __ldcg(src + dv.offs[2 * b]));
        __stcg(dst + dv.offs[2 * b + 1], acc);
      }
    } else {
      for (int b = 0; b < dv.nblocks; ++b) {
        const int4 h = dv.blocks[b];
        const int64_t* bo = dv.offs + h.z;
        block_product<V>(dv.vals + h.w, h.x, h.y,
                         [&](int j) { return __ldcg(src + bo[j]); },
                         [&](int a, V v) { __stcg(dst + bo[h.y + a], v); });
      }
    }
    for (int z = 0; z < dv.nzero; ++z) __stcg(dst + dv.zero_offs[z], czero<V>());
  }
}

// ---- staged (segmented) apply ------------------------------------------------------------
struct SegGeom {
  uint32_t t;          // operator dimension = elements per fiber
  uint32_t lseg;       // contiguous elements per segment (innermost target run)
  uint32_t nseg;       // segments per fiber
  uint32_t pitch;      // shared-memory row pitch per fiber (odd -> conflict-free)
  uint64_t M;          // extent of the rest run adjacent to the inner run (fibers lseg apart)
  int64_t seg_off[64]; // global element offset of segment s
  FiberMap outer;      // remaining untouched axes
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
}

template <typename V>
__global__ void __launch_bounds__(256)
k_apply_blocks_seg(const V* __restrict__ in, V* __restrict__ out, const SegGeom g,
                   const BlockArrays A, const uint32_t desc_cap) {
  if (!A.hdr->seg_handled) return;
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const DescView<V> dv = desc_view<V>(A, s_raw, desc_cap);
  V* buf = (V*)(s_raw + desc_cap) + (size_t)warp * 32 * g.pitch;
  const uint64_t chunks = (g.M + 31) / 32;
  const uint64_t ntasks = g.outer.nfib * chunks;
  // lane-constant decomposition of the coalesced copy index e = it*32 + lane into
  // (fiber q, position r) without divisions: each iteration adds 32 = qa*lseg + rb
  const uint32_t qa = 32 / g.lseg, rb = 32 % g.lseg;
  for (uint64_t task = (uint64_t)blockIdx.x * nwarps + warp; task < ntasks;
       task += (uint64_t)gridDim.x * nwarps) {
    const uint64_t o = task / chunks, c = task - o * chunks;
    const uint64_t m0 = c * 32;
    const uint32_t nvalid = (uint32_t)((g.M - m0) < 32 ? (g.M - m0) : 32);
    const int64_t base = g.outer.base(o) + (int64_t)m0 * g.lseg;
    const uint32_t row_elems = nvalid * g.lseg;
    // ---- stage: nseg coalesced rows -> padded per-fiber rows in shared memory
    for (uint32_t s = 0; s < g.nseg; ++s) {
      const V* src = in + base + g.seg_off[s];
      uint32_t q = (uint32_t)lane / g.lseg, r = (uint32_t)lane % g.lseg;
      for (uint32_t e = lane; e < row_elems; e += 32) {
        V* dst = buf + (size_t)q * g.pitch + s * g.lseg + r;
        if (sizeof(V) == 16) cp_async16(dst, src + e); else cp_async8(dst, src + e);
        q += qa; r += rb;
        if (r >= g.lseg) { r -= g.lseg; ++q; }
      }
    }
    cp_async_wait_all();
    __syncwarp();
    // ---- compute: one fiber per lane, blocks applied in place
    if ((uint32_t)lane < nvalid) {
      V* fib = buf + (size_t)lane * g.pitch;
      for (int b = 0; b < dv.nblocks; ++b) {
        const int4 h = dv.blocks[b];
        const int64_t* bo = dv.offs + h.z;
        block_product<V>(dv.vals + h.w, h.x, h.y,
                         [&](int j) { return fib[(int32_t)bo[j]]; },
                         [&](int a, V v) { fib[(int32_t)bo[h.y + a]] = v; });
      }
      for (int z = 0; z < dv.nzero; ++z) fib[(int32_t)dv.zero_offs[z]] = czero<V>();
    }
    __syncwarp();
    // ---- stream back, same coalesced row pattern
    for (uint32_t s = 0; s < g.nseg; ++s) {
      V* dstg = out + base + g.seg_off[s];
      uint32_t q = (uint32_t)lane / g.lseg, r = (uint32_t)lane % g.lseg;
      for (uint32_t e = lane; e < row_elems; e += 32) {
        __stcg(dstg + e, buf[(size_t)q * g.pitch + s * g.lseg + r]);
        q += qa; r += rb;
        if (r >= g.lseg) { r -= g.lseg; ++q; }
      }
    }
    __syncwarp();
  }
}

template <typename V>
static int build_block_desc_t(const V* op, uint32_t t, bool conj_op, const TargetMap& tm,
                              Scratch* mem, BlockArrays* arr, cudaStream_t s) {
  // descriptor storage: header | blocks[t] | offs[2t] | zero_offs[t] | vals[t*t]
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t b_hdr = 64, b_blk = sizeof(int4) * t, b_off = al(sizeof(int64_t) * 2 * t),
               b_zero = al(sizeof(int64_t) * t), b_val = sizeof(V) * (size_t)t * t;
  PW_TRY(mem->alloc(b_hdr + b_blk + b_off + b_zero + b_val, s));
  char* base = (char*)mem->p;
  arr->hdr = (BlockDescHeader*)base;
  arr->blocks = (int4*)(base + b_hdr);
  arr->offs = (int64_t*)(base + b_hdr + b_blk);
  arr->zero_offs = (int64_t*)(base + b_hdr + b_blk + b_off);
  arr->vals = base + b_hdr + b_blk + b_off + b_zero;
  k_build_blocks<V><<<1, 256, sizeof(int32_t) * 3 * t, s>>>(op, t, conj_op, tm, arr->hdr, arr->blocks,
                                                          arr->offs, (V*)arr->vals, arr->zero_offs);
  PW_LAUNCHED();
  return PW_OK;
}

int build_block_desc(const void* op, uint32_t t, bool conj_op, const TargetMap& tm, int dtype,
                     Scratch* mem, BlockArrays* arr, cudaStream_t stream) {
  if (t > kBlockMaxT) return PW_ERR_UNSUPPORTED;
  if (dtype == PW_C128) return build_block_desc_t<double2>((const double2*)op, t, conj_op, tm, mem, arr, stream);
  if (dtype == PW_C64) return build_block_desc_t<float2>((const float2*)op, t, conj_op, tm, mem, arr, stream);
  return PW_ERR_INVALID;
}

template <typename V>
static int launch_blocks_t(const V* in, V* out, const V* op, const Plan& p, bool conj_op,
                           BlockDesc* d, cudaStream_t s) {
  PW_TRY(build_block_desc_t<V>(op, p.tm.t, conj_op, p.tm, &d->mem, &d->arr, s));
  d->hdr = d->arr.hdr;
  const BlockArrays& A = d->arr;
  const int grid = grid_for(p.fm.nfib, 256, 3);
  k_apply_blocks<V><<<grid, 256, kDescSmemCap, s>>>(in, out, p.fm, A);
  PW_LAUNCHED();
  g_paths[4].fetch_add(1, std::memory_order_relaxed);
  return PW_OK;
}

int launch_blocks(const void* in, void* out, const void* op, const Plan& plan, int dtype,
                  bool conj_op, BlockDesc* desc, cudaStream_t stream) {
  if (plan.tm.t > kBlockMaxT || in == out) return PW_ERR_UNSUPPORTED;
  if (dtype == PW_C128)
    return launch_blocks_t<double2>((const double2*)in, (double2*)out, (const double2*)op, plan, conj_op, desc, stream);
  if (dtype == PW_C64)
    return launch_blocks_t<float2>((const float2*)in, (float2*)out, (const float2*)op, plan, conj_op, desc, stream);
  return PW_ERR_INVALID;
}

// Geometry of the staged kernel; false when the innermost axis is not a target or the
// sizes do not fit.
static bool plan_seg(const int64_t* axes, int naxes, const int* targets, int nt, SegGeom* g,
                     TargetMap* smem_tm) {
  if (naxes < 1 || naxes > kMaxAxes || nt < 1 || nt > kMaxTargets) return false;
  int64_t stride[kMaxAxes];
  bool is_t[kMaxAxes] = {false};
  uint64_t N = 1;
  for (int i = naxes - 1; i >= 0; --i) { stride[i] = (int64_t)N; N *= (uint64_t)axes[i]; }
  uint64_t t = 1;
  for (int k = 0; k < nt; ++k) { is_t[targets[k]] = true; t *= (uint64_t)axes[targets[k]]; }
  int a = naxes - 1;
  while (a >= 0 && axes[a] == 1) --a;
  if (a < 0 || !is_t[a]) return false;
  // inner target run (contiguous, innermost)
  const int inner_hi = a;
  uint64_t lseg = 1;
  while (a >= 0 && (is_t[a] || axes[a] == 1)) { if (is_t[a]) lseg *= (uint64_t)axes[a]; --a; }
  const int inner_lo = a + 1;
  // adjacent rest run
  uint64_t M = 1;
  while (a >= 0 && !is_t[a]) { M *= (uint64_t)axes[a]; --a; }
  const int outer_hi = a;  // axes [0, outer_hi] are "outer"
  if (t > kBlockMaxT || lseg > 0xffffffffull || t / lseg > 64 || M < 2) return false;
  g->t = (uint32_t)t;
  g->lseg = (uint32_t)lseg;
  g->nseg = (uint32_t)(t / lseg);
  g->pitch = (uint32_t)(t | 1);
  g->M = M;
  // outer axes: targets enumerate segments (tensor order, first most significant), the
  // rest forms the outer fiber map
  int outer_t[kMaxAxes], n_ot = 0;
  FiberMap& fm = g->outer;
  fm.ng = 0;
  fm.nfib = 1;
  int i = 0;
  while (i <= outer_hi) {
    if (is_t[i]) { if (axes[i] > 1) outer_t[n_ot++] = i; ++i; continue; }
    uint64_t size = 1;
    int last = i;
    while (i <= outer_hi && !is_t[i]) { size *= (uint64_t)axes[i]; last = i; ++i; }
    if (size == 1) continue;
    if (fm.ng >= kMaxGroups) return false;
    fm.gsize[fm.ng] = size;
    fm.gstride[fm.ng] = stride[last];
    fm.nfib *= size;
    ++fm.ng;
  }
  for (uint32_t s = 0; s < g->nseg; ++s) {
    uint32_t rem = s;
    int64_t off = 0;
    for (int k = n_ot - 1; k >= 0; --k) {
      const uint32_t d = (uint32_t)axes[outer_t[k]];
      off += (int64_t)(rem % d) * stride[outer_t[k]];
      rem /= d;
    }
    g->seg_off[s] = off;
  }
  // shared-memory position of operator index j: inner-run axes keep their contiguous
  // strides, outer target axes step by lseg in segment order
  smem_tm->nt = nt;
  smem_tm->t = (uint32_t)t;
  for (int k = 0; k < nt; ++k) {
    const int ax = targets[k];
    smem_tm->tdim[k] = (uint32_t)axes[ax];
    if (ax >= inner_lo && ax <= inner_hi) {
      smem_tm->tstride[k] = stride[ax];  // innermost run: global strides ARE the in-row strides
    } else {
      int64_t st = (int64_t)lseg;
      for (int q = n_ot - 1; q >= 0 && outer_t[q] != ax; --q) st *= axes[outer_t[q]];
      smem_tm->tstride[k] = axes[ax] > 1 ? st : 0;
    }
  }
  return true;
}

template <typename V>
static int launch_blocks_seg_t(const V* in, V* out, const V* op, const SegGeom& g, const TargetMap& stm,
                               bool conj_op, BlockDesc* d, cudaStream_t s) {
  PW_TRY(build_block_desc_t<V>(op, g.t, conj_op, stm, &d->mem, &d->arr, s));
  d->hdr = d->arr.hdr;
  d->staged = true;
  const size_t per_warp = (size_t)32 * g.pitch * sizeof(V);
  // descriptor copy (worst case for `usable` operators: t blocks, 2t offsets, 8t values)
  uint32_t desc_cap = (uint32_t)(g.t * 16 + 2 * g.t * 8 + g.t * 8 + 8 * g.t * sizeof(V) + 64);
  desc_cap = (desc_cap + 127u) & ~127u;
  int warps = (int)((110 * 1024 - desc_cap) / per_warp);
  if (warps < 1) return PW_ERR_UNSUPPORTED;
  if (warps > 8) warps = 8;
  const size_t smem = per_warp * warps + desc_cap;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    PW_CUDA(cudaFuncSetAttribute(k_apply_blocks_seg<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const uint64_t ntasks = g.outer.nfib * ((g.M + 31) / 32);
  int ctas = (int)((227 * 1024) / (smem + 1024));
  if (ctas > 8) ctas = 8;
  if (ctas < 1) ctas = 1;
  uint64_t grid = (ntasks + warps - 1) / warps;
  if (grid > (uint64_t)kNumSMs * ctas) grid = (uint64_t)kNumSMs * ctas;
  const BlockArrays& A = d->arr;
  k_apply_blocks_seg<V><<<(unsigned)grid, warps * 32, smem, s>>>(in, out, g, A, desc_cap);
  PW_LAUNCHED();
  g_paths[4].fetch_add(1, std::memory_order_relaxed);
  return PW_OK;
}

int launch_blocks_seg(const void* in, void* out, const void* op, const int64_t* axes, int naxes,
                      const int* targets, int nt, int dtype, bool conj_op, BlockDesc* desc,
                      cudaStream_t stream) {
  if (in == out) return PW_ERR_UNSUPPORTED;
  SegGeom g;
  TargetMap stm;
  if (!plan_seg(axes, naxes, targets, nt, &g, &stm)) return PW_ERR_UNSUPPORTED;
  if (dtype == PW_C128)
    return launch_blocks_seg_t<double2>((const double2*)in, (double2*)out, (const double2*)op, g, stm, conj_op, desc, stream);
  if (dtype == PW_C64)
    return launch_blocks_seg_t<float2>((const float2*)in, (float2*)out, (const float2*)op, g, stm, conj_op, desc, stream);
  return PW_ERR_INVALID;
}

// ---- two-sided single pass: rho' = U rho U^dagger ------------------------------------------
// The row side and the column side of a density-matrix update commute, but a thread that
// owns a row-side fiber cannot reach across the column side.  Here a CTA stages tiles of
// (row block k) x (column block l) x (32 adjacent fibers) through shared memory:
//
//   load     cp.async copies the tile (one coalesced 512 B row per operator index pair) into
//            shared memory -- for the NEXT tile while the current one is being computed;
//   stage 1  a warp takes one column of a block pair, multiplies its nc_k values by B_k and
//            writes the results back in place;
//   stage 2  a warp takes one row, multiplies by conj(B_l) and streams the outputs to HBM,
//            again one coalesced row per store.
//
// Every element of rho is read once and written once: ONE HBM pass for U rho U^dagger with a
// dense d <= 8 operator (one 8 x 8 block pair) or a block-structured two-mode operator (beam
// splitter: photon-number sectors, 15 x 15 block pairs for d = 8), against the two passes of
// the row-then-column route.  Small block pairs are packed into 64-slot groups so that the
// eight warps stay busy.  The innermost axis must not be a target (the 32 fibers of a tile are
// then adjacent in memory).  Structure is analysed on the device like for the one-sided block
// kernel; `flags` carries the outcome to the fallback kernels.
constexpr int kTsWarps = 8;
constexpr int kTsSlots = 64;                 // tile = 64 slots x 32 lanes
constexpr uint32_t kTsValCap = 8 * 1024;     // per side: block values (BS d=8: 344 x 16 B)

struct TsFlags { int handled; int any; int ngroups; int use_mma; };

struct TsTables {
  TsFlags* flags;
  int4* groups;        // 2 per group: (slot0, nslots, item1_0, nitems1), (item2_0, nitems2, -, -)
  int4* pairs;         // scratch of the builder: (k, l, slot0, group)
  int64_t* in_off;     // [t*t] element offset of every slot (input side)
  int64_t* out_off;    // [t*t] (output side)
  int4* item1;         // (slot in tile, n, stride, value offset)   stage 1: a column of a pair
  int4* item2;         // (slot in tile, n, value offset, -)        stage 2: a row of a pair
  int* in_row;         // [t*t] row a of every slot inside its block pair (tensor-core kernel: swizzle)
};

template <typename V>
__global__ void __launch_bounds__(256)
k_ts_build(const BlockArrays L, const BlockArrays R, const int* __restrict__ other, const TsTables T,
           const uint32_t val_cap, const int mma_ok) {
  __shared__ int s_handled, s_npairs;
  const BlockDescHeader* hl = L.hdr;
  const BlockDescHeader* hr = R.hdr;
  if (threadIdx.x == 0) {
    const int o = other ? *other : 0;
    int h = (!o && hl->usable && hr->usable && hl->nzero == 0 && hr->nzero == 0 && hl->maxn >= 2 &&
             hl->nblocks == hr->nblocks) ? 1 : 0;
    if (h && ((uint32_t)hl->n_vals * sizeof(V) > val_cap || (uint32_t)hr->n_vals * sizeof(V) > val_cap)) h = 0;
    if (h)
      for (int b = 0; b < hl->nblocks && h; ++b)
        if (L.blocks[b].x != L.blocks[b].y || R.blocks[b].x != R.blocks[b].y) h = 0;  // in place: square blocks
    int ng = 0;
    if (h) {
      // greedy packing of the block pairs (k major) into groups of <= kTsSlots slots
      const int m = hl->nblocks;
      int slot0 = 0, used = 0, gslot0 = 0, it1 = 0, it2 = 0, git1 = 0, git2 = 0, np = 0;
      for (int k = 0; k < m; ++k)
        for (int l = 0; l < m; ++l) {
          const int nk = L.blocks[k].y, nl = R.blocks[l].y, sz = nk * nl;
          if (used + sz > kTsSlots) {
            T.groups[2 * ng] = make_int4(gslot0, used, git1, it1 - git1);
            T.groups[2 * ng + 1] = make_int4(git2, it2 - git2, 0, 0);
            ++ng; gslot0 = slot0; used = 0; git1 = it1; git2 = it2;
          }
          T.pairs[np++] = make_int4(k, l, slot0, (it1 << 16) | it2);  // item offsets fit 16 bits (t <= 128)
          slot0 += sz; used += sz; it1 += nl; it2 += nk;
        }
      T.groups[2 * ng] = make_int4(gslot0, used, git1, it1 - git1);
      T.groups[2 * ng + 1] = make_int4(git2, it2 - git2, 0, 0);
      ++ng;
      s_npairs = np;
    }
    T.flags->handled = h;
    T.flags->any = (h || o) ? 1 : 0;
    T.flags->ngroups = ng;
    // one dense block (d = 5..8 single-mode operator): FP64 tensor cores; many small blocks
    // (beam splitter sectors) run faster on the scalar kernel (measured)
    T.flags->use_mma = (h && mma_ok && hl->nblocks == 1 && hl->maxn >= 5) ? 1 : 0;
    s_handled = h;
  }
  __syncthreads();
  if (!s_handled) return;
  // group start slot of every pair: recover by scanning groups (few); then fill the tables
  const int ng = T.flags->ngroups;
  for (int p = threadIdx.x; p < s_npairs; p += blockDim.x) {
    const int4 pr = T.pairs[p];
    const int k = pr.x, l = pr.y, slot0 = pr.z, it1 = pr.w >> 16, it2 = pr.w & 0xffff;
    int g = 0;
    while (g + 1 < ng && T.groups[2 * (g + 1)].x <= slot0) ++g;
    const int local0 = slot0 - T.groups[2 * g].x;
    const int4 hk = L.blocks[k], hq = R.blocks[l];
    const int nk = hk.y, nl = hq.y;
    const int64_t* ko = L.offs + hk.z;
    const int64_t* lo = R.offs + hq.z;
    for (int a = 0; a < nk; ++a)
      for (int b = 0; b < nl; ++b) {
        T.in_off[slot0 + a * nl + b] = ko[a] + lo[b];
        T.in_row[slot0 + a * nl + b] = a;
        T.out_off[slot0 + a * nl + b] = ko[nk + a] + lo[nl + b];
      }
    for (int j = 0; j < nl; ++j) T.item1[it1 + j] = make_int4(local0 + j, nk | (j << 8), nl, hk.w);
    for (int i2 = 0; i2 < nk; ++i2) T.item2[it2 + i2] = make_int4(local0 + i2 * nl, nl, hq.w, i2);
  }
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename V>
__global__ void __launch_bounds__(kTsWarps * 32, 2)
k_two_sided(const V* __restrict__ in, V* __restrict__ out, const FiberMap fm, const BlockArrays L,
            const BlockArrays R, const TsTables T) {
  if (!T.flags->handled || T.flags->use_mma) return;
  extern __shared__ __align__(16) unsigned char s_raw[];
  V* vl = (V*)s_raw;                       // row-side block values
  V* vr = (V*)(s_raw + kTsValCap);         // column-side block values (already conjugated)
  V* tile0 = (V*)(s_raw + 2 * kTsValCap);  // [2][kTsSlots][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < L.hdr->n_vals; i += blockDim.x) vl[i] = ((const V*)L.vals)[i];
  for (int i = threadIdx.x; i < R.hdr->n_vals; i += blockDim.x) vr[i] = ((const V*)R.vals)[i];
  const uint32_t G = (uint32_t)T.flags->ngroups;
  const uint64_t chunks = (fm.nfib + 31) / 32;
  const uint64_t units = chunks * G;

  auto issue = [&](uint64_t u, V* tile, int64_t* base_out) {
    // chunk index fastest: CTAs that run side by side work on adjacent memory with the same
    // slot offsets (DRAM page and TLB locality)
    const uint32_t g = (uint32_t)(u / chunks);
    const uint64_t c = u - (uint64_t)g * chunks;
    const uint64_t f = c * 32 + lane;
    const int64_t base = f < fm.nfib ? fm.base(f) : -1;
    *base_out = base;
    if (base >= 0) {
      const int4 g0 = __ldg(T.groups + 2 * g);
      const int64_t* off = T.in_off + g0.x;
      for (int sl = warp; sl < g0.y; sl += kTsWarps) {
        const V* src = in + base + __ldg(off + sl);
        if (sizeof(V) == 16) cp_async16(tile + sl * 32 + lane, src); else cp_async8(tile + sl * 32 + lane, src);
      }
    }
    cp_async_commit();
  };

  uint64_t u = blockIdx.x;
  int64_t base_cur = -1, base_next = -1;
  int cur = 0;
  if (u < units) issue(u, tile0, &base_cur);
  for (; u < units; u += gridDim.x) {
    V* tile = tile0 + (size_t)cur * kTsSlots * 32;
    const uint64_t un = u + gridDim.x;
    if (un < units) issue(un, tile0 + (size_t)(cur ^ 1) * kTsSlots * 32, &base_next);
    else cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t g = (uint32_t)(u / chunks);
    const int4 g0 = __ldg(T.groups + 2 * g), g1 = __ldg(T.groups + 2 * g + 1);
    if (base_cur >= 0) {
      for (int it = warp; it < g0.w; it += kTsWarps) {
        const int4 w = __ldg(T.item1 + g0.z + it);
        V* col = tile + w.x * 32 + lane;
        const int stride = w.z * 32;
        block_product<V, 4>(vl + w.w, w.y & 0xff, w.y & 0xff,
                         [&](int a) { return col[a * stride]; },
                         [&](int i, V v) { col[i * stride] = v; });
      }
    }
    __syncthreads();
    if (base_cur >= 0) {
      const int64_t* oo = T.out_off + g0.x;
      for (int it = warp; it < g1.y; it += kTsWarps) {
        const int4 w = __ldg(T.item2 + g1.x + it);
        const V* row = tile + w.x * 32 + lane;
        V* dst = out + base_cur;
        const int64_t* ro = oo + w.x;
        block_product<V, 4>(vr + w.z, w.y, w.y,
                         [&](int b) { return row[b * 32]; },
                         [&](int j, V v) { __stcg(dst + __ldg(ro + j), v); });
      }
    }
    __syncthreads();
    base_cur = base_next;
    cur ^= 1;
  }
  cp_async_wait<0>();
}

// ---- the same pass on the FP64 tensor cores (complex128) --------------------------------------
// The scalar kernel above is bound by what surrounds its FMAs: every complex FMA needs a
// broadcast shared-memory load of an operator entry, and 16 warps per SM cannot hide the FP64
// latency.  mma.sync.m8n8k4.f64 keeps the operator in registers as an A fragment (two complex
// values per lane) and performs 256 FMAs per instruction.  A complex product Y = B X is the real
// product [Yr; Yi] = [Br -Bi; Bi Br] [Xr; Xi]: M = 16 (two m8 tiles: real and imaginary rows),
// K = 16 (four k4 steps: real then imaginary inputs), N = 8 fibers per n-tile, four n-tiles per
// warp item.  The lane that owns (row gid, k/column pair tig) needs X at rows {tig, 4 + tig} of
// fiber gid: one 128-bit shared load yields the real AND imaginary part, i.e. the B fragments of
// two k-steps.  The tile is stored XOR-swizzled (fiber index ^ 2 (row & 3)) so that those loads
// are bank-conflict free.
__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Single dense block (n = 5..8): the slot of operator index pair (a, b) is a * n + b, so no
// item tables are needed; the operator fragments live in registers for the whole kernel.
// Three tile buffers: the loads run two tiles ahead and a tile costs two barriers.
__global__ void __launch_bounds__(kTsWarps * 32, 2)
k_two_sided_mma(const double2* __restrict__ in, double2* __restrict__ out, const FiberMap fm,
                const BlockArrays L, const BlockArrays R, const TsTables T) {
  using V = double2;
  if (!T.flags->handled || !T.flags->use_mma) return;
  extern __shared__ __align__(16) unsigned char s_raw[];
  V* tile0 = (V*)s_raw;                                          // [3][kTsSlots][32], swizzled
  int64_t* s_base = (int64_t*)(tile0 + 3 * kTsSlots * 32);       // [3][32] fiber bases (-1: past the end)
  int64_t* s_off = s_base + 3 * 32;                              // [4][8]: row in, column in, row out, column out
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = L.blocks[0].x;
  if (threadIdx.x < 4 * 8) {
    const int which = threadIdx.x >> 3, e = threadIdx.x & 7;
    int64_t v = 0;
    if (e < n) {
      const int64_t* ko = L.offs + L.blocks[0].z;
      const int64_t* lo = R.offs + R.blocks[0].z;
      v = which == 0 ? ko[e] : which == 1 ? lo[e] : which == 2 ? ko[n + e] : lo[n + e];
    }
    s_off[threadIdx.x] = v;
  }
  // operator fragments: lane (gid, tig) holds W[gid][tig] and W[gid][4 + tig]
  const int gid = lane >> 2, tig = lane & 3;
  V l0 = make_double2(0.0, 0.0), l1 = l0, r0 = l0, r1 = l0;
  if (gid < n) {
    const V* wl = (const V*)L.vals + L.blocks[0].w;
    const V* wr = (const V*)R.vals + R.blocks[0].w;
    if (tig < n) { l0 = wl[gid * n + tig]; r0 = wr[gid * n + tig]; }
    if (4 + tig < n) { l1 = wl[gid * n + 4 + tig]; r1 = wr[gid * n + 4 + tig]; }
  }
  __syncthreads();
  const uint64_t units = (fm.nfib + 31) / 32;
  const int nslots = n * n;

  auto issue = [&](uint64_t u, int buf) {
    if (u < units) {
      const uint64_t f = u * 32 + lane;
      const int64_t base = f < fm.nfib ? fm.base(f) : -1;
      if (warp == 0) s_base[buf * 32 + lane] = base;
      if (base >= 0) {
        V* tile = tile0 + (size_t)buf * kTsSlots * 32;
        for (int sl = warp; sl < nslots; sl += kTsWarps) {
          const int a = sl / n, b = sl - a * n;
          cp_async16(tile + sl * 32 + (lane ^ (2 * (a & 3))), in + base + s_off[a] + s_off[8 + b]);
        }
      }
    }
    cp_async_commit();
  };

  // one item: out[i][f] = sum_a W[i][a] x[a][f]; fragments w0 / w1 precomputed
  auto product = [&](const V w0, const V w1, auto ldx, auto sty) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int f = 8 * nt + gid;
      V x0 = make_double2(0.0, 0.0), x1 = x0;
      if (tig < n) x0 = ldx(tig, f);
      if (4 + tig < n) x1 = ldx(4 + tig, f);
      double re0 = 0.0, re1 = 0.0, im0 = 0.0, im1 = 0.0;
      dmma8x8x4(re0, re1, w0.x, x0.x);
      dmma8x8x4(im0, im1, w0.y, x0.x);
      dmma8x8x4(re0, re1, -w0.y, x0.y);
      dmma8x8x4(im0, im1, w0.x, x0.y);
      dmma8x8x4(re0, re1, w1.x, x1.x);
      dmma8x8x4(im0, im1, w1.y, x1.x);
      dmma8x8x4(re0, re1, -w1.y, x1.y);
      dmma8x8x4(im0, im1, w1.x, x1.y);
      sty(nt, 8 * nt + 2 * tig, make_double2(re0, im0), make_double2(re1, im1));
    }
  };

  uint64_t u = blockIdx.x;
  int cur = 0;
  issue(u, 0);
  issue(u + gridDim.x, 1);
  for (; u < units; u += gridDim.x) {
    V* tile = tile0 + (size_t)cur * kTsSlots * 32;
    cp_async_wait<1>();
    __syncthreads();
    issue(u + 2 * (uint64_t)gridDim.x, cur == 0 ? 2 : cur - 1);  // the buffer of the previous tile
    // stage 1: column j of the tile, in place
    for (int j = warp; j < n; j += kTsWarps) {
      V* col = tile + j * 32;
      V y0[4], y1[4];
      product(l0, l1, [&](int a, int f) { return col[a * n * 32 + (f ^ (2 * (a & 3)))]; },
              [&](int nt, int f, V v0, V v1) { y0[nt] = v0; y1[nt] = v1; (void)f; });
      __syncwarp();  // every lane has read its inputs before the column is overwritten
      if (gid < n) {
        V* dst = col + gid * n * 32;
        const int sw = (2 * (j & 3)) ^ (gid & 1);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int f = 8 * nt + 2 * tig;
          dst[f ^ sw] = y0[nt];
          dst[(f + 1) ^ sw] = y1[nt];
        }
      }
    }
    __syncthreads();
    // stage 2: row i of the tile, streamed to HBM
    const int64_t* sb = s_base + cur * 32;
    for (int i = warp; i < n; i += kTsWarps) {
      const V* row = tile + i * n * 32;
      const int64_t ro = s_off[16 + i] + (gid < n ? s_off[24 + gid] : 0);
      product(r0, r1, [&](int b, int f) { return row[b * 32 + (f ^ (2 * (b & 3)) ^ (i & 1))]; },
              [&](int nt, int f, V v0, V v1) {
                (void)nt;
                if (gid < n) {
                  const int64_t b0 = sb[f], b1 = sb[f + 1];
                  if (b0 >= 0) __stcg(out + b0 + ro, v0);
                  if (b1 >= 0) __stcg(out + b1 + ro, v1);
                }
              });
    }
    cur = cur == 2 ? 0 : cur + 1;
  }
  cp_async_wait<0>();
}

template <typename V>
static int launch_two_sided_t(const V* rho, V* out, const V* op, const FiberMap& fm,
                              const TargetMap& rows, const TargetMap& cols, const int* other,
                              TwoSided* ts, cudaStream_t s) {
  PW_TRY(build_block_desc_t<V>(op, rows.t, false, rows, &ts->memL, &ts->L, s));
  PW_TRY(build_block_desc_t<V>(op, cols.t, true, cols, &ts->memR, &ts->R, s));
  const size_t t = rows.t, tt = t * t;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_flags = 256, b_groups = al(sizeof(int4) * 2 * tt), b_pairs = al(sizeof(int4) * tt),
               b_off = al(sizeof(int64_t) * tt), b_item = al(sizeof(int4) * tt), b_row = al(sizeof(int) * tt);
  PW_TRY(ts->flag_mem.alloc(b_flags + b_groups + b_pairs + 2 * b_off + 2 * b_item + b_row, s));
  char* base = (char*)ts->flag_mem.p;
  TsTables T;
  T.flags = (TsFlags*)base;
  T.groups = (int4*)(base + b_flags);
  T.pairs = (int4*)(base + b_flags + b_groups);
  T.in_off = (int64_t*)(base + b_flags + b_groups + b_pairs);
  T.out_off = (int64_t*)(base + b_flags + b_groups + b_pairs + b_off);
  T.item1 = (int4*)(base + b_flags + b_groups + b_pairs + 2 * b_off);
  T.item2 = (int4*)(base + b_flags + b_groups + b_pairs + 2 * b_off + b_item);
  T.in_row = (int*)(base + b_flags + b_groups + b_pairs + 2 * b_off + 2 * b_item);
  k_ts_build<V><<<1, 256, 0, s>>>(ts->L, ts->R, other, T, kTsValCap,
                                  (sizeof(V) == 16 && options().ts_mma.load(std::memory_order_relaxed)) ? 1 : 0);
  PW_LAUNCHED();
  const size_t smem = 2 * kTsValCap + sizeof(V) * 2 * kTsSlots * 32;
  static bool attr_set = false;
  if (!attr_set) {
    PW_CUDA(cudaFuncSetAttribute(k_two_sided<double2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(2 * kTsValCap + sizeof(double2) * 2 * kTsSlots * 32)));
    PW_CUDA(cudaFuncSetAttribute(k_two_sided<float2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(2 * kTsValCap + sizeof(float2) * 2 * kTsSlots * 32)));
    attr_set = true;
  }
  const int per_sm = options().ts_ctas.load(std::memory_order_relaxed);
  const int grid = kNumSMs * (per_sm > 0 ? per_sm : 2);
  k_two_sided<V><<<grid, kTsWarps * 32, smem, s>>>(rho, out, fm, ts->L, ts->R, T);
  if (sizeof(V) == 16 && options().ts_mma.load(std::memory_order_relaxed)) {
    PW_LAUNCHED();
    static bool attr_mma = false;
    const size_t smem_mma = sizeof(double2) * 3 * kTsSlots * 32 + (3 * 32 + 4 * 8) * sizeof(int64_t);
    if (!attr_mma) {
      PW_CUDA(cudaFuncSetAttribute(k_two_sided_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma));
      attr_mma = true;
    }
    k_two_sided_mma<<<grid, kTsWarps * 32, smem_mma, s>>>((const double2*)rho, (double2*)out, fm, ts->L, ts->R, T);
  }
  PW_LAUNCHED();
  g_paths[4].fetch_add(1, std::memory_order_relaxed);
  ts->skip = &T.flags->any;
  return PW_OK;
}

int launch_two_sided(const void* rho, void* out, const void* op, const FiberMap& fm,
                     const TargetMap& rows, const TargetMap& cols, int dtype, const int* other_handled,
                     TwoSided* ts, cudaStream_t stream) {
  if (rho == out || rows.t != cols.t || rows.t > kBlockMaxT) return PW_ERR_UNSUPPORTED;
  if (dtype == PW_C128)
    return launch_two_sided_t<double2>((const double2*)rho, (double2*)out, (const double2*)op, fm, rows,
                                       cols, other_handled, ts, stream);
  if (dtype == PW_C64)
    return launch_two_sided_t<float2>((const float2*)rho, (float2*)out, (const float2*)op, fm, rows,
                                      cols, other_handled, ts, stream);
  return PW_ERR_INVALID;
}

}  // namespace pw
