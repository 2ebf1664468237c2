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
#include "pw_tf32.cuh"

#include <type_traits>

namespace pw {

// ---- preparation ---------------------------------------------------------------------
// Single CTA.  Label propagation on the bipartite graph (row a -- column b iff O[a,b] != 0)
// converges to label = smallest column index of the component.
template <typename V>
__global__ void __launch_bounds__(256)
k_build_blocks(const V* __restrict__ op, const uint32_t t, const bool conj0, const TargetMap tm0,
               const BlockArrays A0, const bool conj1, const TargetMap tm1, const BlockArrays A1,
               const int* __restrict__ skip) {
  // one CTA per descriptor: gridDim.x = 2 analyses the same operator for two offset maps at once
  // (row side and conjugated column side of U rho U^dagger)
  const bool second = blockIdx.x == 1;
  const bool conj_op = second ? conj1 : conj0;
  const TargetMap& tm = second ? tm1 : tm0;
  const BlockArrays& A = second ? A1 : A0;
  BlockDescHeader* __restrict__ hdr = A.hdr;
  int4* __restrict__ blocks = A.blocks;
  int64_t* __restrict__ offs = A.offs;
  V* __restrict__ vals = (V*)A.vals;
  int64_t* __restrict__ zero_offs = A.zero_offs;
  extern __shared__ __align__(16) int32_t s_i[];
  int4* s_blk = (int4*)s_i;    // [t] block table (copied out at the end)
  int32_t* lr = s_i + 4 * t;   // [t] row labels
  int32_t* lc = lr + t;        // [t] column labels
  int32_t* bid = lc + t;       // [t] label -> block id
  uint32_t* nzr = (uint32_t*)(bid + t);   // [t][4] nonzero columns of every row (bit masks, t <= 128)
  uint32_t* nzc = nzr + 4 * t;            // [t][4] nonzero rows of every column
  int32_t* s_rc = (int32_t*)(nzc + 4 * t);  // [t] rank of column c inside its block
  int32_t* s_rr = s_rc + t;               // [t] rank of row a inside its block
  __shared__ int s_changed;
  __shared__ int s_nblocks, s_maxn, s_nzero;
  const int32_t INF = 0x7fffffff;
  if (skip != nullptr && *skip != 0) {
    // another path already produced the result: nobody may take this descriptor
    if (threadIdx.x == 0) {
      hdr->nblocks = 0; hdr->maxn = 0; hdr->nzero = 0; hdr->usable = 0; hdr->inplace_safe = 0;
      hdr->seg_handled = 0; hdr->n_offs = 0; hdr->n_vals = 0;
    }
    return;
  }
  for (uint32_t i = threadIdx.x; i < 8 * t; i += blockDim.x) nzr[i] = 0;
  for (uint32_t i = threadIdx.x; i < t; i += blockDim.x) { lr[i] = INF; lc[i] = (int32_t)i; }
  __syncthreads();
  // ONE coalesced pass over the operator: the structure analysis runs on bit masks
  for (uint32_t e = threadIdx.x; e < t * t; e += blockDim.x) {
    const V v = op[e];
    if (v.x != 0 || v.y != 0) {
      const uint32_t a = e / t, b = e - a * t;
      atomicOr(&nzr[4 * a + (b >> 5)], 1u << (b & 31));
      atomicOr(&nzc[4 * b + (a >> 5)], 1u << (a & 31));
    }
  }
  __syncthreads();
  for (uint32_t iter = 0; iter < 2 * t + 2; ++iter) {
    if (threadIdx.x == 0) s_changed = 0;
    __syncthreads();
    for (uint32_t a = threadIdx.x; a < t; a += blockDim.x) {
      int32_t m = lr[a];
      for (uint32_t w = 0; w < 4; ++w)
        for (uint32_t bits = nzr[4 * a + w]; bits; bits &= bits - 1)
          m = min(m, lc[32 * w + __ffs(bits) - 1]);
      if (m != lr[a]) { lr[a] = m; s_changed = 1; }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < t; b += blockDim.x) {
      int32_t m = lc[b];
      for (uint32_t w = 0; w < 4; ++w)
        for (uint32_t bits = nzc[4 * b + w]; bits; bits &= bits - 1)
          m = min(m, lr[32 * w + __ffs(bits) - 1]);
      if (m != lc[b]) { lc[b] = m; s_changed = 1; }
    }
    __syncthreads();
    if (!s_changed) break;
    __syncthreads();
  }
  // columns whose label is their own index AND that touch at least one row open a block
  for (uint32_t b = threadIdx.x; b < t; b += blockDim.x) bid[b] = -1;
  __syncthreads();
  for (uint32_t a = threadIdx.x; a < t; a += blockDim.x)
    if (lr[a] != INF) bid[lr[a]] = -2;  // the component of this label has rows
  __syncthreads();
  if (threadIdx.x == 0) {
    int nb = 0;
    for (uint32_t b = 0; b < t; ++b) bid[b] = (lc[b] == (int32_t)b && bid[b] == -2) ? nb++ : -1;
    s_nblocks = nb;
    s_maxn = 0;
    int nz = 0;
    for (uint32_t a = 0; a < t; ++a)
      if (lr[a] == INF) zero_offs[nz++] = tm.offset(a);
    s_nzero = nz;
  }
  __syncthreads();
  // sizes of every block (one thread per component root)
  for (uint32_t root = threadIdx.x; root < t; root += blockDim.x) {
    const int b = bid[root];
    if (b < 0) continue;
    int nc = 0, nr = 0;
    for (uint32_t c = 0; c < t; ++c) nc += (lc[c] == (int32_t)root);
    for (uint32_t a = 0; a < t; ++a) nr += (lr[a] == (int32_t)root);
    atomicMax(&s_maxn, max(nr, nc));
    s_blk[b] = make_int4(nr, nc, 0, 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // exclusive scans for the packed layouts: s_blk[b] = (nr, nc, index offset, value offset)
    int io = 0, vo = 0;
    for (int b = 0; b < s_nblocks; ++b) {
      const int4 h = s_blk[b];
      s_blk[b] = make_int4(h.x, h.y, io, vo);
      io += h.x + h.y;
      vo += h.x * h.y;
    }
    hdr->nblocks = s_nblocks;
    hdr->n_offs = io;
    hdr->n_vals = (s_maxn <= kBlockMaxN) ? vo : 0;
    hdr->maxn = s_maxn;
    hdr->nzero = s_nzero;
    const int usable = (s_maxn <= kBlockMaxN) ? 1 : 0;
    hdr->usable = usable;
    // position a is written by the block of ROW a and read by the block of COLUMN a:
    // in-place is safe iff those coincide wherever both exist (zero rows write too)
    int safe = 1;
    for (uint32_t a = 0; a < t; ++a) {
      const bool col_used = bid[lc[a]] >= 0;   // the component of column a has rows
      if (!col_used) continue;                 // nobody reads position a
      if (lr[a] != lc[a]) safe = 0;            // written by another block (or cleared)
    }
    hdr->inplace_safe = safe;
    hdr->seg_handled = (usable && safe) ? 1 : 0;
  }
  __syncthreads();
  // rank of every column / row inside its block (tensor order), and the element offsets:
  // columns first, then rows
  for (uint32_t i = threadIdx.x; i < t; i += blockDim.x) {
    int rc = -1, rr = -1;
    const int bc = bid[lc[i]];
    if (bc >= 0) {
      rc = 0;
      for (uint32_t c = 0; c < i; ++c) rc += (lc[c] == lc[i]);
      offs[s_blk[bc].z + rc] = tm.offset(i);
    }
    if (lr[i] != INF) {
      const int4 h = s_blk[bid[lr[i]]];
      rr = 0;
      for (uint32_t a = 0; a < i; ++a) rr += (lr[a] == lr[i]);
      offs[h.z + h.y + rr] = tm.offset(i);
    }
    s_rc[i] = rc;
    s_rr[i] = rr;
  }
  __syncthreads();
  // dense block values: one coalesced pass over the operator
  if (s_maxn <= kBlockMaxN)
    for (uint32_t e = threadIdx.x; e < t * t; e += blockDim.x) {
      const uint32_t a = e / t, c = e - a * t;
      if (lr[a] == INF || lr[a] != lc[c]) continue;
      const int4 h = s_blk[bid[lr[a]]];
      V v = op[e];
      if (conj_op) v.y = -v.y;
      vals[h.w + s_rr[a] * h.y + s_rc[c]] = v;
    }
  for (int b = threadIdx.x; b < s_nblocks; b += blockDim.x) blocks[b] = s_blk[b];
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

// complex64, contiguous untouched run of EVEN length: a thread owns two adjacent fibers and moves
// them as one 128-bit access (the complex128 kernel's bytes per thread and per instruction; with
// 8-byte accesses the complex64 kernel sat at 60-72 % of the roof).  `fm` counts fiber PAIRS.
__global__ void __launch_bounds__(256, 3)
k_apply_blocks_c64x2(const float2* __restrict__ in, float2* __restrict__ out, const FiberMap fm,
                     const BlockArrays A) {
  extern __shared__ __align__(16) unsigned char s_desc[];
  if (!A.hdr->usable) return;
  const DescView<float2> dv = desc_view<float2>(A, s_desc, kDescSmemCap);
  const bool scalar_blocks = A.hdr->maxn == 1;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib; f += stride) {
    const int64_t base = fm.base(f);
    const float2* src = in + base;
    float2* dst = out + base;
    if (scalar_blocks) {
      int b = 0;
      for (; b + 4 <= dv.nblocks; b += 4) {
        C64x2 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = ld_c64x2(src + dv.offs[2 * (b + u)]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          C64x2 acc = czero<C64x2>();
          cfma(acc, dv.vals[b + u], x[u]);
          st_c64x2(dst + dv.offs[2 * (b + u) + 1], acc);
        }
      }
      for (; b < dv.nblocks; ++b) {
        C64x2 acc = czero<C64x2>();
        cfma(acc, dv.vals[b], ld_c64x2(src + dv.offs[2 * b]));
        st_c64x2(dst + dv.offs[2 * b + 1], acc);
      }
    } else {
      for (int b = 0; b < dv.nblocks; ++b) {
        const int4 h = dv.blocks[b];
        const int64_t* bo = dv.offs + h.z;
        block_product<float2>(dv.vals + h.w, h.x, h.y,
                              [&](int j) { return ld_c64x2(src + bo[j]); },
                              [&](int a, C64x2 v) { st_c64x2(dst + bo[h.y + a], v); });
      }
    }
    for (int z = 0; z < dv.nzero; ++z) st_c64x2(dst + dv.zero_offs[z], czero<C64x2>());
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
                   const BlockArrays A, const uint32_t desc_cap, const uint32_t fpw,
                   const int* __restrict__ mma_taken) {
  if (!A.hdr->seg_handled || (mma_taken != nullptr && *mma_taken != 0)) return;
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const DescView<V> dv = desc_view<V>(A, s_raw, desc_cap);
  // a warp stages `fpw` (32, 16 or 8) adjacent fibers; 32 / fpw lanes share a fiber and take
  // every (32 / fpw)-th block (blocks are independent): smaller stages -> more resident warps
  // to overlap one warp's load -> compute -> store sequence with the others'
  const uint32_t lpf = 32 / fpw;
  const uint32_t myfib = (uint32_t)lane / lpf, sub = (uint32_t)lane - myfib * lpf;
  V* buf = (V*)(s_raw + desc_cap) + (size_t)warp * fpw * g.pitch;
  const uint64_t chunks = (g.M + fpw - 1) / fpw;
  const uint64_t ntasks = g.outer.nfib * chunks;
  // lane-constant decomposition of the coalesced copy index e = it*32 + lane into
  // (fiber q, position r) without divisions: each iteration adds 32 = qa*lseg + rb
  const uint32_t qa = 32 / g.lseg, rb = 32 % g.lseg;
  for (uint64_t task = (uint64_t)blockIdx.x * nwarps + warp; task < ntasks;
       task += (uint64_t)gridDim.x * nwarps) {
    const uint64_t o = task / chunks, c = task - o * chunks;
    const uint64_t m0 = c * fpw;
    const uint32_t nvalid = (uint32_t)((g.M - m0) < fpw ? (g.M - m0) : fpw);
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
    // ---- compute: blocks applied in place
    if (myfib < nvalid) {
      V* fib = buf + (size_t)myfib * g.pitch;
      for (int b = (int)sub; b < dv.nblocks; b += (int)lpf) {
        const int4 h = dv.blocks[b];
        const int64_t* bo = dv.offs + h.z;
        block_product<V>(dv.vals + h.w, h.x, h.y,
                         [&](int j) { return fib[(int32_t)bo[j]]; },
                         [&](int a, V v) { fib[(int32_t)bo[h.y + a]] = v; });
      }
      for (int z = (int)sub; z < dv.nzero; z += (int)lpf) fib[(int32_t)dv.zero_offs[z]] = czero<V>();
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
static int alloc_block_desc(uint32_t t, Scratch* mem, BlockArrays* arr, cudaStream_t s) {
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
  return PW_OK;
}

template <typename V>
static int build_block_desc_t(const V* op, uint32_t t, bool conj_op, const TargetMap& tm,
                              Scratch* mem, BlockArrays* arr, cudaStream_t s, const int* skip = nullptr) {
  PW_TRY(alloc_block_desc<V>(t, mem, arr, s));
  if (plan_replay()) return PW_OK;   // cached plan: the descriptor is already built
  k_build_blocks<V><<<1, 256, sizeof(int32_t) * 20 * t, s>>>(op, t, conj_op, tm, *arr, conj_op, tm, *arr, skip);
  PW_LAUNCHED();
  return PW_OK;
}

// two descriptors of the same operator in ONE launch (two CTAs)
template <typename V>
static int build_block_desc_pair_t(const V* op, uint32_t t, bool conj0, const TargetMap& tm0, Scratch* mem0,
                                   BlockArrays* arr0, bool conj1, const TargetMap& tm1, Scratch* mem1,
                                   BlockArrays* arr1, cudaStream_t s) {
  PW_TRY(alloc_block_desc<V>(t, mem0, arr0, s));
  PW_TRY(alloc_block_desc<V>(t, mem1, arr1, s));
  if (plan_replay()) return PW_OK;
  k_build_blocks<V><<<2, 256, sizeof(int32_t) * 20 * t, s>>>(op, t, conj0, tm0, *arr0, conj1, tm1, *arr1, nullptr);
  PW_LAUNCHED();
  return PW_OK;
}

int build_block_desc(const void* op, uint32_t t, bool conj_op, const TargetMap& tm, int dtype,
                     Scratch* mem, BlockArrays* arr, cudaStream_t stream, const int* skip) {
  if (t > kBlockMaxT) return PW_ERR_UNSUPPORTED;
  if (dtype == PW_C128) return build_block_desc_t<double2>((const double2*)op, t, conj_op, tm, mem, arr, stream, skip);
  if (dtype == PW_C64) return build_block_desc_t<float2>((const float2*)op, t, conj_op, tm, mem, arr, stream, skip);
  return PW_ERR_INVALID;
}

template <typename V>
static int launch_blocks_t(const V* in, V* out, const V* op, const Plan& p, bool conj_op,
                           BlockDesc* d, cudaStream_t s) {
  PW_TRY(build_block_desc_t<V>(op, p.tm.t, conj_op, p.tm, &d->mem, &d->arr, s));
  d->hdr = d->arr.hdr;
  const BlockArrays& A = d->arr;
  FiberMap pm;
  if (sizeof(V) == 8 && options().c64_pairs.load(std::memory_order_relaxed) && pair_fiber_map(p.fm, in, out, &pm)) {
    const int grid = grid_for(pm.nfib, 256, 3);
    k_apply_blocks_c64x2<<<grid, 256, kDescSmemCap, s>>>((const float2*)in, (float2*)out, pm, A);
  } else {
    const int grid = grid_for(p.fm.nfib, 256, 3);
    k_apply_blocks<V><<<grid, 256, kDescSmemCap, s>>>(in, out, p.fm, A);
  }
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

// tensor-core variant of the staged kernel (complex128; defined at the end of this file):
// enqueues the super-block packing and the kernel; *mma_taken = device flag the scalar kernel
// must honour (nullptr when the variant does not apply)
static int enqueue_seg_mma(const void* in, void* out, const SegGeom& g, BlockDesc* d, int dtype,
                           cudaStream_t s, const int** mma_taken);

template <typename V>
static int launch_blocks_seg_t(const V* in, V* out, const V* op, const SegGeom& g, const TargetMap& stm,
                               bool conj_op, BlockDesc* d, cudaStream_t s, bool mma_only) {
  PW_TRY(build_block_desc_t<V>(op, g.t, conj_op, stm, &d->mem, &d->arr, s));
  d->hdr = d->arr.hdr;
  d->staged = true;
  const int* mma_taken = nullptr;
  PW_TRY(enqueue_seg_mma(in, out, g, d, sizeof(V) == 16 ? PW_C128 : PW_C64, s, &mma_taken));
  if (mma_only) {
    if (mma_taken == nullptr) return PW_ERR_UNSUPPORTED;
    d->mma_flag = mma_taken;
    g_paths[4].fetch_add(1, std::memory_order_relaxed);
    return PW_OK;
  }
  // descriptor copy (worst case for `usable` operators: t blocks, 2t offsets, 8t values)
  uint32_t desc_cap = (uint32_t)(g.t * 16 + 2 * g.t * 8 + g.t * 8 + 8 * g.t * sizeof(V) + 64);
  desc_cap = (desc_cap + 127u) & ~127u;
  // fibers per warp stage: 32 (measured: 16 / 8 with lanes sharing a fiber buy more resident
  // warps but the lanes' different block shapes diverge -- 10.5 / 16.4 ms against 10.7 ms for the
  // loss channel on the innermost mode of rho, 16.9 / 26.8 against 15.1 ms for the beam splitter
  // on the two innermost modes of the vector)
  uint32_t fpw = 32;
  const int forced = options().seg_fpw.load(std::memory_order_relaxed);
  if (forced == 32 || forced == 16 || forced == 8) fpw = (uint32_t)forced;
  const size_t per_warp = (size_t)fpw * g.pitch * sizeof(V);
  int warps = (int)((110 * 1024 - desc_cap) / per_warp);
  if (warps < 1) return PW_ERR_UNSUPPORTED;
  if (warps > 8) warps = 8;
  const size_t smem = per_warp * warps + desc_cap;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    PW_CUDA(cudaFuncSetAttribute(k_apply_blocks_seg<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const uint64_t ntasks = g.outer.nfib * ((g.M + fpw - 1) / fpw);
  int ctas = (int)((227 * 1024) / (smem + 1024));
  if (ctas > 8) ctas = 8;
  if (ctas < 1) ctas = 1;
  uint64_t grid = (ntasks + warps - 1) / warps;
  if (grid > (uint64_t)kNumSMs * ctas) grid = (uint64_t)kNumSMs * ctas;
  const BlockArrays& A = d->arr;
  k_apply_blocks_seg<V><<<(unsigned)grid, warps * 32, smem, s>>>(in, out, g, A, desc_cap, fpw, mma_taken);
  PW_LAUNCHED();
  g_paths[4].fetch_add(1, std::memory_order_relaxed);
  return PW_OK;
}

int launch_blocks_seg(const void* in, void* out, const void* op, const int64_t* axes, int naxes,
                      const int* targets, int nt, int dtype, bool conj_op, BlockDesc* desc,
                      cudaStream_t stream, bool mma_only) {
  if (in == out) return PW_ERR_UNSUPPORTED;
  SegGeom g;
  TargetMap stm;
  if (!plan_seg(axes, naxes, targets, nt, &g, &stm)) return PW_ERR_UNSUPPORTED;
  if (dtype == PW_C128)
    return launch_blocks_seg_t<double2>((const double2*)in, (double2*)out, (const double2*)op, g, stm, conj_op, desc, stream, mma_only);
  if (dtype == PW_C64)
    return launch_blocks_seg_t<float2>((const float2*)in, (float2*)out, (const float2*)op, g, stm, conj_op, desc, stream, mma_only);
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

constexpr int kTsMaxSb = 16;                 // tensor-core kernel: super-blocks per side
constexpr int kTwMaxSb = 8;                  // row-tile variant (innermost column targets)

struct TsFlags { int handled; int any; int ngroups; int use_mma; int nsb; };

struct TsTables {
  TsFlags* flags;
  int4* groups;        // 2 per group: (slot0, nslots, item1_0, nitems1), (item2_0, nitems2, -, -)
  int4* pairs;         // scratch of the builder: (k, l, slot0, group)
  int64_t* in_off;     // [t*t] element offset of every slot (input side)
  int64_t* out_off;    // [t*t] (output side)
  int4* item1;         // (slot in tile, n, stride, value offset)   stage 1: a column of a pair
  int4* item2;         // (slot in tile, n, value offset, -)        stage 2: a row of a pair
  int* in_row;         // [t*t] row a of every slot inside its block pair (tensor-core kernel: swizzle)
  // tensor-core kernel: the blocks of one side packed into super-blocks of <= 8 indices
  int* sb_n;           // [kTsMaxSb] indices per super-block
  int2* sb_blk;        // [t] per block: (super-block, first index inside it)
  int64_t* sb_off;     // [4][kTsMaxSb][8] element offsets: row in, column in, row out, column out
  void* sb_w;          // [2][kTsMaxSb][8][8] zero-padded block-diagonal values: row side, column side (conjugated)
};

template <typename V>
__global__ void __launch_bounds__(256)
k_ts_build(const BlockArrays L, const BlockArrays R, const int* __restrict__ other, const TsTables T,
           const uint32_t val_cap, const int mma_ok, const int inner) {
  __shared__ int s_handled, s_npairs, s_mma, s_square;
  __shared__ int s_bn[kBlockMaxT];  // block sides (thread 0 below must not walk global memory)
  const BlockDescHeader* hl = L.hdr;
  const BlockDescHeader* hr = R.hdr;
  if (threadIdx.x == 0) s_square = 1;
  __syncthreads();
  {
    const int m = hl->nblocks == hr->nblocks ? min(hl->nblocks, (int)kBlockMaxT) : 0;
    for (int b = threadIdx.x; b < m; b += blockDim.x) {
      const int4 x = L.blocks[b], y = R.blocks[b];
      s_bn[b] = x.x;
      if (x.x != x.y || y.x != y.y || x.x != y.x) s_square = 0;  // in place: square blocks
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int o = other ? *other : 0;
    int h = (!o && hl->usable && hr->usable && hl->nzero == 0 && hr->nzero == 0 && hl->maxn >= 2 &&
             hl->nblocks == hr->nblocks) ? 1 : 0;
    if (h && ((uint32_t)hl->n_vals * sizeof(V) > val_cap || (uint32_t)hr->n_vals * sizeof(V) > val_cap)) h = 0;
    if (!s_square) h = 0;
    // tensor cores: first-fit-decreasing packing of the blocks into super-blocks of <= 8
    // indices (beam splitter, cutoff 8: sectors 1..8..1 -> 8 | 7+1 | 7+1 | 6+2 | 6+2 | 5+3 |
    // 5+3 | 4+4, i.e. eight full 8 x 8 tensor-core operands, no padding work)
    int nsb = 0, mma = 0;
    if (h && mma_ok && hl->maxn <= 8) {
      const int m = hl->nblocks;
      int total = 0, fill[kTsMaxSb];
      mma = 1;
      for (int sz = 8; sz >= 1 && mma; --sz)
        for (int b = 0; b < m && mma; ++b) {
          if (s_bn[b] != sz) continue;
          int s = 0;
          while (s < nsb && fill[s] + sz > 8) ++s;
          if (s == nsb) {
            if (nsb == kTsMaxSb) { mma = 0; break; }
            fill[nsb++] = 0;
          }
          T.sb_blk[b] = make_int2(s, fill[s]);
          fill[s] += sz;
          total += sz;
        }
      if (total < 5) mma = 0;  // tiny operators: the scalar kernel / register tiles
      for (int s = 0; s < nsb && mma; ++s) T.sb_n[s] = fill[s];
    }
    // innermost (contiguous) column target: only the tensor-core kernel has the transposing
    // load, and only for ONE dense block whose columns are the contiguous run itself
    // (inner bit 1: the column targets ARE the innermost run -> row-tile kernel; bit 2: one
    //  5..8-level column target with whole chunks of 32 fibers next to it -> 32-fiber tiles)
    if (inner) {
      if (mma && hl->nblocks == 1 && (inner & 2)) mma = 1;
      else if (mma && nsb <= kTwMaxSb && hr->inplace_safe) mma = 2;
      else h = 0;
    }
    if (!h) mma = 0;
    int ng = 0;
    if (h && !mma) {
      // greedy packing of the block pairs (k major) into groups of <= kTsSlots slots
      const int m = hl->nblocks;
      int slot0 = 0, used = 0, gslot0 = 0, it1 = 0, it2 = 0, git1 = 0, git2 = 0, np = 0;
      for (int k = 0; k < m; ++k)
        for (int l = 0; l < m; ++l) {
          const int nk = s_bn[k], nl = s_bn[l], sz = nk * nl;
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
    T.flags->use_mma = mma;
    T.flags->nsb = nsb;
    s_handled = h;
    s_mma = mma;
  }
  __syncthreads();
  if (!s_handled) return;
  if (s_mma) {
    // super-block tables: one thread per (super-block, index) row -- block and local row come
    // from a shared map filled by one thread per block
    __shared__ short s_map[kTsMaxSb * 8][2];
    V* w = (V*)T.sb_w;
    for (int i = threadIdx.x; i < 2 * kTsMaxSb * 64; i += blockDim.x) w[i] = czero<V>();
    for (int i = threadIdx.x; i < 4 * kTsMaxSb * 8; i += blockDim.x) T.sb_off[i] = 0;
    for (int i = threadIdx.x; i < kTsMaxSb * 8; i += blockDim.x) s_map[i][0] = -1;
    __syncthreads();
    for (int b = threadIdx.x; b < hl->nblocks; b += blockDim.x) {
      const int2 sp = T.sb_blk[b];
      for (int a = 0; a < s_bn[b]; ++a) {
        s_map[sp.x * 8 + sp.y + a][0] = (short)b;
        s_map[sp.x * 8 + sp.y + a][1] = (short)a;
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kTsMaxSb * 8; e += blockDim.x) {
      const int b = s_map[e][0], a = s_map[e][1];
      if (b < 0) continue;
      const int4 hk = L.blocks[b], hq = R.blocks[b];
      const int n = hk.x, c0 = (e & 7) - a;  // first index of the block inside its super-block
      const int64_t* ko = L.offs + hk.z;
      const int64_t* lo = R.offs + hq.z;
      const V* vl = (const V*)L.vals + hk.w + a * n;
      const V* vr = (const V*)R.vals + hq.w + a * n;
      T.sb_off[0 * kTsMaxSb * 8 + e] = ko[a];
      T.sb_off[1 * kTsMaxSb * 8 + e] = lo[a];
      T.sb_off[2 * kTsMaxSb * 8 + e] = ko[n + a];
      T.sb_off[3 * kTsMaxSb * 8 + e] = lo[n + a];
      for (int c = 0; c < n; ++c) {
        w[e * 8 + c0 + c] = vl[c];
        w[(kTsMaxSb * 8 + e) * 8 + c0 + c] = vr[c];
      }
    }
    return;
  }
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
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Two complex128 values that are usually adjacent in memory (the two fibers / positions of an
// accumulator fragment): ONE 256-bit store (sm_100: STG.E.256) when they are, two 128-bit stores
// otherwise.
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

// Complex 8 x 8 products of NB independent items, issued STEP-major: the 2 NB accumulator chains
// advance together, so that consecutive DMMAs in program order are independent (a warp issues in
// order).  Measured: 6 % faster than item-major order in the row-tile kernel, no difference in
// the staged and dense kernels, 20 % SLOWER in k_two_sided_mma (there the loads of the next
// fiber octet overlap the DMMAs of the current one) -- the tensor pipe is not the bottleneck.
// Lane (gid, tig): w0 = W[gid][tig], w1 = W[gid][4 + tig]; x0 = X[tig][n = gid], x1 = X[4 + tig][gid];
// y0 = Y[gid][2 tig], y1 = Y[gid][2 tig + 1].  SAMEW: all items share w0[0] / w1[0].
template <int NB, bool SAMEW>
__device__ __forceinline__ void cmma_batch(const double2* w0, const double2* w1, const double2* x0,
                                           const double2* x1, double2* y0, double2* y1) {
#ifndef PW_CMMA_BATCH_4M
  // Three real products per complex product (Gauss): with s = xr + xi,
  //   k = wr s,  yr = k - (wr + wi) xi,  yi = k + (wi - wr) xr
  // -- 6 DMMAs per item instead of 8 (the FP64 tensor pipe is where these kernels stall:
  // profiles/r2_ncu_rows_lines.txt), paid with one DADD per state fragment; the operator-side
  // combinations are formed once per call.  k is accumulated first and seeds both accumulators.
  double re0[NB], re1[NB], im0[NB], im1[NB], s0[NB], s1[NB];
  constexpr int NW = SAMEW ? 1 : NB;
  double n0[NW], d0[NW], n1[NW], d1[NW];
#pragma unroll
  for (int b = 0; b < NW; ++b) {
    n0[b] = -(w0[b].x + w0[b].y); d0[b] = w0[b].y - w0[b].x;
    n1[b] = -(w1[b].x + w1[b].y); d1[b] = w1[b].y - w1[b].x;
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    re0[b] = re1[b] = 0.0;
    s0[b] = x0[b].x + x0[b].y;
    s1[b] = x1[b].x + x1[b].y;
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(re0[b], re1[b], w0[SAMEW ? 0 : b].x, s0[b]);
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(re0[b], re1[b], w1[SAMEW ? 0 : b].x, s1[b]);
#pragma unroll
  for (int b = 0; b < NB; ++b) { im0[b] = re0[b]; im1[b] = re1[b]; }
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(re0[b], re1[b], n0[SAMEW ? 0 : b], x0[b].y);
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(im0[b], im1[b], d0[SAMEW ? 0 : b], x0[b].x);
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(re0[b], re1[b], n1[SAMEW ? 0 : b], x1[b].y);
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(im0[b], im1[b], d1[SAMEW ? 0 : b], x1[b].x);
#else
  double re0[NB], re1[NB], im0[NB], im1[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) re0[b] = re1[b] = im0[b] = im1[b] = 0.0;
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(re0[b], re1[b], w0[SAMEW ? 0 : b].x, x0[b].x);
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(im0[b], im1[b], w0[SAMEW ? 0 : b].y, x0[b].x);
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(re0[b], re1[b], -w0[SAMEW ? 0 : b].y, x0[b].y);
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(im0[b], im1[b], w0[SAMEW ? 0 : b].x, x0[b].y);
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(re0[b], re1[b], w1[SAMEW ? 0 : b].x, x1[b].x);
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(im0[b], im1[b], w1[SAMEW ? 0 : b].y, x1[b].x);
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(re0[b], re1[b], -w1[SAMEW ? 0 : b].y, x1[b].y);
#pragma unroll
  for (int b = 0; b < NB; ++b) dmma8x8x4(im0[b], im1[b], w1[SAMEW ? 0 : b].x, x1[b].y);
#endif
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    y0[b] = make_double2(re0[b], im0[b]);
    y1[b] = make_double2(re1[b], im1[b]);
  }
}

// A unit of work is (super-block pair (K, L)) x (32 fibers): slot (a, b) = a * n_L + b holds
// operator index pair (row index a of K, column index b of L).  A single dense d = 5..8
// operator is one super-block; a cutoff-8 beam splitter is eight, every one a full 8 x 8 operand.
// The operand fragments of a unit come from the zero-padded tables of k_ts_build (L1/L2
// resident).  Three tile buffers: the loads run two tiles ahead and a tile costs two barriers.
//
// `inner`: the column target is the innermost (contiguous) axis.  The 32 fibers of a chunk are
// then n apart and, for a fixed row index a, (fiber f, column index b) is ONE contiguous run of
// 32 n elements.  The tile keeps that order -- element (a, b, f) at a * 256 + f * 8 + (b ^ sigma(f, a))
// -- so that a cp.async instruction's destinations stay inside whole 128-byte shared-memory rows
// (scattering the 16-byte pieces over slots doubles the L2 -> SM sectors: measured); sigma
// permutes inside each row such that the stage 1 reads and writes and the stage 2 reads are all
// bank-conflict free.  The stores of stage 2 are coalesced as they are (8 column outputs x 4
// fiber pairs = four full 128-byte lines per instruction).
__device__ __forceinline__ int ts_sigma(const int f, const int a) {
  return ((f & 1) << 2) ^ (f & 6) ^ (a & 3);
}

// Storage type S of the state (double2, or float2 for complex64): the arithmetic is always the FP64
// tensor-core product; complex64 values are widened when a fragment is read from the tile and
// rounded when a result is written back (the tile and the intermediate U rho live in float2, like
// every intermediate of the reference's complex64 path), so the kernel moves half the bytes with
// the same DMMA count -- complex64 leaves the FMA pipe (33-42 % of the HBM roof) for this path.
// compute type of a storage type: complex64 tiles are multiplied in TF32 x 3 (TF) or widened to FP64
template <typename S, bool TF> struct TsCompute { using V = double2; };
template <> struct TsCompute<float2, true> { using V = float2; };
template <typename V> __device__ __forceinline__ V ts_ldv(const double2* p);
template <> __device__ __forceinline__ double2 ts_ldv<double2>(const double2* p) { return *p; }
template <typename V> __device__ __forceinline__ V ts_ldv(const float2* p);
template <> __device__ __forceinline__ double2 ts_ldv<double2>(const float2* p) { const float2 v = *p; return make_double2((double)v.x, (double)v.y); }
template <> __device__ __forceinline__ float2 ts_ldv<float2>(const float2* p) { return *p; }
template <typename V> __device__ __forceinline__ V ts_ldgv(const double2* p);
template <> __device__ __forceinline__ double2 ts_ldgv<double2>(const double2* p) { return __ldg(p); }
template <typename V> __device__ __forceinline__ V ts_ldgv(const float2* p);
template <> __device__ __forceinline__ double2 ts_ldgv<double2>(const float2* p) { const float2 v = __ldg(p); return make_double2((double)v.x, (double)v.y); }
template <> __device__ __forceinline__ float2 ts_ldgv<float2>(const float2* p) { return __ldg(p); }
__device__ __forceinline__ void ts_st(float2* p, const float2 v) { *p = v; }

__device__ __forceinline__ double2 ts_ld(const double2* p) { return *p; }
__device__ __forceinline__ double2 ts_ld(const float2* p) { const float2 v = *p; return make_double2((double)v.x, (double)v.y); }
__device__ __forceinline__ double2 ts_ldg(const double2* p) { return __ldg(p); }
__device__ __forceinline__ double2 ts_ldg(const float2* p) { const float2 v = __ldg(p); return make_double2((double)v.x, (double)v.y); }
__device__ __forceinline__ void ts_st(double2* p, const double2 v) { *p = v; }
__device__ __forceinline__ void ts_st(float2* p, const double2 v) { *p = make_float2((float)v.x, (float)v.y); }
__device__ __forceinline__ void ts_cp(double2* smem_dst, const double2* gsrc) { cp_async16(smem_dst, gsrc); }
__device__ __forceinline__ void ts_cp(float2* smem_dst, const float2* gsrc) { cp_async8(smem_dst, gsrc); }
__device__ __forceinline__ void st_pair(float2* p0, float2* p1, const float2 f0, const float2 f1,
                                        const bool ok0, const bool ok1) {
  if (ok0 && ok1 && p1 == p0 + 1 && (reinterpret_cast<uintptr_t>(p0) & 15) == 0) {
    __stcg(reinterpret_cast<float4*>(p0), make_float4(f0.x, f0.y, f1.x, f1.y));
  } else {
    if (ok0) __stcg(p0, f0);
    if (ok1) __stcg(p1, f1);
  }
}
__device__ __forceinline__ void st_pair(float2* p0, float2* p1, const double2 v0, const double2 v1,
                                        const bool ok0, const bool ok1) {
  st_pair(p0, p1, make_float2((float)v0.x, (float)v0.y), make_float2((float)v1.x, (float)v1.y), ok0, ok1);
}

template <bool MULTI, bool INNER, typename S, bool TF>
__device__ __forceinline__ void two_sided_mma_body(const S* __restrict__ in, S* __restrict__ out,
                                                   const FiberMap& fm, const TsTables& T,
                                                   unsigned char* s_raw, const int order) {
  using V = typename TsCompute<S, TF>::V;   // float2: TF32 x 3 products; double2: FP64 tensor cores
  constexpr bool kTf = sizeof(V) == 8;
  constexpr bool inner = INNER;
  S* tile0 = (S*)s_raw;                                          // [3][kTsSlots][32], swizzled
  int64_t* s_base = (int64_t*)(tile0 + 3 * kTsSlots * 32);       // [3][32] fiber bases (-1: past the end)
  int64_t* s_off = s_base + 3 * 32;                              // [4][kTsMaxSb][8]
  int* s_n = (int*)(s_off + 4 * kTsMaxSb * 8);                   // [kTsMaxSb]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t nsb = MULTI ? (uint32_t)T.flags->nsb : 1u;
  for (int i = threadIdx.x; i < 4 * kTsMaxSb * 8; i += blockDim.x) s_off[i] = T.sb_off[i];
  if (threadIdx.x < kTsMaxSb) s_n[threadIdx.x] = threadIdx.x < (int)nsb ? T.sb_n[threadIdx.x] : 0;
  __syncthreads();
  const int64_t* rin = s_off;
  const int64_t* cin = s_off + kTsMaxSb * 8;
  const int64_t* rout = s_off + 2 * kTsMaxSb * 8;
  const int64_t* cout = s_off + 3 * kTsMaxSb * 8;
  const int gid = lane >> 2, tig = lane & 3;
  const uint64_t chunks = (fm.nfib + 31) / 32;
  const uint64_t units = chunks * nsb * nsb;
  const S* wtab = (const S*)T.sb_w;

  // chunk index fastest: CTAs that run side by side work on adjacent memory
  // (32-bit arithmetic: the host guarantees units < 2^32)
  const uint32_t chunks32 = (uint32_t)chunks, nsb2 = nsb * nsb;
  auto decode = [&](uint64_t u, uint32_t& K, uint32_t& L, uint64_t& c) {
    if (!MULTI) { K = 0; L = 0; c = u; return; }
    const uint32_t u32 = (uint32_t)u;
    uint32_t p, c32;
    if (order == 0) {
      p = u32 / chunks32;
      c32 = u32 - p * chunks32;
    } else {  // pair index fastest: consecutive units sweep one chunk's whole operator block
      c32 = u32 / nsb2;
      p = u32 - c32 * nsb2;
    }
    c = c32;
    K = p / nsb;
    L = p - K * nsb;
  };

  auto issue = [&](uint64_t u, int buf) {
    if (u < units) {
      uint32_t K, L;
      uint64_t c;
      decode(u, K, L, c);
      const int nK = s_n[K], nL = s_n[L];
      S* tile = tile0 + (size_t)buf * kTsSlots * 32;
      if (!inner) {
        const uint64_t f = c * 32 + lane;
        const int64_t base = f < fm.nfib ? fm.base(f) : -1;
        if (warp == 0) s_base[buf * 32 + lane] = base;
        if (base >= 0) {
          const int64_t* ro = rin + K * 8;
          const int64_t* co = cin + L * 8;
          if (nL == kTsWarps) {
            // full super-block: warp w copies column w of every row
            const S* src = in + base + co[warp];
            S* dst = tile + warp * 32;
            for (int a = 0; a < nK; ++a)
              ts_cp(dst + a * (kTsWarps * 32) + (lane ^ (2 * (a & 3))), src + ro[a]);
          } else {
            int a = 0, b = warp;
            while (b >= nL) { b -= nL; ++a; }
            for (; a < nK;) {
              ts_cp(tile + (a * nL + b) * 32 + (lane ^ (2 * (a & 3))), in + base + ro[a] + co[b]);
              b += kTsWarps;
              while (b >= nL) { b -= nL; ++a; }
            }
          }
        }
      } else {
        // one super-block (K = L = 0, n = nK = nL); fibers of the chunk: base0 + i * n
        const int64_t base0 = fm.base(c * 32);
        const uint64_t left = fm.nfib - c * 32;
        const int nvalid = left < 32 ? (int)left : 32;
        if (warp == 0) s_base[buf * 32 + lane] = lane < nvalid ? base0 + (int64_t)lane * nL : -1;
        for (int a = warp; a < nK; a += kTsWarps) {
          const S* src = in + base0 + rin[a];
          S* dst = tile + a * 256;
          int f = lane / nL, b = lane - f * nL;
          const int qa = 32 / nL, rb = 32 - qa * nL;
          for (int e = lane; e < nvalid * nL; e += 32) {
            ts_cp(dst + f * 8 + (b ^ ts_sigma(f, a)), src + e);
            f += qa; b += rb;
            if (b >= nL) { b -= nL; ++f; }
          }
        }
      }
    }
    cp_async_commit();
  };

  // one item: out[i][f] = sum_a W[i][a] x[a][f]; fragments w0 / w1: lane (gid, tig) holds
  // W[gid][tig] and W[gid][4 + tig]; nx = valid input indices
  // (item-major issue order on purpose: the loads of fiber octet nt + 1 overlap the DMMAs of
  //  octet nt; the step-major batch of cmma_batch measured 20 % slower HERE, 6 % faster in the
  //  row-tile kernel)
  auto product = [&](const V w0, const V w1, const int nx, auto ldx, auto sty) {
    if constexpr (kTf) {
      const TfW w = tf_w(w0, w1);
      V x0[4], x1[4];
      float c[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int f = 8 * nt + gid;
        x0[nt] = czero<V>(); x1[nt] = x0[nt];
        if (tig < nx) x0[nt] = ldx(tig, f);
        if (4 + tig < nx) x1[nt] = ldx(4 + tig, f);
      }
      cprod_tf32_batch<4, false>(w, x0, x1, c);   // term-major: four independent accumulators
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        sty(nt, 8 * nt + 2 * tig, make_float2(c[nt][0], c[nt][2]), make_float2(c[nt][1], c[nt][3]));
    } else {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int f = 8 * nt + gid;
        V x0 = czero<V>(), x1 = x0;
        if (tig < nx) x0 = ldx(tig, f);
        if (4 + tig < nx) x1 = ldx(4 + tig, f);
        double re0 = 0.0, re1 = 0.0, im0, im1;
#ifndef PW_CMMA_4M
        // three real products per complex product (see cmma_batch)
        dmma8x8x4(re0, re1, w0.x, x0.x + x0.y);
        dmma8x8x4(re0, re1, w1.x, x1.x + x1.y);
        im0 = re0; im1 = re1;
        dmma8x8x4(re0, re1, -(w0.x + w0.y), x0.y);
        dmma8x8x4(im0, im1, w0.y - w0.x, x0.x);
        dmma8x8x4(re0, re1, -(w1.x + w1.y), x1.y);
        dmma8x8x4(im0, im1, w1.y - w1.x, x1.x);
#else
        im0 = im1 = 0.0;
        dmma8x8x4(re0, re1, w0.x, x0.x);
        dmma8x8x4(im0, im1, w0.y, x0.x);
        dmma8x8x4(re0, re1, -w0.y, x0.y);
        dmma8x8x4(im0, im1, w0.x, x0.y);
        dmma8x8x4(re0, re1, w1.x, x1.x);
        dmma8x8x4(im0, im1, w1.y, x1.x);
        dmma8x8x4(re0, re1, -w1.y, x1.y);
        dmma8x8x4(im0, im1, w1.x, x1.y);
#endif
        V v0, v1;
        v0.x = re0; v0.y = im0; v1.x = re1; v1.y = im1;
        sty(nt, 8 * nt + 2 * tig, v0, v1);
      }
    }
  };

  uint64_t u = blockIdx.x;
  int cur = 0;
  issue(u, 0);
  issue(u + gridDim.x, 1);
  // operand fragments (zero padded to 8 x 8): per unit, or once for a single super-block
  V l0 = ts_ldgv<V>(wtab + gid * 8 + tig), l1 = ts_ldgv<V>(wtab + gid * 8 + 4 + tig);
  V r0 = ts_ldgv<V>(wtab + kTsMaxSb * 64 + gid * 8 + tig), r1 = ts_ldgv<V>(wtab + kTsMaxSb * 64 + gid * 8 + 4 + tig);
  for (; u < units; u += gridDim.x) {
    S* tile = tile0 + (size_t)cur * kTsSlots * 32;
    uint32_t K, L;
    uint64_t c;
    decode(u, K, L, c);
    const int nK = s_n[K], nL = s_n[L];
    if (MULTI) {
      const S* wl = wtab + (size_t)K * 64 + gid * 8 + tig;
      const S* wr = wtab + (size_t)(kTsMaxSb + L) * 64 + gid * 8 + tig;
      l0 = ts_ldgv<V>(wl); l1 = ts_ldgv<V>(wl + 4); r0 = ts_ldgv<V>(wr); r1 = ts_ldgv<V>(wr + 4);
    }
    cp_async_wait<1>();
    __syncthreads();
    issue(u + 2 * (uint64_t)gridDim.x, cur == 0 ? 2 : cur - 1);  // the buffer of the previous tile
    // stage 1: column j of the tile, in place
    for (int j = warp; j < nL; j += kTsWarps) {
      S* col = tile + j * 32;
      V y0[4], y1[4];
      if (inner)
        product(l0, l1, nK, [&](int a, int f) { return ts_ldv<V>(tile + a * 256 + f * 8 + (j ^ ts_sigma(f, a))); },
                [&](int nt, int f, V v0, V v1) { y0[nt] = v0; y1[nt] = v1; (void)f; });
      else
        product(l0, l1, nK, [&](int a, int f) { return ts_ldv<V>(col + a * nL * 32 + (f ^ (2 * (a & 3)))); },
                [&](int nt, int f, V v0, V v1) { y0[nt] = v0; y1[nt] = v1; (void)f; });
      __syncwarp();  // every lane has read its inputs before the column is overwritten
      if (inner) {
        if (gid < nK) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int f = 8 * nt + 2 * tig;
            ts_st(tile + gid * 256 + f * 8 + (j ^ ts_sigma(f, gid)), y0[nt]);
            ts_st(tile + gid * 256 + (f + 1) * 8 + (j ^ ts_sigma(f + 1, gid)), y1[nt]);
          }
        }
      } else if (gid < nK) {
        S* dst = col + gid * nL * 32;
        const int sw = (2 * (j & 3)) ^ (gid & 1);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int f = 8 * nt + 2 * tig;
          ts_st(dst + (f ^ sw), y0[nt]);
          ts_st(dst + ((f + 1) ^ sw), y1[nt]);
        }
      }
    }
    __syncthreads();
    // stage 2: row i of the tile, streamed to HBM
    const int64_t* sb = s_base + cur * 32;
    if (inner) {
      // innermost column target: the TRANSPOSED product (operands swapped: A = X^T from the
      // tile, B = W^T -- the very same register values) leaves a lane with two ADJACENT output
      // positions of ONE fiber: a 256-bit store per lane, a full 128-byte line per quarter warp
      for (int i = warp; i < nK; i += kTsWarps) {
        const int64_t ro = rout[K * 8 + i] + 2 * tig;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int f = 8 * nt + gid;
          V x0 = czero<V>(), x1 = x0, v0, v1;
          if (tig < nL) x0 = ts_ldv<V>(tile + i * 256 + f * 8 + (tig ^ ts_sigma(f, i)));
          if (4 + tig < nL) x1 = ts_ldv<V>(tile + i * 256 + f * 8 + ((4 + tig) ^ ts_sigma(f, i)));
          if constexpr (kTf) {
            cprod_tf32_t(tf_w(r0, r1), x0, x1, v0, v1);
          } else {
            double re0 = 0.0, re1 = 0.0, im0, im1;
#ifndef PW_CMMA_4M
            // three real products (operands swapped: the state fragment is A, the operator B)
            dmma8x8x4(re0, re1, x0.x + x0.y, r0.x);
            dmma8x8x4(re0, re1, x1.x + x1.y, r1.x);
            im0 = re0; im1 = re1;
            dmma8x8x4(re0, re1, x0.y, -(r0.x + r0.y));
            dmma8x8x4(im0, im1, x0.x, r0.y - r0.x);
            dmma8x8x4(re0, re1, x1.y, -(r1.x + r1.y));
            dmma8x8x4(im0, im1, x1.x, r1.y - r1.x);
#else
            im0 = im1 = 0.0;
            dmma8x8x4(re0, re1, x0.x, r0.x);
            dmma8x8x4(im0, im1, x0.x, r0.y);
            dmma8x8x4(re0, re1, -x0.y, r0.y);
            dmma8x8x4(im0, im1, x0.y, r0.x);
            dmma8x8x4(re0, re1, x1.x, r1.x);
            dmma8x8x4(im0, im1, x1.x, r1.y);
            dmma8x8x4(re0, re1, -x1.y, r1.y);
            dmma8x8x4(im0, im1, x1.y, r1.x);
#endif
            v0.x = re0; v0.y = im0; v1.x = re1; v1.y = im1;
          }
          const int64_t b0 = sb[f];
          S* dst = out + b0 + ro;
          st_pair(dst, dst + 1, v0, v1, b0 >= 0 && 2 * tig < nL, b0 >= 0 && 2 * tig + 1 < nL);
        }
      }
    } else
    for (int i = warp; i < nK; i += kTsWarps) {
      const S* row = tile + i * nL * 32;
      const int64_t ro = rout[K * 8 + i] + cout[L * 8 + gid];
      product(r0, r1, nL,
              [&](int b, int f) {
                return inner ? ts_ldv<V>(tile + i * 256 + f * 8 + (b ^ ts_sigma(f, i)))
                             : ts_ldv<V>(row + b * 32 + (f ^ (2 * (b & 3)) ^ (i & 1)));
              },
              [&](int nt, int f, V v0, V v1) {
                (void)nt;
                if (gid < nL) {
                  const int64_t b0 = sb[f], b1 = sb[f + 1];
                  st_pair(out + b0 + ro, out + b1 + ro, v0, v1, b0 >= 0, b1 >= 0);
                }
              });
    }
    cur = cur == 2 ? 0 : cur + 1;
  }
  cp_async_wait<0>();
}

template <typename S, bool TF>
__global__ void __launch_bounds__(kTsWarps * 32, 2)
k_two_sided_mma(const S* __restrict__ in, S* __restrict__ out, const FiberMap fm,
                const TsTables T, const int inner, const int order) {
  if (!T.flags->handled || T.flags->use_mma != 1) return;
  extern __shared__ __align__(16) unsigned char s_raw[];
  if (inner) two_sided_mma_body<false, true, S, TF>(in, out, fm, T, s_raw, 0);
  else if (T.flags->nsb == 1) two_sided_mma_body<false, false, S, TF>(in, out, fm, T, s_raw, 0);
  else two_sided_mma_body<true, false, S, TF>(in, out, fm, T, s_raw, order);
}


struct TwMeta { int64_t base0; int K; int nvalid; };

// ---- row tiles: the column targets are the innermost (contiguous) run ---------------------------
// With the column targets innermost (beam splitter on the last two modes), the column index
// pairs of a super-block are scattered 16-byte pieces inside every run of t elements; the only
// coalesced unit is the whole run, so a tile holds ALL t positions of its fibers.
// The free (N) dimension of an MMA is any set of eight independent columns.  Stage 1 (row side)
// mixes the rows of the tile, so eight CONSECUTIVE COLUMN POSITIONS of one fiber are as good as
// eight fibers; stage 2 (column side) mixes positions inside a row, so the EIGHT ROWS of one fiber
// are its free dimension.  A tile therefore needs neither 8 nor 32 fibers: (8 rows) x (F fibers
// x all t positions) = 2048 elements = 32 KB, three buffers and two decoupled 8-warp CTAs per SM
// (the shape of k_two_sided_mma), instead of one lock-stepped 16-warp CTA with 64 KB tiles.
//   element (row a, fiber f, position j)  at  a * 256 + f * tp + (j ^ sigma(a)),
//   sigma(a) = 2 (a & 3) ^ (a & 4)
// column side (first, in place): B fragment (k = position cin[tig] / cin[4 + tig], n = row pi(gid)),
//   C fragment (position cout[gid], rows pi(2 tig) = tig and pi(2 tig + 1) = tig + 4),
//   pi(g) = (g >> 1) + 4 (g & 1);
// row side (second, streamed to HBM from the fragments -- the sides commute): B fragment (k = row
//   tig / 4 + tig, n = position 8 q + gid), C fragment (row gid, positions 8 q + 2 tig, + 1: one
//   full sector per lane, one full 128-byte line per quarter warp);
// all three shared-memory access patterns are bank-conflict free for positions that are
// consecutive modulo 8 (beam-splitter sectors: 7 k + n), the copies stay inside whole 128-byte rows.
constexpr int kTnWarps = 8;
constexpr int kTnRow = 256;   // elements per tile row (F fibers x padded t)
constexpr int kTnBufs = 3;

// Column targets = the innermost run (lseg positions) x outer column targets (nseg segments):
// position p = s * lseg + jr of fiber f lives at seg_off[s] + f * lseg + jr.  A beam splitter on
// the last two modes is one segment (lseg = t); on modes (4, 0) it is eight segments of eight.
struct TnSeg { int lseg, nseg; int64_t seg_off[64]; };

// float2 items: TF32 x 3 products (the operand fragments are split once per batch)
template <int NB, bool SAMEW>
__device__ __forceinline__ void cmma_batch(const float2* w0, const float2* w1, const float2* x0,
                                           const float2* x1, float2* y0, float2* y1) {
  static_assert(SAMEW, "float2 batches share their operand");
  const TfW w = tf_w(w0[0], w1[0]);
  float c[NB][4];
  cprod_tf32_batch<NB, false>(w, x0, x1, c);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    y0[b] = make_float2(c[b][0], c[b][2]);
    y1[b] = make_float2(c[b][1], c[b][3]);
  }
}

template <typename S, bool TF>
__global__ void __launch_bounds__(kTnWarps * 32, 2)
k_two_sided_mma_rows(const S* __restrict__ in, S* __restrict__ out, const FiberMap fm,
                     const TsTables T, const int t, const int order, const TnSeg seg) {
  using V = typename TsCompute<S, TF>::V;
  if (!T.flags->handled || T.flags->use_mma != 2) return;
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ int64_t s_seg[64];
  if (threadIdx.x < 64) s_seg[threadIdx.x] = threadIdx.x < seg.nseg ? seg.seg_off[threadIdx.x] : 0;
  S* tile0 = (S*)s_raw;                                       // [kTnBufs][8][kTnRow]
  V* s_w = (V*)(tile0 + kTnBufs * 8 * kTnRow);                // [kTwMaxSb][64] column-side operands (swizzled)
  int64_t* s_roff = (int64_t*)(s_w + kTwMaxSb * 64);          // [2][kTwMaxSb][8] row offsets: in, out
  TwMeta* s_meta = (TwMeta*)(s_roff + 2 * kTwMaxSb * 8);      // [kTnBufs]
  int* s_cpos = (int*)(s_meta + kTnBufs);                     // [2][kTwMaxSb][8] column positions: in, out
  int* s_n = s_cpos + 2 * kTwMaxSb * 8;                       // [kTwMaxSb]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int nsb = T.flags->nsb;
  for (int i = threadIdx.x; i < kTwMaxSb * 64; i += blockDim.x)
    s_w[i ^ (((i >> 3) & 1) << 2)] = ts_ldv<V>((const S*)T.sb_w + kTsMaxSb * 64 + i);
  for (int i = threadIdx.x; i < 2 * kTwMaxSb * 8; i += blockDim.x) {
    const int which = i / (kTwMaxSb * 8), r = i - which * (kTwMaxSb * 8);
    s_roff[i] = T.sb_off[(2 * which) * kTsMaxSb * 8 + r];
    s_cpos[i] = (int)T.sb_off[(2 * which + 1) * kTsMaxSb * 8 + r];
  }
  if (threadIdx.x < kTwMaxSb) s_n[threadIdx.x] = threadIdx.x < nsb ? T.sb_n[threadIdx.x] : 0;
  __syncthreads();
  const int64_t* rin = s_roff;
  const int64_t* rout = s_roff + kTwMaxSb * 8;
  const int* cin = s_cpos;
  const int* cout = s_cpos + kTwMaxSb * 8;
  const int tp = (t + 7) & ~7;
  const int F = kTnRow / tp;                   // fibers per tile (t <= 64: at least 4)
  const int nq = tp >> 3;                      // position octets per fiber
  const int lseg = seg.lseg, nseg = seg.nseg;
  const uint32_t inv_l = 0xffffffffu / (uint32_t)lseg + 1;   // x / lseg for x < 2^16
  const uint32_t inv_nq = 0xffffu / (uint32_t)nq + 1;
  const uint32_t gl = (uint32_t)fm.gsize[fm.ng - 1];
  const uint32_t cpr = (gl + F - 1) / F;
  const uint32_t chunks = (uint32_t)(fm.nfib / gl) * cpr;
  const uint32_t units = chunks * (uint32_t)nsb;
  const S* wl_tab = (const S*)T.sb_w;          // row-side operands (global, L1 resident)

  auto sig = [](int a) { return (2 * (a & 3)) ^ (a & 4); };

  auto issue = [&](uint32_t u, int buf) {
    if (u < units) {
      uint32_t K, c;
      if ((order & 1) == 0) { K = u / chunks; c = u - K * chunks; }       // chunk fastest
      else { c = u / (uint32_t)nsb; K = u - c * (uint32_t)nsb; }    // row super-block fastest
      const uint32_t o = c / cpr, w = c - o * cpr;
      const uint32_t f0 = w * F;
      const int nvalid = (gl - f0) < (uint32_t)F ? (int)(gl - f0) : F;
      const int64_t base0 = fm.base((uint64_t)o * gl + f0);
      if (threadIdx.x == 0) { s_meta[buf].base0 = base0; s_meta[buf].K = (int)K; s_meta[buf].nvalid = nvalid; }
      const int nK = s_n[K];
      S* tile = tile0 + (size_t)buf * 8 * kTnRow;
      const int64_t* ro = rin + K * 8;
      // per segment the F fibers x lseg positions are one contiguous piece
      const int rowlen = nvalid * lseg;
      for (int idx = threadIdx.x; idx < rowlen * nseg; idx += kTnWarps * 32) {
        const int sg = nseg == 1 ? 0 : idx / rowlen;
        const int e = idx - sg * rowlen;
        const int f = (int)__umulhi((uint32_t)e, inv_l), j = sg * lseg + (e - f * lseg);
        S* dst = tile + f * tp;
        const S* src = in + base0 + s_seg[sg] + e;
        for (int a = 0; a < nK; ++a) ts_cp(dst + a * kTnRow + (j ^ sig(a)), src + ro[a]);
      }
    }
    cp_async_commit();
  };


  const int fsw = (gid & 1) << 2;
  const V zero = czero<V>();
  // row-side lane constants: read rows tig / 4 + tig at position octet + gid
  const int rd1a = tig * kTnRow + (gid ^ sig(tig)), rd1b = (4 + tig) * kTnRow + (gid ^ sig(4 + tig));
  // column-side lane constants: read row pi(gid), write rows tig and tig + 4
  const int prow = (gid >> 1) + 4 * (gid & 1);
  const int rd2 = prow * kTnRow, sg2 = sig(prow);
  const int wr2a = tig * kTnRow, wr2b = (4 + tig) * kTnRow, sa = sig(tig), sb4 = sig(4 + tig);

  const uint32_t g = gridDim.x;
  uint32_t u = blockIdx.x;
  int cur = 0;
  issue(u, 0);
  issue(u + g, 1);
  for (; u < units; u += g) {
    S* tile = tile0 + (size_t)cur * 8 * kTnRow;
    cp_async_wait<1>();
    __syncthreads();  // A: tile `cur` is resident; the previous tile has been copied out
    issue(u + 2 * g, cur == 0 ? kTnBufs - 1 : cur - 1);
    const TwMeta meta = s_meta[cur];
    const int K = meta.K, nK = s_n[K];
    const V l0 = ts_ldgv<V>(wl_tab + K * 64 + gid * 8 + tig), l1 = ts_ldgv<V>(wl_tab + K * 64 + gid * 8 + 4 + tig);
    // The two sides commute: the COLUMN side runs first, in place, so that the row side can
    // stream its accumulators straight to HBM -- a lane owns positions 8 q + 2 tig and + 1 of one
    // output row (one full 32-byte sector, a quarter warp one full 128-byte line).  Four items per
    // step: eight independent DMMA chains per warp.
    constexpr int kB = 4;
    // stage 1 (column side): (fiber f, column super-block L), positions mixed inside every row
#ifdef PW_TS_ABLATE  // profiling builds only (make EXTRA=-DPW_TS_ABLATE): ts_order bit 2 = no column side
    const int n2g = (order & 2) ? 0 : nsb * ((F + kB - 1) / kB);
#else
    const int n2g = nsb * ((F + kB - 1) / kB);
#endif
    // an item is (column super-block L, group of kB fibers): operand fragments and positions are
    // loaded once per item and shared by its fibers
    const int ngrp = (F + kB - 1) / kB;
    for (int it = warp; it < n2g; it += kTnWarps) {
      const int L = it / ngrp, f0 = (it - L * ngrp) * kB;
      const int nL = s_n[L];
      const V r0 = s_w[L * 64 + gid * 8 + (tig ^ fsw)], r1 = s_w[L * 64 + gid * 8 + ((4 + tig) ^ fsw)];
      const int c0 = rd2 + (cin[L * 8 + tig] ^ sg2), c1 = rd2 + (cin[L * 8 + 4 + tig] ^ sg2);
      const int pos = cout[L * 8 + gid];
      const int wa = wr2a + (pos ^ sa), wb = wr2b + (pos ^ sb4);
      V x0[kB], x1[kB], y0[kB], y1[kB];
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        const S* fib = tile + (f0 + b) * tp;
        x0[b] = (tig < nL && f0 + b < F) ? ts_ldv<V>(fib + c0) : zero;
        x1[b] = (4 + tig < nL && f0 + b < F) ? ts_ldv<V>(fib + c1) : zero;
      }
      cmma_batch<kB, true>(&r0, &r1, x0, x1, y0, y1);
      __syncwarp();
      if (gid < nL) {
#pragma unroll
        for (int b = 0; b < kB; ++b)
          if (f0 + b < F) {
            S* fib = tile + (f0 + b) * tp;
            ts_st(fib + wa, y0[b]);
            ts_st(fib + wb, y1[b]);
          }
      }
    }
    __syncthreads();  // B
    // stage 2 (row side): (fiber f, position octet q), rows mixed, streamed to HBM
    const int n1 = F * nq;
    S* orow = out + meta.base0 + rout[K * 8 + (gid < nK ? gid : 0)];
    for (int it0 = warp; it0 < n1; it0 += kB * kTnWarps) {
      int fq[kB];
      V x0[kB], x1[kB], y0[kB], y1[kB];
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        const int it = it0 + b * kTnWarps;
        const int itc = it < n1 ? it : it0;
        const int f = (int)(((uint32_t)itc * inv_nq) >> 16), q = itc - f * nq;
        const S* col = tile + f * tp + 8 * q;
        fq[b] = it < n1 ? (f << 8) | q : -1;
        x0[b] = (tig < nK && it < n1) ? ts_ldv<V>(col + rd1a) : zero;
        x1[b] = (4 + tig < nK && it < n1) ? ts_ldv<V>(col + rd1b) : zero;
      }
#ifdef PW_TS_ABLATE  // ts_order bit 4 = no row-side arithmetic (copies the inputs)
      if (order & 4) {
#pragma unroll
        for (int b = 0; b < kB; ++b) { y0[b] = x0[b]; y1[b] = x1[b]; }
      } else
#endif
        cmma_batch<kB, true>(&l0, &l1, x0, x1, y0, y1);
      if (gid < nK) {
#pragma unroll
        for (int b = 0; b < kB; ++b) {
          if (fq[b] < 0) continue;
          const int f = fq[b] >> 8, j = 8 * (fq[b] & 255) + 2 * tig;
          if (f >= meta.nvalid) continue;
          const int sg0 = (int)__umulhi((uint32_t)j, inv_l), sg1 = (int)__umulhi((uint32_t)(j + 1), inv_l);
          S* dst = orow + (int64_t)f * lseg;
          st_pair(dst + s_seg[sg0] + (j - sg0 * lseg), dst + s_seg[sg1] + (j + 1 - sg1 * lseg), y0[b], y1[b],
                  j < t, j + 1 < t);
        }
      }
    }
    cur = cur == kTnBufs - 1 ? 0 : cur + 1;
  }
  cp_async_wait<0>();
}

template <typename V>
static int launch_two_sided_t(const V* rho, V* out, const V* op, const FiberMap& fm,
                              const TargetMap& rows, const TargetMap& cols, const int* other,
                              const int inner, const TnSeg& seg, const TargetMap& cols_pos,
                              TwoSided* ts, cudaStream_t s) {
  // (innermost column targets: the column side is addressed by POSITION inside the fiber row)
  PW_TRY(build_block_desc_pair_t<V>(op, rows.t, false, rows, &ts->memL, &ts->L, true, inner ? cols_pos : cols,
                                    &ts->memR, &ts->R, s));
  const size_t t = rows.t, tt = t * t;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_flags = 256, b_groups = al(sizeof(int4) * 2 * tt), b_pairs = al(sizeof(int4) * tt),
               b_off = al(sizeof(int64_t) * tt), b_item = al(sizeof(int4) * tt), b_row = al(sizeof(int) * tt),
               b_sbn = al(sizeof(int) * kTsMaxSb), b_sbblk = al(sizeof(int2) * t),
               b_sboff = al(sizeof(int64_t) * 4 * kTsMaxSb * 8), b_sbw = al(sizeof(V) * 2 * kTsMaxSb * 64);
  PW_TRY(ts->flag_mem.alloc(b_flags + b_groups + b_pairs + 2 * b_off + 2 * b_item + b_row + b_sbn + b_sbblk +
                            b_sboff + b_sbw, s));
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
  char* sbb = base + b_flags + b_groups + b_pairs + 2 * b_off + 2 * b_item + b_row;
  T.sb_n = (int*)sbb;
  T.sb_blk = (int2*)(sbb + b_sbn);
  T.sb_off = (int64_t*)(sbb + b_sbn + b_sbblk);
  T.sb_w = sbb + b_sbn + b_sbblk + b_sboff;
  // (the tensor-core kernels index their units with 32 bits)
  // complex64: the same FP64 tensor-core kernel with float2 storage (values widened per fragment);
  const bool c128 = sizeof(V) == 16;
  const bool mma_ok = options().ts_mma.load(std::memory_order_relaxed) && fm.nfib < (1ull << 28) &&
                      (c128 || options().ts_mma_c64.load(std::memory_order_relaxed));
  if (inner && !mma_ok) return PW_ERR_UNSUPPORTED;
  if (!plan_replay()) {
    k_ts_build<V><<<1, 256, 0, s>>>(ts->L, ts->R, other, T, kTsValCap, mma_ok ? 1 : 0, inner);
    PW_LAUNCHED();
  }
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
  if (!inner) k_two_sided<V><<<grid, kTsWarps * 32, smem, s>>>(rho, out, fm, ts->L, ts->R, T);
  if (mma_ok) {
    if (!inner) PW_LAUNCHED();
    static bool attr_mma = false;
    const size_t smem_mma = sizeof(V) * 3 * kTsSlots * 32 + (3 * 32 + 4 * kTsMaxSb * 8) * sizeof(int64_t) +
                            kTsMaxSb * sizeof(int);
    if (!attr_mma) {
      PW_CUDA(cudaFuncSetAttribute(k_two_sided_mma<double2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(sizeof(double2) * 3 * kTsSlots * 32 + (3 * 32 + 4 * kTsMaxSb * 8) * sizeof(int64_t) +
                                         kTsMaxSb * sizeof(int))));
      const int sm_f = (int)(sizeof(float2) * 3 * kTsSlots * 32 + (3 * 32 + 4 * kTsMaxSb * 8) * sizeof(int64_t) +
                             kTsMaxSb * sizeof(int));
      PW_CUDA(cudaFuncSetAttribute(k_two_sided_mma<float2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_f));
      PW_CUDA(cudaFuncSetAttribute(k_two_sided_mma<float2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_f));
      attr_mma = true;
    }
    // complex64: TF32 x 3 (default) or widened FP64 products (ts_mma_c64 = 2)
    const bool tf = !c128 && options().ts_mma_c64.load(std::memory_order_relaxed) == 1;
    const int ord = options().ts_order.load(std::memory_order_relaxed);
    if (tf) k_two_sided_mma<V, !std::is_same<V, double2>::value><<<grid, kTsWarps * 32, smem_mma, s>>>(rho, out, fm, T, inner ? 1 : 0, ord);
    else k_two_sided_mma<V, false><<<grid, kTsWarps * 32, smem_mma, s>>>(rho, out, fm, T, inner ? 1 : 0, ord);
    if (inner) {
      PW_LAUNCHED();
      static bool attr_rows = false;
      auto rows_smem = [](size_t es) {
        return es * (size_t)(kTnBufs * 8 * kTnRow) + sizeof(double2) * (size_t)(kTwMaxSb * 64) +
               sizeof(int64_t) * 2 * kTwMaxSb * 8 + sizeof(TwMeta) * kTnBufs + sizeof(int) * (2 * kTwMaxSb * 8 + kTwMaxSb);
      };
      if (!attr_rows) {
        PW_CUDA(cudaFuncSetAttribute(k_two_sided_mma_rows<double2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rows_smem(16)));
        PW_CUDA(cudaFuncSetAttribute(k_two_sided_mma_rows<float2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rows_smem(8)));
        PW_CUDA(cudaFuncSetAttribute(k_two_sided_mma_rows<float2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rows_smem(8)));
        attr_rows = true;
      }
      if (tf) k_two_sided_mma_rows<V, !std::is_same<V, double2>::value><<<2 * kNumSMs, kTnWarps * 32, rows_smem(sizeof(V)), s>>>(rho, out, fm, T, (int)cols.t, ord, seg);
      else k_two_sided_mma_rows<V, false><<<2 * kNumSMs, kTnWarps * 32, rows_smem(sizeof(V)), s>>>(rho, out, fm, T, (int)cols.t, ord, seg);
    }
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
  // innermost axis: untouched (adjacent fibers), or the single column target of a 5..8-level
  // operator with whole chunks of 32 fibers in the run next to it (tensor-core kernel only)
  // bit 1: the column targets are exactly the innermost run of t elements (row tiles);
  // bit 2: ... and ONE 5..8-level target with whole chunks of 32 fibers next to it (32-fiber tiles)
  int inner = 0;
  TnSeg seg;
  seg.lseg = (int)cols.t; seg.nseg = 1; seg.seg_off[0] = 0;
  TargetMap cols_pos = cols;
  if (fm.ng == 0) return PW_ERR_UNSUPPORTED;
  if (fm.gstride[fm.ng - 1] != 1) {
    // the innermost axis is a column target: the column targets below the last untouched run
    // must fill the innermost run exactly (lseg positions); the other column targets enumerate
    // segments (tensor order, last fastest)
    const int64_t L = fm.gstride[fm.ng - 1];
    if ((dtype != PW_C128 && !options().ts_mma_c64.load(std::memory_order_relaxed)) || cols.t < 5 || cols.t > 64 ||
        fm.nfib >= (1ull << 28))
      return PW_ERR_UNSUPPORTED;
    int64_t prod_in = 1;
    int outer[kMaxTargets], n_out = 0;
    for (int k = 0; k < cols.nt; ++k) {
      if (cols.tdim[k] == 1) { cols_pos.tstride[k] = 0; continue; }
      if (cols.tstride[k] < L) prod_in *= cols.tdim[k];
      else outer[n_out++] = k;
    }
    if (prod_in != L || L > 64 || cols.t % (uint32_t)L != 0 || cols.t / (uint32_t)L > 64) return PW_ERR_UNSUPPORTED;
    for (int i = 1; i < n_out; ++i)  // sort by stride, descending (tensor order)
      for (int j = i; j > 0 && cols.tstride[outer[j]] > cols.tstride[outer[j - 1]]; --j) {
        const int tmp = outer[j]; outer[j] = outer[j - 1]; outer[j - 1] = tmp;
      }
    seg.lseg = (int)L;
    seg.nseg = (int)(cols.t / (uint32_t)L);
    int64_t ps = L;
    for (int i = n_out - 1; i >= 0; --i) { cols_pos.tstride[outer[i]] = ps; ps *= cols.tdim[outer[i]]; }
    for (int sg = 0; sg < seg.nseg; ++sg) {
      int rem = sg;
      int64_t off = 0;
      for (int i = n_out - 1; i >= 0; --i) {
        const int dd = (int)cols.tdim[outer[i]];
        off += (int64_t)(rem % dd) * cols.tstride[outer[i]];
        rem /= dd;
      }
      seg.seg_off[sg] = off;
    }
    inner = 1;
    if (cols.nt == 1 && cols.tstride[0] == 1 && cols.t <= 8 && fm.gsize[fm.ng - 1] % 32 == 0) inner |= 2;
  }
  if (dtype == PW_C128)
    return launch_two_sided_t<double2>((const double2*)rho, (double2*)out, (const double2*)op, fm, rows,
                                       cols, other_handled, inner, seg, cols_pos, ts, stream);
  if (dtype == PW_C64)
    return launch_two_sided_t<float2>((const float2*)rho, (float2*)out, (const float2*)op, fm, rows,
                                      cols, other_handled, inner, seg, cols_pos, ts, stream);
  return PW_ERR_INVALID;
}

// ---- staged block kernel on the FP64 tensor cores ---------------------------------------------
// Same geometry as k_apply_blocks_seg (the innermost axis is a target: a fiber's t elements are
// nseg contiguous segments), but a CTA stages 64 adjacent fibers, the blocks are packed into
// super-blocks of <= 8 indices (first-fit decreasing, like the two-sided kernel) and a warp item
// is (super-block) x (32 fibers): 8 shared loads, 32 DMMAs, 8 shared stores, in place.  The tile
// is [fiber][t] (pitch a multiple of 8, position p of fiber f at p ^ sigma(f): loads, fragment
// reads and fragment writes are all bank-conflict free and a cp.async instruction stays inside
// whole 128-byte rows); copy-in and copy-out are the coalesced segment rows of the scalar
// kernel.  The scalar kernel is bound by its per-fiber instruction stream (52 % of the HBM roof
// for the loss channel on the innermost mode of rho); here the arithmetic is 0.125 DMMA per
// element and the kernel is HBM-bound.
constexpr int kSgMaxSb = 16;
// fibers per tile: a multiple of 64 that makes the tile ~32 KB (64 for t >= 32, 512 for t = 4)
static inline int seg_mma_fibers(int t) {
  const int tp = (t + 7) & ~7;
  int f = (2048 / tp) & ~63;
  return f < 64 ? 64 : f;
}

struct SegSb {
  int* flags;      // [0] = 1: the tensor-core kernel runs (and the scalar one exits); [1] = nsb
  int* sb_n;       // [kSgMaxSb]
  int2* sb_blk;    // [t] per block: (super-block, first index)
  int* pos;        // [2][kSgMaxSb][8] positions inside the fiber row: inputs, outputs
  double2* w;      // [kSgMaxSb][8][8] zero-padded block-diagonal values, column c of row r at (c ^ 4 (r & 1))
};

__global__ void __launch_bounds__(128)
k_seg_sb_build(const BlockArrays A, const SegSb S) {
  __shared__ int s_bn[kBlockMaxT];
  __shared__ int s_ok, s_square;
  const BlockDescHeader* h = A.hdr;
  if (threadIdx.x == 0) s_square = 1;
  __syncthreads();
  const int m = h->seg_handled ? min(h->nblocks, (int)kBlockMaxT) : 0;
  for (int b = threadIdx.x; b < m; b += blockDim.x) {
    const int4 x = A.blocks[b];
    s_bn[b] = x.x;
    if (x.x != x.y) s_square = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // all 1 x 1 (diagonal / permutation) operators stay on the scalar kernel
    int ok = (h->seg_handled && h->nzero == 0 && h->maxn >= 2 && h->maxn <= 8 && s_square) ? 1 : 0;
    int nsb = 0, fill[kSgMaxSb];
    for (int sz = 8; sz >= 1 && ok; --sz)
      for (int b = 0; b < m && ok; ++b) {
        if (s_bn[b] != sz) continue;
        int q = 0;
        while (q < nsb && fill[q] + sz > 8) ++q;
        if (q == nsb) {
          if (nsb == kSgMaxSb) { ok = 0; break; }
          fill[nsb++] = 0;
        }
        S.sb_blk[b] = make_int2(q, fill[q]);
        fill[q] += sz;
      }
    for (int q = 0; q < nsb && ok; ++q) S.sb_n[q] = fill[q];
    S.flags[0] = ok;
    S.flags[1] = nsb;
    s_ok = ok;
  }
  __syncthreads();
  if (!s_ok) return;
  for (int i = threadIdx.x; i < kSgMaxSb * 64; i += blockDim.x) S.w[i] = make_double2(0.0, 0.0);
  for (int i = threadIdx.x; i < 2 * kSgMaxSb * 8; i += blockDim.x) S.pos[i] = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < m; b += blockDim.x) {
    const int4 hb = A.blocks[b];
    const int2 sp = S.sb_blk[b];
    const int n = hb.x;
    const int64_t* bo = A.offs + hb.z;   // n column (input) positions, then n row (output) positions
    const double2* v = (const double2*)A.vals + hb.w;
    for (int a = 0; a < n; ++a) {
      const int r = sp.y + a;
      S.pos[sp.x * 8 + r] = (int)bo[a];
      S.pos[(kSgMaxSb + sp.x) * 8 + r] = (int)bo[n + a];
      for (int c = 0; c < n; ++c) S.w[sp.x * 64 + r * 8 + ((sp.y + c) ^ ((r & 1) << 2))] = v[a * n + c];
    }
  }
}

struct SgMeta { int64_t base; int nvalid; int pad; };

__global__ void __launch_bounds__(512, 1)
k_apply_blocks_seg_mma(const double2* __restrict__ in, double2* __restrict__ out, const SegGeom g,
                       const SegSb S, const int kSgFib) {
  using V = double2;
  if (!S.flags[0]) return;
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int t = (int)g.t, tp = (t + 7) & ~7;
  V* s_w = (V*)s_raw;                                   // [kSgMaxSb][64]
  V* tile0 = s_w + kSgMaxSb * 64;                       // [2][kSgFib][tp]
  SgMeta* s_meta = (SgMeta*)(tile0 + 2 * kSgFib * tp);  // [2]
  int* s_pos = (int*)(s_meta + 2);                      // [2][kSgMaxSb][8]
  int* s_n = s_pos + 2 * kSgMaxSb * 8;                  // [kSgMaxSb]
  __shared__ int64_t s_seg[64];                         // segment offsets
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int nsb = S.flags[1];
  if (threadIdx.x < 64) s_seg[threadIdx.x] = threadIdx.x < g.nseg ? g.seg_off[threadIdx.x] : 0;
  for (int i = threadIdx.x; i < kSgMaxSb * 64; i += blockDim.x) s_w[i] = S.w[i];
  for (int i = threadIdx.x; i < 2 * kSgMaxSb * 8; i += blockDim.x) s_pos[i] = S.pos[i];
  if (threadIdx.x < kSgMaxSb) s_n[threadIdx.x] = threadIdx.x < nsb ? S.sb_n[threadIdx.x] : 0;
  __syncthreads();
  const uint32_t lseg = g.lseg;
  const uint32_t inv_l = 0xffffffffu / lseg + 1;        // e / lseg for e < 2^16
  const uint64_t chunks = (g.M + kSgFib - 1) / kSgFib;
  const uint64_t units = g.outer.nfib * chunks;

  auto sigma = [](int f) { return ((f & 1) << 2) ^ (f & 6); };
  // copy the nseg coalesced segment rows of a tile: global <-> [fiber][position ^ sigma(fiber)]
  auto decode = [&](uint64_t u, int64_t& base, int& nvalid) {
    const uint64_t o = u / chunks, c = u - o * chunks;
    const uint64_t m0 = c * kSgFib;
    nvalid = (g.M - m0) < (uint64_t)kSgFib ? (int)(g.M - m0) : kSgFib;
    base = g.outer.base(o) + (int64_t)m0 * lseg;
  };
  auto issue = [&](uint64_t u, int buf) {
    if (u < units) {
      int64_t base;
      int nvalid;
      decode(u, base, nvalid);
      if (threadIdx.x == 0) { s_meta[buf].base = base; s_meta[buf].nvalid = nvalid; }
      V* tile = tile0 + (size_t)buf * kSgFib * tp;
      const int rowlen = nvalid * (int)lseg;
      if (g.nseg >= 4) {
        // many short segments: the (fiber, position) arithmetic once per element column
        for (int e = threadIdx.x; e < rowlen; e += blockDim.x) {
          const int f = (int)__umulhi((uint32_t)e, inv_l);
          const int jr = e - f * (int)lseg, sf_ = sigma(f);
          V* dst = tile + f * tp;
          const V* src = in + base + e;
          for (uint32_t sg = 0; sg < g.nseg; ++sg)
            cp_async16(dst + (((int)(sg * lseg) + jr) ^ sf_), src + s_seg[sg]);
        }
      } else {
        for (uint32_t sg = 0; sg < g.nseg; ++sg) {
          const V* src = in + base + s_seg[sg];
          for (int e = threadIdx.x; e < rowlen; e += blockDim.x) {
            const int f = (int)__umulhi((uint32_t)e, inv_l);
            const int p = (int)(sg * lseg) + (e - f * (int)lseg);
            cp_async16(tile + f * tp + (p ^ sigma(f)), src + e);
          }
        }
      }
    }
    cp_async_commit();
  };

  const int sf = sigma(gid);             // fragment reads: fiber 8 n + gid
  const int fsw = (gid & 1) << 2;
  const V zero = make_double2(0.0, 0.0);
  uint64_t u = blockIdx.x;
  int cur = 0;
  issue(u, 0);
  for (; u < units; u += gridDim.x) {
    V* tile = tile0 + (size_t)cur * kSgFib * tp;
    __syncthreads();  // the other buffer has been copied out
    issue(u + gridDim.x, cur ^ 1);
    cp_async_wait<1>();
    __syncthreads();
    const SgMeta meta = s_meta[cur];
    const int halves = kSgFib >> 5;  // groups of 32 fibers
    for (int it = warp; it < halves * nsb; it += nwarps) {
      const int sb = it / halves, h = it - sb * halves;
      const int n = s_n[sb];
      const V w0 = s_w[sb * 64 + gid * 8 + (tig ^ fsw)], w1 = s_w[sb * 64 + gid * 8 + ((4 + tig) ^ fsw)];
      const int p0 = s_pos[sb * 8 + tig] ^ sf, p1 = s_pos[sb * 8 + 4 + tig] ^ sf;
      V* xf = tile + (32 * h + gid) * tp;
      V x0[4], x1[4], y0[4], y1[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        x0[q] = tig < n ? xf[8 * q * tp + p0] : zero;
        x1[q] = 4 + tig < n ? xf[8 * q * tp + p1] : zero;
      }
      cmma_batch<4, true>(&w0, &w1, x0, x1, y0, y1);
      __syncwarp();  // every lane has read its inputs before the positions are overwritten
      if (gid < n) {
        const int po = s_pos[(kSgMaxSb + sb) * 8 + gid] ^ (2 * tig);   // sigma(8 q + 2 tig) = 2 tig
        V* yf = tile + (32 * h + 2 * tig) * tp;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          yf[8 * q * tp + po] = y0[q];
          yf[(8 * q + 1) * tp + (po ^ 4)] = y1[q];
        }
      }
    }
    __syncthreads();
    {
      const int rowlen = meta.nvalid * (int)lseg;
      if ((lseg & 1) == 0 && g.nseg < 4) {
        // few long even segments: two adjacent positions per thread, one 256-bit store
        // (many short segments: measured slower -- fewer active threads per row)
        for (int e = 2 * threadIdx.x; e < rowlen; e += 2 * blockDim.x) {
          const int f = (int)__umulhi((uint32_t)e, inv_l);
          const int jr = e - f * (int)lseg, sf_ = sigma(f);
          const V* src = tile + f * tp;
          V* dst = out + meta.base + e;
          for (uint32_t sg = 0; sg < g.nseg; ++sg) {
            const int p = (int)(sg * lseg) + jr;
            V* d0 = dst + s_seg[sg];
            st_pair(d0, d0 + 1, src[p ^ sf_], src[(p + 1) ^ sf_], true, true);
          }
        }
      } else if (g.nseg >= 4) {
        for (int e = threadIdx.x; e < rowlen; e += blockDim.x) {
          const int f = (int)__umulhi((uint32_t)e, inv_l);
          const int jr = e - f * (int)lseg, sf_ = sigma(f);
          const V* src = tile + f * tp;
          V* dst = out + meta.base + e;
          for (uint32_t sg = 0; sg < g.nseg; ++sg)
            __stcg(dst + s_seg[sg], src[((int)(sg * lseg) + jr) ^ sf_]);
        }
      } else {
        for (uint32_t sg = 0; sg < g.nseg; ++sg) {
          V* dst = out + meta.base + s_seg[sg];
          for (int e = threadIdx.x; e < rowlen; e += blockDim.x) {
            const int f = (int)__umulhi((uint32_t)e, inv_l);
            const int p = (int)(sg * lseg) + (e - f * (int)lseg);
            __stcg(dst + e, tile[f * tp + (p ^ sigma(f))]);
          }
        }
      }
    }
    cur ^= 1;
  }
  cp_async_wait<0>();
}

static int enqueue_seg_mma(const void* in, void* out, const SegGeom& g, BlockDesc* d, int dtype,
                           cudaStream_t s, const int** mma_taken) {
  *mma_taken = nullptr;
  if (dtype != PW_C128 || g.t > 64 || g.t < 4 || g.lseg > 1024 ||
      !options().seg_mma.load(std::memory_order_relaxed))
    return PW_OK;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_flags = 256, b_n = al(sizeof(int) * kSgMaxSb), b_blk = al(sizeof(int2) * g.t),
               b_pos = al(sizeof(int) * 2 * kSgMaxSb * 8), b_w = al(sizeof(double2) * kSgMaxSb * 64);
  PW_TRY(d->mem2.alloc(b_flags + b_n + b_blk + b_pos + b_w, s));
  char* base = (char*)d->mem2.p;
  SegSb S;
  S.flags = (int*)base;
  S.sb_n = (int*)(base + b_flags);
  S.sb_blk = (int2*)(base + b_flags + b_n);
  S.pos = (int*)(base + b_flags + b_n + b_blk);
  S.w = (double2*)(base + b_flags + b_n + b_blk + b_pos);
  if (!plan_replay()) {
    k_seg_sb_build<<<1, 128, 0, s>>>(d->arr, S);
    PW_LAUNCHED();
  }
  const int tp = ((int)g.t + 7) & ~7;
  const int kSgFib = seg_mma_fibers((int)g.t);
  const size_t smem = sizeof(double2) * (size_t)(kSgMaxSb * 64 + 2 * kSgFib * tp) + sizeof(SgMeta) * 2 +
                      sizeof(int) * (2 * kSgMaxSb * 8 + kSgMaxSb);
  static bool attr_set = false;
  if (!attr_set) {
    PW_CUDA(cudaFuncSetAttribute(k_apply_blocks_seg_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  int nwarps = (kSgFib / 32) * (((int)g.t + 7) / 8);  // one warp per (32 fibers, 8 indices)
  if (nwarps < 4) nwarps = 4;
  if (nwarps > 16) nwarps = 16;
  const int ctas = (smem + 1024) * 2 <= 227 * 1024 ? 2 : 1;
  const uint64_t units = g.outer.nfib * ((g.M + kSgFib - 1) / kSgFib);
  uint64_t grid = (uint64_t)kNumSMs * ctas;
  if (grid > units) grid = units;
  k_apply_blocks_seg_mma<<<(unsigned)grid, nwarps * 32, smem, s>>>((const double2*)in, (double2*)out, g, S, kSgFib);
  PW_LAUNCHED();
  *mma_taken = S.flags;
  return PW_OK;
}

}  // namespace pw
