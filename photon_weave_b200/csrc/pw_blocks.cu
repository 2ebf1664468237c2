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
    hdr->maxn = s_maxn;
    hdr->nzero = s_nzero;
    hdr->usable = (s_maxn <= kBlockMaxN) ? 1 : 0;
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
template <typename V, int MAXN>
__global__ void __launch_bounds__(256, (sizeof(V) == 16 ? (MAXN <= 4 ? 4 : 3) : 4))
k_apply_blocks(const V* __restrict__ in, V* __restrict__ out, const FiberMap fm,
               const BlockDescHeader* __restrict__ hdr, const int4* __restrict__ blocks,
               const int64_t* __restrict__ offs, const V* __restrict__ vals,
               const int64_t* __restrict__ zero_offs) {
  // each instantiation owns max block sizes in (MAXN/2, MAXN] (MAXN = 2 also takes 0..1)
  constexpr int LO = MAXN == 2 ? -1 : MAXN / 2;
  const int maxn = hdr->maxn;
  if (!hdr->usable || maxn <= LO || maxn > MAXN) return;
  const int nblocks = hdr->nblocks, nzero = hdr->nzero;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < fm.nfib; f += stride) {
    const int64_t base = fm.base(f);
    const V* src = in + base;
    V* dst = out + base;
    for (int b = 0; b < nblocks; ++b) {
      const int4 h = __ldg(blocks + b);
      const int nr = h.x, nc = h.y;
      const int64_t* bo = offs + h.z;
      const V* bv = vals + h.w;
      V x[MAXN];
#pragma unroll
      for (int j = 0; j < MAXN; ++j)
        if (j < nc) x[j] = src[__ldg(bo + j)];
      for (int a = 0; a < nr; ++a) {
        V acc = czero<V>();
#pragma unroll
        for (int j = 0; j < MAXN; ++j)
          if (j < nc) cfma(acc, __ldg(bv + a * nc + j), x[j]);
        dst[__ldg(bo + nc + a)] = acc;
      }
    }
    for (int z = 0; z < nzero; ++z) dst[__ldg(zero_offs + z)] = czero<V>();
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
  const int grid = grid_for(p.fm.nfib, 256, 8);
  // three instantiations cover max block sizes [0,2], (2,4], (4,8]; two of them exit at once
  k_apply_blocks<V, 2><<<grid, 256, 0, s>>>(in, out, p.fm, A.hdr, A.blocks, A.offs, (const V*)A.vals, A.zero_offs);
  PW_LAUNCHED();
  k_apply_blocks<V, 4><<<grid, 256, 0, s>>>(in, out, p.fm, A.hdr, A.blocks, A.offs, (const V*)A.vals, A.zero_offs);
  PW_LAUNCHED();
  k_apply_blocks<V, 8><<<grid, 256, 0, s>>>(in, out, p.fm, A.hdr, A.blocks, A.offs, (const V*)A.vals, A.zero_offs);
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

}  // namespace pw
