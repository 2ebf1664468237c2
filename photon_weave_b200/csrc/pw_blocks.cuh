// Block-structured operator descriptor (see pw_blocks.cu).
#pragma once
#include "pw_internal.cuh"

namespace pw {

constexpr int kBlockMaxN = 8;     // largest dense block the register kernel takes
constexpr uint32_t kBlockMaxT = 128;  // largest operator dimension analysed on the device

struct BlockDescHeader {
  int nblocks;
  int maxn;    // largest block side
  int nzero;   // all-zero rows (outputs that are simply cleared)
  int usable;  // 1: the block kernel handles this operator; fallback kernels must exit
  int inplace_safe;  // 1: no block writes a position a DIFFERENT block still has to read
  int seg_handled;   // usable && inplace_safe: the staged (segmented) kernel took it
  int n_offs;        // entries used in offs[]
  int n_vals;        // entries used in vals[]
};

// device pointers of one analysed operator
struct BlockArrays {
  BlockDescHeader* hdr = nullptr;
  int4* blocks = nullptr;        // per block: nr, nc, offset into offs, offset into vals
  int64_t* offs = nullptr;       // per block: nc column offsets then nr row offsets (elements)
  int64_t* zero_offs = nullptr;  // offsets of all-zero rows
  void* vals = nullptr;          // per block: dense nr x nc values (row-major)
};

struct BlockDesc {
  Scratch mem, mem2;  // descriptor; tensor-core tables of the staged kernel
  BlockArrays arr;
  BlockDescHeader* hdr = nullptr;
  bool staged = false;  // launch_blocks_seg was used: the flag to honour is seg_handled
  const int* mma_flag = nullptr;  // staged, tensor-core variant only: its own "taken" flag
  const int* skip_flag() const {
    if (mma_flag) return mma_flag;
    return hdr ? (staged ? &hdr->seg_handled : &hdr->usable) : nullptr;
  }
};

#ifdef __CUDACC__
// Device view of a descriptor (global memory, or a shared-memory copy made by the CTA).
template <typename V>
struct DescView {
  const int4* blocks;
  const int64_t* offs;
  const V* vals;
  const int64_t* zero_offs;
  int nblocks, nzero;
};

// bytes a shared-memory copy of the descriptor needs (16-byte aligned sections)
__device__ __forceinline__ uint32_t desc_smem_bytes(const BlockDescHeader* h, uint32_t vsize) {
  auto al = [](uint32_t x) { return (x + 15u) & ~15u; };
  return al((uint32_t)h->nblocks * 16u) + al((uint32_t)h->n_offs * 8u) +
         al((uint32_t)h->nzero * 8u) + al((uint32_t)h->n_vals * vsize);
}

// Copies the descriptor into `smem` when it fits `cap` bytes (all threads of the CTA must
// call this; contains a __syncthreads) and returns a view of whichever copy to use.
template <typename V>
__device__ __forceinline__ DescView<V> desc_view(const BlockArrays& A, unsigned char* smem, uint32_t cap) {
  DescView<V> v;
  const BlockDescHeader* h = A.hdr;
  v.nblocks = h->nblocks;
  v.nzero = h->nzero;
  v.blocks = A.blocks; v.offs = A.offs; v.vals = (const V*)A.vals; v.zero_offs = A.zero_offs;
  const uint32_t need = desc_smem_bytes(h, (uint32_t)sizeof(V));
  if (smem != nullptr && need <= cap) {
    auto al = [](uint32_t x) { return (x + 15u) & ~15u; };
    int4* sb = (int4*)smem;
    int64_t* so = (int64_t*)(smem + al((uint32_t)h->nblocks * 16u));
    int64_t* sz = (int64_t*)((unsigned char*)so + al((uint32_t)h->n_offs * 8u));
    V* sv = (V*)((unsigned char*)sz + al((uint32_t)h->nzero * 8u));
    for (int i = threadIdx.x; i < h->nblocks; i += blockDim.x) sb[i] = A.blocks[i];
    for (int i = threadIdx.x; i < h->n_offs; i += blockDim.x) so[i] = A.offs[i];
    for (int i = threadIdx.x; i < h->nzero; i += blockDim.x) sz[i] = A.zero_offs[i];
    for (int i = threadIdx.x; i < h->n_vals; i += blockDim.x) sv[i] = ((const V*)A.vals)[i];
    v.blocks = sb; v.offs = so; v.vals = sv; v.zero_offs = sz;
  }
  __syncthreads();
  return v;
}

// y[rows] = B (nr x NC, dense) x[cols]; NC is a compile-time constant so the NC inputs
// live in registers without predication.
// SYNCW: __syncwarp() between the loads and the stores (in-place use where a lane's outputs
// land on positions another lane of the same warp reads)
template <typename V, int NC, int UR, bool SYNCW, typename LoadX, typename StoreY>
__device__ __forceinline__ void block_product_nc(const V* __restrict__ bv, const int nr, LoadX ldx, StoreY sty) {
  using X = decltype(ldx(0));  // V, or C64x2: two adjacent complex64 fibers per thread
  X x[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) x[j] = ldx(j);
  if (SYNCW) __syncwarp();
  // four output rows at a time: eight independent FMA chains per thread keep the FP64 pipe
  // busy at low occupancy (the dependent chain of one row is 2 * NC FMAs long)
  int a = 0;
  if (UR >= 4)
  for (; a + 4 <= nr; a += 4) {
    X acc0 = czero<X>(), acc1 = czero<X>(), acc2 = czero<X>(), acc3 = czero<X>();
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      cfma(acc0, bv[a * NC + j], x[j]);
      cfma(acc1, bv[(a + 1) * NC + j], x[j]);
      cfma(acc2, bv[(a + 2) * NC + j], x[j]);
      cfma(acc3, bv[(a + 3) * NC + j], x[j]);
    }
    sty(a, acc0); sty(a + 1, acc1); sty(a + 2, acc2); sty(a + 3, acc3);
  }
  for (; a < nr; ++a) {
    X acc = czero<X>();
#pragma unroll
    for (int j = 0; j < NC; ++j) cfma(acc, bv[a * NC + j], x[j]);
    sty(a, acc);
  }
}

template <typename V, int UR = 1, bool SYNCW = false, typename LoadX, typename StoreY>
__device__ __forceinline__ void block_product(const V* __restrict__ bv, const int nr, const int nc, LoadX ldx, StoreY sty) {
  switch (nc) {
    case 1: block_product_nc<V, 1, UR, SYNCW>(bv, nr, ldx, sty); break;
    case 2: block_product_nc<V, 2, UR, SYNCW>(bv, nr, ldx, sty); break;
    case 3: block_product_nc<V, 3, UR, SYNCW>(bv, nr, ldx, sty); break;
    case 4: block_product_nc<V, 4, UR, SYNCW>(bv, nr, ldx, sty); break;
    case 5: block_product_nc<V, 5, UR, SYNCW>(bv, nr, ldx, sty); break;
    case 6: block_product_nc<V, 6, UR, SYNCW>(bv, nr, ldx, sty); break;
    case 7: block_product_nc<V, 7, UR, SYNCW>(bv, nr, ldx, sty); break;
    case 8: block_product_nc<V, 8, UR, SYNCW>(bv, nr, ldx, sty); break;
    default: break;  // nc == 0 cannot occur (blocks have >= 1 column); nc > 8 is not `usable`
  }
}
#endif  // __CUDACC__

// Enqueues the structure analysis of a dense t x t operator (t <= kBlockMaxT); element
// offsets come from `tm` (global strides for the direct kernel, tile strides for the
// tiled kernel).
// `skip` (optional device flag): when set the analysis is not run and the descriptor reports
// "not usable".
int build_block_desc(const void* op, uint32_t t, bool conj_op, const TargetMap& tm, int dtype,
                     Scratch* mem, BlockArrays* arr, cudaStream_t stream, const int* skip = nullptr);

// Enqueues structure analysis + the block apply kernels.  `conj_op`: use conj(op).
// Requires in != out.  After this call the caller must enqueue a fallback path whose
// kernels exit when *desc->skip_flag() != 0.
int launch_blocks(const void* in, void* out, const void* op, const Plan& plan, int dtype,
                  bool conj_op, BlockDesc* desc, cudaStream_t stream);

// Staged variant for layouts whose innermost axis is a target: a warp stages 32 adjacent
// fibers in shared memory with coalesced cp.async rows, applies the blocks in place (one
// fiber per lane), and streams the rows back.  Runs when the analysis reports
// usable && inplace_safe (desc->hdr->seg_handled); the fallback must take that as skip flag.
// Returns PW_ERR_UNSUPPORTED when the geometry does not fit.
// `mma_only`: enqueue only the tensor-core variant (complex128, t <= 64; used for t <= 8, where
// the register kernels are the fallback); PW_ERR_UNSUPPORTED when it does not apply.
int launch_blocks_seg(const void* in, void* out, const void* op, const int64_t* axes, int naxes,
                      const int* targets, int nt, int dtype, bool conj_op, BlockDesc* desc,
                      cudaStream_t stream, bool mma_only = false);

// Two-sided single pass rho' = U rho U^dagger for a dense d <= 8 or block-structured operator
// (see pw_blocks.cu).  `fm` enumerates the fibers of the doubled tensor (every axis that is
// neither a row nor a column target), `rows` / `cols` map the operator index to element
// offsets on the two sides.  `other_handled` (device flag or nullptr): a path enqueued
// earlier that already produced `out`.  After the call `ts->skip` is the device flag the
// fallback kernels must honour.  Requires rho != out and an untouched innermost axis.
struct TwoSided {
  Scratch memL, memR, flag_mem;
  BlockArrays L, R;
  const int* skip = nullptr;
};
int launch_two_sided(const void* rho, void* out, const void* op, const FiberMap& fm,
                     const TargetMap& rows, const TargetMap& cols, int dtype, const int* other_handled,
                     TwoSided* ts, cudaStream_t stream);

}  // namespace pw
