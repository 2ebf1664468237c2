// Peer-memory (NVLink) path of the multi-GPU axis swap.
//
// One process per GPU.  Shard buffers are cudaMalloc'ed here and exported with CUDA IPC
// so every rank maps every other rank's buffers.  The axis swap then needs no
// collective library call: a rank READS the piece it needs straight out of its peers'
// buffers over NVLink -- either as a plain gather (pw_gather_chunks) or FUSED with the
// operator application that follows (pw_apply_vec_gather), so the swapped data is
// never written to and re-read from local HBM.
#include <cstring>

#include "pw_internal.cuh"
#include "pw_blocks.cuh"

namespace pw {

constexpr int kMaxPeers = 16;

struct SegIn {
  const void* ptr[kMaxPeers];  // ptr[p] = start of the chunk read from peer p
  uint64_t chunk;              // elements per chunk
};

template <typename V>
__device__ __forceinline__ const V* seg_addr(const SegIn& s, uint64_t e, bool small) {
  uint64_t p, rem;
  if (small) {
    const uint32_t e32 = (uint32_t)e, c32 = (uint32_t)s.chunk;
    const uint32_t q = e32 / c32;
    p = q; rem = e32 - q * c32;
  } else {
    p = e / s.chunk; rem = e - p * s.chunk;
  }
  return (const V*)s.ptr[p] + rem;
}

// out[p * chunk + x] = peer_p[x]: coalesced 128-bit loads over NVLink, local stores
template <typename V>
__global__ void __launch_bounds__(256)
k_gather_chunks(const SegIn s, const int W, const uint64_t rot, V* __restrict__ out) {
  const uint64_t total = s.chunk * (uint64_t)W;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    // rotated start: at any moment the W ranks pull from W different peers
    uint64_t e = i + rot;
    if (e >= total) e -= total;
    const uint64_t p = e / s.chunk;
    out[e] = ((const V*)s.ptr[p])[e - p * s.chunk];
  }
}

// Register apply kernel (t <= 8) whose INPUT is the virtual concatenation of W peer
// chunks: the element offset decides which peer pointer a load goes to.
template <typename V, int T>
__global__ void __launch_bounds__(256, (sizeof(V) == 16 ? (T <= 4 ? 4 : (T <= 6 ? 3 : 2)) : 3))
k_apply_small_gather(const SegIn s, V* __restrict__ out, const V* __restrict__ op,
                     const FiberMap fm, const TargetMap tm, const bool small, const uint64_t rot) {
  __shared__ V s_op[T * T];
  __shared__ int64_t s_off[T];
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) s_op[i] = op[i];
  if (threadIdx.x < T) s_off[threadIdx.x] = tm.offset(threadIdx.x);
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < fm.nfib; i += stride) {
    uint64_t f = i + rot;  // rotated start: ranks hit different peers at any moment
    if (f >= fm.nfib) f -= fm.nfib;
    const int64_t base = fm.base(f);
    int zero = 0;
    asm volatile("" : "+r"(zero));
    const V* sop = s_op + zero;
    V x[T];
#pragma unroll
    for (int j = 0; j < T; ++j) x[j] = *seg_addr<V>(s, (uint64_t)(base + s_off[j]), small);
#pragma unroll
    for (int a = 0; a < T; ++a) {
      V acc = czero<V>();
#pragma unroll
      for (int b = 0; b < T; ++b) cfma(acc, sop[a * T + b], x[b]);
      out[base + s_off[a]] = acc;
    }
  }
}

template <typename V, int T>
static void launch_gather_small(const SegIn& s, V* out, const V* op, const Plan& p, uint64_t rot, cudaStream_t st) {
  const bool small = p.N <= 0xffffffffull;
  k_apply_small_gather<V, T><<<grid_for(p.fm.nfib, 256, 8), 256, 0, st>>>(s, out, op, p.fm, p.tm, small, rot);
}

template <typename V>
static int apply_gather_t(const SegIn& s, V* out, const V* op, const Plan& p, int rank, int W, cudaStream_t st) {
  const uint64_t rot = (p.fm.nfib / (uint64_t)W) * (uint64_t)(rank % W);
  switch (p.tm.t) {
    case 1: launch_gather_small<V, 1>(s, out, op, p, rot, st); break;
    case 2: launch_gather_small<V, 2>(s, out, op, p, rot, st); break;
    case 3: launch_gather_small<V, 3>(s, out, op, p, rot, st); break;
    case 4: launch_gather_small<V, 4>(s, out, op, p, rot, st); break;
    case 5: launch_gather_small<V, 5>(s, out, op, p, rot, st); break;
    case 6: launch_gather_small<V, 6>(s, out, op, p, rot, st); break;
    case 7: launch_gather_small<V, 7>(s, out, op, p, rot, st); break;
    case 8: launch_gather_small<V, 8>(s, out, op, p, rot, st); break;
    default: return PW_ERR_UNSUPPORTED;
  }
  PW_LAUNCHED();
  return PW_OK;
}

// Block-structured operators (8 < t <= 128: beam splitters, two-mode phases) with the same
// gathered input: k_apply_blocks whose loads go through the peer pointers.  One fiber per thread,
// every block gathered into registers, the result written once to local memory.
template <typename V>
__global__ void __launch_bounds__(256, 3)
k_apply_blocks_gather(const SegIn s, V* __restrict__ out, const FiberMap fm, const BlockArrays A,
                      const bool small, const uint64_t rot) {
  extern __shared__ __align__(16) unsigned char s_desc[];
  const DescView<V> dv = desc_view<V>(A, s_desc, 24 * 1024);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < fm.nfib; i += stride) {
    uint64_t f = i + rot;
    if (f >= fm.nfib) f -= fm.nfib;
    const int64_t base = fm.base(f);
    V* dst = out + base;
    for (int b = 0; b < dv.nblocks; ++b) {
      const int4 h = dv.blocks[b];
      const int64_t* bo = dv.offs + h.z;
      block_product<V>(dv.vals + h.w, h.x, h.y,
                       [&](int j) { return *seg_addr<V>(s, (uint64_t)(base + bo[j]), small); },
                       [&](int a, V v) { __stcg(dst + bo[h.y + a], v); });
    }
    for (int z = 0; z < dv.nzero; ++z) __stcg(dst + dv.zero_offs[z], czero<V>());
  }
}

template <typename V>
static int apply_gather_blocks_t(const SegIn& s, V* out, const void* op, const Plan& p, int dtype, int rank,
                                 int W, cudaStream_t st) {
  Scratch mem;
  BlockArrays arr;
  PW_TRY(build_block_desc(op, p.tm.t, false, p.tm, dtype, &mem, &arr, st));
  // the structure analysis runs on the device; this entry point reads its verdict back (one
  // 32-byte copy + stream synchronisation) because the caller needs to know whether to fall
  // back to gather + apply -- the only libpwb200 device entry point that synchronises
  BlockDescHeader h;
  PW_CUDA(cudaMemcpyAsync(&h, arr.hdr, sizeof(h), cudaMemcpyDeviceToHost, st));
  PW_CUDA(cudaStreamSynchronize(st));
  if (!h.usable) return PW_ERR_UNSUPPORTED;
  const uint64_t rot = (p.fm.nfib / (uint64_t)W) * (uint64_t)(rank % W);
  k_apply_blocks_gather<V><<<grid_for(p.fm.nfib, 256, 3), 256, 24 * 1024, st>>>(s, out, p.fm, arr,
                                                                              p.N <= 0xffffffffull, rot);
  PW_LAUNCHED();
  return PW_OK;
}

}  // namespace pw

using namespace pw;

extern "C" int pw_peer_alloc(size_t bytes, void** ptr, unsigned char handle[64]) {
  if (!ptr || !handle) return PW_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  PW_CUDA(cudaMalloc(ptr, bytes));
  cudaIpcMemHandle_t h;
  const int rc = cuda_status(cudaIpcGetMemHandle(&h, *ptr));
  if (rc != PW_OK) { cudaFree(*ptr); *ptr = nullptr; return rc; }
  memcpy(handle, &h, 64);
  return PW_OK;
}

extern "C" int pw_peer_open(const unsigned char handle[64], void** ptr) {
  if (!ptr || !handle) return PW_ERR_INVALID;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  PW_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return PW_OK;
}

extern "C" int pw_peer_close(void* ptr) {
  PW_CUDA(cudaIpcCloseMemHandle(ptr));
  return PW_OK;
}

extern "C" int pw_peer_free(void* ptr) {
  PW_CUDA(cudaFree(ptr));
  return PW_OK;
}

extern "C" int pw_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) return PW_ERR_INVALID;
  // unified addressing: src may be an IPC-mapped peer buffer; the copy runs on a copy engine
  // (no SM), so it overlaps the apply kernels of the pipelined axis swap
  PW_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return PW_OK;
}

static int make_seg(const void* const* srcs, int W, int64_t chunk, SegIn* s) {
  if (!srcs || W < 1 || W > kMaxPeers || chunk < 1) return PW_ERR_INVALID;
  for (int p = 0; p < W; ++p) {
    if (!srcs[p]) return PW_ERR_INVALID;
    s->ptr[p] = srcs[p];
  }
  s->chunk = (uint64_t)chunk;
  return PW_OK;
}

extern "C" int pw_gather_chunks(const void* const* srcs, int W, int64_t chunk_elems, int rank,
                                void* out, int dtype, void* stream) {
  SegIn s;
  PW_TRY(make_seg(srcs, W, chunk_elems, &s));
  if (!out) return PW_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for((uint64_t)chunk_elems * W, 256, 8);
  const uint64_t rot = (uint64_t)chunk_elems * (uint64_t)(((rank % W) + W) % W);
  if (dtype == PW_C128) k_gather_chunks<double2><<<grid, 256, 0, st>>>(s, W, rot, (double2*)out);
  else if (dtype == PW_C64) k_gather_chunks<float2><<<grid, 256, 0, st>>>(s, W, rot, (float2*)out);
  else return PW_ERR_INVALID;
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_apply_vec_gather(const void* const* srcs, int W, int64_t chunk_elems, int rank,
                                   void* out, const void* op, const int64_t* dims, int n,
                                   const int32_t* targets, int nt, int dtype, void* stream) {
  SegIn s;
  PW_TRY(make_seg(srcs, W, chunk_elems, &s));
  if (!out || !op || nt < 1) return PW_ERR_INVALID;
  Plan p;
  PW_TRY(build_plan(dims, n, targets, nt, &p));
  if (p.N != (uint64_t)chunk_elems * (uint64_t)W) return PW_ERR_INVALID;
  const int r = rank < 0 ? 0 : rank;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.tm.t > 8) {  // block-structured operators up to 128 x 128 (dense ones: PW_ERR_UNSUPPORTED)
    if (p.tm.t > kBlockMaxT) return PW_ERR_UNSUPPORTED;
    if (dtype == PW_C128) return apply_gather_blocks_t<double2>(s, (double2*)out, op, p, dtype, r, W, st);
    if (dtype == PW_C64) return apply_gather_blocks_t<float2>(s, (float2*)out, op, p, dtype, r, W, st);
    return PW_ERR_INVALID;
  }
  if (dtype == PW_C128) return apply_gather_t<double2>(s, (double2*)out, (const double2*)op, p, r, W, st);
  if (dtype == PW_C64) return apply_gather_t<float2>(s, (float2*)out, (const float2*)op, p, r, W, st);
  return PW_ERR_INVALID;
}
