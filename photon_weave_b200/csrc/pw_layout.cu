// Device-side layout operations: general axis permutation of a row-major tensor
// (ProductState.reorder, state/composite_envelope.py:201-248; reorder_vector_front /
// reorder_matrix_front, core/adapters.py:210-296) and the pack step of the multi-GPU
// axis swap.
#include "pw_internal.cuh"

namespace pw {

struct PermMap {
  int n;                       // merged axes, OUTPUT order (outermost first)
  uint64_t osize[kMaxAxes];    // extent of output axis k
  int64_t istride[kMaxAxes];   // element stride of that axis in the INPUT tensor
  uint64_t total;
};

template <typename V>
__global__ void __launch_bounds__(256)
k_permute(const V* __restrict__ in, V* __restrict__ out, const PermMap pm) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < pm.total; o += stride) {
    uint64_t r = o;
    int64_t off = 0;
    for (int k = pm.n - 1; k > 0; --k) {
      const uint64_t q = r / pm.osize[k];
      off += (int64_t)(r - q * pm.osize[k]) * pm.istride[k];
      r = q;
    }
    off += (int64_t)r * pm.istride[0];
    out[o] = in[off];
  }
}

}  // namespace pw

using namespace pw;

extern "C" int pw_permute(const void* in, void* out, const int64_t* dims, int n,
                          const int32_t* perm, int dtype, void* stream) {
  if (!in || !out || !dims || !perm || n < 1 || n > kMaxAxes || in == out) return PW_ERR_INVALID;
  int64_t stride[kMaxAxes];
  bool seen[kMaxAxes] = {false};
  uint64_t total = 1;
  for (int i = n - 1; i >= 0; --i) {
    if (dims[i] < 1) return PW_ERR_INVALID;
    stride[i] = (int64_t)total;
    total *= (uint64_t)dims[i];
  }
  PermMap pm;
  pm.n = 0;
  pm.total = total;
  for (int k = 0; k < n; ++k) {
    const int a = perm[k];
    if (a < 0 || a >= n || seen[a]) return PW_ERR_INVALID;
    seen[a] = true;
    if (dims[a] == 1) continue;
    // merge with the previous output axis when they are also adjacent in the input
    if (pm.n > 0 && pm.istride[pm.n - 1] == stride[a] * dims[a]) {
      pm.osize[pm.n - 1] *= (uint64_t)dims[a];
      pm.istride[pm.n - 1] = stride[a];
    } else {
      pm.osize[pm.n] = (uint64_t)dims[a];
      pm.istride[pm.n] = stride[a];
      ++pm.n;
    }
  }
  if (pm.n == 0) { pm.n = 1; pm.osize[0] = 1; pm.istride[0] = 1; }
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for(total, 256, 8);
  if (dtype == PW_C128)
    k_permute<double2><<<grid, 256, 0, s>>>((const double2*)in, (double2*)out, pm);
  else if (dtype == PW_C64)
    k_permute<float2><<<grid, 256, 0, s>>>((const float2*)in, (float2*)out, pm);
  else
    return PW_ERR_INVALID;
  PW_LAUNCHED();
  return PW_OK;
}
