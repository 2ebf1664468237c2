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

// ---------------------------------------------------------------------------------------
// Layout operations either side of an apply:
//   resize of one subsystem (ProductState.resize_fock, state/composite_envelope.py:495-585),
//   kron of two states (CompositeEnvelope.combine -> kron_reduce, core/ops.py:53-64),
//   psi -> psi psi^dagger (state_expand, state/utils/state_transform.py:58-63),
//   rho -> psi for a pure rho (state_contract, state/utils/state_transform.py:122-137).
// ---------------------------------------------------------------------------------------
namespace pw {

// out has the same axes as in except dims[axis] = new_dim: zero padding or truncation.
// For a density matrix the tensor is dims ++ dims and BOTH copies of the axis are resized
// (axis2 >= 0).
template <typename V>
__global__ void __launch_bounds__(256)
k_resize(const V* __restrict__ in, V* __restrict__ out, const uint64_t outer, const uint64_t old_d,
         const uint64_t new_d, const uint64_t mid, const uint64_t inner, const bool two) {
  // tensor viewed as (outer, d, mid, d', inner) with d' present only when `two`
  const uint64_t d2n = two ? new_d : 1, d2o = two ? old_d : 1;
  const uint64_t total = outer * new_d * mid * d2n * inner;
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = o;
    const uint64_t i = r % inner; r /= inner;
    const uint64_t b = r % d2n; r /= d2n;
    const uint64_t m = r % mid; r /= mid;
    const uint64_t a = r % new_d; r /= new_d;
    V v = czero<V>();
    if (a < old_d && b < d2o) v = in[(((r * old_d + a) * mid + m) * d2o + b) * inner + i];
    out[o] = v;
  }
}

template <typename V>
__global__ void __launch_bounds__(256)
k_kron(const V* __restrict__ a, const V* __restrict__ b, V* __restrict__ out, const uint64_t ra,
       const uint64_t ca, const uint64_t rb, const uint64_t cb) {
  // out[(i*rb + j), (k*cb + l)] = a[i,k] * b[j,l]
  const uint64_t cols = ca * cb, total = ra * rb * cols;
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = o / cols, col = o - row * cols;
    const uint64_t i = row / rb, j = row - i * rb, k = col / cb, l = col - k * cb;
    const V x = a[i * ca + k], y = b[j * cb + l];
    V v;
    v.x = x.x * y.x - x.y * y.y;
    v.y = x.x * y.y + x.y * y.x;
    out[o] = v;
  }
}

template <typename V>
__global__ void __launch_bounds__(256)
k_outer(const V* __restrict__ psi, V* __restrict__ out, const uint64_t n) {
  const uint64_t total = n * n;
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = o / n, j = o - i * n;
    const V x = psi[i], y = psi[j];
    V v;
    v.x = x.x * y.x + x.y * y.y;   // x * conj(y)
    v.y = x.y * y.x - x.x * y.y;
    out[o] = v;
  }
}

// Pure rho = psi psi^dagger: psi = rho[:, c] / sqrt(rho[c, c]) for the largest diagonal entry
// c, then the global phase is removed so that psi[0] is real and non-negative (the
// reference's convention, state_transform.py:131-134).  info[0] = Tr(rho^2), info[1] = Tr rho;
// flag = 1 when |Tr(rho^2) - 1| < tol (only then psi is written).
template <typename V>
__global__ void __launch_bounds__(256)
k_contract(const V* __restrict__ rho, const uint64_t n, const double tol, V* __restrict__ psi,
           const double* __restrict__ stats, double* __restrict__ info, int* __restrict__ flag) {
  __shared__ double s_val[256];
  __shared__ uint64_t s_idx[256];
  __shared__ double s_acc[256];
  double best = -1.0, tr = 0.0;
  uint64_t bi = 0;
  for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = (double)rho[i * n + i].x;
    tr += d;
    if (d > best) { best = d; bi = i; }
  }
  s_val[threadIdx.x] = best; s_idx[threadIdx.x] = bi; s_acc[threadIdx.x] = tr;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      if (s_val[threadIdx.x + s] > s_val[threadIdx.x] ||
          (s_val[threadIdx.x + s] == s_val[threadIdx.x] && s_idx[threadIdx.x + s] < s_idx[threadIdx.x])) {
        s_val[threadIdx.x] = s_val[threadIdx.x + s];
        s_idx[threadIdx.x] = s_idx[threadIdx.x + s];
      }
      s_acc[threadIdx.x] += s_acc[threadIdx.x + s];
    }
    __syncthreads();
  }
  const uint64_t c = s_idx[0];
  const double purity = stats[0];  // sum |rho_ij|^2 = Tr(rho^2) for Hermitian rho
  const bool pure = fabs(purity - 1.0) < tol;
  if (threadIdx.x == 0) {
    info[0] = purity;
    info[1] = s_acc[0];
    *flag = pure ? 1 : 0;
  }
  if (!pure) return;
  // phase that makes psi[0] real >= 0: psi0 ~ rho[0, c]
  const V r0 = rho[c];
  const double a0 = sqrt((double)r0.x * r0.x + (double)r0.y * r0.y);
  double px = 1.0, py = 0.0;  // conj(rho[0,c]) / |rho[0,c]|
  if (a0 > 0) { px = (double)r0.x / a0; py = -(double)r0.y / a0; }
  const double inv = 1.0 / sqrt(s_val[0]);
  for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const V v = rho[i * n + c];
    V w;
    w.x = (typename RealOf<V>::type)(((double)v.x * px - (double)v.y * py) * inv);
    w.y = (typename RealOf<V>::type)(((double)v.x * py + (double)v.y * px) * inv);
    psi[i] = w;
  }
}

// out[0] = number of entries exactly equal to 1 + 0i, out[1] = index of the first of them
// (-1 if none): the Vector -> Label test of state_contract (state_transform.py:138-144).
template <typename V>
__global__ void __launch_bounds__(256)
k_find_label(const V* __restrict__ psi, const int64_t n, long long* __restrict__ out) {
  __shared__ long long s_cnt[256];
  __shared__ long long s_first[256];
  long long cnt = 0, first = -1;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const V v = psi[i];
    if (v.x == 1 && v.y == 0) {
      ++cnt;
      if (first < 0) first = i;
    }
  }
  s_cnt[threadIdx.x] = cnt; s_first[threadIdx.x] = first;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + s];
      const long long a = s_first[threadIdx.x], b = s_first[threadIdx.x + s];
      s_first[threadIdx.x] = a < 0 ? b : (b < 0 ? a : (a < b ? a : b));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = s_cnt[0]; out[1] = s_first[0]; }
}

}  // namespace pw

extern "C" int pw_resize_axis(const void* in, void* out, const int64_t* dims, int n, int axis,
                              int64_t new_dim, int is_matrix, int dtype, void* stream) {
  if (!in || !out || !dims || n < 1 || n > PW_MAX_SUBSYSTEMS || axis < 0 || axis >= n || new_dim < 1 || in == out)
    return PW_ERR_INVALID;
  uint64_t outer = 1, inner = 1, N = 1;
  for (int i = 0; i < n; ++i) {
    if (dims[i] < 1) return PW_ERR_INVALID;
    if (i < axis) outer *= (uint64_t)dims[i];
    if (i > axis) inner *= (uint64_t)dims[i];
    N *= (uint64_t)dims[i];
  }
  const uint64_t old_d = (uint64_t)dims[axis];
  // vector: (outer, d, 1, -, inner); matrix: rows (outer, d, inner) x cols (outer, d', inner)
  const uint64_t mid = is_matrix ? inner * outer : 1;
  const uint64_t total = outer * (uint64_t)new_dim * mid * (is_matrix ? (uint64_t)new_dim : 1) * inner;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for(total, 256, 8);
  if (dtype == PW_C128)
    k_resize<double2><<<grid, 256, 0, s>>>((const double2*)in, (double2*)out, outer, old_d, (uint64_t)new_dim, mid, inner, is_matrix != 0);
  else if (dtype == PW_C64)
    k_resize<float2><<<grid, 256, 0, s>>>((const float2*)in, (float2*)out, outer, old_d, (uint64_t)new_dim, mid, inner, is_matrix != 0);
  else
    return PW_ERR_INVALID;
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_kron(const void* a, const void* b, void* out, int64_t rows_a, int64_t cols_a,
                       int64_t rows_b, int64_t cols_b, int dtype, void* stream) {
  if (!a || !b || !out || rows_a < 1 || cols_a < 1 || rows_b < 1 || cols_b < 1) return PW_ERR_INVALID;
  const uint64_t total = (uint64_t)rows_a * rows_b * cols_a * cols_b;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for(total, 256, 8);
  if (dtype == PW_C128)
    k_kron<double2><<<grid, 256, 0, s>>>((const double2*)a, (const double2*)b, (double2*)out, rows_a, cols_a, rows_b, cols_b);
  else if (dtype == PW_C64)
    k_kron<float2><<<grid, 256, 0, s>>>((const float2*)a, (const float2*)b, (float2*)out, rows_a, cols_a, rows_b, cols_b);
  else
    return PW_ERR_INVALID;
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_expand_vec(const void* psi, void* rho, int64_t n, int dtype, void* stream) {
  if (!psi || !rho || n < 1) return PW_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for((uint64_t)n * n, 256, 8);
  if (dtype == PW_C128) k_outer<double2><<<grid, 256, 0, s>>>((const double2*)psi, (double2*)rho, (uint64_t)n);
  else if (dtype == PW_C64) k_outer<float2><<<grid, 256, 0, s>>>((const float2*)psi, (float2*)rho, (uint64_t)n);
  else return PW_ERR_INVALID;
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_contract_mat(const void* rho, int64_t n, double tol, void* psi, double* info,
                               int* flag, int dtype, void* stream) {
  if (!rho || !psi || !info || !flag || n < 1) return PW_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  Scratch stats;
  PW_TRY(stats.alloc(2 * sizeof(double), s));
  PW_TRY(pw_norm_stats(rho, n * n, dtype, (double*)stats.p, stream));  // full-grid reduction
  if (dtype == PW_C128) k_contract<double2><<<1, 256, 0, s>>>((const double2*)rho, (uint64_t)n, tol, (double2*)psi, (const double*)stats.p, info, flag);
  else if (dtype == PW_C64) k_contract<float2><<<1, 256, 0, s>>>((const float2*)rho, (uint64_t)n, tol, (float2*)psi, (const double*)stats.p, info, flag);
  else return PW_ERR_INVALID;
  PW_LAUNCHED();
  return PW_OK;
}

extern "C" int pw_find_label(const void* psi, int64_t n, int dtype, int64_t* out, void* stream) {
  if (!psi || !out || n < 1) return PW_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PW_C128) k_find_label<double2><<<1, 256, 0, s>>>((const double2*)psi, n, (long long*)out);
  else if (dtype == PW_C64) k_find_label<float2><<<1, 256, 0, s>>>((const float2*)psi, n, (long long*)out);
  else return PW_ERR_INVALID;
  PW_LAUNCHED();
  return PW_OK;
}
