// Host-buffer entry point: the end-to-end path a caller with NumPy/host arrays
// takes (H2D upload, the operation list on the device, D2H download).
#include <vector>

#include "pw_internal.cuh"

using namespace pw;

namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
};
struct Stream {
  cudaStream_t s = nullptr;
  ~Stream() { if (s) cudaStreamDestroy(s); }
};
}  // namespace

// Runs an operation list on device buffers, ping-ponging between `cur` and `other`
// (other == nullptr: in place).  Adjacent single-subsystem vector operators on different
// axes commute; with `fuse` they are applied as ONE pass (pw_apply_vec_pair).
static int run_ops(void** cur, void** other, int is_matrix, const int64_t* dims, int n,
                   const pw_host_op* ops, void* const* dev_ops, int nops, int dtype, int fuse,
                   cudaStream_t s) {
  int i = 0;
  while (i < nops) {
    const pw_host_op& o = ops[i];
    void* dst = *other ? *other : *cur;
    int rc = PW_ERR_UNSUPPORTED;
    int used = 1;
    if (fuse && *other && o.kind == 0 && o.nt == 1 && i + 1 < nops && ops[i + 1].kind == 0 &&
        ops[i + 1].nt == 1 && ops[i + 1].targets[0] != o.targets[0]) {
      rc = pw_apply_vec_pair(*cur, dst, dev_ops[i], dev_ops[i + 1], dims, n, o.targets[0],
                             ops[i + 1].targets[0], dtype, s);
      if (rc == PW_OK) used = 2;
    }
    if (rc == PW_ERR_UNSUPPORTED) {
      if (o.kind == 0)
        rc = pw_apply_vec(*cur, dst, dev_ops[i], dims, n, o.targets, o.nt, dtype, s);
      else if (o.kind == 1)
        rc = pw_apply_mat(*cur, dst, dev_ops[i], dims, n, o.targets, o.nt, dtype, s);
      else
        rc = pw_kraus_mat(*cur, dst, dev_ops[i], o.K, dims, n, o.targets, o.nt, dtype, s);
    }
    if (rc != PW_OK) return rc;
    if (*other) { void* t = *cur; *cur = *other; *other = t; }
    i += used;
  }
  (void)is_matrix;
  return PW_OK;
}

static int check_ops(int is_matrix, const int64_t* dims, int n, const pw_host_op* ops, int nops) {
  for (int i = 0; i < nops; ++i) {
    const pw_host_op& o = ops[i];
    if (o.nt < 1 || o.nt > PW_MAX_TARGETS || !o.op) return PW_ERR_INVALID;
    if ((o.kind != 0) != (is_matrix != 0)) return PW_ERR_INVALID;
    if (o.kind == 2 && o.K < 1) return PW_ERR_INVALID;
    for (int k = 0; k < o.nt; ++k)
      if (o.targets[k] < 0 || o.targets[k] >= n) return PW_ERR_INVALID;
  }
  (void)dims;
  return PW_OK;
}

extern "C" int pw_circuit_device(void* state, void* scratch, int is_matrix, const int64_t* dims,
                                 int n, const pw_host_op* ops, int nops, int dtype, int fuse,
                                 void* stream, int* result_in_scratch) {
  if (!state || !dims || n < 1 || n > PW_MAX_SUBSYSTEMS || nops < 0 || (nops && !ops)) return PW_ERR_INVALID;
  if (dtype != PW_C64 && dtype != PW_C128) return PW_ERR_INVALID;
  PW_TRY(check_ops(is_matrix, dims, n, ops, nops));
  std::vector<void*> dev_ops(nops);
  for (int i = 0; i < nops; ++i) dev_ops[i] = const_cast<void*>(ops[i].op);
  void* cur = state;
  void* other = scratch;
  PW_TRY(run_ops(&cur, &other, is_matrix, dims, n, ops, dev_ops.data(), nops, dtype, fuse,
                 (cudaStream_t)stream));
  if (result_in_scratch) *result_in_scratch = (cur != state) ? 1 : 0;
  return PW_OK;
}

extern "C" int pw_circuit_host(const void* h_state, void* h_out, int is_matrix,
                               const int64_t* dims, int n, const pw_host_op* ops, int nops,
                               int dtype, int device) {
  if (!h_state || !h_out || !dims || n < 1 || n > PW_MAX_SUBSYSTEMS || nops < 0 || (nops && !ops))
    return PW_ERR_INVALID;
  if (dtype != PW_C64 && dtype != PW_C128) return PW_ERR_INVALID;
  PW_CUDA(cudaSetDevice(device));
  const size_t es = dtype == PW_C128 ? 16 : 8;
  uint64_t N = 1;
  for (int i = 0; i < n; ++i) {
    if (dims[i] < 1) return PW_ERR_INVALID;
    N *= (uint64_t)dims[i];
  }
  const uint64_t count = is_matrix ? N * N : N;
  const size_t bytes = count * es;

  Stream st;
  PW_CUDA(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
  DevBuf state, other;
  PW_CUDA(cudaMalloc(&state.p, bytes));
  // ping-pong when a second buffer fits (lets the block-structured path run); else in place
  if (cudaMalloc(&other.p, bytes) != cudaSuccess) {
    other.p = nullptr;
    (void)cudaGetLastError();
  }

  // operators: one small upload each, queued ahead of the state so they overlap it
  std::vector<DevBuf> dops(nops);
  for (int i = 0; i < nops; ++i) {
    const pw_host_op& o = ops[i];
    if (o.nt < 1 || o.nt > PW_MAX_TARGETS || !o.op) return PW_ERR_INVALID;
    if ((o.kind != 0) != (is_matrix != 0)) return PW_ERR_INVALID;
    uint64_t t = 1;
    for (int k = 0; k < o.nt; ++k) {
      if (o.targets[k] < 0 || o.targets[k] >= n) return PW_ERR_INVALID;
      t *= (uint64_t)dims[o.targets[k]];
    }
    const int K = o.kind == 2 ? o.K : 1;
    if (K < 1) return PW_ERR_INVALID;
    const size_t ob = (size_t)K * t * t * es;
    PW_CUDA(cudaMalloc(&dops[i].p, ob));
    PW_CUDA(cudaMemcpyAsync(dops[i].p, o.op, ob, cudaMemcpyHostToDevice, st.s));
  }
  PW_CUDA(cudaMemcpyAsync(state.p, h_state, bytes, cudaMemcpyHostToDevice, st.s));
  {
    std::vector<void*> dev_ops(nops);
    for (int i = 0; i < nops; ++i) dev_ops[i] = dops[i].p;
    const int fuse = pw::options().fuse_pairs.load(std::memory_order_relaxed);
    const int rc = run_ops(&state.p, &other.p, is_matrix, dims, n, ops, dev_ops.data(), nops, dtype,
                           fuse, st.s);
    if (rc != PW_OK) {
      cudaStreamSynchronize(st.s);
      return rc;
    }
  }
  PW_CUDA(cudaMemcpyAsync(h_out, state.p, bytes, cudaMemcpyDeviceToHost, st.s));
  PW_CUDA(cudaStreamSynchronize(st.s));
  return PW_OK;
}
