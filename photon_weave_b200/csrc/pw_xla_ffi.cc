// XLA typed-FFI shim over the C ABI of libpwb200 (include/pw_b200.h): one handler per seam
// operation of photon_weave/core/adapters.py, so that the reference's JAX code can call the
// sm_100a kernels as `jax.ffi.ffi_call` custom calls with zero-copy device buffers
// (BASELINE.json north_star; SURVEY.md section 8b "On top (i)").
//
// GATED BUILD: this file needs the XLA FFI headers that ship inside jaxlib
// (`python -c "import jax.ffi; print(jax.ffi.include_dir())"`).  jax / jaxlib are NOT part of the
// build image, so `make` skips it there; on a machine with jax:
//
//     make -C photon_weave_b200/csrc xla        # -> photon_weave_b200/lib/libpwb200_xla.so
//
// and photon_weave_b200/core/jax_ffi.py registers the handlers below.  Handlers only enqueue on
// the stream XLA hands them, capture no Python state and are re-entrant (XLA calls them from its
// own executor thread).  Static arguments (dims, targets) arrive as int64 array attributes.
#include <cuda_runtime_api.h>

#include <cstdint>
#include <string>
#include <vector>

#include "pw_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using I64s = ffi::Span<const int64_t>;

// PW_C64 / PW_C128 from the buffer's element type; -1 otherwise
int dtype_of(const ffi::AnyBuffer& b) {
  switch (b.element_type()) {
    case ffi::DataType::C64: return PW_C64;
    case ffi::DataType::C128: return PW_C128;
    default: return -1;
  }
}

ffi::Error status(int rc, const char* what) {
  if (rc == PW_OK) return ffi::Error::Success();
  std::string msg = std::string(what) + ": " + pw_status_string(rc);
  if (rc == PW_ERR_INVALID) return ffi::Error::InvalidArgument(msg);
  return ffi::Error::Internal(msg);
}

std::vector<int32_t> to_i32(I64s v) {
  std::vector<int32_t> out(v.size());
  for (size_t i = 0; i < v.size(); ++i) out[i] = static_cast<int32_t>(v[i]);
  return out;
}

int64_t prod(I64s dims) {
  int64_t n = 1;
  for (int64_t d : dims) n *= d;
  return n;
}

// ---- apply ----------------------------------------------------------------------------------

ffi::Error ApplyVecImpl(cudaStream_t stream, ffi::AnyBuffer psi, ffi::AnyBuffer op, I64s dims,
                        I64s targets, ffi::Result<ffi::AnyBuffer> out) {
  const int dt = dtype_of(psi);
  if (dt < 0 || dtype_of(op) != dt) return ffi::Error::InvalidArgument("pw_apply_vec: complex64/128 buffers of one dtype expected");
  if ((int64_t)psi.element_count() != prod(dims)) return ffi::Error::InvalidArgument("pw_apply_vec: state size != prod(dims)");
  auto tg = to_i32(targets);
  return status(pw_apply_vec(psi.untyped_data(), out->untyped_data(), op.untyped_data(), dims.begin(),
                             (int)dims.size(), tg.data(), (int)tg.size(), dt, stream), "pw_apply_vec");
}

ffi::Error KrausMatImpl(cudaStream_t stream, ffi::AnyBuffer rho, ffi::AnyBuffer ops, I64s dims,
                        I64s targets, ffi::Result<ffi::AnyBuffer> out) {
  const int dt = dtype_of(rho);
  if (dt < 0 || dtype_of(ops) != dt) return ffi::Error::InvalidArgument("pw_kraus_mat: complex64/128 buffers of one dtype expected");
  const int64_t n = prod(dims);
  if ((int64_t)rho.element_count() != n * n) return ffi::Error::InvalidArgument("pw_kraus_mat: state size != prod(dims)^2");
  auto tg = to_i32(targets);
  const auto od = ops.dimensions();  // (K, t, t); a single (t, t) operator is K = 1
  const int K = od.size() == 3 ? (int)od[0] : 1;
  if (K == 1)
    return status(pw_apply_mat(rho.untyped_data(), out->untyped_data(), ops.untyped_data(), dims.begin(),
                               (int)dims.size(), tg.data(), (int)tg.size(), dt, stream), "pw_apply_mat");
  return status(pw_kraus_mat(rho.untyped_data(), out->untyped_data(), ops.untyped_data(), K, dims.begin(),
                             (int)dims.size(), tg.data(), (int)tg.size(), dt, stream), "pw_kraus_mat");
}

// ---- reductions -------------------------------------------------------------------------------

ffi::Error TraceOutMatImpl(cudaStream_t stream, ffi::AnyBuffer rho, I64s dims, I64s keep,
                           ffi::Result<ffi::AnyBuffer> out) {
  const int dt = dtype_of(rho);
  if (dt < 0) return ffi::Error::InvalidArgument("pw_trace_out_mat: complex buffer expected");
  auto kp = to_i32(keep);
  return status(pw_trace_out_mat(rho.untyped_data(), out->untyped_data(), dims.begin(), (int)dims.size(),
                                 kp.data(), (int)kp.size(), dt, stream), "pw_trace_out_mat");
}

ffi::Error ReducedFromVecImpl(cudaStream_t stream, ffi::AnyBuffer psi, I64s dims, I64s keep,
                              ffi::Result<ffi::AnyBuffer> out) {
  const int dt = dtype_of(psi);
  if (dt < 0) return ffi::Error::InvalidArgument("pw_reduced_from_vec: complex buffer expected");
  auto kp = to_i32(keep);
  return status(pw_reduced_from_vec(psi.untyped_data(), out->untyped_data(), dims.begin(), (int)dims.size(),
                                    kp.data(), (int)kp.size(), dt, stream), "pw_reduced_from_vec");
}

ffi::Error CrossReducedImpl(cudaStream_t stream, ffi::AnyBuffer x, ffi::AnyBuffer y, I64s axes, I64s keep,
                            ffi::Result<ffi::AnyBuffer> out) {
  const int dt = dtype_of(x);
  if (dt < 0 || dtype_of(y) != dt) return ffi::Error::InvalidArgument("pw_cross_reduced: complex buffers of one dtype expected");
  auto kp = to_i32(keep);
  return status(pw_cross_reduced(x.untyped_data(), y.untyped_data(), out->untyped_data(), axes.begin(),
                                 (int)axes.size(), kp.data(), (int)kp.size(), dt, stream), "pw_cross_reduced");
}

ffi::Error ProbsImpl(cudaStream_t stream, ffi::AnyBuffer state, I64s dims, I64s targets, int64_t is_matrix,
                     ffi::Result<ffi::AnyBuffer> probs) {
  const int dt = dtype_of(state);
  if (dt < 0) return ffi::Error::InvalidArgument("pw_probs: complex buffer expected");
  auto tg = to_i32(targets);
  auto fn = is_matrix ? pw_probs_mat : pw_probs_vec;
  return status(fn(state.untyped_data(), probs->untyped_data(), dims.begin(), (int)dims.size(), tg.data(),
                   (int)tg.size(), dt, stream), is_matrix ? "pw_probs_mat" : "pw_probs_vec");
}

// ---- sampling + collapse ----------------------------------------------------------------------

// key: uint32[2] on the device (jax.random.key_data of the traced key); outcome: int64 scalar
ffi::Error DrawImpl(cudaStream_t stream, ffi::AnyBuffer probs, ffi::AnyBuffer key,
                    ffi::Result<ffi::AnyBuffer> outcome) {
  int dt;
  switch (probs.element_type()) {
    case ffi::DataType::F32: dt = PW_C64; break;
    case ffi::DataType::F64: dt = PW_C128; break;
    default: return ffi::Error::InvalidArgument("pw_draw: float32/float64 probabilities expected");
  }
  if (key.element_type() != ffi::DataType::U32 || key.element_count() != 2)
    return ffi::Error::InvalidArgument("pw_draw: key must be uint32[2]");
  return status(pw_draw_devkey(probs.untyped_data(), (int64_t)probs.element_count(),
                               static_cast<const uint32_t*>(key.untyped_data()), dt,
                               static_cast<int64_t*>(outcome->untyped_data()), stream), "pw_draw");
}

ffi::Error CollapseImpl(cudaStream_t stream, ffi::AnyBuffer state, ffi::AnyBuffer outcome, ffi::AnyBuffer probs,
                        I64s dims, I64s targets, int64_t is_matrix, ffi::Result<ffi::AnyBuffer> post) {
  const int dt = dtype_of(state);
  if (dt < 0) return ffi::Error::InvalidArgument("pw_collapse: complex buffer expected");
  auto tg = to_i32(targets);
  auto fn = is_matrix ? pw_collapse_mat : pw_collapse_vec;
  return status(fn(state.untyped_data(), post->untyped_data(), static_cast<const int64_t*>(outcome.untyped_data()),
                   probs.untyped_data(), dims.begin(), (int)dims.size(), tg.data(), (int)tg.size(), dt, stream),
                is_matrix ? "pw_collapse_mat" : "pw_collapse_vec");
}

ffi::Error PovmProbsImpl(cudaStream_t stream, ffi::AnyBuffer rho_t, ffi::AnyBuffer ops,
                         ffi::Result<ffi::AnyBuffer> probs, ffi::Result<ffi::AnyBuffer> traces) {
  const int dt = dtype_of(rho_t);
  if (dt < 0 || dtype_of(ops) != dt) return ffi::Error::InvalidArgument("pw_povm_probs: complex buffers of one dtype expected");
  const auto od = ops.dimensions();
  if (od.size() != 3) return ffi::Error::InvalidArgument("pw_povm_probs: operators must be (K, t, t)");
  return status(pw_povm_probs(rho_t.untyped_data(), ops.untyped_data(), (int)od[0], od[1], probs->untyped_data(),
                              traces->untyped_data(), dt, stream), "pw_povm_probs");
}

ffi::Error PovmSelectImpl(cudaStream_t stream, ffi::AnyBuffer ops, ffi::AnyBuffer outcome, ffi::AnyBuffer traces,
                          ffi::Result<ffi::AnyBuffer> op_out) {
  const int dt = dtype_of(ops);
  if (dt < 0) return ffi::Error::InvalidArgument("pw_povm_select: complex buffer expected");
  const auto od = ops.dimensions();
  if (od.size() != 3) return ffi::Error::InvalidArgument("pw_povm_select: operators must be (K, t, t)");
  return status(pw_povm_select(ops.untyped_data(), od[1], static_cast<const int64_t*>(outcome.untyped_data()),
                               traces.untyped_data(), op_out->untyped_data(), dt, stream), "pw_povm_select");
}

}  // namespace

#define PW_STREAM .Ctx<ffi::PlatformStream<cudaStream_t>>()
#define PW_BUF .Arg<ffi::AnyBuffer>()
#define PW_OUT .Ret<ffi::AnyBuffer>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(PwApplyVec, ApplyVecImpl,
                              ffi::Ffi::Bind() PW_STREAM PW_BUF PW_BUF.Attr<I64s>("dims").Attr<I64s>("targets") PW_OUT);
XLA_FFI_DEFINE_HANDLER_SYMBOL(PwKrausMat, KrausMatImpl,
                              ffi::Ffi::Bind() PW_STREAM PW_BUF PW_BUF.Attr<I64s>("dims").Attr<I64s>("targets") PW_OUT);
XLA_FFI_DEFINE_HANDLER_SYMBOL(PwTraceOutMat, TraceOutMatImpl,
                              ffi::Ffi::Bind() PW_STREAM PW_BUF.Attr<I64s>("dims").Attr<I64s>("keep") PW_OUT);
XLA_FFI_DEFINE_HANDLER_SYMBOL(PwReducedFromVec, ReducedFromVecImpl,
                              ffi::Ffi::Bind() PW_STREAM PW_BUF.Attr<I64s>("dims").Attr<I64s>("keep") PW_OUT);
XLA_FFI_DEFINE_HANDLER_SYMBOL(PwCrossReduced, CrossReducedImpl,
                              ffi::Ffi::Bind() PW_STREAM PW_BUF PW_BUF.Attr<I64s>("axes").Attr<I64s>("keep") PW_OUT);
XLA_FFI_DEFINE_HANDLER_SYMBOL(PwProbs, ProbsImpl,
                              ffi::Ffi::Bind() PW_STREAM PW_BUF.Attr<I64s>("dims").Attr<I64s>("targets")
                                  .Attr<int64_t>("is_matrix") PW_OUT);
XLA_FFI_DEFINE_HANDLER_SYMBOL(PwDraw, DrawImpl, ffi::Ffi::Bind() PW_STREAM PW_BUF PW_BUF PW_OUT);
XLA_FFI_DEFINE_HANDLER_SYMBOL(PwCollapse, CollapseImpl,
                              ffi::Ffi::Bind() PW_STREAM PW_BUF PW_BUF PW_BUF.Attr<I64s>("dims").Attr<I64s>("targets")
                                  .Attr<int64_t>("is_matrix") PW_OUT);
XLA_FFI_DEFINE_HANDLER_SYMBOL(PwPovmProbs, PovmProbsImpl, ffi::Ffi::Bind() PW_STREAM PW_BUF PW_BUF PW_OUT PW_OUT);
XLA_FFI_DEFINE_HANDLER_SYMBOL(PwPovmSelect, PovmSelectImpl, ffi::Ffi::Bind() PW_STREAM PW_BUF PW_BUF PW_BUF PW_OUT);
