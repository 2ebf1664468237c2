// TEST INFRASTRUCTURE ONLY -- a compile-time stand-in for jaxlib's xla/ffi/api/ffi.h (absent from
// the build image) that declares the small part of the typed-FFI surface pw_xla_ffi.cc uses, as
// documented for jaxlib 0.4.31+ / 0.8: Error, DataType, Span, AnyBuffer, Result, PlatformStream,
// Ffi::Bind().Ctx/Arg/Attr/Ret and XLA_FFI_DEFINE_HANDLER_SYMBOL.  It lets
// tests/test_xla_ffi_source.py type-check the shim's handler signatures against its bindings
// (argument order, attribute types, result count); it does not and cannot run anything.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>

struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;

namespace xla::ffi {

enum class DataType { F32, F64, C64, C128, S64, U32 };

class Error {
 public:
  static Error Success() { return Error(); }
  static Error InvalidArgument(std::string m) { return Error(std::move(m)); }
  static Error Internal(std::string m) { return Error(std::move(m)); }
  bool success() const { return msg_.empty(); }

 private:
  Error() = default;
  explicit Error(std::string m) : msg_(std::move(m)) {}
  std::string msg_;
};

template <typename T>
class Span {
 public:
  Span(const std::remove_const_t<T>* p = nullptr, size_t n = 0) : p_(p), n_(n) {}
  const std::remove_const_t<T>* begin() const { return p_; }
  const std::remove_const_t<T>* end() const { return p_ + n_; }
  size_t size() const { return n_; }
  const std::remove_const_t<T>& operator[](size_t i) const { return p_[i]; }

 private:
  const std::remove_const_t<T>* p_;
  size_t n_;
};

class AnyBuffer {
 public:
  DataType element_type() const { return DataType::C128; }
  Span<const int64_t> dimensions() const { return {}; }
  size_t element_count() const { return 0; }
  void* untyped_data() const { return nullptr; }
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }

 private:
  T v_;
};

template <typename S>
struct PlatformStream {};

namespace internal {
template <typename T> struct ArgTag { using type = T; };
template <typename T> struct RetTag { using type = Result<T>; };
template <typename T> struct AttrTag { using type = T; };
template <typename T> struct CtxTag;
template <typename S> struct CtxTag<PlatformStream<S>> { using type = S; };

template <typename... Tags>
struct Binding {
  template <typename T> Binding<Tags..., CtxTag<T>> Ctx() const { return {}; }
  template <typename T> Binding<Tags..., ArgTag<T>> Arg() const { return {}; }
  template <typename T> Binding<Tags..., RetTag<T>> Ret() const { return {}; }
  template <typename T> Binding<Tags..., AttrTag<T>> Attr(const char*) const { return {}; }
  // the handler must be callable with exactly the bound parameter list
  template <typename Fn>
  static constexpr bool Matches() { return std::is_invocable_r_v<Error, Fn, typename Tags::type...>; }
};
}  // namespace internal

struct Ffi {
  static internal::Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                          \
  static_assert(decltype(binding)::template Matches<decltype(&impl)>(),                              \
                #name ": handler signature does not match its binding");                             \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*) { return nullptr; }
