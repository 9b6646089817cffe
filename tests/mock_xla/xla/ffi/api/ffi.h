// MOCK of the binding surface of xla/ffi/api/ffi.h that ffi/sde4_xla_ffi.cc uses -- TEST INFRASTRUCTURE.
// It exists so that the shim's `__has_include` guard and the SIGNATURES of its handlers can be compiled in an image
// without jaxlib: XLA_FFI_DEFINE_HANDLER_SYMBOL static-asserts that the implementation is invocable with exactly the
// C++ types the binding declares (context, attributes, arguments, results, in order), as the real decoder would pass.
// Nothing here executes a handler; the exported symbol returns a null error.
#ifndef MOCK_XLA_FFI_API_FFI_H_
#define MOCK_XLA_FFI_API_FFI_H_
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <type_traits>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla {
namespace ffi {

enum DataType { F32, U32, S32, U8 };
template <DataType>
struct NativeType;
template <>
struct NativeType<F32> { using type = float; };
template <>
struct NativeType<U32> { using type = uint32_t; };
template <>
struct NativeType<S32> { using type = int32_t; };
template <>
struct NativeType<U8> { using type = uint8_t; };

template <class T>
class Span {
 public:
  Span() : p_(nullptr), n_(0) {}
  Span(const T* p, size_t n) : p_(p), n_(n) {}
  const T* data() const { return p_; }
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return p_[i]; }
  const T* begin() const { return p_; }
  const T* end() const { return p_ + n_; }

 private:
  const T* p_;
  size_t n_;
};

template <DataType dt>
class Buffer {
 public:
  using T = typename NativeType<dt>::type;
  T* typed_data() const { return data_; }
  void* untyped_data() const { return data_; }
  Span<const int64_t> dimensions() const { return Span<const int64_t>(dims_, rank_); }
  size_t element_count() const {
    size_t n = 1;
    for (size_t i = 0; i < rank_; ++i) n *= (size_t)dims_[i];
    return n;
  }

 private:
  T* data_ = nullptr;
  const int64_t* dims_ = nullptr;
  size_t rank_ = 0;
};

template <class T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};
template <DataType dt>
using ResultBuffer = Result<Buffer<dt>>;

class Error {
 public:
  static Error Success() { return Error(true, ""); }
  static Error Internal(std::string m) { return Error(false, std::move(m)); }
  static Error InvalidArgument(std::string m) { return Error(false, std::move(m)); }
  bool success() const { return ok_; }
  const std::string& message() const { return msg_; }

 private:
  Error(bool ok, std::string m) : ok_(ok), msg_(std::move(m)) {}
  bool ok_;
  std::string msg_;
};

template <class S>
struct PlatformStream {};

// ---- binding: a type list of the C++ parameters the decoder will pass to the implementation ----
template <class... Ps>
struct Binding {
  template <class C>
  struct CtxParam;
  template <class S>
  struct CtxParam<PlatformStream<S>> { using type = S; };
  template <class C>
  Binding<Ps..., typename CtxParam<C>::type> Ctx() const { return {}; }
  template <class T>
  Binding<Ps..., T> Attr(const char*) const { return {}; }
  template <class T>
  Binding<Ps..., T> Arg() const { return {}; }
  template <class T>
  Binding<Ps..., Result<T>> Ret() const { return {}; }
  template <class F>
  static constexpr bool Accepts() { return std::is_invocable_r<Error, F, Ps...>::value; }
};
struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                                                        \
  static_assert(decltype(binding)::template Accepts<decltype(&impl)>(),                                              \
                #impl " is not callable with the parameter list its binding declares");                             \
  extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame*) { return nullptr; }

#endif  // MOCK_XLA_FFI_API_FFI_H_
