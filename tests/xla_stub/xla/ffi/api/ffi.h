// MINIMAL STAND-IN for jaxlib's `xla/ffi/api/ffi.h` -- TEST INFRASTRUCTURE ONLY.
//
// jax / jaxlib are not installable in this image, so the real header (jax.ffi.include_dir()) is absent.  This stub declares just
// the part of the public XLA FFI C++ API that gpjax_b200/csrc/xla_ffi_shim.cc uses, with the same names and shapes, so that the
// shim is type-checked in CI (tests/test_abi_symbols.py::test_xla_ffi_shim_compiles_against_the_stub): every handler's parameter
// list is checked against its Ffi::Bind() chain by a static_assert, exactly the check the real binder performs.  It executes
// nothing.  Restated from the published API (xla/ffi/api/ffi.h, XLA FFI "external" API v1): Buffer<dtype>, Result<T>, Error /
// ErrorCode, PlatformStream<T>, ScratchAllocator, Ffi::Bind().Ctx/Arg/Ret/Attr, XLA_FFI_DEFINE_HANDLER_SYMBOL.
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <type_traits>
#include <utility>

namespace xla::ffi {

enum DataType { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
template <DataType dt> struct NativeTypeOf;
template <> struct NativeTypeOf<F64> { using type = double; };
template <> struct NativeTypeOf<F32> { using type = float; };
template <> struct NativeTypeOf<S32> { using type = int32_t; };
template <> struct NativeTypeOf<S64> { using type = int64_t; };
template <> struct NativeTypeOf<S8> { using type = int8_t; };
template <> struct NativeTypeOf<U8> { using type = uint8_t; };

template <typename T>
class Span {
public:
    Span(const T* d, size_t n) : d_(d), n_(n) {}
    const T& operator[](size_t i) const { return d_[i]; }
    size_t size() const { return n_; }
    const T* begin() const { return d_; }
    const T* end() const { return d_ + n_; }

private:
    const T* d_;
    size_t n_;
};

template <DataType dt>
class Buffer {
public:
    using T = typename NativeTypeOf<dt>::type;
    Span<int64_t> dimensions() const { return Span<int64_t>(dims_, rank_); }
    size_t element_count() const {
        size_t n = 1;
        for (size_t i = 0; i < rank_; ++i) n *= (size_t)dims_[i];
        return n;
    }
    size_t size_bytes() const { return element_count() * sizeof(T); }
    T* typed_data() const { return data_; }
    void* untyped_data() const { return data_; }

private:
    T* data_ = nullptr;
    const int64_t* dims_ = nullptr;
    size_t rank_ = 0;
};

template <typename T>
class Result {
public:
    T* operator->() { return &v_; }
    T& operator*() { return v_; }

private:
    T v_;
};

enum class ErrorCode { kOk, kCancelled, kUnknown, kInvalidArgument, kNotFound, kUnimplemented, kInternal, kResourceExhausted };
class Error {
public:
    Error() = default;
    Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
    static Error Success() { return Error(); }
    static Error InvalidArgument(std::string m) { return Error(ErrorCode::kInvalidArgument, std::move(m)); }
    static Error Internal(std::string m) { return Error(ErrorCode::kInternal, std::move(m)); }
    bool success() const { return code_ == ErrorCode::kOk; }

private:
    ErrorCode code_ = ErrorCode::kOk;
    std::string message_;
};

template <typename T> struct PlatformStream {};
class ScratchAllocator {
public:
    std::optional<void*> Allocate(size_t size, size_t alignment = 1) { (void)size; (void)alignment; return std::nullopt; }
};

namespace internal {
template <typename T> struct CtxParam { using type = T; };
template <typename T> struct CtxParam<PlatformStream<T>> { using type = T; };
template <typename Fn, typename... Ps>
struct Handler {
    Fn fn;
};
}  // namespace internal

template <typename... Ps>
class Binding {
public:
    template <typename T> Binding<Ps..., typename internal::CtxParam<T>::type> Ctx() const { return {}; }
    template <typename T> Binding<Ps..., T> Arg() const { return {}; }
    template <typename T> Binding<Ps..., Result<T>> Ret() const { return {}; }
    template <typename T> Binding<Ps..., T> Attr(const char*) const { return {}; }
    template <typename Fn>
    internal::Handler<Fn, Ps...> To(Fn fn) const {
        static_assert(std::is_invocable_r_v<Error, Fn, Ps...>,
                      "XLA FFI: the handler's parameter list does not match its Ffi::Bind() chain");
        return {fn};
    }
};
struct Ffi {
    static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

// the real macro defines `extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*)`; the stub keeps the symbol and the signature check
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)  \
    extern "C" const void* name() {                          \
        static auto handler = (binding).To(impl);            \
        return &handler;                                     \
    }
