// TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the part of <xla/ffi/api/ffi.h> (shipped in
// jaxlib, absent from this image) that jax_finufft_b200/csrc/xla_ffi_shim.cc uses, so that the
// shim is parsed, its 18 handlers are instantiated and their attribute decoding is executed by
// tests/test_xla_shim_mock.py.  It mimics the typed-FFI binder's interface (Ffi::Bind().Ctx<>()
// .Attr<T>(name)...RemainingArgs().Ret<>().To(fn)), not XLA's implementation: a call frame here
// is a plain struct the test driver fills.  The real build recipe is INTEGRATION.md section 2.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <tuple>
#include <utility>
#include <variant>
#include <vector>

struct XLA_FFI_Error {
  int code;  // 0 ok, 3 invalid argument, 13 internal (absl status codes, as XLA uses)
  std::string message;
};
struct XLA_FFI_CallFrame {
  void *stream = nullptr;
  std::map<std::string, std::variant<float, double, int64_t>> attrs;
  std::vector<void *> args;
  void *ret = nullptr;
  std::vector<std::string> decoded;  // attribute names in the order the handler asked for them
};
using XLA_FFI_Handler = XLA_FFI_Error *(XLA_FFI_CallFrame *);

namespace xla::ffi {

class Error {
 public:
  static Error Success() { return Error(0, ""); }
  static Error InvalidArgument(std::string m) { return Error(3, std::move(m)); }
  static Error Internal(std::string m) { return Error(13, std::move(m)); }
  bool success() const { return code_ == 0; }
  int code() const { return code_; }
  const std::string &message() const { return msg_; }

 private:
  Error(int c, std::string m) : code_(c), msg_(std::move(m)) {}
  int code_;
  std::string msg_;
};

class AnyBuffer {
 public:
  explicit AnyBuffer(void *p = nullptr) : p_(p) {}
  void *untyped_data() const { return p_; }

 private:
  void *p_;
};
template <typename T> class Result {
 public:
  explicit Result(T v) : v_(v) {}
  T *operator->() { return &v_; }

 private:
  T v_;
};
class RemainingArgs {
 public:
  explicit RemainingArgs(std::vector<void *> a) : a_(std::move(a)) {}
  size_t size() const { return a_.size(); }
  template <typename T> std::optional<T> get(size_t i) const {
    if (i >= a_.size()) return std::nullopt;
    return T(a_[i]);
  }

 private:
  std::vector<void *> a_;
};
template <typename S> struct PlatformStream {};

namespace detail {
template <typename T> struct AttrTag { std::string name; };
template <typename S> struct CtxTag {};
struct RemainingTag {};
template <typename T> struct RetTag {};

template <typename T> struct Decode;
template <typename T> struct Decode<AttrTag<T>> {
  using type = T;
  static bool get(XLA_FFI_CallFrame *f, const AttrTag<T> &t, T *out, std::string *err) {
    f->decoded.push_back(t.name);
    auto it = f->attrs.find(t.name);
    if (it == f->attrs.end()) { *err = "missing attribute " + t.name; return false; }
    if (!std::holds_alternative<T>(it->second)) { *err = "attribute " + t.name + " has the wrong type"; return false; }
    *out = std::get<T>(it->second);
    return true;
  }
};
template <typename S> struct Decode<CtxTag<PlatformStream<S>>> {
  using type = S;
  static bool get(XLA_FFI_CallFrame *f, const CtxTag<PlatformStream<S>> &, S *out, std::string *) {
    *out = reinterpret_cast<S>(f->stream);
    return true;
  }
};
template <> struct Decode<RemainingTag> {
  using type = RemainingArgs;
};
template <typename T> struct Decode<RetTag<T>> {
  using type = Result<T>;
};
}  // namespace detail

template <typename Fn, typename... Tags> class Handler {
 public:
  Handler(std::tuple<Tags...> tags, Fn fn) : tags_(std::move(tags)), fn_(fn) {}
  XLA_FFI_Error *Call(XLA_FFI_CallFrame *frame) {
    std::string err;
    Error e = Invoke(frame, &err, std::index_sequence_for<Tags...>{});
    if (e.success()) return nullptr;
    return new XLA_FFI_Error{e.code(), e.message()};
  }

 private:
  template <typename Tag> auto One(XLA_FFI_CallFrame *f, const Tag &t, std::string *err, bool *ok) {
    if constexpr (std::is_same_v<Tag, detail::RemainingTag>) {
      return RemainingArgs(f->args);
    } else if constexpr (std::is_same_v<Tag, detail::RetTag<AnyBuffer>>) {
      return Result<AnyBuffer>(AnyBuffer(f->ret));
    } else {
      typename detail::Decode<Tag>::type v{};
      if (!detail::Decode<Tag>::get(f, t, &v, err)) *ok = false;
      return v;
    }
  }
  template <size_t... I> Error Invoke(XLA_FFI_CallFrame *f, std::string *err, std::index_sequence<I...>) {
    bool ok = true;
    // braced init: evaluated left to right, i.e. in binding order
    std::tuple<decltype(One(f, std::get<I>(tags_), err, &ok))...> vals{One(f, std::get<I>(tags_), err, &ok)...};
    if (!ok) return Error::InvalidArgument(*err);
    return std::apply(fn_, std::move(vals));
  }
  std::tuple<Tags...> tags_;
  Fn fn_;
};

template <typename... Tags> class Binding {
 public:
  explicit Binding(std::tuple<Tags...> t = {}) : tags_(std::move(t)) {}
  template <typename C> auto Ctx() && { return Add(detail::CtxTag<C>{}); }
  template <typename T> auto Attr(std::string name) && { return Add(detail::AttrTag<T>{std::move(name)}); }
  auto RemainingArgs() && { return Add(detail::RemainingTag{}); }
  template <typename T> auto Ret() && { return Add(detail::RetTag<T>{}); }
  template <typename Fn> auto To(Fn fn) && {
    return std::make_unique<Handler<Fn, Tags...>>(std::move(tags_), fn);
  }

 private:
  template <typename Tag> auto Add(Tag t) {
    return Binding<Tags..., Tag>(std::tuple_cat(std::move(tags_), std::make_tuple(std::move(t))));
  }
  std::tuple<Tags...> tags_;
};

struct Ffi {
  static Binding<> Bind() { return Binding<>(); }
};

}  // namespace xla::ffi
