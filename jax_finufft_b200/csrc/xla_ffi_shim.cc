// xla_ffi_shim.cc -- the Python extension module `jax_finufft.jax_finufft_gpu` rebuilt on
// libb200nufft.so.  Drop-in for lib/jax_finufft_gpu.cc + lib/kernels.cc.cu +
// lib/cufinufft_wrapper.{h,cc}: exports `registrations() -> dict[str, PyCapsule]` with the same
// 18 custom-call targets, the same typed attributes and the same operand order, so
// src/jax_finufft/lowering.py:15-21,176-178 registers and lowers them unchanged.
//
// NOT built into the library in this repository's image: it needs <xla/ffi/api/ffi.h> (shipped
// inside jaxlib) and nanobind, neither of which is installed here.  Everything
// with behaviour lives below the C ABI (ffi_core.cpp: b2n_ffi_call) and is exercised by
// tests/test_ffi_core.py; this file only adapts XLA's call frame to that one function.
// tests/test_xla_shim_mock.py compiles it against a mock of the xla::ffi / nanobind surface it
// uses and runs the 18 handlers (schema, eps width, operand count, error forwarding) -- JAX
// itself has never driven it.  Build recipe: INTEGRATION.md §2.
//
// Shape of the adapter: instead of one hand-written wrapper per (dim, type, precision) as
// upstream, ONE handler template parameterised by the target index; operands arrive through
// ffi::RemainingArgs, so arity (1 + dim, or 1 + 2 dim for type 3) is checked at run time against
// b2n_ffi_arity().
#if defined(B2N_BUILD_XLA_SHIM)

#include <cuda_runtime_api.h>
#include <nanobind/nanobind.h>
#include <xla/ffi/api/ffi.h>

#include <cstdint>
#include <string>

#include "../../include/b200nufft.h"

namespace ffi = xla::ffi;
namespace nb = nanobind;

namespace {

template <int I, typename Eps>
ffi::Error call(cudaStream_t stream, Eps eps, int64_t iflag, int64_t n_tot, int64_t n_transf, int64_t n_j,
                int64_t n_k_1, int64_t n_k_2, int64_t n_k_3, int64_t modeord, double upsampfac,
                int64_t gpu_method, int64_t gpu_sort, int64_t gpu_kerevalmeth, int64_t gpu_maxbatchsize,
                int64_t debug, ffi::RemainingArgs args, ffi::Result<ffi::AnyBuffer> out) {
  const char *name = b2n_ffi_targets()[I];
  const int arity = b2n_ffi_arity(name);
  if ((int)args.size() != arity)
    return ffi::Error::InvalidArgument(std::string(name) + ": expected " + std::to_string(arity) + " operands");
  const void *ops[7];
  for (int i = 0; i < arity; i++) {
    auto b = args.get<ffi::AnyBuffer>(i);
    if (!b.has_value()) return ffi::Error::InvalidArgument(std::string(name) + ": operand is not a buffer");
    ops[i] = b->untyped_data();
  }
  const b2n_ffi_attrs a = {(double)eps, iflag,     n_tot,      n_transf, n_j,
                           n_k_1,       n_k_2,     n_k_3,      modeord,  upsampfac,
                           gpu_method,  gpu_sort,  gpu_kerevalmeth, gpu_maxbatchsize, debug};
  const int rc = b2n_ffi_call(name, (void *)stream, &a, ops, arity, out->untyped_data());
  if (rc > 1) return ffi::Error::Internal(std::string(b2n_strerror(rc)) + " (code " + std::to_string(rc) + ")");
  // The reference blocks until the stream drains (lib/kernels.cc.cu:83) because it destroys its
  // plan; ours is cached and stream-ordered, so the call returns as soon as the work is enqueued.
  return ffi::Error::Success();
}

template <int I, typename Eps> XLA_FFI_Error *handler(XLA_FFI_CallFrame *frame) {
  // `Eps` makes every type after .Attr<Eps> dependent: the member templates need `template`
  static auto *h = ffi::Ffi::Bind()
                       .Ctx<ffi::PlatformStream<cudaStream_t>>()
                       .template Attr<Eps>("eps")
                       .template Attr<int64_t>("iflag")
                       .template Attr<int64_t>("n_tot")
                       .template Attr<int64_t>("n_transf")
                       .template Attr<int64_t>("n_j")
                       .template Attr<int64_t>("n_k_1")
                       .template Attr<int64_t>("n_k_2")
                       .template Attr<int64_t>("n_k_3")
                       .template Attr<int64_t>("modeord")
                       .template Attr<double>("upsampfac")
                       .template Attr<int64_t>("gpu_method")
                       .template Attr<int64_t>("gpu_sort")
                       .template Attr<int64_t>("gpu_kerevalmeth")
                       .template Attr<int64_t>("gpu_maxbatchsize")
                       .template Attr<int64_t>("debug")
                       .RemainingArgs()
                       .template Ret<ffi::AnyBuffer>()
                       .To(call<I, Eps>)
                       .release();
  return h->Call(frame);
}

// b2n_ffi_targets() alternates single ("...f", float eps attribute) and double precision
template <int I> void add(nb::dict &d) {
  if constexpr (I < 18) {
    XLA_FFI_Handler *fn = (I % 2 == 0) ? &handler<I, float> : &handler<I, double>;
    d[b2n_ffi_targets()[I]] = nb::capsule(reinterpret_cast<void *>(fn));
    add<I + 1>(d);
  }
}

nb::dict registrations() {
  nb::dict d;
  add<0>(d);
  return d;
}

}  // namespace

NB_MODULE(jax_finufft_gpu, m) { m.def("registrations", &registrations); }

#endif  // B2N_BUILD_XLA_SHIM
