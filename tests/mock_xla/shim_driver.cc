// TEST INFRASTRUCTURE ONLY -- drives jax_finufft_b200/csrc/xla_ffi_shim.cc (compiled against the
// mock headers in include/) on a machine without a GPU: registrations() must hold the 18 targets,
// every handler must decode the reference's attribute schema in the reference's order
// (lib/jax_finufft_gpu.cc:28-60) with a float `eps` for the ...f targets and a double one
// otherwise, check the operand count, and turn library error codes into ffi::Error::Internal.
// Prints one line per check; exit code 0 = all passed.
#include <nanobind/nanobind.h>
#include <xla/ffi/api/ffi.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200nufft.h"

void mock_nb_module_jax_finufft_gpu(nanobind::module_ &m);

static int failures = 0;
#define CHECK(cond, ...)                                  \
  do {                                                    \
    if (!(cond)) {                                        \
      failures++;                                         \
      std::printf("FAIL %s:%d: ", __FILE__, __LINE__);    \
      std::printf(__VA_ARGS__);                           \
      std::printf("\n");                                  \
    }                                                     \
  } while (0)

static XLA_FFI_CallFrame frame(bool single, int n_ops, long long n_transf) {
  static double buf[64];
  XLA_FFI_CallFrame f;
  if (single) f.attrs["eps"] = 1e-6f; else f.attrs["eps"] = 1e-6;
  for (const char *k : {"iflag", "n_tot", "n_j", "n_k_1", "n_k_2", "n_k_3", "modeord", "gpu_method", "gpu_sort",
                        "gpu_kerevalmeth", "gpu_maxbatchsize", "debug"})
    f.attrs[k] = int64_t(1);
  f.attrs["n_transf"] = int64_t(n_transf);
  f.attrs["upsampfac"] = 2.0;
  f.args.assign(n_ops, buf);
  f.ret = buf;
  return f;
}

int main() {
  nanobind::module_ m;
  mock_nb_module_jax_finufft_gpu(m);
  CHECK(m.registrations != nullptr, "module defines no registrations()");
  nanobind::dict d = m.registrations();
  CHECK(d.items.size() == 18, "expected 18 targets, got %zu", d.items.size());
  const std::vector<std::string> schema = {"eps", "iflag", "n_tot", "n_transf", "n_j", "n_k_1", "n_k_2", "n_k_3", "modeord",
                                           "upsampfac", "gpu_method", "gpu_sort", "gpu_kerevalmeth", "gpu_maxbatchsize", "debug"};
  const char *const *names = b2n_ffi_targets();
  for (int i = 0; names[i]; i++) {
    const std::string name = names[i];
    auto it = d.items.find(name);
    CHECK(it != d.items.end() && it->second.ptr, "%s not registered", name.c_str());
    if (it == d.items.end()) continue;
    auto *h = reinterpret_cast<XLA_FFI_Handler *>(it->second.ptr);
    const bool single = name.back() == 'f';
    const int arity = b2n_ffi_arity(name.c_str());
    // (1) wrong operand count -> InvalidArgument naming the target
    XLA_FFI_CallFrame f1 = frame(single, arity + 1, 1);
    XLA_FFI_Error *e = h(&f1);
    CHECK(e && e->code == 3 && e->message.find(name) != std::string::npos, "%s: operand count not checked", name.c_str());
    CHECK(f1.decoded == schema, "%s: attribute schema/order differs from lib/jax_finufft_gpu.cc:28-60", name.c_str());
    delete e;
    // (2) eps of the other width must be refused by the typed decoder
    XLA_FFI_CallFrame f2 = frame(!single, arity, 1);
    e = h(&f2);
    CHECK(e && e->code == 3 && e->message.find("eps") != std::string::npos, "%s: eps width not enforced", name.c_str());
    delete e;
    // (3) a library error (n_transf = 0 -> code 9, rejected before any device work) -> Internal
    XLA_FFI_CallFrame f3 = frame(single, arity, 0);
    e = h(&f3);
    CHECK(e && e->code == 13 && e->message.find("code 9") != std::string::npos, "%s: library error not forwarded (%s)",
          name.c_str(), e ? e->message.c_str() : "no error");
    delete e;
    std::printf("ok %s (arity %d, eps %s)\n", name.c_str(), arity, single ? "f32" : "f64");
  }
  std::printf("%s\n", failures ? "FAILED" : "ALL OK");
  return failures ? 1 : 0;
}
