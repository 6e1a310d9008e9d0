// TEST INFRASTRUCTURE ONLY -- stand-in for the three nanobind names xla_ffi_shim.cc uses
// (nb::dict, nb::capsule, NB_MODULE / module_::def); see ../xla/ffi/api/ffi.h.
#pragma once
#include <map>
#include <string>

namespace nanobind {
struct capsule {
  void *ptr = nullptr;
  capsule() = default;
  explicit capsule(void *p) : ptr(p) {}
};
struct dict {
  std::map<std::string, capsule> items;
  capsule &operator[](const char *k) { return items[k]; }
};
struct module_ {
  dict (*registrations)() = nullptr;
  void def(const char *, dict (*fn)()) { registrations = fn; }
};
}  // namespace nanobind
#define NB_MODULE(name, m)                               \
  void mock_nb_module_##name(nanobind::module_ &m);      \
  void mock_nb_module_##name(nanobind::module_ &m)
