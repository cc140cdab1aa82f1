// cuda_abi.cpp — run-time binding of the C ABI (see cuda_abi.hpp).
#include "cuda_abi.hpp"

#include <dlfcn.h>

#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>

namespace eagle_mpc {

static std::string default_lib_path() {
  if (const char* e = std::getenv("EMPC_LIB")) return e;
  Dl_info info;
  if (dladdr((const void*)&default_lib_path, &info) && info.dli_fname) {
    const std::string self(info.dli_fname);
    const std::size_t slash = self.find_last_of('/');
    return (slash == std::string::npos ? std::string(".") : self.substr(0, slash)) + "/libempc_b200.so";
  }
  return "libempc_b200.so";
}

template <class F>
static void bind(void* lib, const char* name, F& out) {
  void* p = dlsym(lib, name);
  if (!p) throw std::runtime_error(std::string("libempc_b200.so does not export ") + name);
  out = reinterpret_cast<F>(p);
}

const CudaAbi& cuda_abi() {
  static CudaAbi abi;
  static std::once_flag once;
  static std::string error;
  std::call_once(once, [] {
    const std::string path = default_lib_path();
    void* lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!lib) { error = "cannot load the CUDA path (" + path + "): " + dlerror() + " -- there is no CPU fallback"; return; }
    try {
      bind(lib, "empc_last_error", abi.last_error);
      bind(lib, "empc_default_params", abi.default_params);
      try { bind(lib, "empc_box_params", abi.box_params); }   // (a CUDA library older than the Box solvers: only they are refused)
      catch (const std::exception&) { abi.box_params = nullptr; }
      bind(lib, "empc_create", abi.create);
      bind(lib, "empc_destroy", abi.destroy);
      bind(lib, "empc_set_x0", abi.set_x0);
      bind(lib, "empc_set_candidate", abi.set_candidate);
      bind(lib, "empc_set_params", abi.set_params);
      bind(lib, "empc_update_costs", abi.update_costs);
      bind(lib, "empc_solve", abi.solve);
      bind(lib, "empc_get_solution", abi.get_solution);
      bind(lib, "empc_get_K", abi.get_K);
      bind(lib, "empc_get_k", abi.get_k);
      bind(lib, "empc_enable_iteration_log", abi.enable_iteration_log);
      bind(lib, "empc_get_iteration_log", abi.get_iteration_log);
    } catch (const std::exception& e) { error = e.what(); }
  });
  if (!error.empty()) throw std::runtime_error(error);
  return abi;
}

}  // namespace eagle_mpc
