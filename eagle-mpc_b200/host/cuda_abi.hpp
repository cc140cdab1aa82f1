// cuda_abi.hpp — the host-side mirror's only door to the CUDA path: the C ABI of include/empc_b200.h, bound at run time.
//
// libempc_host.so (YAML / URDF / Trajectory / MPC controllers: no compute) does not link against libempc_b200.so (CUDA
// kernels + C ABI).  SolverSbFDDP binds the entry points it needs with dlopen/dlsym the first time a solver is
// constructed, so problem construction works on machines without a GPU (and without loading any CUDA code), while
// anything that solves fails loudly when the CUDA library is missing.  There is no CPU fallback.
#pragma once
#include "../../include/empc_b200.h"

namespace eagle_mpc {

struct CudaAbi {
  decltype(&empc_last_error) last_error;
  decltype(&empc_default_params) default_params;
  decltype(&empc_box_params) box_params;
  decltype(&empc_create) create;
  decltype(&empc_destroy) destroy;
  decltype(&empc_set_x0) set_x0;
  decltype(&empc_set_candidate) set_candidate;
  decltype(&empc_set_params) set_params;
  decltype(&empc_update_costs) update_costs;
  decltype(&empc_solve) solve;
  decltype(&empc_get_solution) get_solution;
  decltype(&empc_get_K) get_K;
  decltype(&empc_get_k) get_k;
  decltype(&empc_enable_iteration_log) enable_iteration_log;
  decltype(&empc_get_iteration_log) get_iteration_log;
};

// Throws std::runtime_error when libempc_b200.so cannot be loaded ($EMPC_LIB, else next to libempc_host.so).
const CudaAbi& cuda_abi();

}  // namespace eagle_mpc
