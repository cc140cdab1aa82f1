// solver.cu — C-ABI implementation (include/empc_b200.h): device memory ownership, kernel launches, batch iteration loop.
//
// Host-side control flow mirrors SolverSbFDDP::solve (src/sbfddp.cpp:192-226) for a whole batch at once: every OCP
// carries its own state machine on the device (kernels.cuh: OcpState), the host only loops
//   calc_diff -> backward -> rollout -> decide
// until no OCP is active.  No CPU fallback: every entry point fails with EMPC_ERR_CUDA if the device is unusable.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "retarget.cuh"

using namespace empc;

static thread_local std::string g_last_error;
static int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return fail(EMPC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

struct empc_solver {
  int device = 0, B = 0, T = 0;
  int na = 0, nr = 0, nq = 0, nv = 0, nx = 0, ndx = 0, nu = 0, tile = 0;
  int n_costs = 0, n_pool = 0, n_node_maps = 0, n_costsets = 0;
  empc_solver_params_t P;
  DevModel hmodel;
  Buffers bf;
  std::vector<void*> allocs;
  // device copies that need host-side access
  DevModel* d_model = nullptr;
  empc_cost_t* d_costs = nullptr;
  double* d_pool = nullptr;
  int* d_costset_begin = nullptr;
  int* d_node_costset = nullptr;
  int* d_ocp_map = nullptr;
  double* d_x0 = nullptr;
  double *d_xs_init = nullptr, *d_us_init = nullptr;
  // batched MPC instances (retarget.cuh)
  double* d_ref_table = nullptr;  // (n_ref + 2) x nx: reference trajectory + rail hover row + carrot tail row
  size_t cap_ref_table = 0;
  long long* d_t_stages = nullptr; unsigned char* d_is_transition = nullptr; int n_stages = 0;  // carrot schedule
  size_t cap_t_stages = 0, cap_is_transition = 0;
  // weighted schedule storage (the const views the kernels see are in wsched)
  long long *d_ws_t_ini = nullptr, *d_ws_t_end = nullptr; unsigned char *d_ws_match = nullptr, *d_ws_task = nullptr; double* d_ws_base = nullptr;
  size_t cap_ws_stage = 0, cap_ws_entry = 0;
  // empc_plant_step staging (host-pointer plant), grown on demand
  double *d_plant_x = nullptr, *d_plant_u = nullptr, *d_plant_xn = nullptr; size_t cap_plant = 0;
  int n_ref = 0, dt_ref_ms = 0;
  long long* d_times = nullptr;   // n_node_maps controller times
  // contact dynamics (contact.cuh): has_contact = some cost set's model carries a contact; has_coupled = some cost set
  // holds a contact-force cost (backward_kernel<D, true>)
  // rk4: IntegratedActionModelRK4 (rk4.cuh).  overlay = has_contact || rk4 selects the <.., true> instantiations of the
  // rollout / decide / trial-cost kernels
  bool has_contact = false, has_coupled = false, rk4 = false, overlay = false;
  int width_a = RO_WIDTH_A;  // stage-A width of the line search (rollout.cuh)
  // small batches (one wave of node_calc_kernel or less): node_cost_kernel runs beside node_calc_kernel on a second stream
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  double *d_pack = nullptr, *h_pack = nullptr;  // empc_get_solution staging (small batches)
  WeightedScheduleDev wsched = {0, 0, nullptr, nullptr, 0, 0.0, 0.0, nullptr, nullptr, nullptr};
  int init_feasible = 0;
  size_t cap_iter_log = 0;  // records allocated for bf.iter_log (batch x bf.log_cap used)
  // empc_solve_stream staging: job table and per-job results on the device
  double *d_st_x0 = nullptr, *d_st_xs = nullptr, *d_st_us = nullptr, *d_st_uss = nullptr, *d_st_sc = nullptr; int* d_st_int = nullptr;
  size_t cap_st_x0 = 0, cap_st_xs = 0, cap_st_us = 0, cap_st_uss = 0, cap_st_sc = 0, cap_st_int = 0;
  cudaStream_t stream = nullptr;
  int* h_active = nullptr;  // pinned: [group][slot][2]
  static constexpr int MAXG = 4;
  int n_groups = 1;
  cudaStream_t gstream[MAXG] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_it[MAXG][2] = {}, ev_skew[MAXG] = {}, ev_init = nullptr, ev_gdone[MAXG] = {};
  // stats
  long long launches = 0, total_iterations = 0;
  int timing = 0;
  double ms_by_kernel[4] = {0, 0, 0, 0};
  long long units_by_kernel[4] = {0, 0, 0, 0};  // OCPs processed by each kernel family, summed over launches
  double solve_ms = 0;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_solve[2] = {nullptr, nullptr};
};

template <class Tp>
static cudaError_t dalloc(empc_solver* h, Tp** p, size_t n) {
  cudaError_t e = cudaMalloc((void**)p, n * sizeof(Tp));
  if (e == cudaSuccess) { h->allocs.push_back(*p); e = cudaMemsetAsync(*p, 0, n * sizeof(Tp), h->stream); }
  return e;
}

// (re)allocation of a buffer whose size may change between calls: reuse when it fits, otherwise free the old one
template <class Tp>
static cudaError_t dreuse(empc_solver* h, Tp** p, size_t* cap, size_t n) {
  if (*p && *cap >= n) return cudaSuccess;
  if (*p) {
    cudaFree(*p);  // synchronises with work that may still read it
    h->allocs.erase(std::remove(h->allocs.begin(), h->allocs.end(), (void*)*p), h->allocs.end());
    *p = nullptr; *cap = 0;
  }
  cudaError_t e = dalloc(h, p, n);
  if (e == cudaSuccess) *cap = n;
  return e;
}

// ---- (NA, NR) dispatch: one instantiation per eagle-mpc platform family ---------------------------------------------
#ifdef EMPC_DEBUG_ONLY_ARM3  /* development builds: one platform family, a quarter of the compile time */
#define EMPC_DISPATCH(h, CALL)                                               \
  do {                                                                       \
    if ((h)->na == 3 && (h)->nr == 6) { using D = Dim<3, 6>; CALL; }         \
    else return fail(EMPC_ERR_UNSUPPORTED, "debug build: flying_arm_3 only");                              \
  } while (0)
#else
#define EMPC_DISPATCH(h, CALL)                                               \
  do {                                                                       \
    if ((h)->na == 0 && (h)->nr == 4) { using D = Dim<0, 4>; CALL; }         \
    else if ((h)->na == 0 && (h)->nr == 6) { using D = Dim<0, 6>; CALL; }    \
    else if ((h)->na == 2 && (h)->nr == 6) { using D = Dim<2, 6>; CALL; }    \
    else if ((h)->na == 3 && (h)->nr == 6) { using D = Dim<3, 6>; CALL; }    \
    else if ((h)->na == 5 && (h)->nr == 6) { using D = Dim<5, 6>; CALL; }    \
    else return fail(EMPC_ERR_UNSUPPORTED, "unsupported (arm joints, rotors) combination");                \
  } while (0)
#endif

static int packet_doubles(int na, int nr) {
  int out = 0;
#define EMPC_PK(NA_, NR_) if (na == NA_ && nr == NR_) out = Pk<Dim<NA_, NR_>>::SIZE;
  EMPC_PK(0, 4) EMPC_PK(0, 6) EMPC_PK(2, 6) EMPC_PK(3, 6) EMPC_PK(5, 6)
#undef EMPC_PK
  return out;
}
static bool supported(int na, int nr) {
  return (na == 0 && nr == 4) || (na == 0 && nr == 6) || (na == 2 && nr == 6) || (na == 3 && nr == 6) || (na == 5 && nr == 6);
}

// Dynamic shared memory opt-in of the kernels that need more than 48 KB.  The attribute belongs to the (function, device)
// pair, so it is set for the handle's device every time a handle is created (one process may hold handles on several GPUs).
template <class K>
static cudaError_t opt_in_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
template <class D>
static cudaError_t set_kernel_attributes() {
  cudaError_t e;
  if ((e = opt_in_smem(node_diff_kernel<D>, sizeof(double) * DiffCfg<D>::SMEM_DOUBLES)) != cudaSuccess) return e;
  if ((e = opt_in_smem(backward_kernel<D, false>, sizeof(double) * BwCfg<D>::TOTAL)) != cudaSuccess) return e;
  if ((e = opt_in_smem(backward_kernel<D, true>, sizeof(double) * BwCfg<D>::TOTAL)) != cudaSuccess) return e;
  if ((e = opt_in_smem(backward_kernel<D, true, true>, sizeof(double) * BwCfg<D>::TOTAL)) != cudaSuccess) return e;
  if ((e = opt_in_smem(rollout_kernel<D, RO_WIDTH_A, false>, sizeof(double) * RoCfg<D, RO_WIDTH_A>::SMEM_DOUBLES)) != cudaSuccess) return e;
  if ((e = opt_in_smem(rollout_kernel<D, RO_WIDTH_A, true>, sizeof(double) * RoCfg<D, RO_WIDTH_A>::SMEM_DOUBLES)) != cudaSuccess) return e;
  if ((e = opt_in_smem(rollout_kernel<D, 8, true>, sizeof(double) * RoCfg<D, 8>::SMEM_DOUBLES)) != cudaSuccess) return e;
  return opt_in_smem(rollout_kernel<D, 8, false>, sizeof(double) * RoCfg<D, 8>::SMEM_DOUBLES);
}

static void inertia_matrix(double m, const double* c, const double* Ic, double* Y) {
  const double S[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
  double SS[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) SS[3 * i + j] = S[3 * i] * S[j] + S[3 * i + 1] * S[3 + j] + S[3 * i + 2] * S[6 + j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Y[6 * i + j] = (i == j) ? m : 0.0;
      Y[6 * i + 3 + j] = -m * S[3 * i + j];
      Y[6 * (3 + i) + j] = m * S[3 * i + j];
      Y[6 * (3 + i) + 3 + j] = Ic[3 * i + j] - m * SS[3 * i + j];
    }
}

extern "C" {

const char* empc_last_error(void) { return g_last_error.c_str(); }

void empc_default_params(empc_solver_params_t* p) {
  std::memset(p, 0, sizeof(*p));
  p->maxiter = 100; p->stop_gap_norm = 0; p->squash_quirk = 0;
  p->convergence_init = 1e-2; p->convergence_stop = 1e-3; p->convergence_mult = 1e-1;
  p->smooth_init = 0.1; p->smooth_mult = 0.5; p->barrier_weight = 1e-3;
  p->reg_init = 1e-9; p->reg_min = 1e-9; p->reg_max = 1e9; p->reg_factor = 10;
  p->th_acceptstep = 0.1; p->th_acceptnegstep = 2; p->th_grad = 1e-12; p->th_gaptol = 1e-16;
  p->th_stepdec = 0.5; p->th_stepinc = 0.01; p->th_stop_gaps = 1.0;
  p->solver_type = EMPC_SOLVER_SBFDDP;
  p->th_stop = 5e-5; p->boxqp_maxiter = 100; p->boxqp_th_acceptstep = 0.1; p->boxqp_th_grad = 1e-5; p->boxqp_reg = 0.0;
}

void empc_box_params(empc_solver_params_t* p, int32_t solver_type) {
  empc_default_params(p);
  p->solver_type = solver_type;
  p->stop_criteria = EMPC_STOP_CRITERIA_QU_NORM; p->stop_test = EMPC_STOP_TEST_FEASIBLE;  // upstream SolverFDDP / SolverDDP::solve
}

int empc_destroy(empc_solver_t* h) {
  if (!h) return EMPC_OK;
  cudaSetDevice(h->device);
  for (void* p : h->allocs) cudaFree(p);
  if (h->h_active) cudaFreeHost(h->h_active);
  if (h->h_pack) cudaFreeHost(h->h_pack);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_solve) if (e) cudaEventDestroy(e);
  for (int g = 0; g < empc_solver::MAXG; ++g) {
    if (h->gstream[g]) cudaStreamDestroy(h->gstream[g]);
    for (auto& e : h->ev_it[g]) if (e) cudaEventDestroy(e);
    if (h->ev_skew[g]) cudaEventDestroy(h->ev_skew[g]);
    if (h->ev_gdone[g]) cudaEventDestroy(h->ev_gdone[g]);
  }
  if (h->ev_init) cudaEventDestroy(h->ev_init);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return EMPC_OK;
}

int empc_create(const empc_problem_desc_t* d, int32_t batch, int32_t device, empc_solver_t** out) {
  if (!d || !out || batch <= 0) return fail(EMPC_ERR_INVALID, "null argument or batch <= 0");
  const empc_robot_t& r = d->robot;
  if (r.n_joints < 1 || r.n_joints > EMPC_MAX_JOINTS) return fail(EMPC_ERR_INVALID, "bad joint count");
  for (int i = 1; i < r.n_joints; ++i)
    if (r.parent[i] != i - 1) return fail(EMPC_ERR_UNSUPPORTED, "device path supports serial-chain arms only");
  if (!supported(r.n_joints - 1, d->n_rotors)) return fail(EMPC_ERR_UNSUPPORTED, "unsupported (arm joints, rotors) combination");
  if (d->T < 1 || d->n_node_maps < 1) return fail(EMPC_ERR_INVALID, "bad horizon / node maps");
  // contact dynamics: one ContactModel3D / 6D per cost set, zero Baumgarte gains; a friction-cone cost needs the contact
  // of its own cost set, on the same frame
  std::vector<unsigned char> coupled((size_t)std::max(1, d->n_costsets), 0);
  bool any_contact = false, any_coupled = false;
  if (d->n_contacts > 0) {
    if (!d->contacts || !d->costset_contact) return fail(EMPC_ERR_INVALID, "n_contacts > 0 without contact tables");
    for (int c = 0; c < d->n_contacts; ++c) {
      const empc_contact_t& ct = d->contacts[c];
      if (ct.type != EMPC_CONTACT_3D && ct.type != EMPC_CONTACT_6D) return fail(EMPC_ERR_INVALID, "bad contact type");
      if (ct.frame < 0 || ct.frame >= r.n_frames) return fail(EMPC_ERR_INVALID, "contact frame out of range");
      if (ct.gains[0] != 0.0 || ct.gains[1] != 0.0) return fail(EMPC_ERR_UNSUPPORTED, "non-zero contact gains are not supported");
    }
    for (int cs = 0; cs < d->n_costsets; ++cs) {
      if (d->costset_contact[cs] >= d->n_contacts) return fail(EMPC_ERR_INVALID, "costset_contact out of range");
      if (d->costset_contact[cs] >= 0) any_contact = true;
    }
  }
  if (d->integrator != EMPC_INTEGRATOR_EULER && d->integrator != EMPC_INTEGRATOR_RK4) return fail(EMPC_ERR_INVALID, "bad integrator");
  const bool is_rk4 = d->integrator == EMPC_INTEGRATOR_RK4;
  if (is_rk4 && any_contact) return fail(EMPC_ERR_UNSUPPORTED, "IntegratedActionModelRK4 over contact dynamics is not supported");
  if (is_rk4) { any_coupled = true; for (auto& c : coupled) c = 1; }  // the RK4 pull-back fills Lxu and the whole Luu of every node
  for (int cs = 0; cs < d->n_costsets; ++cs)
    for (int c = d->costset_begin[cs]; c < d->costset_begin[cs + 1]; ++c)
      if (d->costs[c].type == EMPC_COST_CONTACT_FRICTION_CONE) {
        const int ci = (d->n_contacts > 0) ? d->costset_contact[cs] : -1;
        if (ci < 0 || d->contacts[ci].frame != d->costs[c].frame)
          return fail(EMPC_ERR_INVALID, "a friction-cone cost needs a contact on the same frame in its cost set");
        coupled[(size_t)cs] = 1; any_coupled = true;
      }
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(EMPC_ERR_CUDA, "no such CUDA device");
  CK(cudaSetDevice(device));

  empc_solver* h = new empc_solver();
  h->device = device; h->B = batch; h->T = d->T;
  h->na = r.n_joints - 1; h->nr = d->n_rotors;
  h->nq = 7 + h->na; h->nv = 6 + h->na; h->nx = h->nq + h->nv; h->ndx = 2 * h->nv; h->nu = h->nr + h->na;
  empc_default_params(&h->P);
  // 8 trials per OCP in stage A while that still leaves at most one rollout warp per SM sub-partition
  h->width_a = ((long long)batch * RO_WIDTH_A_SMALL <= 32LL * 148 * 4) ? RO_WIDTH_A_SMALL : RO_WIDTH_A;
  if (const char* e = std::getenv("EMPC_RO_WIDTH_A")) { const int w = std::atoi(e); if (w == RO_WIDTH_A || w == RO_WIDTH_A_SMALL) h->width_a = w; }
  h->n_costs = d->n_costs; h->n_pool = d->n_pool; h->n_node_maps = d->n_node_maps; h->n_costsets = d->n_costsets;
  h->has_contact = any_contact; h->has_coupled = any_coupled; h->rk4 = is_rk4; h->overlay = any_contact || is_rk4;

  DevModel& M = h->hmodel;
  std::memset(&M, 0, sizeof(M));
  M.nj = r.n_joints; M.na = h->na; M.nq = h->nq; M.nv = h->nv; M.nx = h->nx; M.ndx = h->ndx; M.nu = h->nu; M.nr = h->nr;
  M.T = d->T; M.use_squash = d->use_squash; M.n_frames = r.n_frames; M.dt = d->dt; M.integrator = d->integrator;
  const int ndx = h->ndx, nu = h->nu;
  M.oFx = 0; M.oFu = ndx * ndx; M.oLxx = M.oFu + ndx * nu; M.oLxu = M.oLxx + ndx * ndx; M.oLuu = M.oLxu + ndx * nu;
  M.oLx = M.oLuu + nu * nu; M.oLu = M.oLx + ndx;
  M.tile = M.oLu + nu; M.tile += M.tile & 1;
  h->tile = M.tile;
  for (int i = 0; i < r.n_joints; ++i) {
    std::memcpy(M.jR[i], r.jplace_R[i], 72); std::memcpy(M.jp[i], r.jplace_p[i], 24); std::memcpy(M.axis[i], r.axis[i], 24);
    inertia_matrix(r.mass[i], r.com[i], r.inertia[i], M.Y[i]);
    M.mass[i] = r.mass[i]; std::memcpy(M.com[i], r.com[i], 24); std::memcpy(M.Ic[i], r.inertia[i], 72);
  }
  for (int i = 0; i < 3; ++i) { M.a0[i] = -r.gravity[i]; M.a0[3 + i] = 0; }
  for (int f = 0; f < r.n_frames; ++f) {
    M.frame_joint[f] = r.frame_joint[f];
    std::memcpy(M.fR[f], r.frame_R[f], 72); std::memcpy(M.fp[f], r.frame_p[f], 24);
  }
  std::memcpy(M.tau_f, d->tau_f, sizeof(double) * 6 * h->nr);
  for (int i = 0; i < nu; ++i) {
    M.u_lb[i] = d->u_lb[i]; M.u_ub[i] = d->u_ub[i];
    const double mid = 0.5 * (d->u_lb[i] + d->u_ub[i]), dd = 0.5 * (d->u_ub[i] - d->u_lb[i]);
    M.bar_lb[i] = mid - 1.0 * dd; M.bar_ub[i] = mid + 1.0 * dd;  // crocoddyl::ActivationBounds ctor, beta = 1
  }
  M.barrier_weight = h->P.barrier_weight;

  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) {
#define EMPC_ATTR(NA_, NR_) if (h->na == NA_ && h->nr == NR_) e = set_kernel_attributes<Dim<NA_, NR_>>();
#ifdef EMPC_DEBUG_ONLY_ARM3
    EMPC_ATTR(3, 6)
#else
    EMPC_ATTR(0, 4) EMPC_ATTR(0, 6) EMPC_ATTR(2, 6) EMPC_ATTR(3, 6) EMPC_ATTR(5, 6)
#endif
#undef EMPC_ATTR
  }
  if (e != cudaSuccess) { std::string m_ = cudaGetErrorString(e); empc_destroy(h); return fail(EMPC_ERR_CUDA, m_); }
#define CKH(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e_); empc_destroy(h); return fail(EMPC_ERR_CUDA, m_); } } while (0)
  const size_t B = batch, T = d->T, T1 = T + 1, nx = h->nx, tile = h->tile;
  CKH(cudaMallocHost((void**)&h->h_active, empc_solver::MAXG * 2 * 2 * sizeof(int)));
  for (int g = 0; g < empc_solver::MAXG; ++g) {
    CKH(cudaStreamCreateWithFlags(&h->gstream[g], cudaStreamNonBlocking));
    for (auto& e : h->ev_it[g]) CKH(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CKH(cudaEventCreateWithFlags(&h->ev_skew[g], cudaEventDisableTiming));
    CKH(cudaEventCreateWithFlags(&h->ev_gdone[g], cudaEventDisableTiming));
  }
  CKH(cudaEventCreateWithFlags(&h->ev_init, cudaEventDisableTiming));
  // batch groups: the rollout is latency-bound (one wave), calc_diff/backward are throughput-bound; skewed groups on
  // separate streams let one group's rollout overlap the others' calc_diff + backward
  h->n_groups = 1;  // measured on B200: grouping does not pay at B=4096 (profiles/), kept as an opt-in (EMPC_GROUPS)
  if (const char* e = std::getenv("EMPC_GROUPS")) h->n_groups = std::max(1, std::min((int)empc_solver::MAXG, std::atoi(e)));
  for (auto& ev : h->ev) CKH(cudaEventCreate(&ev));
  for (auto& ev : h->ev_solve) CKH(cudaEventCreate(&ev));
  CKH(dalloc(h, &h->d_model, 1));
  CKH(dalloc(h, &h->d_costs, (size_t)std::max(1, d->n_costs)));
  CKH(dalloc(h, &h->d_pool, (size_t)std::max(1, d->n_pool)));
  CKH(dalloc(h, &h->d_costset_begin, (size_t)d->n_costsets + 1));
  CKH(dalloc(h, &h->d_node_costset, (size_t)d->n_node_maps * T1));
  CKH(dalloc(h, &h->d_ocp_map, B));
  CKH(dalloc(h, &h->d_x0, B * nx));
  CKH(dalloc(h, &h->d_xs_init, B * T1 * nx));
  CKH(dalloc(h, &h->d_us_init, B * T * nu));
  Buffers& bf = h->bf;
  std::memset(&bf, 0, sizeof(bf));
  bf.B = batch; bf.T = d->T; bf.b0 = 0; bf.nb = batch;
  CKH(dalloc(h, &bf.st, B));
  CKH(dalloc(h, &bf.xs, B * T1 * nx));
  CKH(dalloc(h, &bf.us, B * T * nu));
  CKH(dalloc(h, &bf.xs_try0, B * nx));
  CKH(dalloc(h, &bf.tiles, B * T1 * tile));
  CKH(dalloc(h, &bf.node_dense, B * T1));
  CKH(dalloc(h, &bf.packets, (B * T1 + 7) / 8 * 8 * (size_t)packet_doubles(h->na, h->nr)));
  CKH(dalloc(h, &bf.xnext, B * T1 * nx));
  CKH(dalloc(h, &bf.node_cost, B * T1));
  CKH(dalloc(h, &bf.fs, B * T1 * ndx));
  CKH(dalloc(h, &bf.gap_inf, B * T1));
  CKH(dalloc(h, &bf.gap_l1, B * T1));
  CKH(dalloc(h, &bf.K, B * T * nu * ndx));
  CKH(dalloc(h, &bf.k, B * T * nu));
  CKH(dalloc(h, &bf.Vx, B * T1 * ndx));
  CKH(dalloc(h, &bf.g, B * T1 * ndx));
  CKH(dalloc(h, &bf.nodesc, B * T1 * 4));
  CKH(dalloc(h, &bf.xs_try, (size_t)EMPC_N_ALPHAS * B * T1 * nx));
  CKH(dalloc(h, &bf.us_try, (size_t)EMPC_N_ALPHAS * B * T * nu));
  CKH(dalloc(h, &bf.cost_try, B * EMPC_N_ALPHAS));
  CKH(dalloc(h, &bf.trial_node_cost, (size_t)EMPC_N_ALPHAS * B * T1));
  CKH(dalloc(h, &bf.dv, B * EMPC_N_ALPHAS));
  CKH(dalloc(h, &bf.ok, B * EMPC_N_ALPHAS));
  CKH(dalloc(h, &bf.us_squash, B * T * nu));
  CKH(dalloc(h, &bf.qu2, B * T1));
  CKH(dalloc(h, &bf.n_active, 2 * empc_solver::MAXG));
  bf.model = h->d_model; bf.ct.costs = h->d_costs; bf.ct.pool = h->d_pool; bf.ct.costset_begin = h->d_costset_begin;
  bf.node_costset = h->d_node_costset; bf.ocp_map = h->d_ocp_map; bf.x0 = h->d_x0;
  CKH(cudaMemcpyAsync(h->d_model, &M, sizeof(M), cudaMemcpyHostToDevice, h->stream));
  if (d->n_costs) CKH(cudaMemcpyAsync(h->d_costs, d->costs, sizeof(empc_cost_t) * d->n_costs, cudaMemcpyHostToDevice, h->stream));
  if (d->n_pool) CKH(cudaMemcpyAsync(h->d_pool, d->pool, sizeof(double) * d->n_pool, cudaMemcpyHostToDevice, h->stream));
  CKH(cudaMemcpyAsync(h->d_costset_begin, d->costset_begin, sizeof(int) * (d->n_costsets + 1), cudaMemcpyHostToDevice, h->stream));
  CKH(cudaMemcpyAsync(h->d_node_costset, d->node_costset, sizeof(int) * d->n_node_maps * T1, cudaMemcpyHostToDevice, h->stream));
  {  // overlay tables (contact / coupled-cost flags per cost set): also read by the Box solvers' kernel instantiations
    empc_contact_t* d_contacts = nullptr; int* d_cc = nullptr; unsigned char* d_cpl = nullptr;
    std::vector<int> cc((size_t)d->n_costsets, -1);
    if (any_contact) std::copy(d->costset_contact, d->costset_contact + d->n_costsets, cc.begin());
    CKH(dalloc(h, &d_contacts, (size_t)std::max(1, d->n_contacts)));
    CKH(dalloc(h, &d_cc, (size_t)d->n_costsets));
    CKH(dalloc(h, &d_cpl, (size_t)d->n_costsets));
    if (any_contact) CKH(cudaMemcpyAsync(d_contacts, d->contacts, sizeof(empc_contact_t) * d->n_contacts, cudaMemcpyHostToDevice, h->stream));
    CKH(cudaMemcpy(d_cc, cc.data(), sizeof(int) * d->n_costsets, cudaMemcpyHostToDevice));
    CKH(cudaMemcpyAsync(d_cpl, coupled.data(), (size_t)d->n_costsets, cudaMemcpyHostToDevice, h->stream));
    bf.ct.contacts = d_contacts; bf.ct.costset_contact = d_cc; bf.ct.costset_coupled = d_cpl;
  }
  CKH(cudaStreamSynchronize(h->stream));
  // default x0 / candidate: state.zero()
  {
    std::vector<double> x0(B * nx, 0.0);
    for (size_t b = 0; b < B; ++b) x0[b * nx + 6] = 1.0;
    CKH(cudaMemcpy(h->d_x0, x0.data(), sizeof(double) * x0.size(), cudaMemcpyHostToDevice));
  }
  *out = h;
  int rc = empc_set_candidate(h, nullptr, nullptr, 0);
  if (rc != EMPC_OK) { empc_destroy(h); *out = nullptr; return rc; }
  return EMPC_OK;
}

int empc_get_dims(const empc_solver_t* h, empc_dims_t* o) {
  if (!h || !o) return fail(EMPC_ERR_INVALID, "null");
  o->nq = h->nq; o->nv = h->nv; o->nx = h->nx; o->ndx = h->ndx; o->nu = h->nu; o->T = h->T; o->batch = h->B; o->tile = h->tile;
  return EMPC_OK;
}

int empc_set_x0(empc_solver_t* h, const double* x0) {
  if (!h || !x0) return fail(EMPC_ERR_INVALID, "null");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_x0, x0, sizeof(double) * h->B * h->nx, cudaMemcpyDefault, h->stream));  // x0: host or device memory
  CK(cudaStreamSynchronize(h->stream));
  return EMPC_OK;
}

static int load_candidate(empc_solver* h) {
  const size_t nxs = (size_t)h->B * (h->T + 1) * h->nx, nus = (size_t)h->B * h->T * h->nu;
  CK(cudaMemcpyAsync(h->bf.xs, h->d_xs_init, sizeof(double) * nxs, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(h->bf.us, h->d_us_init, sizeof(double) * nus, cudaMemcpyDeviceToDevice, h->stream));
  return EMPC_OK;
}

int empc_set_candidate(empc_solver_t* h, const double* xs, const double* us, int32_t is_feasible) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  CK(cudaSetDevice(h->device));
  const size_t nxs = (size_t)h->B * (h->T + 1) * h->nx, nus = (size_t)h->B * h->T * h->nu;
  if (xs) CK(cudaMemcpyAsync(h->d_xs_init, xs, sizeof(double) * nxs, cudaMemcpyDefault, h->stream));
  else {
    zero_candidate_kernel<<<(unsigned)((nxs + 255) / 256), 256, 0, h->stream>>>(h->d_xs_init, nxs / h->nx, h->nx);
    CK(cudaGetLastError());
  }
  if (us) CK(cudaMemcpyAsync(h->d_us_init, us, sizeof(double) * nus, cudaMemcpyDefault, h->stream));
  else CK(cudaMemsetAsync(h->d_us_init, 0, sizeof(double) * nus, h->stream));
  h->init_feasible = is_feasible ? 1 : 0;
  int rc = load_candidate(h);
  if (rc) return rc;
  init_state_kernel<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->bf, h->P, h->init_feasible, h->nx);
  override_state_kernel<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->bf, 0.0, h->init_feasible, 0);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return EMPC_OK;
}

int empc_set_params(empc_solver_t* h, const empc_solver_params_t* p) {
  if (!h || !p) return fail(EMPC_ERR_INVALID, "null");
  if (p->maxiter < 1) return fail(EMPC_ERR_INVALID, "maxiter < 1");
  if (p->solver_type != EMPC_SOLVER_SBFDDP && p->solver_type != EMPC_SOLVER_BOXFDDP && p->solver_type != EMPC_SOLVER_BOXDDP)
    return fail(EMPC_ERR_INVALID, "solver_type is not one of EMPC_SOLVER_*");
  if (p->solver_type != EMPC_SOLVER_SBFDDP && p->boxqp_maxiter < 1) return fail(EMPC_ERR_INVALID, "boxqp_maxiter < 1");
  h->P = *p;
  if (h->hmodel.barrier_weight != p->barrier_weight) {  // the device copy of the model only changes with this field
    h->hmodel.barrier_weight = p->barrier_weight;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpy(h->d_model, &h->hmodel, sizeof(DevModel), cudaMemcpyHostToDevice));
  }
  return EMPC_OK;
}

int empc_set_node_maps(empc_solver_t* h, const int32_t* m) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  CK(cudaSetDevice(h->device));
  if (m) {
    for (int b = 0; b < h->B; ++b) if (m[b] < 0 || m[b] >= h->n_node_maps) return fail(EMPC_ERR_INVALID, "node map out of range");
    CK(cudaMemcpy(h->d_ocp_map, m, sizeof(int) * h->B, cudaMemcpyHostToDevice));
  } else CK(cudaMemset(h->d_ocp_map, 0, sizeof(int) * h->B));
  return EMPC_OK;
}

int empc_update_costs(empc_solver_t* h, int32_t first, int32_t n, const empc_cost_t* costs, int32_t pool_off, int32_t n_pool, const double* pool) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  if (n < 0 || first < 0 || first + n > h->n_costs || n_pool < 0 || pool_off < 0 || pool_off + n_pool > h->n_pool)
    return fail(EMPC_ERR_INVALID, "cost / pool range out of bounds");
  CK(cudaSetDevice(h->device));
  if (n) CK(cudaMemcpyAsync(h->d_costs + first, costs, sizeof(empc_cost_t) * n, cudaMemcpyHostToDevice, h->stream));
  if (n_pool) CK(cudaMemcpyAsync(h->d_pool + pool_off, pool, sizeof(double) * n_pool, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return EMPC_OK;
}

int empc_update_node_costsets(empc_solver_t* h, const int32_t* nc) {
  if (!h || !nc) return fail(EMPC_ERR_INVALID, "null");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpy(h->d_node_costset, nc, sizeof(int) * h->n_node_maps * (h->T + 1), cudaMemcpyHostToDevice));
  return EMPC_OK;
}

// ---- batched MPC instances: private cost tables per instance, device-side rail retargeting (retarget.cuh) ----
int empc_replicate_instances(empc_solver_t* h, int32_t n) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  if (n < 1) return fail(EMPC_ERR_INVALID, "n_instances < 1");
  if (h->n_node_maps != 1) return fail(EMPC_ERR_INVALID, "instances are replicated from a problem with a single node map");
  if (h->rk4) return fail(EMPC_ERR_UNSUPPORTED, "batched MPC instances with the RK4 integrator are not supported (cost sets are flagged per set, not per instance)");
  if (h->has_contact) return fail(EMPC_ERR_UNSUPPORTED, "MPC instances with contact stages are not supported (the reference's controllers refuse them, src/mpc-controllers/carrot-mpc.cpp:204-206)");
  if ((long long)n * std::max(h->n_pool, h->n_costs) > 0x7fffffffLL / 2) return fail(EMPC_ERR_INVALID, "replicated tables exceed the 32-bit offsets");
  CK(cudaSetDevice(h->device));
  const int T1 = h->T + 1;
  empc_cost_t* costs = nullptr; double* pool = nullptr; int *begin = nullptr, *node_set = nullptr;
  CK(dalloc(h, &costs, (size_t)std::max(1, n * h->n_costs)));
  CK(dalloc(h, &pool, (size_t)std::max(1, n * h->n_pool)));
  CK(dalloc(h, &begin, (size_t)n * h->n_costsets + 1));
  CK(dalloc(h, &node_set, (size_t)n * T1));
  CK(dalloc(h, &h->d_times, (size_t)n));
  replicate_tables_kernel<<<296, 256, 0, h->stream>>>(h->d_costs, h->n_costs, h->d_pool, h->n_pool, h->d_costset_begin, h->n_costsets,
                                                       h->d_node_costset, T1, n, costs, pool, begin, node_set);
  CK(cudaGetLastError());
  std::vector<int> map((size_t)h->B);
  for (int b = 0; b < h->B; ++b) map[(size_t)b] = b % n;
  CK(cudaMemcpyAsync(h->d_ocp_map, map.data(), sizeof(int) * h->B, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  // the single-map tables stay allocated until empc_destroy (they are small); the handle now points at the copies
  h->d_costs = costs; h->d_pool = pool; h->d_costset_begin = begin; h->d_node_costset = node_set;
  h->bf.ct.costs = costs; h->bf.ct.pool = pool; h->bf.ct.costset_begin = begin; h->bf.node_costset = node_set;
  h->n_costs *= n; h->n_pool *= n; h->n_costsets *= n; h->n_node_maps = n;
  return EMPC_OK;
}

int empc_set_reference_trajectory(empc_solver_t* h, const double* state_ref, int32_t n_ref, int32_t dt_ref_ms) {
  if (!h || !state_ref) return fail(EMPC_ERR_INVALID, "null");
  if (n_ref < 1 || dt_ref_ms < 1) return fail(EMPC_ERR_INVALID, "empty reference trajectory / dt_ref < 1 ms");
  CK(cudaSetDevice(h->device));
  const int nx = h->nx, nq = h->nq;
  std::vector<double> tab((size_t)(n_ref + 2) * nx, 0.0);
  std::copy(state_ref, state_ref + (size_t)n_ref * nx, tab.begin());
  // hover row (rail-mpc.cpp:180-186): last configuration, zero velocity, quaternion rebuilt from its (z, w) pair only
  // -- the x / y components copied along with the configuration stay
  const double* last = state_ref + (size_t)(n_ref - 1) * nx;
  double* hov = tab.data() + (size_t)n_ref * nx;
  std::copy(last, last + nq, hov);
  const double w = last[6], z = last[5], norm = std::sqrt(w * w + z * z);
  hov[5] = z / norm; hov[6] = w / norm;
  // carrot tail row (carrot-mpc.cpp:376-388): last configuration as it is, zero velocity
  std::copy(last, last + nq, hov + nx);
  CK(dreuse(h, &h->d_ref_table, &h->cap_ref_table, tab.size()));
  CK(cudaMemcpyAsync(h->d_ref_table, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->n_ref = n_ref; h->dt_ref_ms = dt_ref_ms;
  return EMPC_OK;
}

int empc_rail_retarget(empc_solver_t* h, const int64_t* times_ms, int32_t dt_node_ms) {
  if (!h || !times_ms) return fail(EMPC_ERR_INVALID, "null");
  if (!h->d_ref_table) return fail(EMPC_ERR_INVALID, "empc_set_reference_trajectory has not been called");
  if (dt_node_ms < 1) return fail(EMPC_ERR_INVALID, "dt_node < 1 ms");
  for (int m = 0; m < h->n_node_maps; ++m) if (times_ms[m] < 0) return fail(EMPC_ERR_INVALID, "negative controller time");
  CK(cudaSetDevice(h->device));
  if (!h->d_times) CK(dalloc(h, &h->d_times, (size_t)h->n_node_maps));
  static_assert(sizeof(long long) == sizeof(int64_t), "controller times are 64-bit");
  CK(cudaMemcpyAsync(h->d_times, times_ms, sizeof(int64_t) * h->n_node_maps, cudaMemcpyHostToDevice, h->stream));
  const int T1 = h->T + 1, total = h->n_node_maps * T1;
  rail_retarget_kernel<<<(total + 127) / 128, 128, 0, h->stream>>>(h->d_costs, h->d_pool, h->d_costset_begin, h->d_node_costset, T1,
                                                                   h->n_node_maps, h->d_times, dt_node_ms, h->d_ref_table, h->n_ref,
                                                                   h->dt_ref_ms, h->nx);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return EMPC_OK;
}

int empc_set_carrot_schedule(empc_solver_t* h, int32_t n_stages, const int64_t* t_stages, const uint8_t* is_transition) {
  if (!h || !t_stages || !is_transition) return fail(EMPC_ERR_INVALID, "null");
  if (n_stages < 1) return fail(EMPC_ERR_INVALID, "n_stages < 1");
  CK(cudaSetDevice(h->device));
  CK(dreuse(h, &h->d_t_stages, &h->cap_t_stages, (size_t)n_stages + 1));
  CK(dreuse(h, &h->d_is_transition, &h->cap_is_transition, (size_t)n_stages));
  CK(cudaMemcpyAsync(h->d_t_stages, t_stages, sizeof(int64_t) * ((size_t)n_stages + 1), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_is_transition, is_transition, (size_t)n_stages, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->n_stages = n_stages;
  return EMPC_OK;
}

int empc_carrot_retarget(empc_solver_t* h, const int64_t* times_ms, int32_t dt_node_ms) {
  if (!h || !times_ms) return fail(EMPC_ERR_INVALID, "null");
  if (!h->d_ref_table) return fail(EMPC_ERR_INVALID, "empc_set_reference_trajectory has not been called");
  if (!h->d_t_stages) return fail(EMPC_ERR_INVALID, "empc_set_carrot_schedule has not been called");
  if (dt_node_ms < 1) return fail(EMPC_ERR_INVALID, "dt_node < 1 ms");
  for (int m = 0; m < h->n_node_maps; ++m) if (times_ms[m] < 0) return fail(EMPC_ERR_INVALID, "negative controller time");
  CK(cudaSetDevice(h->device));
  if (!h->d_times) CK(dalloc(h, &h->d_times, (size_t)h->n_node_maps));
  CK(cudaMemcpyAsync(h->d_times, times_ms, sizeof(int64_t) * h->n_node_maps, cudaMemcpyHostToDevice, h->stream));
  const int T1 = h->T + 1, total = h->n_node_maps * T1;
  carrot_retarget_kernel<<<(total + 127) / 128, 128, 0, h->stream>>>(h->d_costs, h->d_pool, h->d_costset_begin, h->d_node_costset, T1,
                                                                     h->n_node_maps, h->d_times, dt_node_ms, h->d_ref_table, h->n_ref,
                                                                     h->dt_ref_ms, h->nx, h->n_stages, h->d_t_stages, h->d_is_transition);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return EMPC_OK;
}

int empc_set_weighted_schedule(empc_solver_t* h, const empc_weighted_schedule_t* s) {
  if (!h || !s) return fail(EMPC_ERR_INVALID, "null");
  if (s->n_stages < 1 || s->n_slots < 1 || !s->t_ini || !s->t_end || !s->match || !s->task || !s->base)
    return fail(EMPC_ERR_INVALID, "incomplete weighted schedule");
  CK(cudaSetDevice(h->device));
  const size_t ns = (size_t)s->n_stages, ne = ns * (size_t)s->n_slots;
  {
    size_t c1 = h->cap_ws_stage, c2 = h->cap_ws_stage, c3 = h->cap_ws_entry, c4 = h->cap_ws_entry, c5 = h->cap_ws_entry;
    CK(dreuse(h, &h->d_ws_t_ini, &c1, ns)); CK(dreuse(h, &h->d_ws_t_end, &c2, ns));
    CK(dreuse(h, &h->d_ws_match, &c3, ne)); CK(dreuse(h, &h->d_ws_task, &c4, ne)); CK(dreuse(h, &h->d_ws_base, &c5, ne));
    h->cap_ws_stage = std::min(c1, c2); h->cap_ws_entry = std::min(c3, std::min(c4, c5));
  }
  long long *t_ini = h->d_ws_t_ini, *t_end = h->d_ws_t_end; unsigned char *match = h->d_ws_match, *task = h->d_ws_task; double* base = h->d_ws_base;
  CK(cudaMemcpyAsync(t_ini, s->t_ini, sizeof(int64_t) * ns, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(t_end, s->t_end, sizeof(int64_t) * ns, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(match, s->match, ne, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(task, s->task, ne, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(base, s->base, sizeof(double) * ne, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->wsched = WeightedScheduleDev{s->n_stages, s->n_slots, t_ini, t_end, (long long)s->duration, s->alpha, s->beta, match, task, base};
  return EMPC_OK;
}

int empc_weighted_retarget(empc_solver_t* h, const int64_t* times_ms, int32_t dt_node_ms) {
  if (!h || !times_ms) return fail(EMPC_ERR_INVALID, "null");
  if (!h->wsched.base) return fail(EMPC_ERR_INVALID, "empc_set_weighted_schedule has not been called");
  if (dt_node_ms < 1) return fail(EMPC_ERR_INVALID, "dt_node < 1 ms");
  for (int m = 0; m < h->n_node_maps; ++m) if (times_ms[m] < 0) return fail(EMPC_ERR_INVALID, "negative controller time");
  CK(cudaSetDevice(h->device));
  if (!h->d_times) CK(dalloc(h, &h->d_times, (size_t)h->n_node_maps));
  CK(cudaMemcpyAsync(h->d_times, times_ms, sizeof(int64_t) * h->n_node_maps, cudaMemcpyHostToDevice, h->stream));
  weighted_retarget_kernel<<<h->n_node_maps, 128, sizeof(int) * (h->T + 1), h->stream>>>(h->d_costs, h->d_costset_begin, h->d_node_costset,
                                                                                        h->T + 1, h->n_node_maps, h->d_times, dt_node_ms,
                                                                                        h->wsched);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return EMPC_OK;
}

}  // extern "C"

// ---- launches -------------------------------------------------------------------------------------------------------
template <class D>
static cudaError_t launch_calc_diff(empc_solver* h, int force, double smooth, const Buffers* gb = nullptr, cudaStream_t st = nullptr) {
  const Buffers& bf = gb ? *gb : h->bf;
  if (!st) st = h->stream;
  const int T1 = h->T + 1;
  const long long n = (long long)bf.nb * T1;
  // the two thread-per-node halves are independent (different packet fields); when they do not fill the GPU they overlap
  if (h->rk4) {  // IntegratedActionModelRK4: the whole node model is rk4_node_kernel (rk4.cuh)
    rk4_node_kernel<D><<<(unsigned)((n + EMPC_RK4_THREADS - 1) / EMPC_RK4_THREADS), EMPC_RK4_THREADS, 0, st>>>(bf, force, smooth, h->hmodel);
    h->launches++;
    return cudaGetLastError();
  }
  const bool fork = h->side_stream && !gb && st == h->stream && n <= 148LL * NC_THREADS * EMPC_NC_BLOCKS;
  cudaError_t e;
  if (fork) {
    if ((e = cudaEventRecord(h->ev_fork, st)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0)) != cudaSuccess) return e;
  }
  node_calc_kernel<D><<<(unsigned)((n + NC_THREADS - 1) / NC_THREADS), NC_THREADS, 0, st>>>(bf, force, smooth, h->hmodel);
  h->launches++;
  node_cost_kernel<D><<<(unsigned)((n + 127) / 128), 128, 0, fork ? h->side_stream : st>>>(bf, force, smooth, h->hmodel);
  h->launches++;
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (fork) {
    if ((e = cudaEventRecord(h->ev_join, h->side_stream)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(st, h->ev_join, 0)) != cudaSuccess) return e;
  }
  using W = DiffCfg<D>;
  const size_t smem = sizeof(double) * W::SMEM_DOUBLES;
  const long long n_first = (long long)bf.b0 * T1, n_last = n_first + n - 1;
  const long long groups = n_last / Pk<D>::GROUP - n_first / Pk<D>::GROUP + 1;
  node_diff_kernel<D><<<(unsigned)groups, W::THREADS, smem, st>>>(bf, force, smooth, h->hmodel);
  h->launches++;
  if (h->has_contact) {  // the nodes of contact stages: calc + calcDiff again under the contact dynamics (contact.cuh)
    contact_node_kernel<D><<<(unsigned)((n + EMPC_CONTACT_THREADS - 1) / EMPC_CONTACT_THREADS), EMPC_CONTACT_THREADS, 0, st>>>(bf, force, smooth, h->hmodel);
    h->launches++;
  }
  return cudaGetLastError();
}
// the Box solvers run on the overlay instantiation of the rollout kernel (the clamp of the trial controls lives there), so
// that the rollout of the SbFDDP free path stays as it is; their node costs need nothing from the overlays, so decide_kernel
// and trial_cost_kernel stay on the free instantiation (12x faster than the overlay one, which carries the contact solve)
static inline bool box_solver(const empc_solver* h) { return h->P.solver_type != EMPC_SOLVER_SBFDDP; }
static inline bool use_overlay(const empc_solver* h) { return h->overlay || box_solver(h); }
template <class D>
static cudaError_t launch_backward(empc_solver* h, int force, const Buffers* gb = nullptr, cudaStream_t st = nullptr) {
  const Buffers& bf = gb ? *gb : h->bf;
  if (!st) st = h->stream;
  using S = BwCfg<D>;
  const size_t smem = sizeof(double) * S::TOTAL;
  BwParams P{h->P.reg_max, h->P.reg_factor, h->P.th_gaptol, force, h->P.stop_criteria == EMPC_STOP_CRITERIA_QU_NORM,
             h->P.boxqp_maxiter, h->P.boxqp_th_acceptstep, h->P.boxqp_th_grad, h->P.boxqp_reg};
  if (box_solver(h)) backward_kernel<D, true, true><<<bf.nb, 32, smem, st>>>(bf, P);  // SolverBoxFDDP / SolverBoxDDP gains
  else if (h->has_coupled) backward_kernel<D, true><<<bf.nb, 32, smem, st>>>(bf, P);
  else backward_kernel<D, false><<<bf.nb, 32, smem, st>>>(bf, P);  // one warp per OCP
  h->launches++;
  return cudaGetLastError();
}
template <class D, int W>
static cudaError_t launch_rollout_w(empc_solver* h, const RoParams& P, const Buffers& bf, cudaStream_t st) {
  using S = RoCfg<D, W>;
  const size_t smem = sizeof(double) * S::SMEM_DOUBLES;
  // one warp per block = 32/W OCPs x W step lengths
  if (use_overlay(h)) rollout_kernel<D, W, true><<<(bf.nb + S::OCPS - 1) / S::OCPS, 32, smem, st>>>(bf, P, h->hmodel);
  else rollout_kernel<D, W, false><<<(bf.nb + S::OCPS - 1) / S::OCPS, 32, smem, st>>>(bf, P, h->hmodel);
  h->launches++;
  return cudaGetLastError();
}
// stage 0: step lengths [0, width_a) of every active OCP; stage 1: the remaining ones, pending OCPs only
template <class D>
static cudaError_t launch_rollout(empc_solver* h, int stage, int force, int feasible, int ddp, double smooth, const Buffers* gb = nullptr,
                                  cudaStream_t st = nullptr) {
  const Buffers& bf = gb ? *gb : h->bf;
  if (!st) st = h->stream;
  RoParams P{force, feasible, ddp, smooth, stage == 0 ? 0 : h->width_a, box_solver(h) ? 1 : 0};
  cudaError_t e = (stage == 0 && h->width_a == RO_WIDTH_A) ? launch_rollout_w<D, RO_WIDTH_A>(h, P, bf, st) : launch_rollout_w<D, 8>(h, P, bf, st);
  if (e != cudaSuccess) return e;
  // In a solve, decide_kernel evaluates the trial costs lazily in line-search order.  The tile-level parity hook wants
  // cost_try of EVERY step length: node costs in parallel over trials and nodes, then the ordered per-trial sums.
  if (!force) return cudaSuccess;
  const int width = (stage == 0) ? h->width_a : EMPC_N_ALPHAS - h->width_a;
  const long long n_thr = (long long)bf.nb * width * (h->T + 1);
  if (h->overlay) trial_cost_kernel<D, true><<<(unsigned)((n_thr + 127) / 128), 128, 0, st>>>(bf, P, width, h->hmodel);
  else trial_cost_kernel<D, false><<<(unsigned)((n_thr + 127) / 128), 128, 0, st>>>(bf, P, width, h->hmodel);
  h->launches++;
  const long long n_warps = (long long)bf.nb * width;
  trial_sum_kernel<<<(unsigned)((n_warps + 3) / 4), 128, 0, st>>>(bf, P, width);
  h->launches++;
  return cudaGetLastError();
}
template <class D>
static cudaError_t launch_decide(empc_solver* h, int stage, const Buffers* gb = nullptr, cudaStream_t st = nullptr) {
  const Buffers& bf = gb ? *gb : h->bf;
  if (!st) st = h->stream;
  DecideParams dp{h->P, stage, h->width_a};
  // one block per OCP, one node per thread and round: the block size (a multiple of 32, at most 256) that wastes the fewest
  // thread slots over the ceil((T+1)/threads) rounds of the trial-cost evaluation; ties go to the larger block
  const int T1 = h->T + 1;
  int threads = 128; double best = 0.0;
  for (int n = 32; n <= 256; n += 32) {
    const int rounds = (T1 + n - 1) / n;
    const double eff = (double)T1 / ((double)rounds * n);
    if (eff >= best - 1e-12) { best = eff; threads = n; }
  }
  if (h->overlay) decide_kernel<D, true><<<bf.nb, threads, 0, st>>>(bf, dp, h->hmodel);
  else decide_kernel<D, false><<<bf.nb, threads, 0, st>>>(bf, dp, h->hmodel);
  h->launches++;
  return cudaGetLastError();
}
template <class D>
static cudaError_t launch_squash_out(empc_solver* h) {
  const int n = h->B * h->T;
  squash_out_kernel<D><<<(n + 127) / 128, 128, 0, h->stream>>>(h->bf);
  h->launches++;
  return cudaGetLastError();
}

template <class D>
static int solve_impl(empc_solver* h) {
  if (D::TILE != h->tile) return fail(EMPC_ERR_INVALID, "tile size mismatch");
  h->launches = 0; h->total_iterations = 0;
  for (double& m : h->ms_by_kernel) m = 0;
  for (long long& u : h->units_by_kernel) u = 0;
  CK(cudaEventRecord(h->ev_solve[0], h->stream));
  long long n_act = h->B, n_recalc = h->B;  // OCPs entering the next loop iteration
  init_state_kernel<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->bf, h->P, h->init_feasible, h->nx);
  h->launches++;
  CK(cudaGetLastError());
  // upper bound on batch iterations: every pass of every phase can use maxiter iterations
  int passes = 1;
  for (double c = h->P.convergence_init; c >= h->P.convergence_stop && passes < 64; c *= h->P.convergence_mult) passes++;
  // every pass runs at most maxiter iterations plus one loop that only ends it (computeDirection giving up at reg_max)
  const long long max_loops = (long long)passes * ((long long)h->P.maxiter + 1) + 2;
  bool finished = false;
  const bool trace = std::getenv("EMPC_TRACE_OCP") != nullptr;
  if (h->timing || trace || h->n_groups == 1) {
    // serial schedule with per-kernel CUDA-event timing (used for the roofline numbers and for tracing)
    for (long long it = 0; it < max_loops; ++it) {
      if (h->timing) CK(cudaEventRecord(h->ev[0], h->stream));
      CK(launch_calc_diff<D>(h, 0, 0.0));
      if (h->timing) CK(cudaEventRecord(h->ev[1], h->stream));
      CK(launch_backward<D>(h, 0));
      if (h->timing) CK(cudaEventRecord(h->ev[2], h->stream));
      CK(launch_rollout<D>(h, 0, 0, 0, 0, 0.0));
      if (h->timing) CK(cudaEventRecord(h->ev[3], h->stream));
      CK(cudaMemsetAsync(h->bf.n_active, 0, 2 * sizeof(int), h->stream));
      CK(launch_decide<D>(h, 0));
      // stage B of the line search: no-ops unless some OCP rejected every stage-A step length
      CK(launch_rollout<D>(h, 1, 0, 0, 0, 0.0));
      CK(launch_decide<D>(h, 1));
      if (h->timing) CK(cudaEventRecord(h->ev[4], h->stream));
      CK(cudaMemcpyAsync(h->h_active, h->bf.n_active, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      h->units_by_kernel[0] += n_recalc; h->units_by_kernel[1] += n_act; h->units_by_kernel[2] += n_act; h->units_by_kernel[3] += n_act;
      n_act = h->h_active[0]; n_recalc = h->h_active[1];
      if (h->timing) {
        for (int k = 0; k < 4; ++k) { float ms = 0; CK(cudaEventElapsedTime(&ms, h->ev[k], h->ev[k + 1])); h->ms_by_kernel[k] += ms; }
      }
      if (trace) {
        const int b = std::atoi(std::getenv("EMPC_TRACE_OCP"));
        if (b >= 0 && b < h->B) {
          OcpState s;
          CK(cudaMemcpy(&s, h->bf.st + b, sizeof(s), cudaMemcpyDeviceToHost));
          std::printf("[gpu ph%d] loop=%lld it=%d cost=%.15e prev=%.15e step=%g xreg=%g feas=%d wasf=%d stop=%.6e gap=%.6e acc=%d dg=%.6e dq=%.6e smooth=%g tot=%d\n",
                      s.phase, it, s.iter, s.cost, s.cost_prev, s.steplength, s.xreg, s.is_feasible, s.was_feasible, s.stop, s.gap_inf, s.accepted, s.dg, s.dq, s.smooth, s.total_iters);
        }
      }
      if (*h->h_active == 0) { finished = true; break; }
    }
  } else {
    // pipelined schedule: G batch groups on their own streams, started one backward pass apart
    const int G = h->n_groups;
    Buffers gb[empc_solver::MAXG];
    bool alive[empc_solver::MAXG];
    long long g_act[empc_solver::MAXG], g_recalc[empc_solver::MAXG];
    CK(cudaEventRecord(h->ev_init, h->stream));
    for (int g = 0; g < G; ++g) {
      gb[g] = h->bf;
      gb[g].b0 = (int)((long long)h->B * g / G);
      gb[g].nb = (int)((long long)h->B * (g + 1) / G) - gb[g].b0;
      gb[g].n_active = h->bf.n_active + 2 * g;
      alive[g] = gb[g].nb > 0;
      g_act[g] = g_recalc[g] = gb[g].nb;
      CK(cudaStreamWaitEvent(h->gstream[g], h->ev_init, 0));
    }
    for (long long it = 0; it < max_loops + 1; ++it) {
      bool any = false;
      const int slot = (int)(it & 1);
      for (int g = 0; g < G; ++g) {
        if (!alive[g]) continue;
        any = true;
        cudaStream_t st = h->gstream[g];
        if (it == 0 && g > 0) CK(cudaStreamWaitEvent(st, h->ev_skew[g - 1], 0));
        CK(launch_calc_diff<D>(h, 0, 0.0, &gb[g], st));
        CK(launch_backward<D>(h, 0, &gb[g], st));
        if (it == 0) CK(cudaEventRecord(h->ev_skew[g], st));
        CK(launch_rollout<D>(h, 0, 0, 0, 0, 0.0, &gb[g], st));
        CK(cudaMemsetAsync(gb[g].n_active, 0, 2 * sizeof(int), st));
        CK(launch_decide<D>(h, 0, &gb[g], st));
        CK(launch_rollout<D>(h, 1, 0, 0, 0, 0.0, &gb[g], st));
        CK(launch_decide<D>(h, 1, &gb[g], st));
        CK(cudaMemcpyAsync(h->h_active + (g * 2 + slot) * 2, gb[g].n_active, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(h->ev_it[g][slot], st));
      }
      if (!any) { finished = true; break; }
      if (it >= 1) {  // lagged termination check: look at the counters of the previous batch-iteration
        const int ps = (int)((it - 1) & 1);
        for (int g = 0; g < G; ++g) {
          if (!alive[g]) continue;
          CK(cudaEventSynchronize(h->ev_it[g][ps]));
          const int* c = h->h_active + (g * 2 + ps) * 2;
          h->units_by_kernel[0] += g_recalc[g]; h->units_by_kernel[1] += g_act[g]; h->units_by_kernel[2] += g_act[g]; h->units_by_kernel[3] += g_act[g];
          g_act[g] = c[0]; g_recalc[g] = c[1];
          if (c[0] == 0) alive[g] = false;  // the iteration already queued for this group is a no-op (every OCP exits early)
        }
      }
    }
    for (int g = 0; g < G; ++g) {
      CK(cudaEventRecord(h->ev_gdone[g], h->gstream[g]));
      CK(cudaStreamWaitEvent(h->stream, h->ev_gdone[g], 0));
    }
  }
  CK(launch_squash_out<D>(h));
  CK(cudaEventRecord(h->ev_solve[1], h->stream));
  // total inner iterations over the batch
  std::vector<OcpState> st(h->B);
  CK(cudaMemcpyAsync(st.data(), h->bf.st, sizeof(OcpState) * h->B, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  long long tot = 0;
  for (const OcpState& s : st) tot += s.total_iters;
  h->total_iterations = tot;
  { float ms = 0; CK(cudaEventElapsedTime(&ms, h->ev_solve[0], h->ev_solve[1])); h->solve_ms = ms; }
  if (!finished)  // cannot happen with a terminating schedule (convergence_mult < 1); never report a truncated solve as done
    return fail(EMPC_ERR_INVALID, "solve stopped after " + std::to_string(max_loops) + " batch iterations with OCPs still active (check convergence_mult / maxiter)");
  return EMPC_OK;
}

// empc_solve_stream: the batch-iteration loop of solve_impl (serial schedule) with a refill step after the line search
template <class D>
static int stream_impl(empc_solver* h, int n_jobs, const double* x0, double* xs, double* us, double* us_squash, double* cost,
                       double* stop, int32_t* iters, int32_t* feasible) {
  if (D::TILE != h->tile) return fail(EMPC_ERR_INVALID, "tile size mismatch");
  h->launches = 0; h->total_iterations = 0;
  for (double& m : h->ms_by_kernel) m = 0;
  for (long long& u : h->units_by_kernel) u = 0;
  const size_t J = (size_t)n_jobs, T = h->T, T1 = T + 1, nx = h->nx, nu = h->nu;
  // device staging of the job table and of the per-job results (kept in the handle, grown on demand)
  StreamBuffers sb;
  sb.n_jobs = n_jobs;
  CK(dreuse(h, &h->d_st_x0, &h->cap_st_x0, std::max<size_t>(1, J * nx)));
  CK(dreuse(h, &h->d_st_sc, &h->cap_st_sc, std::max<size_t>(1, 2 * J)));
  CK(dreuse(h, &h->d_st_int, &h->cap_st_int, std::max<size_t>(1, 2 * J) + (size_t)h->B + 1));
  if (xs) CK(dreuse(h, &h->d_st_xs, &h->cap_st_xs, std::max<size_t>(1, J * T1 * nx)));
  if (us) CK(dreuse(h, &h->d_st_us, &h->cap_st_us, std::max<size_t>(1, J * T * nu)));
  if (us_squash) CK(dreuse(h, &h->d_st_uss, &h->cap_st_uss, std::max<size_t>(1, J * T * nu)));
  sb.job_x0 = h->d_st_x0; sb.out_xs = xs ? h->d_st_xs : nullptr; sb.out_us = us ? h->d_st_us : nullptr;
  sb.out_us_squash = us_squash ? h->d_st_uss : nullptr;
  sb.out_cost = h->d_st_sc; sb.out_stop = h->d_st_sc + J;
  sb.out_iters = h->d_st_int; sb.out_feasible = h->d_st_int + J; sb.slot_job = h->d_st_int + 2 * J; sb.queue_next = sb.slot_job + h->B;
  if (J) CK(cudaMemcpyAsync(h->d_st_x0, x0, sizeof(double) * J * nx, cudaMemcpyDefault, h->stream));
  const int first = std::min(n_jobs, h->B);
  CK(cudaMemcpyAsync(sb.queue_next, &first, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CK(cudaEventRecord(h->ev_solve[0], h->stream));
  stream_refill_kernel<D><<<h->B, 128, 0, h->stream>>>(h->bf, sb, h->P, 1);
  h->launches++;
  CK(cudaGetLastError());
  // every job runs at most passes x (maxiter + 1) batch-iterations, and a slot serves ceil(n_jobs / B) jobs at worst... the
  // queue hands jobs out dynamically, so the bound is on the total work of the busiest slot: all jobs
  int passes = 1;
  for (double c = h->P.convergence_init; c >= h->P.convergence_stop && passes < 64; c *= h->P.convergence_mult) passes++;
  const long long per_job = (long long)passes * ((long long)h->P.maxiter + 1) + 2;
  const long long max_loops = per_job * ((long long)(n_jobs + h->B - 1) / std::max(1, h->B) + 1) + 2;
  long long n_act = first, n_recalc = first;
  bool finished = first == 0;
  for (long long it = 0; it < max_loops && !finished; ++it) {
    CK(launch_calc_diff<D>(h, 0, 0.0));
    CK(launch_backward<D>(h, 0));
    CK(launch_rollout<D>(h, 0, 0, 0, 0, 0.0));
    CK(cudaMemsetAsync(h->bf.n_active, 0, 2 * sizeof(int), h->stream));
    CK(launch_decide<D>(h, 0));
    CK(launch_rollout<D>(h, 1, 0, 0, 0, 0.0));
    CK(launch_decide<D>(h, 1));
    stream_refill_kernel<D><<<h->B, 128, 0, h->stream>>>(h->bf, sb, h->P, 0);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->h_active, h->bf.n_active, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->units_by_kernel[0] += n_recalc; h->units_by_kernel[1] += n_act; h->units_by_kernel[2] += n_act; h->units_by_kernel[3] += n_act;
    n_act = h->h_active[0]; n_recalc = h->h_active[1];
    if (n_act == 0) finished = true;
  }
  CK(cudaEventRecord(h->ev_solve[1], h->stream));
  if (!finished) return fail(EMPC_ERR_INVALID, "streaming solve stopped with OCPs still active (check convergence_mult / maxiter)");
  // results -> caller (host or device memory)
  std::vector<int32_t> it_host(J);
  if (J) {
    if (xs) CK(cudaMemcpyAsync(xs, h->d_st_xs, sizeof(double) * J * T1 * nx, cudaMemcpyDefault, h->stream));
    if (us) CK(cudaMemcpyAsync(us, h->d_st_us, sizeof(double) * J * T * nu, cudaMemcpyDefault, h->stream));
    if (us_squash) CK(cudaMemcpyAsync(us_squash, h->d_st_uss, sizeof(double) * J * T * nu, cudaMemcpyDefault, h->stream));
    if (cost) CK(cudaMemcpyAsync(cost, sb.out_cost, sizeof(double) * J, cudaMemcpyDefault, h->stream));
    if (stop) CK(cudaMemcpyAsync(stop, sb.out_stop, sizeof(double) * J, cudaMemcpyDefault, h->stream));
    if (iters) CK(cudaMemcpyAsync(iters, sb.out_iters, sizeof(int32_t) * J, cudaMemcpyDefault, h->stream));
    if (feasible) CK(cudaMemcpyAsync(feasible, sb.out_feasible, sizeof(int32_t) * J, cudaMemcpyDefault, h->stream));
    CK(cudaMemcpyAsync(it_host.data(), sb.out_iters, sizeof(int32_t) * J, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  long long tot = 0;
  for (int32_t v : it_host) tot += (long long)v + 1;  // iter_ = total_iters_ - 1 (src/sbfddp.cpp:222)
  h->total_iterations = tot;
  { float ms = 0; CK(cudaEventElapsedTime(&ms, h->ev_solve[0], h->ev_solve[1])); h->solve_ms = ms; }
  return EMPC_OK;
}

extern "C" {

int empc_solve_stream(empc_solver_t* h, int32_t n_jobs, const double* x0, double* xs, double* us, double* us_squash, double* cost,
                      double* stop, int32_t* iters, int32_t* feasible) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  if (n_jobs < 0 || (n_jobs > 0 && !x0)) return fail(EMPC_ERR_INVALID, "n_jobs < 0 or no initial states");
  if (h->n_node_maps != 1) return fail(EMPC_ERR_INVALID, "streaming solves share one problem: not for replicated MPC instances");
  CK(cudaSetDevice(h->device));
  EMPC_DISPATCH(h, return stream_impl<D>(h, n_jobs, x0, xs, us, us_squash, cost, stop, iters, feasible));
  return EMPC_OK;
}

int empc_solve(empc_solver_t* h) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  CK(cudaSetDevice(h->device));
  EMPC_DISPATCH(h, return solve_impl<D>(h));
  return EMPC_OK;
}

int empc_reset(empc_solver_t* h) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  CK(cudaSetDevice(h->device));
  int rc = load_candidate(h);
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return EMPC_OK;
}

// ---- outputs ---------------------------------------------------------------------------------------------------------
static int d2h(const empc_solver* h, void* dst, const void* src, size_t bytes) {
  if (!h || !dst) return fail(EMPC_ERR_INVALID, "null");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, h->stream));  // dst: host or device memory (unified addressing)
  CK(cudaStreamSynchronize(h->stream));
  return EMPC_OK;
}
int empc_get_xs(const empc_solver_t* h, double* o) { return d2h(h, o, h->bf.xs, sizeof(double) * h->B * (h->T + 1) * h->nx); }
int empc_get_us(const empc_solver_t* h, double* o) { return d2h(h, o, h->bf.us, sizeof(double) * h->B * h->T * h->nu); }
int empc_get_us_squash(const empc_solver_t* h, double* o) { return d2h(h, o, h->bf.us_squash, sizeof(double) * h->B * h->T * h->nu); }
int empc_get_K(const empc_solver_t* h, double* o) { return d2h(h, o, h->bf.K, sizeof(double) * h->B * h->T * h->nu * h->ndx); }
int empc_get_k(const empc_solver_t* h, double* o) { return d2h(h, o, h->bf.k, sizeof(double) * h->B * h->T * h->nu); }
int empc_get_tiles(const empc_solver_t* h, double* o) { return d2h(h, o, h->bf.tiles, sizeof(double) * h->B * (h->T + 1) * h->tile); }
int empc_get_xnext(const empc_solver_t* h, double* o) { return d2h(h, o, h->bf.xnext, sizeof(double) * h->B * (h->T + 1) * h->nx); }
int empc_get_node_cost(const empc_solver_t* h, double* o) { return d2h(h, o, h->bf.node_cost, sizeof(double) * h->B * (h->T + 1)); }
int empc_get_gaps(const empc_solver_t* h, double* o) { return d2h(h, o, h->bf.fs, sizeof(double) * h->B * (h->T + 1) * h->ndx); }
int empc_get_Vx(const empc_solver_t* h, double* o) { return d2h(h, o, h->bf.Vx, sizeof(double) * h->B * (h->T + 1) * h->ndx); }
int empc_get_Vxx_fs(const empc_solver_t* h, double* o) { return d2h(h, o, h->bf.g, sizeof(double) * h->B * (h->T + 1) * h->ndx); }

static int get_states(const empc_solver* h, std::vector<OcpState>& st) {
  st.resize(h->B);
  return d2h(h, st.data(), h->bf.st, sizeof(OcpState) * h->B);
}
int empc_get_cost(const empc_solver_t* h, double* o) {
  std::vector<OcpState> st; int rc = get_states(h, st); if (rc) return rc;
  for (int b = 0; b < h->B; ++b) o[b] = st[b].cost;
  return EMPC_OK;
}
int empc_get_iters(const empc_solver_t* h, int32_t* o) {
  std::vector<OcpState> st; int rc = get_states(h, st); if (rc) return rc;
  for (int b = 0; b < h->B; ++b) o[b] = st[b].iters_out;
  return EMPC_OK;
}
int empc_get_stop(const empc_solver_t* h, double* o) {
  std::vector<OcpState> st; int rc = get_states(h, st); if (rc) return rc;
  for (int b = 0; b < h->B; ++b) o[b] = st[b].stop;
  return EMPC_OK;
}
int empc_get_feasible(const empc_solver_t* h, int32_t* o) {
  std::vector<OcpState> st; int rc = get_states(h, st); if (rc) return rc;
  for (int b = 0; b < h->B; ++b) o[b] = st[b].is_feasible;
  return EMPC_OK;
}
int empc_get_cost_tables(const empc_solver_t* h, empc_cost_t* costs, double* pool) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  int rc = 0;
  if (costs && h->n_costs && (rc = d2h(h, costs, h->d_costs, sizeof(empc_cost_t) * h->n_costs))) return rc;
  if (pool && h->n_pool && (rc = d2h(h, pool, h->d_pool, sizeof(double) * h->n_pool))) return rc;
  return EMPC_OK;
}
int empc_get_solution(empc_solver_t* h, double* xs, double* us, double* us_squash, double* cost, double* stop, int32_t* iters,
                      int32_t* feasible) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  const size_t nxs = (size_t)h->B * (h->T + 1) * h->nx, nus = (size_t)h->B * h->T * h->nu, total = nxs + 2 * nus + 4 * (size_t)h->B;
  if (total * sizeof(double) > (8u << 20)) {  // large batches: plain copies straight into the caller's buffers
    int rc = 0;
    if (xs && (rc = empc_get_xs(h, xs))) return rc;
    if (us && (rc = empc_get_us(h, us))) return rc;
    if (us_squash && (rc = empc_get_us_squash(h, us_squash))) return rc;
    if (cost || stop || iters || feasible) {
      std::vector<OcpState> st; rc = get_states(h, st); if (rc) return rc;
      for (int b = 0; b < h->B; ++b) {
        if (cost) cost[b] = st[b].cost;
        if (stop) stop[b] = st[b].stop;
        if (iters) iters[b] = st[b].iters_out;
        if (feasible) feasible[b] = st[b].is_feasible;
      }
    }
    return EMPC_OK;
  }
  CK(cudaSetDevice(h->device));
  if (!h->d_pack) {
    CK(dalloc(h, &h->d_pack, total));
    CK(cudaMallocHost((void**)&h->h_pack, total * sizeof(double)));
  }
  pack_solution_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 592), 256, 0, h->stream>>>(h->bf, h->d_pack, nxs, nus);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->h_pack, h->d_pack, total * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const double* p = h->h_pack;
  if (xs) std::memcpy(xs, p, nxs * sizeof(double));
  if (us) std::memcpy(us, p + nxs, nus * sizeof(double));
  if (us_squash) std::memcpy(us_squash, p + nxs + nus, nus * sizeof(double));
  const double* sc = p + nxs + 2 * nus;
  for (int b = 0; b < h->B; ++b) {
    if (cost) cost[b] = sc[4 * b];
    if (stop) stop[b] = sc[4 * b + 1];
    if (iters) iters[b] = (int32_t)sc[4 * b + 2];
    if (feasible) feasible[b] = (int32_t)sc[4 * b + 3];
  }
  return EMPC_OK;
}
int empc_get_reg(const empc_solver_t* h, double* o) {
  std::vector<OcpState> st; int rc = get_states(h, st); if (rc) return rc;
  for (int b = 0; b < h->B; ++b) o[b] = st[b].xreg;
  return EMPC_OK;
}
int empc_get_dgdq(const empc_solver_t* h, double* o) {
  std::vector<OcpState> st; int rc = get_states(h, st); if (rc) return rc;
  for (int b = 0; b < h->B; ++b) { o[2 * b] = st[b].dg; o[2 * b + 1] = st[b].dq; }
  return EMPC_OK;
}
int empc_enable_iteration_log(empc_solver_t* h, int32_t capacity) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  if (capacity < 0) return fail(EMPC_ERR_INVALID, "negative log capacity");
  CK(cudaSetDevice(h->device));
  if (capacity == 0) { h->bf.log_cap = 0; return EMPC_OK; }  // the buffer (if any) is kept for the next enable
  CK(dreuse(h, &h->bf.iter_log, &h->cap_iter_log, (size_t)h->B * (size_t)capacity));
  h->bf.log_cap = capacity;
  return EMPC_OK;
}
int empc_get_iteration_log(const empc_solver_t* h, int32_t ocp, empc_iter_record_t* out, int32_t max_records, int32_t* n_records) {
  if (!h || !n_records || (!out && max_records > 0)) return fail(EMPC_ERR_INVALID, "null");
  if (ocp < 0 || ocp >= h->B) return fail(EMPC_ERR_INVALID, "OCP index out of range");
  *n_records = 0;
  if (!h->bf.iter_log || h->bf.log_cap <= 0) return EMPC_OK;
  CK(cudaSetDevice(h->device));
  OcpState st;
  CK(cudaMemcpyAsync(&st, h->bf.st + ocp, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const int cap = h->bf.log_cap, have = std::min(st.log_count, cap), n = std::min(have, (int)max_records);
  if (n <= 0) return EMPC_OK;
  std::vector<empc_iter_record_t> ring((size_t)cap);
  CK(cudaMemcpyAsync(ring.data(), h->bf.iter_log + (size_t)ocp * cap, sizeof(empc_iter_record_t) * cap, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const int first = st.log_count - have;  // oldest record still in the ring
  for (int i = 0; i < n; ++i) out[i] = ring[(size_t)((first + i) % cap)];
  *n_records = n;
  return EMPC_OK;
}
int empc_get_total_iterations(const empc_solver_t* h, int64_t* total) {
  if (!h || !total) return fail(EMPC_ERR_INVALID, "null");
  *total = h->total_iterations;
  return EMPC_OK;
}
int empc_get_launch_stats(const empc_solver_t* h, int64_t* launches, double* ms) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  if (launches) *launches = h->launches;
  if (ms) for (int k = 0; k < 4; ++k) ms[k] = h->ms_by_kernel[k];
  return EMPC_OK;
}
int empc_get_solve_stats(const empc_solver_t* h, double* solve_ms, int64_t* units_by_kernel) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  if (solve_ms) *solve_ms = h->solve_ms;
  if (units_by_kernel) for (int k = 0; k < 4; ++k) units_by_kernel[k] = h->units_by_kernel[k];
  return EMPC_OK;
}
int empc_enable_kernel_timing(empc_solver_t* h, int32_t on) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  h->timing = on ? 1 : 0;
  return EMPC_OK;
}

}  // extern "C"

// ---- plant (closed-loop drivers) ----
template <class D>
static int plant_impl(empc_solver* h, const double* x, const double* u, double dt, double* xnext, int n) {
  if (!h->d_plant_x || h->cap_plant < (size_t)n) {  // persistent staging, grown on demand
    size_t c1 = h->cap_plant, c2 = h->cap_plant, c3 = h->cap_plant;
    CK(dreuse(h, &h->d_plant_x, &c1, (size_t)n * h->nx)); CK(dreuse(h, &h->d_plant_u, &c2, (size_t)n * h->nu));
    CK(dreuse(h, &h->d_plant_xn, &c3, (size_t)n * h->nx));
    h->cap_plant = (size_t)n;
  }
  double *dx_ = h->d_plant_x, *du_ = h->d_plant_u, *dxn_ = h->d_plant_xn;
  CK(cudaMemcpyAsync(dx_, x, sizeof(double) * n * h->nx, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(du_, u, sizeof(double) * n * h->nu, cudaMemcpyHostToDevice, h->stream));
  plant_rk4_kernel<D><<<(n + 63) / 64, 64, 0, h->stream>>>(h->d_model, dx_, du_, dt, dxn_, n, (size_t)h->nu);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(xnext, dxn_, sizeof(double) * n * h->nx, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return EMPC_OK;
}
// device-resident closed loop: x0[b] <- RK4(x0[b], us_squash[b][0], dt)
template <class D>
static int plant_advance_impl(empc_solver* h, double dt, double* x_plant, double* u_applied) {
  plant_rk4_kernel<D><<<(h->B + 63) / 64, 64, 0, h->stream>>>(h->d_model, h->d_x0, h->bf.us_squash, dt, h->d_x0, h->B, (size_t)h->T * h->nu);
  CK(cudaGetLastError());
  if (x_plant) CK(cudaMemcpyAsync(x_plant, h->d_x0, sizeof(double) * h->B * h->nx, cudaMemcpyDeviceToHost, h->stream));
  if (u_applied)
    CK(cudaMemcpy2DAsync(u_applied, sizeof(double) * h->nu, h->bf.us_squash, sizeof(double) * h->T * h->nu, sizeof(double) * h->nu, (size_t)h->B,
                         cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return EMPC_OK;
}
extern "C" {
int empc_plant_advance(empc_solver_t* h, double dt, double* x_plant, double* u_applied) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  if (!(dt > 0)) return fail(EMPC_ERR_INVALID, "dt <= 0");
  CK(cudaSetDevice(h->device));
  EMPC_DISPATCH(h, return plant_advance_impl<D>(h, dt, x_plant, u_applied));
  return EMPC_OK;
}
int empc_plant_step(empc_solver_t* h, const double* x, const double* u, double dt, double* xnext, int32_t n) {
  if (!h || !x || !u || !xnext || n <= 0) return fail(EMPC_ERR_INVALID, "null / bad count");
  CK(cudaSetDevice(h->device));
  EMPC_DISPATCH(h, return plant_impl<D>(h, x, u, dt, xnext, n));
  return EMPC_OK;
}

// ---- tile-level parity hooks ---------------------------------------------------------------------------------------
int empc_phase_calc_diff(empc_solver_t* h, double smooth) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  CK(cudaSetDevice(h->device));
  EMPC_DISPATCH(h, CK(launch_calc_diff<D>(h, 1, smooth)));
  CK(cudaStreamSynchronize(h->stream));
  return EMPC_OK;
}
int empc_phase_backward(empc_solver_t* h, double xreg, int32_t is_feasible, int32_t* ok) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  CK(cudaSetDevice(h->device));
  override_state_kernel<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->bf, xreg, is_feasible, 1);
  CK(cudaGetLastError());
  EMPC_DISPATCH(h, CK(launch_backward<D>(h, 1)));
  CK(cudaStreamSynchronize(h->stream));
  if (ok) {
    std::vector<OcpState> st; int rc = get_states(h, st); if (rc) return rc;
    for (int b = 0; b < h->B; ++b) ok[b] = st[b].bw_fail ? 0 : 1;
  }
  return EMPC_OK;
}
#ifdef EMPC_BW_PROFILE
// diagnostic builds only (scripts/diag/backward_phases.py): clock64() cycles per phase of backward_kernel, block 0
int empc_debug_backward_profile(unsigned long long* out16, int reset) {
  if (out16) CK(cudaMemcpyFromSymbol(out16, g_bw_prof, sizeof(unsigned long long) * 16));
  if (reset) { unsigned long long z[16] = {0}; CK(cudaMemcpyToSymbol(g_bw_prof, z, sizeof(z))); }
  return EMPC_OK;
}
#endif
int empc_phase_rollout(empc_solver_t* h, double smooth, int32_t is_feasible, int32_t ddp) {
  if (!h) return fail(EMPC_ERR_INVALID, "null");
  CK(cudaSetDevice(h->device));
  EMPC_DISPATCH(h, CK(launch_rollout<D>(h, 0, 1, is_feasible, ddp, smooth)));
  EMPC_DISPATCH(h, CK(launch_rollout<D>(h, 1, 1, is_feasible, ddp, smooth)));
  CK(cudaStreamSynchronize(h->stream));
  return EMPC_OK;
}
int empc_get_trial(const empc_solver_t* h, int32_t ai, double* xs_try, double* us_try, double* cost_try, double* dv, int32_t* ok) {
  if (!h || ai < 0 || ai >= EMPC_N_ALPHAS) return fail(EMPC_ERR_INVALID, "bad alpha index");
  const size_t B = h->B, T1 = h->T + 1;
  int rc;
  if (xs_try && (rc = d2h(h, xs_try, h->bf.xs_try + (size_t)ai * B * T1 * h->nx, sizeof(double) * B * T1 * h->nx))) return rc;
  if (us_try && (rc = d2h(h, us_try, h->bf.us_try + (size_t)ai * B * h->T * h->nu, sizeof(double) * B * h->T * h->nu))) return rc;
  std::vector<double> c(B * EMPC_N_ALPHAS), d(B * EMPC_N_ALPHAS);
  std::vector<int> o(B * EMPC_N_ALPHAS);
  if ((rc = d2h(h, c.data(), h->bf.cost_try, sizeof(double) * c.size()))) return rc;
  if ((rc = d2h(h, d.data(), h->bf.dv, sizeof(double) * d.size()))) return rc;
  if ((rc = d2h(h, o.data(), h->bf.ok, sizeof(int) * o.size()))) return rc;
  for (size_t b = 0; b < B; ++b) {
    if (cost_try) cost_try[b] = c[b * EMPC_N_ALPHAS + ai];
    if (dv) dv[b] = d[b * EMPC_N_ALPHAS + ai];
    if (ok) ok[b] = o[b * EMPC_N_ALPHAS + ai];
  }
  return EMPC_OK;
}

}  // extern "C"
