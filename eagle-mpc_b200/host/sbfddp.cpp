// sbfddp.cpp — SolverSbFDDP facade: same constructor / solve / setCandidate / getters as the reference
// (include/eagle_mpc/sbfddp.hpp:39-52, src/sbfddp.cpp), with the whole iteration running in CUDA behind the C ABI.
#include <cstring>

#include <cstdio>

#include "cuda_abi.hpp"
#include "eagle_mpc.hpp"

namespace eagle_mpc {

static void ck(int rc, const char* what) {
  if (rc != EMPC_OK) throw std::runtime_error(std::string(what) + ": " + cuda_abi().last_error());
}

SolverSbFDDP::SolverSbFDDP(const std::shared_ptr<ShootingProblem>& problem,
                           const std::shared_ptr<SquashingModelSmoothSat>& squashing_model, int batch, int device)
    : problem_(problem), squashing_model_(squashing_model), batch_(batch), device_(device) {
  if (!squashing_model_) throw std::invalid_argument("SolverSbFDDP needs the squashing model (src/sbfddp.cpp:5)");
  init();
}

SolverSbFDDP::SolverSbFDDP(const std::shared_ptr<ShootingProblem>& problem,
                           const std::shared_ptr<SquashingModelSmoothSat>& squashing_model, int batch, int device, int solver_type)
    : problem_(problem), squashing_model_(squashing_model), batch_(batch), device_(device), solver_type_(solver_type) {
  init();
}

void SolverSbFDDP::init() {
  if (solver_type_ == EMPC_SOLVER_SBFDDP) {
    cuda_abi().default_params(&params_);
    barrierInit();
  } else {
    // the box solvers take the problem as it is (createProblem(dt, squash = false, ...): plain actuation, control limits on
    // every model, src/trajectory.cpp:96-100,131-132)
    if (problem_->terminalModel->squash)
      throw std::invalid_argument("SolverBoxFDDP / SolverBoxDDP: the problem was created with the squashing actuation (createProblem(dt, squash = false))");
    if (!cuda_abi().box_params) throw std::runtime_error("the loaded CUDA library has no Box solvers (empc_box_params is missing)");
    cuda_abi().box_params(&params_, solver_type_);
  }
  flatten_problem(*problem_, flat_);
  ck(cuda_abi().create(&flat_.desc, batch_, device_, &handle_), "empc_create");
  ck(cuda_abi().set_params(handle_, &params_), "empc_set_params");
  const std::size_t T = problem_->get_T();
  const std::size_t ndx = (std::size_t)problem_->state->get_ndx();
  nu_ = squashing_model_ ? squashing_model_->get_ns() : (std::size_t)flat_.desc.n_rotors + ndx / 2 - 6;
  const std::size_t nu = nu_;
  xs_.assign(T + 1, problem_->state->zero());
  us_.assign(T, VectorXd(nu, 0.0));
  us_squash_.assign(T, VectorXd(nu, 0.0));
  k_.assign(T, VectorXd(nu, 0.0));
  K_.assign(T, std::vector<double>(nu * ndx, 0.0));
  syncX0();
}

SolverSbFDDP::~SolverSbFDDP() { if (handle_) cuda_abi().destroy(handle_); }

// src/sbfddp.cpp:169-190: add the "barrier" cost (weight 1e-3) to every running model that does not have it yet.
// One shared cost object; its activation weights follow the smoothing schedule on the device.
void sbfddp_barrier_init(ShootingProblem& problem, std::size_t ns, double barrier_weight) {
  auto barrier = std::make_shared<CostModelResidual>();
  barrier->type = CostModelTypes::CostModelSquashBarrier;
  barrier->activation.type = ActivationModelTypes::ActivationModelWeightedQuadraticBarrier;
  barrier->activation.nr = ns;
  for (auto& m : problem.runningModels) {
    auto& costs = m->costs->get_costs();
    if (costs.find("barrier") == costs.end()) m->costs->addCost("barrier", barrier, barrier_weight);
  }
}
void SolverSbFDDP::barrierInit() { sbfddp_barrier_init(*problem_, squashing_model_->get_ns(), params_.barrier_weight); }

void SolverSbFDDP::syncX0() {
  const std::size_t nx = problem_->x0.size();
  std::vector<double> x0((std::size_t)batch_ * nx);
  for (int b = 0; b < batch_; ++b) std::copy(problem_->x0.begin(), problem_->x0.end(), x0.begin() + (std::size_t)b * nx);
  ck(cuda_abi().set_x0(handle_, x0.data()), "empc_set_x0");
}

void SolverSbFDDP::pushCosts(int first, int n) {
  ck(cuda_abi().update_costs(handle_, first, n, flat_.costs.data() + first, 0, (int)flat_.pool.size(), flat_.pool.data()), "empc_update_costs");
}
void SolverSbFDDP::pushAllCosts() { pushCosts(0, (int)flat_.costs.size()); }

void SolverSbFDDP::setCandidate(const std::vector<VectorXd>& xs_warm, const std::vector<VectorXd>& us_warm, bool is_feasible) {
  const std::size_t T = problem_->get_T(), nx = (std::size_t)problem_->state->get_nx(), nu = nu_;
  std::vector<double> xs, us;
  if (!xs_warm.empty()) {
    if (xs_warm.size() != T + 1)
      throw std::invalid_argument("Warm start state has wrong dimension, got " + std::to_string(xs_warm.size()) + " expecting " + std::to_string(T + 1));
    xs.resize((std::size_t)batch_ * (T + 1) * nx);
    for (int b = 0; b < batch_; ++b)
      for (std::size_t t = 0; t <= T; ++t) std::copy(xs_warm[t].begin(), xs_warm[t].end(), xs.begin() + ((std::size_t)b * (T + 1) + t) * nx);
  }
  if (!us_warm.empty()) {
    if (us_warm.size() != T)
      throw std::invalid_argument("Warm start control has wrong dimension, got " + std::to_string(us_warm.size()) + " expecting " + std::to_string(T));
    us.resize((std::size_t)batch_ * T * nu);
    for (int b = 0; b < batch_; ++b)
      for (std::size_t t = 0; t < T; ++t) std::copy(us_warm[t].begin(), us_warm[t].end(), us.begin() + ((std::size_t)b * T + t) * nu);
  }
  ck(cuda_abi().set_candidate(handle_, xs.empty() ? nullptr : xs.data(), us.empty() ? nullptr : us.data(), is_feasible ? 1 : 0), "empc_set_candidate");
}

void SolverSbFDDP::fetch(bool with_gains) {
  const std::size_t T = problem_->get_T(), nx = (std::size_t)problem_->state->get_nx(), nu = nu_,
                    ndx = (std::size_t)problem_->state->get_ndx(), B = (std::size_t)batch_;
  std::vector<double> buf(B * (T + 1) * nx), ub(B * T * nu), sb(B * T * nu), c(B), s(B);
  std::vector<int32_t> it(B), fe(B);
  ck(cuda_abi().get_solution(handle_, buf.data(), ub.data(), sb.data(), c.data(), s.data(), it.data(), fe.data()), "empc_get_solution");
  for (std::size_t t = 0; t <= T; ++t) xs_[t].assign(buf.begin() + t * nx, buf.begin() + (t + 1) * nx);
  for (std::size_t t = 0; t < T; ++t) us_[t].assign(ub.begin() + t * nu, ub.begin() + (t + 1) * nu);
  for (std::size_t t = 0; t < T; ++t) us_squash_[t].assign(sb.begin() + t * nu, sb.begin() + (t + 1) * nu);
  buf.resize(B * T * nu);
  if (with_gains) {
    ck(cuda_abi().get_k(handle_, buf.data()), "empc_get_k");
    for (std::size_t t = 0; t < T; ++t) k_[t].assign(buf.begin() + t * nu, buf.begin() + (t + 1) * nu);
    buf.resize(B * T * nu * ndx);
    ck(cuda_abi().get_K(handle_, buf.data()), "empc_get_K");
    for (std::size_t t = 0; t < T; ++t) K_[t].assign(buf.begin() + t * nu * ndx, buf.begin() + (t + 1) * nu * ndx);
  }
  cost_ = c[0]; stop_ = s[0]; iter_ = (std::size_t)it[0]; is_feasible_ = fe[0] != 0;
}

// crocoddyl::CallbackVerbose, fed from the device-side iteration log
void CallbackVerbose::operator()(const empc_iter_record_t& r) {
  if (r.total_iter % 10 == 0) std::printf("iter \t cost \t      stop \t    grad \t  xreg \t      ureg \t step \t feas\n");
  std::printf("%4d  %.5e  %.5e  %.5e  %.5e  %.5e   %.4f     %d\n", r.total_iter, r.cost, r.stop, -r.d1, r.xreg, r.xreg, r.steplength,
              r.is_feasible);
}

void SolverSbFDDP::setCallbacks(const std::vector<std::shared_ptr<CallbackAbstract>>& callbacks) {
  callbacks_ = callbacks;
  const int passes = 8;  // FDDP passes + DDP clean-up of one solve() (2 + 1 with the default convergence schedule)
  ck(cuda_abi().enable_iteration_log(handle_, callbacks_.empty() ? 0 : passes * (params_.maxiter + 1)), "empc_enable_iteration_log");
}

// The reference runs its callbacks inside the iteration loop (src/sbfddp.cpp:303-307); here the loop runs on the device and
// leaves one record per iteration, which are replayed through the callbacks in order once the solve has returned.
void SolverSbFDDP::replayCallbacks() {
  if (callbacks_.empty()) return;
  std::vector<empc_iter_record_t> rec((std::size_t)8 * (std::size_t)(params_.maxiter + 1));
  int32_t n = 0;
  ck(cuda_abi().get_iteration_log(handle_, 0, rec.data(), (int32_t)rec.size(), &n), "empc_get_iteration_log");
  for (int32_t i = 0; i < n; ++i)
    for (auto& cb : callbacks_) (*cb)(rec[(std::size_t)i]);
}

bool SolverSbFDDP::solve(const std::vector<VectorXd>& init_xs, const std::vector<VectorXd>& init_us, std::size_t maxiter,
                         bool is_feasible, double /*regInit: ignored, src/sbfddp.cpp:196 vs :210*/) {
  syncX0();
  setCandidate(init_xs, init_us, is_feasible);
  params_.maxiter = (int)maxiter;
  ck(cuda_abi().set_params(handle_, &params_), "empc_set_params");
  ck(cuda_abi().solve(handle_), "empc_solve");
  fetch(true);
  replayCallbacks();
  return true;  // the reference always returns true (src/sbfddp.cpp:225)
}

bool SolverSbFDDP::solveWarm(std::size_t maxiter) {
  syncX0();
  params_.maxiter = (int)maxiter;
  ck(cuda_abi().set_params(handle_, &params_), "empc_set_params");
  ck(cuda_abi().solve(handle_), "empc_solve");
  fetch(false);
  replayCallbacks();
  return true;
}

bool SolverSbFDDP::solveBatch(const double* x0, const double* xs, const double* us, std::size_t maxiter, bool is_feasible) {
  if (x0) ck(cuda_abi().set_x0(handle_, x0), "empc_set_x0");
  ck(cuda_abi().set_candidate(handle_, xs, us, is_feasible ? 1 : 0), "empc_set_candidate");
  params_.maxiter = (int)maxiter;
  ck(cuda_abi().set_params(handle_, &params_), "empc_set_params");
  ck(cuda_abi().solve(handle_), "empc_solve");
  fetch(false);
  replayCallbacks();
  return true;
}

}  // namespace eagle_mpc
