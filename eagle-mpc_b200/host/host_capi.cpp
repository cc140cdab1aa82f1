// host_capi.cpp — plain-C access to the host-side mirror (for the Python front-end / tests; no compute here).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>

#include "eagle_mpc.hpp"
#include "mpc.hpp"

using namespace eagle_mpc;

static thread_local std::string g_err;
#define GUARD(body, failret)            \
  try { body }                          \
  catch (const std::exception& e) { g_err = e.what(); return failret; }
#define GUARD_BEGIN try {
#define GUARD_END(failret) } catch (const std::exception& e) { g_err = e.what(); return failret; }

struct HostTrajectory { std::shared_ptr<Trajectory> traj; };
struct HostFlat { std::shared_ptr<ShootingProblem> problem; FlatProblem flat; };
struct HostSolver { std::shared_ptr<Trajectory> traj; std::shared_ptr<ShootingProblem> problem; std::unique_ptr<SolverSbFDDP> solver; };

extern "C" {

const char* empc_host_last_error(void) { return g_err.c_str(); }
void empc_host_set_dirs(const char* yaml_dir, const char* robot_dir) {
  if (yaml_dir) set_yaml_dir(yaml_dir);
  if (robot_dir) set_robot_data_dir(robot_dir);
}

// ParserYaml -> "key\tvalue\n" dump (caller frees with empc_host_free_str)
char* empc_host_parse_yaml(const char* path) {
  GUARD({
    ParserYaml p(getYamlPath(path));
    std::string out;
    for (auto& kv : p.get_params()) out += kv.first + "\t" + kv.second + "\n";
    char* c = new char[out.size() + 1];
    std::memcpy(c, out.c_str(), out.size() + 1);
    return c;
  }, nullptr)
}
void empc_host_free_str(char* s) { delete[] s; }

void* empc_host_trajectory_create(const char* yaml_path) {
  GUARD({
    auto t = Trajectory::create();
    t->autoSetup(getYamlPath(yaml_path));
    return new HostTrajectory{t};
  }, nullptr)
}
void empc_host_trajectory_free(void* t) { delete (HostTrajectory*)t; }
int empc_host_trajectory_info(void* t, int32_t* out /* nq nv nu n_stages duration n_rotors */) {
  auto& tr = ((HostTrajectory*)t)->traj;
  out[0] = tr->get_robot_model()->nq; out[1] = tr->get_robot_model()->nv; out[2] = (int)tr->get_actuation_nu();
  out[3] = (int)tr->get_stages().size(); out[4] = (int)tr->get_duration(); out[5] = (int)tr->get_platform_params()->n_rotors_;
  return 0;
}
// stage names, one per line, in trajectory order (caller frees with empc_host_free_str)
char* empc_host_trajectory_stage_names(void* t) {
  std::string out;
  for (const auto& st : ((HostTrajectory*)t)->traj->get_stages()) out += st->get_name() + "\n";
  char* c = new char[out.size() + 1];
  std::copy(out.begin(), out.end(), c); c[out.size()] = 0;
  return c;
}
int empc_host_trajectory_platform(void* t, double* tau_f /* 6*n_rotors */, double* u_lb, double* u_ub) {
  auto& pf = ((HostTrajectory*)t)->traj->get_platform_params();
  std::copy(pf->tau_f_.begin(), pf->tau_f_.end(), tau_f);
  std::copy(pf->u_lb.begin(), pf->u_lb.end(), u_lb);
  std::copy(pf->u_ub.begin(), pf->u_ub.end(), u_ub);
  return 0;
}

// createProblem + (optionally) SolverSbFDDP::barrierInit + flatten.  Returns a handle owning the POD description.
void* empc_host_flatten(void* t, int32_t dt_ms, int32_t squash, const char* integrator, int32_t add_barrier) {
  GUARD({
    auto& tr = ((HostTrajectory*)t)->traj;
    std::unique_ptr<HostFlat> hf(new HostFlat());  // released only once nothing can throw any more
    hf->problem = tr->createProblem((std::size_t)dt_ms, squash != 0, integrator);
    if (add_barrier) sbfddp_barrier_init(*hf->problem, tr->get_squash()->get_ns(), 1e-3);
    flatten_problem(*hf->problem, hf->flat);
    return hf.release();
  }, nullptr)
}
void empc_host_flat_free(void* f) { delete (HostFlat*)f; }
const empc_problem_desc_t* empc_host_flat_desc(void* f) { return &((HostFlat*)f)->flat.desc; }
int empc_host_flat_x0(void* f, double* x0) {
  auto& p = ((HostFlat*)f)->problem;
  std::copy(p->x0.begin(), p->x0.end(), x0);
  return 0;
}
// names of the costs of cost set `s`, "\n"-separated, in evaluation order
char* empc_host_flat_cost_names(void* f, int32_t s) {
  auto* hf = (HostFlat*)f;
  std::string out;
  std::vector<std::string> names(hf->flat.slots[s].size());
  const int base = hf->flat.costset_begin[s];
  for (auto& kv : hf->flat.slots[s]) names[kv.second.cost_index - base] = kv.first;
  for (auto& n : names) out += n + "\n";
  char* c = new char[out.size() + 1];
  std::memcpy(c, out.c_str(), out.size() + 1);
  return c;
}

// ---- SolverSbFDDP facade (needs a GPU) ----
void* empc_host_solver_create(void* t, int32_t dt_ms, int32_t squash, const char* integrator, int32_t batch, int32_t device) {
  GUARD({
    auto& tr = ((HostTrajectory*)t)->traj;
    std::unique_ptr<HostSolver> hs(new HostSolver());
    hs->traj = tr;
    hs->problem = tr->createProblem((std::size_t)dt_ms, squash != 0, integrator);
    hs->solver.reset(new SolverSbFDDP(hs->problem, tr->get_squash(), batch, device));
    return hs.release();
  }, nullptr)
}
// trajectory.createProblem(dt, False, integrator) + crocoddyl.SolverBoxFDDP(problem) / SolverBoxDDP(problem)
// (examples/python/trajectory.py:19-24); solver_type = EMPC_SOLVER_BOXFDDP or EMPC_SOLVER_BOXDDP
void* empc_host_box_solver_create(void* t, int32_t dt_ms, int32_t solver_type, const char* integrator, int32_t batch, int32_t device) {
  GUARD({
    auto& tr = ((HostTrajectory*)t)->traj;
    std::unique_ptr<HostSolver> hs(new HostSolver());
    hs->traj = tr;
    hs->problem = tr->createProblem((std::size_t)dt_ms, false, integrator);
    if (solver_type == EMPC_SOLVER_BOXFDDP) hs->solver.reset(new SolverBoxFDDP(hs->problem, batch, device));
    else if (solver_type == EMPC_SOLVER_BOXDDP) hs->solver.reset(new SolverBoxDDP(hs->problem, batch, device));
    else throw std::invalid_argument("solver_type is not a box solver");
    return hs.release();
  }, nullptr)
}
void empc_host_solver_free(void* s) { delete (HostSolver*)s; }
empc_solver_t* empc_host_solver_handle(void* s) { return ((HostSolver*)s)->solver->handle(); }
int empc_host_solver_set_convergence_init(void* s, double c) { ((HostSolver*)s)->solver->set_convergence_init(c); return 0; }
// solve([], [], maxiter) — the reference driver's call (examples/python/trajectory.py:26)
int empc_host_solver_solve(void* s, int32_t maxiter) {
  GUARD({ ((HostSolver*)s)->solver->solve({}, {}, (std::size_t)maxiter); return 0; }, 1)
}
int empc_host_solver_solve_batch(void* s, const double* x0, const double* xs, const double* us, int32_t maxiter) {
  GUARD({ ((HostSolver*)s)->solver->solveBatch(x0, xs, us, (std::size_t)maxiter); return 0; }, 1)
}
int empc_host_solver_result(void* s, double* xs, double* us, double* us_squash, double* cost, int32_t* iter, int32_t* feasible) {
  auto& sv = ((HostSolver*)s)->solver;
  const auto& X = sv->get_xs(); const auto& U = sv->get_us(); const auto& S = sv->getSquashControls();
  for (std::size_t t = 0; t < X.size(); ++t) std::copy(X[t].begin(), X[t].end(), xs + t * X[t].size());
  for (std::size_t t = 0; t < U.size(); ++t) { std::copy(U[t].begin(), U[t].end(), us + t * U[t].size()); std::copy(S[t].begin(), S[t].end(), us_squash + t * S[t].size()); }
  *cost = sv->get_cost(); *iter = (int)sv->get_iter(); *feasible = sv->get_is_feasible() ? 1 : 0;
  return 0;
}

// ---- MPC controllers (src/mpc-controllers/{carrot,rail,weighted}-mpc.cpp).  The entry points after the three
// constructors work on any controller; they keep their historical "carrot" names. ----
struct HostCarrot { std::shared_ptr<Trajectory> traj; std::unique_ptr<MpcAbstract> mpc; };

static std::vector<VectorXd> unpack_states(const double* state_ref, int32_t n_ref, std::size_t nx) {
  std::vector<VectorXd> ref((std::size_t)n_ref);
  for (int i = 0; i < n_ref; ++i) ref[(std::size_t)i].assign(state_ref + (std::size_t)i * nx, state_ref + (std::size_t)(i + 1) * nx);
  return ref;
}

// RailMpc(state_ref, dt_ref, yaml_path): no trajectory object; nx is the row length of state_ref
void* empc_host_rail_create(const double* state_ref, int32_t n_ref, int32_t nx, int32_t dt_ref, const char* yaml_path, int32_t create_solver) {
  GUARD_BEGIN
  std::unique_ptr<HostCarrot> hc(new HostCarrot());
  hc->mpc.reset(new RailMpc(unpack_states(state_ref, n_ref, (std::size_t)nx), (std::size_t)dt_ref, getYamlPath(yaml_path), create_solver != 0));
  return hc.release();
  GUARD_END(nullptr)
}
// WeightedMpc(trajectory, dt_ref, yaml_path): merges the transition stages of `trajectory` in place, like the reference
void* empc_host_weighted_create(void* t, int32_t dt_ref, const char* yaml_path, int32_t create_solver) {
  GUARD_BEGIN
  auto& tr = ((HostTrajectory*)t)->traj;
  std::unique_ptr<HostCarrot> hc(new HostCarrot());
  hc->traj = tr;
  hc->mpc.reset(new WeightedMpc(tr, (std::size_t)dt_ref, getYamlPath(yaml_path), create_solver != 0));
  return hc.release();
  GUARD_END(nullptr)
}

void* empc_host_carrot_create(void* t, const double* state_ref, int32_t n_ref, int32_t dt_ref, const char* yaml_path, int32_t create_solver) {
  GUARD_BEGIN
  auto& tr = ((HostTrajectory*)t)->traj;
  const std::size_t nx = (std::size_t)tr->get_robot_state()->get_nx();
  std::unique_ptr<HostCarrot> hc(new HostCarrot());
  hc->traj = tr;
  hc->mpc.reset(new CarrotMpc(tr, unpack_states(state_ref, n_ref, nx), (std::size_t)dt_ref, getYamlPath(yaml_path), create_solver != 0));
  return hc.release();
  GUARD_END(nullptr)
}
// CarrotMpc stage table (t_stages: n_stages + 1 entries) and transition flags; null arrays => only *n_stages
int empc_host_carrot_schedule(void* m, int32_t* n_stages, int64_t* t_stages, uint8_t* is_transition) {
  GUARD_BEGIN
    auto* c = dynamic_cast<CarrotMpc*>(((HostCarrot*)m)->mpc.get());
    if (!c) throw std::runtime_error("not a CarrotMpc");
    const auto& stages = c->get_trajectory()->get_stages();
    *n_stages = (int32_t)stages.size();
    if (t_stages) {
      for (std::size_t i = 0; i < c->get_t_stages().size(); ++i) t_stages[i] = (int64_t)c->get_t_stages()[i];
      for (std::size_t i = 0; i < stages.size(); ++i) is_transition[i] = stages[i]->get_is_transition() ? 1 : 0;
    }
    return 0;
  GUARD_END(1)
}
// WeightedMpc::schedule() in flat arrays; call once with null arrays for dims = {n_stages, n_slots}
int empc_host_weighted_schedule(void* m, int32_t* dims, int64_t* t_ini, int64_t* t_end, int64_t* duration, double* alpha_beta,
                                uint8_t* match, uint8_t* task, double* base) {
  GUARD_BEGIN
    auto* w = dynamic_cast<WeightedMpc*>(((HostCarrot*)m)->mpc.get());
    if (!w) throw std::runtime_error("not a WeightedMpc");
    const WeightedSchedule s = w->schedule();
    dims[0] = (int32_t)s.t_ini.size(); dims[1] = (int32_t)s.n_slots;
    if (t_ini) {
      std::copy(s.t_ini.begin(), s.t_ini.end(), t_ini); std::copy(s.t_end.begin(), s.t_end.end(), t_end);
      *duration = s.duration; alpha_beta[0] = s.alpha; alpha_beta[1] = s.beta;
      std::copy(s.match.begin(), s.match.end(), match); std::copy(s.task.begin(), s.task.end(), task);
      std::copy(s.base.begin(), s.base.end(), base);
    }
    return 0;
  GUARD_END(1)
}
void empc_host_carrot_free(void* m) { delete (HostCarrot*)m; }
int empc_host_carrot_info(void* m, int32_t* out /* knots dt iters n_costs n_pool */) {
  auto& c = ((HostCarrot*)m)->mpc;
  out[0] = (int)c->get_knots(); out[1] = (int)c->get_dt(); out[2] = (int)c->get_iters();
  out[3] = (int)c->flat().costs.size(); out[4] = (int)c->flat().pool.size();
  return 0;
}
const empc_problem_desc_t* empc_host_carrot_desc(void* m) { return &((HostCarrot*)m)->mpc->flat().desc; }
empc_solver_t* empc_host_carrot_handle(void* m) {
  auto& s = ((HostCarrot*)m)->mpc->get_solver();
  return s ? s->handle() : nullptr;
}
int empc_host_carrot_update(void* m, int32_t t_ms) {
  GUARD({ ((HostCarrot*)m)->mpc->updateProblem((std::size_t)t_ms); return 0; }, 1)
}
// current cost records / pool (after updateProblem), for the tests' oracle
int empc_host_carrot_costs(void* m, empc_cost_t* costs, double* pool) {
  FlatProblem& f = ((HostCarrot*)m)->mpc->flat();
  std::copy(f.costs.begin(), f.costs.end(), costs);
  std::copy(f.pool.begin(), f.pool.end(), pool);
  return 0;
}
// problem.x0 = x0 ; solver.solve(xs, us, maxiter) with explicit warm start (NULL xs => warm start from the previous solution)
int empc_host_carrot_solve(void* m, const double* x0, const double* xs, const double* us, int32_t maxiter, double convergence_init) {
  GUARD_BEGIN
  auto& c = ((HostCarrot*)m)->mpc;
  auto& sv = c->get_solver();
  if (!sv) throw std::runtime_error("the MPC controller was created without a solver");
  const std::size_t nx = (std::size_t)c->get_robot_state()->get_nx(), T = c->get_problem()->get_T(), nu = c->get_squash()->get_ns();
  c->get_problem()->set_x0(VectorXd(x0, x0 + nx));
  sv->set_convergence_init(convergence_init);
  if (xs && us) {
    std::vector<VectorXd> X(T + 1);
    std::vector<VectorXd> U(T);
    for (std::size_t t = 0; t <= T; ++t) X[t].assign(xs + t * nx, xs + (t + 1) * nx);
    for (std::size_t t = 0; t < T; ++t) U[t].assign(us + t * nu, us + (t + 1) * nu);
    sv->solve(X, U, (std::size_t)maxiter);
  } else {
    sv->solveWarm((std::size_t)maxiter);
  }
  return 0;
  GUARD_END(1)
}
int empc_host_carrot_result(void* m, double* xs, double* us, double* us_squash, double* cost, int32_t* iter) {
  auto& sv = ((HostCarrot*)m)->mpc->get_solver();
  const auto& X = sv->get_xs(); const auto& U = sv->get_us(); const auto& S = sv->getSquashControls();
  for (std::size_t t = 0; t < X.size(); ++t) std::copy(X[t].begin(), X[t].end(), xs + t * X[t].size());
  for (std::size_t t = 0; t < U.size(); ++t) { std::copy(U[t].begin(), U[t].end(), us + t * U[t].size()); std::copy(S[t].begin(), S[t].end(), us_squash + t * S[t].size()); }
  *cost = sv->get_cost(); *iter = (int)sv->get_iter();
  return 0;
}

}  // extern "C"
