// mpc.cpp — MpcAbstract + CarrotMpc / RailMpc / WeightedMpc (src/mpc-base.cpp, src/mpc-controllers/*.cpp),
// bug-compatible where the reference's behaviour is observable: integer-division interpolation
// (piecewise-constant reference), t_stages clamped to >= dt, carrot_tail never deactivated, dimension checks that
// never throw, prefix matching of stage names, the hover quaternion of RailMpc that keeps the x / y components.
#include <algorithm>
#include <cmath>

#include "mpc.hpp"

namespace eagle_mpc {

MpcAbstract::MpcAbstract(const std::string& yaml_path) {
  ParserYaml parser(yaml_path);
  params_server_ = std::make_shared<ParamsServer>(parser.get_params());
  initializeRobotObjects();
  loadParams();
  int_models_.reserve(params_.knots);
}

void MpcAbstract::initializeRobotObjects() {
  robot_model_path_ = getUrdfPath(params_server_->getParam<std::string>("robot/urdf"));
  robot_model_ = buildModelFromUrdf(robot_model_path_);
  platform_params_ = std::make_shared<MultiCopterBaseParams>();
  platform_params_->autoSetup("robot/platform/", params_server_, robot_model_);
  robot_state_ = std::make_shared<StateMultibody>();
  robot_state_->pinocchio = robot_model_;
  nu_ = platform_params_->n_rotors_ + (std::size_t)(robot_model_->nv - 6);
  squash_ = std::make_shared<SquashingModelSmoothSat>();
  squash_->u_lb = platform_params_->u_lb; squash_->u_ub = platform_params_->u_ub; squash_->ns = nu_;
}

void MpcAbstract::loadParams() {
  const std::string p = "mpc_controller/";
  params_.integrator_type = params_server_->getParam<std::string>(p + "integration_method");
  if (params_.integrator_type != "IntegratedActionModelEuler" && params_.integrator_type != "IntegratedActionModelRK4")
    throw std::out_of_range("map::at");
  params_.knots = (std::size_t)params_server_->getParam<int>(p + "knots");
  params_.iters = (std::size_t)params_server_->getParam<int>(p + "iters");
  params_.dt = (std::size_t)params_server_->getParam<int>(p + "dt");
  const std::string solver = params_server_->getParam<std::string>(p + "solver");
  if (solver == "SolverSbFDDP") params_.solver_type = SolverTypes::SolverSbFDDP;
  else if (solver == "SolverBoxFDDP") params_.solver_type = SolverTypes::SolverBoxFDDP;
  else if (solver == "SolverBoxDDP") params_.solver_type = SolverTypes::SolverBoxDDP;
  else throw std::out_of_range("map::at");
  try { params_.callback = params_server_->getParam<bool>(p + "callback"); }
  catch (const std::exception&) { params_.callback = false; }
}

void MpcAbstract::checkHotPathSupport() const {
  // all three SolverTypes run on the device (SolverBoxFDDP / SolverBoxDDP: crocoddyl's box solvers, csrc/backward.cuh BOX)
}

std::shared_ptr<ActionModel> MpcAbstract::makeKnotModel(const std::shared_ptr<CostModelSum>& costs) const {
  auto iam = std::make_shared<ActionModel>();
  iam->costs = costs;
  iam->dt = double(params_.dt) / 1000.;
  iam->rk4 = params_.integrator_type == "IntegratedActionModelRK4";  // src/mpc-controllers/carrot-mpc.cpp:211-222
  // src/mpc-controllers/carrot-mpc.cpp:188-193: the squashing actuation only under SolverSbFDDP
  iam->squash = params_.solver_type == SolverTypes::SolverSbFDDP;
  iam->u_lb = platform_params_->u_lb; iam->u_ub = platform_params_->u_ub;
  return iam;
}

void MpcAbstract::finishProblem() {
  problem_ = std::make_shared<ShootingProblem>();
  problem_->x0 = robot_state_->zero();
  problem_->runningModels.assign(int_models_.begin(), int_models_.end() - 1);
  problem_->terminalModel = int_models_.back();
  problem_->state = robot_state_;
  problem_->platform = platform_params_;
  if (!defer_solver_) attachSolver();
  else {
    if (params_.solver_type == SolverTypes::SolverSbFDDP) sbfddp_barrier_init(*problem_, squash_->get_ns(), 1e-3);
    flatten_problem(*problem_, flat_local_);
  }
}

void MpcAbstract::attachSolver() {
  if (solver_) return;
  switch (params_.solver_type) {  // src/mpc-controllers/carrot-mpc.cpp:232-242
    case SolverTypes::SolverSbFDDP: solver_ = std::make_shared<SolverSbFDDP>(problem_, squash_, 1, 0); break;
    case SolverTypes::SolverBoxFDDP: solver_ = std::make_shared<SolverBoxFDDP>(problem_, 1, 0); break;
    case SolverTypes::SolverBoxDDP: solver_ = std::make_shared<SolverBoxDDP>(problem_, 1, 0); break;
  }
}

FlatProblem& MpcAbstract::flat() { return solver_ ? solver_->flat() : flat_local_; }

// copy one cost item of knot `knot` into the flat tables and remember the dirty range
void MpcAbstract::syncCost(std::size_t knot, const std::string& name) {
  FlatProblem& fl = flat();
  const int set = fl.node_costset[knot];
  const FlatProblem::Slot& sl = fl.slots[set].at(name);
  fill_cost_record(*int_models_[knot]->costs->get_costs().at(name), fl.costs[sl.cost_index], fl.pool.data());
  dirty_lo_ = std::min(dirty_lo_, sl.cost_index); dirty_hi_ = std::max(dirty_hi_, sl.cost_index);
}

void MpcAbstract::endUpdate() {
  if (solver_ && dirty_hi_ >= dirty_lo_) solver_->pushCosts(dirty_lo_, dirty_hi_ - dirty_lo_ + 1);
}

// ---- CarrotMpc --------------------------------------------------------------------------------------------------------
CarrotMpc::CarrotMpc(const std::shared_ptr<Trajectory>& trajectory, const std::vector<VectorXd>& state_ref, std::size_t dt_ref,
                     const std::string& yaml_path, bool create_solver)
    : MpcAbstract(yaml_path), trajectory_(trajectory) {
  defer_solver_ = !create_solver;
  state_ref_ = state_ref;
  for (std::size_t i = 0; i < state_ref_.size(); ++i) t_ref_.push_back(dt_ref * i);
  loadCostParams();
  const auto& stages = trajectory_->get_stages();
  t_stages_.reserve(stages.size() + 1);
  t_stages_.push_back(0);
  for (std::size_t i = 1; i < stages.size(); ++i) {
    const std::size_t duration = stages[i - 1]->get_duration() <= params_.dt ? params_.dt : stages[i - 1]->get_duration();
    t_stages_.push_back(t_stages_.back() + duration);
  }
  const std::size_t duration = stages.back()->get_duration() <= params_.dt ? params_.dt : stages.back()->get_duration();
  t_stages_.push_back(t_stages_.back() + duration);
  createProblem();
  update_vars_.state_ref = robot_state_->zero();
}

void CarrotMpc::loadCostParams() {
  auto dbl = [&](const char* key, double def) {
    try { return params_server_->getParam<double>(std::string("mpc_controller/") + key); }
    catch (const std::exception&) { return def; }
  };
  auto vec = [&](const char* key, std::size_t n) {
    try { return converter<VectorXd>::convert(params_server_->getParam<std::string>(std::string("mpc_controller/") + key)); }
    catch (const std::exception&) { return VectorXd(n, 1.0); }
  };
  const std::size_t ndx = (std::size_t)robot_state_->get_ndx();
  carrot_weight_ = dbl("carrot_weight", 10.0);
  carrot_tail_weight_ = dbl("carrot_tail_weight", 5.0);
  carrot_tail_act_weights_ = vec("carrot_tail_act_weights", ndx);
  control_reg_weight_ = dbl("carrot_control_reg_weight", 1e-2);
  control_reg_act_weights_ = vec("carrot_control_reg_act_weights", nu_);
  state_reg_weight_ = dbl("carrot_state_reg_weight", 1e-3);
  state_ref_act_weights_ = vec("carrot_state_ref_act_weights", ndx);
  state_limits_weight_ = dbl("carrot_state_limits_weight", 100);
  state_limits_act_weights_ = vec("carrot_state_limits_act_weights", ndx);
  // mandatory (no try/catch in the reference, :162-171); the size checks there construct an exception without throwing
  state_limits_l_bound_ = converter<VectorXd>::convert(params_server_->getParam<std::string>("mpc_controller/carrot_state_limits_l_bound"));
  state_limits_u_bound_ = converter<VectorXd>::convert(params_server_->getParam<std::string>("mpc_controller/carrot_state_limits_u_bound"));
  for (const VectorXd* v : {&carrot_tail_act_weights_, &state_ref_act_weights_, &state_limits_act_weights_, &state_limits_l_bound_, &state_limits_u_bound_})
    if (v->size() != ndx) throw std::runtime_error("CarrotMPC: a state-sized parameter vector has dimension " + std::to_string(v->size()) +
                                                   ", should be " + std::to_string(ndx) + " (the reference would read out of bounds here)");
  if (control_reg_act_weights_.size() != nu_) throw std::runtime_error("CarrotMPC: control regularization weights have the wrong dimension");
}

std::shared_ptr<CostModelSum> CarrotMpc::createCosts() const {
  auto costs = std::make_shared<CostModelSum>();
  const std::size_t ndx = (std::size_t)robot_state_->get_ndx();
  auto state_cost = [&](ActivationModelTypes t, const VectorXd& w) {
    auto c = std::make_shared<CostModelResidual>();
    c->type = CostModelTypes::CostModelState;
    c->activation.type = t; c->activation.nr = ndx; c->activation.weights = w;
    c->reference = robot_state_->zero();
    return c;
  };
  costs->addCost("state_reg", state_cost(ActivationModelTypes::ActivationModelWeightedQuad, state_ref_act_weights_), state_reg_weight_, true);
  auto control = std::make_shared<CostModelResidual>();
  control->type = CostModelTypes::CostModelControl;
  control->activation.type = ActivationModelTypes::ActivationModelWeightedQuad; control->activation.nr = nu_;
  control->activation.weights = control_reg_act_weights_;
  control->reference.assign(nu_, 0.0);
  costs->addCost("control_reg", control, control_reg_weight_, true);
  auto limits = state_cost(ActivationModelTypes::ActivationModelWeightedQuadraticBarrier, state_limits_act_weights_);
  limits->activation.lb = state_limits_l_bound_; limits->activation.ub = state_limits_u_bound_;
  for (std::size_t i = 0; i < ndx; ++i) {  // crocoddyl::ActivationBounds(lb, ub, 1)
    const double m = 0.5 * (limits->activation.lb[i] + limits->activation.ub[i]), d = 0.5 * (limits->activation.ub[i] - limits->activation.lb[i]);
    limits->activation.lb[i] = m - 1.0 * d; limits->activation.ub[i] = m + 1.0 * d;
  }
  costs->addCost("state_limits", limits, state_limits_weight_, true);
  costs->addCost("carrot_state", state_cost(ActivationModelTypes::ActivationModelQuad, VectorXd()), carrot_weight_, false);
  costs->addCost("carrot_tail", state_cost(ActivationModelTypes::ActivationModelWeightedQuad, carrot_tail_act_weights_), carrot_tail_weight_, false);
  return costs;
}

void CarrotMpc::createProblem() {
  if (trajectory_->get_has_contact()) throw std::runtime_error("Carrot with contact has not been implemented");
  checkHotPathSupport();
  for (std::size_t i = 0; i < params_.knots; ++i) int_models_.push_back(makeKnotModel(createCosts()));  // one model per knot (:195-225)
  finishProblem();
}

void CarrotMpc::computeActiveStage(std::size_t t) {
  update_vars_.idx_stage = std::size_t(std::upper_bound(t_stages_.begin(), t_stages_.end(), t) - t_stages_.begin()) - 1;
}

void CarrotMpc::computeStateReference(std::size_t time) {
  update_vars_.idx_state = std::size_t(std::upper_bound(t_ref_.begin(), t_ref_.end(), time) - t_ref_.begin());
  const std::size_t nq = (std::size_t)robot_state_->get_nq();
  if (update_vars_.idx_state >= state_ref_.size()) {
    update_vars_.state_ref = robot_state_->zero();
    std::copy(state_ref_.back().begin(), state_ref_.back().begin() + (long)nq, update_vars_.state_ref.begin());
  } else {
    // alpha = (time - t_ref[i-1]) / (t_ref[i] - t_ref[i-1]) in integer arithmetic == 0  (:391-392): piecewise constant
    update_vars_.state_ref = state_ref_[update_vars_.idx_state - 1];
  }
}

void CarrotMpc::updateFreeCosts(std::size_t idx) {
  auto& costs = int_models_[idx]->costs->get_costs();
  const auto& stages = trajectory_->get_stages();
  if (update_vars_.idx_stage < stages.size()) {
    if (!stages[update_vars_.idx_stage]->get_is_transition() || (idx == int_models_.size() - 1)) {
      costs.at("carrot_state")->active = true;
      computeStateReference(update_vars_.node_time);
      costs.at("carrot_state")->cost->reference = update_vars_.state_ref;
    } else {
      costs.at("carrot_state")->active = false;
    }
    syncCost(idx, "carrot_state");
  } else {
    costs.at("carrot_state")->active = false;
    costs.at("carrot_tail")->active = true;
    computeStateReference(update_vars_.node_time);
    costs.at("carrot_tail")->cost->reference = update_vars_.state_ref;
    syncCost(idx, "carrot_state");
    syncCost(idx, "carrot_tail");
  }
}

void CarrotMpc::updateProblem(const std::size_t& current_time) {
  computeActiveStage(current_time);
  update_vars_.idx_last_stage = update_vars_.idx_stage;
  beginUpdate();
  for (std::size_t i = 0; i < int_models_.size(); ++i) {
    update_vars_.node_time = current_time + i * params_.dt;
    computeActiveStage(update_vars_.node_time);
    updateFreeCosts(i);
    update_vars_.idx_last_stage = update_vars_.idx_stage;
  }
  endUpdate();
}

// ---- RailMpc (src/mpc-controllers/rail-mpc.cpp) -------------------------------------------------------------------------
RailMpc::RailMpc(const std::vector<VectorXd>& state_ref, std::size_t dt_ref, const std::string& yaml_path, bool create_solver)
    : MpcAbstract(yaml_path) {
  defer_solver_ = !create_solver;
  state_ref_ = state_ref;
  for (std::size_t i = 0; i < state_ref_.size(); ++i) t_ref_.push_back(dt_ref * i);
  const std::size_t ndx = (std::size_t)robot_state_->get_ndx();
  try { state_weight_ = params_server_->getParam<double>("mpc_controller/rail_weight"); }
  catch (const std::exception&) { state_weight_ = 10; }
  try { state_activation_weights_ = converter<VectorXd>::convert(params_server_->getParam<std::string>("mpc_controller/rail_activation_weights")); }
  catch (const std::exception&) { state_activation_weights_ = VectorXd(ndx, 1.0); }
  // (:41-45 constructs a std::runtime_error without throwing it; a wrong size would be read out of bounds there)
  if (state_activation_weights_.size() != ndx)
    throw std::runtime_error("RailMPC: the dimension for the state activation weights vector is " +
                             std::to_string(state_activation_weights_.size()) + ", should be " + std::to_string(ndx));
  try { control_weight_ = params_server_->getParam<double>("mpc_controller/rail_control_weight"); }
  catch (const std::exception&) { control_weight_ = 1e-1; }
  createProblem();
  update_vars_.state_ref = robot_state_->zero();
}

std::shared_ptr<CostModelSum> RailMpc::createCosts() const {
  auto costs = std::make_shared<CostModelSum>();
  auto rail = std::make_shared<CostModelResidual>();
  rail->type = CostModelTypes::CostModelState;
  rail->activation.type = ActivationModelTypes::ActivationModelWeightedQuad;
  rail->activation.nr = (std::size_t)robot_state_->get_ndx();
  rail->activation.weights = state_activation_weights_;
  rail->reference = robot_state_->zero();
  costs->addCost("rail_state", rail, state_weight_, true);
  auto control = std::make_shared<CostModelResidual>();  // CostModelResidual(state, ResidualModelControl): ActivationModelQuad
  control->type = CostModelTypes::CostModelControl;
  control->activation.type = ActivationModelTypes::ActivationModelQuad; control->activation.nr = nu_;
  control->reference.assign(nu_, 0.0);
  costs->addCost("control", control, control_weight_, true);
  return costs;
}

void RailMpc::createProblem() {
  checkHotPathSupport();
  for (std::size_t i = 0; i < params_.knots; ++i) int_models_.push_back(makeKnotModel(createCosts()));
  finishProblem();
}

void RailMpc::computeStateReference(std::size_t time) {
  update_vars_.idx_state = std::size_t(std::upper_bound(t_ref_.begin(), t_ref_.end(), time) - t_ref_.begin());
  const std::size_t nq = (std::size_t)robot_state_->get_nq();
  if (update_vars_.idx_state >= state_ref_.size()) {
    // hover at the last configuration with a yaw-only attitude (:180-186); the quaternion is rebuilt from (w, z) only,
    // but the x / y components copied with head(nq) stay in the reference
    const VectorXd& last = state_ref_.back();
    update_vars_.state_ref = robot_state_->zero();
    std::copy(last.begin(), last.begin() + (long)nq, update_vars_.state_ref.begin());
    const double w = last[6], z = last[5];
    const double norm = std::sqrt(w * w + z * z);  // Eigen::Quaterniond::normalize(): coefficient-wise division by the norm
    update_vars_.state_ref[5] = z / norm;
    update_vars_.state_ref[6] = w / norm;
  } else {
    // alpha = (time - t_ref[i-1]) / (t_ref[i] - t_ref[i-1]) is evaluated in std::size_t arithmetic => 0 (:188-189):
    // pinocchio::interpolate(q0, q1, 0) = q0 and v0 + 0 (v1 - v0) = v0, i.e. a piecewise-constant reference
    update_vars_.state_ref = state_ref_[update_vars_.idx_state - 1];
  }
}

void RailMpc::updateFreeCosts(std::size_t idx) {
  computeStateReference(update_vars_.node_time);
  int_models_[idx]->costs->get_costs().at("rail_state")->cost->reference = update_vars_.state_ref;
  syncCost(idx, "rail_state");
}

void RailMpc::updateProblem(const std::size_t& current_time) {
  beginUpdate();
  for (std::size_t i = 0; i < int_models_.size(); ++i) {
    update_vars_.node_time = current_time + i * params_.dt;
    updateFreeCosts(i);
  }
  endUpdate();
}

// ---- WeightedMpc (src/mpc-controllers/weighted-mpc.cpp) -----------------------------------------------------------------
WeightedMpc::WeightedMpc(const std::shared_ptr<Trajectory>& trajectory, std::size_t /*dt_ref*/, const std::string& yaml_path,
                         bool create_solver)
    : MpcAbstract(yaml_path), trajectory_(trajectory), cost_factory_(std::make_shared<CostModelFactory>()) {
  defer_solver_ = !create_solver;
  auto dbl = [&](const char* key, double def) {
    try { return params_server_->getParam<double>(std::string("mpc_controller/") + key); }
    catch (const std::exception&) { return def; }
  };
  alpha_ = dbl("weighted_alpha", 20.0);
  beta_ = dbl("weighted_beta", 1.0);
  state_reg_ = dbl("weighted_state_reg", 1e-1);
  control_reg_ = dbl("weighted_control_reg", 1e-1);
  // transition stages are merged into the stage that follows them (:63-75): the trajectory object is modified in place
  for (std::size_t i = 0; i < trajectory_->get_stages().size();) {
    const auto& stages = trajectory_->get_stages();
    if (stages[i]->get_is_transition()) {
      if (i + 1 >= stages.size()) throw std::runtime_error("WeightedMpc: the last stage of the trajectory is a transition stage");
      stages[i + 1]->set_duration(stages[i]->get_duration() + stages[i + 1]->get_duration());
      stages[i + 1]->set_t_ini(stages[i]->get_t_ini());
      trajectory_->removeStage(i);
      t_stages_.push_back(trajectory_->get_stages()[i]->get_t_ini());
      ++i;
    } else {
      t_stages_.push_back(stages[i]->get_t_ini());
      ++i;
    }
  }
  createProblem();
}

std::shared_ptr<CostModelSum> WeightedMpc::createCosts() const {
  auto costs = std::make_shared<CostModelSum>();
  for (const auto& stage : trajectory_->get_stages()) {
    if (stage->get_is_transition()) continue;
    const std::string path_to_stage = "stages/" + stage->get_name();
    for (const auto& ctype : stage->get_cost_types()) {
      CostModelTypes cost_type = ctype.second;
      auto cost = cost_factory_->create(path_to_stage + "/costs/" + ctype.first + "/", trajectory_->get_params_server(), robot_state_,
                                        nu_, cost_type);
      costs->addCost(stage->get_name() + "/" + ctype.first, cost, stage->get_costs()->get_costs().at(ctype.first)->weight, false);
    }
  }
  return costs;
}

void WeightedMpc::createProblem() {
  if (trajectory_->get_has_contact()) throw std::runtime_error("Weighted with contact has not been implemented");
  checkHotPathSupport();
  for (std::size_t i = 0; i < params_.knots; ++i) int_models_.push_back(makeKnotModel(createCosts()));
  finishProblem();
}

void WeightedMpc::computeActiveStage(std::size_t current_time) {
  update_vars_.idx_stage = std::size_t(std::upper_bound(t_stages_.begin(), t_stages_.end(), current_time) - t_stages_.begin()) - 1;
}

void WeightedMpc::computeActiveStage(std::size_t current_time, std::size_t last_stage) {
  computeActiveStage(current_time);
  if (update_vars_.idx_stage == last_stage + 2) update_vars_.idx_stage -= 1;
}

void WeightedMpc::computeWeight(std::size_t time) {
  // saturate the weight once the nodes are beyond the end of the trajectory (:229-241)
  if (time > trajectory_->get_duration()) update_vars_.weight_time = 0.0;
  else {
    const auto& st = trajectory_->get_stages()[update_vars_.idx_stage];
    update_vars_.weight_time = ((int)time - ((int)st->get_t_ini() + (int)st->get_duration())) / 1000.0;
  }
  update_vars_.weight = std::exp(alpha_ * update_vars_.weight_time);
}

void WeightedMpc::updateFreeCosts(std::size_t idx) {
  auto& costs = int_models_[idx]->costs->get_costs();
  const std::string& ns = update_vars_.name_stage;
  for (auto& kv : costs) {
    const std::string& name = kv.first;
    if (name.compare(0, ns.size(), ns) == 0) {  // prefix match, like the reference (:205)
      kv.second->active = true;
      if (name.compare(ns.size(), 4, "/reg") != 0 && name.compare(ns.size(), 7, "/limits") != 0) {
        computeWeight(update_vars_.node_time);
        kv.second->weight = trajectory_->get_stages()[update_vars_.idx_stage]->get_costs()->get_costs().at(name.substr(ns.size() + 1))->weight *
                            update_vars_.weight * beta_;
      }
    } else if (name != "barrier") {
      kv.second->active = false;
    }
    if (name != "barrier") syncCost(idx, name);
  }
}

WeightedSchedule WeightedMpc::schedule() const {
  WeightedSchedule s;
  const auto& stages = trajectory_->get_stages();
  s.duration = (std::int64_t)trajectory_->get_duration();
  s.alpha = alpha_; s.beta = beta_;
  std::vector<std::string> names;  // the cost names of a knot in map order, barrier excluded
  for (const auto& kv : int_models_.front()->costs->get_costs())
    if (kv.first != "barrier") names.push_back(kv.first);
  s.n_slots = names.size();
  if (t_stages_.size() != stages.size()) throw std::runtime_error("WeightedMpc: stage table out of sync");
  for (std::size_t si = 0; si < stages.size(); ++si) {
    const auto& st = stages[si];
    s.t_ini.push_back((std::int64_t)t_stages_[si]);  // what computeActiveStage searches
    s.t_end.push_back((std::int64_t)st->get_t_ini() + (std::int64_t)st->get_duration());
    const std::string& ns = st->get_name();
    for (const auto& name : names) {
      const bool m = name.compare(0, ns.size(), ns) == 0;
      const bool task = m && name.compare(ns.size(), 4, "/reg") != 0 && name.compare(ns.size(), 7, "/limits") != 0;
      double base = 0.0;
      if (task) {
        // updateFreeCosts looks the cost up in the stage by the rest of the name; with stage names that are prefixes of
        // one another that lookup throws in the reference, and so does the schedule
        base = st->get_costs()->get_costs().at(name.substr(ns.size() + 1))->weight;
      }
      s.match.push_back(m); s.task.push_back(task); s.base.push_back(base);
    }
  }
  return s;
}

void WeightedMpc::updateProblem(const std::size_t& current_time) {
  computeActiveStage(current_time);
  update_vars_.idx_last_stage = update_vars_.idx_stage;
  beginUpdate();
  for (std::size_t i = 0; i < int_models_.size(); ++i) {
    update_vars_.node_time = current_time + i * params_.dt;
    computeActiveStage(update_vars_.node_time, update_vars_.idx_last_stage);
    update_vars_.name_stage = trajectory_->get_stages()[update_vars_.idx_stage]->get_name();
    updateFreeCosts(i);
    update_vars_.idx_last_stage = update_vars_.idx_stage;
  }
  endUpdate();
}

}  // namespace eagle_mpc
