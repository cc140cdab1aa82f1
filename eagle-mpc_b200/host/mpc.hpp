// mpc.hpp — MPC controllers on top of the B200 SbFDDP path (host-side mirror of include/eagle_mpc/mpc-base.hpp and
// include/eagle_mpc/mpc-controllers/{carrot,rail,weighted}-mpc.hpp).  updateProblem() edits the per-knot cost tables
// exactly like the reference (src/mpc-controllers/carrot-mpc.cpp:298-401, rail-mpc.cpp:154-200,
// weighted-mpc.cpp:173-245) and pushes only the changed records to the device through empc_update_costs.
#pragma once
#include <cstdint>

#include "eagle_mpc.hpp"

namespace eagle_mpc {

enum class SolverTypes { SolverSbFDDP, SolverBoxFDDP, SolverBoxDDP };

struct MpcParams {
  std::string integrator_type;
  std::size_t knots = 0, iters = 0, dt = 0;
  SolverTypes solver_type = SolverTypes::SolverSbFDDP;
  bool callback = false;
};

class MpcAbstract {
 public:
  explicit MpcAbstract(const std::string& yaml_path);
  virtual ~MpcAbstract() {}
  virtual void createProblem() = 0;
  virtual void updateProblem(const std::size_t& current_time) = 0;

  const std::shared_ptr<RobotModel>& get_robot_model() const { return robot_model_; }
  const std::string& get_robot_model_path() const { return robot_model_path_; }
  const std::shared_ptr<MultiCopterBaseParams>& get_platform_params() const { return platform_params_; }
  const std::shared_ptr<StateMultibody>& get_robot_state() const { return robot_state_; }
  const std::shared_ptr<SquashingModelSmoothSat>& get_squash() const { return squash_; }
  const std::vector<std::shared_ptr<ActionModel>>& get_int_models() const { return int_models_; }
  const std::shared_ptr<ShootingProblem>& get_problem() const { return problem_; }
  const std::shared_ptr<SolverSbFDDP>& get_solver() const { return solver_; }
  const std::size_t& get_dt() const { return params_.dt; }
  const std::size_t& get_knots() const { return params_.knots; }
  const std::size_t& get_iters() const { return params_.iters; }
  void attachSolver();  // builds the SolverSbFDDP (needs a GPU)
  // the flattened problem (kept in sync by updateProblem): for the solver, and for the tests' oracle
  FlatProblem& flat();

 protected:
  void initializeRobotObjects();
  void loadParams();
  void checkHotPathSupport() const;  // only SolverSbFDDP + Euler are part of the B200 hot path
  std::shared_ptr<ActionModel> makeKnotModel(const std::shared_ptr<CostModelSum>& costs) const;
  void finishProblem();  // ShootingProblem(zero, int_models[:-1], int_models[-1]) + solver (or a flat copy without one)
  void syncCost(std::size_t knot, const std::string& name);  // one cost item -> flat tables, extends the dirty range
  void beginUpdate() { dirty_lo_ = 1 << 30; dirty_hi_ = -1; }
  void endUpdate();  // pushes the dirty records to the device
  std::shared_ptr<ParamsServer> params_server_;
  std::shared_ptr<RobotModel> robot_model_;
  std::string robot_model_path_;
  std::shared_ptr<MultiCopterBaseParams> platform_params_;
  std::shared_ptr<StateMultibody> robot_state_;
  std::shared_ptr<SquashingModelSmoothSat> squash_;
  std::size_t nu_ = 0;
  MpcParams params_;
  std::vector<std::shared_ptr<ActionModel>> int_models_;
  std::shared_ptr<ShootingProblem> problem_;
  std::shared_ptr<SolverSbFDDP> solver_;
  bool defer_solver_ = false;  // tests flatten the problem on machines without a GPU
  FlatProblem flat_local_;     // used when no solver is attached
  int dirty_lo_ = 1 << 30, dirty_hi_ = -1;
};

class CarrotMpc : public MpcAbstract {
 public:
  CarrotMpc(const std::shared_ptr<Trajectory>& trajectory, const std::vector<VectorXd>& state_ref, std::size_t dt_ref,
            const std::string& yaml_path, bool create_solver = true);
  void createProblem() override;
  void updateProblem(const std::size_t& current_time) override;
  const std::shared_ptr<Trajectory>& get_trajectory() const { return trajectory_; }
  const std::vector<VectorXd>& get_state_ref() const { return state_ref_; }
  const std::vector<std::size_t>& get_t_stages() const { return t_stages_; }
  const std::vector<std::size_t>& get_t_ref() const { return t_ref_; }

 private:
  void loadCostParams();
  std::shared_ptr<CostModelSum> createCosts() const;
  void computeActiveStage(std::size_t t);
  void updateFreeCosts(std::size_t idx);
  void computeStateReference(std::size_t time);

  std::shared_ptr<Trajectory> trajectory_;
  std::vector<VectorXd> state_ref_;
  std::vector<std::size_t> t_ref_, t_stages_;
  double carrot_weight_ = 10, carrot_tail_weight_ = 5, control_reg_weight_ = 1e-2, state_reg_weight_ = 1e-3, state_limits_weight_ = 100;
  VectorXd carrot_tail_act_weights_, control_reg_act_weights_, state_ref_act_weights_, state_limits_act_weights_,
      state_limits_l_bound_, state_limits_u_bound_;
  struct { std::size_t node_time = 0, idx_stage = 0, idx_last_stage = 0, idx_state = 0; VectorXd state_ref; } update_vars_;
};

// RailMpc (include/eagle_mpc/mpc-controllers/rail-mpc.hpp:26-61): every knot tracks the reference trajectory itself.
class RailMpc : public MpcAbstract {
 public:
  RailMpc(const std::vector<VectorXd>& state_ref, std::size_t dt_ref, const std::string& yaml_path, bool create_solver = true);
  void createProblem() override;
  void updateProblem(const std::size_t& current_time) override;
  const std::vector<VectorXd>& get_state_ref() const { return state_ref_; }
  const std::vector<std::size_t>& get_t_ref() const { return t_ref_; }

 private:
  std::shared_ptr<CostModelSum> createCosts() const;
  void updateFreeCosts(std::size_t idx);
  void computeStateReference(std::size_t time);

  std::vector<VectorXd> state_ref_;
  std::vector<std::size_t> t_ref_;
  VectorXd state_activation_weights_;
  double state_weight_ = 10, control_weight_ = 1e-1;
  struct { std::size_t node_time = 0, idx_state = 0; VectorXd state_ref; } update_vars_;
};

// WeightedMpc (include/eagle_mpc/mpc-controllers/weighted-mpc.hpp:27-71): every knot carries the costs of every
// (non-transition) stage of the trajectory; the ones of the stage active at the knot's time are switched on, the task
// costs with the weight exp(alpha (t - t_end_of_stage)) * beta.
// The weight schedule of a WeightedMpc in flat form, for the device-side retargeting (empc_weighted_retarget): per
// (stage, cost slot of a knot) what updateFreeCosts would do.  Slots are the costs of a knot in name order, barrier excluded.
struct WeightedSchedule {
  std::vector<std::int64_t> t_ini, t_end;  // per (merged, non-transition) stage, ms
  std::int64_t duration = 0;               // trajectory duration, ms
  double alpha = 0, beta = 0;
  std::size_t n_slots = 0;
  std::vector<std::uint8_t> match, task;   // n_stages x n_slots: prefix match / weight follows the schedule
  std::vector<double> base;                // n_stages x n_slots: the stage's own weight of that cost
};

class WeightedMpc : public MpcAbstract {
 public:
  WeightedMpc(const std::shared_ptr<Trajectory>& trajectory, std::size_t dt_ref, const std::string& yaml_path, bool create_solver = true);
  void createProblem() override;
  void updateProblem(const std::size_t& current_time) override;
  const std::shared_ptr<Trajectory>& get_trajectory() const { return trajectory_; }
  const std::vector<std::size_t>& get_t_stages() const { return t_stages_; }
  WeightedSchedule schedule() const;

 private:
  std::shared_ptr<CostModelSum> createCosts() const;
  void computeActiveStage(std::size_t current_time);
  void computeActiveStage(std::size_t current_time, std::size_t last_stage);
  void updateFreeCosts(std::size_t idx);
  void computeWeight(std::size_t time);

  std::shared_ptr<Trajectory> trajectory_;
  std::shared_ptr<CostModelFactory> cost_factory_;
  std::vector<std::size_t> t_stages_;
  double alpha_ = 20.0, beta_ = 1.0, state_reg_ = 1e-1, control_reg_ = 1e-1;
  struct { std::size_t node_time = 0, idx_stage = 0, idx_last_stage = 0; std::string name_stage; double weight = 0, weight_time = 0; } update_vars_;
};

}  // namespace eagle_mpc
