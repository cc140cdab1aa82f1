// eagle_mpc.hpp — host-side mirror of eagle-mpc's C++ surfaces for the SbFDDP hot path (drop-in boundary).
//
// Same class names, method names, argument meaning and error behaviour as the reference, re-implemented without
// yaml-cpp / Pinocchio / Crocoddyl / Boost / Eigen (none exist in this environment):
//   ParserYaml, ParamsServer, converter<T>      include/eagle_mpc/utils/{parser_yaml,params_server,converter}.hpp
//   MultiCopterBaseParams                       include/eagle_mpc/multicopter-base-params.hpp
//   Stage, Trajectory                           include/eagle_mpc/{stage,trajectory}.hpp
//   Cost / Activation factories                 include/eagle_mpc/factory/{cost,activation}.hpp
//   SolverSbFDDP                                include/eagle_mpc/sbfddp.hpp
//   MpcAbstract, CarrotMpc, RailMpc, WeightedMpc include/eagle_mpc/mpc-base.hpp, mpc-controllers/*.hpp
// The crocoddyl objects the reference passes around (ShootingProblem, ActionModel, CostModelSum, ...) are replaced by
// small host structs that carry exactly what the kernels need; the solver flattens them into an `empc_problem_desc_t`
// and calls the CUDA path through the C ABI (include/empc_b200.h).  There is no CPU solver here.
#pragma once
#include <cstddef>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/empc_b200.h"

namespace eagle_mpc {

typedef std::vector<double> VectorXd;

// ---------------------------------------------------------------------------------------------------------------------
class MissingValueException : public std::runtime_error {
 public:
  explicit MissingValueException(const std::string& msg) : std::runtime_error(msg) {}
};

template <typename T>
struct converter {
  static T convert(const std::string& val);
};

class ParamsServer {
 public:
  ParamsServer() {}
  explicit ParamsServer(const std::map<std::string, std::string>& params) : params_(params) {}
  void addParam(const std::string& key, const std::string& value) { params_.insert({key, value}); }
  bool has(const std::string& key) const { return params_.find(key) != params_.end(); }
  template <typename T>
  T getParam(const std::string& key) const {
    auto it = params_.find(key);
    if (it == params_.end())
      throw MissingValueException("The following key: '" + key + "' has not been found in the parameters server.");
    return converter<T>::convert(it->second);
  }
  const std::map<std::string, std::string>& get_params() const { return params_; }

 private:
  std::map<std::string, std::string> params_;
};

// YAML (subset) -> flat "/"-keyed string map with the reference's key layout (src/utils/parser_yaml.cpp)
class ParserYaml {
 public:
  ParserYaml(const std::string& file, const std::string& path_root = "", bool freely_parse = false);
  const std::map<std::string, std::string>& get_params() const { return params_; }

 private:
  std::map<std::string, std::string> params_;
};

// Directories that EAGLE_MPC_YAML_DIR / EAGLE_MPC_ROBOT_DATA_DIR were baked to in the reference (config/path.hpp.in);
// here they are process-wide settings (defaults: <repo>/yaml and <repo>/fixtures/urdf, overridable by environment
// variables of the same names).
void set_yaml_dir(const std::string& dir);
void set_robot_data_dir(const std::string& dir);
std::string getYamlPath(const std::string& yaml_path);
std::string getUrdfPath(const std::string& urdf_path);

// ---------------------------------------------------------------------------------------------------------------------
// What pinocchio::urdf::buildModel(path, JointModelFreeFlyer(), model) provides to eagle-mpc.
struct RobotModel {
  int nq = 0, nv = 0;
  int njoints = 0;  // free-flyer + revolute joints (universe excluded)
  std::vector<int> parent;
  std::vector<std::vector<double>> jplace_R, jplace_p, axis, com, inertia;
  std::vector<double> mass;
  std::vector<std::string> joint_names;
  std::vector<double> effortLimit;  // nv
  struct Frame { std::string name; int joint; std::vector<double> R, p; };
  std::vector<Frame> frames;
  std::size_t getFrameId(const std::string& name) const;  // frames.size() when missing (pinocchio semantics)
};
std::shared_ptr<RobotModel> buildModelFromUrdf(const std::string& path);

struct StateMultibody {
  std::shared_ptr<RobotModel> pinocchio;
  int get_nq() const { return pinocchio->nq; }
  int get_nv() const { return pinocchio->nv; }
  int get_nx() const { return pinocchio->nq + pinocchio->nv; }
  int get_ndx() const { return 2 * pinocchio->nv; }
  VectorXd zero() const;
};

// ---------------------------------------------------------------------------------------------------------------------
class MultiCopterBaseParams {
 public:
  MultiCopterBaseParams() {}
  void autoSetup(const std::string& path_to_platform, const std::shared_ptr<ParamsServer>& server);
  void autoSetup(const std::string& path_to_platform, const std::shared_ptr<ParamsServer>& server,
                 const std::shared_ptr<RobotModel>& robot_model);
  void setControlLimits(const std::shared_ptr<RobotModel>& robot_model);

  double cf_ = 0, cm_ = 0, max_thrust_ = 0, min_thrust_ = 0, max_prop_speed_ = 0, min_prop_speed_ = 0;
  std::size_t n_rotors_ = 0;
  std::vector<double> tau_f_;  // 6 x n_rotors, row-major
  std::string base_link_name_;
  VectorXd u_lb, u_ub;
  std::vector<std::vector<double>> rotors_R_, rotors_p_;
  std::vector<int> rotors_spin_dir_;
};

struct SquashingModelSmoothSat {  // carries the bounds; the smoothing lives in the solver's per-OCP device state
  VectorXd u_lb, u_ub;
  std::size_t ns = 0;
  const VectorXd& get_s_lb() const { return u_lb; }
  const VectorXd& get_s_ub() const { return u_ub; }
  std::size_t get_ns() const { return ns; }
};

// ---------------------------------------------------------------------------------------------------------------------
enum class CostModelTypes {
  CostModelState, CostModelControl, CostModelFramePlacement, CostModelFrameRotation, CostModelFrameVelocity,
  CostModelFrameTranslation, CostModelContactFrictionCone, CostModelSquashBarrier
};
enum class ActivationModelTypes {
  ActivationModelQuad, ActivationModelQuadFlatExp, ActivationModelQuadFlatLog, ActivationModelSmooth1Norm,
  ActivationModelSmooth2Norm, ActivationModelWeightedQuad, ActivationModelQuadraticBarrier,
  ActivationModelWeightedQuadraticBarrier
};

struct ActivationModel {
  ActivationModelTypes type = ActivationModelTypes::ActivationModelQuad;
  std::size_t nr = 0;
  VectorXd weights, lb, ub;  // bounds already passed through crocoddyl::ActivationBounds(lb, ub, beta=1)
};
struct CostModelResidual {
  CostModelTypes type = CostModelTypes::CostModelState;
  ActivationModel activation;
  std::size_t frame_id = 0;  // index into RobotModel::frames
  VectorXd reference;        // layout documented in include/empc_b200.h
};
struct CostItem {
  std::string name;
  std::shared_ptr<CostModelResidual> cost;
  double weight = 0;
  bool active = true;
};
class CostModelSum {  // crocoddyl::CostModelSum: std::map => iteration in name order
 public:
  void addCost(const std::string& name, const std::shared_ptr<CostModelResidual>& cost, double weight, bool active = true);
  void removeCost(const std::string& name) { costs_.erase(name); }
  std::map<std::string, std::shared_ptr<CostItem>>& get_costs() { return costs_; }
  const std::map<std::string, std::shared_ptr<CostItem>>& get_costs() const { return costs_; }

 private:
  std::map<std::string, std::shared_ptr<CostItem>> costs_;
};

// crocoddyl::ContactModel3D / ContactModel6D as src/factory/contacts.cpp:32-81 builds them, and ContactModelMultiple
enum class ContactModelTypes { ContactModel3D, ContactModel6D };
struct ContactModel {
  ContactModelTypes type = ContactModelTypes::ContactModel3D;
  std::size_t frame_id = 0;         // index into RobotModel::frames
  double gains[2] = {0, 0};         // Baumgarte gains
  double position[3] = {0, 0, 0};   // reference position
  double rotation[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // reference orientation (6D)
};
class ContactModelMultiple {
 public:
  void addContact(const std::string& name, const std::shared_ptr<ContactModel>& contact) { contacts_[name] = contact; }
  const std::map<std::string, std::shared_ptr<ContactModel>>& get_contacts() const { return contacts_; }

 private:
  std::map<std::string, std::shared_ptr<ContactModel>> contacts_;
};
class ContactModelFactory {
 public:
  std::shared_ptr<ContactModel> create(const std::string& path_to_contact, const std::shared_ptr<ParamsServer>& server,
                                       const std::shared_ptr<StateMultibody>& state, ContactModelTypes& contact_type) const;
};

class ActivationModelFactory {
 public:
  std::shared_ptr<ActivationModel> create(const std::string& path_to_cost, const std::shared_ptr<ParamsServer>& server,
                                          std::size_t nr) const;
};
class CostModelFactory {
 public:
  std::shared_ptr<CostModelResidual> create(const std::string& path_to_cost, const std::shared_ptr<ParamsServer>& server,
                                            const std::shared_ptr<StateMultibody>& state, std::size_t nu,
                                            CostModelTypes& cost_type) const;
};

// IntegratedActionModelEuler( DifferentialActionModelFreeFwdDynamics(state, actuation, costs), dt )
struct ActionModel {
  std::shared_ptr<CostModelSum> costs;
  // DifferentialActionModelContactFwdDynamics(state, actuation, contacts, costs, 0, true) when the trajectory has a
  // contact stage (src/factory/diff-action.cpp:30-32), DifferentialActionModelFreeFwdDynamics otherwise (contacts null)
  std::shared_ptr<ContactModelMultiple> contacts;
  bool rk4 = false;  // IntegratedActionModelRK4 instead of IntegratedActionModelEuler (src/factory/int-action.cpp:24-35)
  double dt = 0;     // seconds
  bool squash = true;
  VectorXd u_lb, u_ub;
};
struct ShootingProblem {
  VectorXd x0;
  std::vector<std::shared_ptr<ActionModel>> runningModels;
  std::shared_ptr<ActionModel> terminalModel;
  std::shared_ptr<StateMultibody> state;
  std::shared_ptr<MultiCopterBaseParams> platform;
  std::size_t get_T() const { return runningModels.size(); }
  const VectorXd& get_x0() const { return x0; }
  void set_x0(const VectorXd& x) { x0 = x; }
};

class Trajectory;
class Stage : public std::enable_shared_from_this<Stage> {
 public:
  static std::shared_ptr<Stage> create(const std::shared_ptr<Trajectory>& trajectory);
  void autoSetup(const std::string& path_to_stages, const std::map<std::string, std::string>& stage,
                 const std::shared_ptr<ParamsServer>& server, std::size_t t_ini);
  void set_t_ini(std::size_t t) { t_ini_ = t; }
  void set_duration(std::size_t d) { duration_ = d; }
  const std::shared_ptr<Trajectory>& get_trajectory() const { return trajectory_; }
  const std::shared_ptr<CostModelSum>& get_costs() const { return costs_; }
  const std::shared_ptr<ContactModelMultiple>& get_contacts() const { return contacts_; }
  const std::map<std::string, ContactModelTypes>& get_contact_types() const { return contact_types_; }
  const std::map<std::string, CostModelTypes>& get_cost_types() const { return cost_types_; }
  std::size_t get_duration() const { return duration_; }
  std::size_t get_t_ini() const { return t_ini_; }
  const std::string& get_name() const { return name_; }
  bool get_is_terminal() const { return is_terminal_; }
  bool get_is_transition() const { return is_transition_; }
  bool has_contacts() const { return !contacts_->get_contacts().empty(); }

 private:
  explicit Stage(const std::shared_ptr<Trajectory>& trajectory);
  std::shared_ptr<Trajectory> trajectory_;
  std::shared_ptr<CostModelSum> costs_;
  std::shared_ptr<ContactModelMultiple> contacts_;
  std::map<std::string, CostModelTypes> cost_types_;
  std::map<std::string, ContactModelTypes> contact_types_;
  std::string name_;
  std::size_t duration_ = 0, t_ini_ = 0;
  bool is_terminal_ = false, is_transition_ = false;
};

struct ProblemParams {
  std::size_t dt = 0;
  bool use_squash = false;
  std::string integrator;
};

class Trajectory : public std::enable_shared_from_this<Trajectory> {
 public:
  static std::shared_ptr<Trajectory> create();
  void autoSetup(const std::string& yaml_path);
  std::shared_ptr<ShootingProblem> createProblem() const;
  std::shared_ptr<ShootingProblem> createProblem(std::size_t dt, bool squash, const std::string& integration_method) const;
  void removeStage(std::size_t idx_stage);
  void set_initial_state(const VectorXd& initial_state);

  const std::vector<std::shared_ptr<Stage>>& get_stages() const { return stages_; }
  std::vector<std::shared_ptr<Stage>>& stages_mut() { return stages_; }
  const std::shared_ptr<RobotModel>& get_robot_model() const { return robot_model_; }
  const std::string& get_robot_model_path() const { return robot_model_path_; }
  const std::shared_ptr<MultiCopterBaseParams>& get_platform_params() const { return platform_params_; }
  const std::shared_ptr<StateMultibody>& get_robot_state() const { return robot_state_; }
  const std::shared_ptr<SquashingModelSmoothSat>& get_squash() const { return squash_; }
  std::size_t get_actuation_nu() const { return nu_; }
  const VectorXd& get_initial_state() const { return initial_state_; }
  const std::shared_ptr<ParamsServer>& get_params_server() const { return params_server_; }
  bool get_has_contact() const { return has_contact_; }
  std::size_t get_duration() const { return duration_; }

 private:
  Trajectory();
  std::vector<std::shared_ptr<Stage>> stages_;
  std::shared_ptr<RobotModel> robot_model_;
  std::string robot_model_path_;
  std::shared_ptr<MultiCopterBaseParams> platform_params_;
  std::shared_ptr<StateMultibody> robot_state_;
  std::shared_ptr<SquashingModelSmoothSat> squash_;
  std::size_t nu_ = 0;
  VectorXd initial_state_;
  std::shared_ptr<ParamsServer> params_server_;
  bool has_contact_ = false;
  std::size_t duration_ = 0;
  ProblemParams problem_params_;
};

// ---------------------------------------------------------------------------------------------------------------------
// Flattened problem (what goes over the C ABI).  Owns the arrays the desc points to.
struct FlatProblem {
  empc_problem_desc_t desc;
  std::vector<int32_t> costset_begin, node_costset, costset_contact;
  std::vector<empc_cost_t> costs;
  std::vector<empc_contact_t> contacts;
  std::vector<double> pool;
  // where each (model, cost name) landed, for MPC retargeting
  struct Slot { int cost_index; int ref_off, w_off, lb_off, ub_off; };
  std::vector<std::map<std::string, Slot>> slots;     // per cost set
  std::vector<const ActionModel*> set_models;         // cost set -> model
  void finalize();  // (re)points desc at the vectors
};
void flatten_problem(const ShootingProblem& problem, FlatProblem& out);
void fill_cost_record(const CostItem& item, empc_cost_t& rec, double* pool);  // rewrite one record in place
// SolverSbFDDP::barrierInit (src/sbfddp.cpp:169-190) as a free function, so a problem can be flattened without a GPU
void sbfddp_barrier_init(ShootingProblem& problem, std::size_t ns, double barrier_weight);

// SolverSbFDDP : the hot path.  `batch` independent copies of the problem (different x0 / warm starts) are solved at
// once on `device`; batch = 1 reproduces the reference's single-OCP interface.
// crocoddyl::CallbackAbstract / CallbackVerbose as eagle-mpc uses them (setCallbacks, src/mpc-controllers/carrot-mpc.cpp:244-247,
// bindings/python/eagle_mpc/sbfddp.hpp:63).  A callback receives the record of one iteration (empc_iter_record_t).
class CallbackAbstract {
 public:
  virtual ~CallbackAbstract() = default;
  virtual void operator()(const empc_iter_record_t& record) = 0;
};
class CallbackVerbose : public CallbackAbstract {
 public:
  void operator()(const empc_iter_record_t& record) override;
};

class SolverSbFDDP {
 public:
  SolverSbFDDP(const std::shared_ptr<ShootingProblem>& problem, const std::shared_ptr<SquashingModelSmoothSat>& squashing_model,
               int batch = 1, int device = 0);
  virtual ~SolverSbFDDP();
  // crocoddyl::SolverAbstract::solve signature (include/eagle_mpc/sbfddp.hpp:43-47); regInit is ignored like upstream.
  bool solve(const std::vector<VectorXd>& init_xs = {}, const std::vector<VectorXd>& init_us = {}, std::size_t maxiter = 100,
             bool is_feasible = false, double regInit = 1e-9);
  void setCandidate(const std::vector<VectorXd>& xs_warm = {}, const std::vector<VectorXd>& us_warm = {}, bool is_feasible = false);
  // batched variants: flat arrays batch*(T+1)*nx / batch*T*nu (nullptr => zero guess)
  bool solveBatch(const double* x0, const double* xs, const double* us, std::size_t maxiter = 100, bool is_feasible = false);
  bool solveWarm(std::size_t maxiter);  // re-solve from the candidate left on the device (MPC warm start)

  void setCallbacks(const std::vector<std::shared_ptr<CallbackAbstract>>& callbacks);
  const std::vector<std::shared_ptr<CallbackAbstract>>& getCallbacks() const { return callbacks_; }

  const std::vector<VectorXd>& get_xs() const { return xs_; }
  const std::vector<VectorXd>& get_us() const { return us_; }
  const std::vector<VectorXd>& getSquashControls() const { return us_squash_; }
  const std::vector<std::vector<double>>& get_K() const { return K_; }
  const std::vector<VectorXd>& get_k() const { return k_; }
  double get_cost() const { return cost_; }
  std::size_t get_iter() const { return iter_; }
  double get_stop() const { return stop_; }
  bool get_is_feasible() const { return is_feasible_; }
  double get_convergence_init() const { return params_.convergence_init; }
  void set_convergence_init(double c) { params_.convergence_init = c; }
  const std::shared_ptr<ShootingProblem>& get_problem() const { return problem_; }
  empc_solver_t* handle() const { return handle_; }
  FlatProblem& flat() { return flat_; }
  int batch() const { return batch_; }
  void syncX0();                          // push problem_->x0 (batch 1) to the device
  void pushCosts(int first, int n);       // after host-side edits of flat_.costs / flat_.pool
  void pushAllCosts();
  void fetch(bool with_gains = true);     // device -> host mirrors (OCP 0)
  empc_solver_params_t& params() { return params_; }

 protected:
  // solver_type = EMPC_SOLVER_BOXFDDP / EMPC_SOLVER_BOXDDP: crocoddyl's box solvers on the same device path (no squashing model,
  // no barrier cost, upstream stop rule; see SolverBoxFDDP / SolverBoxDDP below)
  SolverSbFDDP(const std::shared_ptr<ShootingProblem>& problem, const std::shared_ptr<SquashingModelSmoothSat>& squashing_model,
               int batch, int device, int solver_type);

 private:
  void init();
  void barrierInit();
  void replayCallbacks();
  std::size_t nu_ = 0;
  std::vector<std::shared_ptr<CallbackAbstract>> callbacks_;
  std::shared_ptr<ShootingProblem> problem_;
  std::shared_ptr<SquashingModelSmoothSat> squashing_model_;
  FlatProblem flat_;
  empc_solver_t* handle_ = nullptr;
  empc_solver_params_t params_;
  int batch_ = 1, device_ = 0;
  int solver_type_ = EMPC_SOLVER_SBFDDP;
  std::vector<VectorXd> xs_, us_, us_squash_, k_;
  std::vector<std::vector<double>> K_;
  double cost_ = 0, stop_ = 0;
  std::size_t iter_ = 0;
  bool is_feasible_ = false;
};

// crocoddyl::SolverBoxFDDP / crocoddyl::SolverBoxDDP, the other two SolverTypes of the reference
// (include/eagle_mpc/mpc-base.hpp:36-47; built at src/mpc-controllers/carrot-mpc.cpp:236-241, rail-mpc.cpp:118-123,
// weighted-mpc.cpp:135-140 and by examples/python/trajectory.py:24 for problems created with squash = false).  Same surface
// as SolverSbFDDP (solve, setCandidate, setCallbacks, get_xs / get_us / get_K / get_k ...), same device path: the box QP of
// computeGains runs inside backward_kernel<D, true, true>, the rollouts clamp the trial controls.
class SolverBoxFDDP : public SolverSbFDDP {
 public:
  explicit SolverBoxFDDP(const std::shared_ptr<ShootingProblem>& problem, int batch = 1, int device = 0)
      : SolverSbFDDP(problem, nullptr, batch, device, EMPC_SOLVER_BOXFDDP) {}
};
class SolverBoxDDP : public SolverSbFDDP {
 public:
  explicit SolverBoxDDP(const std::shared_ptr<ShootingProblem>& problem, int batch = 1, int device = 0)
      : SolverSbFDDP(problem, nullptr, batch, device, EMPC_SOLVER_BOXDDP) {}
};

}  // namespace eagle_mpc
