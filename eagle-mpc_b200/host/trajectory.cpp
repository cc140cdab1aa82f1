// trajectory.cpp — problem factory: platform parameters, cost/activation factories, Stage, Trajectory, and the
// flattening of a ShootingProblem into the POD description the kernels consume.
//
// Mirrors (same defaults, error messages and quirks):
//   src/multicopter-base-params.cpp:27-101   tau_f, control bounds
//   src/factory/activation.cpp:17-105        activation factory (default Quad, unit weights, bound dimension checks)
//   src/factory/cost.cpp:17-171              cost factory (reference defaults, quaternion normalisation, frame lookup)
//   src/stage.cpp:26-71                      Stage::autoSetup ("active" key presence => inactive cost)
//   src/trajectory.cpp:21-143                Trajectory::autoSetup / createProblem (knot rule, shared stage models)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>

#include "eagle_mpc.hpp"

namespace eagle_mpc {

static void quat_to_R(const double* q /*xyzw*/, double* R) {
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double x = q[0] / n, y = q[1] / n, z = q[2] / n, w = q[3] / n;  // quat.normalize()
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

// ---- MultiCopterBaseParams ------------------------------------------------------------------------------------------
void MultiCopterBaseParams::autoSetup(const std::string& path, const std::shared_ptr<ParamsServer>& server) {
  try {
    cf_ = server->getParam<double>(path + "cf");
    cm_ = server->getParam<double>(path + "cm");
    max_thrust_ = server->getParam<double>(path + "max_thrust");
    min_thrust_ = server->getParam<double>(path + "min_thrust");
    max_prop_speed_ = std::sqrt(max_thrust_ / cf_);
    min_prop_speed_ = std::sqrt(min_thrust_ / cf_);
    base_link_name_ = server->getParam<std::string>(path + "base_link_name");
    n_rotors_ = (std::size_t)server->getParam<int>(path + "n_rotors");
    std::vector<std::string> rotors = server->getParam<std::vector<std::string>>(path + "rotors");
    if (n_rotors_ != rotors.size())
      throw std::runtime_error("'n_rotors' field and the number of rotor poses specified must be the same.");
    for (std::size_t i = 0; i < n_rotors_; ++i) {
      auto rotor = converter<std::map<std::string, std::string>>::convert(rotors[i]);
      VectorXd t = converter<VectorXd>::convert(rotor.at("translation"));
      VectorXd o = converter<VectorXd>::convert(rotor.at("orientation"));
      VectorXd sd = converter<VectorXd>::convert(rotor.at("spin_direction"));
      std::vector<double> R(9);
      quat_to_R(o.data(), R.data());
      rotors_R_.push_back(R);
      rotors_p_.push_back(t);
      rotors_spin_dir_.push_back((int)sd[0]);
    }
  } catch (const std::exception& e) {
    std::cerr << e.what() << '\n';  // the reference swallows the exception and carries on (:63-65)
  }
  tau_f_.assign(6 * n_rotors_, 0.0);
  for (std::size_t i = 0; i < rotors_R_.size(); ++i) {
    const double* R = rotors_R_[i].data();
    const double* p = rotors_p_[i].data();
    const double tw[3] = {R[2], R[5], R[8]};  // R e3
    const double k = rotors_spin_dir_[i] * cm_ / cf_;
    const double cr[3] = {p[1] * tw[2] - p[2] * tw[1], p[2] * tw[0] - p[0] * tw[2], p[0] * tw[1] - p[1] * tw[0]};
    for (int r = 0; r < 3; ++r) {
      tau_f_[r * n_rotors_ + i] = tw[r];
      tau_f_[(3 + r) * n_rotors_ + i] = cr[r] + k * tw[r];
    }
  }
}
void MultiCopterBaseParams::autoSetup(const std::string& path, const std::shared_ptr<ParamsServer>& server,
                                      const std::shared_ptr<RobotModel>& robot_model) {
  autoSetup(path, server);
  setControlLimits(robot_model);
}
void MultiCopterBaseParams::setControlLimits(const std::shared_ptr<RobotModel>& m) {
  const std::size_t n_arm = (std::size_t)m->nq - 7;
  u_lb.assign(n_arm + n_rotors_, 0.0);
  u_ub = u_lb;
  for (std::size_t i = 0; i < n_rotors_; ++i) { u_lb[i] = min_thrust_; u_ub[i] = max_thrust_; }
  for (std::size_t i = 0; i < n_arm; ++i) {
    const double e = m->effortLimit[m->effortLimit.size() - n_arm + i];
    u_lb[n_rotors_ + i] = -e; u_ub[n_rotors_ + i] = e;
  }
}

// ---- factories --------------------------------------------------------------------------------------------------------
void CostModelSum::addCost(const std::string& name, const std::shared_ptr<CostModelResidual>& cost, double weight, bool active) {
  auto item = std::make_shared<CostItem>();
  item->name = name; item->cost = cost; item->weight = weight; item->active = active;
  if (!costs_.insert({name, item}).second) std::cout << "Warning: this cost item named " << name << " already existed." << std::endl;
}

static const std::map<std::string, ActivationModelTypes> ActivationModelTypes_map = {
    {"ActivationModelQuad", ActivationModelTypes::ActivationModelQuad},
    {"ActivationModelQuadFlatExp", ActivationModelTypes::ActivationModelQuadFlatExp},
    {"ActivationModelQuadFlatLog", ActivationModelTypes::ActivationModelQuadFlatLog},
    {"ActivationModelSmooth1Norm", ActivationModelTypes::ActivationModelSmooth1Norm},
    {"ActivationModelSmooth2Norm", ActivationModelTypes::ActivationModelSmooth2Norm},
    {"ActivationModelWeightedQuad", ActivationModelTypes::ActivationModelWeightedQuad},
    {"ActivationModelQuadraticBarrier", ActivationModelTypes::ActivationModelQuadraticBarrier},
    {"ActivationModelWeightedQuadraticBarrier", ActivationModelTypes::ActivationModelWeightedQuadraticBarrier}};
static const std::map<std::string, CostModelTypes> CostModelTypes_map = {
    {"CostModelState", CostModelTypes::CostModelState},
    {"CostModelControl", CostModelTypes::CostModelControl},
    {"CostModelFramePlacement", CostModelTypes::CostModelFramePlacement},
    {"CostModelFrameRotation", CostModelTypes::CostModelFrameRotation},
    {"CostModelFrameVelocity", CostModelTypes::CostModelFrameVelocity},
    {"CostModelFrameTranslation", CostModelTypes::CostModelFrameTranslation},
    {"CostModelContactFrictionCone", CostModelTypes::CostModelContactFrictionCone}};

// crocoddyl::ActivationBounds(lb, ub, beta = 1): lb = m - beta d, ub = m + beta d
static void activation_bounds(VectorXd& lb, VectorXd& ub) {
  for (std::size_t i = 0; i < lb.size(); ++i) {
    const double m = 0.5 * (lb[i] + ub[i]), d = 0.5 * (ub[i] - lb[i]);
    lb[i] = m - 1.0 * d; ub[i] = m + 1.0 * d;
  }
}

std::shared_ptr<ActivationModel> ActivationModelFactory::create(const std::string& path, const std::shared_ptr<ParamsServer>& server,
                                                                std::size_t nr) const {
  auto act = std::make_shared<ActivationModel>();
  act->nr = nr;
  std::string name;
  try { name = server->getParam<std::string>(path + "activation"); }
  catch (const std::exception&) { name = "ActivationModelQuad"; }
  auto weights_or_ones = [&]() {
    VectorXd w;
    try { w = converter<VectorXd>::convert(server->getParam<std::string>(path + "weights")); }
    catch (const std::exception&) { w.assign(nr, 1.0); }
    if (w.size() != nr)
      throw std::runtime_error("Weights vector @" + path + "weights has dimension " + std::to_string(w.size()) + ". Should be " + std::to_string(nr));
    return w;
  };
  auto bounds = [&]() {
    act->lb = converter<VectorXd>::convert(server->getParam<std::string>(path + "l_bound"));
    act->ub = converter<VectorXd>::convert(server->getParam<std::string>(path + "u_bound"));
  };
  auto check_bounds = [&]() {
    if (act->lb.size() != nr)
      throw std::runtime_error("l_bound vector @" + path + "l_bound has dimension " + std::to_string(act->lb.size()) + ". Should be " + std::to_string(nr));
    if (act->ub.size() != nr)
      throw std::runtime_error("u_bound vector @" + path + "u_bound has dimension " + std::to_string(act->ub.size()) + ". Should be " + std::to_string(nr));
  };
  const auto it = ActivationModelTypes_map.find(name);
  if (it == ActivationModelTypes_map.end()) throw std::out_of_range("map::at");  // ActivationModelTypes_map.at(name)
  act->type = it->second;
  switch (act->type) {
    case ActivationModelTypes::ActivationModelQuad: break;
    case ActivationModelTypes::ActivationModelWeightedQuad: act->weights = weights_or_ones(); break;
    case ActivationModelTypes::ActivationModelQuadraticBarrier:
      bounds(); check_bounds(); activation_bounds(act->lb, act->ub); break;
    case ActivationModelTypes::ActivationModelWeightedQuadraticBarrier:
      bounds(); act->weights = weights_or_ones(); check_bounds(); activation_bounds(act->lb, act->ub); break;
    default:
      throw std::runtime_error("Activation '" + name + "' @" + path + "activation not found");
  }
  return act;
}

std::shared_ptr<CostModelResidual> CostModelFactory::create(const std::string& path, const std::shared_ptr<ParamsServer>& server,
                                                            const std::shared_ptr<StateMultibody>& state, std::size_t nu,
                                                            CostModelTypes& cost_type) const {
  ActivationModelFactory af;
  auto cost = std::make_shared<CostModelResidual>();
  try {
    cost_type = CostModelTypes_map.at(server->getParam<std::string>(path + "type"));
  } catch (const std::exception&) {
    throw std::runtime_error("Cost " + server->getParam<std::string>(path + "type") + " not found. Please make sure the specified cost exists.");
  }
  cost->type = cost_type;
  auto frame_of = [&]() {
    const std::string link_name = server->getParam<std::string>(path + "link_name");
    const std::size_t id = state->pinocchio->getFrameId(link_name);
    if (id == state->pinocchio->frames.size()) throw std::runtime_error("Link " + link_name + "does no exists");
    return id;
  };
  auto vec = [&](const char* key) { return converter<VectorXd>::convert(server->getParam<std::string>(path + key)); };
  switch (cost_type) {
    case CostModelTypes::CostModelState: {
      cost->activation = *af.create(path, server, (std::size_t)state->get_ndx());
      try { cost->reference = vec("reference"); }
      catch (const std::exception&) { cost->reference = state->zero(); }
      if ((int)cost->reference.size() != state->get_nx())
        throw std::runtime_error("State reference vector @" + path + "reference has dimension " + std::to_string(cost->reference.size()) +
                                 ". Should be " + std::to_string(state->get_nx()));
    } break;
    case CostModelTypes::CostModelControl: {
      cost->activation = *af.create(path, server, nu);
      try { cost->reference = vec("reference"); }
      catch (const std::exception&) { cost->reference.assign(nu, 0.0); }
      if (cost->reference.size() != nu)
        throw std::runtime_error("Control reference vector @" + path + "reference has dimension " + std::to_string(cost->reference.size()) +
                                 ". Should be " + std::to_string(nu));
    } break;
    case CostModelTypes::CostModelFramePlacement: {
      cost->activation = *af.create(path, server, 6);
      VectorXd position = vec("position"), orientation = vec("orientation");
      cost->frame_id = frame_of();
      cost->reference.assign(12, 0.0);
      quat_to_R(orientation.data(), cost->reference.data());
      for (int i = 0; i < 3; ++i) cost->reference[9 + i] = position[i];
    } break;
    case CostModelTypes::CostModelFrameRotation: {
      cost->activation = *af.create(path, server, 3);
      VectorXd orientation = vec("orientation");
      cost->frame_id = frame_of();
      cost->reference.assign(9, 0.0);
      quat_to_R(orientation.data(), cost->reference.data());
    } break;
    case CostModelTypes::CostModelFrameVelocity: {
      cost->activation = *af.create(path, server, 6);
      VectorXd linear = vec("linear"), angular = vec("angular");
      cost->frame_id = frame_of();
      cost->reference = {linear[0], linear[1], linear[2], angular[0], angular[1], angular[2]};
    } break;
    case CostModelTypes::CostModelFrameTranslation: {
      cost->activation = *af.create(path, server, 3);
      cost->reference = vec("position");
      cost->frame_id = frame_of();
    } break;
    case CostModelTypes::CostModelContactFrictionCone: {
      // crocoddyl::FrictionCone(n_surf, mu, 4, false) + ActivationModelQuadraticBarrier(bounds(cone.lb, cone.ub)) +
      // ResidualModelContactFrictionCone (src/factory/cost.cpp:149-167).  reference = A, 5 x 3 row-major:
      // rows (-mu z +- t_i)^T c_R_o for the nf/2 = 2 tangents t_i = (cos, sin, 0)(2 pi i / nf), then n_surf^T;
      // bounds (-max, 0] for the cone facets, [0, max) for the normal force (min_nforce = 0, max_nforce = max)
      VectorXd n = vec("n_surf");
      const double mu = server->getParam<double>(path + "mu");
      if (n.size() != 3) throw std::runtime_error("n_surf @" + path + "n_surf must have 3 entries");
      const double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      for (int i = 0; i < 3; ++i) n[i] /= nn;
      // c_R_o = Quaternion::FromTwoVectors(n_surf, UnitZ): rotation about n x z by the angle between them
      double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      {
        const double ax[3] = {n[1], -n[0], 0.0};  // n x z
        const double sn = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1]), cs = n[2];
        if (sn > 1e-12) {
          const double k[3] = {ax[0] / sn, ax[1] / sn, 0.0};
          const double K[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
          for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
              double kk = 0;
              for (int l = 0; l < 3; ++l) kk += K[3 * i + l] * K[3 * l + j];
              R[3 * i + j] = (i == j ? 1.0 : 0.0) + sn * K[3 * i + j] + (1.0 - cs) * kk;
            }
        } else if (cs < 0) {  // n = -z: half turn about x
          R[4] = -1; R[8] = -1;
        }
      }
      cost->reference.assign(15, 0.0);
      const double z[3] = {0, 0, 1};
      for (int i = 0; i < 2; ++i) {
        const double th = 2.0 * M_PI * double(i) / 4.0;
        const double ts[3] = {std::cos(th), std::sin(th), 0.0};
        for (int c = 0; c < 3; ++c) {
          double p = 0, q = 0;
          for (int l = 0; l < 3; ++l) { p += (-mu * z[l] + ts[l]) * R[3 * l + c]; q += (-mu * z[l] - ts[l]) * R[3 * l + c]; }
          cost->reference[3 * (2 * i) + c] = p;
          cost->reference[3 * (2 * i + 1) + c] = q;
        }
      }
      for (int c = 0; c < 3; ++c) cost->reference[12 + c] = n[c];
      cost->frame_id = frame_of();
      cost->activation.type = ActivationModelTypes::ActivationModelQuadraticBarrier;
      cost->activation.nr = 5;
      const double big = std::numeric_limits<double>::max();
      cost->activation.lb = {-big, -big, -big, -big, 0.0};
      cost->activation.ub = {0.0, 0.0, 0.0, 0.0, big};
      activation_bounds(cost->activation.lb, cost->activation.ub);
    } break;
    default:
      throw std::runtime_error("cost type not supported");
  }
  return cost;
}

// src/factory/contacts.cpp:17-81
std::shared_ptr<ContactModel> ContactModelFactory::create(const std::string& path, const std::shared_ptr<ParamsServer>& server,
                                                          const std::shared_ptr<StateMultibody>& state,
                                                          ContactModelTypes& contact_type) const {
  static const std::map<std::string, ContactModelTypes> ContactModelTypes_map = {
      {"ContactModel3D", ContactModelTypes::ContactModel3D}, {"ContactModel6D", ContactModelTypes::ContactModel6D}};
  try {
    contact_type = ContactModelTypes_map.at(server->getParam<std::string>(path + "type"));
  } catch (const std::exception&) {
    throw std::runtime_error("Contact " + server->getParam<std::string>(path + "type") + "not found. Please make sure the specified contact exists.");
  }
  auto contact = std::make_shared<ContactModel>();
  contact->type = contact_type;
  VectorXd position = converter<VectorXd>::convert(server->getParam<std::string>(path + "position"));
  for (int i = 0; i < 3; ++i) contact->position[i] = position[i];
  if (contact_type == ContactModelTypes::ContactModel6D) {
    VectorXd orientation = converter<VectorXd>::convert(server->getParam<std::string>(path + "orientation"));
    quat_to_R(orientation.data(), contact->rotation);  // Eigen::Quaterniond(orientation).normalize().toRotationMatrix()
  }
  const std::string link_name = server->getParam<std::string>(path + "link_name");
  contact->frame_id = state->pinocchio->getFrameId(link_name);
  if (contact->frame_id == state->pinocchio->frames.size()) throw std::runtime_error("Link " + link_name + "does no exists");
  try {
    VectorXd gains = converter<VectorXd>::convert(server->getParam<std::string>(path + "gains"));
    contact->gains[0] = gains[0]; contact->gains[1] = gains[1];
  } catch (const std::exception&) {
    contact->gains[0] = contact->gains[1] = 0;  // "Set to the zero gains vector"
  }
  return contact;
}

// ---- Stage ------------------------------------------------------------------------------------------------------------
Stage::Stage(const std::shared_ptr<Trajectory>& trajectory)
    : trajectory_(trajectory), costs_(std::make_shared<CostModelSum>()), contacts_(std::make_shared<ContactModelMultiple>()) {}
std::shared_ptr<Stage> Stage::create(const std::shared_ptr<Trajectory>& trajectory) { return std::shared_ptr<Stage>(new Stage(trajectory)); }

void Stage::autoSetup(const std::string& path_to_stages, const std::map<std::string, std::string>& stage,
                      const std::shared_ptr<ParamsServer>& server, std::size_t t_ini) {
  const std::string path_to_stage = path_to_stages + stage.at("name") + "/";
  name_ = stage.at("name");
  duration_ = std::size_t(converter<int>::convert(stage.at("duration")));
  t_ini_ = t_ini;
  is_transition_ = converter<bool>::convert(stage.at("transition"));
  if (stage.count("contacts")) {  // src/stage.cpp:38-47
    ContactModelFactory ctf;
    for (const std::string& contact_name : converter<std::vector<std::string>>::convert(stage.at("contacts"))) {
      ContactModelTypes contact_type;
      auto contact = ctf.create(path_to_stage + "contacts/" + contact_name + "/", server, trajectory_->get_robot_state(), contact_type);
      contacts_->addContact(contact_name, contact);
      contact_types_.insert({contact_name, contact_type});
    }
  }
  CostModelFactory cf;
  for (const std::string& cost_name : converter<std::vector<std::string>>::convert(stage.at("costs"))) {
    const double weight = server->getParam<double>(path_to_stage + "costs/" + cost_name + "/weight");
    bool active = false;
    try { server->getParam<double>(path_to_stage + "costs/" + cost_name + "/active"); }  // presence => inactive (:55-61)
    catch (const std::exception&) { active = true; }
    CostModelTypes cost_type;
    auto cost = cf.create(path_to_stage + "costs/" + cost_name + "/", server, trajectory_->get_robot_state(),
                          trajectory_->get_actuation_nu(), cost_type);
    costs_->addCost(cost_name, cost, weight, active);
    cost_types_.insert({cost_name, cost_type});
  }
}

// ---- Trajectory -------------------------------------------------------------------------------------------------------
Trajectory::Trajectory() {}
std::shared_ptr<Trajectory> Trajectory::create() { return std::shared_ptr<Trajectory>(new Trajectory()); }

void Trajectory::autoSetup(const std::string& yaml_path) {
  ParserYaml parser(yaml_path);
  params_server_ = std::make_shared<ParamsServer>(parser.get_params());
  robot_model_path_ = getUrdfPath(params_server_->getParam<std::string>("robot/urdf"));
  robot_model_ = buildModelFromUrdf(robot_model_path_);
  platform_params_ = std::make_shared<MultiCopterBaseParams>();
  platform_params_->autoSetup("robot/platform/", params_server_, robot_model_);
  try {
    problem_params_.use_squash = params_server_->getParam<bool>("problem_params/use_squash");
    problem_params_.dt = (std::size_t)params_server_->getParam<int>("problem_params/dt");
    problem_params_.integrator = params_server_->getParam<std::string>("problem_params/integrator");
  } catch (const std::exception&) {
    problem_params_.use_squash = false; problem_params_.dt = 0; problem_params_.integrator = "";
  }
  robot_state_ = std::make_shared<StateMultibody>();
  robot_state_->pinocchio = robot_model_;
  nu_ = platform_params_->n_rotors_ + (std::size_t)(robot_model_->nv - 6);  // ActuationModelMultiCopterBase::nu
  squash_ = std::make_shared<SquashingModelSmoothSat>();
  squash_->u_lb = platform_params_->u_lb; squash_->u_ub = platform_params_->u_ub; squash_->ns = nu_;
  try { initial_state_ = params_server_->getParam<VectorXd>("initial_state"); }
  catch (const std::exception&) { initial_state_ = robot_state_->zero(); }
  if ((int)initial_state_.size() != robot_state_->get_nx())
    throw std::runtime_error("The specified initial state has wrong dimension. Should be " + std::to_string(robot_state_->get_nx()) +
                             " and it has " + std::to_string(initial_state_.size()));
  auto stages_params = params_server_->getParam<std::vector<std::map<std::string, std::string>>>("stages");
  std::size_t time = 0;
  bool stage_duration_0 = false;
  for (const auto& stage_param : stages_params) {
    std::shared_ptr<Stage> stage = Stage::create(shared_from_this());
    stage->autoSetup("stages/", stage_param, params_server_, time);
    if (!stage_duration_0 && stage->get_duration() == 0) stage_duration_0 = true;
    else if (stage_duration_0 && stage->get_duration() == 0)
      throw std::runtime_error("Two consecutives stages cannot have duration 0. Please, unify them in a single stage.");
    else stage_duration_0 = false;
    time += stage->get_duration();
    stages_.push_back(stage);
    if (!has_contact_) has_contact_ = stage->has_contacts();
  }
  duration_ = time;
}

std::shared_ptr<ShootingProblem> Trajectory::createProblem() const {
  if (problem_params_.integrator == "")
    throw std::runtime_error("Problem parameters not specified in the YAML file. Try calling createProblem() by passing the problem parameters.");
  return createProblem(problem_params_.dt, problem_params_.use_squash, problem_params_.integrator);
}

std::shared_ptr<ShootingProblem> Trajectory::createProblem(std::size_t dt, bool squash, const std::string& integration_method) const {
  const bool rk4 = integration_method == "IntegratedActionModelRK4";
  if (!rk4 && integration_method != "IntegratedActionModelEuler") throw std::out_of_range("map::at");  // IntegratedActionModelTypes_map.at()
  if (rk4 && has_contact_)
    throw std::runtime_error("IntegratedActionModelRK4 over contact dynamics is not supported by the B200 hot path");
  auto problem = std::make_shared<ShootingProblem>();
  bool last_duration0 = false;
  for (auto stage = stages_.begin(); stage != stages_.end(); ++stage) {
    auto iam = std::make_shared<ActionModel>();
    iam->costs = (*stage)->get_costs();
    if (has_contact_) iam->contacts = (*stage)->get_contacts();  // dam_factory_->create(has_contact_, squash, *stage), :115-116
    iam->dt = double(dt) / 1000.;
    iam->rk4 = rk4;
    iam->squash = squash;
    std::size_t n_knots;
    if ((*stage)->get_duration() / dt == 0 && std::next(stage) != stages_.end()) { n_knots = 1; last_duration0 = true; }
    else {
      n_knots = (*stage)->get_duration() / dt;
      if (last_duration0) n_knots -= 1;
      last_duration0 = false;
    }
    iam->u_lb = platform_params_->u_lb; iam->u_ub = platform_params_->u_ub;
    problem->terminalModel = iam;
    for (std::size_t k = 0; k < n_knots; ++k) problem->runningModels.push_back(iam);  // one model shared by the stage
  }
  problem->x0 = initial_state_;
  problem->state = robot_state_;
  problem->platform = platform_params_;
  return problem;
}

void Trajectory::removeStage(std::size_t idx) { stages_.erase(stages_.begin() + (long)idx); }
void Trajectory::set_initial_state(const VectorXd& x) { initial_state_ = x; }

// ---- flattening -------------------------------------------------------------------------------------------------------
static int cost_type_code(CostModelTypes t) {
  switch (t) {
    case CostModelTypes::CostModelState: return EMPC_COST_STATE;
    case CostModelTypes::CostModelControl: return EMPC_COST_CONTROL;
    case CostModelTypes::CostModelFramePlacement: return EMPC_COST_FRAME_PLACEMENT;
    case CostModelTypes::CostModelFrameRotation: return EMPC_COST_FRAME_ROTATION;
    case CostModelTypes::CostModelFrameVelocity: return EMPC_COST_FRAME_VELOCITY;
    case CostModelTypes::CostModelFrameTranslation: return EMPC_COST_FRAME_TRANSLATION;
    case CostModelTypes::CostModelSquashBarrier: return EMPC_COST_SQUASH_BARRIER;
    case CostModelTypes::CostModelContactFrictionCone: return EMPC_COST_CONTACT_FRICTION_CONE;
    default: throw std::runtime_error("cost type not supported by the B200 hot path");
  }
}
static int act_code(ActivationModelTypes t) {
  switch (t) {
    case ActivationModelTypes::ActivationModelQuad: return EMPC_ACT_QUAD;
    case ActivationModelTypes::ActivationModelWeightedQuad: return EMPC_ACT_WEIGHTED_QUAD;
    case ActivationModelTypes::ActivationModelQuadraticBarrier: return EMPC_ACT_QUAD_BARRIER;
    case ActivationModelTypes::ActivationModelWeightedQuadraticBarrier: return EMPC_ACT_WEIGHTED_QUAD_BARRIER;
    default: throw std::runtime_error("activation type not supported by the B200 hot path");
  }
}

void FlatProblem::finalize() {
  desc.n_costsets = (int)costset_begin.size() - 1;
  desc.n_costs = (int)costs.size();
  desc.n_pool = (int)pool.size();
  desc.costset_begin = costset_begin.data();
  desc.costs = costs.data();
  desc.pool = pool.data();
  desc.node_costset = node_costset.data();
  desc.n_contacts = (int)contacts.size();
  desc.contacts = contacts.empty() ? nullptr : contacts.data();
  desc.costset_contact = contacts.empty() ? nullptr : costset_contact.data();
}

void fill_cost_record(const CostItem& item, empc_cost_t& rec, double* pool) {
  rec.weight = item.weight;
  rec.active = item.active ? 1 : 0;
  const CostModelResidual& c = *item.cost;
  if (rec.ref_off >= 0) std::copy(c.reference.begin(), c.reference.end(), pool + rec.ref_off);
  if (rec.w_off >= 0) std::copy(c.activation.weights.begin(), c.activation.weights.end(), pool + rec.w_off);
  if (rec.lb_off >= 0) std::copy(c.activation.lb.begin(), c.activation.lb.end(), pool + rec.lb_off);
  if (rec.ub_off >= 0) std::copy(c.activation.ub.begin(), c.activation.ub.end(), pool + rec.ub_off);
}

void flatten_problem(const ShootingProblem& problem, FlatProblem& out) {
  const RobotModel& rm = *problem.state->pinocchio;
  const MultiCopterBaseParams& pf = *problem.platform;
  std::memset(&out.desc, 0, sizeof(out.desc));
  empc_robot_t& r = out.desc.robot;
  if (rm.njoints > EMPC_MAX_JOINTS) throw std::runtime_error("robot has too many joints for the B200 path");
  if (pf.n_rotors_ > EMPC_MAX_ROTORS) throw std::runtime_error("too many rotors for the B200 path");
  r.n_joints = rm.njoints;
  for (int i = 0; i < rm.njoints; ++i) {
    r.parent[i] = rm.parent[i];
    std::copy(rm.jplace_R[i].begin(), rm.jplace_R[i].end(), r.jplace_R[i]);
    std::copy(rm.jplace_p[i].begin(), rm.jplace_p[i].end(), r.jplace_p[i]);
    std::copy(rm.axis[i].begin(), rm.axis[i].end(), r.axis[i]);
    r.mass[i] = rm.mass[i];
    std::copy(rm.com[i].begin(), rm.com[i].end(), r.com[i]);
    std::copy(rm.inertia[i].begin(), rm.inertia[i].end(), r.inertia[i]);
  }
  r.gravity[0] = 0; r.gravity[1] = 0; r.gravity[2] = -9.81;
  // only the frames the costs reference travel to the device
  std::map<std::size_t, int> frame_slot;
  auto slot_of = [&](std::size_t fid) {
    auto it = frame_slot.find(fid);
    if (it != frame_slot.end()) return it->second;
    const int s = (int)frame_slot.size();
    if (s >= EMPC_MAX_FRAMES) throw std::runtime_error("too many distinct frames referenced by costs");
    const RobotModel::Frame& f = rm.frames[fid];
    r.frame_joint[s] = f.joint;
    std::copy(f.R.begin(), f.R.end(), r.frame_R[s]);
    std::copy(f.p.begin(), f.p.end(), r.frame_p[s]);
    frame_slot[fid] = s;
    return s;
  };
  out.desc.n_rotors = (int)pf.n_rotors_;
  std::copy(pf.tau_f_.begin(), pf.tau_f_.end(), out.desc.tau_f);
  for (std::size_t i = 0; i < pf.u_lb.size(); ++i) { out.desc.u_lb[i] = pf.u_lb[i]; out.desc.u_ub[i] = pf.u_ub[i]; }
  const std::size_t T = problem.get_T();
  out.desc.T = (int)T;
  out.desc.dt = problem.terminalModel->dt;
  out.desc.use_squash = problem.terminalModel->squash ? 1 : 0;
  out.desc.integrator = problem.terminalModel->rk4 ? EMPC_INTEGRATOR_RK4 : EMPC_INTEGRATOR_EULER;
  out.desc.n_node_maps = 1;
  out.costset_begin.assign(1, 0);
  out.costs.clear(); out.pool.clear(); out.node_costset.assign(T + 1, 0); out.slots.clear(); out.set_models.clear();
  out.contacts.clear(); out.costset_contact.clear();
  std::map<const ActionModel*, int> set_of;
  auto add_model = [&](const ActionModel* m) {
    auto it = set_of.find(m);
    if (it != set_of.end()) return it->second;
    const int s = (int)out.set_models.size();
    set_of[m] = s;
    out.set_models.push_back(m);
    out.slots.emplace_back();
    if (m->dt != out.desc.dt) throw std::runtime_error("all action models must share the same time step");
    if (m->rk4 != problem.terminalModel->rk4) throw std::runtime_error("all action models must share the same integrator");
    // the model's ContactModelMultiple: one contact per model is what the corpus uses and what the kernels cover
    int contact_index = -1;
    if (m->contacts && !m->contacts->get_contacts().empty()) {
      if (m->contacts->get_contacts().size() > 1) throw std::runtime_error("more than one contact per stage is not supported by the B200 hot path");
      const ContactModel& c = *m->contacts->get_contacts().begin()->second;
      if (c.gains[0] != 0.0 || c.gains[1] != 0.0) throw std::runtime_error("non-zero contact gains are not supported by the B200 hot path");
      empc_contact_t rec;
      std::memset(&rec, 0, sizeof(rec));
      rec.type = c.type == ContactModelTypes::ContactModel6D ? EMPC_CONTACT_6D : EMPC_CONTACT_3D;
      rec.frame = slot_of(c.frame_id);
      rec.gains[0] = c.gains[0]; rec.gains[1] = c.gains[1];
      std::copy(c.position, c.position + 3, rec.ref_p);
      std::copy(c.rotation, c.rotation + 9, rec.ref_R);
      contact_index = (int)out.contacts.size();
      out.contacts.push_back(rec);
    }
    out.costset_contact.push_back(contact_index);
    for (const auto& kv : m->costs->get_costs()) {  // std::map => sorted by name, crocoddyl's iteration order
      const CostItem& item = *kv.second;
      const CostModelResidual& c = *item.cost;
      empc_cost_t rec;
      rec.type = cost_type_code(c.type);
      rec.activation = act_code(c.activation.type);
      const bool is_frame = (rec.type >= EMPC_COST_FRAME_PLACEMENT && rec.type <= EMPC_COST_FRAME_TRANSLATION) ||
                            rec.type == EMPC_COST_CONTACT_FRICTION_CONE;
      if (rec.type == EMPC_COST_CONTACT_FRICTION_CONE &&
          (contact_index < 0 || out.contacts[contact_index].frame != slot_of(c.frame_id)))
        throw std::runtime_error("CostModelContactFrictionCone '" + kv.first + "' needs a contact on the same link in its stage");
      rec.frame = is_frame ? slot_of(c.frame_id) : 0;
      auto reserve = [&](std::size_t n) { if (!n) return -1; const int off = (int)out.pool.size(); out.pool.resize(out.pool.size() + n, 0.0); return off; };
      rec.ref_off = reserve(c.reference.size());
      rec.w_off = reserve(c.activation.weights.size());
      rec.lb_off = reserve(c.activation.lb.size());
      rec.ub_off = reserve(c.activation.ub.size());
      fill_cost_record(item, rec, out.pool.data());
      out.slots[s][kv.first] = {(int)out.costs.size(), rec.ref_off, rec.w_off, rec.lb_off, rec.ub_off};
      out.costs.push_back(rec);
    }
    out.costset_begin.push_back((int)out.costs.size());
    return s;
  };
  for (std::size_t t = 0; t < T; ++t) out.node_costset[t] = add_model(problem.runningModels[t].get());
  out.node_costset[T] = add_model(problem.terminalModel.get());
  r.n_frames = (int)frame_slot.size();
  out.finalize();
}

}  // namespace eagle_mpc
