// pybind.cpp — Python front-end with the reference's names (bindings/python/eagle_mpc/eagle-mpc.cpp:41-53,
// trajectory.hpp, sbfddp.hpp:30-79, mpc-controllers/*.hpp): eagle_mpc.Trajectory / SolverSbFDDP / CallbackVerbose /
// CarrotMpc / RailMpc / WeightedMpc over the host-side C++ mirror (eagle_mpc.hpp, mpc.hpp).  The reference binds with
// Boost.Python and hands crocoddyl objects around; here the same surface is bound with pybind11, states and controls
// travel as numpy arrays, and solve() runs the CUDA path behind the C ABI (no CPU fallback: constructing a solver
// without a usable GPU raises RuntimeError).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "eagle_mpc.hpp"
#include "mpc.hpp"

namespace py = pybind11;
using namespace eagle_mpc;

static py::array_t<double> rows(const std::vector<VectorXd>& v) {
  const std::size_t n = v.size(), m = n ? v[0].size() : 0;
  py::array_t<double> a({n, m});
  auto r = a.mutable_unchecked<2>();
  for (std::size_t i = 0; i < n; ++i)
    for (std::size_t j = 0; j < m; ++j) r(i, j) = v[i][j];
  return a;
}
static py::array_t<double> vec(const VectorXd& v) { return py::array_t<double>(v.size(), v.data()); }
static std::vector<VectorXd> to_rows(const py::object& o) {
  std::vector<VectorXd> out;
  if (o.is_none()) return out;
  for (auto row : o) out.push_back(py::cast<VectorXd>(row));
  return out;
}

// Python-side callbacks: any object with __call__(record dict); CallbackVerbose prints like crocoddyl's
class PyCallback : public CallbackAbstract {
 public:
  explicit PyCallback(py::object f) : f_(std::move(f)) {}
  void operator()(const empc_iter_record_t& r) override {
    py::dict d;
    d["iter"] = r.total_iter; d["cost"] = r.cost; d["stop"] = r.stop; d["steplength"] = r.steplength; d["xreg"] = r.xreg;
    d["ureg"] = r.xreg; d["is_feasible"] = r.is_feasible != 0; d["phase"] = r.phase; d["accepted"] = r.accepted;
    d["d0"] = r.d0; d["d1"] = r.d1; d["smooth"] = r.smooth;
    f_(d);
  }

 private:
  py::object f_;
};

PYBIND11_MODULE(_eagle_mpc, m) {
  m.doc() = "eagle-mpc front-end (reference names) over the B200 SbFDDP path";
  m.def("set_yaml_dir", &set_yaml_dir);
  m.def("set_robot_data_dir", &set_robot_data_dir);

  py::class_<RobotModel, std::shared_ptr<RobotModel>>(m, "RobotModel")
      .def_readonly("nq", &RobotModel::nq).def_readonly("nv", &RobotModel::nv).def_readonly("njoints", &RobotModel::njoints)
      .def_property_readonly("effortLimit", [](const RobotModel& r) { return vec(r.effortLimit); })
      .def("getFrameId", &RobotModel::getFrameId)
      .def_property_readonly("frame_names", [](const RobotModel& r) { std::vector<std::string> n; for (auto& f : r.frames) n.push_back(f.name); return n; });
  py::class_<MultiCopterBaseParams, std::shared_ptr<MultiCopterBaseParams>>(m, "MultiCopterBaseParams")
      .def_readonly("cf", &MultiCopterBaseParams::cf_).def_readonly("cm", &MultiCopterBaseParams::cm_)
      .def_readonly("max_thrust", &MultiCopterBaseParams::max_thrust_).def_readonly("min_thrust", &MultiCopterBaseParams::min_thrust_)
      .def_readonly("n_rotors", &MultiCopterBaseParams::n_rotors_).def_readonly("base_link_name", &MultiCopterBaseParams::base_link_name_)
      .def_property_readonly("tau_f", [](const MultiCopterBaseParams& p) {
        py::array_t<double> a({(std::size_t)6, p.n_rotors_});
        auto r = a.mutable_unchecked<2>();
        for (std::size_t i = 0; i < 6; ++i) for (std::size_t j = 0; j < p.n_rotors_; ++j) r(i, j) = p.tau_f_[i * p.n_rotors_ + j];
        return a; })
      .def_property_readonly("u_lb", [](const MultiCopterBaseParams& p) { return vec(p.u_lb); })
      .def_property_readonly("u_ub", [](const MultiCopterBaseParams& p) { return vec(p.u_ub); });
  py::class_<SquashingModelSmoothSat, std::shared_ptr<SquashingModelSmoothSat>>(m, "SquashingModelSmoothSat")
      .def_property_readonly("s_lb", [](const SquashingModelSmoothSat& s) { return vec(s.u_lb); })
      .def_property_readonly("s_ub", [](const SquashingModelSmoothSat& s) { return vec(s.u_ub); })
      .def_property_readonly("ns", &SquashingModelSmoothSat::get_ns);
  py::class_<ShootingProblem, std::shared_ptr<ShootingProblem>>(m, "ShootingProblem")
      .def_property_readonly("T", &ShootingProblem::get_T)
      .def_property("x0", [](const ShootingProblem& p) { return vec(p.x0); }, [](ShootingProblem& p, const VectorXd& x) {
        if (x.size() != p.x0.size()) throw std::runtime_error("x0 has the wrong dimension");
        p.x0 = x; })
      .def_property_readonly("nx", [](const ShootingProblem& p) { return p.state->get_nx(); })
      .def_property_readonly("ndx", [](const ShootingProblem& p) { return p.state->get_ndx(); });
  py::class_<Stage, std::shared_ptr<Stage>>(m, "Stage")
      .def_property_readonly("name", &Stage::get_name).def_property_readonly("duration", &Stage::get_duration)
      .def_property_readonly("t_ini", &Stage::get_t_ini).def_property_readonly("is_transition", &Stage::get_is_transition)
      .def_property_readonly("cost_names", [](const Stage& s) { std::vector<std::string> n; for (auto& c : s.get_costs()->get_costs()) n.push_back(c.first); return n; });

  py::class_<Trajectory, std::shared_ptr<Trajectory>>(m, "Trajectory")
      .def(py::init([]() { return Trajectory::create(); }))
      .def("autoSetup", &Trajectory::autoSetup, py::arg("yaml_path"))
      .def("createProblem", [](const Trajectory& t) { return t.createProblem(); })
      .def("createProblem", [](const Trajectory& t, std::size_t dt, bool squash, const std::string& integ) { return t.createProblem(dt, squash, integ); },
           py::arg("dt"), py::arg("squash"), py::arg("integration_method"))
      .def("removeStage", &Trajectory::removeStage)
      .def_property_readonly("stages", &Trajectory::get_stages)
      .def_property_readonly("robot_model", &Trajectory::get_robot_model)
      .def_property_readonly("robot_model_path", &Trajectory::get_robot_model_path)
      .def_property_readonly("platform_params", &Trajectory::get_platform_params)
      .def_property_readonly("squash", &Trajectory::get_squash)
      .def_property_readonly("duration", &Trajectory::get_duration)
      .def_property("initial_state", [](const Trajectory& t) { return vec(t.get_initial_state()); },
                    [](Trajectory& t, const VectorXd& x) { t.set_initial_state(x); });

  py::class_<CallbackAbstract, std::shared_ptr<CallbackAbstract>>(m, "CallbackAbstract");
  py::class_<CallbackVerbose, CallbackAbstract, std::shared_ptr<CallbackVerbose>>(m, "CallbackVerbose").def(py::init<>());

  py::class_<SolverSbFDDP, std::shared_ptr<SolverSbFDDP>>(m, "SolverSbFDDP")
      .def(py::init<const std::shared_ptr<ShootingProblem>&, const std::shared_ptr<SquashingModelSmoothSat>&, int, int>(),
           py::arg("problem"), py::arg("squashing"), py::arg("batch") = 1, py::arg("device") = 0)
      .def("solve", [](SolverSbFDDP& s, const py::object& xs, const py::object& us, std::size_t maxiter, bool feas, double reg) {
        return s.solve(to_rows(xs), to_rows(us), maxiter, feas, reg); },
           py::arg("init_xs") = py::list(), py::arg("init_us") = py::list(), py::arg("maxiter") = 100, py::arg("isFeasible") = false,
           py::arg("regInit") = 1e-9)
      .def("setCandidate", [](SolverSbFDDP& s, const py::object& xs, const py::object& us, bool feas) { s.setCandidate(to_rows(xs), to_rows(us), feas); },
           py::arg("xs") = py::list(), py::arg("us") = py::list(), py::arg("isFeasible") = false)
      .def("setCallbacks", [](SolverSbFDDP& s, const py::list& cbs) {
        std::vector<std::shared_ptr<CallbackAbstract>> v;
        for (auto c : cbs) {
          if (py::isinstance<CallbackAbstract>(c)) v.push_back(py::cast<std::shared_ptr<CallbackAbstract>>(c));
          else v.push_back(std::make_shared<PyCallback>(py::reinterpret_borrow<py::object>(c)));
        }
        s.setCallbacks(v); })
      .def_property_readonly("xs", [](const SolverSbFDDP& s) { return rows(s.get_xs()); })
      .def_property_readonly("us", [](const SolverSbFDDP& s) { return rows(s.get_us()); })
      .def_property_readonly("us_squash", [](const SolverSbFDDP& s) { return rows(s.getSquashControls()); })
      .def("getSquashControls", [](const SolverSbFDDP& s) { return rows(s.getSquashControls()); })
      .def_property_readonly("k", [](const SolverSbFDDP& s) { return rows(s.get_k()); })
      .def_property_readonly("K", [](const SolverSbFDDP& s) {
        const auto& K = s.get_K();
        const std::size_t T = K.size(), nu = s.get_k().empty() ? 0 : s.get_k()[0].size(), ndx = nu ? K[0].size() / nu : 0;
        py::array_t<double> a({T, nu, ndx});
        auto r = a.mutable_unchecked<3>();
        for (std::size_t t = 0; t < T; ++t) for (std::size_t i = 0; i < nu; ++i) for (std::size_t j = 0; j < ndx; ++j) r(t, i, j) = K[t][i * ndx + j];
        return a; })
      .def_property_readonly("cost", &SolverSbFDDP::get_cost)
      .def_property_readonly("iter", &SolverSbFDDP::get_iter)
      .def_property_readonly("stop", &SolverSbFDDP::get_stop)
      .def_property_readonly("isFeasible", &SolverSbFDDP::get_is_feasible)
      .def_property_readonly("problem", &SolverSbFDDP::get_problem)
      .def_property("convergence_init", &SolverSbFDDP::get_convergence_init, &SolverSbFDDP::set_convergence_init)
      .def_property("stop_criteria", [](SolverSbFDDP& s) { return s.params().stop_criteria; }, [](SolverSbFDDP& s, int v) { s.params().stop_criteria = v; })
      .def_property("stop_test", [](SolverSbFDDP& s) { return s.params().stop_test; }, [](SolverSbFDDP& s, int v) { s.params().stop_test = v; })
      .def_property_readonly("handle", [](const SolverSbFDDP& s) { return (std::uintptr_t)s.handle(); });

  // crocoddyl.SolverBoxFDDP(problem) / crocoddyl.SolverBoxDDP(problem) of the reference's drivers (examples/python/trajectory.py:24,
  // examples/python/mpc.py:26): same attributes as SolverSbFDDP, constructed from a problem created with squash = False
  py::class_<SolverBoxFDDP, SolverSbFDDP, std::shared_ptr<SolverBoxFDDP>>(m, "SolverBoxFDDP")
      .def(py::init<const std::shared_ptr<ShootingProblem>&, int, int>(), py::arg("problem"), py::arg("batch") = 1, py::arg("device") = 0);
  py::class_<SolverBoxDDP, SolverSbFDDP, std::shared_ptr<SolverBoxDDP>>(m, "SolverBoxDDP")
      .def(py::init<const std::shared_ptr<ShootingProblem>&, int, int>(), py::arg("problem"), py::arg("batch") = 1, py::arg("device") = 0);

  auto mpc_base = py::class_<MpcAbstract, std::shared_ptr<MpcAbstract>>(m, "MpcAbstract")
      .def("updateProblem", [](MpcAbstract& c, std::size_t t) { c.updateProblem(t); }, py::arg("current_time"))
      .def_property_readonly("robot_model", &MpcAbstract::get_robot_model)
      .def_property_readonly("platform_params", &MpcAbstract::get_platform_params)
      .def_property_readonly("squash", &MpcAbstract::get_squash)
      .def_property_readonly("problem", &MpcAbstract::get_problem)
      .def_property_readonly("solver", &MpcAbstract::get_solver)
      .def_property_readonly("dt", &MpcAbstract::get_dt)
      .def_property_readonly("knots", &MpcAbstract::get_knots)
      .def_property_readonly("iters", &MpcAbstract::get_iters);
  (void)mpc_base;
  py::class_<CarrotMpc, MpcAbstract, std::shared_ptr<CarrotMpc>>(m, "CarrotMpc")
      .def(py::init([](const std::shared_ptr<Trajectory>& tr, const py::object& ref, std::size_t dt_ref, const std::string& yaml) {
        return std::make_shared<CarrotMpc>(tr, to_rows(ref), dt_ref, getYamlPath(yaml), true); }),
           py::arg("trajectory"), py::arg("state_ref"), py::arg("dt_ref"), py::arg("yaml_path"));
  py::class_<RailMpc, MpcAbstract, std::shared_ptr<RailMpc>>(m, "RailMpc")
      .def(py::init([](const py::object& ref, std::size_t dt_ref, const std::string& yaml) {
        return std::make_shared<RailMpc>(to_rows(ref), dt_ref, getYamlPath(yaml), true); }),
           py::arg("state_ref"), py::arg("dt_ref"), py::arg("yaml_path"));
  py::class_<WeightedMpc, MpcAbstract, std::shared_ptr<WeightedMpc>>(m, "WeightedMpc")
      .def(py::init([](const std::shared_ptr<Trajectory>& tr, std::size_t dt_ref, const std::string& yaml) {
        return std::make_shared<WeightedMpc>(tr, dt_ref, getYamlPath(yaml), true); }),
           py::arg("trajectory"), py::arg("dt_ref"), py::arg("yaml_path"));
}
