"""eagle_mpc — the reference's Python package name and surface (bindings/python/eagle_mpc/__init__.py) over the B200 path.

    import eagle_mpc
    trajectory = eagle_mpc.Trajectory(); trajectory.autoSetup(yaml)
    problem = trajectory.createProblem(dt, True, "IntegratedActionModelEuler")
    solver = eagle_mpc.SolverSbFDDP(problem, trajectory.squash)
    solver.setCallbacks([eagle_mpc.CallbackVerbose()])     # crocoddyl.CallbackVerbose() in the reference's scripts
    solver.solve([], [], maxiter=100)

    problem = trajectory.createProblem(dt, False, "IntegratedActionModelEuler")   # useSquash = False
    solver = eagle_mpc.SolverBoxFDDP(problem)               # crocoddyl.SolverBoxFDDP(problem) in the reference's scripts

The extension module (_eagle_mpc, pybind11 over eagle-mpc_b200/host/) binds the CUDA library at run time; there is no CPU
fallback: constructing a solver without a usable GPU raises RuntimeError."""
import os as _os

_PKG = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # eagle-mpc_b200/
_os.environ.setdefault("EMPC_LIB", _os.path.join(_PKG, "lib", "libempc_b200.so"))

from ._eagle_mpc import (CallbackVerbose, CarrotMpc, MpcAbstract, MultiCopterBaseParams, RailMpc, RobotModel,  # noqa: E402,F401
                         ShootingProblem, SolverBoxDDP, SolverBoxFDDP, SolverSbFDDP, SquashingModelSmoothSat, Stage, Trajectory, WeightedMpc,
                         set_robot_data_dir, set_yaml_dir)
from . import utils  # noqa: E402,F401

set_yaml_dir(utils.path.EAGLE_MPC_YAML_DIR)
set_robot_data_dir(utils.path.EAGLE_MPC_ROBOT_DATA_DIR)
