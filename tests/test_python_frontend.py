"""The eagle_mpc Python front-end (pybind11, reference names: bindings/python/eagle_mpc/*.hpp)."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eagle-mpc_b200", "python"))
import eagle_mpc  # noqa: E402
from eagle_mpc.utils.path import EAGLE_MPC_YAML_DIR  # noqa: E402

host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")


def _trajectory(rel="hexacopter370_flying_arm_3/trajectories/displacement.yaml"):
    tr = eagle_mpc.Trajectory()
    tr.autoSetup(EAGLE_MPC_YAML_DIR + "/" + rel)
    return tr


def test_reference_surface_without_a_gpu():
    tr = _trajectory()
    problem = tr.createProblem(20, True, "IntegratedActionModelEuler")
    assert problem.T == 400 and problem.nx == 19 and problem.ndx == 18
    assert np.array_equal(problem.x0, [0, 0, 0, 0, 0, 0, 1] + [0] * 12)
    assert tr.squash.ns == 9 and tr.platform_params.n_rotors == 6 and tr.platform_params.tau_f.shape == (6, 6)
    assert [s.name for s in tr.stages][:2] == ["nav_wp1", "wp_1"] and tr.stages[0].is_transition
    assert tr.robot_model.nq == 10 and tr.robot_model.getFrameId("flying_arm_3__gripper") < len(tr.robot_model.frame_names)
    x = problem.x0; x[0] = 0.5
    problem.x0 = x
    assert problem.x0[0] == 0.5
    assert tr.createProblem(20, True, "IntegratedActionModelRK4").T == 400   # both integrators of src/factory/int-action.cpp
    with pytest.raises((RuntimeError, IndexError)):   # errors of the factories surface as exceptions, like the reference's
        tr.createProblem(20, True, "IntegratedActionModelMidpoint")
    try:
        import torch
        gpu = torch.cuda.is_available()
    except ImportError:
        gpu = False
    if not gpu:
        with pytest.raises(RuntimeError):   # no CPU fallback
            eagle_mpc.SolverSbFDDP(tr.createProblem(20, True, "IntegratedActionModelEuler"), tr.squash)


@pytest.mark.gpu
def test_solver_matches_the_c_abi_and_replays_callbacks():
    tr = _trajectory("hexacopter370/trajectories/passthrough.yaml")
    problem = tr.createProblem(20, True, "IntegratedActionModelEuler")
    solver = eagle_mpc.SolverSbFDDP(problem, tr.squash)
    seen = []
    solver.setCallbacks([eagle_mpc.CallbackVerbose(), lambda rec: seen.append(rec)])
    assert solver.solve([], [], maxiter=100)
    fp = host.Trajectory("hexacopter370/trajectories/passthrough.yaml").createProblem(20)
    g = capi.BatchSolver(fp, 1)
    g.set_x0(fp.x0); g.set_candidate(None, None, False); g.solve()
    assert solver.iter == g.iters()[0] and len(seen) == solver.iter + 1
    assert np.array_equal(solver.xs, g.xs()[0]) and np.array_equal(solver.us, g.us()[0])
    assert np.array_equal(solver.us_squash, g.us_squash()[0]) and np.array_equal(solver.K, g.K()[0]) and np.array_equal(solver.k, g.k()[0])
    assert solver.cost == g.cost()[0] and seen[-1]["cost"] == solver.cost and [r["iter"] for r in seen] == list(range(len(seen)))


@pytest.mark.gpu
def test_reference_scripts_run():
    for script, args in (("reference_trajectory.py", []), ("reference_mpc.py", ["carrot", "40"]), ("reference_mpc.py", ["rail", "20"])):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script)] + args, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        assert "final position" in out.stdout


@pytest.mark.gpu
def test_contact_trajectory_through_the_front_end():
    """examples/python/trajectory.py on eagle_catch.yaml: a contact trajectory (DifferentialActionModelContactFwdDynamics
    for every stage, src/trajectory.cpp:115-116) built and solved through the reference's Python names"""
    rel = "hexacopter370_flying_arm_3/trajectories/eagle_catch.yaml"
    tr = _trajectory(rel)
    problem = tr.createProblem(20, True, "IntegratedActionModelEuler")
    solver = eagle_mpc.SolverSbFDDP(problem, tr.squash)
    assert solver.solve([], [], maxiter=100)
    fp = host.Trajectory(rel).createProblem(20)
    assert fp.desc.n_contacts == 1
    g = capi.BatchSolver(fp, 1)
    g.set_x0(fp.x0); g.set_candidate(None, None, False); g.solve()
    assert solver.iter == g.iters()[0]
    assert np.array_equal(solver.xs, g.xs()[0]) and np.array_equal(solver.us, g.us()[0])
