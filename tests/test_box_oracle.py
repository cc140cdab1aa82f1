"""SolverBoxFDDP / SolverBoxDDP of the oracle (oracle/oracle.cpp: box_qp, the box branch of backward_pass, clamped rollouts)
against the independent numpy twin (oracle/twin.py: box_qp, riccati_sweep(box=...), rollout(clamp=...)) and against the
optimality conditions of the box QP themselves.  CPU only.

The reference selects these solvers by SolverTypes (include/eagle_mpc/mpc-base.hpp:36-47, src/mpc-controllers/carrot-mpc.cpp:236-241,
examples/python/trajectory.py:24); they live in Crocoddyl, which is not in the reference tree, so the oracle restates the
published algorithm (Tassa's projected-Newton box QP as crocoddyl::BoxQP runs it) and the twin restates it a second time.
"""
import importlib
import os
import sys

import numpy as np
import pytest

import oracle_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import twin  # noqa: E402

host = importlib.import_module("eagle-mpc_b200.host")
abi = importlib.import_module("eagle-mpc_b200.abi")
YAML_ROOT = os.path.join(ROOT, "yaml")
URDF_ROOT = os.path.join(ROOT, "fixtures", "urdf")


def rel_err(a, b):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def random_state(tw, rng, amp):
    rob = tw.rob
    dx = rng.uniform(-amp, amp, size=rob.ndx)
    x = np.zeros(rob.nq + rob.nv); x[6] = 1.0
    return twin.integrate(rob, x, dx)


def blocks_of(tile, fp_o):
    """the oracle's node tile -> Fx, Fu, Lx, Lu, Lxx, Lxu, Luu (layout: DESIGN.md section 3)"""
    ndx, nu = fp_o.ndx, fp_o.nu
    o = 0
    out = {}
    for key, shape in (("Fx", (ndx, ndx)), ("Fu", (ndx, nu)), ("Lxx", (ndx, ndx)), ("Lxu", (ndx, nu)), ("Luu", (nu, nu)), ("Lx", (ndx,)), ("Lu", (nu,))):
        n = int(np.prod(shape))
        out[key] = tile[o:o + n].reshape(shape); o += n
    return out


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_box_qp_optimality_conditions(seed):
    """x* of the twin's and the oracle's box QP: inside the box, gradient ~ 0 on the free set, pushing outwards on the clamped"""
    rng = np.random.default_rng(seed)
    n = 9
    A = rng.normal(size=(n, n)); H = A @ A.T + 0.1 * np.eye(n)
    q = rng.normal(size=n) * 3
    lb, ub = -rng.uniform(0.05, 0.5, n), rng.uniform(0.05, 0.5, n)
    x, free, clamped, Hff_inv = twin.box_qp(H, q, lb, ub, np.zeros(n))
    g = q + H @ x
    assert np.all(x >= lb) and np.all(x <= ub)
    assert clamped.size > 0, "the case is meant to clamp"
    assert np.abs(g[free]).max() <= 1e-5
    for j in clamped:
        assert (x[j] == lb[j] and g[j] > 0) or (x[j] == ub[j] and g[j] < 0)
    assert np.allclose(Hff_inv @ H[np.ix_(free, free)], np.eye(free.size), atol=1e-9)
    # brute-force check on a projected-gradient descent of the same problem
    y = np.zeros(n); L = np.linalg.eigvalsh(H).max()
    for _ in range(20000):
        y = np.clip(y - (q + H @ y) / L, lb, ub)
    assert np.abs(x - y).max() <= 1e-5


@pytest.mark.parametrize("rel,dt", [("hexacopter370_flying_arm_3/trajectories/displacement.yaml", 20), ("iris/trajectories/loop.yaml", 20)])
def test_box_backward_and_rollout_equal_twin(rel, dt):
    fp = host.Trajectory(rel).createProblem(dt, False, "IntegratedActionModelEuler")
    assert fp.desc.use_squash == 0
    tw = twin.Problem(rel, YAML_ROOT, URDF_ROOT, dt, use_squash=False)
    rob, T = tw.rob, 10
    rng = np.random.default_rng(5)
    xreg = 1e-6
    x0 = random_state(tw, rng, 0.1)
    xs_full = np.array([random_state(tw, rng, 0.1) for _ in range(fp.T + 1)])
    # controls close to the limits so that the QP clamps some of them
    us_full = np.where(rng.uniform(size=(fp.T, tw.nu)) < 0.5, tw.u_lb + 0.02 * (tw.u_ub - tw.u_lb), tw.u_ub - 0.02 * (tw.u_ub - tw.u_lb))
    o = ob.Oracle(fp)
    p = ob.box_params(abi.SOLVER_BOXFDDP)
    o.set_params(p)
    o.set_x0(x0); o.set_candidate(xs_full, us_full, True)
    o.phase_calc_diff(0.1)
    T0 = fp.T - T
    nodes = [tw.calc_diff(tw.node_stage[t], xs_full[t], us_full[t], 0.1) for t in range(T0, fp.T)]
    term = tw.calc_diff(tw.node_stage[fp.T], xs_full[fp.T], None, 0.1, terminal=True)
    fs = [np.zeros(rob.ndx)] * (T + 1)
    k_prev = [np.zeros(tw.nu)] * T
    n_clamped = 0
    for sweep in range(2):   # the second sweep is warm-started by the first one's k
        assert o.phase_backward(xreg, True)
        K, k, Vx, Vxx, Qus = twin.riccati_sweep(nodes, term, fs, xreg, True, box=(us_full[T0:], tw.u_lb, tw.u_ub, k_prev))
        Ko, ko, Vxo, Vxxo, Quo = o.get("K"), o.get("k"), o.get("Vx"), o.get("Vxx"), o.get("Qu")
        for t in range(T):
            assert rel_err(Ko[T0 + t], K[t]) <= 1e-9, (sweep, t, rel_err(Ko[T0 + t], K[t]))
            assert rel_err(ko[T0 + t], k[t]) <= 1e-9
            assert rel_err(Quo[T0 + t], Qus[t]) <= 1e-9
            assert rel_err(Vxo[T0 + t], Vx[t]) <= 1e-9 and rel_err(Vxxo[T0 + t], Vxx[t]) <= 1e-9
            # the step stays inside the box and clamped controls have no feedback
            du = -ko[T0 + t]
            assert np.all(us_full[T0 + t] + du >= tw.u_lb - 1e-12) and np.all(us_full[T0 + t] + du <= tw.u_ub + 1e-12)
            cl = np.flatnonzero(Quo[T0 + t] == 0.0)
            n_clamped += cl.size
            assert np.all(Ko[T0 + t][cl] == 0.0)
        k_prev = k
    assert n_clamped > 0, "no control was clamped: the case does not exercise the box QP"
    # clamped rollout on a tamer candidate (all rotors close to full thrust, arm at rest: no spin-up): oracle vs twin
    xs_r = np.array([random_state(tw, rng, 0.02) for _ in range(fp.T + 1)])
    us_r = np.concatenate([0.99 * tw.u_ub[:tw.nr], np.zeros(tw.nu - tw.nr)]) + 0.004 * (tw.u_ub - tw.u_lb) * rng.uniform(-1, 1, size=(fp.T, tw.nu))
    o.set_candidate(xs_r, us_r, True)
    o.phase_calc_diff(0.1)
    assert o.phase_backward(1e-3, True)
    o.phase_rollout(0.1, True, False, 0)   # (the candidate may still run away further down the horizon: the first nodes are compared)
    xt_o, ut_o = o.get("xs_try"), o.get("us_try")
    n_chk = 8
    assert np.abs(xt_o[:n_chk]).max() < 1e2
    Ko, ko = o.get("K"), o.get("k")

    class Head:
        pass
    hd = Head(); hd.rob, hd.calc, hd.node_stage = tw.rob, tw.calc, tw.node_stage[:n_chk] + [tw.node_stage[n_chk]]
    xs_t, us_t, _ = twin.rollout(hd, x0, xs_r[:n_chk + 1], us_r[:n_chk], Ko[:n_chk], ko[:n_chk], np.zeros((n_chk + 1, rob.ndx)), 1.0, 0.1, True,
                                 clamp=(tw.u_lb, tw.u_ub))
    assert rel_err(xt_o[:n_chk], xs_t[:n_chk]) <= 1e-9 and rel_err(ut_o[:n_chk], us_t[:n_chk]) <= 1e-9
    assert np.all(ut_o[:n_chk] >= tw.u_lb) and np.all(ut_o[:n_chk] <= tw.u_ub)
    assert ((ut_o[:n_chk] == tw.u_lb) | (ut_o[:n_chk] == tw.u_ub)).any(), "no trial control was clamped"


@pytest.mark.parametrize("solver_type", [abi.SOLVER_BOXFDDP, abi.SOLVER_BOXDDP])
def test_box_solve_properties(solver_type):
    """a full oracle solve: controls inside the limits with some of them ON the limits, feasible, cost below the first
    iteration's, one pass (no smoothing schedule, no clean-up phase)"""
    fp = host.Trajectory("iris/trajectories/loop.yaml").createProblem(20, False, "IntegratedActionModelEuler")
    o = ob.Oracle(fp)
    p = ob.box_params(solver_type); p.maxiter = 60
    o.set_params(p); o.set_x0(np.array(fp.x0)); o.solve(None, None)
    us = o.get("us"); nu = us.shape[1]
    lb, ub = np.array(fp.desc.u_lb[:nu]), np.array(fp.desc.u_ub[:nu])
    assert np.all(us >= lb) and np.all(us <= ub)
    assert ((us == lb) | (us == ub)).sum() > 0
    log = o.iteration_log()
    assert len(log) == int(o.get("iter")) + 1 <= 60
    assert all(r.phase == (0 if solver_type == abi.SOLVER_BOXFDDP else 1) for r in log)
    assert o.get("feasible") == 1.0 and log[-1].cost < log[0].cost
    assert np.array_equal(o.get("us_squash"), us)


def box_mpc_yaml(tmp_path, solver="SolverBoxFDDP"):
    """the flying arm's mpc.yaml with the solver entry switched (mpc_controller/solver, src/mpc-base.cpp:53)"""
    src = open(os.path.join(YAML_ROOT, "hexacopter370_flying_arm_3/mpc/mpc.yaml")).read()
    assert 'solver: "SolverSbFDDP"' in src
    out = tmp_path / ("mpc_" + solver + ".yaml")
    out.write_text(src.replace('solver: "SolverSbFDDP"', 'solver: "%s"' % solver))
    return str(out)


@pytest.mark.parametrize("solver", ["SolverBoxFDDP", "SolverBoxDDP"])
def test_mpc_controller_with_a_box_solver_builds_the_unsquashed_problem(solver, tmp_path):
    """src/mpc-controllers/carrot-mpc.cpp:188-193: the squashing actuation only under SolverSbFDDP; no barrier cost either
    (barrierInit belongs to SolverSbFDDP)"""
    mpcmod = importlib.import_module("eagle-mpc_b200.mpc")
    tr = host.Trajectory("hexacopter370_flying_arm_3/trajectories/displacement.yaml")
    xs = np.tile(np.array([0, 0, 0, 0, 0, 0, 1.0] + [0.0] * 12), (50, 1))
    mpc = mpcmod.CarrotMpc(tr, xs, 20, box_mpc_yaml(tmp_path, solver), create_solver=False)
    assert mpc.desc.use_squash == 0
    assert abi.COST_SQUASH_BARRIER not in {c.type for c in mpc.cost_tables()[0]}
    ref = mpcmod.CarrotMpc(tr, xs, 20, "hexacopter370_flying_arm_3/mpc/mpc.yaml", create_solver=False)
    assert ref.desc.use_squash == 1 and abi.COST_SQUASH_BARRIER in {c.type for c in ref.cost_tables()[0]}


GB = np.load(os.path.join(ROOT, "tests", "golden", "twin_box.npz"))


@pytest.mark.parametrize("ci", range(int(GB["n_cases"])))
def test_oracle_box_sweep_equals_twin_golden(ci):
    """the committed twin fixture of the box sweeps (scripts/make_twin_box_golden.py) against the oracle"""
    key = f"c{ci}"
    yaml, dt = str(GB[key + "_yaml"]), int(GB[key + "_dt"])
    tail, xreg, smooth = int(GB["tail"]), float(GB["xreg"]), float(GB["smooth"])
    fp = host.Trajectory(yaml).createProblem(dt, False, "IntegratedActionModelEuler")
    xs, us = GB[key + "_xs"], GB[key + "_us"]
    o = ob.Oracle(fp)
    o.set_params(ob.box_params(abi.SOLVER_BOXFDDP))
    o.set_x0(xs[0]); o.set_candidate(xs, us, True)
    o.phase_calc_diff(smooth)
    T0 = fp.T - tail
    for sweep in range(2):
        assert o.phase_backward(xreg, True)
        for name, got in (("K", o.get("K")[T0:]), ("k", o.get("k")[T0:]), ("Vx", o.get("Vx")[T0:fp.T])):
            assert rel_err(got, GB[f"{key}_s{sweep}_{name}"]) <= 1e-9, (yaml, sweep, name, rel_err(got, GB[f"{key}_s{sweep}_{name}"]))
