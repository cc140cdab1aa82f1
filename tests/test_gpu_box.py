"""GPU parity (through the C ABI) for the reference's other two SolverTypes, crocoddyl's SolverBoxFDDP / SolverBoxDDP
(include/eagle_mpc/mpc-base.hpp:36-47, src/mpc-controllers/carrot-mpc.cpp:236-241, examples/python/trajectory.py:19-24): the
problem is created without the squashing actuation, computeGains solves a box QP per node (backward_kernel<D, true, true>,
bw_box_qp) and the rollouts clamp the trial controls.  Checker: the CPU oracle (oracle/oracle.cpp box_qp, itself held to the
numpy twin in tests/test_box_oracle.py).  Phase level: K, k, Vx of box sweeps (cold and warm-started) and clamped trials;
solver level: the oracle's iteration path and solution on the YAML problems, plus the Python front-ends."""
import importlib
import os

import numpy as np
import pytest

import oracle_binding as ob
import parity

pytestmark = pytest.mark.gpu
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads")
abi = importlib.import_module("eagle-mpc_b200.abi")
EULER = "IntegratedActionModelEuler"


def rel(a, b):
    return parity.rel(a, b)


def limits(fp):
    return np.array(fp.desc.u_lb[:fp.nu]), np.array(fp.desc.u_ub[:fp.nu])


def candidate_near_limits(fp, B, seed):
    """states around the YAML initial state, every control 2 % inside one of its limits: the QP clamps some of them"""
    rng = np.random.default_rng(seed)
    T = fp.T
    lb, ub = limits(fp)
    xs = np.tile(fp.x0, (B, T + 1, 1))
    xs[:, :, :3] += rng.uniform(-0.1, 0.1, size=(B, T + 1, 3))
    q = rng.normal(size=(B, T + 1, 4)) * 0.05 + np.array([0, 0, 0, 1.0])
    xs[:, :, 3:7] = q / np.linalg.norm(q, axis=-1, keepdims=True)
    xs[:, :, 7:fp.nq] += rng.uniform(-0.1, 0.1, size=(B, T + 1, fp.nq - 7))
    xs[:, :, fp.nq:] = rng.uniform(-0.1, 0.1, size=(B, T + 1, fp.nv))
    us = np.where(rng.uniform(size=(B, T, fp.nu)) < 0.5, lb + 0.02 * (ub - lb), ub - 0.02 * (ub - lb))
    x0 = xs[:, 0].copy()
    x0[:, :3] += 0.01
    return x0, xs, us


@pytest.mark.parametrize("yaml", ["hexacopter370_flying_arm_3/trajectories/displacement.yaml", "iris/trajectories/loop.yaml",
                                  "hextilt_flying_arm_5/trajectories/push_slide.yaml"])
def test_box_phases(yaml):
    fp = host.Trajectory(yaml).createProblem(20, False, EULER)
    assert fp.desc.use_squash == 0
    B = 3
    lb, ub = limits(fp)
    x0, xs, us = candidate_near_limits(fp, B, 23)
    p = capi.box_params(abi.SOLVER_BOXFDDP)
    g = capi.BatchSolver(fp, B)
    g.set_params(p)
    g.set_x0(x0); g.set_candidate(xs, us, True)
    oracles, yard = [], []
    for b in range(B):
        for nofma, into in ((False, oracles), (True, yard)):   # yard: the oracle without FMA contraction, the rounding yardstick
            o = ob.Oracle(fp, nofma=nofma)
            o.set_params(ob.box_params(abi.SOLVER_BOXFDDP))
            o.set_x0(x0[b]); o.set_candidate(xs[b], us[b], True)
            o.phase_calc_diff(0.1)
            into.append(o)
    g.phase_calc_diff(0.1)
    n_clamped = 0
    # cold sweep (warm start k = 0), warm sweep (the first one's k), a sweep of an INFEASIBLE candidate (plain gains), and a
    # third box sweep at another regularisation
    # (xreg = 1e-2 for the box sweeps: at 1e-6 the flying arm's Quu is so ill-conditioned on this random candidate that the
    #  oracle with and without FMA contraction differ by 4e-8 on k)
    for xreg, feasible in ((1e-2, True), (1e-2, True), (1e-6, False), (1e-3, True)):
        ok = g.phase_backward(xreg, feasible)
        K, k, Vx = g.K(), g.k(), g.Vx()
        for b, o in enumerate(oracles):
            ook = o.phase_backward(xreg, feasible)
            assert ok[b] == ook, (xreg, feasible, b)
            if not ook:
                continue
            # bar: 1e-8, or 4 x the oracle's own sensitivity to rounding on this sweep (Quu at xreg = 1e-6 is ill-conditioned
            # for a random candidate 2 % off the limits over 400 nodes)
            assert yard[b].phase_backward(xreg, feasible) == ook
            for key, got in (("K", K[b]), ("k", k[b]), ("Vx", Vx[b])):
                tol = max(1e-8, 4 * rel(yard[b].get(key), o.get(key)))
                assert rel(got, o.get(key)) <= tol, (xreg, feasible, key, rel(got, o.get(key)), tol)
            if feasible:
                # the QP step keeps the controls inside the box; a clamped control has no feedback row
                assert np.all(us[b] - k[b] >= lb - 1e-12) and np.all(us[b] - k[b] <= ub + 1e-12)
                rows_gpu = np.all(K[b] == 0.0, axis=-1)
                assert np.array_equal(rows_gpu, np.all(o.get("K") == 0.0, axis=-1)), "different active sets"
                n_clamped += int(rows_gpu.sum())
    assert n_clamped > 0, "no control was clamped: the case does not exercise the box QP"
    # clamped rollouts on a tamer candidate (rotors close to full thrust, arm at rest: the trials do not spin up)
    rng = np.random.default_rng(29)
    nr = fp.desc.n_rotors
    us = np.concatenate([0.99 * ub[:nr], np.zeros(fp.nu - nr)]) + 0.004 * (ub - lb) * rng.uniform(-1, 1, size=(B, fp.T, fp.nu))
    xs = np.tile(fp.x0, (B, fp.T + 1, 1))
    xs[:, :, :3] += rng.uniform(-0.02, 0.02, size=(B, fp.T + 1, 3))
    xs[:, :, fp.nq:] = rng.uniform(-0.02, 0.02, size=(B, fp.T + 1, fp.nv))
    g.set_x0(xs[:, 0].copy()); g.set_candidate(xs, us, True)
    g.phase_calc_diff(0.1)
    assert np.all(g.phase_backward(1e-3, True) == 1)
    for b, o in enumerate(oracles):
        o.set_x0(xs[b, 0]); o.set_candidate(xs[b], us[b], True)
        o.phase_calc_diff(0.1)
        assert o.phase_backward(1e-3, True)
    n_at_limit = 0
    g.phase_rollout(0.1, True, False)
    for ai in (0, 2, 9):
        xt, ut, ct, dv, okt = g.trial(ai)
        for b, o in enumerate(oracles):
            ook = o.phase_rollout(0.1, True, False, ai)
            xo, uo = o.get("xs_try"), o.get("us_try")
            wild = np.nonzero(np.abs(xo).max(axis=1) > 500.0)[0]   # (a runaway trial amplifies rounding without bound)
            n_ok = int(wild[0]) if wild.size else fp.T + 1
            if n_ok == fp.T + 1:
                assert okt[b] == ook
            else:
                n_ok = min(n_ok, 6)   # a trial that runs away (|x| > 500 within the horizon) amplifies rounding: its first nodes only
            assert n_ok >= 3, (ai, n_ok)
            assert rel(xt[b][:n_ok], xo[:n_ok]) < 1e-8, (ai, n_ok, rel(xt[b][:n_ok], xo[:n_ok]))
            assert rel(ut[b][:n_ok - 1], uo[:n_ok - 1]) < 1e-8
            assert np.all(ut[b][:n_ok - 1] >= lb) and np.all(ut[b][:n_ok - 1] <= ub)
            n_at_limit += int(((ut[b][:n_ok - 1] == lb) | (ut[b][:n_ok - 1] == ub)).sum())
            if n_ok == fp.T + 1 and ook:
                assert rel(ct[b], o.get("cost_try")) < 1e-8
    assert n_at_limit > 0, "no trial control was clamped"


SOLVES = [("iris_loop", "iris/trajectories/loop.yaml", abi.SOLVER_BOXFDDP, 100, 5100),
          ("iris_loop", "iris/trajectories/loop.yaml", abi.SOLVER_BOXDDP, 100, 5200),
          ("iris_hover", "iris/trajectories/hover.yaml", abi.SOLVER_BOXFDDP, 100, 5300),
          ("hexacopter370_passthrough", "hexacopter370/trajectories/passthrough.yaml", abi.SOLVER_BOXFDDP, 100, 5350),
          ("flying_arm_3_displacement", "hexacopter370_flying_arm_3/trajectories/displacement.yaml", abi.SOLVER_BOXFDDP, 100, 5400),
          ("flying_arm_3_displacement", "hexacopter370_flying_arm_3/trajectories/displacement.yaml", abi.SOLVER_BOXDDP, 30, 5500),
          ("hextilt_flying_arm_5_push_slide", "hextilt_flying_arm_5/trajectories/push_slide.yaml", abi.SOLVER_BOXFDDP, 100, 5600)]


@pytest.mark.parametrize("name,yaml,solver_type,maxiter,seed", SOLVES)
def test_box_solve(name, yaml, solver_type, maxiter, seed):
    fp = host.Trajectory(yaml).createProblem(20, False, EULER)
    B = 3
    x0 = wl.noisy_x0(fp.x0, B, seed)
    x0[0] = fp.x0
    pg = capi.box_params(solver_type); pg.maxiter = maxiter
    po = ob.box_params(solver_type); po.maxiter = maxiter
    g = capi.BatchSolver(fp, B)
    g.set_params(pg)
    g.enable_iteration_log(512)
    g.set_x0(x0); g.set_candidate(None, None, False)
    g.solve()
    got = {"xs": g.xs(), "us": g.us(), "K": g.K(), "k": g.k(), "cost": g.cost(), "us_squash": g.us_squash()}
    iters, feas = g.iters(), g.feasible()
    lb, ub = limits(fp)
    assert np.all(got["us"] >= lb) and np.all(got["us"] <= ub)
    assert np.array_equal(got["us_squash"], got["us"])   # no squashing function on this path
    worst = {}
    for b in range(B):
        log = g.iteration_log(b)
        assert all(r.phase == (0 if solver_type == abi.SOLVER_BOXFDDP else 1) for r in log)
        for key, d_gpu, d_self in parity.check_ocp((name, solver_type, b), fp, x0[b], {k_: v[b] for k_, v in got.items()}, iters[b], feas[b],
                                                   params=po, log=log):
            w = worst.setdefault(key, [0.0, 0.0])
            w[0] = max(w[0], d_gpu); w[1] = max(w[1], d_self)
    at_limit = int(((got["us"] == lb) | (got["us"] == ub)).sum())
    print(name, "box solver", solver_type, "iters", iters.tolist(), "controls on a limit", at_limit,
          {k_: f"gpu {v[0]:.1e} / self {v[1]:.1e}" for k_, v in worst.items()})


def test_box_front_ends():
    """the reference driver's path (examples/python/trajectory.py:19-26 with useSquash = False): createProblem(dt, False, ...)
    + SolverBoxFDDP(problem).solve([], [], maxiter), through the ctypes facade and the pybind11 module; equal to the C ABI"""
    yaml = "iris/trajectories/loop.yaml"
    tr = host.Trajectory(yaml)
    fp = tr.createProblem(20, False, EULER)
    s = host.SolverBoxFDDP(tr, 20, EULER)
    s.solve(60)
    xs, us, uss, cost, it, fe = s.result(fp.T)
    g = capi.BatchSolver(fp, 1)
    p = capi.box_params(abi.SOLVER_BOXFDDP); p.maxiter = 60
    g.set_params(p); g.set_x0(fp.x0); g.set_candidate(None, None, False); g.solve()
    assert it == g.iters()[0] and np.array_equal(xs, g.xs()[0]) and np.array_equal(us, g.us()[0]) and cost == g.cost()[0]
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "eagle-mpc_b200", "python"))
    import eagle_mpc
    from eagle_mpc.utils.path import EAGLE_MPC_YAML_DIR
    t2 = eagle_mpc.Trajectory()
    t2.autoSetup(EAGLE_MPC_YAML_DIR + "/" + yaml)
    problem = t2.createProblem(20, False, EULER)
    sb = eagle_mpc.SolverBoxFDDP(problem)
    sb.solve([], [], 60)
    assert sb.iter == it and np.array_equal(np.asarray(sb.xs), xs) and np.array_equal(np.asarray(sb.us), us)
    with pytest.raises(Exception):
        eagle_mpc.SolverBoxFDDP(t2.createProblem(20, True, EULER))   # a squashed problem is not a box-solver problem


def test_carrot_mpc_on_box_fddp_closed_loop(tmp_path):
    """mpc.yaml with solver: "SolverBoxFDDP" (src/mpc-base.cpp:53, src/mpc-controllers/carrot-mpc.cpp:236-238): the controller builds
    the unsquashed problem, owns a SolverBoxFDDP and steps the closed loop of examples/python/mpc.py like the oracle does"""
    from test_box_oracle import box_mpc_yaml
    mpcmod = importlib.import_module("eagle-mpc_b200.mpc")
    traj = "hexacopter370_flying_arm_3/trajectories/displacement.yaml"
    tr = host.Trajectory(traj)
    fp = tr.createProblem(20, False, EULER)
    po = ob.box_params(abi.SOLVER_BOXFDDP); po.maxiter = 100
    o = ob.Oracle(fp); o.set_params(po); o.set_x0(fp.x0); o.solve()
    xs, us = o.get("xs"), o.get("us")
    yaml = box_mpc_yaml(tmp_path)
    # 12 steps: the loop is then still moving.  Once it has settled (|k| ~ 1e-6, from step ~15 on) the box QP of every node
    # leaves at its first test with the warm start (gradient 3e-7 < th_grad: traced on the device, scripts/dev/dbg_box_mpc2.py),
    # i.e. k = -k of the previous sweep, and the FDDP line search then compares costs that differ by less than the resolution
    # of a double: which step length the reference accepts is decided by the rounding of its own cost sums, and an
    # implementation with other rounding lands on the other side now and then (2e-8 on the controls when it happens)
    n_steps = 12
    mpc_g = mpcmod.CarrotMpc(tr, xs, 20, yaml, create_solver=True)
    lat_g, st_g, u_g, it_g = mpcmod.closed_loop(mpc_g, xs, us, xs[0], n_steps, record=True)
    mpc_o = mpcmod.CarrotMpc(tr, xs, 20, yaml, create_solver=False)
    lat_o, st_o, u_o, it_o = ob.oracle_closed_loop(mpc_o, xs, us, xs[0], n_steps, record=True, params=ob.box_params(abi.SOLVER_BOXFDDP))
    assert it_g == it_o
    lb, ub = limits(fp)
    assert np.all(u_g >= lb) and np.all(u_g <= ub)
    _l, st_y, u_y, _it = ob.oracle_closed_loop(mpcmod.CarrotMpc(tr, xs, 20, yaml, create_solver=False), xs, us, xs[0], n_steps, record=True,
                                               params=ob.box_params(abi.SOLVER_BOXFDDP), nofma=True)
    print("box carrot closed loop: u gpu %.1e / self %.1e, x gpu %.1e / self %.1e" % ob.closed_loop_bar(u_g, st_g, (st_o, u_o), (st_y, u_y)))


def test_box_stream_equals_plain_batch():
    """empc_solve_stream under a Box solver: every job is a fresh solver (its QP warm starts k_ = 0 whichever slot it lands in),
    bit-identical to the same OCP in a plain batch on a fresh handle"""
    fp = host.Trajectory("iris/trajectories/hover.yaml").createProblem(20, False, EULER)
    jobs, slots = 21, 4
    x0 = wl.noisy_x0(fp.x0, jobs, 5700)
    p = capi.box_params(abi.SOLVER_BOXFDDP); p.maxiter = 40
    ref = capi.BatchSolver(fp, jobs)
    ref.set_params(p); ref.set_x0(x0); ref.set_candidate(None, None, False); ref.solve()
    g = capi.BatchSolver(fp, slots)
    g.set_params(p)
    out = g.solve_stream(x0)
    assert len(set(out["iters"].tolist())) > 1, "the jobs are meant to take different numbers of iterations"
    assert np.array_equal(out["iters"], ref.iters()) and np.array_equal(out["feasible"], ref.feasible())
    assert np.array_equal(out["cost"], ref.cost()) and np.array_equal(out["xs"], ref.xs()) and np.array_equal(out["us"], ref.us())
    lb, ub = limits(fp)
    assert ((out["us"] == lb) | (out["us"] == ub)).any()


@pytest.mark.parametrize("name,yaml,integ,maxiter", [("move_arm_rk4", "hexacopter370_flying_arm_3/trajectories/move_arm.yaml", "IntegratedActionModelRK4", 30),
                                                      ("eagle_catch_contact", "hexacopter370_flying_arm_3/trajectories/eagle_catch.yaml", EULER, 12)])
def test_box_over_the_other_overlays(name, yaml, integ, maxiter):
    """SolverBoxFDDP on a problem that also uses an overlay node model — createProblem(dt, False, "IntegratedActionModelRK4"), or
    a trajectory with a contact stage: the box sweep (which carries the Lxu / dense Luu terms those nodes produce) and the
    clamped rollout (which calls their dynamics) against the oracle's iteration path"""
    fp = host.Trajectory(yaml).createProblem(20, False, integ)
    B = 2
    x0 = np.tile(fp.x0, (B, 1))
    x0[1] = wl.noisy_x0(fp.x0, 1, 5800)[0]
    pg = capi.box_params(abi.SOLVER_BOXFDDP); pg.maxiter = maxiter
    po = ob.box_params(abi.SOLVER_BOXFDDP); po.maxiter = maxiter
    g = capi.BatchSolver(fp, B)
    g.set_params(pg); g.enable_iteration_log(512)
    g.set_x0(x0); g.set_candidate(None, None, False); g.solve()
    got = {"xs": g.xs(), "us": g.us(), "K": g.K(), "k": g.k(), "cost": g.cost(), "us_squash": g.us_squash()}
    iters, feas = g.iters(), g.feasible()
    lb, ub = limits(fp)
    assert np.all(got["us"] >= lb) and np.all(got["us"] <= ub)
    worst = {}
    for b in range(B):
        for key, d_gpu, d_self in parity.check_ocp((name, b), fp, x0[b], {k_: v[b] for k_, v in got.items()}, iters[b], feas[b], params=po,
                                                   log=g.iteration_log(b), perturb=1e-12 if "contact" in name else 0.0):
            w = worst.setdefault(key, [0.0, 0.0])
            w[0] = max(w[0], d_gpu); w[1] = max(w[1], d_self)
    print(name, "box iters", iters.tolist(), {k_: f"gpu {v[0]:.1e} / self {v[1]:.1e}" for k_, v in worst.items()})


GB = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "twin_box.npz"))


@pytest.mark.parametrize("ci", range(int(GB["n_cases"])))
def test_box_sweep_equals_twin_golden(ci):
    """backward_kernel<D, true, true> against gains generated by the independent numpy twin (tests/golden/twin_box.npz, made by
    scripts/make_twin_box_golden.py: complex-step node blocks, dense Riccati recursion, its own projected-Newton box QP) over
    the last nodes of the horizon, cold and warm-started.  No oracle code runs here: the fixture is the checker."""
    key = f"c{ci}"
    yaml, dt = str(GB[key + "_yaml"]), int(GB[key + "_dt"])
    tail, xreg, smooth = int(GB["tail"]), float(GB["xreg"]), float(GB["smooth"])
    fp = host.Trajectory(yaml).createProblem(dt, False, EULER)
    xs, us = GB[key + "_xs"], GB[key + "_us"]
    g = capi.BatchSolver(fp, 2)
    g.set_params(capi.box_params(abi.SOLVER_BOXFDDP))
    g.set_x0(np.stack([xs[0], xs[0]])); g.set_candidate(np.stack([xs, xs]), np.stack([us, us]), True)
    g.phase_calc_diff(smooth)
    T0 = fp.T - tail
    for sweep in range(2):
        # (sweep 1 is warm-started by sweep 0's k, on the device as in the fixture; the nodes before the tail do not enter it)
        assert np.all(g.phase_backward(xreg, True) == 1)
        K, k, Vx = g.K()[1], g.k()[1], g.Vx()[1]
        for name, got in (("K", K[T0:]), ("k", k[T0:]), ("Vx", Vx[T0:fp.T])):
            ref = GB[f"{key}_s{sweep}_{name}"]
            assert rel(got, ref) <= 1e-9, (yaml, sweep, name, rel(got, ref))
        assert np.array_equal(np.all(K[T0:] == 0.0, axis=-1), np.all(GB[f"{key}_s{sweep}_K"] == 0.0, axis=-1)), "different active sets"
