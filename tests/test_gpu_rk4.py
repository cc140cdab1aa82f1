"""GPU parity (through the C ABI) for IntegratedActionModelRK4 inside the OCP (src/factory/int-action.cpp:29-31;
createProblem(dt, squash, "IntegratedActionModelRK4")), against the CPU oracle (itself held to the twin's complex step in
tests/test_twin.py).  Phase level: tiles (Lxu and the dense Luu of the RK4 pull-back included) / xnext / node cost / gaps,
backward K, k, Vx, rollout trials; solver level: the oracle's iteration path and solution."""
import importlib

import numpy as np
import pytest

import oracle_binding as ob
import parity
from test_gpu_contact import random_candidate, rel

pytestmark = pytest.mark.gpu
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads")
abi = importlib.import_module("eagle-mpc_b200.abi")
RK4 = "IntegratedActionModelRK4"


@pytest.mark.parametrize("yaml", ["hexacopter370_flying_arm_3/trajectories/move_arm.yaml", "hexacopter370/trajectories/hover.yaml",
                                  "hextilt_flying_arm_5/trajectories/push_slide.yaml"])
def test_rk4_phases(yaml):
    fp = host.Trajectory(yaml).createProblem(20, True, RK4)
    assert fp.desc.integrator == abi.INTEGRATOR_RK4
    B = 2
    x0, xs, us = random_candidate(fp, B, 11)
    g = capi.BatchSolver(fp, B)
    g.set_x0(x0); g.set_candidate(xs, us, False)
    oracles = []
    for b in range(B):
        o = ob.Oracle(fp)
        o.set_x0(x0[b]); o.set_candidate(xs[b], us[b], False)
        oracles.append(o)
    smooth = 0.1
    g.phase_calc_diff(smooth)
    tiles, xnext, ncost, gaps = g.tiles(), g.xnext(), g.node_cost(), g.gaps()
    off = fp.tile_offsets()
    worst = {}
    for b, o in enumerate(oracles):
        o.phase_calc_diff(smooth)
        worst["xnext"] = max(worst.get("xnext", 0), rel(xnext[b, :-1], o.get("xnext")[:-1]))
        worst["cost"] = max(worst.get("cost", 0), rel(ncost[b], o.get("node_cost")))
        worst["fs"] = max(worst.get("fs", 0), rel(gaps[b], o.get("fs")))
        ot = o.get("tiles")
        for name, size in (("Fx", fp.ndx * fp.ndx), ("Fu", fp.ndx * fp.nu), ("Lxx", fp.ndx * fp.ndx), ("Lxu", fp.ndx * fp.nu),
                           ("Luu", fp.nu * fp.nu), ("Lx", fp.ndx), ("Lu", fp.nu)):
            a = tiles[b, :-1, off[name]:off[name] + size]; c = ot[:-1, off[name]:off[name] + size]
            worst[name] = max(worst.get(name, 0), rel(a, c))
        # the terminal node contributes its cost, Lx and Lxx only (SolverDDP::backwardPass)
        for name, size in (("Lxx", fp.ndx * fp.ndx), ("Lx", fp.ndx)):
            worst[name] = max(worst[name], rel(tiles[b, -1, off[name]:off[name] + size], ot[-1, off[name]:off[name] + size]))
    print(yaml, {k: f"{v:.1e}" for k, v in worst.items()})
    for k, v in worst.items():
        assert v < 1e-10, (k, v)
    for feasible in (False, True):
        ok = g.phase_backward(1e-6, feasible)
        K, k, Vx = g.K(), g.k(), g.Vx()
        for b, o in enumerate(oracles):
            ook = o.phase_backward(1e-6, feasible)
            assert ok[b] == ook
            if not ook:
                continue
            assert rel(K[b], o.get("K")) < 1e-8, rel(K[b], o.get("K"))
            assert rel(k[b], o.get("k")) < 1e-8
            assert rel(Vx[b], o.get("Vx")) < 1e-8
    g.phase_backward(1e-6, False)
    for o in oracles:
        o.phase_backward(1e-6, False)
    for feasible, ddp in ((False, False), (True, False)):
        g.phase_rollout(smooth, feasible, ddp)
        for ai in (0, 3, 9):
            xt, ut, ct, dv, okt = g.trial(ai)
            for b, o in enumerate(oracles):
                ook = o.phase_rollout(smooth, feasible, ddp, ai)
                assert okt[b] == ook
                if not ook:
                    continue
                xo, uo = o.get("xs_try"), o.get("us_try")
                wild = np.nonzero(np.abs(xo).max(axis=1) > 50.0)[0]   # (a runaway trial amplifies rounding without bound)
                n_ok = int(wild[0]) if wild.size else fp.T + 1
                assert n_ok >= 10, (ai, n_ok)
                assert rel(xt[b][:n_ok], xo[:n_ok]) < 1e-8, (ai, n_ok, rel(xt[b][:n_ok], xo[:n_ok]))
                assert rel(ut[b][:n_ok - 1], uo[:n_ok - 1]) < 1e-8
                if n_ok == fp.T + 1:
                    assert rel(ct[b], o.get("cost_try")) < 1e-8, (ai, ct[b], o.get("cost_try"))


# (flying_arm_3 displacement.yaml is left out on purpose: its third waypoint asks for a 180 degree yaw, and from the zero
#  guess the RK4 stage states sit on the cut of log3 with a yaw of +-1e-17 — rounding noise decides the sign of that
#  cost gradient, so two implementations of the same mathematics take different first steps.  Under Euler the cost is
#  evaluated at the guess itself (yaw exactly 0) and the path is reproducible: tests/test_gpu_parity_yaml.py.)
SOLVES = {"hexacopter370_passthrough": ("hexacopter370/trajectories/passthrough.yaml", 1500),
          "flying_arm_3_move_arm": ("hexacopter370_flying_arm_3/trajectories/move_arm.yaml", 2100),
          "hextilt_flying_arm_5_push_slide": ("hextilt_flying_arm_5/trajectories/push_slide.yaml", 4096)}


@pytest.mark.parametrize("name,B", [("hexacopter370_passthrough", 2), ("flying_arm_3_move_arm", 2), ("hextilt_flying_arm_5_push_slide", 2)])
def test_rk4_solve(name, B):
    yaml, seed0 = SOLVES[name]
    dt = 20
    fp = host.Trajectory(yaml).createProblem(dt, True, RK4)
    x0 = wl.noisy_x0(fp.x0, B, seed0)
    x0[0] = fp.x0
    g = capi.BatchSolver(fp, B)
    g.enable_iteration_log(512)
    g.set_x0(x0); g.set_candidate(None, None, False)
    g.solve()
    got = {"xs": g.xs(), "us": g.us(), "K": g.K(), "k": g.k(), "cost": g.cost(), "us_squash": g.us_squash()}
    iters, feas = g.iters(), g.feasible()
    worst = {}
    for b in range(B):
        for key, d_gpu, d_self in parity.check_ocp((name, b), fp, x0[b], {k_: v[b] for k_, v in got.items()}, iters[b], feas[b],
                                                       log=g.iteration_log(b)):
            w = worst.setdefault(key, [0.0, 0.0])
            w[0] = max(w[0], d_gpu); w[1] = max(w[1], d_self)
    print(name, "RK4 iters", iters.tolist(), {k_: f"gpu {v[0]:.1e} / self {v[1]:.1e}" for k_, v in worst.items()})
