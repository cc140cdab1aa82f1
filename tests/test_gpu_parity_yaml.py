"""GPU vs oracle on the reference's named YAML problems (synthetic URDFs), through the C ABI.

Bar (BASELINE.json north_star): identical iteration count and stopping decision, <= 1e-9 relative on cost, xs, us, K, k.
"""
import importlib

import numpy as np
import pytest

import oracle_binding as ob
import parity

pytestmark = pytest.mark.gpu
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads")



# Bar and yardstick: tests/parity.py (identical iteration count / feasibility, <= 1e-9 relative on cost, xs, us, K, k,
# us_squash, scaled by the oracle's own FMA / no-FMA sensitivity where a problem is ill-conditioned; every key is always
# checked against a bounded tolerance).  On the BASELINE.json batches (flying_arm_3 displacement, hextilt push_slide)
# d_self < 1e-9, so the plain 1e-9 bar applies; iris_px4 hover (yaw nearly unobservable, d_self ~ 7e-4) and perturbed
# hover starts (100+ crawling iterations) are the ill-conditioned ones (DESIGN.md "Parity").
@pytest.mark.parametrize("name,B", [("hexacopter370_hover", 4), ("hexacopter370_passthrough", 3),
                                    ("hexacopter370_flying_arm_3_displacement", 4),
                                    ("hextilt_flying_arm_5_push_slide", 4), ("iris_px4_hover", 3),
                                    ("iris_px4_displacement", 2)])
def test_named_problem(name, B):
    yaml, dt, seed0 = wl.CONFIGS[name]
    tr = host.Trajectory(yaml)
    fp = tr.createProblem(dt)
    x0 = wl.noisy_x0(fp.x0, B, seed0)
    x0[0] = fp.x0  # OCP 0 is the reference's own single-OCP case (unperturbed initial state)
    g = capi.BatchSolver(fp, B)
    g.enable_iteration_log(512)
    g.set_x0(x0)
    g.set_candidate(None, None, False)
    g.solve()
    got = {"xs": g.xs(), "us": g.us(), "K": g.K(), "k": g.k(), "cost": g.cost(), "us_squash": g.us_squash()}
    iters, feas = g.iters(), g.feasible()
    worst = {}
    for b in range(B):
        for key, d_gpu, d_self in parity.check_ocp((name, b), fp, x0[b], {k_: v[b] for k_, v in got.items()}, iters[b], feas[b],
                                                       log=g.iteration_log(b)):
            w = worst.setdefault(key, [0.0, 0.0])
            w[0] = max(w[0], d_gpu); w[1] = max(w[1], d_self)
    print(name, "iters", iters.tolist(), {k_: f"gpu {v[0]:.1e} / self {v[1]:.1e}" for k_, v in worst.items()})


@pytest.mark.parametrize("criteria,test", [(1, 1), (1, 0), (0, 1)])
@pytest.mark.parametrize("name", ["hexacopter370_passthrough", "hextilt_flying_arm_5_push_slide"])
def test_stop_rule_policies(name, criteria, test):
    """set_stoppingCriteria / set_stoppingTest (src/sbfddp.cpp:28-29): upstream's sum ||Qu||^2 criterion and feasibility
    test are selectable beside the fork's (inferred) cost-reduction / gaps pair; same decisions as the oracle."""
    yaml, dt, seed0 = wl.CONFIGS[name]
    fp = host.Trajectory(yaml).createProblem(dt)
    B = 2
    x0 = wl.noisy_x0(fp.x0, B, seed0 + 77)
    p = capi.default_params()
    p.stop_criteria, p.stop_test, p.maxiter = criteria, test, 40
    g = capi.BatchSolver(fp, B)
    g.set_params(p)
    g.set_x0(x0); g.set_candidate(None, None, False); g.solve()
    got = {"xs": g.xs(), "us": g.us(), "K": g.K(), "k": g.k(), "cost": g.cost(), "us_squash": g.us_squash(), "stop": g.stop()}
    iters, feas = g.iters(), g.feasible()
    po = ob.default_params()
    po.stop_criteria, po.stop_test, po.maxiter = criteria, test, 40
    for b in range(B):
        parity.check_ocp((name, criteria, test, b), fp, x0[b], {k_: v[b] for k_, v in got.items()}, iters[b], feas[b],
                         params=po, keys=parity.KEYS + ("stop",))
    print(name, "criteria", criteria, "test", test, "iters", iters.tolist(), "stop", got["stop"].tolist())
