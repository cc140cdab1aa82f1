"""GPU path against the committed golden fixtures (tests/golden/oracle_named_problems.json, written by
tests/golden/make_golden.py from the CPU oracle): no oracle code runs here, only the C ABI and the JSON numbers."""
import importlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads")

# iris_px4_hover is ill-conditioned (DESIGN.md "Parity": the oracle does not reproduce itself across FMA settings there),
# so its trajectories are not compared to fixed numbers; every other named problem is held to the 1e-9 bar.
NAMES = ["hexacopter370_hover", "hexacopter370_passthrough", "hexacopter370_flying_arm_3_displacement",
         "hextilt_flying_arm_5_push_slide", "iris_px4_displacement"]


@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_golden(name):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_named_problems.json")))[name]
    yaml, dt, _ = wl.CONFIGS[name]
    fp = host.Trajectory(yaml).createProblem(dt)
    recs = gold["ocps"]
    x0 = np.array([r["x0"] for r in recs])
    g = capi.BatchSolver(fp, len(recs))
    assert (g.T, g.nx, g.nu) == (gold["T"], gold["nx"], gold["nu"])
    g.set_x0(x0); g.set_candidate(None, None, False); g.solve()
    xs, us, K, cost, iters, feas = g.xs(), g.us(), g.K(), g.cost(), g.iters(), g.feasible()
    # the packed one-call getter returns the same bits as the individual ones
    for a, b_ in zip(g.solution(), (xs, us, g.us_squash(), cost, g.stop(), iters, feas)):
        assert np.array_equal(a, b_)
    for b, r in enumerate(recs):
        if name == "hexacopter370_hover" and b > 0:
            # perturbed hover starts crawl for 30+ iterations: iteration count pinned, values to 1e-8
            tol = 1e-8
        else:
            tol = 1e-9
        assert iters[b] == r["iter"] and feas[b] == r["feasible"], (name, b, iters[b], r["iter"])
        assert abs(cost[b] - r["cost"]) <= tol * max(1.0, abs(r["cost"]))
        sx = max(1.0, np.abs(xs[b]).max()); su = max(1.0, np.abs(us[b]).max())
        assert np.abs(xs[b, -1] - np.array(r["xs_T"])).max() <= tol * sx
        assert np.abs(xs[b, g.T // 2] - np.array(r["xs_mid"])).max() <= tol * sx
        for t, u in r["us_samples"].items():
            assert np.abs(us[b, int(t)] - np.array(u)).max() <= tol * su, (name, b, t)
        for t, kf in r["K_fro"].items():
            assert abs(np.linalg.norm(K[b, int(t)]) - kf) <= 1e-8 * max(1.0, kf), (name, b, t)
        assert abs(xs[b].sum() - r["xs_sum"]) <= tol * sx * xs[b].size
        assert abs(us[b].sum() - r["us_sum"]) <= tol * su * us[b].size
