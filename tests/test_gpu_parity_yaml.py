"""GPU vs oracle on the reference's named YAML problems (synthetic URDFs), through the C ABI.

Bar (BASELINE.json north_star): identical iteration count and stopping decision, <= 1e-9 relative on cost, xs, us, K, k.
"""
import importlib

import numpy as np
import pytest

import oracle_binding as ob

pytestmark = pytest.mark.gpu
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads")

TOL = 1e-9


def rel(a, b):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


# (name, batch, tolerance for the perturbed OCPs).  OCP 0 is always the reference's own case (YAML initial state) and is
# held to 1e-9, as are the BASELINE.json batches (configs 2 and 4).  The perturbed *hover* variants are not named
# configs: from a tilted, moving start the algorithm (as the reference defines it) crawls with alpha = 1/16..1/32 and
# its DDP clean-up phase accepts a full step that multiplies the cost by ~100 before recovering over 100+ iterations;
# rounding-level differences are amplified along such paths, so only the decisions (iteration counts, feasibility) and a
# loose tolerance are asserted there (DESIGN.md "Parity").
@pytest.mark.parametrize("name,B,tol_rest", [("hexacopter370_hover", 4, 1e-6), ("hexacopter370_passthrough", 3, 1e-9),
                                             ("hexacopter370_flying_arm_3_displacement", 4, 1e-9),
                                             ("hextilt_flying_arm_5_push_slide", 4, 1e-9), ("iris_px4_hover", 3, 2e-2),
                                             ("iris_px4_displacement", 2, 1e-9)])
def test_named_problem(name, B, tol_rest):
    yaml, dt, seed0 = wl.CONFIGS[name]
    tr = host.Trajectory(yaml)
    fp = tr.createProblem(dt)
    x0 = wl.noisy_x0(fp.x0, B, seed0)
    x0[0] = fp.x0  # OCP 0 is the reference's own single-OCP case (unperturbed initial state)
    g = capi.BatchSolver(fp, B)
    g.set_x0(x0)
    g.set_candidate(None, None, False)
    g.solve()
    xs, us, K, k, cost, iters, feas, uss = g.xs(), g.us(), g.K(), g.k(), g.cost(), g.iters(), g.feasible(), g.us_squash()
    worst = {}
    worst0 = {}
    for b in range(B):
        o = ob.Oracle(fp)
        o.set_x0(x0[b])
        o.solve()
        assert int(o.get("iter")) == iters[b], (name, b, o.get("iter"), iters[b])
        assert int(o.get("feasible")) == feas[b]
        for key, a, c in (("cost", cost[b], o.get("cost")), ("xs", xs[b], o.get("xs")), ("us", us[b], o.get("us")),
                          ("K", K[b], o.get("K")), ("k", k[b], o.get("k")), ("us_squash", uss[b], o.get("us_squash"))):
            w = worst0 if b == 0 else worst
            w[key] = max(w.get(key, 0.0), rel(a, c))
    print(name, "ocp0", {k_: f"{v:.2e}" for k_, v in worst0.items()}, "rest", {k_: f"{v:.2e}" for k_, v in worst.items()},
          "iters", iters.tolist())
    for key, v in worst0.items():
        assert v < TOL, (name, key, v)
    for key, v in worst.items():
        assert v < tol_rest, (name, key, v)
