"""BASELINE.json configs 2 and 4 at their full sizes on one GPU: hexacopter370_flying_arm_3 displacement (T = 400,
B = 4096 OCPs) and hextilt_flying_arm_5 push_slide (T = 100, B = 16384 OCPs).

The oracle cannot solve 4096 OCPs in test time, so the full batch is checked through properties that do not depend on
the batch size, and a sample of its OCPs against the oracle:
  * every OCP terminates feasible with closed shooting gaps (xs[t+1] = f(xs[t], us[t]) re-evaluated by the calc kernels),
    unit quaternions, squashed controls inside the actuator bounds;
  * the work counter (sum of inner iterations) equals the per-OCP iteration counts;
  * batch independence: OCPs taken out of the big batch and solved in a batch of their own give bit-identical results
    (an OCP's arithmetic does not depend on its neighbours or its position);
  * the sampled OCPs match the oracle: identical iteration count, cost / xs / us / K / k / us_squash within the bar of
    tests/parity.py (1e-9, scaled only by the oracle's own measured rounding sensitivity).
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


def rel(a, b):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.mark.parametrize("name,B,T,dims", [("hexacopter370_flying_arm_3_displacement", 4096, 400, (19, 18, 9)),
                                           ("hextilt_flying_arm_5_push_slide", 16384, 100, (23, 22, 11))])
def test_full_batch_properties_and_sampled_parity(name, B, T, dims):
    yaml, dt, seed0 = wl.CONFIGS[name]
    tr = host.Trajectory(yaml)
    fp = tr.createProblem(dt)
    x0 = wl.noisy_x0(fp.x0, B, seed0)
    g = capi.BatchSolver(fp, B)
    assert g.T == T and (g.nx, g.ndx, g.nu) == dims
    g.set_x0(x0); g.set_candidate(None, None, False)
    g.solve()
    xs, us, uss, cost, iters, feas = g.xs(), g.us(), g.us_squash(), g.cost(), g.iters(), g.feasible()
    assert np.isfinite(cost).all() and np.isfinite(xs).all() and np.isfinite(us).all()
    for a, b_ in zip(g.solution(), (xs, us, uss, cost, g.stop(), iters, feas)):   # large-batch path of empc_get_solution
        assert np.array_equal(a, b_)
    assert (feas == 1).all()
    assert (iters >= 1).all() and (iters < 100).all()           # nobody ran into maxiter
    assert g.total_iterations() == int((iters + 1).sum())       # iter_ = total_iters_ - 1 (src/sbfddp.cpp:222)
    assert np.abs(np.linalg.norm(xs[:, :, 3:7], axis=2) - 1.0).max() < 1e-12
    assert np.array_equal(xs[:, 0], x0)
    lb = np.array(list(fp.desc.u_lb)[:g.nu]); ub = np.array(list(fp.desc.u_ub)[:g.nu])
    assert (uss >= lb - 1e-12).all() and (uss <= ub + 1e-12).all()
    # shooting gaps of the solution, re-evaluated: fs[t+1] = diff(xs[t+1], f(xs[t], us[t]))
    g.phase_calc_diff(0.05)
    assert np.abs(g.gaps()).max() < 1e-9
    # batch independence + sampled parity
    sample = np.array([0, 1, 777, B // 2, B - 763, B - 1])
    g.close()
    s = capi.BatchSolver(fp, len(sample))
    s.set_x0(x0[sample]); s.set_candidate(None, None, False); s.solve()
    assert np.array_equal(s.iters(), iters[sample])
    assert np.array_equal(s.cost(), cost[sample])
    assert np.array_equal(s.xs(), xs[sample]) and np.array_equal(s.us(), us[sample])
    K, k, su = s.K(), s.k(), s.us_squash()
    for j, b in enumerate(sample):
        got = {"cost": cost[b], "xs": xs[b], "us": us[b], "K": K[j], "k": k[j], "us_squash": su[j]}
        parity.check_ocp((name, int(b)), fp, x0[b], got, iters[b], feas[b])
    print(name, "full batch: iterations min/median/max", iters.min(), int(np.median(iters)), iters.max(),
          "total", int((iters + 1).sum()))
