"""Solver-level parity helper shared by the -m gpu tests (test infrastructure).

Bar (BASELINE.json north_star): identical iteration count and stopping decision, <= 1e-9 relative on cost, xs, us, K, k.
Where a problem is ill-conditioned the bar is scaled by the oracle's own sensitivity to rounding: the same oracle source
compiled with and without FMA contraction (liboracle.so / liboracle_nofma.so) gives the yardstick d_self = |o - o_nofma|,
and the GPU must stay within max(1e-9, 4 d_self) of the oracle (16 d_self where d_self > 1e-6: there a single sample of
the sensitivity is only an order of magnitude).  The tolerance is therefore always bounded by a measured quantity: there
is no path on which a key goes unchecked.  If the two oracle builds do not even agree on the iteration count (a stop
test decided by the last bits), the GPU must reproduce the iteration count and feasibility of ONE of them and is compared
against that build, the distance between the two builds being the yardstick.
"""
import numpy as np

import oracle_binding as ob

TOL = 1e-9
KEYS = ("cost", "xs", "us", "K", "k", "us_squash")


def rel(a, b):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def oracle_pair(fp, x0, params=None, xs=None, us=None):
    out = []
    for nofma in (False, True):
        o = ob.Oracle(fp, nofma=nofma)
        if params is not None:
            o.set_params(params)
        o.set_x0(x0)
        o.solve(xs, us)
        out.append(o)
    return out


def check_ocp(tag, fp, x0, got, iters, feas, params=None, keys=KEYS, xs=None, us=None):
    """got: dict key -> this OCP's array from the GPU.  Returns [(key, d_gpu, d_self)]."""
    o, o2 = oracle_pair(fp, x0, params, xs, us)
    it = (int(o.get("iter")), int(o2.get("iter")))
    ref = o
    if it[0] != it[1]:
        assert int(iters) in it, (tag, "iterations", int(iters), it)
        ref = o if int(iters) == it[0] else o2
    assert int(ref.get("iter")) == int(iters), (tag, "iterations", int(iters), it)
    assert int(ref.get("feasible")) == int(feas), (tag, "feasible")
    report = []
    for key in keys:
        d_self = rel(o2.get(key), o.get(key))
        d_gpu = rel(got[key], ref.get(key))
        factor = 16 if d_self > 1e-6 else 4
        assert d_gpu <= max(TOL, factor * d_self), (tag, key, d_gpu, d_self)
        report.append((key, d_gpu, d_self))
    return report
