"""Two-stage line search (rollout.cuh): stage A tries alpha = 1 .. 1/8 for every OCP, stage B the remaining step lengths
only for the OCPs that rejected all of stage A.  A strict acceptance threshold forces small step lengths, so the `pending`
hand-over between decide stage 0 / rollout stage B / decide stage 1 is exercised; the iteration path must still be the
oracle's (which tries the step lengths one after the other, src/sbfddp.cpp:260-290)."""
import importlib

import numpy as np
import pytest

import oracle_binding as ob
import synth

pytestmark = pytest.mark.gpu
capi = importlib.import_module("eagle-mpc_b200.capi")


def rel(a, b):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.mark.parametrize("width_a", [4, 8])
@pytest.mark.parametrize("na,nr,T,th", [(3, 6, 20, 0.999), (0, 4, 25, 0.9999), (2, 6, 15, 0.99999)])
def test_small_steps_take_stage_b(na, nr, T, th, width_a, monkeypatch):
    # stage A rolls out 4 step lengths for batches that fill the GPU and 8 for smaller ones; both splits are exercised here
    monkeypatch.setenv("EMPC_RO_WIDTH_A", str(width_a))
    B = 6
    h = synth.make_problem(seed=40 + na, na=na, n_rotors=nr, T=T, all_costs=False)
    rng = np.random.default_rng(9)
    x0 = np.zeros((B, h.nx)); x0[:, 6] = 1
    x0[:, :3] = rng.uniform(-0.4, 0.4, size=(B, 3))
    x0[:, 7:h.nq] = rng.uniform(-0.3, 0.3, size=(B, h.na))
    pg = capi.default_params(); pg.maxiter = 12; pg.th_acceptstep = th
    po = ob.default_params(); po.maxiter = 12; po.th_acceptstep = th
    g = capi.BatchSolver(h, B)
    g.set_params(pg)
    g.set_x0(x0)
    g.set_candidate(None, None, False)
    g.solve()
    xs, us, cost, iters, feas = g.xs(), g.us(), g.cost(), g.iters(), g.feasible()
    small_steps = 0
    for b in range(B):
        o = ob.Oracle(h)
        o.set_params(po)
        o.set_x0(x0[b])
        o.solve()
        assert int(o.get("iter")) == iters[b], (b, o.get("iter"), iters[b])
        assert int(o.get("feasible")) == feas[b]
        assert rel(cost[b], o.get("cost")) < 1e-8
        assert rel(xs[b], o.get("xs")) < 1e-6
        assert rel(us[b], o.get("us")) < 1e-6
        small_steps += 1
    assert small_steps == B
    # the regularisation must have moved, i.e. steps <= th_stepinc (1/128 and below: stage B) were taken or rejected
    assert (g.reg() > 1e-9).any()
