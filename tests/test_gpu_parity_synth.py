"""GPU parity (through the C ABI) against the CPU oracle on synthetic problems with every cost type.

Phase-level: calc_diff tiles / xnext / cost / gaps, backward K,k,Vx,Vxx.fs,dg,dq, rollout trials for all step lengths.
Solver-level: identical iteration count, stopping state, cost/xs/us/K/k within 1e-9 relative.
"""
import importlib

import numpy as np
import pytest

import oracle_binding as ob
import synth

pytestmark = pytest.mark.gpu
empc = importlib.import_module("eagle-mpc_b200")


def rel(a, b):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def make(na, nr, T=10, B=5, seed=0, all_costs=True):
    from importlib import import_module
    capi = import_module("eagle-mpc_b200.capi")
    h = synth.make_problem(seed=seed, na=na, n_rotors=nr, T=T, all_costs=all_costs)
    rng = np.random.default_rng(seed + 100)
    x0 = np.stack([synth.random_state(rng, h, scale=0.3) for _ in range(B)])
    xs = np.stack([[synth.random_state(rng, h, scale=0.3) for _ in range(T + 1)] for _ in range(B)])
    us = rng.uniform(-1, 13, size=(B, T, h.nu))
    us[:, :, nr:] = rng.uniform(-2.5, 2.5, size=(B, T, h.na))
    g = capi.BatchSolver(h, B)
    g.set_x0(x0)
    g.set_candidate(xs, us, False)
    oracles = []
    for b in range(B):
        o = ob.Oracle(h)
        o.set_x0(x0[b]); o.set_candidate(xs[b], us[b], False)
        oracles.append(o)
    return h, g, oracles, x0, xs, us


@pytest.mark.parametrize("na,nr", [(0, 4), (0, 6), (2, 6), (3, 6), (5, 6)])
def test_phases(na, nr):
    h, g, oracles, x0, xs, us = make(na, nr, seed=na)
    B = g.B
    smooth = 0.1
    g.phase_calc_diff(smooth)
    tiles, xnext, ncost, gaps = g.tiles(), g.xnext(), g.node_cost(), g.gaps()
    for b, o in enumerate(oracles):
        o.phase_calc_diff(smooth)
        assert rel(xnext[b, :-1], o.get("xnext")[:-1]) < 1e-12
        assert rel(ncost[b], o.get("node_cost")) < 1e-12
        assert rel(gaps[b], o.get("fs")) < 1e-11
        ot = o.get("tiles")
        off = h.tile_offsets()
        for name, size in (("Fx", h.ndx * h.ndx), ("Fu", h.ndx * h.nu), ("Lxx", h.ndx * h.ndx), ("Lxu", h.ndx * h.nu),
                           ("Luu", h.nu * h.nu), ("Lx", h.ndx), ("Lu", h.nu)):
            a = tiles[b, :, off[name]:off[name] + size]; c = ot[:, off[name]:off[name] + size]
            assert rel(a, c) < 1e-10, (name, rel(a, c))
    # backward pass (infeasible: gap terms on) with a comfortable regularisation
    for feasible in (False, True):
        ok = g.phase_backward(1e-6, feasible)
        K, k, Vx, gv, dgdq = g.K(), g.k(), g.Vx(), g.Vxx_fs(), g.dgdq()
        for b, o in enumerate(oracles):
            ook = o.phase_backward(1e-6, feasible)
            assert ok[b] == ook
            if not ook:
                continue
            assert rel(K[b], o.get("K")) < 1e-9
            assert rel(k[b], o.get("k")) < 1e-9
            assert rel(Vx[b], o.get("Vx")) < 1e-9
            Vxx = o.get("Vxx"); fs = o.get("fs")
            assert rel(gv[b], np.einsum("tij,tj->ti", Vxx, fs)) < 1e-9
            if not feasible:
                assert rel(dgdq[b], o.get("dgdq")) < 1e-9
    # rollouts, FDDP infeasible / feasible / DDP, every step length
    ok = g.phase_backward(1e-6, False)
    for o in oracles:
        o.phase_backward(1e-6, False)
    for feasible, ddp in ((False, False), (True, False), (False, True)):
        g.phase_rollout(smooth, feasible, ddp)
        for ai in (0, 1, 4, 9):
            xt, ut, ct, dv, okt = g.trial(ai)
            for b, o in enumerate(oracles):
                if ddp:
                    ob.lib.orc_set_xs_try0(o.p, ob.dp(np.ascontiguousarray(x0[b])))
                ook = o.phase_rollout(smooth, feasible, ddp, ai)
                assert okt[b] == ook
                if not ook:
                    continue
                assert rel(xt[b], o.get("xs_try")) < 1e-9
                assert rel(ut[b], o.get("us_try")) < 1e-9
                assert rel(ct[b], o.get("cost_try")) < 1e-9
                if not ddp and not feasible:
                    assert abs(dv[b] - o.get("dv")) < 1e-9 * max(1.0, abs(o.get("dv")))


@pytest.mark.parametrize("na,nr,T", [(0, 4, 30), (3, 6, 40), (5, 6, 25)])
def test_full_solve(na, nr, T):
    capi = importlib.import_module("eagle-mpc_b200.capi")
    B = 6
    h = synth.make_problem(seed=20 + na, na=na, n_rotors=nr, T=T, all_costs=False)
    rng = np.random.default_rng(5)
    x0 = np.zeros((B, h.nx)); x0[:, 6] = 1
    x0[:, :3] = rng.uniform(-0.3, 0.3, size=(B, 3))
    x0[:, 7:h.nq] = rng.uniform(-0.2, 0.2, size=(B, h.na))
    g = capi.BatchSolver(h, B)
    g.enable_iteration_log(1024)
    g.set_x0(x0)
    g.set_candidate(None, None, False)
    g.solve()
    xs, us, K, k, cost, iters, feas, stop, uss = g.xs(), g.us(), g.K(), g.k(), g.cost(), g.iters(), g.feasible(), g.stop(), g.us_squash()
    assert g.total_iterations() == int((iters + 1).sum())
    # random synthetic problems crawl for ~200 iterations and are chaotic in the DDP clean-up (two builds of the oracle drift
    # apart to 1e-4): every OCP's iteration path is checked strictly up to the reference's own reproducibility horizon
    # (tests/parity.py); no OCP is skipped.
    import parity
    got = {"xs": xs, "us": us, "K": K, "k": k, "cost": cost, "us_squash": uss, "stop": stop}
    for b in range(B):
        # (two more yardstick samples a few ulp away: the four standard ones spread over two orders of magnitude on these
        #  chaotic solves — 1.5e-11 .. 5.4e-10 on K of one OCP, 1.6e-9 with the extra pair — so a larger sample is a fairer bar)
        parity.check_ocp(("synth", na, b), h, x0[b], {k_: v[b] for k_, v in got.items()}, iters[b], feas[b],
                         keys=parity.KEYS + ("stop",), log=g.iteration_log(b), perturb=1e-15)


def test_diverging_rollout_is_a_forward_error():
    """crocoddyl::raiseIfNaN trips on NaN, +-inf AND any value >= 1e30: a rollout whose states leave that range without ever
    producing a NaN must be skipped like the reference's forward_error (src/sbfddp.cpp:265-267), and a finite but huge cost_try
    likewise.  x0 carries a finite base velocity of 2e30: every step length fails at the first node, on both sides."""
    capi = importlib.import_module("eagle-mpc_b200.capi")
    h = synth.make_problem(seed=31, na=3, n_rotors=6, T=8, all_costs=False)
    B = 3
    x0 = np.zeros((B, h.nx)); x0[:, 6] = 1
    x0[1, h.nq] = 2e30          # OCP 1: finite, but beyond the raiseIfNaN range
    x0[2, h.nq + 1] = -3e30     # OCP 2: the same with a negative entry (the test is on the infinity norm)
    xs = np.repeat(x0[:, None, :], h.T + 1, axis=1)
    us = np.zeros((B, h.T, h.nu))
    g = capi.BatchSolver(h, B)
    g.set_x0(x0); g.set_candidate(xs, us, False)
    g.phase_calc_diff(0.1)
    g.phase_rollout(0.1, True, False)
    for ai in (0, 3, 9):
        _xt, _ut, _c, _dv, ok = g.trial(ai)
        assert ok.tolist() == [1, 0, 0], (ai, ok)
    for b in range(B):
        o = ob.Oracle(h); o.set_x0(x0[b]); o.set_candidate(xs[b], us[b], False)
        o.phase_calc_diff(0.1)
        assert bool(o.phase_rollout(0.1, True, False, 0)) == (b == 0)


def test_device_pointers_at_the_c_abi():
    """empc_set_x0 / empc_set_candidate / empc_get_* take host or device memory (the NCCL data plane hands device buffers
    straight to the handle): identical results either way."""
    torch = pytest.importorskip("torch")
    capi = importlib.import_module("eagle-mpc_b200.capi")
    abi = importlib.import_module("eagle-mpc_b200.abi")
    import ctypes as C
    h = synth.make_problem(seed=23, na=3, n_rotors=6, T=20, all_costs=False)
    B = 4
    rng = np.random.default_rng(9)
    x0 = np.zeros((B, h.nx)); x0[:, 6] = 1; x0[:, :3] = rng.uniform(-0.2, 0.2, size=(B, 3))
    p = capi.default_params(); p.maxiter = 6
    g = capi.BatchSolver(h, B); g.set_params(p)
    g.set_x0(x0); g.set_candidate(None, None, False); g.solve()
    xs_host, us_host = g.xs(), g.us()
    g2 = capi.BatchSolver(h, B); g2.set_params(p)
    x0_d = torch.from_numpy(x0).cuda()
    g2.set_x0_ptr(x0_d.data_ptr())
    g2.set_candidate(None, None, False)
    g2.solve()
    xs_d = torch.empty((B, h.T + 1, h.nx), dtype=torch.float64, device="cuda")
    us_d = torch.empty((B, h.T, h.nu), dtype=torch.float64, device="cuda")
    g2.get_into("xs", xs_d.data_ptr()); g2.get_into("us", us_d.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(xs_d.cpu().numpy(), xs_host) and np.array_equal(us_d.cpu().numpy(), us_host)
    # a device-resident candidate
    lib = capi.lib()
    xs_c, us_c = torch.from_numpy(xs_host).cuda(), torch.from_numpy(us_host).cuda()
    rc = lib.empc_set_candidate(g2.h, C.cast(xs_c.data_ptr(), abi.c_double_p), C.cast(us_c.data_ptr(), abi.c_double_p), 0)
    assert rc == 0
    g.set_candidate(xs_host, us_host, False)
    g.solve(); g2.solve()
    assert np.array_equal(g.xs(), g2.xs()) and np.array_equal(g.iters(), g2.iters())
