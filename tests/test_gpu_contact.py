"""GPU parity (through the C ABI) on the contact-dynamics path: DifferentialActionModelContactFwdDynamics
with a ContactModel3D (eagle_catch.yaml, monkey_bar.yaml) or a ContactModel6D (variant of monkey_bar.yaml) and the
friction-cone cost, against the CPU oracle (itself held to the independent complex-step twin in tests/test_twin.py).

Phase level: tiles (incl. Lxu and the full Luu of the friction-cone nodes) / xnext / node cost / gaps, backward K, k, Vx,
rollout trials of several step lengths.  Solver level: same iteration count, cost / xs / us / K / k within the bar.
"""
import importlib
import os

import numpy as np
import pytest

import contact_variants
import oracle_binding as ob
import parity

pytestmark = pytest.mark.gpu
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
YAML_ROOT = os.path.join(ROOT, "yaml")
CATCH = "hexacopter370_flying_arm_3/trajectories/eagle_catch.yaml"
MONKEY = "hexacopter370_flying_arm_3/trajectories/monkey_bar.yaml"


def rel(a, b):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    assert np.isfinite(a).all(), "non-finite values from the device"
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def problem(case, tmp_path, monkeypatch):
    rel_path = case
    if case in ("6d", "hextilt"):
        root, rel_path = contact_variants.six_d(tmp_path) if case == "6d" else contact_variants.hextilt_push(tmp_path)
        monkeypatch.setenv("EAGLE_MPC_YAML_DIR", root)
    return host.Trajectory(rel_path).createProblem(20)


def random_candidate(fp, B, seed):
    """states around the YAML initial state (moving, arm bent), controls inside and outside the squashing box"""
    rng = np.random.default_rng(seed)
    T = fp.T
    xs = np.tile(fp.x0, (B, T + 1, 1))
    xs[:, :, :3] += rng.uniform(-0.3, 0.3, size=(B, T + 1, 3))
    q = rng.normal(size=(B, T + 1, 4)) * 0.15 + np.array([0, 0, 0, 1.0])
    xs[:, :, 3:7] = q / np.linalg.norm(q, axis=-1, keepdims=True)
    xs[:, :, 7:fp.nq] += rng.uniform(-0.6, 0.6, size=(B, T + 1, fp.nq - 7))
    xs[:, :, fp.nq:] = rng.uniform(-0.8, 0.8, size=(B, T + 1, fp.nv))
    us = rng.uniform(-1, 13, size=(B, T, fp.nu))
    us[:, :, fp.nu - (fp.nq - 7):] = rng.uniform(-2.5, 2.5, size=(B, T, fp.nq - 7))
    x0 = xs[:, 0].copy()
    x0[:, :3] += 0.01
    return x0, xs, us


@pytest.mark.parametrize("case", [CATCH, MONKEY, "6d", "hextilt"])
def test_contact_phases(case, tmp_path, monkeypatch):
    fp = problem(case, tmp_path, monkeypatch)
    assert fp.desc.n_contacts == 1
    B = 3
    x0, xs, us = random_candidate(fp, B, 7)
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
    lxu_seen = 0.0
    for b, o in enumerate(oracles):
        o.phase_calc_diff(smooth)
        worst["xnext"] = max(worst.get("xnext", 0), rel(xnext[b, :-1], o.get("xnext")[:-1]))
        worst["cost"] = max(worst.get("cost", 0), rel(ncost[b], o.get("node_cost")))
        worst["fs"] = max(worst.get("fs", 0), rel(gaps[b], o.get("fs")))
        ot = o.get("tiles")
        for name, size in (("Fx", fp.ndx * fp.ndx), ("Fu", fp.ndx * fp.nu), ("Lxx", fp.ndx * fp.ndx), ("Lxu", fp.ndx * fp.nu),
                           ("Luu", fp.nu * fp.nu), ("Lx", fp.ndx), ("Lu", fp.nu)):
            a = tiles[b, :, off[name]:off[name] + size]; c = ot[:, off[name]:off[name] + size]
            worst[name] = max(worst.get(name, 0), rel(a, c))
            if name == "Lxu":
                lxu_seen = max(lxu_seen, np.abs(c).max())
    print(case, {k: f"{v:.1e}" for k, v in worst.items()}, "max |Lxu|", lxu_seen)
    for k, v in worst.items():
        assert v < 1e-9, (k, v)
    if case in (CATCH, "hextilt"):
        assert lxu_seen > 0, "the random candidate should activate a facet of the friction cone"
    # backward pass with and without the gap terms
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
    # rollouts through the contact nodes, several step lengths of both stages
    g.phase_backward(1e-6, False)
    for o in oracles:
        o.phase_backward(1e-6, False)
    for feasible, ddp in ((False, False), (True, False)):
        g.phase_rollout(smooth, feasible, ddp)
        for ai in (0, 2, 5, 9):
            xt, ut, ct, dv, okt = g.trial(ai)
            for b, o in enumerate(oracles):
                ook = o.phase_rollout(smooth, feasible, ddp, ai)
                assert okt[b] == ook
                if not ook:
                    continue
                # A rollout from a random candidate through 70 constrained nodes can run away (joint velocities of 1e4
                # rad/s), and a runaway amplifies rounding differences without bound: the trial is compared up to the
                # node where the oracle's own trial leaves the physically meaningful range, its cost only if it never does.
                xo, uo = o.get("xs_try"), o.get("us_try")
                wild = np.nonzero(np.abs(xo).max(axis=1) > 50.0)[0]
                n_ok = int(wild[0]) if wild.size else fp.T + 1
                assert n_ok >= 10, (ai, n_ok)
                assert rel(xt[b][:n_ok], xo[:n_ok]) < 1e-8, (ai, n_ok, rel(xt[b][:n_ok], xo[:n_ok]))
                assert rel(ut[b][:n_ok - 1], uo[:n_ok - 1]) < 1e-8
                if n_ok == fp.T + 1:
                    assert rel(ct[b], o.get("cost_try")) < 1e-8, (ai, ct[b], o.get("cost_try"))


@pytest.mark.parametrize("case", [CATCH, MONKEY])
def test_contact_trajectory_solve(case, tmp_path, monkeypatch):
    """the reference's two contact trajectories end to end (examples/python/trajectory.py on eagle_catch.yaml /
    monkey_bar.yaml): OCP 0 from the YAML initial state, the others from perturbed ones"""
    fp = problem(case, tmp_path, monkeypatch)
    B = 3
    x0 = wl.noisy_x0(fp.x0, B, 4242)
    x0[0] = fp.x0
    # monkey_bar.yaml starts INSIDE its contact stage (the gripper holds the bar): a perturbed initial state breaks that
    # holonomic constraint at position level and the solve is chaotic from the first iterations (regularisation ramps,
    # step lengths of 2^-9: the oracle built with and without FMA contraction is 1e-11 apart after one iteration, 4e-9
    # after ten, 1e-6 after twenty, and ends at costs that differ in the third digit).  Its perturbed OCPs are therefore
    # held to the oracle over the reproducible prefix of the iteration path and to sanity at the end; OCP 0 — the YAML's
    # own initial state, the reference's use case — is compared over the whole solve like every other problem.
    chaotic = (lambda b: case == MONKEY and b > 0)
    g = capi.BatchSolver(fp, B)
    g.enable_iteration_log(512)
    g.set_x0(x0); g.set_candidate(None, None, False)
    g.solve()
    got = {"xs": g.xs(), "us": g.us(), "K": g.K(), "k": g.k(), "cost": g.cost(), "us_squash": g.us_squash()}
    iters, feas = g.iters(), g.feasible()
    worst = {}
    for b in range(B):
        # yardstick: besides the rounding-level samples, two starts 1e-12 (relative) away — the scale at which two faithful
        # restatements of the contact solve differ (oracle vs twin: 4e-12 on a node)
        for key, d_gpu, d_self in parity.check_ocp((case, b), fp, x0[b], {k_: v[b] for k_, v in got.items()}, iters[b], feas[b],
                                                       log=g.iteration_log(b), perturb=1e-12,
                                                       max_horizon=6 if chaotic(b) else None):
            w = worst.setdefault(key, [0.0, 0.0])
            w[0] = max(w[0], d_gpu); w[1] = max(w[1], d_self)
        if chaotic(b):
            # the end of a chaotic solve: a feasible trajectory of the CONTACT dynamics — the oracle's node model maps
            # (xs[t], us[t]) of the device's solution onto its xs[t + 1], node by node
            assert feas[b] == 1
            o = ob.Oracle(fp)
            smooth = g.iteration_log(b)[-1].smooth
            d = fp.desc
            for t in range(fp.T):
                xnext = o.node_eval(d.node_costset[t], smooth, got["xs"][b][t], got["us"][b][t])[0]
                assert rel(got["xs"][b][t + 1], xnext) < 1e-9, (b, t, rel(got["xs"][b][t + 1], xnext))
    print(case, "iters", iters.tolist(), {k_: f"gpu {v[0]:.1e} / self {v[1]:.1e}" for k_, v in worst.items()})
