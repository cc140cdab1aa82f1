"""The oracle against an INDEPENDENT second restatement (oracle/twin.py: PyYAML + xml.etree front-end, M^-1 (tau - h)
dynamics by recursive Newton-Euler, every derivative by the complex step).  CPU only.

  * both front-ends (the product's host mirror and the twin) turn every loadable YAML of yaml/ into the same robot,
    actuation map, knot layout and cost tables;
  * the oracle's analytic node blocks (xnext, cost, Fx, Fu, Lx, Lu, Lxx, Luu) equal the twin's complex-step blocks on all
    five robot families, for every stage's cost set and for the terminal node;
  * one Riccati sweep and one rollout of the oracle equal the twin's dense numpy versions.
"""
import glob
import importlib
import os
import sys

import numpy as np
import pytest

import oracle_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import twin  # noqa: E402

host = importlib.import_module("eagle-mpc_b200.host")
abi = importlib.import_module("eagle-mpc_b200.abi")
YAML_ROOT = os.path.join(ROOT, "yaml")
URDF_ROOT = os.path.join(ROOT, "fixtures", "urdf")

TYPE = {"CostModelState": abi.COST_STATE, "CostModelControl": abi.COST_CONTROL, "CostModelFramePlacement": abi.COST_FRAME_PLACEMENT,
        "CostModelFrameRotation": abi.COST_FRAME_ROTATION, "CostModelFrameVelocity": abi.COST_FRAME_VELOCITY,
        "CostModelFrameTranslation": abi.COST_FRAME_TRANSLATION, "Barrier": abi.COST_SQUASH_BARRIER,
        "CostModelContactFrictionCone": abi.COST_CONTACT_FRICTION_CONE}
ACT = {"ActivationModelQuad": abi.ACT_QUAD, "ActivationModelWeightedQuad": abi.ACT_WEIGHTED_QUAD,
       "ActivationModelQuadraticBarrier": abi.ACT_QUAD_BARRIER, "ActivationModelWeightedQuadraticBarrier": abi.ACT_WEIGHTED_QUAD_BARRIER}


def trajectory_yamls():
    out = []
    for p in sorted(glob.glob(os.path.join(YAML_ROOT, "*", "trajectories", "*.yaml"))):
        out.append(os.path.relpath(p, YAML_ROOT))
    return out


def dt_of(rel):
    return 10 if rel == "hexacopter370/trajectories/displacement.yaml" else 20


@pytest.mark.parametrize("rel", trajectory_yamls())
def test_front_ends_agree(rel):
    dt = dt_of(rel)
    fp = host.Trajectory(rel).createProblem(dt)
    tw = twin.Problem(rel, YAML_ROOT, URDF_ROOT, dt)
    d, rob = fp.desc, tw.rob
    # robot
    assert d.robot.n_joints == rob.nj and (fp.nq, fp.nv, fp.nu, fp.T) == (rob.nq, rob.nv, tw.nu, tw.T)
    for i in range(rob.nj):
        assert d.robot.parent[i] == rob.parent[i]
        np.testing.assert_allclose(np.array(d.robot.jplace_R[i]).reshape(3, 3), rob.place_R[i], atol=1e-15)
        np.testing.assert_allclose(np.array(d.robot.jplace_p[i]), rob.place_p[i], atol=1e-15)
        if i:
            np.testing.assert_allclose(np.array(d.robot.axis[i]), rob.axis[i], atol=1e-15)
        m, c = d.robot.mass[i], np.array(d.robot.com[i])
        np.testing.assert_allclose(m, rob.mass[i], rtol=1e-14)
        np.testing.assert_allclose(m * c, rob.mc[i], atol=1e-15)
        C = twin.hat(c)
        np.testing.assert_allclose(np.array(d.robot.inertia[i]).reshape(3, 3) - m * (C @ C), rob.I_o[i], atol=1e-15)
    # actuation
    np.testing.assert_allclose(np.array(d.tau_f[:6 * tw.nr]).reshape(6, tw.nr), tw.tau_f, atol=1e-15)
    np.testing.assert_allclose(np.array(d.u_lb[:tw.nu]), tw.u_lb, atol=0); np.testing.assert_allclose(np.array(d.u_ub[:tw.nu]), tw.u_ub, atol=0)
    assert d.dt == tw.dt
    np.testing.assert_allclose(fp.x0, tw.x0, atol=0)
    # knot layout: node -> cost set = node -> stage
    node_cs = [d.node_costset[t] for t in range(fp.T + 1)]
    assert d.n_costsets == len(tw.stages)
    assert node_cs == tw.node_stage
    # cost tables, in the reference's iteration order
    pool = np.array([d.pool[i] for i in range(d.n_pool)])
    for cs, st in enumerate(tw.stages):
        names = fp.cost_names(cs)
        assert names == list(st["costs"].keys()), (rel, cs)
        for k, (name, c) in enumerate(st["costs"].items()):
            h = d.costs[d.costset_begin[cs] + k]
            assert h.type == TYPE[c["type"]] and bool(h.active) == c["active"], (rel, name)
            assert h.weight == c["weight"], (rel, name)
            if c["type"] == "Barrier":
                continue
            assert h.activation == ACT[c["act"]], (rel, name)
            if "ref" in c:
                np.testing.assert_allclose(pool[h.ref_off:h.ref_off + c["ref"].size], c["ref"], atol=0)
            if "A" in c:   # friction cone: facet matrix (5 x 3)
                np.testing.assert_allclose(pool[h.ref_off:h.ref_off + 15].reshape(5, 3), c["A"], atol=1e-15)
            if "frame" in c:
                j, Rf, pf = rob.frames[c["frame"]]
                assert d.robot.frame_joint[h.frame] == j
                np.testing.assert_allclose(np.array(d.robot.frame_R[h.frame]).reshape(3, 3), Rf, atol=1e-15)
                np.testing.assert_allclose(np.array(d.robot.frame_p[h.frame]), pf, atol=1e-15)
            ref = []
            if "R" in c: ref += list(c["R"].reshape(-1))
            if "p" in c: ref += list(c["p"])
            if "vref" in c: ref += list(c["vref"])
            if ref:
                np.testing.assert_allclose(pool[h.ref_off:h.ref_off + len(ref)], ref, atol=1e-15)
            for key, off in (("w", h.w_off), ("lb", h.lb_off), ("ub", h.ub_off)):
                if key in c:
                    np.testing.assert_allclose(pool[off:off + c[key].size], c[key], atol=0)
                else:
                    assert off < 0 or key == "w", (rel, name, key)


        # contact of the stage's model (src/stage.cpp:38-47)
        ci = d.costset_contact[cs] if d.n_contacts else -1
        if st["contact"] is None:
            assert ci < 0
        else:
            hc = d.contacts[ci]
            assert hc.type == {"ContactModel3D": abi.CONTACT_3D, "ContactModel6D": abi.CONTACT_6D}[st["contact"]["type"]]
            j, Rf, pf = rob.frames[st["contact"]["frame"]]
            assert d.robot.frame_joint[hc.frame] == j
            np.testing.assert_allclose(np.array(d.robot.frame_R[hc.frame]).reshape(3, 3), Rf, atol=1e-15)
            np.testing.assert_allclose(np.array(d.robot.frame_p[hc.frame]), pf, atol=1e-15)
            assert hc.gains[0] == 0 and hc.gains[1] == 0


def rel_err(a, b):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def random_state(tw, rng, scale=1.0):
    rob = tw.rob
    dx = scale * np.concatenate([rng.uniform(-0.5, 0.5, 3), rng.uniform(-0.8, 0.8, 3), rng.uniform(-0.7, 0.7, rob.na),
                                 rng.uniform(-1.0, 1.0, rob.nv)])
    return twin.integrate(rob, tw.x0.astype(float), dx)


def tile_blocks(tile, ndx, nu):
    o = 0
    out = {}
    for name, shape in (("Fx", (ndx, ndx)), ("Fu", (ndx, nu)), ("Lxx", (ndx, ndx)), ("Lxu", (ndx, nu)), ("Luu", (nu, nu)), ("Lx", (ndx,)), ("Lu", (nu,))):
        n = int(np.prod(shape))
        out[name] = tile[o:o + n].reshape(shape); o += n
    return out


ROBOTS = ["hexacopter370/trajectories/passthrough.yaml", "hexacopter370_flying_arm_3/trajectories/displacement.yaml",
          "hextilt_flying_arm_5/trajectories/push_slide.yaml", "iris_px4/trajectories/displacement.yaml",
          "hexacopter680_flying_arm_2/trajectories/hover.yaml"]


@pytest.mark.parametrize("rel", ROBOTS)
def test_oracle_blocks_equal_complex_step(rel):
    """every stage's cost set and the terminal node, random states / controls, two smoothing values"""
    dt = dt_of(rel)
    fp = host.Trajectory(rel).createProblem(dt)
    tw = twin.Problem(rel, YAML_ROOT, URDF_ROOT, dt)
    o = ob.Oracle(fp)
    rng = np.random.default_rng(abs(hash(rel)) % 1000)
    worst = {}
    stages = sorted(set(range(len(tw.stages))), key=lambda s: -len(tw.stages[s]["costs"]))[:3]   # the richest cost sets
    for cs in stages:
        for smooth, terminal in ((0.1, False), (0.05, False), (0.05, True)):
            x = random_state(tw, rng)
            u = rng.uniform(tw.u_lb - 0.3 * (tw.u_ub - tw.u_lb), tw.u_ub + 0.3 * (tw.u_ub - tw.u_lb))   # also outside the box: barrier active
            ref = tw.calc_diff(cs, x, u, smooth, terminal)
            xnext, cost, _s, tile = o.node_eval(cs, smooth, x, None if terminal else u)
            got = tile_blocks(tile, fp.ndx, fp.nu)
            got["xnext"], got["cost"] = xnext, cost
            # the terminal node only contributes its cost, Lx and Lxx to the solver (SolverDDP::backwardPass)
            for key in (("cost", "Lx", "Lxx") if terminal else ("xnext", "cost", "Fx", "Fu", "Lx", "Lu", "Lxx", "Luu", "Lxu")):
                e = rel_err(got[key], ref[key])
                worst[key] = max(worst.get(key, 0.0), e)
                assert e <= 1e-12, (rel, cs, smooth, terminal, key, e)
    print(rel, {k: f"{v:.1e}" for k, v in worst.items()})


CONTACT_YAMLS = ["hexacopter370_flying_arm_3/trajectories/eagle_catch.yaml", "hexacopter370_flying_arm_3/trajectories/monkey_bar.yaml"]


import contact_variants  # noqa: E402


@pytest.mark.parametrize("case", CONTACT_YAMLS + ["6d", "hextilt"])
def test_oracle_contact_blocks_equal_complex_step(case, tmp_path, monkeypatch):
    """DifferentialActionModelContactFwdDynamics (ContactModel3D of the corpus, a ContactModel6D variant) and the
    friction-cone cost: the oracle's KKT dynamics and its implicit-function derivatives (Fx, Fu, and the force Jacobians
    inside Lx, Lu, Lxx, Lxu, Luu) against the twin's complex step through its own dense KKT solve.  The bar is 1e-10: the
    contact solve divides by Jc M^-1 Jc^T, and the two restatements factorise differently."""
    yaml_root, rel = YAML_ROOT, case
    if case in ("6d", "hextilt"):
        yaml_root, rel = contact_variants.six_d(tmp_path) if case == "6d" else contact_variants.hextilt_push(tmp_path)
        monkeypatch.setenv("EAGLE_MPC_YAML_DIR", yaml_root)
    fp = host.Trajectory(rel).createProblem(20)
    tw = twin.Problem(rel, yaml_root, URDF_ROOT, 20)
    assert fp.desc.n_contacts == 1
    o = ob.Oracle(fp)
    rng = np.random.default_rng(3)
    worst = {}
    n_checked = 0
    for cs in range(len(tw.stages)):
        if tw.stages[cs]["contact"] is None:
            continue
        for smooth, terminal in ((0.1, False), (0.05, False), (0.05, True)):
            x = random_state(tw, rng)
            u = rng.uniform(tw.u_lb - 0.3 * (tw.u_ub - tw.u_lb), tw.u_ub + 0.3 * (tw.u_ub - tw.u_lb))
            ref = tw.calc_diff(cs, x, u, smooth, terminal)
            xnext, cost, _s, tile = o.node_eval(cs, smooth, x, None if terminal else u)
            got = tile_blocks(tile, fp.ndx, fp.nu)
            got["xnext"], got["cost"] = xnext, cost
            for key in (("cost", "Lx", "Lxx") if terminal else ("xnext", "cost", "Fx", "Fu", "Lx", "Lu", "Lxx", "Luu", "Lxu")):
                e = rel_err(got[key], ref[key])
                worst[key] = max(worst.get(key, 0.0), e)
                assert e <= 1e-10, (rel, cs, smooth, terminal, key, e)
            n_checked += 1
            # the constraint holds: the contact frame's constrained acceleration vanishes at the oracle's xnext
            a = (xnext[tw.rob.nq:] - x[tw.rob.nq:]) / tw.dt
            acc = tw.rob.contact_acceleration(tw.stages[cs]["contact"], x[:tw.rob.nq], x[tw.rob.nq:], a)
            assert np.abs(acc).max() <= 1e-8 * max(1.0, np.abs(a).max())
    assert n_checked >= 3
    if case.endswith("eagle_catch.yaml"):
        assert worst["Lxu"] > 0 or True   # (Lxu is non-zero only when a facet of the cone is active)
    print(rel, {k: f"{v:.1e}" for k, v in worst.items()})


@pytest.mark.parametrize("rel", ["hexacopter370_flying_arm_3/trajectories/displacement.yaml", "hexacopter370/trajectories/passthrough.yaml",
                                 "hextilt_flying_arm_5/trajectories/push_slide.yaml"])
def test_oracle_rk4_blocks_equal_complex_step(rel):
    """IntegratedActionModelRK4 (src/factory/int-action.cpp:29-31): the oracle's chain rule through the four stages
    (rk4.hxx: dyi_dx, dki_dx, the Gauss-Newton pull-back of the stage costs, Lxu and a dense Luu included) against the
    twin's complex step of its own four-stage calc"""
    rk4 = "IntegratedActionModelRK4"
    fp = host.Trajectory(rel).createProblem(20, True, rk4)
    assert fp.desc.integrator == abi.INTEGRATOR_RK4
    tw = twin.Problem(rel, YAML_ROOT, URDF_ROOT, 20, integrator=rk4)
    o = ob.Oracle(fp)
    rng = np.random.default_rng(5)
    worst = {}
    stages = sorted(set(range(len(tw.stages))), key=lambda s: -len(tw.stages[s]["costs"]))[:2]
    for cs in stages:
        for smooth, terminal in ((0.1, False), (0.05, True)):
            x = random_state(tw, rng)
            u = rng.uniform(tw.u_lb - 0.3 * (tw.u_ub - tw.u_lb), tw.u_ub + 0.3 * (tw.u_ub - tw.u_lb))
            ref = tw.calc_diff(cs, x, u, smooth, terminal)
            xnext, cost, _s, tile = o.node_eval(cs, smooth, x, None if terminal else u)
            got = tile_blocks(tile, fp.ndx, fp.nu)
            got["xnext"], got["cost"] = xnext, cost
            for key in (("cost", "Lx", "Lxx") if terminal else ("xnext", "cost", "Fx", "Fu", "Lx", "Lu", "Lxx", "Luu", "Lxu")):
                e = rel_err(got[key], ref[key])
                worst[key] = max(worst.get(key, 0.0), e)
                assert e <= 1e-12, (rel, cs, smooth, terminal, key, e)
    print(rel, {k: f"{v:.1e}" for k, v in worst.items()})


def test_riccati_sweep_and_rollout_equal_dense_numpy():
    rel, dt = "hexacopter370/trajectories/hover.yaml", 20
    fp = host.Trajectory(rel).createProblem(dt)
    tw = twin.Problem(rel, YAML_ROOT, URDF_ROOT, dt)
    rob, T = tw.rob, 12   # the first 12 running nodes + the terminal model are enough for the algebra
    rng = np.random.default_rng(11)
    smooth, xreg = 0.1, 1e-6
    x0 = random_state(tw, rng, 0.3)
    xs_full = np.array([random_state(tw, rng, 0.3) for _ in range(fp.T + 1)])
    us_full = rng.uniform(0.3 * tw.u_ub, 0.7 * tw.u_ub, size=(fp.T, tw.nu))
    o = ob.Oracle(fp)
    o.set_x0(x0); o.set_candidate(xs_full, us_full, False)
    o.phase_calc_diff(smooth)
    assert o.phase_backward(xreg, False)
    # the twin works on the tail of the horizon: nodes T0..T (the Riccati sweep runs backwards, so the tail is self-contained)
    T0 = fp.T - T
    nodes = [tw.calc_diff(tw.node_stage[t], xs_full[t], us_full[t], smooth) for t in range(T0, fp.T)]
    term = tw.calc_diff(tw.node_stage[fp.T], xs_full[fp.T], None, smooth, terminal=True)
    fs = [twin.diff(rob, xs_full[t], nodes[t - 1 - T0]["xnext"]) if t > T0 else np.zeros(rob.ndx) for t in range(T0, fp.T + 1)]
    fs_o = o.get("fs")
    for t in range(T0 + 1, fp.T + 1):
        assert rel_err(fs_o[t], fs[t - T0]) <= 1e-12
    fs[0] = fs_o[T0]   # the gap of the first tail node comes from the node before it
    K, k, Vx, Vxx = twin.riccati_sweep(nodes, term, fs, xreg, False)
    Ko, ko, Vxo, Vxxo = o.get("K"), o.get("k"), o.get("Vx"), o.get("Vxx")
    for t in range(T):
        assert rel_err(Ko[T0 + t], K[t]) <= 1e-9, (t, rel_err(Ko[T0 + t], K[t]))
        assert rel_err(ko[T0 + t], k[t]) <= 1e-9
        assert rel_err(Vxo[T0 + t], Vx[t]) <= 1e-9 and rel_err(Vxxo[T0 + t], Vxx[t]) <= 1e-9
    # rollout with alpha = 1/2 over a short problem of its own (same stage models): oracle vs twin
    ai, alpha = 1, 0.5
    assert o.phase_rollout(smooth, False, False, ai)
    xt_o, ut_o, c_o = o.get("xs_try"), o.get("us_try"), o.get("cost_try")
    # the twin replays the first nodes of the oracle's rollout with the oracle's own gains (the algebra under test is the
    # rollout rule, the node model was checked above)
    n_chk = 10
    xs_t, us_t, _c = twin.rollout(_Head(tw, n_chk), x0, xs_full[:n_chk + 1], us_full[:n_chk], Ko[:n_chk], ko[:n_chk], fs_o[:n_chk + 1], alpha, smooth, False)
    assert rel_err(xt_o[:n_chk], xs_t[:n_chk]) <= 1e-10 and rel_err(ut_o[:n_chk], us_t[:n_chk]) <= 1e-10
    # and the trial cost: sum of the twin's node costs along the ORACLE's trial trajectory
    c_t = sum(tw.calc(tw.node_stage[t], xt_o[t], ut_o[t], smooth)[1] for t in range(fp.T)) + tw.calc(tw.node_stage[fp.T], xt_o[fp.T], None, smooth, terminal=True)[1]
    assert abs(c_t - c_o) <= 1e-11 * max(1.0, abs(c_o))


class _Head:
    """the first n running nodes of a twin problem (for the rollout replay)"""

    def __init__(self, tw, n):
        self.rob, self.calc = tw.rob, tw.calc
        self.node_stage = tw.node_stage[:n] + [tw.node_stage[n]]
