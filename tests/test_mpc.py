"""CarrotMpc: host-side retargeting (CPU) and closed-loop parity GPU vs oracle (GPU)."""
import ctypes as C
import importlib

import numpy as np
import pytest

import oracle_binding as ob

host = importlib.import_module("eagle-mpc_b200.host")
mpcmod = importlib.import_module("eagle-mpc_b200.mpc")
abi = importlib.import_module("eagle-mpc_b200.abi")

TRAJ = "hexacopter370_flying_arm_3/trajectories/displacement.yaml"
MPC = "hexacopter370_flying_arm_3/mpc/mpc.yaml"


def _trajectory_solution():
    tr = host.Trajectory(TRAJ)
    fp = tr.createProblem(20)
    p = ob.default_params(); p.maxiter = 400   # examples/python/mpc.py:29
    o = ob.Oracle(fp); o.set_params(p); o.set_x0(fp.x0); o.solve()
    return tr, fp, o.get("xs"), o.get("us")


def test_carrot_retargeting_matches_reference_rules():
    tr, fp, xs, us = _trajectory_solution()
    mpc = mpcmod.CarrotMpc(tr, xs, 20, MPC, create_solver=False)
    assert (mpc.knots, mpc.dt, mpc.iters) == (30, 30, 2) and mpc.T == 29
    assert mpc.desc.n_costsets == 30   # one model per knot (src/mpc-controllers/carrot-mpc.cpp:195-225)
    names = {c.type for c in mpc.cost_tables()[0]}
    assert abi.COST_SQUASH_BARRIER in names

    def carrot_state(knot):
        costs, pool = mpc.cost_tables()
        begin = np.ctypeslib.as_array(mpc.desc.costset_begin, shape=(31,))
        # names sorted: [barrier,] carrot_state, carrot_tail, control_reg, state_limits, state_reg
        base = begin[knot] + (1 if knot < 29 else 0)
        c = costs[base]
        return c.active, pool[c.ref_off:c.ref_off + mpc.nx], costs[base + 1].active

    # t = 0: stage 0 (nav_wp1) is a transition stage => carrot_state inactive on running knots, active on the last knot
    mpc.updateProblem(0)
    for k in range(29):
        assert carrot_state(k)[0] == 0
    act, ref, _ = carrot_state(29)
    # last knot time = 29*30 = 870 ms -> piecewise-constant reference: state_ref[upper_bound(t_ref, 870) - 1] = xs[43]
    assert act == 1 and np.array_equal(ref, xs[43])
    # t = 1500 ms: knots whose time falls into the 0-duration stage wp_1 (clamped to [2000, 2030)) get the carrot
    mpc.updateProblem(1500)
    active = [k for k in range(30) if carrot_state(k)[0]]
    assert 17 in active and 29 in active and 16 not in active and 18 not in active   # 1500 + 17*30 = 2010
    # past the end of the trajectory: carrot_tail switches on with the last q and zero velocity
    mpc.updateProblem(9000)
    act, ref, tail = carrot_state(29)
    assert act == 0 and tail == 1


@pytest.mark.gpu
def test_carrot_closed_loop_gpu_vs_oracle():
    tr, fp, xs, us = _trajectory_solution()
    n_steps = 40
    mpc_g = mpcmod.CarrotMpc(tr, xs, 20, MPC, create_solver=True)
    lat_g, st_g, u_g, it_g = mpcmod.closed_loop(mpc_g, xs, us, xs[0], n_steps, record=True)
    mpc_o = mpcmod.CarrotMpc(tr, xs, 20, MPC, create_solver=False)
    lat_o, st_o, u_o, it_o = ob.oracle_closed_loop(mpc_o, xs, us, xs[0], n_steps, record=True)
    assert it_g == it_o
    _l, st_y, u_y, it_y = ob.oracle_closed_loop(mpcmod.CarrotMpc(tr, xs, 20, MPC, create_solver=False), xs, us, xs[0], n_steps, record=True, nofma=True)
    print("carrot closed loop: u gpu %.1e / self %.1e, x gpu %.1e / self %.1e" % ob.closed_loop_bar(u_g, st_g, (st_o, u_o), (st_y, u_y)))
    print("p50 latency gpu %.3f ms, oracle %.3f ms" % (1e3 * np.median(lat_g), 1e3 * np.median(lat_o)))


# ---- RailMpc / WeightedMpc (BASELINE.json config 5: iris_px4) ---------------------------------------------------------
IRIS_TRAJ = "iris_px4/trajectories/displacement.yaml"
IRIS_MPC = "iris_px4/mpc/mpc.yaml"


def _iris_solution():
    tr = host.Trajectory(IRIS_TRAJ)
    fp = tr.createProblem(20)
    p = ob.default_params(); p.maxiter = 400
    o = ob.Oracle(fp); o.set_params(p); o.set_x0(fp.x0); o.solve()
    return tr, fp, o.get("xs"), o.get("us")


def _cost_records(mpc, knot):
    costs, pool = mpc.cost_tables()
    begin = np.ctypeslib.as_array(mpc.desc.costset_begin, shape=(mpc.knots + 1,))
    return [costs[i] for i in range(begin[knot], begin[knot + 1])], pool


def test_rail_retargeting_matches_reference_rules():
    tr, fp, xs, us = _iris_solution()
    mpc = mpcmod.RailMpc(xs, 20, IRIS_MPC, create_solver=False)
    assert (mpc.knots, mpc.dt, mpc.iters) == (40, 20, 2) and mpc.T == 39
    # per knot: [barrier,] control (Quad, weight rail_control_weight), rail_state (WeightedQuad, weight rail_weight)
    recs, pool = _cost_records(mpc, 0)
    assert [r.type for r in recs] == [abi.COST_SQUASH_BARRIER, abi.COST_CONTROL, abi.COST_STATE]
    assert recs[1].weight == 1e-1 and recs[2].weight == 1000 and recs[2].activation == abi.ACT_WEIGHTED_QUAD
    assert [r.type for r in _cost_records(mpc, 39)[0]] == [abi.COST_CONTROL, abi.COST_STATE]   # terminal knot: no barrier
    # knot i at time t + i dt tracks state_ref[upper_bound(t_ref, time) - 1] (integer-division interpolation, rail-mpc.cpp:188)
    mpc.updateProblem(130)
    for k in (0, 7, 39):
        recs, pool = _cost_records(mpc, k)
        ref = pool[recs[-1].ref_off:recs[-1].ref_off + mpc.nx]
        assert np.array_equal(ref, xs[(130 + 20 * k) // 20])
    # beyond the end: hover at the last configuration, yaw-only quaternion from (w, z), zero velocity (:180-186)
    t_end = 20 * (len(xs) - 1)
    mpc.updateProblem(t_end + 500)
    recs, pool = _cost_records(mpc, 5)
    ref = pool[recs[-1].ref_off:recs[-1].ref_off + mpc.nx]
    last = xs[-1]
    n = np.hypot(last[6], last[5])
    assert np.array_equal(ref[:5], last[:5]) and ref[5] == last[5] / n and ref[6] == last[6] / n
    assert np.all(ref[7:] == 0)


def test_weighted_retargeting_matches_reference_rules():
    tr = host.Trajectory(IRIS_TRAJ)
    n_before = len(tr.stage_names())
    mpc = mpcmod.WeightedMpc(tr, 20, IRIS_MPC, create_solver=False)
    # the transition stages were merged into their successors, in place (weighted-mpc.cpp:63-75)
    names = tr.stage_names()
    assert len(names) < n_before and not any(n.startswith("nav_") for n in names)
    assert mpc.knots == 40
    recs0, _ = _cost_records(mpc, 0)
    n_stage_costs = len(recs0) - 1                      # minus the solver's barrier
    assert n_stage_costs >= 4 * 3                       # every stage's costs live in every knot
    # t = 0: all 40 knots (0 .. 780 ms) fall into the first merged stage (wp_1, 0 .. 2000 ms)
    mpc.updateProblem(0)
    for k in (0, 20, 39):
        recs, _ = _cost_records(mpc, k)
        act = [r for r in recs if r.active and r.type != abi.COST_SQUASH_BARRIER]
        assert len(act) == 4                            # wp_1/{motion,placement}_base_link, wp_1/reg_{control,state}
        w = sorted(r.weight for r in act if r.type in (abi.COST_FRAME_PLACEMENT, abi.COST_FRAME_VELOCITY))
        # task costs: stage weight * exp(alpha (t - t_end_of_stage)) * beta with alpha = 20, beta = 1 (defaults)
        expw = np.exp(20.0 * ((20 * k) - 2000) / 1000.0)
        assert len(w) == 2 and w[0] > 0 and abs(w[0] / expw - round(w[0] / expw)) < 1e-9
    # regularisation costs keep their stage weight
    recs, _ = _cost_records(mpc, 3)
    regs = [r for r in recs if r.active and r.type in (abi.COST_STATE, abi.COST_CONTROL)]
    assert len(regs) == 2
    # later: knots straddle the switch from the first to the second merged stage
    mpc.updateProblem(1700)
    first = [sum(1 for r in _cost_records(mpc, k)[0] if r.active and r.type == abi.COST_FRAME_PLACEMENT) for k in range(40)]
    assert all(v == 1 for v in first)                   # exactly one stage active per knot


def _closed_loop_pair(make, xs, us, n_steps):
    mpc_g = make(True)
    lat_g, st_g, u_g, it_g = mpcmod.closed_loop(mpc_g, xs, us, xs[0], n_steps, record=True)
    mpc_o = make(False)
    lat_o, st_o, u_o, it_o = ob.oracle_closed_loop(mpc_o, xs, us, xs[0], n_steps, record=True)
    assert it_g == it_o
    _l, st_y, u_y, it_y = ob.oracle_closed_loop(make(False), xs, us, xs[0], n_steps, record=True, nofma=True)
    print("closed loop: u gpu %.1e / self %.1e, x gpu %.1e / self %.1e" % ob.closed_loop_bar(u_g, st_g, (st_o, u_o), (st_y, u_y)))
    return 1e3 * np.median(lat_g), 1e3 * np.median(lat_o)


@pytest.mark.gpu
def test_rail_closed_loop_gpu_vs_oracle():
    tr, fp, xs, us = _iris_solution()
    g, o = _closed_loop_pair(lambda s: mpcmod.RailMpc(xs, 20, IRIS_MPC, create_solver=s), xs, us, 30)
    print("rail p50 latency gpu %.3f ms, oracle %.3f ms" % (g, o))


@pytest.mark.gpu
def test_weighted_closed_loop_gpu_vs_oracle():
    _tr, fp, xs, us = _iris_solution()
    g, o = _closed_loop_pair(lambda s: mpcmod.WeightedMpc(host.Trajectory(IRIS_TRAJ), 20, IRIS_MPC, create_solver=s), xs, us, 30)
    print("weighted p50 latency gpu %.3f ms, oracle %.3f ms" % (g, o))


@pytest.mark.gpu
def test_batched_rail_instances_match_oracle():
    """Config 5 shape: one retargeted RailMpc problem, several warm-started instances with different initial states."""
    capi = importlib.import_module("eagle-mpc_b200.capi")
    wl = importlib.import_module("eagle-mpc_b200.workloads")
    _tr, fp, xs, us = _iris_solution()
    mpc = mpcmod.RailMpc(xs, 20, IRIS_MPC, create_solver=False)
    t0 = 1000
    mpc.updateProblem(t0)
    T, i0, B = mpc.knots - 1, t0 // 20, 5
    x0 = wl.noisy_x0(xs[i0], B, 9000)
    xs_w, us_w = xs[i0:i0 + T + 1], us[i0:i0 + T]
    g = capi.BatchSolver(mpc, B)
    costs, pool = mpc.cost_tables()
    g.update_costs(0, costs, 0, pool)
    pg = capi.default_params(); pg.maxiter = mpc.iters; pg.convergence_init = 1e-3
    g.set_params(pg)
    xs_b = np.broadcast_to(xs_w, (B,) + xs_w.shape).copy(); xs_b[:, 0] = x0
    us_b = np.broadcast_to(us_w, (B,) + us_w.shape).copy()
    g.set_x0(x0); g.set_candidate(xs_b, us_b, False); g.solve()
    gx, gu, gc, gi = g.xs(), g.us(), g.cost(), g.iters()
    po = ob.default_params(); po.maxiter = mpc.iters; po.convergence_init = 1e-3
    for b in range(B):
        o = ob.Oracle(mpc); o.set_params(po)
        ob.lib.orc_update_costs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.Cost), C.c_int, C.c_int, abi.c_double_p]
        ob.lib.orc_update_costs(o.p, 0, len(costs), costs, 0, len(pool), ob.dp(pool))
        o.set_x0(x0[b]); o.solve(xs_b[b], us_b[b])
        assert int(o.get("iter")) == gi[b]
        assert abs(gc[b] - o.get("cost")) <= 1e-9 * max(1.0, abs(o.get("cost")))
        assert np.abs(gx[b] - o.get("xs")).max() <= 1e-9 * max(1.0, np.abs(o.get("xs")).max())
        assert np.abs(gu[b] - o.get("us")).max() <= 1e-9 * max(1.0, np.abs(o.get("us")).max())


@pytest.mark.gpu
def test_device_rail_retarget_instances_at_different_times():
    """several rail controllers at different times in one handle, retargeted by one kernel
    (empc_rail_retarget).  Every instance must equal the oracle solving the problem that the host mirror of
    RailMpc::updateProblem retargets at that instance's time, including the hover tail past the end of the reference."""
    capi = importlib.import_module("eagle-mpc_b200.capi")
    wl = importlib.import_module("eagle-mpc_b200.workloads")
    _tr, fp, xs, us = _iris_solution()
    mpc = mpcmod.RailMpc(xs, 20, IRIS_MPC, create_solver=False)
    t_end = 20 * (len(xs) - 1)
    times = [0, 210, 1000, 3999, t_end - 300, t_end + 5000]   # not all multiples of the reference spacing; tail cases
    B, T = len(times), mpc.knots - 1
    x0 = np.zeros((B, mpc.nx)); xs_b = np.zeros((B, T + 1, mpc.nx)); us_b = np.zeros((B, T, mpc.nu))
    for b, t0 in enumerate(times):
        idx = np.minimum(t0 // 20 + np.arange(T + 1), len(xs) - 1)
        x0[b] = wl.noisy_x0(xs[idx[0]], 1, 9000 + b)[0]
        xs_b[b] = xs[idx]; xs_b[b, 0] = x0[b]
        us_b[b] = us[np.minimum(idx[:-1], len(us) - 1)]
    mpc.updateProblem(123456)   # the tables the handle starts from: every knot on the hover state
    g = capi.BatchSolver(mpc, B)
    costs, pool = mpc.cost_tables()
    g.update_costs(0, costs, 0, pool)
    g.replicate_instances(B)
    g.set_reference_trajectory(xs, 20)
    g.rail_retarget(times, mpc.dt)
    pg = capi.default_params(); pg.maxiter = mpc.iters; pg.convergence_init = 1e-3
    g.set_params(pg)
    g.set_x0(x0); g.set_candidate(xs_b, us_b, False); g.solve()
    gx, gu, gc, gi = g.xs(), g.us(), g.cost(), g.iters()
    po = ob.default_params(); po.maxiter = mpc.iters; po.convergence_init = 1e-3
    ob.lib.orc_update_costs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.Cost), C.c_int, C.c_int, abi.c_double_p]
    costs_seen = []
    for b, t0 in enumerate(times):
        mpc.updateProblem(t0)
        costs, pool = mpc.cost_tables()
        o = ob.Oracle(mpc); o.set_params(po)
        ob.lib.orc_update_costs(o.p, 0, len(costs), costs, 0, len(pool), ob.dp(pool))
        o.set_x0(x0[b]); o.solve(xs_b[b], us_b[b])
        assert int(o.get("iter")) == gi[b], (b, t0)
        assert abs(gc[b] - o.get("cost")) <= 1e-9 * max(1.0, abs(o.get("cost"))), (b, t0)
        assert np.abs(gx[b] - o.get("xs")).max() <= 1e-9 * max(1.0, np.abs(o.get("xs")).max()), (b, t0)
        assert np.abs(gu[b] - o.get("us")).max() <= 1e-9 * max(1.0, np.abs(o.get("us")).max()), (b, t0)
        costs_seen.append(float(o.get("cost")))
    assert len({round(c, 6) for c in costs_seen}) > 3   # the instances really solve different problems
    # a second retarget of the same handle (all instances moved on by one controller period) is picked up
    g.rail_retarget([t + 20 for t in times], mpc.dt)
    g.set_x0(x0); g.set_candidate(xs_b, us_b, False); g.solve()
    mpc.updateProblem(times[2] + 20)
    costs, pool = mpc.cost_tables()
    o = ob.Oracle(mpc); o.set_params(po)
    ob.lib.orc_update_costs(o.p, 0, len(costs), costs, 0, len(pool), ob.dp(pool))
    o.set_x0(x0[2]); o.solve(xs_b[2], us_b[2])
    assert abs(g.cost()[2] - o.get("cost")) <= 1e-9 * max(1.0, abs(o.get("cost")))


@pytest.mark.gpu
def test_device_weighted_retarget_instances_at_different_times():
    """Same for WeightedMpc: empc_weighted_retarget walks the knots of every instance on the device (active stage with the
    zero-duration rule, exp(alpha dt) weights, saturation past the end); per instance the solve equals the oracle on the
    problem the host mirror retargets at that time."""
    capi = importlib.import_module("eagle-mpc_b200.capi")
    wl = importlib.import_module("eagle-mpc_b200.workloads")
    _tr, fp, xs, us = _iris_solution()
    mpc = mpcmod.WeightedMpc(host.Trajectory(IRIS_TRAJ), 20, IRIS_MPC, create_solver=False)
    times = [0, 510, 1990, 3999, 7500, 9000]   # stage changes inside the horizon, the end of the trajectory, beyond it
    B, T = len(times), mpc.knots - 1
    x0 = np.zeros((B, mpc.nx)); xs_b = np.zeros((B, T + 1, mpc.nx)); us_b = np.zeros((B, T, mpc.nu))
    for b, t0 in enumerate(times):
        idx = np.minimum(t0 // 20 + np.arange(T + 1), len(xs) - 1)
        x0[b] = wl.noisy_x0(xs[idx[0]], 1, 9000 + b)[0]
        xs_b[b] = xs[idx]; xs_b[b, 0] = x0[b]
        us_b[b] = us[np.minimum(idx[:-1], len(us) - 1)]
    mpc.updateProblem(2500)
    g = capi.BatchSolver(mpc, B)
    costs, pool = mpc.cost_tables()
    g.update_costs(0, costs, 0, pool)
    g.replicate_instances(B)
    g.set_weighted_schedule(mpc.schedule())
    g.weighted_retarget(times, mpc.dt)
    pg = capi.default_params(); pg.maxiter = mpc.iters; pg.convergence_init = 1e-3
    g.set_params(pg)
    g.set_x0(x0); g.set_candidate(xs_b, us_b, False); g.solve()
    gx, gu, gc, gi = g.xs(), g.us(), g.cost(), g.iters()
    po = ob.default_params(); po.maxiter = mpc.iters; po.convergence_init = 1e-3
    ob.lib.orc_update_costs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.Cost), C.c_int, C.c_int, abi.c_double_p]
    costs_seen = []
    for b, t0 in enumerate(times):
        mpc.updateProblem(t0)
        costs, pool = mpc.cost_tables()
        o = ob.Oracle(mpc); o.set_params(po)
        ob.lib.orc_update_costs(o.p, 0, len(costs), costs, 0, len(pool), ob.dp(pool))
        o.set_x0(x0[b]); o.solve(xs_b[b], us_b[b])
        assert int(o.get("iter")) == gi[b], (b, t0)
        assert abs(gc[b] - o.get("cost")) <= 1e-9 * max(1.0, abs(o.get("cost"))), (b, t0)
        assert np.abs(gx[b] - o.get("xs")).max() <= 1e-9 * max(1.0, np.abs(o.get("xs")).max()), (b, t0)
        assert np.abs(gu[b] - o.get("us")).max() <= 1e-9 * max(1.0, np.abs(o.get("us")).max()), (b, t0)
        costs_seen.append(float(o.get("cost")))
    assert len({round(c, 6) for c in costs_seen}) > 3


def test_weighted_schedule_reproduces_updateProblem():
    """WeightedMpc.schedule() (what empc_weighted_retarget consumes) replayed in Python gives the cost tables that the
    host mirror of WeightedMpc::updateProblem writes, at times inside, across and beyond the stages."""
    import math
    mpc = mpcmod.WeightedMpc(host.Trajectory(IRIS_TRAJ), 20, IRIS_MPC, create_solver=False)
    sch = mpc.schedule()
    begin = np.ctypeslib.as_array(mpc.desc.costset_begin, shape=(mpc.knots + 1,))
    assert sch["match"].shape == sch["task"].shape == sch["base"].shape == (4, 16)
    assert (sch["match"].sum(axis=0) == 1).all()        # iris displacement: no stage name is a prefix of another

    def stage_of(t):
        return int(np.searchsorted(sch["t_ini"], t, side="right")) - 1

    for t0 in (0, 510, 1990, 3999, 7500, 9000):
        mpc.updateProblem(t0)
        costs, _pool = mpc.cost_tables()
        last = stage_of(t0)
        for i in range(mpc.knots):
            nt = t0 + i * mpc.dt
            st = stage_of(nt)
            if st == last + 2:
                st -= 1
            wt = 0.0 if nt > sch["duration"] else (nt - int(sch["t_end"][st])) / 1000.0
            w = math.exp(sch["alpha"] * wt)
            slot = 0
            for c in range(begin[i], begin[i + 1]):
                if costs[c].type == abi.COST_SQUASH_BARRIER:
                    continue
                assert costs[c].active == int(sch["match"][st, slot]), (t0, i, slot)
                if sch["task"][st, slot]:
                    assert costs[c].weight == sch["base"][st, slot] * w * sch["beta"], (t0, i, slot)
                slot += 1
            assert slot == 16
            last = st


@pytest.mark.gpu
def test_device_carrot_retarget_instances_at_different_times():
    """Same for CarrotMpc (flying arm, config 3's controller): empc_carrot_retarget switches the carrot on the knots that
    fall into non-transition stages (and on the last knot), moves its reference, and hands over to the tail past the end."""
    capi = importlib.import_module("eagle-mpc_b200.capi")
    wl = importlib.import_module("eagle-mpc_b200.workloads")
    tr, fp, xs, us = _trajectory_solution()
    mpc = mpcmod.CarrotMpc(tr, xs, 20, MPC, create_solver=False)
    t_st, is_tr = mpc.schedule()
    assert len(t_st) == len(is_tr) + 1 and t_st[0] == 0 and is_tr[0] == 1 and is_tr[1] == 0
    times = [0, 1500, 1985, 5200, 7900, 9000]   # before / across the zero-duration way-point stages, the end, the tail
    B, T = len(times), mpc.knots - 1
    x0 = np.zeros((B, mpc.nx)); xs_b = np.zeros((B, T + 1, mpc.nx)); us_b = np.zeros((B, T, mpc.nu))
    for b, t0 in enumerate(times):
        idx = np.minimum((t0 + mpc.dt * np.arange(T + 1)) // 20, len(xs) - 1)
        x0[b] = wl.noisy_x0(xs[idx[0]], 1, 2024 + b)[0]
        xs_b[b] = xs[idx]; xs_b[b, 0] = x0[b]
        us_b[b] = us[np.minimum(idx[:-1], len(us) - 1)]
    g = capi.BatchSolver(mpc, B)   # tables as created: carrot and tail off everywhere
    g.replicate_instances(B)
    g.set_reference_trajectory(xs, 20)
    g.set_carrot_schedule((t_st, is_tr))
    g.carrot_retarget(times, mpc.dt)
    pg = capi.default_params(); pg.maxiter = mpc.iters; pg.convergence_init = 1e-3
    g.set_params(pg)
    g.set_x0(x0); g.set_candidate(xs_b, us_b, False); g.solve()
    gx, gu, gc, gi = g.xs(), g.us(), g.cost(), g.iters()
    po = ob.default_params(); po.maxiter = mpc.iters; po.convergence_init = 1e-3
    ob.lib.orc_update_costs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.Cost), C.c_int, C.c_int, abi.c_double_p]
    costs_seen = []
    for b, t0 in enumerate(times):
        # a fresh controller per instance: "carrot_tail" is never switched off again (carrot-mpc.cpp:349-357), so the
        # tables depend on the controller's history; the device instances start from the tables as created
        mpc_b = mpcmod.CarrotMpc(tr, xs, 20, MPC, create_solver=False)
        mpc_b.updateProblem(t0)
        costs, pool = mpc_b.cost_tables()
        o = ob.Oracle(mpc_b); o.set_params(po)
        ob.lib.orc_update_costs(o.p, 0, len(costs), costs, 0, len(pool), ob.dp(pool))
        o.set_x0(x0[b]); o.solve(xs_b[b], us_b[b])
        assert int(o.get("iter")) == gi[b], (b, t0)
        assert abs(gc[b] - o.get("cost")) <= 1e-9 * max(1.0, abs(o.get("cost"))), (b, t0)
        assert np.abs(gx[b] - o.get("xs")).max() <= 1e-9 * max(1.0, np.abs(o.get("xs")).max()), (b, t0)
        assert np.abs(gu[b] - o.get("us")).max() <= 1e-9 * max(1.0, np.abs(o.get("us")).max()), (b, t0)
        costs_seen.append(float(o.get("cost")))
    assert len({round(c, 6) for c in costs_seen}) > 3


@pytest.mark.gpu
def test_batched_device_closed_loop_rail():
    """Device-resident closed loop of several rail controllers that are at different points of the trajectory
    (retarget -> warm solve -> RK4 plant advance, no host round trip of states or warm starts) against the oracle twin
    of examples/python/mpc.py run once per instance."""
    capi = importlib.import_module("eagle-mpc_b200.capi")
    _tr, fp, xs, us = _iris_solution()
    mpc = mpcmod.RailMpc(xs, 20, IRIS_MPC, create_solver=False)
    starts = [0, 1000, 3010, 7000]
    n_steps, dt_sim = 12, 2
    B, T = len(starts), mpc.knots - 1
    x0 = np.zeros((B, mpc.nx)); xs_b = np.zeros((B, T + 1, mpc.nx)); us_b = np.zeros((B, T, mpc.nu))
    for b, t0 in enumerate(starts):
        idx = np.minimum(t0 // 20 + np.arange(T + 1), len(xs) - 1)
        x0[b] = xs[idx[0]]; xs_b[b] = xs[idx]; us_b[b] = us[np.minimum(idx[:-1], len(us) - 1)]
    g = capi.BatchSolver(mpc, B)
    g.replicate_instances(B)
    g.set_reference_trajectory(xs, 20)
    times = np.array(starts, dtype=np.int64)
    g.rail_retarget(times, mpc.dt)
    pg = capi.default_params(); pg.maxiter = 100; pg.convergence_init = 1e-2
    g.set_params(pg); g.set_x0(x0); g.set_candidate(xs_b, us_b, False); g.solve()
    pg.maxiter = mpc.iters; pg.convergence_init = 1e-3
    g.set_params(pg)
    st_g, u_g, it_g = [x0.copy()], [], []
    for _ in range(n_steps):
        g.rail_retarget(times, mpc.dt)
        g.solve()
        it_g.append(g.iters().copy())
        x, u = g.plant_advance(dt_sim / 1000.0)
        st_g.append(x); u_g.append(u)
        times += dt_sim
    st_g, u_g, it_g = np.array(st_g), np.array(u_g), np.array(it_g)
    for b, t0 in enumerate(starts):
        mpc_o = mpcmod.RailMpc(xs, 20, IRIS_MPC, create_solver=False)
        _lat, st_o, u_o, it_o = ob.oracle_closed_loop(mpc_o, xs, us, x0[b], n_steps, dt_sim_ms=dt_sim, record=True, t_start=t0,
                                                       xs_warm=xs_b[b], us_warm=us_b[b])
        assert list(it_g[:, b]) == it_o, (b, t0)
        _l, st_y, u_y, it_y = ob.oracle_closed_loop(mpcmod.RailMpc(xs, 20, IRIS_MPC, create_solver=False), xs, us, x0[b], n_steps, dt_sim_ms=dt_sim,
                                                    record=True, t_start=t0, xs_warm=xs_b[b], us_warm=us_b[b], nofma=True)
        ob.closed_loop_bar(u_g[:, b], st_g[:, b], (st_o, u_o), (st_y, u_y))


# ---- golden cost tables (tests/golden/mpc_cost_tables.json, written by tests/golden/make_golden_mpc.py) --------------------
GOLD_CASES = {"carrot": (TRAJ, MPC), "rail": (IRIS_TRAJ, IRIS_MPC), "weighted": (IRIS_TRAJ, IRIS_MPC)}


def _gold_controller(kind, tr, xs):
    traj_yaml, mpc_yaml = GOLD_CASES[kind]
    if kind == "carrot":
        return mpcmod.CarrotMpc(tr, xs, 20, mpc_yaml, create_solver=False)
    if kind == "rail":
        return mpcmod.RailMpc(xs, 20, mpc_yaml, create_solver=False)
    return mpcmod.WeightedMpc(host.Trajectory(traj_yaml), 20, mpc_yaml, create_solver=False)


def _check_table(costs, pool, nx, gold, tag):
    """active flags always; weights and reference checksums of the active records (inactive ones keep stale values)"""
    assert len(costs) == len(gold), tag
    for i, (c, g) in enumerate(zip(costs, gold)):
        assert (c.type, c.active) == (g[0], g[1]), (tag, i)
        if c.active:
            assert abs(c.weight - g[2]) <= 4e-16 * abs(g[2]), (tag, i, c.weight, g[2])   # device exp(): last-bit differences
            if c.type == abi.COST_STATE and c.ref_off >= 0:
                ref = pool[c.ref_off:c.ref_off + nx]
                assert abs(float(np.dot(ref, np.arange(1, nx + 1))) - g[3]) <= 1e-9 * max(1.0, abs(g[3])), (tag, i)  # the reference states come from an oracle solve: not bit-stable across oracle builds


@pytest.mark.parametrize("kind", ["carrot", "rail", "weighted"])
def test_host_retargeting_matches_golden_tables(kind):
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpc_cost_tables.json")))[kind]
    tr, _fp, xs, _us = _trajectory_solution() if kind == "carrot" else _iris_solution()
    for t, g in gold.items():
        mpc = _gold_controller(kind, tr, xs)
        mpc.updateProblem(int(t))
        costs, pool = mpc.cost_tables()
        _check_table(costs, pool, mpc.nx, g, (kind, t))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["carrot", "rail", "weighted"])
def test_device_retargeting_matches_golden_tables(kind):
    """The tables the retarget kernels leave on the device, read back and compared with the committed golden tables
    (no oracle, no host retargeting at run time)."""
    import json, os
    capi = importlib.import_module("eagle-mpc_b200.capi")
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpc_cost_tables.json")))[kind]
    tr, _fp, xs, _us = _trajectory_solution() if kind == "carrot" else _iris_solution()
    mpc = _gold_controller(kind, tr, xs)
    times = [int(t) for t in gold]
    g = capi.BatchSolver(mpc, len(times))
    g.replicate_instances(len(times))
    if kind == "rail":
        g.set_reference_trajectory(xs, 20); g.rail_retarget(times, mpc.dt)
    elif kind == "carrot":
        g.set_reference_trajectory(xs, 20); g.set_carrot_schedule(mpc.schedule()); g.carrot_retarget(times, mpc.dt)
    else:
        g.set_weighted_schedule(mpc.schedule()); g.weighted_retarget(times, mpc.dt)
    costs, pool = g.cost_tables(mpc.n_costs * len(times), mpc.n_pool * len(times))
    for m, t in enumerate(times):
        _check_table(costs[m * mpc.n_costs:(m + 1) * mpc.n_costs], pool, mpc.nx, gold[str(t)], (kind, t))
