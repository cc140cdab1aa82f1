"""CarrotMpc: host-side retargeting (CPU) and closed-loop parity GPU vs oracle (GPU)."""
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
    assert np.abs(u_g - u_o).max() <= 1e-7 * max(1.0, np.abs(u_o).max())
    assert np.abs(st_g - st_o).max() <= 1e-8 * max(1.0, np.abs(st_o).max())
    print("p50 latency gpu %.3f ms, oracle %.3f ms" % (1e3 * np.median(lat_g), 1e3 * np.median(lat_o)))
