#!/usr/bin/env python3
"""Where does the time of one single-instance carrot-MPC solve go (B = 1, T = 29, iters 2)?"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
host = importlib.import_module("eagle-mpc_b200.host"); capi = importlib.import_module("eagle-mpc_b200.capi")
mpcmod = importlib.import_module("eagle-mpc_b200.mpc")
traj = "hexacopter370_flying_arm_3/trajectories/displacement.yaml"
tr = host.Trajectory(traj); fp = tr.createProblem(20)
s1 = capi.BatchSolver(fp, 1); p = capi.default_params(); p.maxiter = 400
s1.set_params(p); s1.set_x0(fp.x0); s1.set_candidate(None, None, False); s1.solve()
xs, us = s1.xs()[0], s1.us()[0]; s1.close()
mpc = mpcmod.CarrotMpc(host.Trajectory(traj), xs, 20, "hexacopter370_flying_arm_3/mpc/mpc.yaml", create_solver=False)
mpc.updateProblem(500)
for B in (1, 32):
    g = capi.BatchSolver(mpc, B)
    costs, pool = mpc.cost_tables(); g.update_costs(0, costs, 0, pool)
    T = mpc.knots - 1
    idx = np.minimum((500 + mpc.dt * np.arange(T + 1)) // 20, len(xs) - 1)
    xs_b = np.broadcast_to(xs[idx], (B, T + 1, g.nx)).copy(); us_b = np.broadcast_to(us[idx[:-1]], (B, T, g.nu)).copy()
    pr = capi.default_params(); pr.maxiter = 2; pr.convergence_init = 1e-3
    g.set_params(pr); g.set_x0(xs_b[:, 0]); g.set_candidate(xs_b, us_b, False); g.solve()
    for timing in (False, True):
        g.enable_kernel_timing(timing)
        w = []; tot = np.zeros(4)
        for _ in range(50):
            g.reset(); t1 = time.perf_counter(); g.solve(); w.append(time.perf_counter() - t1)
            nl, ms = g.launch_stats(); tot += ms
        dev_ms, _u = g.solve_stats()
        print("B", B, "timing", timing, "wall p50 ms", round(1e3 * float(np.median(w)), 3), "device ms (last)", round(dev_ms, 3), "launches", nl,
              "by kernel [calc_diff backward rollout decide]", np.round(tot / 50, 3), "iters", g.total_iterations())
    g.close()
