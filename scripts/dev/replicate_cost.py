#!/usr/bin/env python3
"""How much does giving every instance private cost tables cost?  rail, iris_px4, same controller time for all instances:
shared tables vs replicated tables (identical problems, identical iterations), per kernel family."""
import importlib, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
host = importlib.import_module("eagle-mpc_b200.host"); capi = importlib.import_module("eagle-mpc_b200.capi")
mpcmod = importlib.import_module("eagle-mpc_b200.mpc"); wl = importlib.import_module("eagle-mpc_b200.workloads")
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bench_mpc_sweep as sw

tr = host.Trajectory(sw.TRAJ); fp = tr.createProblem(20)
s1 = capi.BatchSolver(fp, 1); p = capi.default_params(); p.maxiter = 400
s1.set_params(p); s1.set_x0(fp.x0); s1.set_candidate(None, None, False); s1.solve()
xs, us = s1.xs()[0], s1.us()[0]; s1.close()
B, t0 = 1024, 1000
tmp = tempfile.mkdtemp()
for knots in (50, 200):
    for mode in ("shared", "replicated", "replicated+spread"):
        mpc = mpcmod.RailMpc(xs, 20, sw.yaml_with_knots(knots, tmp), create_solver=False)
        mpc.updateProblem(t0)
        T = mpc.knots - 1
        times = np.full(B, t0, dtype=np.int64) if mode != "replicated+spread" else (7 * np.arange(B)) % 8000
        idx = np.minimum((times // 20)[:, None] + np.arange(T + 1)[None, :], len(xs) - 1)
        xs_b = xs[idx]; us_b = us[np.minimum(idx[:, :-1], len(us) - 1)]
        x0 = np.stack([wl.noisy_x0(xs[idx[b, 0]], 1, 9000 + b)[0] for b in range(B)]); xs_b[:, 0] = x0
        g = capi.BatchSolver(mpc, B)
        costs, pool = mpc.cost_tables(); g.update_costs(0, costs, 0, pool)
        if mode != "shared":
            g.replicate_instances(B); g.set_reference_trajectory(xs, 20); g.rail_retarget(times, mpc.dt)
        pr = capi.default_params(); pr.maxiter = mpc.iters; pr.convergence_init = 1e-3
        g.set_params(pr); g.set_x0(x0); g.set_candidate(xs_b, us_b, False); g.solve()
        g.enable_kernel_timing(True)
        tot = np.zeros(4); n = 0; wall = 0
        for _ in range(5):
            g.reset(); t1 = time.perf_counter(); g.solve(); wall += time.perf_counter() - t1
            nl, ms = g.launch_stats(); tot += ms; n += 1
        print(knots, mode, "wall ms", round(1e3 * wall / n, 2), "launches", nl, "by kernel", np.round(tot / n, 2), "iters/inst", g.total_iterations() / B)
        g.close()
