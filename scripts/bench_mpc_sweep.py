#!/usr/bin/env python3
"""BASELINE.json config 5: iris_px4 Rail / Weighted MPC, horizon sweep, B batched warm-started instances.

For every controller and every `knots` value: one MPC problem (cost tables retargeted by the host-side controller at time
t0), B instances that differ in their initial state (trajectory state at t0 + the benchmark's noise recipe, seeds 9000+b),
warm-started from the trajectory slice, `iters` SbFDDP iterations each (mpc.yaml: 2).  Prints one JSON line per case.
Not a bench.py line (bench.py measures config 2); evidence for profiles/.
"""
import argparse, importlib, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
mpcmod = importlib.import_module("eagle-mpc_b200.mpc")
wl = importlib.import_module("eagle-mpc_b200.workloads")

TRAJ = "iris_px4/trajectories/displacement.yaml"
MPC_YAML = os.path.join(ROOT, "yaml", "iris_px4", "mpc", "mpc.yaml")


def yaml_with_knots(knots, tmpdir):
    txt = open(MPC_YAML).read().replace("knots: 40", f"knots: {knots}")
    p = os.path.join(tmpdir, f"mpc_{knots}.yaml")
    open(p, "w").write(txt)
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--knots", type=int, nargs="+", default=[50, 100, 200, 400])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--t0", type=int, default=1000, help="controller time (ms) at which the instances are solved")
    ap.add_argument("--spread", action="store_true",
                    help="every instance at its own controller time, retargeted on the device (empc_rail_retarget / empc_weighted_retarget)")
    ap.add_argument("--closed-loop", action="store_true",
                    help="device-resident closed loop of B controllers at different times: retarget -> warm solve -> RK4 plant")
    args = ap.parse_args()
    # reference trajectory: iris_px4 displacement solved by the B200 path itself (B = 1, maxiter 400)
    tr = host.Trajectory(TRAJ)
    fp = tr.createProblem(20)
    s1 = capi.BatchSolver(fp, 1)
    p = capi.default_params(); p.maxiter = 400
    s1.set_params(p); s1.set_x0(fp.x0); s1.set_candidate(None, None, False); s1.solve()
    xs, us = s1.xs()[0], s1.us()[0]
    s1.close()
    tmpdir = tempfile.mkdtemp()
    B = args.batch
    if args.spread:
        return spread(args, xs, us, tmpdir)
    if args.closed_loop:
        return closed_loop(args, xs, us)
    for kind in ("rail", "weighted"):
        for knots in args.knots:
            y = yaml_with_knots(knots, tmpdir)
            mpc = (mpcmod.RailMpc(xs, 20, y, create_solver=False) if kind == "rail"
                   else mpcmod.WeightedMpc(host.Trajectory(TRAJ), 20, y, create_solver=False))
            mpc.updateProblem(args.t0)
            T = mpc.knots - 1
            i0 = args.t0 // 20
            idx = np.minimum(i0 + np.arange(T + 1), len(xs) - 1)
            xs_w = xs[idx]
            us_w = us[np.minimum(idx[:-1], len(us) - 1)]
            x0 = wl.noisy_x0(xs[i0], B, 9000)
            g = capi.BatchSolver(mpc, B)
            costs, pool = mpc.cost_tables()
            g.update_costs(0, costs, 0, pool)
            pr = capi.default_params(); pr.maxiter = mpc.iters; pr.convergence_init = 1e-3
            g.set_params(pr)
            xs_b = np.broadcast_to(xs_w, (B,) + xs_w.shape).copy()
            xs_b[:, 0] = x0
            us_b = np.broadcast_to(us_w, (B,) + us_w.shape).copy()
            g.set_x0(x0); g.set_candidate(xs_b, us_b, False)
            g.solve()  # warm-up
            t_tot, it_tot = 0.0, 0
            for _ in range(args.steps):
                g.reset()
                t1 = time.perf_counter(); g.solve(); t_tot += time.perf_counter() - t1
                it_tot += g.total_iterations()
            print(json.dumps({"controller": kind, "knots": knots, "T": T, "batch": B, "iters_per_instance": it_tot / args.steps / B,
                              "ms_per_batched_mpc_step": 1e3 * t_tot / args.steps, "ocp_iterations_per_s": it_tot / t_tot,
                              "us_per_instance_step": 1e6 * t_tot / args.steps / B}))
            g.close()


def closed_loop(args, xs_iris, us_iris):
    """examples/python/mpc.py:49-61 for B controllers at once, everything device-resident between steps: per step one
    retarget kernel, one warm-started batched solve (mpc.yaml iters), one RK4 plant kernel (2 ms); only the B controller
    times go to the device and the applied controls / plant states come back."""
    B, dt_sim = args.batch, 2
    cases = [("rail", "iris_px4", TRAJ, MPC_YAML, xs_iris, us_iris)]
    arm_traj = "hexacopter370_flying_arm_3/trajectories/displacement.yaml"
    tr = host.Trajectory(arm_traj); fp = tr.createProblem(20)
    s1 = capi.BatchSolver(fp, 1); p = capi.default_params(); p.maxiter = 400
    s1.set_params(p); s1.set_x0(fp.x0); s1.set_candidate(None, None, False); s1.solve()
    xa, ua = s1.xs()[0], s1.us()[0]; s1.close()
    cases.append(("carrot", "hexacopter370_flying_arm_3", arm_traj, "hexacopter370_flying_arm_3/mpc/mpc.yaml", xa, ua))
    cases.append(("weighted", "iris_px4", TRAJ, MPC_YAML, xs_iris, us_iris))
    for kind, robot, traj, yaml, xs, us in cases:
        if kind == "rail":
            mpc = mpcmod.RailMpc(xs, 20, yaml, create_solver=False)
        elif kind == "carrot":
            mpc = mpcmod.CarrotMpc(host.Trajectory(traj), xs, 20, yaml, create_solver=False)
        else:
            mpc = mpcmod.WeightedMpc(host.Trajectory(traj), 20, yaml, create_solver=False)
        T = mpc.knots - 1
        t_end = 20 * (len(xs) - 1)
        times = ((7 * np.arange(B)) % t_end).astype(np.int64)
        idx = np.minimum((times[:, None] + mpc.dt * np.arange(T + 1)[None, :]) // 20, len(xs) - 1)
        xs_b = xs[idx]; us_b = us[np.minimum(idx[:, :-1], len(us) - 1)]
        x0 = xs_b[:, 0].copy()
        g = capi.BatchSolver(mpc, B)
        g.replicate_instances(B)
        if kind == "rail":
            g.set_reference_trajectory(xs, 20); retarget = g.rail_retarget
        elif kind == "carrot":
            g.set_reference_trajectory(xs, 20); g.set_carrot_schedule(mpc.schedule()); retarget = g.carrot_retarget
        else:
            g.set_weighted_schedule(mpc.schedule()); retarget = g.weighted_retarget
        retarget(times, mpc.dt)
        pr = capi.default_params(); pr.maxiter = 100; pr.convergence_init = 1e-2
        g.set_params(pr); g.set_x0(x0); g.set_candidate(xs_b, us_b, False); g.solve()
        pr.maxiter = mpc.iters; pr.convergence_init = 1e-3
        g.set_params(pr)
        lat, its = [], 0
        n_steps = 10 * args.steps
        for step in range(n_steps + 5):
            t1 = time.perf_counter()
            retarget(times, mpc.dt); g.solve(); g.plant_advance(dt_sim / 1000.0)
            t2 = time.perf_counter()
            if step >= 5:
                lat.append(t2 - t1); its += g.total_iterations()
            times += dt_sim
        lat = np.array(lat)
        print(json.dumps({"controller": kind, "robot": robot, "mode": "device-resident closed loop", "knots": mpc.knots, "batch": B,
                          "steps": n_steps, "iters_per_instance_step": its / n_steps / B,
                          "ms_per_batched_step_p50": 1e3 * float(np.median(lat)), "ms_per_batched_step_p95": 1e3 * float(np.percentile(lat, 95)),
                          "instance_steps_per_s": B / float(np.median(lat)), "us_per_instance_step": 1e6 * float(np.median(lat)) / B}))
        g.close()


def spread(args, xs, us, tmpdir):
    """B controllers at different times (7 ms apart, wrapping over the trajectory): device-side retargeting of all
    instances by one kernel vs the host-side updateProblem loop run once per instance."""
    B = args.batch
    t_end = 20 * (len(xs) - 1)
    times = (7 * np.arange(B)) % t_end
    for kind, knots in [(k, n) for k in ("rail", "weighted") for n in args.knots]:
        rail = kind == "rail"
        mpc = (mpcmod.RailMpc(xs, 20, yaml_with_knots(knots, tmpdir), create_solver=False) if rail
               else mpcmod.WeightedMpc(host.Trajectory(TRAJ), 20, yaml_with_knots(knots, tmpdir), create_solver=False))
        T = mpc.knots - 1
        idx = np.minimum((times // 20)[:, None] + np.arange(T + 1)[None, :], len(xs) - 1)
        xs_b = xs[idx]
        us_b = us[np.minimum(idx[:, :-1], len(us) - 1)]
        x0 = np.stack([wl.noisy_x0(xs[idx[b, 0]], 1, 9000 + b)[0] for b in range(B)])
        xs_b[:, 0] = x0
        mpc.updateProblem(0)
        g = capi.BatchSolver(mpc, B)
        costs, pool = mpc.cost_tables()
        g.update_costs(0, costs, 0, pool)
        g.replicate_instances(B)
        if rail:
            g.set_reference_trajectory(xs, 20)
            retarget = g.rail_retarget
        else:
            g.set_weighted_schedule(mpc.schedule())
            retarget = g.weighted_retarget
        pr = capi.default_params(); pr.maxiter = mpc.iters; pr.convergence_init = 1e-3
        g.set_params(pr)
        retarget(times, mpc.dt)
        g.set_x0(x0); g.set_candidate(xs_b, us_b, False)
        g.solve()  # warm-up
        t_ret = t_sol = 0.0
        it_tot = 0
        for s in range(args.steps):
            g.reset()
            t1 = time.perf_counter(); retarget(times + 20 * (s + 1), mpc.dt); t2 = time.perf_counter()
            g.solve(); t3 = time.perf_counter()
            t_ret += t2 - t1; t_sol += t3 - t2
            it_tot += g.total_iterations()
        n_host = min(B, 64)
        t1 = time.perf_counter()
        for b in range(n_host):
            mpc.updateProblem(int(times[b]))
        t_host = (time.perf_counter() - t1) / n_host
        print(json.dumps({"controller": kind, "mode": "instances at different times, device retarget", "knots": knots, "T": T,
                          "batch": B, "iters_per_instance": it_tot / args.steps / B,
                          "ms_retarget_all_instances": 1e3 * t_ret / args.steps,
                          "ms_host_updateProblem_all_instances": 1e3 * t_host * B,
                          "ms_per_batched_mpc_step": 1e3 * (t_ret + t_sol) / args.steps,
                          "ocp_iterations_per_s": it_tot / (t_ret + t_sol),
                          "us_per_instance_step": 1e6 * (t_ret + t_sol) / args.steps / B}))
        g.close()


if __name__ == "__main__":
    main()
