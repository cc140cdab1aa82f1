#!/usr/bin/env python3
"""Counterpart of the reference's examples/python/mpc.py: solve the trajectory (maxiter 400), build a carrot / rail /
weighted MPC controller on it, and run the closed loop against the RK4 plant (2 ms).  With --instances N > 1 the loop
runs N controllers that sit at different times of the trajectory in one handle, device-resident: per step one retarget
kernel, one warm-started batched solve, one plant kernel.  Needs a CUDA device."""
import argparse
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
mpcmod = importlib.import_module("eagle-mpc_b200.mpc")

ROBOTS = {"flying_arm_3": ("hexacopter370_flying_arm_3/trajectories/displacement.yaml", "hexacopter370_flying_arm_3/mpc/mpc.yaml"),
          "iris_px4": ("iris_px4/trajectories/displacement.yaml", "iris_px4/mpc/mpc.yaml")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robot", choices=sorted(ROBOTS), default="flying_arm_3")
    ap.add_argument("--controller", choices=["carrot", "rail", "weighted"], default="carrot")
    ap.add_argument("--instances", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    traj_yaml, mpc_yaml = ROBOTS[args.robot]
    dt_traj, dt_sim = 20, 2
    trajectory = host.Trajectory(traj_yaml)
    problem = trajectory.createProblem(dt_traj)
    s = capi.BatchSolver(problem, 1)
    p = capi.default_params(); p.maxiter = 400                      # examples/python/mpc.py:29
    s.set_params(p); s.set_x0(problem.x0); s.set_candidate(None, None, False); s.solve()
    xs, us = s.xs()[0], s.us()[0]
    s.close()
    single = args.instances == 1
    if args.controller == "carrot":
        mpc = mpcmod.CarrotMpc(trajectory, xs, dt_traj, mpc_yaml, create_solver=single)
    elif args.controller == "rail":
        mpc = mpcmod.RailMpc(xs, dt_traj, mpc_yaml, create_solver=single)
    else:
        mpc = mpcmod.WeightedMpc(host.Trajectory(traj_yaml), dt_traj, mpc_yaml, create_solver=single)
    if single:
        lat, states, controls, iters = mpcmod.closed_loop(mpc, xs, us, xs[0], args.steps, dt_sim_ms=dt_sim, record=True)
        print(f"{args.controller} MPC, {mpc.knots} knots: step p50 {1e3 * np.median(lat):.3f} ms, p95 {1e3 * np.percentile(lat, 95):.3f} ms; "
              f"position after {args.steps * dt_sim} ms {np.round(states[-1][:3], 4)} (trajectory {np.round(xs[args.steps * dt_sim // dt_traj][:3], 4)})")
        return
    B, T = args.instances, mpc.knots - 1
    t_end = dt_traj * (len(xs) - 1)
    times = ((t_end // B) * np.arange(B)).astype(np.int64)          # the instances are spread over the trajectory
    idx = np.minimum((times[:, None] + mpc.dt * np.arange(T + 1)[None, :]) // dt_traj, len(xs) - 1)
    xs_b = xs[idx]; us_b = us[np.minimum(idx[:, :-1], len(us) - 1)]
    g = capi.BatchSolver(mpc, B)
    g.replicate_instances(B)
    if args.controller == "rail":
        g.set_reference_trajectory(xs, dt_traj); retarget = g.rail_retarget
    elif args.controller == "carrot":
        g.set_reference_trajectory(xs, dt_traj); g.set_carrot_schedule(mpc.schedule()); retarget = g.carrot_retarget
    else:
        g.set_weighted_schedule(mpc.schedule()); retarget = g.weighted_retarget
    retarget(times, mpc.dt)
    p = capi.default_params(); p.maxiter = 100; p.convergence_init = 1e-2
    g.set_params(p); g.set_x0(xs_b[:, 0].copy()); g.set_candidate(xs_b, us_b, False); g.solve()
    p.maxiter = mpc.iters; p.convergence_init = 1e-3
    g.set_params(p)
    lat = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        retarget(times, mpc.dt); g.solve(); x, u = g.plant_advance(dt_sim / 1000.0)
        lat.append(time.perf_counter() - t0)
        times += dt_sim
    err = np.linalg.norm(x[:, :3] - xs[np.minimum(times // dt_traj, len(xs) - 1), :3], axis=1)
    print(f"{B} {args.controller} controllers, {mpc.knots} knots: batched step p50 {1e3 * np.median(lat):.3f} ms "
          f"({1e6 * np.median(lat) / B:.2f} us per controller-step); position error vs the trajectory after {args.steps} steps: "
          f"median {np.median(err):.4f} m, max {err.max():.4f} m")


if __name__ == "__main__":
    main()
