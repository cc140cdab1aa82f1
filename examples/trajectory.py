#!/usr/bin/env python3
"""Counterpart of the reference's examples/python/trajectory.py on the B200 path: build the shooting problem of a
trajectory YAML (dt = 20 ms, squashing, Euler), solve it with SbFDDP (maxiter 100) and print what the reference's script
looks at.  With --batch N the same problem is solved for N perturbed initial states at once.  Needs a CUDA device."""
import argparse
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--yaml", default="hexacopter370_flying_arm_3/trajectories/displacement.yaml")
    ap.add_argument("--dt", type=int, default=20)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--maxiter", type=int, default=100)
    args = ap.parse_args()
    trajectory = host.Trajectory(args.yaml)                 # eagle_mpc.Trajectory(); autoSetup(yaml)
    problem = trajectory.createProblem(args.dt)             # createProblem(dt, squash=True, "IntegratedActionModelEuler")
    solver = capi.BatchSolver(problem, args.batch)          # eagle_mpc.SolverSbFDDP(problem, trajectory.squash)
    x0 = np.vstack([problem.x0[None, :], wl.noisy_x0(problem.x0, args.batch - 1, 2024)]) if args.batch > 1 else problem.x0[None, :]
    p = capi.default_params(); p.maxiter = args.maxiter
    solver.set_params(p)
    solver.set_x0(x0)
    solver.set_candidate(None, None, False)                 # solver.solve([], [], maxiter)
    solver.solve()
    xs, us, uss, cost, stop, iters, feas = solver.solution()
    for b in range(min(args.batch, 8)):
        print(f"OCP {b}: iterations {iters[b]}, cost {cost[b]:.6f}, stop {stop[b]:.3e}, feasible {bool(feas[b])}, "
              f"final position {np.round(xs[b, -1, :3], 4)}, max squashed thrust {uss[b, :, :problem.desc.n_rotors].max():.3f}")


if __name__ == "__main__":
    main()
