"""Throughput of the overlays (contact dynamics, RK4 integrator, the Box solvers) next to the tuned Euler / free path on the same
trajectories: OCP-iterations/s of a batched solve, CUDA events around empc_solve.  Not part of bench.py's headline (no
BASELINE.json config uses them); numbers go to profiles/."""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads")
abi = importlib.import_module("eagle-mpc_b200.abi")

CASES = [
    ("eagle_catch (contact, 10 of 160 knots)", "hexacopter370_flying_arm_3/trajectories/eagle_catch.yaml", "IntegratedActionModelEuler", 1024),
    ("eagle_catch_nc (same trajectory, no contact)", "hexacopter370_flying_arm_3/trajectories/eagle_catch_nc.yaml", "IntegratedActionModelEuler", 1024),
    ("monkey_bar (contact, 70 of 95 knots)", "hexacopter370_flying_arm_3/trajectories/monkey_bar.yaml", "IntegratedActionModelEuler", 1024),
    ("move_arm, Euler", "hexacopter370_flying_arm_3/trajectories/move_arm.yaml", "IntegratedActionModelEuler", 1024),
    ("move_arm, RK4", "hexacopter370_flying_arm_3/trajectories/move_arm.yaml", "IntegratedActionModelRK4", 1024),
    # crocoddyl's box solvers on the unsquashed problem (squash = False), against SbFDDP on the same trajectory
    ("displacement, SolverSbFDDP", "hexacopter370_flying_arm_3/trajectories/displacement.yaml", "IntegratedActionModelEuler", 1024),
    ("displacement, SolverBoxFDDP", "hexacopter370_flying_arm_3/trajectories/displacement.yaml", "IntegratedActionModelEuler", 1024, abi.SOLVER_BOXFDDP),
    ("displacement, SolverBoxDDP", "hexacopter370_flying_arm_3/trajectories/displacement.yaml", "IntegratedActionModelEuler", 1024, abi.SOLVER_BOXDDP),
    ("iris loop, SolverBoxFDDP", "iris/trajectories/loop.yaml", "IntegratedActionModelEuler", 1024, abi.SOLVER_BOXFDDP),
]

out = []
for case in CASES:
    label, yaml, integ, B = case[:4]
    box = case[4] if len(case) > 4 else None
    fp = host.Trajectory(yaml).createProblem(20, box is None, integ)
    x0 = np.tile(fp.x0, (B, 1)) if "monkey" in yaml else wl.noisy_x0(fp.x0, B, 777)
    g = capi.BatchSolver(fp, B)
    if box is not None:
        g.set_params(capi.box_params(box))
    ms = []
    for rep in range(3):
        g.set_x0(x0); g.set_candidate(None, None, False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.solve(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    iters = g.iters()
    rec = {"case": label, "T": fp.T, "batch": B, "integrator": integ[21:], "iterations_per_ocp_mean": float((iters + 1).mean()),
           "ms_per_solve": float(np.median(ms)), "ocp_iterations_per_s": float((iters + 1).sum() / (np.median(ms) * 1e-3))}
    print(json.dumps(rec)); out.append(rec)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_overlays.json"), "w"), indent=1)
