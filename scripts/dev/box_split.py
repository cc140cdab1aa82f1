"""per-kernel split of a Box-solver solve vs SbFDDP on the same trajectory (CUDA events inside the library)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
host = importlib.import_module("eagle-mpc_b200.host"); capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads"); abi = importlib.import_module("eagle-mpc_b200.abi")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for yaml in ("hexacopter370_flying_arm_3/trajectories/displacement.yaml", "iris/trajectories/loop.yaml"):
    for box in (None, abi.SOLVER_BOXFDDP):
        fp = host.Trajectory(yaml).createProblem(20, box is None, "IntegratedActionModelEuler")
        g = capi.BatchSolver(fp, B)
        if box is not None:
            p = capi.box_params(box); p.maxiter = 20; g.set_params(p)
        g.enable_kernel_timing(True)
        x0 = wl.noisy_x0(fp.x0, B, 777)
        for rep in range(2):
            g.set_x0(x0); g.set_candidate(None, None, False); g.solve()
        n, ms = g.launch_stats()
        sms, units = g.solve_stats()
        its = int((g.iters() + 1).max())
        print(yaml.split("/")[0], "box" if box else "sbfddp", "batch-iterations", its, "solve ms %.1f" % sms,
              "per batch-iteration ms: calc_diff %.2f backward %.2f rollout %.2f decide %.2f" % tuple(ms / its))
