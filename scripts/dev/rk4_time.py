import importlib, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
host = importlib.import_module("eagle-mpc_b200.host"); capi = importlib.import_module("eagle-mpc_b200.capi"); wl = importlib.import_module("eagle-mpc_b200.workloads")
fp = host.Trajectory("hexacopter370_flying_arm_3/trajectories/move_arm.yaml").createProblem(20, True, "IntegratedActionModelRK4")
B = 1024
x0 = wl.noisy_x0(fp.x0, B, 777)
g = capi.BatchSolver(fp, B)
for rep in range(2):
    g.set_x0(x0); g.set_candidate(None, None, False); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.solve(); e1.record(); torch.cuda.synchronize()
print(os.environ.get("EMPC_RK4_BLOCKS"), "ms", e0.elapsed_time(e1), "iters mean", (g.iters() + 1).mean())
