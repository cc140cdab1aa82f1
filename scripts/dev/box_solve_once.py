"""one SolverBoxFDDP solve of the flying-arm displacement batch (for ncu captures of the Box instantiation)"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
host = importlib.import_module("eagle-mpc_b200.host"); capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads"); abi = importlib.import_module("eagle-mpc_b200.abi")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
fp = host.Trajectory("hexacopter370_flying_arm_3/trajectories/displacement.yaml").createProblem(20, False, "IntegratedActionModelEuler")
g = capi.BatchSolver(fp, B)
p = capi.box_params(abi.SOLVER_BOXFDDP); p.maxiter = 12; g.set_params(p)
g.set_x0(wl.noisy_x0(fp.x0, B, 777)); g.set_candidate(None, None, False); g.solve()
print("iters", int((g.iters() + 1).max()))
