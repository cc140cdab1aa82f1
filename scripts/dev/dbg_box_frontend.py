import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eagle-mpc_b200", "python"))
host = importlib.import_module("eagle-mpc_b200.host"); capi = importlib.import_module("eagle-mpc_b200.capi"); abi = importlib.import_module("eagle-mpc_b200.abi")
import eagle_mpc
from eagle_mpc.utils.path import EAGLE_MPC_YAML_DIR
yaml = "iris/trajectories/loop.yaml"; EULER = "IntegratedActionModelEuler"
L = capi.lib()
def log_of(handle, n=512):
    rec = (abi.IterRecord * n)(); cnt = C.c_int32(0)
    L.empc_get_iteration_log(handle, 0, rec, n, C.byref(cnt))
    return [rec[i] for i in range(cnt.value)]
tr = host.Trajectory(yaml)
s = host.SolverBoxFDDP(tr, 20, EULER)
L.empc_enable_iteration_log(s.handle, 512)
s.solve(60)
la = log_of(s.handle)
t2 = eagle_mpc.Trajectory(); t2.autoSetup(EAGLE_MPC_YAML_DIR + "/" + yaml)
problem = t2.createProblem(20, False, EULER)
sb = eagle_mpc.SolverBoxFDDP(problem)
L.empc_enable_iteration_log(C.c_void_p(sb.handle), 512)
sb.solve([], [], 60)
lb = log_of(C.c_void_p(sb.handle))
print("ctypes iters", len(la), "pybind iters", len(lb), sb.iter, sb.cost, sb.stop)
for i in range(min(6, len(la), len(lb))):
    print(i, la[i].cost, lb[i].cost, la[i].stop, lb[i].stop, la[i].steplength, lb[i].steplength, la[i].is_feasible, lb[i].is_feasible, la[i].phase, lb[i].phase)
print("last", la[-1].cost, la[-1].stop, lb[-1].cost, lb[-1].stop)
sb2 = eagle_mpc.SolverBoxFDDP(problem); sb2.solve([], [], 60); print("second pybind solver:", sb2.iter, sb2.cost)
