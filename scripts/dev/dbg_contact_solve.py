import sys, os, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_binding as ob
from test_gpu_contact import MONKEY
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads")
fp = host.Trajectory(MONKEY).createProblem(20)
x0 = wl.noisy_x0(fp.x0, 3, 4242)[1:2]
g = capi.BatchSolver(fp, 1); g.enable_iteration_log(512); g.set_x0(x0); g.set_candidate(None, None, False); g.solve()
o = ob.Oracle(fp); o.set_x0(x0[0]); o.solve()
lg, lo = g.iteration_log(0), o.iteration_log()
for i in range(0, 13):
    a, b = lg[i], lo[i]
    print(i, "gpu", a.accepted, a.steplength, f"{a.cost:.12f} d0 {a.d0:.6e} d1 {a.d1:.6e} xreg {a.xreg:g}", "| orc", b.accepted, b.steplength, f"{b.cost:.12f} d0 {b.d0:.6e} d1 {b.d1:.6e} xreg {b.xreg:g}")
