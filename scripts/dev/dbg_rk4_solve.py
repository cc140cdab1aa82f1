import sys, os, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_binding as ob
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads")
yaml, dt, seed0 = wl.CONFIGS["hexacopter370_flying_arm_3_displacement"]
fp = host.Trajectory(yaml).createProblem(dt, True, "IntegratedActionModelRK4")
x0 = fp.x0[None].copy()
g = capi.BatchSolver(fp, 1); g.enable_iteration_log(512); g.set_x0(x0); g.set_candidate(None, None, False); g.solve()
o = ob.Oracle(fp); o.set_x0(x0[0]); o.solve()
lg, lo = g.iteration_log(0), o.iteration_log()
print(len(lg), len(lo))
for i in range(0, 4):
    a, b = lg[i], lo[i]
    print(i, "gpu", a.accepted, a.steplength, f"{a.cost:.12f} d0 {a.d0:.6e} d1 {a.d1:.6e} xreg {a.xreg:g}", "| orc", b.accepted, b.steplength, f"{b.cost:.12f} d0 {b.d0:.6e} d1 {b.d1:.6e} xreg {b.xreg:g}")
from test_gpu_contact import rel
g = capi.BatchSolver(fp, 1); g.set_x0(x0); g.set_candidate(None, None, False)
o = ob.Oracle(fp); o.set_x0(x0[0]); o.set_candidate(None, None, False)
g.phase_calc_diff(0.1); o.phase_calc_diff(0.1)
tiles = g.tiles()[0]; ot = o.get("tiles"); off = fp.tile_offsets()
print("xnext", rel(g.xnext()[0][:-1], o.get("xnext")[:-1]), "cost", rel(g.node_cost()[0], o.get("node_cost")), "fs", rel(g.gaps()[0], o.get("fs")))
for name, size in (("Fx", fp.ndx * fp.ndx), ("Fu", fp.ndx * fp.nu), ("Lxx", fp.ndx * fp.ndx), ("Lxu", fp.ndx * fp.nu), ("Luu", fp.nu * fp.nu), ("Lx", fp.ndx), ("Lu", fp.nu)):
    a = tiles[:, off[name]:off[name] + size]; c = ot[:, off[name]:off[name] + size]
    err = np.abs(a - c).max(axis=1) / np.maximum(1.0, np.abs(c).max())
    print(name, "worst", err.max(), "bad nodes", np.nonzero(err > 1e-9)[0][:10].tolist())
ok = g.phase_backward(1e-9, False); ook = o.phase_backward(1e-9, False)
print("bw ok", ok, ook, "K", rel(g.K()[0], o.get("K")), "k", rel(g.k()[0], o.get("k")), "Vx", rel(g.Vx()[0], o.get("Vx")), "dgdq", g.dgdq()[0], o.get("dgdq"))
