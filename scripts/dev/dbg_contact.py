import sys, os, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_binding as ob
from test_gpu_contact import random_candidate, CATCH, MONKEY
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
case = sys.argv[1] if len(sys.argv) > 1 else CATCH
fp = host.Trajectory(case).createProblem(20)
B = 2
x0, xs, us = random_candidate(fp, B, 7)
g = capi.BatchSolver(fp, B); g.set_x0(x0); g.set_candidate(xs, us, False)
o = ob.Oracle(fp); o.set_x0(x0[0]); o.set_candidate(xs[0], us[0], False)
g.phase_calc_diff(0.1); o.phase_calc_diff(0.1)
nc = g.node_cost()[0]; oc = o.get("node_cost")
d = fp.desc
cs = [d.node_costset[t] for t in range(fp.T + 1)]
cc = [d.costset_contact[c] for c in range(d.n_costsets)]
print("costset_contact", cc)
for t in range(fp.T + 1):
    if abs(nc[t] - oc[t]) > 1e-9 * max(1, abs(oc[t])) or cc[cs[t]] >= 0:
        print(t, cs[t], cc[cs[t]], nc[t], oc[t])
tiles = g.tiles()[0]; ot = o.get("tiles"); off = fp.tile_offsets()
for name, size in (("Fx", fp.ndx * fp.ndx), ("Fu", fp.ndx * fp.nu), ("Lxx", fp.ndx * fp.ndx), ("Lxu", fp.ndx * fp.nu), ("Luu", fp.nu * fp.nu), ("Lx", fp.ndx), ("Lu", fp.nu)):
    a = tiles[:, off[name]:off[name] + size]; c = ot[:, off[name]:off[name] + size]
    err = np.abs(a - c).max(axis=1) / np.maximum(1.0, np.abs(c).max())
    bad = np.nonzero(err > 1e-9)[0]
    print(name, "worst", err.max(), "bad nodes", bad[:20].tolist())
print("gpu NaN nodes:", np.nonzero(np.isnan(tiles).any(axis=1))[0].tolist(), " oracle NaN nodes:", np.nonzero(np.isnan(ot).any(axis=1))[0].tolist())
for t in (71, 72):
    for name, size in (("Fx", fp.ndx * fp.ndx), ("Fu", fp.ndx * fp.nu), ("Lxx", fp.ndx * fp.ndx), ("Lxu", fp.ndx * fp.nu), ("Luu", fp.nu * fp.nu), ("Lx", fp.ndx), ("Lu", fp.nu)):
        a = tiles[t, off[name]:off[name] + size]; c = ot[t, off[name]:off[name] + size]
        print(t, name, "gpu nan count", int(np.isnan(a).sum()), "maxdiff", np.nanmax(np.abs(a - c)), "max", np.abs(c).max())
    a = tiles[t, off["Fx"]:off["Fx"] + fp.ndx * fp.ndx].reshape(fp.ndx, fp.ndx); c = ot[t, off["Fx"]:off["Fx"] + fp.ndx * fp.ndx].reshape(fp.ndx, fp.ndx)
    np.set_printoptions(precision=3, linewidth=250, suppress=True)
    print((np.abs(a - c) > 1e-8).astype(int))
xn = g.xnext()[0]; oxn = o.get("xnext")
print("xnext err contact nodes", [float(np.abs(xn[t] - oxn[t]).max()) for t in range(70, 81)])
