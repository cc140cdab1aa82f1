import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, importlib
import synth
capi = importlib.import_module("eagle-mpc_b200.capi")
na,nr,T=0,4,30
B=6
h = synth.make_problem(seed=20 + na, na=na, n_rotors=nr, T=T, all_costs=False)
rng = np.random.default_rng(5)
x0 = np.zeros((B, h.nx)); x0[:, 6] = 1
x0[:, :3] = rng.uniform(-0.3, 0.3, size=(B, 3))
x0[:, 7:h.nq] = rng.uniform(-0.2, 0.2, size=(B, h.na))
g = capi.BatchSolver(h, B)
g.set_x0(x0); g.set_candidate(None, None, False)
g.solve()
print(g.iters())
