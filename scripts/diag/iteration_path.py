"""Diagnostic: where does the GPU's iteration path leave the oracle's on a synthetic problem?  (GPU box only)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob
import synth
capi = importlib.import_module("eagle-mpc_b200.capi")
na, nr, T = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (5, 6, 25)
B = 6
h = synth.make_problem(seed=20 + na, na=na, n_rotors=nr, T=T, all_costs=False)
rng = np.random.default_rng(5)
x0 = np.zeros((B, h.nx)); x0[:, 6] = 1
x0[:, :3] = rng.uniform(-0.3, 0.3, size=(B, 3))
x0[:, 7:h.nq] = rng.uniform(-0.2, 0.2, size=(B, h.na))
g = capi.BatchSolver(h, B)
g.enable_iteration_log(1024)
g.set_x0(x0); g.set_candidate(None, None, False); g.solve()
print("gpu iters", g.iters().tolist())
for b in range(B):
    o = ob.Oracle(h); o.set_x0(x0[b]); o.solve()
    o2 = ob.Oracle(h, nofma=True); o2.set_x0(x0[b]); o2.solve()
    lg, lo, l2 = g.iteration_log(b), o.iteration_log(), o2.iteration_log()
    print(f"OCP {b}: gpu {len(lg)} orc {len(lo)} nofma {len(l2)}")
    shown = 0
    for i in range(min(len(lg), len(lo), len(l2))):
        a, r, r2 = lg[i], lo[i], l2[i]
        dg = abs(a.cost - r.cost) / max(1, abs(r.cost)); ds = abs(r2.cost - r.cost) / max(1, abs(r.cost))
        dec = (a.accepted, a.is_feasible, a.xreg) == (r.accepted, r.is_feasible, r.xreg)
        dec2 = (r2.accepted, r2.is_feasible, r2.xreg) == (r.accepted, r.is_feasible, r.xreg)
        if (not dec or not dec2 or dg > 1e-9 or ds > 1e-9 or i % 20 == 0) and shown < 40:
            shown += 1
            print(f"  it {i:3d} ph{r.phase} acc g/o/o2 {a.accepted}/{r.accepted}/{r2.accepted} xreg {a.xreg:.0e}/{r.xreg:.0e}/{r2.xreg:.0e} cost {r.cost:.12e} d_gpu {dg:.1e} d_self {ds:.1e} stop {a.stop:.3e}/{r.stop:.3e} d0 {a.d0:.6e}/{r.d0:.6e}")
