"""lockstep box-MPC loop: the device and the oracle solve the same step from the oracle's plant state; where do k / K differ?"""
import ctypes as C, importlib, os, sys, pathlib, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob
from test_box_oracle import box_mpc_yaml
host = importlib.import_module("eagle-mpc_b200.host"); abi = importlib.import_module("eagle-mpc_b200.abi"); mpcmod = importlib.import_module("eagle-mpc_b200.mpc")
capi = importlib.import_module("eagle-mpc_b200.capi")
L = capi.lib()
tr = host.Trajectory("hexacopter370_flying_arm_3/trajectories/displacement.yaml")
fp = tr.createProblem(20, False, "IntegratedActionModelEuler")
po = ob.box_params(1); po.maxiter = 100
o0 = ob.Oracle(fp); o0.set_params(po); o0.set_x0(fp.x0); o0.solve()
xs, us = o0.get("xs"), o0.get("us")
yaml = box_mpc_yaml(pathlib.Path(tempfile.mkdtemp()))
mg = mpcmod.CarrotMpc(tr, xs, 20, yaml, create_solver=True)
mo = mpcmod.CarrotMpc(tr, xs, 20, yaml, create_solver=False)
T = mg.knots - 1; nu = mg.nu; ndx = 18
o = ob.Oracle(mo)
ob.lib.orc_update_costs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.Cost), C.c_int, C.c_int, abi.c_double_p]
ob.lib.orc_plant_step.argtypes = [C.c_void_p, abi.c_double_p, abi.c_double_p, C.c_double, abi.c_double_p]
def push():
    costs, pool = mo.cost_tables(); ob.lib.orc_update_costs(o.p, 0, len(costs), costs, 0, len(pool), ob.dp(pool))
p = ob.box_params(1)
mg.updateProblem(0); mo.updateProblem(0); push()
p.maxiter = 100; o.set_params(p); o.set_x0(xs[0]); o.solve(xs[:T + 1], us[:T])
mg.solve(xs[0], xs[:T + 1], us[:T], maxiter=100, convergence_init=1e-2)
p.maxiter = mg.iters; o.set_params(p)
x = xs[0].copy(); t = 0
lb = np.array(fp.desc.u_lb[:nu]); ub = np.array(fp.desc.u_ub[:nu])
for step in range(int(os.environ.get("NSTEPS", "20"))):
    print("STEP", step, flush=True)
    mg.updateProblem(t); mo.updateProblem(t); push()
    o.set_x0(x); o.solve(o.get("xs"), o.get("us"))
    mg.solve(x, None, None, maxiter=mg.iters, convergence_init=1e-3)
    kg = np.zeros((T, nu)); Kg = np.zeros((T, nu, ndx))
    L.empc_get_k(mg.handle, abi.as_double_p(kg)); L.empc_get_K(mg.handle, abi.as_double_p(Kg))
    ko, Ko = o.get("k"), o.get("K")
    _xs, usg, uss, _c, itg = mg.result()
    d = np.abs(kg - ko).max(axis=1)
    tt = int(d.argmax())
    print(step, "iters", itg, int(o.get("iter")), "max|dk| %.2e at node %d" % (d.max(), tt), "max|dus| %.2e" % np.abs(usg - o.get("us")).max(),
          "zero K rows gpu/oracle:", int((np.abs(Kg).max(axis=2) == 0).sum()), int((np.abs(Ko).max(axis=2) == 0).sum()))
    if d.max() > 1e-10:
        print("   k gpu   ", kg[tt]); print("   k oracle", ko[tt]); print("   us      ", o.get("us")[tt]); print("   zero rows gpu", np.flatnonzero(np.abs(Kg[tt]).max(axis=1) == 0), "oracle", np.flatnonzero(np.abs(Ko[tt]).max(axis=1) == 0))
    u = o.get("us_squash")[0].copy(); xn = np.zeros_like(x)
    ob.lib.orc_plant_step(o.p, ob.dp(np.ascontiguousarray(x)), ob.dp(np.ascontiguousarray(u)), 0.002, ob.dp(xn)); x = xn; t += 2
