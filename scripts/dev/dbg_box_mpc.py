import importlib, os, sys, pathlib, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob
from test_box_oracle import box_mpc_yaml
host = importlib.import_module("eagle-mpc_b200.host"); abi = importlib.import_module("eagle-mpc_b200.abi"); mpcmod = importlib.import_module("eagle-mpc_b200.mpc")
tr = host.Trajectory("hexacopter370_flying_arm_3/trajectories/displacement.yaml")
fp = tr.createProblem(20, False, "IntegratedActionModelEuler")
po = ob.box_params(1); po.maxiter = 100
o = ob.Oracle(fp); o.set_params(po); o.set_x0(fp.x0); o.solve()
xs, us = o.get("xs"), o.get("us")
yaml = box_mpc_yaml(pathlib.Path(tempfile.mkdtemp()))
n = 20
mg = mpcmod.CarrotMpc(tr, xs, 20, yaml, create_solver=True)
_, st_g, u_g, it_g = mpcmod.closed_loop(mg, xs, us, xs[0], n, record=True)
mo = mpcmod.CarrotMpc(tr, xs, 20, yaml, create_solver=False)
_, st_o, u_o, it_o = ob.oracle_closed_loop(mo, xs, us, xs[0], n, record=True, params=ob.box_params(1))
my = mpcmod.CarrotMpc(tr, xs, 20, yaml, create_solver=False)
_, st_y, u_y, it_y = ob.oracle_closed_loop(my, xs, us, xs[0], n, record=True, params=ob.box_params(1), nofma=True)
print("iters", it_g == it_o, it_g)
for k in range(n):
    print(k, "du gpu %.2e self %.2e   dx gpu %.2e self %.2e" % (np.abs(u_g[k] - u_o[k]).max(), np.abs(u_y[k] - u_o[k]).max(), np.abs(st_g[k + 1] - st_o[k + 1]).max(), np.abs(st_y[k + 1] - st_o[k + 1]).max()), "argmax", int(np.abs(u_g[k] - u_o[k]).argmax()))
