"""Diagnostic: clock64() cycles per phase of one node of backward_kernel (block 0), library built with -DEMPC_BW_PROFILE.
usage: EMPC_LIB=build_var/libempc_prof.so python scripts/diag/backward_phases.py"""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
host = importlib.import_module("eagle-mpc_b200.host"); capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads")
name = sys.argv[1] if len(sys.argv) > 1 else "hexacopter370_flying_arm_3_displacement"
yaml, dt, seed0 = wl.CONFIGS[name]
fp = host.Trajectory(yaml).createProblem(dt)
L = capi.lib()
names = ["loop-back + sync", "F^T[V|Vx] + stores (+ Lxx / cost-block loads issued)", "(F^T V) F + Qux/Quu stores", "Cholesky (+ F loads issued)",
         "gain solves", "Quuk, Vx, Qxu K, symmetrise, V store", "Vxx fs, Vx", "outputs (K, k, Vx, dots)"]
for k in (1, 4, 8):
    B = 148 * k
    g = capi.BatchSolver(fp, B)
    g.set_x0(wl.noisy_x0(fp.x0, B, seed0)); g.set_candidate(None, None, False)
    g.phase_calc_diff(0.1)
    g.phase_backward(1e-9, False)
    buf = (C.c_ulonglong * 16)()
    L.empc_debug_backward_profile(buf, 1)
    g.phase_backward(1e-9, False)
    L.empc_debug_backward_profile(buf, 0)
    v = np.array(list(buf), dtype=np.float64) / fp.T
    order = [8, 1, 2, 3, 4, 5, 6, 7]
    print(f"--- {k} warps per SM (B = {B}): {sum(v[i] for i in order):.0f} cycles per node")
    for nm, i in zip(names, order):
        print(f"   {nm:60s} {v[i]:8.0f}")
    g.close()
