"""Diagnostic: device time of one backward_kernel launch vs the number of OCPs (= warps) in flight.  (GPU box only)
B = 148 k OCPs put k warps on every SM (one per sub-partition up to k = 4): the curve separates per-OCP latency from
throughput limits.  usage: backward_occupancy.py [workload]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
host = importlib.import_module("eagle-mpc_b200.host"); capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads")
name = sys.argv[1] if len(sys.argv) > 1 else "hexacopter370_flying_arm_3_displacement"
yaml, dt, seed0 = wl.CONFIGS[name]
fp = host.Trajectory(yaml).createProblem(dt)
for k in (1, 2, 4, 6, 8, 10, 12, 16, 24, 28):
    B = 148 * k
    g = capi.BatchSolver(fp, B)
    g.set_x0(wl.noisy_x0(fp.x0, B, seed0)); g.set_candidate(None, None, False)
    g.phase_calc_diff(0.1)
    g.phase_backward(1e-9, False)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); g.phase_backward(1e-9, False); ts.append(time.perf_counter() - t0)
    t = min(ts) * 1e3
    print(f"warps/SM {k:3d}  B {B:5d}  backward {t:8.3f} ms   per node per OCP-slot {t*1e-3*1.965e9/fp.T:9.0f} cycles   ms per 4096 OCPs {t*4096/B:7.2f}", flush=True)
    g.close()
