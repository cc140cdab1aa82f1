#!/usr/bin/env python3
"""Generates tests/golden/twin_box.npz: the gains of crocoddyl's box solvers (SolverBoxFDDP / SolverBoxDDP::computeGains: box QP
per node, feedback from the free block, k = -du) over the LAST nodes of a horizon, computed by the independent numpy twin
(oracle/twin.py: complex-step node blocks, dense Riccati recursion, its own projected-Newton box_qp) at seeded candidates
whose controls sit 2 % inside their limits.  The -m gpu test tests/test_gpu_box.py::test_box_sweep_equals_twin_golden compares
the CUDA sweep (backward_kernel<D, true, true>) against this file without running any oracle code.
Run from the repo root: python scripts/make_twin_box_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import twin  # noqa: E402

CASES = [("iris/trajectories/loop.yaml", 20), ("hexacopter370_flying_arm_3/trajectories/displacement.yaml", 20),
         ("hexacopter370/trajectories/passthrough.yaml", 20)]
TAIL, XREG, SMOOTH = 10, 1e-2, 0.1


def main():
    out = {"n_cases": np.array(len(CASES)), "tail": np.array(TAIL), "xreg": np.array(XREG), "smooth": np.array(SMOOTH)}
    for ci, (rel, dt) in enumerate(CASES):
        tw = twin.Problem(rel, os.path.join(ROOT, "yaml"), os.path.join(ROOT, "fixtures", "urdf"), dt, use_squash=False)
        rob, T = tw.rob, tw.T
        rng = np.random.default_rng(777 + ci)
        xs = np.zeros((T + 1, rob.nx))
        for t in range(T + 1):
            xs[t] = twin.integrate(rob, tw.x0.astype(float), rng.uniform(-0.1, 0.1, rob.ndx))
        us = np.where(rng.uniform(size=(T, tw.nu)) < 0.5, tw.u_lb + 0.02 * (tw.u_ub - tw.u_lb), tw.u_ub - 0.02 * (tw.u_ub - tw.u_lb))
        T0 = T - TAIL
        nodes = [tw.calc_diff(tw.node_stage[t], xs[t], us[t], SMOOTH) for t in range(T0, T)]
        term = tw.calc_diff(tw.node_stage[T], xs[T], None, SMOOTH, terminal=True)
        fs = [np.zeros(rob.ndx)] * (TAIL + 1)
        k_prev = [np.zeros(tw.nu)] * TAIL
        key = f"c{ci}"
        out[key + "_yaml"] = np.array(rel); out[key + "_dt"] = np.array(dt); out[key + "_xs"] = xs; out[key + "_us"] = us
        n_clamped = 0
        for sweep in range(2):   # cold (warm start k = 0), then warm-started by the first sweep's k
            K, k, Vx, Vxx, Qus = twin.riccati_sweep(nodes, term, fs, XREG, True, box=(us[T0:], tw.u_lb, tw.u_ub, k_prev))
            out[f"{key}_s{sweep}_K"] = np.array(K); out[f"{key}_s{sweep}_k"] = np.array(k); out[f"{key}_s{sweep}_Vx"] = np.array(Vx[:TAIL])
            n_clamped += int(sum((np.abs(Kt).max(axis=1) == 0).sum() for Kt in K))
            k_prev = k
        assert n_clamped > 0, "no control clamped: the fixture would not exercise the box QP"
        print(rel, "T", T, "tail", TAIL, "clamped rows over both sweeps", n_clamped)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "twin_box.npz"), **out)


if __name__ == "__main__":
    main()
