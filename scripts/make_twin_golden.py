#!/usr/bin/env python3
"""Generates tests/golden/twin_nodes.npz: node blocks (xnext, cost, Fx, Fu, Lx, Lu, Lxx, Luu) computed by the independent
numpy twin (oracle/twin.py: PyYAML / xml.etree front-end, RNEA-based forward dynamics, complex-step derivatives) at
seeded random candidate trajectories of the five robot families.  The -m gpu test test_gpu_twin_golden.py compares the CUDA
kernels' tiles against this file without running any oracle code.  Run from the repo root: python scripts/make_twin_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import twin  # noqa: E402

CASES = [("hexacopter370/trajectories/passthrough.yaml", 20), ("hexacopter370_flying_arm_3/trajectories/displacement.yaml", 20),
         ("hextilt_flying_arm_5/trajectories/push_slide.yaml", 20), ("iris_px4/trajectories/displacement.yaml", 20),
         ("hexacopter680_flying_arm_2/trajectories/hover.yaml", 20)]
SMOOTH = 0.07


def main():
    out = {}
    for ci, (rel, dt) in enumerate(CASES):
        tw = twin.Problem(rel, os.path.join(ROOT, "yaml"), os.path.join(ROOT, "fixtures", "urdf"), dt)
        rob, T = tw.rob, tw.T
        rng = np.random.default_rng(4242 + ci)
        xs = np.zeros((T + 1, rob.nx)); us = np.zeros((T, tw.nu))
        for t in range(T + 1):
            dx = np.concatenate([rng.uniform(-0.5, 0.5, 3), rng.uniform(-0.7, 0.7, 3), rng.uniform(-0.6, 0.6, rob.na), rng.uniform(-1, 1, rob.nv)])
            xs[t] = twin.integrate(rob, tw.x0.astype(float), dx)
        span = tw.u_ub - tw.u_lb
        us[:] = rng.uniform(tw.u_lb - 0.2 * span, tw.u_ub + 0.2 * span, size=(T, tw.nu))   # partly outside the box: barrier active
        # nodes: the first node of every distinct stage (at most 5 stages, the richest cost sets first) + the terminal node
        first = {}
        for t, s in enumerate(tw.node_stage[:T]):
            first.setdefault(s, t)
        stages = sorted(first, key=lambda s: -len(tw.stages[s]["costs"]))[:5]
        nodes = sorted(first[s] for s in stages) + [T]
        key = f"c{ci}"
        out[key + "_yaml"] = np.array(rel); out[key + "_dt"] = np.array(dt); out[key + "_smooth"] = np.array(SMOOTH)
        out[key + "_xs"] = xs; out[key + "_us"] = us; out[key + "_nodes"] = np.array(nodes)
        for t in nodes:
            term = t == T
            b = tw.calc_diff(tw.node_stage[t], xs[t], None if term else us[t], SMOOTH, terminal=term)
            for name in (("cost", "Lx", "Lxx") if term else ("xnext", "cost", "Fx", "Fu", "Lx", "Lu", "Lxx", "Luu")):
                out[f"{key}_n{t}_{name}"] = np.asarray(b[name])
        print(rel, "T", T, "nodes", nodes)
    out["n_cases"] = np.array(len(CASES))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "twin_nodes.npz"), **out)


if __name__ == "__main__":
    main()
