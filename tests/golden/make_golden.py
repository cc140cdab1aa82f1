#!/usr/bin/env python3
"""Generates tests/golden/*.json from the CPU oracle (oracle/liboracle.so) through the host-side factory.

The reference ships no golden vectors (parity unpinned), so these fixtures pin the *in-repo*
oracle's behaviour on the named YAML problems with the synthetic URDFs: iteration count, stopping state, final cost and
samples of xs / us / K.  Regenerate with:  python tests/golden/make_golden.py
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402

host = importlib.import_module("eagle-mpc_b200.host")
wl = importlib.import_module("eagle-mpc_b200.workloads")


def main():
    out = {}
    for name, (yaml, dt, seed0) in wl.CONFIGS.items():
        fp = host.Trajectory(yaml).createProblem(dt)
        rec = {"T": fp.T, "nx": fp.nx, "nu": fp.nu, "ocps": []}
        x0s = np.vstack([fp.x0[None, :], wl.noisy_x0(fp.x0, 2, seed0)])  # the YAML state + the first two noisy starts
        for x0 in x0s:
            o = ob.Oracle(fp)
            o.set_x0(x0)
            o.solve()
            xs, us, K = o.get("xs"), o.get("us"), o.get("K")
            ts = sorted(set([0, fp.T // 4, fp.T // 2, fp.T - 1]))
            rec["ocps"].append({
                "x0": x0.tolist(), "iter": int(o.get("iter")), "feasible": int(o.get("feasible")),
                "cost": float(o.get("cost")), "stop": float(o.get("stop")),
                "xs_T": xs[-1].tolist(), "xs_mid": xs[fp.T // 2].tolist(),
                "us_samples": {str(t): us[t].tolist() for t in ts},
                "K_fro": {str(t): float(np.linalg.norm(K[t])) for t in ts},
                "xs_sum": float(xs.sum()), "us_sum": float(us.sum()),
            })
        out[name] = rec
    with open(os.path.join(HERE, "oracle_named_problems.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out), "problems")


if __name__ == "__main__":
    main()
