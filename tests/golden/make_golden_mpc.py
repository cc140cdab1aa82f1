#!/usr/bin/env python3
"""Generates tests/golden/mpc_cost_tables.json: the per-knot cost records (active flag, weight, checksum of the reference)
that the host mirrors of {Carrot,Rail,Weighted}Mpc::updateProblem write at a few controller times.  The reference
trajectories come from the CPU oracle (maxiter 400, examples/python/mpc.py:29).  Pins the retargeting rules (integer-division
reference index, transition stages, tails, exp(alpha dt) weights) against regressions:  python tests/golden/make_golden_mpc.py
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
mpcmod = importlib.import_module("eagle-mpc_b200.mpc")

CASES = {
    "carrot": ("hexacopter370_flying_arm_3/trajectories/displacement.yaml", "hexacopter370_flying_arm_3/mpc/mpc.yaml"),
    "rail": ("iris_px4/trajectories/displacement.yaml", "iris_px4/mpc/mpc.yaml"),
    "weighted": ("iris_px4/trajectories/displacement.yaml", "iris_px4/mpc/mpc.yaml"),
}
TIMES = [0, 1500, 1990, 7900, 9000]


def reference_trajectory(traj_yaml):
    tr = host.Trajectory(traj_yaml)
    fp = tr.createProblem(20)
    p = ob.default_params(); p.maxiter = 400
    o = ob.Oracle(fp); o.set_params(p); o.set_x0(fp.x0); o.solve()
    return tr, o.get("xs")


def make_controller(kind, tr, xs, traj_yaml, mpc_yaml):
    if kind == "carrot":
        return mpcmod.CarrotMpc(tr, xs, 20, mpc_yaml, create_solver=False)
    if kind == "rail":
        return mpcmod.RailMpc(xs, 20, mpc_yaml, create_solver=False)
    return mpcmod.WeightedMpc(host.Trajectory(traj_yaml), 20, mpc_yaml, create_solver=False)


def tables(mpc):
    costs, pool = mpc.cost_tables()
    out = []
    for c in costs:
        ref = pool[c.ref_off:c.ref_off + mpc.nx] if (c.ref_off >= 0 and c.type == 0) else np.zeros(0)
        out.append([int(c.type), int(c.active), float(c.weight), float(np.dot(ref, np.arange(1, ref.size + 1)))])
    return out


def main():
    out = {}
    for kind, (traj_yaml, mpc_yaml) in CASES.items():
        tr, xs = reference_trajectory(traj_yaml)
        rec = {}
        for t in TIMES:
            mpc = make_controller(kind, tr if kind == "carrot" else None, xs, traj_yaml, mpc_yaml)  # fresh: no history
            mpc.updateProblem(t)
            rec[str(t)] = tables(mpc)
        out[kind] = rec
    with open(os.path.join(HERE, "mpc_cost_tables.json"), "w") as f:
        json.dump(out, f)
    print("wrote", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
