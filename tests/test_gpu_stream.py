"""empc_solve_stream: more OCPs than slots, slots refilled from a queue as OCPs finish.  Every job must come out
bit-identical to the same OCP solved in a plain batch (an OCP's arithmetic does not depend on its slot or its neighbours)."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
wl = importlib.import_module("eagle-mpc_b200.workloads")


@pytest.mark.parametrize("name,slots,jobs", [("hexacopter370_hover", 8, 37), ("hexacopter370_hover", 16, 5),
                                             ("hextilt_flying_arm_5_push_slide", 4, 11)])
def test_stream_equals_plain_batch(name, slots, jobs):
    yaml, dt, seed0 = wl.CONFIGS[name]
    fp = host.Trajectory(yaml).createProblem(dt)
    x0 = wl.noisy_x0(fp.x0, jobs, seed0)
    ref = capi.BatchSolver(fp, jobs)
    ref.set_x0(x0); ref.set_candidate(None, None, False); ref.solve()
    g = capi.BatchSolver(fp, slots)
    out = g.solve_stream(x0)
    assert np.array_equal(out["iters"], ref.iters()) and np.array_equal(out["feasible"], ref.feasible())
    assert np.array_equal(out["cost"], ref.cost()) and np.array_equal(out["stop"], ref.stop())
    assert np.array_equal(out["xs"], ref.xs()) and np.array_equal(out["us"], ref.us()) and np.array_equal(out["us_squash"], ref.us_squash())
    assert g.total_iterations() == int((out["iters"] + 1).sum())
    if len(set(out["iters"].tolist())) > 1:
        print(name, "iterations per job", sorted(set(out["iters"].tolist()))[:6], "...")
    # the handle is reusable: a plain solve afterwards, and a second stream without trajectories
    g.set_x0(np.resize(x0, (slots, fp.nx))); g.set_candidate(None, None, False); g.solve()   # (np.resize repeats the rows)
    n = min(slots, jobs)
    assert np.array_equal(g.iters()[:n], ref.iters()[:n]) and np.array_equal(g.xs()[:n], ref.xs()[:n])
    out2 = g.solve_stream(x0, want_trajectories=False)
    assert np.array_equal(out2["iters"], out["iters"]) and np.array_equal(out2["cost"], out["cost"])
    assert g.solve_stream(np.zeros((0, fp.nx)))["iters"].size == 0
