#!/bin/bash
# compute-sanitizer passes over a small solve of every robot family (memcheck, racecheck, synccheck)
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import importlib, sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import synth
capi = importlib.import_module("eagle-mpc_b200.capi")
for na, nr in ((3, 6), (0, 4), (2, 6), (5, 6)):
    h = synth.make_problem(seed=20 + na, na=na, n_rotors=nr, T=12, all_costs=True)
    B = 5
    rng = np.random.default_rng(5)
    x0 = np.zeros((B, h.nx)); x0[:, 6] = 1
    x0[:, :3] = rng.uniform(-0.3, 0.3, size=(B, 3))
    g = capi.BatchSolver(h, B)
    p = capi.default_params(); p.maxiter = 4; p.stop_criteria = 1
    g.enable_iteration_log(16)
    g.set_params(p); g.set_x0(x0); g.set_candidate(None, None, False); g.solve()
    assert len(g.iteration_log(0)) > 0
    g.phase_calc_diff(0.1); g.phase_backward(1e-6, False); g.phase_rollout(0.1, False, False)
    print("ok", na, nr, g.iters().tolist())
    # batched-MPC entry points (memory safety only: made-up schedules on the synthetic problem)
    g.replicate_instances(B)
    ref = np.zeros((20, h.nx)); ref[:, 6] = 1; ref[:, 0] = np.linspace(0, 1, 20)
    g.set_reference_trajectory(ref, 20)
    times = np.array([0, 30, 150, 390, 5000])
    g.rail_retarget(times, 20)
    g.set_carrot_schedule((np.array([0, 100, 120, 400]), np.array([1, 0, 0], dtype=np.uint8)))
    g.carrot_retarget(times, 20)
    nslots = max(int(np.diff(np.ctypeslib.as_array(h.desc.costset_begin, shape=(h.desc.n_costsets + 1,))).max()), 1)
    sch = {"t_ini": np.array([0, 200]), "t_end": np.array([200, 400]), "duration": 400, "alpha": 2.0, "beta": 1.0,
           "match": np.ones((2, nslots), dtype=np.uint8), "task": np.tile(np.arange(nslots) % 2, (2, 1)).astype(np.uint8),
           "base": np.full((2, nslots), 0.5)}
    g.set_weighted_schedule(sch); g.weighted_retarget(times, 20)
    g.set_x0(x0); g.set_candidate(None, None, False); g.solve()
    g.plant_advance(0.002)
    print("ok batched-mpc entry points", na, nr, g.iters().tolist())
    g.close()
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|error|hazard" | head -12
done
