#!/bin/bash
# compute-sanitizer memcheck over the overlay node models: a contact trajectory (eagle_catch: contact_node_kernel,
# node_dyn_contact, contact_force, backward_kernel<D, true>), an RK4 problem (rk4_node_kernel, node_dyn_rk4, rk4_node_cost) and
# a Box-solver solve (backward_kernel<D, true, true> with bw_box_qp, clamped overlay rollouts; racecheck too: lane 0 writes the
# shared-memory results the whole warp reads)
mkdir -p gpurun_out
cat > /tmp/san_overlay.py <<'PY'
import importlib, sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
for yaml, integ in (("hexacopter370_flying_arm_3/trajectories/eagle_catch.yaml", "IntegratedActionModelEuler"),
                    ("hexacopter370/trajectories/hover.yaml", "IntegratedActionModelRK4")):
    fp = host.Trajectory(yaml).createProblem(20, True, integ)
    B = 2
    g = capi.BatchSolver(fp, B)
    p = capi.default_params(); p.maxiter = 2
    g.set_params(p); g.set_x0(np.tile(fp.x0, (B, 1))); g.set_candidate(None, None, False); g.solve()
    g.phase_calc_diff(0.1); g.phase_backward(1e-6, False); g.phase_rollout(0.1, False, False)
    print(yaml, integ, "iters", g.iters().tolist())
abi = importlib.import_module("eagle-mpc_b200.abi")
fp = host.Trajectory("iris/trajectories/loop.yaml").createProblem(20, False, "IntegratedActionModelEuler")
g = capi.BatchSolver(fp, 2)
p = capi.box_params(abi.SOLVER_BOXFDDP); p.maxiter = 6
g.set_params(p); g.set_x0(np.tile(fp.x0, (2, 1))); g.set_candidate(None, None, False); g.solve()
us = g.us(); lb = np.array(fp.desc.u_lb[:fp.nu]); ub = np.array(fp.desc.u_ub[:fp.nu])
print("box solve iters", g.iters().tolist(), "controls on a limit", int(((us == lb) | (us == ub)).sum()))
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python /tmp/san_overlay.py 2>&1 | tail -8
echo "memcheck rc=$?"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python /tmp/san_overlay.py 2>&1 | tail -30
echo "racecheck rc=$?"
