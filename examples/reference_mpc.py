#!/usr/bin/env python3
"""The reference's examples/python/mpc.py on the eagle_mpc front-end of this repo: trajectory -> carrot / rail / weighted MPC
controller -> closed loop with the RK4 plant (2 ms), per-step updateProblem + solve timings.  Needs a CUDA device."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "eagle-mpc_b200", "python"))
import eagle_mpc  # noqa: E402
from eagle_mpc.utils.path import EAGLE_MPC_YAML_DIR  # noqa: E402
from eagle_mpc.utils.simulator import AerialSimulator  # noqa: E402

dt = 20  # ms
useSquash = True
robotName = 'hexacopter370_flying_arm_3'
trajectoryName = 'displacement'
mpcName = sys.argv[1] if len(sys.argv) > 1 else 'carrot'
nSteps = int(sys.argv[2]) if len(sys.argv) > 2 else None

trajectory = eagle_mpc.Trajectory()
trajectory.autoSetup(EAGLE_MPC_YAML_DIR + "/" + robotName + "/trajectories/" + trajectoryName + ".yaml")
problem = trajectory.createProblem(dt, useSquash, "IntegratedActionModelEuler")
solver = eagle_mpc.SolverSbFDDP(problem, trajectory.squash)
solver.solve([], [], maxiter=400)

mpcPath = EAGLE_MPC_YAML_DIR + "/" + robotName + "/mpc/mpc.yaml"
if mpcName == 'rail':
    mpcController = eagle_mpc.RailMpc(solver.xs, dt, mpcPath)
elif mpcName == 'weighted':
    mpcController = eagle_mpc.WeightedMpc(trajectory, dt, mpcPath)
else:
    mpcController = eagle_mpc.CarrotMpc(trajectory, solver.xs, dt, mpcPath)

mpcController.updateProblem(0)
mpcController.solver.solve(solver.xs[:mpcController.problem.T + 1], solver.us[:mpcController.problem.T])
mpcController.solver.convergence_init = 1e-3

dtSimulator = 2
simulator = AerialSimulator(mpcController.solver, dtSimulator, solver.xs[0])
t = 0
updateTime, solveTime = [], []
for i in range(0, nSteps if nSteps is not None else int(problem.T * dt * 1.2)):
    mpcController.problem.x0 = simulator.states[-1]
    start = time.time()
    mpcController.updateProblem(int(t))
    updateTime.append(time.time() - start)
    start = time.time()
    mpcController.solver.solve(mpcController.solver.xs, mpcController.solver.us, mpcController.iters)
    solveTime.append(time.time() - start)
    control = np.copy(mpcController.solver.us_squash[0])
    simulator.simulateStep(control)
    t += dtSimulator
print(f"{mpcName}: {len(solveTime)} steps, updateProblem p50 {1e3 * np.median(updateTime):.3f} ms, solve p50 {1e3 * np.median(solveTime):.3f} ms, "
      f"final position {np.round(simulator.states[-1][:3], 4)}")
