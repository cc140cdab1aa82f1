#!/usr/bin/env python3
"""The reference's examples/python/trajectory.py, statement for statement, on the eagle_mpc front-end of this repo
(eagle-mpc_b200/python/eagle_mpc).  The only differences: crocoddyl.CallbackVerbose -> eagle_mpc.CallbackVerbose (the
callbacks are replayed from the device-side iteration log) and no Gepetto display.  Needs a CUDA device."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "eagle-mpc_b200", "python"))
import eagle_mpc  # noqa: E402
from eagle_mpc.utils.path import EAGLE_MPC_YAML_DIR  # noqa: E402

dt = 20  # ms
useSquash = True
robotName = 'hexacopter370_flying_arm_3'
trajectoryName = 'displacement'

trajectory = eagle_mpc.Trajectory()
trajectory.autoSetup(EAGLE_MPC_YAML_DIR + "/" + robotName + "/trajectories/" + trajectoryName + ".yaml")
problem = trajectory.createProblem(dt, useSquash, "IntegratedActionModelEuler")

solver = eagle_mpc.SolverSbFDDP(problem, trajectory.squash)

solver.setCallbacks([eagle_mpc.CallbackVerbose()])
solver.solve([], [], maxiter=100)
print("iterations", solver.iter, "cost", solver.cost, "final position", solver.xs[-1][:3])
