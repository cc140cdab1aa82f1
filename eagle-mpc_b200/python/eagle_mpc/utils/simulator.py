"""AerialSimulator (bindings/python/eagle_mpc/utils/simulator.py:7-29): RK4 plant with the plain multicopter actuation.
The reference integrates with pinocchio on the CPU; here a step is the RK4 plant kernel behind empc_plant_step, run on the
solver handle of the controller the robot model came from."""
import ctypes as C

import numpy as np


class AerialSimulator:
    def __init__(self, solver, dt_ms, x0):
        """solver: the eagle_mpc.SolverSbFDDP whose robot is simulated (the reference passes robot_model and
        platform_params; the plant kernel takes both from the solver's problem)"""
        import os
        self._lib = C.CDLL(os.environ["EMPC_LIB"])
        self._lib.empc_plant_step.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double,
                                              C.POINTER(C.c_double), C.c_int32]
        self._lib.empc_last_error.restype = C.c_char_p
        self._h = C.c_void_p(solver.handle)
        self.dt = dt_ms / 1000.0
        self.states = [np.array(x0, dtype=np.float64)]
        self.controls = []

    def simulateStep(self, control):
        x = np.ascontiguousarray(self.states[-1]); u = np.ascontiguousarray(control, dtype=np.float64)
        out = np.zeros_like(x)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        if self._lib.empc_plant_step(self._h, dp(x), dp(u), self.dt, dp(out), 1):
            raise RuntimeError(self._lib.empc_last_error().decode())
        self.controls.append(u.copy()); self.states.append(out)
        return out
