"""Python front-end of the host-side mirror (Trajectory / SolverSbFDDP), over the plain-C helpers in host_capi.cpp."""
import ctypes as C

import numpy as np

from . import abi
import os

from .capi import EmpcError

HOST_LIB_PATH = os.environ.get("EMPC_HOST_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libempc_host.so"))

_h = None


def hlib():
    global _h
    if _h is None:
        # the host-side mirror is its own shared object: no CUDA code, so YAML / problem construction (and the CPU arm of
        # bench.py) never loads the CUDA library; SolverSbFDDP binds libempc_b200.so at run time (host/cuda_abi.cpp)
        if not os.path.exists(HOST_LIB_PATH):
            raise EmpcError(f"{HOST_LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(HOST_LIB_PATH)
        L.empc_host_last_error.restype = C.c_char_p
        L.empc_host_parse_yaml.restype = C.c_void_p
        L.empc_host_parse_yaml.argtypes = [C.c_char_p]
        L.empc_host_free_str.argtypes = [C.c_void_p]
        L.empc_host_set_dirs.argtypes = [C.c_char_p, C.c_char_p]
        L.empc_host_trajectory_create.restype = C.c_void_p
        L.empc_host_trajectory_create.argtypes = [C.c_char_p]
        L.empc_host_trajectory_free.argtypes = [C.c_void_p]
        L.empc_host_trajectory_info.argtypes = [C.c_void_p, abi.c_int32_p]
        L.empc_host_trajectory_stage_names.restype = C.c_void_p
        L.empc_host_trajectory_stage_names.argtypes = [C.c_void_p]
        L.empc_host_trajectory_platform.argtypes = [C.c_void_p, abi.c_double_p, abi.c_double_p, abi.c_double_p]
        L.empc_host_flatten.restype = C.c_void_p
        L.empc_host_flatten.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
        L.empc_host_flat_free.argtypes = [C.c_void_p]
        L.empc_host_flat_desc.restype = C.POINTER(abi.ProblemDesc)
        L.empc_host_flat_desc.argtypes = [C.c_void_p]
        L.empc_host_flat_x0.argtypes = [C.c_void_p, abi.c_double_p]
        L.empc_host_flat_cost_names.restype = C.c_void_p
        L.empc_host_flat_cost_names.argtypes = [C.c_void_p, C.c_int32]
        L.empc_host_solver_create.restype = C.c_void_p
        L.empc_host_solver_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int32, C.c_int32]
        L.empc_host_box_solver_create.restype = C.c_void_p
        L.empc_host_box_solver_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int32, C.c_int32]
        L.empc_host_solver_free.argtypes = [C.c_void_p]
        L.empc_host_solver_handle.restype = C.c_void_p
        L.empc_host_solver_handle.argtypes = [C.c_void_p]
        L.empc_host_solver_set_convergence_init.argtypes = [C.c_void_p, C.c_double]
        L.empc_host_solver_solve.argtypes = [C.c_void_p, C.c_int32]
        L.empc_host_solver_solve_batch.argtypes = [C.c_void_p, abi.c_double_p, abi.c_double_p, abi.c_double_p, C.c_int32]
        L.empc_host_solver_result.argtypes = [C.c_void_p, abi.c_double_p, abi.c_double_p, abi.c_double_p, abi.c_double_p,
                                              abi.c_int32_p, abi.c_int32_p]
        _h = L
    return _h


def _err():
    return hlib().empc_host_last_error().decode()


def _take_str(p):
    if not p:
        raise EmpcError(_err())
    s = C.string_at(p).decode()
    hlib().empc_host_free_str(p)
    return s


def parse_yaml(path):
    """ParserYaml(path).get_params() as a dict (the reference's flat '/'-keyed string map)."""
    out = {}
    for line in _take_str(hlib().empc_host_parse_yaml(path.encode())).splitlines():
        k, _, v = line.partition("\t")
        out[k] = v
    return out


class FlatProblem(abi.DescView):
    """A flattened ShootingProblem: `.desc` is the POD description both the CUDA path and the oracle consume."""

    def __init__(self, ptr, owner):
        self._p = C.c_void_p(ptr)
        self._owner = owner
        self.desc = hlib().empc_host_flat_desc(self._p).contents
        x0 = np.zeros(self.nx)
        hlib().empc_host_flat_x0(self._p, abi.as_double_p(x0))
        self.x0 = x0

    def cost_names(self, costset):
        return _take_str(hlib().empc_host_flat_cost_names(self._p, costset)).splitlines()

    def __del__(self):
        try:
            if getattr(self, "_p", None) and self._p.value:
                hlib().empc_host_flat_free(self._p)
                self._p = C.c_void_p()
        except Exception:  # interpreter shutdown: the library handle is already gone
            pass


class Trajectory:
    """eagle_mpc.Trajectory: autoSetup(yaml) at construction; createProblem flattens for the kernels."""

    def __init__(self, yaml_path):
        p = hlib().empc_host_trajectory_create(yaml_path.encode())
        if not p:
            raise EmpcError(_err())
        self._p = C.c_void_p(p)
        info = np.zeros(6, dtype=np.int32)
        hlib().empc_host_trajectory_info(self._p, abi.as_int32_p(info))
        self.nq, self.nv, self.nu, self.n_stages, self.duration, self.n_rotors = (int(v) for v in info)
        self.tau_f = np.zeros((6, self.n_rotors)); self.u_lb = np.zeros(self.nu); self.u_ub = np.zeros(self.nu)
        hlib().empc_host_trajectory_platform(self._p, abi.as_double_p(self.tau_f), abi.as_double_p(self.u_lb),
                                             abi.as_double_p(self.u_ub))

    def stage_names(self):
        """current stages (WeightedMpc merges the transition stages of the trajectory it is given, in place)"""
        return _take_str(hlib().empc_host_trajectory_stage_names(self._p)).split()

    def createProblem(self, dt, squash=True, integrator="IntegratedActionModelEuler", add_barrier=None):
        """add_barrier: SolverSbFDDP::barrierInit on the problem before flattening (default: exactly when squash is set — the
        box solvers take the problem as createProblem leaves it).  The "barrier" cost lands in the stages' shared cost sums
        and stays there, as it does in the reference (src/sbfddp.cpp:181-186 on the models of src/trajectory.cpp:102-143)."""
        if add_barrier is None:
            add_barrier = bool(squash)
        p = hlib().empc_host_flatten(self._p, int(dt), int(squash), integrator.encode(), int(add_barrier))
        if not p:
            raise EmpcError(_err())
        return FlatProblem(p, self)

    def __del__(self):
        try:
            if getattr(self, "_p", None) and self._p.value:
                hlib().empc_host_trajectory_free(self._p)
                self._p = C.c_void_p()
        except Exception:  # interpreter shutdown: the library handle is already gone
            pass


class SolverSbFDDP:
    """eagle_mpc.SolverSbFDDP facade (single OCP when batch=1), CUDA behind it."""

    def __init__(self, trajectory, dt, squash=True, integrator="IntegratedActionModelEuler", batch=1, device=0):
        p = hlib().empc_host_solver_create(trajectory._p, int(dt), int(squash), integrator.encode(), batch, device)
        if not p:
            raise EmpcError(_err())
        self._p = C.c_void_p(p)
        self._traj = trajectory
        self.batch = batch

    def set_convergence_init(self, c):
        hlib().empc_host_solver_set_convergence_init(self._p, c)

    @property
    def handle(self):
        """the empc_solver_t* behind the facade (for the C-ABI getters of capi.py)"""
        return C.c_void_p(hlib().empc_host_solver_handle(self._p))

    def solve(self, maxiter=100):
        if hlib().empc_host_solver_solve(self._p, maxiter):
            raise EmpcError(_err())

    def result(self, T):
        nx, nu = self._traj.nq + self._traj.nv, self._traj.nu
        xs = np.zeros((T + 1, nx)); us = np.zeros((T, nu)); uss = np.zeros((T, nu))
        cost = np.zeros(1); it = np.zeros(1, dtype=np.int32); fe = np.zeros(1, dtype=np.int32)
        hlib().empc_host_solver_result(self._p, abi.as_double_p(xs), abi.as_double_p(us), abi.as_double_p(uss),
                                       abi.as_double_p(cost), abi.as_int32_p(it), abi.as_int32_p(fe))
        return xs, us, uss, cost[0], int(it[0]), bool(fe[0])

    def __del__(self):
        try:
            if getattr(self, "_p", None) and self._p.value:
                hlib().empc_host_solver_free(self._p)
                self._p = C.c_void_p()
        except Exception:  # interpreter shutdown: the library handle is already gone
            pass


class SolverBoxFDDP(SolverSbFDDP):
    """crocoddyl.SolverBoxFDDP on the problem created with squash = False (examples/python/trajectory.py:19-24), CUDA behind it."""
    SOLVER_TYPE = abi.SOLVER_BOXFDDP

    def __init__(self, trajectory, dt, integrator="IntegratedActionModelEuler", batch=1, device=0):
        p = hlib().empc_host_box_solver_create(trajectory._p, int(dt), self.SOLVER_TYPE, integrator.encode(), batch, device)
        if not p:
            raise EmpcError(_err())
        self._p = C.c_void_p(p)
        self._traj = trajectory
        self.batch = batch


class SolverBoxDDP(SolverBoxFDDP):
    """crocoddyl.SolverBoxDDP, same construction."""
    SOLVER_TYPE = abi.SOLVER_BOXDDP
