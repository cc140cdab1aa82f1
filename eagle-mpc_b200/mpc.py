"""Closed-loop MPC driver over the host-side controller mirrors (examples/python/mpc.py of the reference).

`CarrotMpc` / `RailMpc` / `WeightedMpc` wrap the C++ controllers (host/mpc.cpp): updateProblem(t) retargets the per-knot
costs, solve() runs the B200 SbFDDP path for the single MPC instance, and the RK4 plant runs as a CUDA kernel behind
`empc_plant_step`.
"""
import ctypes as C
import time

import numpy as np

from . import abi
from .capi import EmpcError, lib as _lib
from .host import hlib, _err


def _mlib():
    L = hlib()
    if not getattr(L, "_mpc_ready", False):
        L.empc_host_carrot_create.restype = C.c_void_p
        L.empc_host_carrot_create.argtypes = [C.c_void_p, abi.c_double_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
        L.empc_host_rail_create.restype = C.c_void_p
        L.empc_host_rail_create.argtypes = [abi.c_double_p, C.c_int32, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
        L.empc_host_weighted_create.restype = C.c_void_p
        L.empc_host_weighted_create.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32]
        L.empc_host_carrot_schedule.argtypes = [C.c_void_p, abi.c_int32_p, C.POINTER(C.c_int64), C.POINTER(C.c_uint8)]
        L.empc_host_weighted_schedule.argtypes = [C.c_void_p, abi.c_int32_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                                  C.POINTER(C.c_int64), abi.c_double_p, C.POINTER(C.c_uint8),
                                                  C.POINTER(C.c_uint8), abi.c_double_p]
        L.empc_host_carrot_free.argtypes = [C.c_void_p]
        L.empc_host_carrot_info.argtypes = [C.c_void_p, abi.c_int32_p]
        L.empc_host_carrot_desc.restype = C.POINTER(abi.ProblemDesc)
        L.empc_host_carrot_desc.argtypes = [C.c_void_p]
        L.empc_host_carrot_handle.restype = C.c_void_p
        L.empc_host_carrot_handle.argtypes = [C.c_void_p]
        L.empc_host_carrot_update.argtypes = [C.c_void_p, C.c_int32]
        L.empc_host_carrot_costs.argtypes = [C.c_void_p, C.POINTER(abi.Cost), abi.c_double_p]
        L.empc_host_carrot_solve.argtypes = [C.c_void_p, abi.c_double_p, abi.c_double_p, abi.c_double_p, C.c_int32, C.c_double]
        L.empc_host_carrot_result.argtypes = [C.c_void_p, abi.c_double_p, abi.c_double_p, abi.c_double_p, abi.c_double_p, abi.c_int32_p]
        L._mpc_ready = True
    return L


class _MpcBase(abi.DescView):
    """Shared surface of the controllers: updateProblem / solve / result / plant_step over an opaque C++ object."""

    def _adopt(self, p, create_solver):
        L = _mlib()
        if not p:
            raise EmpcError(_err())
        self._p = C.c_void_p(p)
        info = np.zeros(5, dtype=np.int32)
        L.empc_host_carrot_info(self._p, abi.as_int32_p(info))
        self.knots, self.dt, self.iters, self.n_costs, self.n_pool = (int(v) for v in info)
        self.desc = L.empc_host_carrot_desc(self._p).contents
        self.handle = C.c_void_p(L.empc_host_carrot_handle(self._p)) if create_solver else None

    def refresh_sizes(self):
        info = np.zeros(5, dtype=np.int32)
        _mlib().empc_host_carrot_info(self._p, abi.as_int32_p(info))
        self.n_costs, self.n_pool = int(info[3]), int(info[4])

    def updateProblem(self, t_ms):
        if _mlib().empc_host_carrot_update(self._p, int(t_ms)):
            raise EmpcError(_err())

    def cost_tables(self):
        costs = (abi.Cost * self.n_costs)()
        pool = np.zeros(self.n_pool)
        _mlib().empc_host_carrot_costs(self._p, costs, abi.as_double_p(pool))
        return costs, pool

    def solve(self, x0, xs=None, us=None, maxiter=100, convergence_init=1e-2):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        xs = None if xs is None else np.ascontiguousarray(xs, dtype=np.float64)
        us = None if us is None else np.ascontiguousarray(us, dtype=np.float64)
        rc = _mlib().empc_host_carrot_solve(self._p, abi.as_double_p(x0), None if xs is None else abi.as_double_p(xs),
                                            None if us is None else abi.as_double_p(us), int(maxiter), float(convergence_init))
        if rc:
            raise EmpcError(_err())

    def result(self):
        T = self.knots - 1
        xs = np.zeros((T + 1, self.nx)); us = np.zeros((T, self.nu)); uss = np.zeros((T, self.nu))
        cost = np.zeros(1); it = np.zeros(1, dtype=np.int32)
        _mlib().empc_host_carrot_result(self._p, abi.as_double_p(xs), abi.as_double_p(us), abi.as_double_p(uss),
                                        abi.as_double_p(cost), abi.as_int32_p(it))
        return xs, us, uss, cost[0], int(it[0])

    def plant_step(self, x, u, dt_s):
        x = np.ascontiguousarray(x, dtype=np.float64); u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.zeros_like(x)
        _lib().empc_plant_step.argtypes = [C.c_void_p, abi.c_double_p, abi.c_double_p, C.c_double, abi.c_double_p, C.c_int32]
        rc = _lib().empc_plant_step(self.handle, abi.as_double_p(x), abi.as_double_p(u), float(dt_s), abi.as_double_p(out), 1)
        if rc:
            raise EmpcError(_lib().empc_last_error().decode())
        return out

    def __del__(self):
        if getattr(self, "_p", None) and self._p.value:
            _mlib().empc_host_carrot_free(self._p)
            self._p = C.c_void_p()


class CarrotMpc(_MpcBase):
    """eagle_mpc.CarrotMpc(trajectory, state_ref, dt_ref, yaml_path)"""

    def __init__(self, trajectory, state_ref, dt_ref, yaml_path, create_solver=True):
        self._traj = trajectory
        ref = np.ascontiguousarray(state_ref, dtype=np.float64)
        self._adopt(_mlib().empc_host_carrot_create(trajectory._p, abi.as_double_p(ref), ref.shape[0], int(dt_ref),
                                                    yaml_path.encode(), int(create_solver)), create_solver)

    def schedule(self):
        """(t_stages [n_stages + 1] in ms, is_transition [n_stages]) for BatchSolver.set_carrot_schedule"""
        L = _mlib()
        n = np.zeros(1, dtype=np.int32)
        if L.empc_host_carrot_schedule(self._p, abi.as_int32_p(n), None, None):
            raise EmpcError(_err())
        t = np.zeros(int(n[0]) + 1, dtype=np.int64); tr = np.zeros(int(n[0]), dtype=np.uint8)
        if L.empc_host_carrot_schedule(self._p, abi.as_int32_p(n), t.ctypes.data_as(C.POINTER(C.c_int64)),
                                       tr.ctypes.data_as(C.POINTER(C.c_uint8))):
            raise EmpcError(_err())
        return t, tr


class RailMpc(_MpcBase):
    """eagle_mpc.RailMpc(state_ref, dt_ref, yaml_path)"""

    def __init__(self, state_ref, dt_ref, yaml_path, create_solver=True):
        ref = np.ascontiguousarray(state_ref, dtype=np.float64)
        self._adopt(_mlib().empc_host_rail_create(abi.as_double_p(ref), ref.shape[0], ref.shape[1], int(dt_ref),
                                                  yaml_path.encode(), int(create_solver)), create_solver)


class WeightedMpc(_MpcBase):
    """eagle_mpc.WeightedMpc(trajectory, dt_ref, yaml_path) — merges the trajectory's transition stages in place"""

    def __init__(self, trajectory, dt_ref, yaml_path, create_solver=True):
        self._traj = trajectory
        self._adopt(_mlib().empc_host_weighted_create(trajectory._p, int(dt_ref), yaml_path.encode(), int(create_solver)),
                    create_solver)

    def schedule(self):
        """The weight schedule in flat arrays (for BatchSolver.set_weighted_schedule): dict with t_ini, t_end (ms per
        stage), duration, alpha, beta and the n_stages x n_slots tables match / task / base."""
        L = _mlib()
        dims = np.zeros(2, dtype=np.int32)
        if L.empc_host_weighted_schedule(self._p, abi.as_int32_p(dims), None, None, None, None, None, None, None):
            raise EmpcError(_err())
        ns, nc = int(dims[0]), int(dims[1])
        t_ini = np.zeros(ns, dtype=np.int64); t_end = np.zeros(ns, dtype=np.int64); dur = C.c_int64()
        ab = np.zeros(2); match = np.zeros((ns, nc), dtype=np.uint8); task = np.zeros((ns, nc), dtype=np.uint8)
        base = np.zeros((ns, nc))
        i64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        u8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))
        if L.empc_host_weighted_schedule(self._p, abi.as_int32_p(dims), i64(t_ini), i64(t_end), C.byref(dur), abi.as_double_p(ab),
                                         u8(match), u8(task), abi.as_double_p(base)):
            raise EmpcError(_err())
        return {"t_ini": t_ini, "t_end": t_end, "duration": int(dur.value), "alpha": float(ab[0]), "beta": float(ab[1]),
                "match": match, "task": task, "base": base}


def closed_loop(mpc, xs_traj, us_traj, x_start, n_steps, dt_sim_ms=2, record=False):
    """examples/python/mpc.py:40-61: warm-up solve from the trajectory slice, then `n_steps` MPC steps against the RK4
    plant (dt_sim_ms).  Returns per-step latencies (s) of updateProblem+solve, and optionally the states / controls."""
    T = mpc.knots - 1
    mpc.updateProblem(0)
    mpc.solve(x_start, xs_traj[:T + 1], us_traj[:T], maxiter=100, convergence_init=1e-2)
    lat = []
    x = np.array(x_start, dtype=np.float64)
    states, controls, iters = [x.copy()], [], []
    t = 0
    for _ in range(n_steps):
        t0 = time.perf_counter()
        mpc.updateProblem(int(t))
        mpc.solve(x, None, None, maxiter=mpc.iters, convergence_init=1e-3)
        lat.append(time.perf_counter() - t0)
        _xs, _us, uss, _c, it = mpc.result()
        u = uss[0].copy()
        x = mpc.plant_step(x, u, dt_sim_ms / 1000.0)
        t += dt_sim_ms
        if record:
            states.append(x.copy()); controls.append(u); iters.append(it)
    return np.array(lat), np.array(states), np.array(controls), iters
