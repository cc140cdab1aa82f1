"""ctypes binding of the C ABI (include/empc_b200.h) exported by lib/libempc_b200.so.

There is no CPU fallback: if the shared library is missing or no CUDA device is usable, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EMPC_LIB", os.path.join(_HERE, "lib", "libempc_b200.so"))
_lib = None


class EmpcError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EmpcError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.empc_last_error.restype = C.c_char_p
        L.empc_create.argtypes = [C.POINTER(abi.ProblemDesc), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        for name in ("empc_destroy", "empc_solve", "empc_reset"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.empc_get_dims.argtypes = [C.c_void_p, C.POINTER(abi.Dims)]
        L.empc_set_x0.argtypes = [C.c_void_p, abi.c_double_p]
        L.empc_set_candidate.argtypes = [C.c_void_p, abi.c_double_p, abi.c_double_p, C.c_int32]
        L.empc_set_params.argtypes = [C.c_void_p, C.POINTER(abi.SolverParams)]
        L.empc_set_node_maps.argtypes = [C.c_void_p, abi.c_int32_p]
        L.empc_update_costs.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(abi.Cost), C.c_int32, C.c_int32,
                                        abi.c_double_p]
        L.empc_update_node_costsets.argtypes = [C.c_void_p, abi.c_int32_p]
        for name in ("xs", "us", "us_squash", "K", "k", "cost", "stop", "reg", "tiles", "xnext", "node_cost", "gaps",
                     "Vx", "Vxx_fs", "dgdq"):
            getattr(L, "empc_get_" + name).argtypes = [C.c_void_p, abi.c_double_p]
        for name in ("iters", "feasible"):
            getattr(L, "empc_get_" + name).argtypes = [C.c_void_p, abi.c_int32_p]
        L.empc_replicate_instances.argtypes = [C.c_void_p, C.c_int32]
        L.empc_set_reference_trajectory.argtypes = [C.c_void_p, abi.c_double_p, C.c_int32, C.c_int32]
        L.empc_rail_retarget.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int32]
        L.empc_plant_advance.argtypes = [C.c_void_p, C.c_double, abi.c_double_p, abi.c_double_p]
        L.empc_set_carrot_schedule.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_uint8)]
        L.empc_carrot_retarget.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int32]
        L.empc_set_weighted_schedule.argtypes = [C.c_void_p, C.POINTER(abi.WeightedSchedule)]
        L.empc_weighted_retarget.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int32]
        L.empc_get_cost_tables.argtypes = [C.c_void_p, C.POINTER(abi.Cost), abi.c_double_p]
        L.empc_get_solution.argtypes = [C.c_void_p, abi.c_double_p, abi.c_double_p, abi.c_double_p, abi.c_double_p,
                                        abi.c_double_p, abi.c_int32_p, abi.c_int32_p]
        L.empc_get_total_iterations.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.empc_get_launch_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), abi.c_double_p]
        L.empc_enable_kernel_timing.argtypes = [C.c_void_p, C.c_int32]
        L.empc_get_solve_stats.argtypes = [C.c_void_p, abi.c_double_p, C.POINTER(C.c_int64)]
        L.empc_phase_calc_diff.argtypes = [C.c_void_p, C.c_double]
        L.empc_phase_backward.argtypes = [C.c_void_p, C.c_double, C.c_int32, abi.c_int32_p]
        L.empc_phase_rollout.argtypes = [C.c_void_p, C.c_double, C.c_int32, C.c_int32]
        L.empc_solve_stream.argtypes = [C.c_void_p, C.c_int32, abi.c_double_p, abi.c_double_p, abi.c_double_p, abi.c_double_p,
                                        abi.c_double_p, abi.c_double_p, abi.c_int32_p, abi.c_int32_p]
        L.empc_enable_iteration_log.argtypes = [C.c_void_p, C.c_int32]
        L.empc_get_iteration_log.argtypes = [C.c_void_p, C.c_int32, C.POINTER(abi.IterRecord), C.c_int32, abi.c_int32_p]
        L.empc_get_trial.argtypes = [C.c_void_p, C.c_int32, abi.c_double_p, abi.c_double_p, abi.c_double_p,
                                     abi.c_double_p, abi.c_int32_p]
        _lib = L
    return _lib


def default_params():
    p = abi.SolverParams()
    lib().empc_default_params(C.byref(p))
    return p


def box_params(solver_type=abi.SOLVER_BOXFDDP):
    """Defaults of crocoddyl::SolverBoxFDDP / SolverBoxDDP (empc_box_params)."""
    p = abi.SolverParams()
    lib().empc_box_params(C.byref(p), solver_type)
    return p


def _ck(rc):
    if rc != 0:
        raise EmpcError(f"empc error {rc}: {lib().empc_last_error().decode()}")


class BatchSolver:
    """A batch of independent SbFDDP solves of one shooting problem on one GPU (handle of the C ABI)."""

    def __init__(self, holder, batch, device=0):
        self.holder = holder
        self.h = C.c_void_p()
        _ck(lib().empc_create(C.byref(holder.desc), batch, device, C.byref(self.h)))
        d = abi.Dims()
        _ck(lib().empc_get_dims(self.h, C.byref(d)))
        self.nq, self.nv, self.nx, self.ndx, self.nu, self.T, self.B, self.tile = (
            d.nq, d.nv, d.nx, d.ndx, d.nu, d.T, d.batch, d.tile)

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            lib().empc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs ----
    def set_x0(self, x0):
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(self.B, self.nx)
        _ck(lib().empc_set_x0(self.h, abi.as_double_p(x0)))

    def set_candidate(self, xs=None, us=None, feasible=False):
        xs = None if xs is None else np.ascontiguousarray(xs, dtype=np.float64).reshape(self.B, self.T + 1, self.nx)
        us = None if us is None else np.ascontiguousarray(us, dtype=np.float64).reshape(self.B, self.T, self.nu)
        _ck(lib().empc_set_candidate(self.h, None if xs is None else abi.as_double_p(xs),
                                     None if us is None else abi.as_double_p(us), int(feasible)))

    def set_params(self, p):
        _ck(lib().empc_set_params(self.h, C.byref(p)))

    def set_node_maps(self, m):
        m = np.ascontiguousarray(m, dtype=np.int32)
        _ck(lib().empc_set_node_maps(self.h, abi.as_int32_p(m)))

    def update_costs(self, first, costs, pool_off, pool):
        n = len(costs)
        arr = (abi.Cost * max(n, 1))(*costs)
        pool = np.ascontiguousarray(pool, dtype=np.float64)
        _ck(lib().empc_update_costs(self.h, first, n, arr, pool_off, len(pool), abi.as_double_p(pool)))

    def update_node_costsets(self, nc):
        nc = np.ascontiguousarray(nc, dtype=np.int32)
        _ck(lib().empc_update_node_costsets(self.h, abi.as_int32_p(nc)))

    # ---- batched MPC instances (device-side retargeting) ----
    def replicate_instances(self, n):
        """private cost tables for n controller instances; OCP b then belongs to instance b % n"""
        _ck(lib().empc_replicate_instances(self.h, int(n)))
        self.n_instances = int(n)

    def set_reference_trajectory(self, state_ref, dt_ref_ms):
        ref = np.ascontiguousarray(state_ref, dtype=np.float64).reshape(-1, self.nx)
        _ck(lib().empc_set_reference_trajectory(self.h, abi.as_double_p(ref), ref.shape[0], int(dt_ref_ms)))

    def rail_retarget(self, times_ms, dt_node_ms):
        """RailMpc.updateProblem(times_ms[m]) for every instance m, on the device"""
        t = np.ascontiguousarray(times_ms, dtype=np.int64)
        assert t.size == getattr(self, "n_instances", 1)
        _ck(lib().empc_rail_retarget(self.h, t.ctypes.data_as(C.POINTER(C.c_int64)), int(dt_node_ms)))

    def set_carrot_schedule(self, sch):
        """sch: CarrotMpc.schedule() = (t_stages [n_stages + 1], is_transition [n_stages])"""
        t = np.ascontiguousarray(sch[0], dtype=np.int64); tr = np.ascontiguousarray(sch[1], dtype=np.uint8)
        assert t.size == tr.size + 1
        _ck(lib().empc_set_carrot_schedule(self.h, tr.size, t.ctypes.data_as(C.POINTER(C.c_int64)),
                                           tr.ctypes.data_as(C.POINTER(C.c_uint8))))

    def carrot_retarget(self, times_ms, dt_node_ms):
        """CarrotMpc.updateProblem(times_ms[m]) for every instance m, on the device"""
        t = np.ascontiguousarray(times_ms, dtype=np.int64)
        assert t.size == getattr(self, "n_instances", 1)
        _ck(lib().empc_carrot_retarget(self.h, t.ctypes.data_as(C.POINTER(C.c_int64)), int(dt_node_ms)))

    def set_weighted_schedule(self, sch):
        """sch: WeightedMpc.schedule()"""
        t_ini = np.ascontiguousarray(sch["t_ini"], dtype=np.int64); t_end = np.ascontiguousarray(sch["t_end"], dtype=np.int64)
        match = np.ascontiguousarray(sch["match"], dtype=np.uint8); task = np.ascontiguousarray(sch["task"], dtype=np.uint8)
        base = np.ascontiguousarray(sch["base"], dtype=np.float64)
        s = abi.WeightedSchedule(match.shape[0], match.shape[1], t_ini.ctypes.data_as(C.POINTER(C.c_int64)),
                                 t_end.ctypes.data_as(C.POINTER(C.c_int64)), int(sch["duration"]), float(sch["alpha"]),
                                 float(sch["beta"]), match.ctypes.data_as(C.POINTER(C.c_uint8)),
                                 task.ctypes.data_as(C.POINTER(C.c_uint8)), abi.as_double_p(base))
        _ck(lib().empc_set_weighted_schedule(self.h, C.byref(s)))

    def weighted_retarget(self, times_ms, dt_node_ms):
        """WeightedMpc.updateProblem(times_ms[m]) for every instance m, on the device"""
        t = np.ascontiguousarray(times_ms, dtype=np.int64)
        assert t.size == getattr(self, "n_instances", 1)
        _ck(lib().empc_weighted_retarget(self.h, t.ctypes.data_as(C.POINTER(C.c_int64)), int(dt_node_ms)))

    def plant_advance(self, dt_s, fetch=True):
        """x0[b] <- RK4 plant(x0[b], us_squash[b][0], dt) on the device; returns (x_plant, u_applied) if fetch"""
        if not fetch:
            _ck(lib().empc_plant_advance(self.h, float(dt_s), None, None))
            return None
        x = np.zeros((self.B, self.nx)); u = np.zeros((self.B, self.nu))
        _ck(lib().empc_plant_advance(self.h, float(dt_s), abi.as_double_p(x), abi.as_double_p(u)))
        return x, u

    # ---- hot path ----
    def solve(self):
        _ck(lib().empc_solve(self.h))

    def solve_stream(self, x0_jobs, want_trajectories=True):
        """solve len(x0_jobs) independent OCPs through this handle's slots (refilled as OCPs finish); results per job"""
        x0 = np.ascontiguousarray(x0_jobs, dtype=np.float64).reshape(-1, self.nx)
        J = x0.shape[0]
        out = {"cost": np.zeros(J), "stop": np.zeros(J), "iters": np.zeros(J, dtype=np.int32), "feasible": np.zeros(J, dtype=np.int32)}
        if want_trajectories:
            out["xs"] = np.zeros((J, self.T + 1, self.nx)); out["us"] = np.zeros((J, self.T, self.nu)); out["us_squash"] = np.zeros((J, self.T, self.nu))
        p = lambda k: abi.as_double_p(out[k]) if k in out else None  # noqa: E731
        _ck(lib().empc_solve_stream(self.h, J, abi.as_double_p(x0), p("xs"), p("us"), p("us_squash"), p("cost"), p("stop"),
                                    abi.as_int32_p(out["iters"]), abi.as_int32_p(out["feasible"])))
        return out

    def reset(self):
        _ck(lib().empc_reset(self.h))

    def enable_kernel_timing(self, on=True):
        _ck(lib().empc_enable_kernel_timing(self.h, int(on)))

    # ---- outputs ----
    def _get(self, name, shape, dtype=np.float64):
        out = np.zeros(shape, dtype=dtype)
        ptr = abi.as_double_p(out) if dtype == np.float64 else abi.as_int32_p(out)
        _ck(getattr(lib(), "empc_get_" + name)(self.h, ptr))
        return out

    def xs(self): return self._get("xs", (self.B, self.T + 1, self.nx))
    def us(self): return self._get("us", (self.B, self.T, self.nu))
    def us_squash(self): return self._get("us_squash", (self.B, self.T, self.nu))
    def K(self): return self._get("K", (self.B, self.T, self.nu, self.ndx))
    def k(self): return self._get("k", (self.B, self.T, self.nu))
    def cost(self): return self._get("cost", (self.B,))
    def stop(self): return self._get("stop", (self.B,))
    def reg(self): return self._get("reg", (self.B,))
    def iters(self): return self._get("iters", (self.B,), np.int32)
    def feasible(self): return self._get("feasible", (self.B,), np.int32)
    def tiles(self): return self._get("tiles", (self.B, self.T + 1, self.tile))
    def xnext(self): return self._get("xnext", (self.B, self.T + 1, self.nx))
    def node_cost(self): return self._get("node_cost", (self.B, self.T + 1))
    def gaps(self): return self._get("gaps", (self.B, self.T + 1, self.ndx))
    def Vx(self): return self._get("Vx", (self.B, self.T + 1, self.ndx))
    def Vxx_fs(self): return self._get("Vxx_fs", (self.B, self.T + 1, self.ndx))
    def dgdq(self): return self._get("dgdq", (self.B, 2))

    def cost_tables(self, n_costs, n_pool):
        """the device's cost records and pool (sizes: those of the problem times the number of instances)"""
        costs = (abi.Cost * n_costs)(); pool = np.zeros(n_pool)
        _ck(lib().empc_get_cost_tables(self.h, costs, abi.as_double_p(pool)))
        return costs, pool

    def solution(self):
        """(xs, us, us_squash, cost, stop, iters, feasible) in one call (one packed copy for small batches)"""
        xs = np.zeros((self.B, self.T + 1, self.nx)); us = np.zeros((self.B, self.T, self.nu)); uss = np.zeros_like(us)
        cost = np.zeros(self.B); stop = np.zeros(self.B)
        it = np.zeros(self.B, dtype=np.int32); fe = np.zeros(self.B, dtype=np.int32)
        _ck(lib().empc_get_solution(self.h, abi.as_double_p(xs), abi.as_double_p(us), abi.as_double_p(uss), abi.as_double_p(cost),
                                    abi.as_double_p(stop), abi.as_int32_p(it), abi.as_int32_p(fe)))
        return xs, us, uss, cost, stop, it, fe

    def enable_iteration_log(self, capacity):
        """keep the last `capacity` iteration records per OCP (the stand-in for setCallbacks); 0 switches it off"""
        _ck(lib().empc_enable_iteration_log(self.h, int(capacity)))
        self._log_cap = int(capacity)

    def iteration_log(self, ocp):
        """records of OCP `ocp` from the last solve, oldest first (list of abi.IterRecord)"""
        cap = max(getattr(self, "_log_cap", 0), 1)
        rec = (abi.IterRecord * cap)()
        n = np.zeros(1, dtype=np.int32)
        _ck(lib().empc_get_iteration_log(self.h, int(ocp), rec, cap, abi.as_int32_p(n)))
        return [rec[i] for i in range(int(n[0]))]

    def total_iterations(self):
        v = C.c_int64()
        _ck(lib().empc_get_total_iterations(self.h, C.byref(v)))
        return v.value

    def launch_stats(self):
        n = C.c_int64()
        ms = np.zeros(4)
        _ck(lib().empc_get_launch_stats(self.h, C.byref(n), abi.as_double_p(ms)))
        return n.value, ms

    def solve_stats(self):
        """(device ms of the last solve, OCPs processed per kernel family summed over launches)"""
        ms = np.zeros(1)
        units = (C.c_int64 * 4)()
        _ck(lib().empc_get_solve_stats(self.h, abi.as_double_p(ms), units))
        return float(ms[0]), np.array(list(units), dtype=np.int64)

    def get_into(self, name, ptr):
        """empc_get_<name> into a caller-owned host buffer (e.g. pinned memory); ptr is an integer address"""
        fn = getattr(lib(), "empc_get_" + name)
        _ck(fn(self.h, C.cast(ptr, fn.argtypes[1])))

    def set_x0_ptr(self, ptr):
        """host pointer (e.g. pinned) to batch*nx doubles"""
        _ck(lib().empc_set_x0(self.h, C.cast(ptr, abi.c_double_p)))

    # ---- phase hooks ----
    def phase_calc_diff(self, smooth):
        _ck(lib().empc_phase_calc_diff(self.h, smooth))

    def phase_backward(self, xreg, feasible):
        ok = np.zeros(self.B, dtype=np.int32)
        _ck(lib().empc_phase_backward(self.h, xreg, int(feasible), abi.as_int32_p(ok)))
        return ok

    def phase_rollout(self, smooth, feasible, ddp):
        _ck(lib().empc_phase_rollout(self.h, smooth, int(feasible), int(ddp)))

    def trial(self, ai):
        xs = np.zeros((self.B, self.T + 1, self.nx)); us = np.zeros((self.B, self.T, self.nu))
        c = np.zeros(self.B); dv = np.zeros(self.B); ok = np.zeros(self.B, dtype=np.int32)
        _ck(lib().empc_get_trial(self.h, ai, abi.as_double_p(xs), abi.as_double_p(us), abi.as_double_p(c),
                                 abi.as_double_p(dv), abi.as_int32_p(ok)))
        return xs, us, c, dv, ok
