"""ctypes binding of oracle/liboracle.so — the CPU checker.  Test infrastructure only."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
abi = importlib.import_module("eagle-mpc_b200.abi")
dp = abi.as_double_p


def _load(name="liboracle.so"):
    path = os.path.join(ROOT, "oracle", name)
    src = [os.path.join(ROOT, "oracle", f) for f in ("oracle.cpp", "oracle_model.hpp", "oracle_math.hpp")]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(path)
    lib.orc_create.restype = C.c_void_p
    lib.orc_create.argtypes = [C.POINTER(abi.ProblemDesc)]
    lib.orc_destroy.argtypes = [C.c_void_p]
    lib.orc_get.argtypes = [C.c_void_p, C.c_char_p, abi.c_double_p]
    lib.orc_phase_calc_diff.argtypes = [C.c_void_p, C.c_double]
    lib.orc_phase_backward.argtypes = [C.c_void_p, C.c_double, C.c_int]
    lib.orc_phase_rollout.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
    lib.orc_node_eval.argtypes = [C.c_void_p, C.c_int, C.c_double] + [abi.c_double_p] * 6
    lib.orc_solve_batch.restype = C.c_double
    lib.orc_solve_batch.argtypes = [C.POINTER(abi.ProblemDesc), C.POINTER(abi.SolverParams), abi.c_double_p, C.c_int, C.c_int,
                                    abi.c_int32_p, abi.c_double_p]
    lib.orc_get_iteration_log.argtypes = [C.c_void_p, C.POINTER(abi.IterRecord), C.c_int]
    return lib


lib = _load()
_lib_nofma = None


def lib_nofma():
    """The same oracle compiled with -ffp-contract=off (conditioning yardstick)."""
    global _lib_nofma
    if _lib_nofma is None:
        _lib_nofma = _load("liboracle_nofma.so")
    return _lib_nofma


def default_params():
    p = abi.SolverParams()
    lib.orc_default_params(C.byref(p))
    return p


def box_params(solver_type=abi.SOLVER_BOXFDDP):
    """crocoddyl::SolverBoxFDDP / SolverBoxDDP defaults (orc_box_params)"""
    p = abi.SolverParams()
    lib.orc_box_params(C.byref(p), solver_type)
    return p


def solve_batch(holder, x0, nthreads=1, params=None):
    """(seconds, iterations per OCP, final costs) of the CPU oracle over a batch of initial states."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = x0.shape[0]
    p = params or default_params()
    iters = np.zeros(n, dtype=np.int32); cost = np.zeros(n)
    sec = lib.orc_solve_batch(C.byref(holder.desc), C.byref(p), dp(x0), n, nthreads, abi.as_int32_p(iters), dp(cost))
    return sec, iters, cost


class Oracle:
    def __init__(self, holder, nofma=False):
        self.h = holder
        self.lib = lib_nofma() if nofma else lib
        self.p = C.c_void_p(self.lib.orc_create(C.byref(holder.desc)))
        d = (C.c_int32 * 7)()
        self.lib.orc_dims(self.p, d)
        self.nq, self.nv, self.nx, self.ndx, self.nu, self.T, self.tile = list(d)

    def __del__(self):
        if getattr(self, "p", None):
            self.lib.orc_destroy(self.p)
            self.p = None

    def set_params(self, p):
        self.lib.orc_set_params(self.p, C.byref(p))

    def set_x0(self, x0):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        self.lib.orc_set_x0(self.p, dp(x0))

    def set_candidate(self, xs=None, us=None, feasible=False):
        xs = None if xs is None else np.ascontiguousarray(xs, dtype=np.float64)
        us = None if us is None else np.ascontiguousarray(us, dtype=np.float64)
        self.lib.orc_set_candidate(self.p, None if xs is None else dp(xs), None if us is None else dp(us), int(feasible))

    def solve(self, xs=None, us=None, feasible=False):
        xs = None if xs is None else np.ascontiguousarray(xs, dtype=np.float64)
        us = None if us is None else np.ascontiguousarray(us, dtype=np.float64)
        self.lib.orc_solve(self.p, None if xs is None else dp(xs), None if us is None else dp(us), int(feasible))

    def get(self, name):
        T, nx, ndx, nu = self.T, self.nx, self.ndx, self.nu
        shapes = {
            "xs": (T + 1, nx), "us": (T, nu), "xs_try": (T + 1, nx), "us_try": (T, nu), "us_squash": (T, nu),
            "K": (T, nu, ndx), "k": (T, nu), "Qu": (T, nu), "qp_stats": (4,), "Vx": (T + 1, ndx), "Vxx": (T + 1, ndx, ndx), "fs": (T + 1, ndx),
            "tiles": (T + 1, self.tile), "xnext": (T + 1, nx), "node_cost": (T + 1,), "cost": (1,),
            "cost_try": (1,), "stop": (1,), "xreg": (1,), "dgdq": (2,), "dv": (1,), "iter": (1,), "feasible": (1,),
        }
        out = np.zeros(shapes[name])
        rc = self.lib.orc_get(self.p, name.encode(), dp(out))
        assert rc == 0, name
        return out if out.size > 1 else out.reshape(-1)[0]

    def iteration_log(self):
        """one record per iteration of the last solve (what CallbackVerbose would have seen)"""
        n = self.lib.orc_get_iteration_log(self.p, None, 0)
        rec = (abi.IterRecord * max(n, 1))()
        self.lib.orc_get_iteration_log(self.p, rec, n)
        return [rec[i] for i in range(n)]

    def phase_calc_diff(self, smooth):
        self.lib.orc_phase_calc_diff(self.p, smooth)

    def phase_backward(self, xreg, feasible):
        return self.lib.orc_phase_backward(self.p, xreg, int(feasible))

    def phase_rollout(self, smooth, feasible, ddp, alpha_index):
        return self.lib.orc_phase_rollout(self.p, smooth, int(feasible), int(ddp), alpha_index)

    def node_eval(self, costset, smooth, x, u, diff=True):
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
        xnext = np.zeros(self.nx)
        cost = np.zeros(1)
        s = np.zeros(self.nu)
        tile = np.zeros(self.tile) if diff else None
        self.lib.orc_node_eval(self.p, costset, smooth, dp(x), None if u is None else dp(u), dp(xnext), dp(cost), dp(s),
                          None if tile is None else dp(tile))
        return xnext, cost[0], s, tile

    def integrate(self, x, dx):
        x = np.ascontiguousarray(x, dtype=np.float64); dx = np.ascontiguousarray(dx, dtype=np.float64)
        out = np.zeros(self.nx)
        self.lib.orc_state_integrate(self.p, dp(x), dp(dx), dp(out))
        return out

    def diff(self, x0, x1):
        x0 = np.ascontiguousarray(x0, dtype=np.float64); x1 = np.ascontiguousarray(x1, dtype=np.float64)
        out = np.zeros(self.ndx)
        self.lib.orc_state_diff(self.p, dp(x0), dp(x1), dp(out))
        return out

    def aba(self, q, v, tau):
        q, v, tau = (np.ascontiguousarray(a, dtype=np.float64) for a in (q, v, tau))
        a = np.zeros(self.nv)
        self.lib.orc_aba(self.p, dp(q), dp(v), dp(tau), dp(a))
        return a

    def rnea(self, q, v, a):
        q, v, a = (np.ascontiguousarray(z, dtype=np.float64) for z in (q, v, a))
        tau = np.zeros(self.nv)
        self.lib.orc_rnea(self.p, dp(q), dp(v), dp(a), dp(tau))
        return tau

    def aba_derivatives(self, q, v, tau):
        q, v, tau = (np.ascontiguousarray(z, dtype=np.float64) for z in (q, v, tau))
        nv = self.nv
        a = np.zeros(nv); aq = np.zeros((nv, nv)); av = np.zeros((nv, nv)); Minv = np.zeros((nv, nv))
        self.lib.orc_aba_derivatives(self.p, dp(q), dp(v), dp(tau), dp(a), dp(aq), dp(av), dp(Minv))
        return a, aq, av, Minv


def exp6(nu):
    nu = np.ascontiguousarray(nu, dtype=np.float64); R = np.zeros((3, 3)); p = np.zeros(3)
    lib.orc_exp6(dp(nu), dp(R), dp(p)); return R, p


def log6(R, p):
    R = np.ascontiguousarray(R, dtype=np.float64); p = np.ascontiguousarray(p, dtype=np.float64); nu = np.zeros(6)
    lib.orc_log6(dp(R), dp(p), dp(nu)); return nu


def Jlog6(R, p):
    R = np.ascontiguousarray(R, dtype=np.float64); p = np.ascontiguousarray(p, dtype=np.float64); J = np.zeros((6, 6))
    lib.orc_Jlog6(dp(R), dp(p), dp(J)); return J


def Jexp6(nu):
    nu = np.ascontiguousarray(nu, dtype=np.float64); J = np.zeros((6, 6))
    lib.orc_Jexp6(dp(nu), dp(J)); return J


def exp3(w):
    w = np.ascontiguousarray(w, dtype=np.float64); R = np.zeros((3, 3)); lib.orc_exp3(dp(w), dp(R)); return R


def log3(R):
    R = np.ascontiguousarray(R, dtype=np.float64); w = np.zeros(3); lib.orc_log3(dp(R), dp(w)); return w


def Jlog3(R):
    R = np.ascontiguousarray(R, dtype=np.float64); J = np.zeros((3, 3)); lib.orc_Jlog3(dp(R), dp(J)); return J


def quat_to_R(q):
    q = np.ascontiguousarray(q, dtype=np.float64); R = np.zeros((3, 3)); lib.orc_quat_to_R(dp(q), dp(R)); return R


def R_to_quat(R):
    R = np.ascontiguousarray(R, dtype=np.float64); q = np.zeros(4); lib.orc_R_to_quat(dp(R), dp(q)); return q


def closed_loop_bar(got_u, got_st, ref, yard, floor_u=1e-9, floor_st=1e-9):
    """Closed-loop parity bar: the GPU loop (got_*) must stay within max(floor, 4 x the reference's own sensitivity) of the
    oracle loop `ref` = (states, controls); the sensitivity is measured as the distance between `ref` and `yard`, the same
    loop run by the oracle compiled without FMA contraction (the closed loop feeds every solve's rounding into the next
    plant state, so the bar has to be a measured one).  Returns the four distances (u gpu, u self, x gpu, x self)."""
    st_o, u_o = ref
    st_y, u_y = yard
    su, sx = max(1.0, np.abs(u_o).max()), max(1.0, np.abs(st_o).max())
    du_self, dx_self = np.abs(u_y - u_o).max() / su, np.abs(st_y - st_o).max() / sx
    du, dx = np.abs(got_u - u_o).max() / su, np.abs(got_st - st_o).max() / sx
    assert du <= max(floor_u, 4 * du_self), ("controls", du, du_self)
    assert dx <= max(floor_st, 4 * dx_self), ("states", dx, dx_self)
    return du, du_self, dx, dx_self


def oracle_closed_loop(mpc, xs_traj, us_traj, x_start, n_steps, dt_sim_ms=2, record=False, t_start=0, xs_warm=None, us_warm=None,
                       params=None, nofma=False):
    """CPU twin of eagle-mpc_b200.mpc.closed_loop: same host-side CarrotMpc retargeting (created without a solver),
    oracle solves and oracle RK4 plant."""
    import time
    T = mpc.knots - 1
    o = Oracle(mpc, nofma=nofma)   # nofma: the rounding yardstick of closed_loop_bar
    L = o.lib
    L.orc_update_costs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.Cost), C.c_int, C.c_int, abi.c_double_p]
    L.orc_plant_step.argtypes = [C.c_void_p, abi.c_double_p, abi.c_double_p, C.c_double, abi.c_double_p]

    def push():
        costs, pool = mpc.cost_tables()
        L.orc_update_costs(o.p, 0, len(costs), costs, 0, len(pool), dp(pool))

    p = default_params() if params is None else params   # params: e.g. box_params(...) for a controller built on a Box solver
    mpc.updateProblem(int(t_start)); push()
    p.maxiter = 100; p.convergence_init = 1e-2
    o.set_params(p); o.set_x0(x_start)
    o.solve(xs_traj[:T + 1] if xs_warm is None else xs_warm, us_traj[:T] if us_warm is None else us_warm)
    p.maxiter = mpc.iters; p.convergence_init = 1e-3
    o.set_params(p)
    x = np.array(x_start, dtype=np.float64)
    lat, states, controls, iters = [], [x.copy()], [], []
    t = int(t_start)
    for _ in range(n_steps):
        t0 = time.perf_counter()
        mpc.updateProblem(int(t)); push()
        o.set_x0(x)
        o.solve(o.get("xs"), o.get("us"))
        lat.append(time.perf_counter() - t0)
        u = o.get("us_squash")[0].copy()
        xn = np.zeros_like(x)
        L.orc_plant_step(o.p, dp(np.ascontiguousarray(x)), dp(np.ascontiguousarray(u)), dt_sim_ms / 1000.0, dp(xn))
        x = xn
        t += dt_sim_ms
        if record:
            states.append(x.copy()); controls.append(u); iters.append(int(o.get("iter")))
    return np.array(lat), np.array(states), np.array(controls), iters
