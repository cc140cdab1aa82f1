"""Synthetic problem descriptions for unit tests (no YAML involved): random serial-chain UAMs with every cost type."""
import ctypes as C
import importlib

import numpy as np

abi = importlib.import_module("eagle-mpc_b200.abi")


def rand_rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def set_mat(dst, M):
    flat = np.asarray(M, dtype=float).reshape(-1)
    for i, v in enumerate(flat):
        dst[i] = float(v)


def make_robot(rng, na, n_frames=2, branching=False):
    r = abi.Robot()
    r.n_joints = 1 + na
    r.n_frames = n_frames
    for i in range(1 + na):
        r.parent[i] = -1 if i == 0 else (i - 1)
        if branching and i >= 3:
            r.parent[i] = int(rng.integers(0, i))
        R = np.eye(3) if i == 0 else rand_rot(rng)
        p = np.zeros(3) if i == 0 else rng.uniform(-0.2, 0.2, size=3)
        set_mat(r.jplace_R[i], R)
        set_mat(r.jplace_p[i], p)
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        set_mat(r.axis[i], ax)
        r.mass[i] = float(rng.uniform(0.2, 1.5) if i > 0 else rng.uniform(1.0, 3.0))
        set_mat(r.com[i], rng.uniform(-0.05, 0.05, size=3))
        A = rng.normal(size=(3, 3))
        I = (A @ A.T + 3 * np.eye(3)) * (0.002 if i > 0 else 0.02)
        set_mat(r.inertia[i], I)
    set_mat(r.gravity, [0, 0, -9.81])
    for f in range(n_frames):
        r.frame_joint[f] = 0 if f == 0 else na  # base link, and the last arm link
        set_mat(r.frame_R[f], np.eye(3) if f == 0 else rand_rot(rng))
        set_mat(r.frame_p[f], np.zeros(3) if f == 0 else rng.uniform(-0.1, 0.1, size=3))
    return r


class PoolBuilder:
    def __init__(self):
        self.data = []

    def add(self, v):
        off = len(self.data)
        self.data.extend(np.asarray(v, dtype=float).reshape(-1).tolist())
        return off


def make_problem(seed=0, na=3, n_rotors=6, T=12, dt=0.02, branching=False, all_costs=True):
    """Returns a DescHolder.  Cost set 0: running (with barrier), set 1: waypoint node (all frame costs, barrier),
    set 2: terminal (no barrier)."""
    rng = np.random.default_rng(seed)
    d = abi.ProblemDesc()
    d.robot = make_robot(rng, na, branching=branching)
    nv, nq = 6 + na, 7 + na
    nx, ndx, nu = nq + nv, 2 * nv, n_rotors + na
    d.n_rotors = n_rotors
    d.use_squash = 1
    # rotors on a circle, slightly tilted, alternating spin (same construction as multicopter-base-params.cpp:67-78)
    cf, cm = 4.1e-6, 7.0e-8
    tau_f = np.zeros((6, n_rotors))
    for i in range(n_rotors):
        ang = 2 * np.pi * i / n_rotors
        pos = np.array([0.2 * np.cos(ang), 0.2 * np.sin(ang), 0.01])
        tilt = rand_rot(rng) if False else np.eye(3)
        e3 = tilt @ np.array([0.05 * np.cos(ang + 1), 0.05 * np.sin(ang + 1), 1.0])
        e3 /= np.linalg.norm(e3)
        spin = -1 if i % 2 == 0 else 1
        tau_f[:3, i] = e3
        tau_f[3:, i] = np.cross(pos, e3) + spin * cm / cf * e3
    set_mat(d.tau_f, tau_f)
    for i in range(nu):
        d.u_lb[i] = 0.1 if i < n_rotors else -2.0
        d.u_ub[i] = 12.0 if i < n_rotors else 2.0
    d.dt = dt
    d.T = T
    d.n_node_maps = 1

    pool = PoolBuilder()
    costs = []

    def xref():
        x = np.zeros(nx)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        x[:3] = rng.uniform(-1, 1, size=3)
        x[3:7] = q
        x[7:nq] = rng.uniform(-0.5, 0.5, size=na)
        x[nq:] = rng.uniform(-0.3, 0.3, size=nv)
        return x

    def cost(type_, act, weight, frame=0, ref=None, w=None, lb=None, ub=None, active=1):
        c = abi.Cost()
        c.type, c.activation, c.frame, c.active, c.weight = type_, act, frame, active, weight
        c.ref_off = pool.add(ref) if ref is not None else -1
        c.w_off = pool.add(w) if w is not None else -1
        c.lb_off = pool.add(lb) if lb is not None else -1
        c.ub_off = pool.add(ub) if ub is not None else -1
        return c

    def se3ref():
        return np.concatenate([rand_rot(rng).reshape(-1), rng.uniform(-1, 1, size=3)])

    zero_x = np.zeros(nx); zero_x[6] = 1
    lim_w = np.zeros(ndx); lim_w[6:nv] = 1; lim_w[nv + 6:] = 1
    lim_ub = np.zeros(ndx); lim_ub[6:nv] = 0.3; lim_ub[nv + 6:] = 0.5
    # set 0: running — names sorted: barrier, limits_state, reg_control, reg_state
    set0 = [
        cost(abi.COST_SQUASH_BARRIER, abi.ACT_WEIGHTED_QUAD_BARRIER, 1e-3),
        cost(abi.COST_STATE, abi.ACT_WEIGHTED_QUAD_BARRIER, 100.0, ref=zero_x, w=lim_w, lb=-lim_ub, ub=lim_ub),
        cost(abi.COST_CONTROL, abi.ACT_WEIGHTED_QUAD, 1e-2, ref=np.zeros(nu), w=rng.uniform(0.5, 2, size=nu)),
        cost(abi.COST_STATE, abi.ACT_WEIGHTED_QUAD, 1e-1, ref=xref(), w=rng.uniform(0.5, 2, size=ndx)),
    ]
    set1 = [
        cost(abi.COST_SQUASH_BARRIER, abi.ACT_WEIGHTED_QUAD_BARRIER, 1e-3),
        cost(abi.COST_FRAME_VELOCITY, abi.ACT_QUAD, 10.0, frame=1, ref=rng.uniform(-0.2, 0.2, size=6)),
        cost(abi.COST_FRAME_ROTATION, abi.ACT_QUAD, 20.0, frame=0, ref=rand_rot(rng).reshape(-1)),
        cost(abi.COST_FRAME_PLACEMENT, abi.ACT_QUAD, 30.0, frame=1, ref=se3ref()),
        cost(abi.COST_FRAME_PLACEMENT, abi.ACT_WEIGHTED_QUAD, 15.0, frame=0, ref=se3ref(), w=rng.uniform(0.5, 2, size=6)),
        cost(abi.COST_FRAME_TRANSLATION, abi.ACT_QUAD, 25.0, frame=1, ref=rng.uniform(-1, 1, size=3)),
        cost(abi.COST_FRAME_VELOCITY, abi.ACT_QUAD_BARRIER, 5.0, frame=0, ref=np.zeros(6), lb=-0.05 * np.ones(6), ub=0.05 * np.ones(6)),
        cost(abi.COST_STATE, abi.ACT_QUAD, 0.5, ref=xref()),
        cost(abi.COST_CONTROL, abi.ACT_QUAD, 1e-2, ref=rng.uniform(0, 1, size=nu), active=0),
        cost(abi.COST_CONTROL, abi.ACT_QUAD, 2e-2, ref=rng.uniform(0, 1, size=nu)),
    ]
    set2 = [
        cost(abi.COST_FRAME_PLACEMENT, abi.ACT_QUAD, 300.0, frame=1, ref=se3ref()),
        cost(abi.COST_FRAME_VELOCITY, abi.ACT_QUAD, 100.0, frame=1, ref=np.zeros(6)),
        cost(abi.COST_CONTROL, abi.ACT_WEIGHTED_QUAD, 1e-2, ref=np.zeros(nu), w=np.ones(nu)),
        cost(abi.COST_STATE, abi.ACT_WEIGHTED_QUAD, 1e-1, ref=zero_x, w=np.ones(ndx)),
    ]
    if not all_costs:
        set1 = list(set0)
    sets = [set0, set1, set2]
    begin = np.cumsum([0] + [len(s) for s in sets]).astype(np.int32)
    flat = [c for s in sets for c in s]
    arr = (abi.Cost * len(flat))(*flat)
    node_costset = np.zeros(T + 1, dtype=np.int32)
    node_costset[T // 2] = 1
    node_costset[T] = 2
    return abi.DescHolder(d, begin, arr, np.array(pool.data), node_costset)


def random_state(rng, h, scale=1.0):
    x = np.zeros(h.nx)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    x[:3] = rng.uniform(-1, 1, size=3) * scale
    x[3:7] = q
    x[7:h.nq] = rng.uniform(-1, 1, size=h.na) * scale
    x[h.nq:] = rng.uniform(-1, 1, size=h.nv) * scale
    return x
