"""ctypes mirror of include/empc_b200.h (POD structs only).

Used by the Python front-end of the package and by the tests' oracle binding; contains no compute.
"""
import ctypes as C

import numpy as np

MAX_JOINTS = 8
MAX_FRAMES = 16
MAX_ROTORS = 8
MAX_NU = 16
N_ALPHAS = 10

COST_STATE, COST_CONTROL, COST_FRAME_PLACEMENT, COST_FRAME_ROTATION, COST_FRAME_VELOCITY, \
    COST_FRAME_TRANSLATION, COST_SQUASH_BARRIER, COST_CONTACT_FRICTION_CONE = range(8)
CONTACT_3D, CONTACT_6D = 1, 2
INTEGRATOR_EULER, INTEGRATOR_RK4 = 0, 1
ACT_QUAD, ACT_WEIGHTED_QUAD, ACT_QUAD_BARRIER, ACT_WEIGHTED_QUAD_BARRIER = range(4)

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class Robot(C.Structure):
    _fields_ = [
        ("n_joints", C.c_int32),
        ("n_frames", C.c_int32),
        ("parent", C.c_int32 * MAX_JOINTS),
        ("jplace_R", (C.c_double * 9) * MAX_JOINTS),
        ("jplace_p", (C.c_double * 3) * MAX_JOINTS),
        ("axis", (C.c_double * 3) * MAX_JOINTS),
        ("mass", C.c_double * MAX_JOINTS),
        ("com", (C.c_double * 3) * MAX_JOINTS),
        ("inertia", (C.c_double * 9) * MAX_JOINTS),
        ("gravity", C.c_double * 3),
        ("frame_joint", C.c_int32 * MAX_FRAMES),
        ("frame_R", (C.c_double * 9) * MAX_FRAMES),
        ("frame_p", (C.c_double * 3) * MAX_FRAMES),
    ]


class WeightedSchedule(C.Structure):
    """empc_weighted_schedule_t"""
    _fields_ = [
        ("n_stages", C.c_int32), ("n_slots", C.c_int32),
        ("t_ini", C.POINTER(C.c_int64)), ("t_end", C.POINTER(C.c_int64)),
        ("duration", C.c_int64),
        ("alpha", C.c_double), ("beta", C.c_double),
        ("match", C.POINTER(C.c_uint8)), ("task", C.POINTER(C.c_uint8)),
        ("base", C.POINTER(C.c_double)),
    ]


class Cost(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("activation", C.c_int32),
        ("frame", C.c_int32),
        ("active", C.c_int32),
        ("weight", C.c_double),
        ("ref_off", C.c_int32),
        ("w_off", C.c_int32),
        ("lb_off", C.c_int32),
        ("ub_off", C.c_int32),
    ]


class Contact(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("frame", C.c_int32),
        ("gains", C.c_double * 2),
        ("ref_p", C.c_double * 3),
        ("ref_R", C.c_double * 9),
    ]


class ProblemDesc(C.Structure):
    _fields_ = [
        ("robot", Robot),
        ("n_rotors", C.c_int32),
        ("use_squash", C.c_int32),
        ("tau_f", C.c_double * (6 * MAX_ROTORS)),
        ("u_lb", C.c_double * MAX_NU),
        ("u_ub", C.c_double * MAX_NU),
        ("dt", C.c_double),
        ("T", C.c_int32),
        ("n_costsets", C.c_int32),
        ("n_costs", C.c_int32),
        ("n_pool", C.c_int32),
        ("n_node_maps", C.c_int32),
        ("costset_begin", c_int32_p),
        ("costs", C.POINTER(Cost)),
        ("pool", c_double_p),
        ("node_costset", c_int32_p),
        ("n_contacts", C.c_int32),
        ("integrator", C.c_int32),
        ("contacts", C.POINTER(Contact)),
        ("costset_contact", c_int32_p),
    ]


class SolverParams(C.Structure):
    _fields_ = [
        ("maxiter", C.c_int32),
        ("stop_gap_norm", C.c_int32),
        ("squash_quirk", C.c_int32),
        ("stop_criteria", C.c_int32),
        ("stop_test", C.c_int32),
        ("solver_type", C.c_int32),
        ("convergence_init", C.c_double),
        ("convergence_stop", C.c_double),
        ("convergence_mult", C.c_double),
        ("smooth_init", C.c_double),
        ("smooth_mult", C.c_double),
        ("barrier_weight", C.c_double),
        ("reg_init", C.c_double),
        ("reg_min", C.c_double),
        ("reg_max", C.c_double),
        ("reg_factor", C.c_double),
        ("th_acceptstep", C.c_double),
        ("th_acceptnegstep", C.c_double),
        ("th_grad", C.c_double),
        ("th_gaptol", C.c_double),
        ("th_stepdec", C.c_double),
        ("th_stepinc", C.c_double),
        ("th_stop_gaps", C.c_double),
        ("th_stop", C.c_double),
        ("boxqp_th_acceptstep", C.c_double),
        ("boxqp_th_grad", C.c_double),
        ("boxqp_reg", C.c_double),
        ("boxqp_maxiter", C.c_int32),
        ("reserved", C.c_int32),
    ]


STOP_CRITERIA_COST_REDUCTION, STOP_CRITERIA_QU_NORM = 0, 1
STOP_TEST_GAPS, STOP_TEST_FEASIBLE = 0, 1
SOLVER_SBFDDP, SOLVER_BOXFDDP, SOLVER_BOXDDP = 0, 1, 2


class IterRecord(C.Structure):
    """empc_iter_record_t"""
    _fields_ = [("iter", C.c_int32), ("total_iter", C.c_int32), ("phase", C.c_int32), ("accepted", C.c_int32),
                ("is_feasible", C.c_int32), ("reserved", C.c_int32), ("cost", C.c_double), ("stop", C.c_double),
                ("steplength", C.c_double), ("xreg", C.c_double), ("d0", C.c_double), ("d1", C.c_double),
                ("smooth", C.c_double)]


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nq", "nv", "nx", "ndx", "nu", "T", "batch", "tile")]


def as_double_p(a):
    return a.ctypes.data_as(c_double_p)


def as_int32_p(a):
    return a.ctypes.data_as(c_int32_p)


class DescView:
    """Dimension helpers over a ProblemDesc (`self.desc`)."""

    @property
    def na(self):
        return self.desc.robot.n_joints - 1

    @property
    def nv(self):
        return 6 + self.na

    @property
    def nq(self):
        return 7 + self.na

    @property
    def nx(self):
        return self.nq + self.nv

    @property
    def ndx(self):
        return 2 * self.nv

    @property
    def nu(self):
        return self.desc.n_rotors + self.na

    @property
    def T(self):
        return self.desc.T

    @property
    def tile(self):
        ndx, nu = self.ndx, self.nu
        t = 2 * ndx * ndx + 2 * ndx * nu + nu * nu + ndx + nu
        return t + (t & 1)

    def tile_offsets(self):
        ndx, nu = self.ndx, self.nu
        o = {}
        o["Fx"] = 0
        o["Fu"] = o["Fx"] + ndx * ndx
        o["Lxx"] = o["Fu"] + ndx * nu
        o["Lxu"] = o["Lxx"] + ndx * ndx
        o["Luu"] = o["Lxu"] + ndx * nu
        o["Lx"] = o["Luu"] + nu * nu
        o["Lu"] = o["Lx"] + ndx
        return o


class DescHolder(DescView):
    """Owns the numpy arrays a ProblemDesc points to (keeps them alive)."""

    def __init__(self, desc, costset_begin, costs, pool, node_costset):
        self.desc = desc
        self.costset_begin = np.ascontiguousarray(costset_begin, dtype=np.int32)
        self.costs = costs  # ctypes array of Cost
        self.pool = np.ascontiguousarray(pool, dtype=np.float64)
        self.node_costset = np.ascontiguousarray(node_costset, dtype=np.int32)
        desc.costset_begin = as_int32_p(self.costset_begin)
        desc.costs = C.cast(self.costs, C.POINTER(Cost))
        desc.pool = as_double_p(self.pool)
        desc.node_costset = as_int32_p(self.node_costset)
        desc.n_costsets = len(self.costset_begin) - 1
        desc.n_costs = len(self.costs)
        desc.n_pool = len(self.pool)
