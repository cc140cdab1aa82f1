/*
 * empc_b200.h — C ABI of the B200-native SbFDDP hot path (drop-in for eagle-mpc's solver path).
 *
 * Everything that crosses this boundary is plain-old-data: fixed-size structs, pointers and sizes.
 * No C++ types, no torch types, no exceptions.  The host-side mirror of eagle-mpc's C++ surfaces
 * (Trajectory / Stage / factories / SolverSbFDDP / MPC controllers, see eagle-mpc_b200/host/) flattens a
 * problem into an `empc_problem_desc_t` and drives the CUDA kernels through the functions below.
 *
 * Reference interfaces each entry point replaces (paths relative to the eagle-mpc tree):
 *   empc_create            SolverSbFDDP::SolverSbFDDP(problem, squashing)        src/sbfddp.cpp:5-38 (+ barrierInit :169-190)
 *   empc_set_x0            crocoddyl::ShootingProblem::set_x0                     examples/python/mpc.py:50
 *   empc_set_candidate     crocoddyl::SolverAbstract::setCandidate                src/sbfddp.cpp:199
 *   empc_solve_stream      a queue of independent SolverSbFDDP::solve calls       src/sbfddp.cpp:192-226
 *   empc_set_params        set_convergence_init / solve(maxiter)                  include/eagle_mpc/sbfddp.hpp:43-52
 *   empc_update_costs      {Carrot,Rail,Weighted}Mpc::updateProblem               src/mpc-controllers/carrot-mpc.cpp:298-401
 *   empc_solve             SolverSbFDDP::solve                                    src/sbfddp.cpp:192-226
 *   empc_get_*             get_xs/get_us/get_K/get_k/get_cost/get_iter, getSquashControls  src/sbfddp.cpp:487
 *   empc_phase_*           computeDirection / tryStep (tile-level parity hooks)   src/sbfddp.cpp:244,264
 *
 * Conventions (Crocoddyl 1.x / Pinocchio 2.x, restated in DESIGN.md):
 *   x = [q ; v],  q = [p(3), quat xyzw(4), theta(na)],  v = [v_lin(3), omega(3) (body frame), theta_dot(na)]
 *   nq = 7+na, nv = 6+na, nx = nq+nv, ndx = 2 nv, nu = n_rotors + na.
 *   All matrices are row-major doubles.
 */
#ifndef EMPC_B200_H
#define EMPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMPC_MAX_JOINTS 8   /* free-flyer + up to 7 revolute joints */
#define EMPC_MAX_FRAMES 16
#define EMPC_MAX_ROTORS 8
#define EMPC_MAX_NU 16
#define EMPC_N_ALPHAS 10    /* line-search step lengths 2^0 .. 2^-9 (crocoddyl SolverDDP ctor) */

/* status codes */
enum {
  EMPC_OK = 0,
  EMPC_ERR_INVALID = 1,  /* bad argument / unsupported problem */
  EMPC_ERR_CUDA = 2,     /* CUDA runtime error (see empc_last_error) */
  EMPC_ERR_UNSUPPORTED = 3
};

/* cost (residual) types — src/factory/cost.cpp:38-168 */
enum {
  EMPC_COST_STATE = 0,
  EMPC_COST_CONTROL = 1,
  EMPC_COST_FRAME_PLACEMENT = 2,
  EMPC_COST_FRAME_ROTATION = 3,
  EMPC_COST_FRAME_VELOCITY = 4,
  EMPC_COST_FRAME_TRANSLATION = 5,
  EMPC_COST_SQUASH_BARRIER = 6, /* the "barrier" cost SolverSbFDDP adds, src/sbfddp.cpp:169-190 */
  EMPC_COST_CONTACT_FRICTION_CONE = 7 /* src/factory/cost.cpp:149-167: r = A f on the force of the stage's contact */
};

/* integrators — src/factory/int-action.cpp:24-35, selected by createProblem(dt, squash, integration_method)
 * (src/trajectory.cpp:102-104) or by mpc_controller/integration_method (src/mpc-base.cpp:42-43) */
enum { EMPC_INTEGRATOR_EULER = 0, EMPC_INTEGRATOR_RK4 = 1 };

/* contact models — src/factory/contacts.cpp:32-81 (DifferentialActionModelContactFwdDynamics, src/factory/diff-action.cpp:30-32) */
enum { EMPC_CONTACT_3D = 1, EMPC_CONTACT_6D = 2 };
typedef struct empc_contact {
  int32_t type;       /* EMPC_CONTACT_3D: the frame origin is held (3 constraint rows); _6D: the whole frame (6 rows) */
  int32_t frame;      /* index into empc_robot_t::frame_* */
  double gains[2];    /* Baumgarte gains; the corpus uses (0, 0) and the kernels require it */
  double ref_p[3];    /* reference position (used by the gains only) */
  double ref_R[9];    /* reference orientation, 6D (used by the gains only) */
} empc_contact_t;

/* activation types — src/factory/activation.cpp:34-101 */
enum {
  EMPC_ACT_QUAD = 0,
  EMPC_ACT_WEIGHTED_QUAD = 1,
  EMPC_ACT_QUAD_BARRIER = 2,
  EMPC_ACT_WEIGHTED_QUAD_BARRIER = 3
};

/* Rigid-body tree: joint 0 is the free-flyer (Pinocchio "root_joint"), joints 1.. are revolute. */
typedef struct empc_robot {
  int32_t n_joints;                       /* 1 + number of arm joints */
  int32_t n_frames;                       /* operational frames referenced by costs */
  int32_t parent[EMPC_MAX_JOINTS];        /* parent joint index, -1 for joint 0 */
  double jplace_R[EMPC_MAX_JOINTS][9];    /* joint placement in the parent joint frame (rotation) */
  double jplace_p[EMPC_MAX_JOINTS][3];    /* ... translation */
  double axis[EMPC_MAX_JOINTS][3];        /* revolute axis, unit, joint frame (unused for joint 0) */
  double mass[EMPC_MAX_JOINTS];           /* body inertia attached to joint i, expressed in joint frame */
  double com[EMPC_MAX_JOINTS][3];
  double inertia[EMPC_MAX_JOINTS][9];     /* rotational inertia about the COM, 3x3 */
  double gravity[3];                      /* world linear gravity, (0,0,-9.81) */
  int32_t frame_joint[EMPC_MAX_FRAMES];   /* joint each frame hangs on */
  double frame_R[EMPC_MAX_FRAMES][9];     /* frame placement in that joint's frame */
  double frame_p[EMPC_MAX_FRAMES][3];
} empc_robot_t;

/* One cost term of a CostModelSum.  Offsets index `pool` (doubles); -1 = absent.
 * Reference layouts: STATE nx | CONTROL nu | FRAME_PLACEMENT R(9)+p(3) | FRAME_ROTATION R(9) |
 * FRAME_VELOCITY lin(3)+ang(3) | FRAME_TRANSLATION p(3) | SQUASH_BARRIER none |
 * CONTACT_FRICTION_CONE A (5 x 3 row-major: crocoddyl::FrictionCone(n_surf, mu, 4, false)), lb / ub 5 each.
 * Activation vectors (weights / lower / upper bound) have the residual's dimension. */
typedef struct empc_cost {
  int32_t type;
  int32_t activation;
  int32_t frame;
  int32_t active;
  double weight;
  int32_t ref_off;
  int32_t w_off;
  int32_t lb_off;
  int32_t ub_off;
} empc_cost_t;

/* Flat shooting problem.  A "cost set" is one CostModelSum (costs already in the order the reference
 * iterates them: std::map<std::string,...> => sorted by name).  node_costset has n_node_maps*(T+1)
 * entries; entry [m*(T+1)+t] is the cost set of node t (t = T is the terminal node) under map m. */
typedef struct empc_problem_desc {
  empc_robot_t robot;
  int32_t n_rotors;
  int32_t use_squash;                       /* 1: ActuationSquashingModel (required by the solver) */
  double tau_f[6 * EMPC_MAX_ROTORS];        /* 6 x n_rotors, row-major (src/multicopter-base-params.cpp:67-78) */
  double u_lb[EMPC_MAX_NU];
  double u_ub[EMPC_MAX_NU];
  double dt;                                /* seconds */
  int32_t T;                                /* number of running nodes */
  int32_t n_costsets;
  int32_t n_costs;
  int32_t n_pool;
  int32_t n_node_maps;
  const int32_t* costset_begin;             /* n_costsets+1 */
  const empc_cost_t* costs;                 /* n_costs */
  const double* pool;                       /* n_pool */
  const int32_t* node_costset;              /* n_node_maps*(T+1) */
  /* Contact dynamics.  n_contacts = 0: DifferentialActionModelFreeFwdDynamics everywhere.  Otherwise
   * costset_contact[c] is the contact of the model that owns cost set c (one ContactModel per stage, as in the
   * corpus), or -1: that model has an empty ContactModelMultiple and behaves like the free dynamics. */
  int32_t n_contacts;
  int32_t integrator;                       /* EMPC_INTEGRATOR_*: IntegratedActionModelEuler / RK4 (src/factory/int-action.cpp:24-35) */
  const empc_contact_t* contacts;           /* n_contacts */
  const int32_t* costset_contact;           /* n_costsets, or NULL when n_contacts = 0 */
} empc_problem_desc_t;

/* Stop rules.  The PepMS Crocoddyl fork defines stoppingCriteria()/stoppingTest() behind two enums that eagle-mpc sets to
 * StopCriteriaCostReduction / StopTestGaps (src/sbfddp.cpp:28-29); the fork is not in the reference tree, so their
 * meaning is inferred and upstream Crocoddyl's rule is selectable beside it:
 *   COST_REDUCTION  stop_ = |cost_prev_ - cost_|                       (fork, inferred; default)
 *   QU_NORM         stop_ = sum_t ||Qu_t||^2                           (upstream SolverDDP::stoppingCriteria)
 *   GAPS            converged when stop_ < th_stop_ && gap norm < th_stop_gaps_   (fork, inferred; default)
 *   FEASIBLE        converged when was_feasible_ && stop_ < th_stop_   (upstream SolverFDDP::solve) */
enum { EMPC_STOP_CRITERIA_COST_REDUCTION = 0, EMPC_STOP_CRITERIA_QU_NORM = 1 };
enum { EMPC_STOP_TEST_GAPS = 0, EMPC_STOP_TEST_FEASIBLE = 1 };

/* Solver families of SolverTypes (include/eagle_mpc/mpc-base.hpp:36).  The Box solvers are crocoddyl's: ONE SolverFDDP::solve
 * (BOXFDDP) or SolverDDP::solve (BOXDDP) pass on the problem built WITHOUT the squashing actuation (use_squash = 0, no
 * barrier cost), where computeGains solves the box-constrained QP  min 1/2 du' Quu du + Qu' du,  u_lb <= us + du <= u_ub
 * (projected Newton, crocoddyl::BoxQP) whenever the candidate is feasible, and the rollouts clamp us_try to [u_lb, u_ub].
 * Their stop rule is upstream's: set stop_criteria = QU_NORM and stop_test = FEASIBLE (empc_box_params does). */
enum { EMPC_SOLVER_SBFDDP = 0, EMPC_SOLVER_BOXFDDP = 1, EMPC_SOLVER_BOXDDP = 2 };

/* Solver constants: eagle-mpc's (src/sbfddp.cpp:5-29) and crocoddyl SolverDDP/FDDP defaults. */
typedef struct empc_solver_params {
  int32_t maxiter;          /* solve(maxiter), default 100 */
  int32_t stop_gap_norm;    /* fork policy knob: 0 = max_t ||fs_t||_inf, 1 = sum_t ||fs_t||_1 */
  int32_t squash_quirk;     /* oracle only: emulate us_squash = "last calc on the data"  */
  int32_t stop_criteria;    /* set_stoppingCriteria (src/sbfddp.cpp:28): EMPC_STOP_CRITERIA_* */
  int32_t stop_test;        /* set_stoppingTest (src/sbfddp.cpp:29): EMPC_STOP_TEST_* (FDDP passes; the DDP clean-up always
                               uses stoppingTestFeasible, src/sbfddp.cpp:387) */
  int32_t solver_type;      /* EMPC_SOLVER_*: SolverSbFDDP (default) or crocoddyl's SolverBoxFDDP / SolverBoxDDP, the other two
                               entries of SolverTypes (include/eagle_mpc/mpc-base.hpp:36-47, src/mpc-controllers/carrot-mpc.cpp:232-242) */
  double convergence_init;  /* 1e-2 */
  double convergence_stop;  /* 1e-3 */
  double convergence_mult;  /* 1e-1 */
  double smooth_init;       /* 0.1 */
  double smooth_mult;       /* 0.5 */
  double barrier_weight;    /* 1e-3 */
  double reg_init;          /* 1e-9 */
  double reg_min;           /* 1e-9 */
  double reg_max;           /* 1e9 */
  double reg_factor;        /* 10 */
  double th_acceptstep;     /* 0.1 */
  double th_acceptnegstep;  /* 2 */
  double th_grad;           /* 1e-12 */
  double th_gaptol;         /* 1e-16 */
  double th_stepdec;        /* 0.5 */
  double th_stepinc;        /* 0.01 */
  double th_stop_gaps;      /* 1.0 */
  /* Box solvers only (crocoddyl SolverBoxFDDP / SolverBoxDDP constructors: th_stop_ = 5e-5, BoxQP(nu, 100, 0.1, 1e-5, 0)) */
  double th_stop;           /* 5e-5 */
  double boxqp_th_acceptstep; /* 0.1 */
  double boxqp_th_grad;     /* 1e-5 */
  double boxqp_reg;         /* 0 */
  int32_t boxqp_maxiter;    /* 100 */
  int32_t reserved;
} empc_solver_params_t;

/* Fills `p` with the reference defaults listed above. */
void empc_default_params(empc_solver_params_t* p);
/* Defaults of crocoddyl::SolverBoxFDDP / SolverBoxDDP (solver_type = EMPC_SOLVER_BOXFDDP or EMPC_SOLVER_BOXDDP): the reference
 * defaults above with the upstream stop rule and th_stop_ = 5e-5. */
void empc_box_params(empc_solver_params_t* p, int32_t solver_type);

/* Derived dimensions of a problem. */
typedef struct empc_dims {
  int32_t nq, nv, nx, ndx, nu, T, batch;
  int32_t tile;       /* doubles per node tile: Fx|Fu|Lxx|Lxu|Luu|Lx|Lu, padded to an even count */
} empc_dims_t;

typedef struct empc_solver empc_solver_t; /* opaque handle: owns all device memory */

/* ---- lifetime ---- */
int empc_create(const empc_problem_desc_t* desc, int32_t batch, int32_t device, empc_solver_t** out);
int empc_destroy(empc_solver_t* h);
const char* empc_last_error(void);
int empc_get_dims(const empc_solver_t* h, empc_dims_t* out);

/* ---- inputs (host pointers, copied; caller keeps ownership) ---- */
int empc_set_x0(empc_solver_t* h, const double* x0 /* batch*nx */);
/* xs: batch*(T+1)*nx or NULL (=> state.zero()); us: batch*T*nu or NULL (=> 0). */
int empc_set_candidate(empc_solver_t* h, const double* xs, const double* us, int32_t is_feasible);
int empc_set_params(empc_solver_t* h, const empc_solver_params_t* p);
/* per-OCP node map selection, batch entries in [0,n_node_maps); NULL => all 0 */
int empc_set_node_maps(empc_solver_t* h, const int32_t* ocp_map);
/* MPC retargeting: overwrite `n` cost records starting at `first_cost` and `n_pool` doubles at `pool_off`. */
int empc_update_costs(empc_solver_t* h, int32_t first_cost, int32_t n, const empc_cost_t* costs,
                      int32_t pool_off, int32_t n_pool, const double* pool);
int empc_update_node_costsets(empc_solver_t* h, const int32_t* node_costset /* n_node_maps*(T+1) */);

/* ---- batched MPC instances: one handle, `n_instances` controllers at different times ----
 * empc_replicate_instances gives every instance a private copy of the cost tables of a single-node-map problem (one
 * model per knot, src/mpc-controllers/rail-mpc.cpp:66-79): afterwards n_node_maps = n_instances, OCP b uses instance
 * b % n_instances (change with empc_set_node_maps), and the cost / pool indices of empc_update_costs address instance m
 * at m*n_costs + c and m*n_pool + i.
 * empc_set_reference_trajectory uploads the state reference of a rail controller (n_ref states of nx doubles, dt_ref_ms
 * apart: RailMpc(state_ref, dt_ref, yaml), rail-mpc.cpp:26-35).
 * empc_rail_retarget is RailMpc::updateProblem(current_time) (rail-mpc.cpp:154-200) for all instances at once, on the
 * device: times_ms[m] is the controller time of instance m, dt_node_ms the knot spacing (mpc_controller/dt). */
int empc_replicate_instances(empc_solver_t* h, int32_t n_instances);
int empc_set_reference_trajectory(empc_solver_t* h, const double* state_ref, int32_t n_ref, int32_t dt_ref_ms);
int empc_rail_retarget(empc_solver_t* h, const int64_t* times_ms /* n_node_maps */, int32_t dt_node_ms);

/* CarrotMpc::updateProblem(current_time) (src/mpc-controllers/carrot-mpc.cpp:298-401) for all instances at once, on the
 * device.  t_stages (n_stages + 1 entries, ms) is the controller's stage table (:60-75: durations clamped to at least one
 * knot spacing), is_transition[s] the stage's flag.  Per knot: stage = upper_bound(t_stages, node_time) - 1; inside the
 * trajectory "carrot_state" is on (with the piecewise-constant state reference of empc_set_reference_trajectory) unless
 * the stage is a transition stage and the knot is not the last one; past the last stage "carrot_state" is off and
 * "carrot_tail" on, tracking the last configuration at rest (:376-388).  "carrot_tail" is never switched off (:349-357). */
int empc_set_carrot_schedule(empc_solver_t* h, int32_t n_stages, const int64_t* t_stages, const uint8_t* is_transition);
int empc_carrot_retarget(empc_solver_t* h, const int64_t* times_ms /* n_node_maps */, int32_t dt_node_ms);

/* The weight schedule of a WeightedMpc (src/mpc-controllers/weighted-mpc.cpp:173-245) in flat form.  A "slot" is a cost of
 * a knot in the order of its cost set, the squashing barrier excluded; every knot carries the same slots.  For stage s and
 * slot c: match = the cost's name starts with the stage's name (:205) => active, otherwise inactive; task = its weight
 * follows the schedule base * exp(alpha * (node_time - t_end[s]) / 1000) * beta (:207-214; exponent 0 once node_time is
 * past `duration`, :229-241); "/reg" and "/limits" costs keep their weight. */
typedef struct empc_weighted_schedule {
  int32_t n_stages, n_slots;
  const int64_t* t_ini;   /* n_stages: start of each (merged) stage, ms */
  const int64_t* t_end;   /* n_stages: t_ini + duration of the stage */
  int64_t duration;       /* trajectory duration, ms */
  double alpha, beta;
  const uint8_t* match;   /* n_stages * n_slots */
  const uint8_t* task;    /* n_stages * n_slots */
  const double* base;     /* n_stages * n_slots */
} empc_weighted_schedule_t;
int empc_set_weighted_schedule(empc_solver_t* h, const empc_weighted_schedule_t* s);
/* WeightedMpc::updateProblem(current_time) for all instances at once, on the device. */
int empc_weighted_retarget(empc_solver_t* h, const int64_t* times_ms /* n_node_maps */, int32_t dt_node_ms);

/* ---- the hot path ---- */
/* Full SbFDDP solve of the whole batch (squash-smoothing schedule, FDDP passes, DDP clean-up). */
int empc_solve(empc_solver_t* h);
/* Same, but inputs (x0, warm start) are already device-resident from a previous set_* / solve:
 * restarts every OCP from its stored x0 and initial candidate (used by bench.py's `value` leg). */
int empc_reset(empc_solver_t* h);

/* RK4 plant step of the closed-loop drivers (bindings/python/eagle_mpc/utils/simulator.py:24-29): n independent
 * instances, x (n*nx), u (n*nu, thrusts and joint torques actually applied), dt in seconds; host pointers. */
int empc_plant_step(empc_solver_t* h, const double* x, const double* u, double dt, double* xnext, int32_t n);

/* Device-resident closed loop for a batch of controller instances (examples/python/mpc.py:49-61, all instances at once):
 * every OCP's plant state (its x0) is advanced by one RK4 step of dt seconds under the first squashed control of its
 * current solution, x0[b] <- plant(x0[b], us_squash[b][0], dt), and becomes the initial state of the next empc_solve,
 * which warm-starts from the solution left on the device (solver.solve(solver.xs, solver.us, iters)).  A closed-loop
 * step is: empc_*_retarget(times) -> empc_solve -> empc_plant_advance.  x_plant (batch*nx) and u_applied (batch*nu) are
 * optional host outputs (NULL to skip the copies). */
int empc_plant_advance(empc_solver_t* h, double dt, double* x_plant, double* u_applied);

/* ---- outputs (host pointers) ---- */
int empc_get_xs(const empc_solver_t* h, double* xs /* batch*(T+1)*nx */);
int empc_get_us(const empc_solver_t* h, double* us /* batch*T*nu */);
int empc_get_us_squash(const empc_solver_t* h, double* us_squash /* batch*T*nu */);
int empc_get_K(const empc_solver_t* h, double* K /* batch*T*nu*ndx */);
int empc_get_k(const empc_solver_t* h, double* k /* batch*T*nu */);
int empc_get_cost(const empc_solver_t* h, double* cost /* batch */);
int empc_get_iters(const empc_solver_t* h, int32_t* iters /* batch: iter_ as left by solve() */);
int empc_get_stop(const empc_solver_t* h, double* stop /* batch */);
int empc_get_feasible(const empc_solver_t* h, int32_t* feasible /* batch */);
int empc_get_reg(const empc_solver_t* h, double* xreg /* batch */);
/* The cost tables as they currently are on the device (after empc_update_costs / empc_replicate_instances / the retarget
 * kernels): n_costs records and n_pool doubles, both as passed to empc_create times the number of instances.  For tests. */
int empc_get_cost_tables(const empc_solver_t* h, empc_cost_t* costs, double* pool);
/* Everything SolverSbFDDP::solve leaves behind for the MPC loop (get_xs, get_us, getSquashControls, get_cost, get_stop,
 * get_iter, is_feasible) in one call; any pointer may be NULL.  Small batches travel as one packed device-to-host copy. */
int empc_get_solution(empc_solver_t* h, double* xs, double* us, double* us_squash, double* cost, double* stop, int32_t* iters,
                      int32_t* feasible);
/* ---- iteration log: the device-side stand-in for setCallbacks({CallbackVerbose}) (src/sbfddp.cpp:303-307,
 * src/mpc-controllers/carrot-mpc.cpp:244-247, bindings/python/eagle_mpc/sbfddp.hpp:63).  The whole iteration loop runs on
 * the device, so instead of calling back into the host every iteration each OCP appends one record per inner iteration
 * (at the point where the reference runs its callbacks: after the line search, the regularisation update and
 * stoppingCriteria) to a ring of `capacity` records; the host reads it after empc_solve. ---- */
typedef struct empc_iter_record {
  int32_t iter;         /* iter_ of the inner solve (SolverAbstract::get_iter inside the callback) */
  int32_t total_iter;   /* iterations executed by this solve() before this one, all passes */
  int32_t phase;        /* 0: FDDP pass, 1: DDP clean-up */
  int32_t accepted;     /* index of the accepted step length (steplength = 2^-accepted), -1: none accepted */
  int32_t is_feasible;
  int32_t reserved;
  double cost, stop, steplength, xreg; /* ureg_ = xreg_ */
  double d0, d1;        /* expected improvement of the last trial: CallbackVerbose prints grad = -d1 */
  double smooth;        /* squashing smoothing of the pass */
} empc_iter_record_t;
/* capacity = records kept per OCP (the most recent ones); 0 switches the log off (default). */
int empc_enable_iteration_log(empc_solver_t* h, int32_t capacity);
/* records of OCP `ocp` from the last solve, oldest first; *n_records = number written (<= max_records). */
int empc_get_iteration_log(const empc_solver_t* h, int32_t ocp, empc_iter_record_t* out, int32_t max_records,
                           int32_t* n_records);

/* ---- streaming solve: more OCPs than slots.  The handle's `batch` OCP slots are refilled from a queue of `n_jobs`
 * initial states as soon as an OCP finishes (on the device, between batch-iterations), so a workload whose OCPs need very
 * different numbers of iterations does not run at the pace of its slowest member (every job of the reference is an
 * independent SolverSbFDDP::solve, src/sbfddp.cpp:192-226; nothing couples them).  Each job is solve([], [], maxiter) of the
 * handle's problem from its own x0 with a fresh solver state; results are per job, in job order, bit-identical to solving
 * the job in a plain batch.  x0: n_jobs*nx.  xs (n_jobs*(T+1)*nx), us, us_squash (n_jobs*T*nu), cost, stop, iters,
 * feasible (n_jobs) may each be NULL.  Pointers: host or device memory.  K / k are not kept per job. ---- */
int empc_solve_stream(empc_solver_t* h, int32_t n_jobs, const double* x0, double* xs, double* us, double* us_squash,
                      double* cost, double* stop, int32_t* iters, int32_t* feasible);

/* total inner iterations executed by the last solve, summed over the batch (the benchmark's work unit) */
int empc_get_total_iterations(const empc_solver_t* h, int64_t* total);
/* number of kernel launches issued by the last solve and device time (ms) spent per kernel family:
 * [0]=calc_diff [1]=backward [2]=rollout [3]=decide */
int empc_get_launch_stats(const empc_solver_t* h, int64_t* launches, double* ms_by_kernel /* 4 or NULL */);
int empc_enable_kernel_timing(empc_solver_t* h, int32_t on);
/* device time of the last empc_solve (CUDA events on the solver's stream) and, per kernel family, the number of OCPs
 * each launch processed summed over the launches (x T = node-iterations) */
int empc_get_solve_stats(const empc_solver_t* h, double* solve_ms, int64_t* units_by_kernel /* 4 or NULL */);

/* ---- tile-level parity hooks (one phase on the current candidate of every OCP) ---- */
int empc_phase_calc_diff(empc_solver_t* h, double smooth);      /* calc + calcDiff + gaps */
int empc_phase_backward(empc_solver_t* h, double xreg, int32_t is_feasible, int32_t* ok /* batch */);
int empc_phase_rollout(empc_solver_t* h, double smooth, int32_t is_feasible, int32_t ddp);
int empc_get_tiles(const empc_solver_t* h, double* tiles /* batch*(T+1)*tile */);
int empc_get_xnext(const empc_solver_t* h, double* xnext /* batch*(T+1)*nx (row T unused) */);
int empc_get_node_cost(const empc_solver_t* h, double* c /* batch*(T+1) */);
int empc_get_gaps(const empc_solver_t* h, double* fs /* batch*(T+1)*ndx */);
int empc_get_Vx(const empc_solver_t* h, double* Vx /* batch*(T+1)*ndx */);
int empc_get_Vxx_fs(const empc_solver_t* h, double* g /* batch*(T+1)*ndx */);
int empc_get_dgdq(const empc_solver_t* h, double* dgdq /* batch*2 */);
int empc_get_trial(const empc_solver_t* h, int32_t alpha_index, double* xs_try, double* us_try,
                   double* cost_try /* batch */, double* dv /* batch */, int32_t* ok /* batch */);

#ifdef __cplusplus
}
#endif
#endif /* EMPC_B200_H */
