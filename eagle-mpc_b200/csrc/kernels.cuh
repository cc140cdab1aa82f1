// kernels.cuh — the four kernel families of the batched SbFDDP iteration (FP64, sm_100a).
//
//   node_calc_kernel  thread per (OCP, node): calc + gaps + serial half of calcDiff  (src/sbfddp.cpp:244,332 -> SolverDDP::calcDiff)
//   node_diff_kernel  8/10/16 lanes per (OCP, node): column-parallel half of calcDiff, writes the node tiles
//   backward_kernel   warp per OCP: Riccati sweep, Quu LLT, gains, expected-improvement sums, regularisation retry
//                                                                              (SolverDDP::backwardPass/computeGains, :245-253)
//   rollout_kernel    thread per (OCP, step length): all 10 alphas concurrently (SolverFDDP::forwardPass, :416-460)
//   decide_kernel     block per OCP: line-search acceptance, regularisation, stop tests, smoothing schedule
//                                                                              (src/sbfddp.cpp:192-226, :260-312, :348-390)
#pragma once
#include "node.cuh"

namespace empc {

enum { PHASE_FDDP = 0, PHASE_DDP = 1, PHASE_DONE = 2 };

struct OcpState {
  double smooth;       // squashing smoothing currently applied (SquashingModelSmoothSat::smooth_)
  double smooth_next;  // smooth_ member of SolverSbFDDP (already multiplied for the next pass)
  double convergence, th_stop;
  double xreg, cost, cost_prev, stop, steplength;
  double dg, dq;       // FDDP expected improvement (with gap terms)
  double dg0, dq0;     // DDP expected improvement (src/sbfddp.cpp:395-408)
  double gap_inf, gap_l1;
  int phase, iter, total_iters, is_feasible, was_feasible, recalc, bw_fail, iters_out;
  int accepted;
  int pending;  // every stage-A step length was rejected: stage B of the line search has to run (rollout.cuh)
  double qu2;   // sum_t ||Qu_t||^2 of the last backward pass (upstream's stoppingCriteria, EMPC_STOP_CRITERIA_QU_NORM)
  double d0_last, d1_last;  // expected-improvement pair of the last trial the line search evaluated (iteration log)
  int log_count;            // iteration records written by this solve (iteration log ring)
  int pad_;
};

struct Buffers {
  // problem
  const DevModel* model;
  CostTables ct;
  const int* node_costset;
  const int* ocp_map;
  int B, T;
  int b0, nb;  // OCP window [b0, b0+nb) this launch works on (batch groups pipelined on separate streams)
  // per-OCP
  OcpState* st;
  const double* x0;
  double* xs; double* us;
  double* xs_try0;
  // per node
  double* packets;  // calcdiff.cuh: per-node hand-over from node_calc_kernel to node_diff_kernel (AoSoA, groups of 8)
  unsigned char* node_dense;  // per node: the tile currently holds a dense Lxx (frame costs), see node_diff_kernel
  double* tiles; double* xnext; double* node_cost; double* fs; double* gap_inf; double* gap_l1;
  double* K; double* k; double* Vx; double* g; double* nodesc;
  // trials
  double* xs_try; double* us_try; double* cost_try; double* dv; int* ok;
  double* trial_node_cost;  // [alpha][OCP][T+1] node costs of the trial trajectories
  double* us_squash;
  double* qu2;  // per node ||Qu_t||^2
  int* n_active;
  empc_iter_record_t* iter_log;  // [OCP][log_cap] ring, or nullptr
  int log_cap;
};

// ---------------------------------------------------------------------------------------------------------------------
#include "calcdiff.cuh"

#include "contact.cuh"

#include "rk4.cuh"

#include "backward.cuh"

// ---------------------------------------------------------------------------------------------------------------------
#include "rollout.cuh"

// ---------------------------------------------------------------------------------------------------------------------
struct DecideParams {
  empc_solver_params_t P;
  int stage;    // 0: after rollout stage A (step lengths [0, width_a)); 1: after stage B, pending OCPs only
  int width_a;  // step lengths rolled out by stage A: 4 for large batches, 8 when the GPU has idle sub-partitions anyway
};

__device__ __forceinline__ void increase_reg(OcpState& st, const empc_solver_params_t& P) {
  st.xreg *= P.reg_factor;
  if (st.xreg > P.reg_max) st.xreg = P.reg_max;
}
__device__ __forceinline__ void decrease_reg(OcpState& st, const empc_solver_params_t& P) {
  st.xreg /= P.reg_factor;
  if (st.xreg < P.reg_min) st.xreg = P.reg_min;
}
// end of solveFDDP / solveDDP for one OCP: advance the SbFDDP outer schedule (src/sbfddp.cpp:205-222)
__device__ __forceinline__ void end_inner_solve(OcpState& st, const empc_solver_params_t& P) {
  st.total_iters += st.iter + 1;
  if (P.solver_type != EMPC_SOLVER_SBFDDP) {  // crocoddyl's SolverBoxFDDP / SolverBoxDDP: solve() is this one pass
    st.phase = PHASE_DONE;
    st.iters_out = st.total_iters - 1;
    return;
  }
  bool start_ddp = false;
  if (st.phase == PHASE_FDDP) {
    st.smooth_next *= P.smooth_mult;
    st.convergence *= P.convergence_mult;
    if (st.convergence >= P.convergence_stop) {
      // next FDDP pass: squashingUpdate/barrierUpdate, th_stop_ = convergence_, solveFDDP(maxiter,false,reg_init)
      st.smooth = st.smooth_next;
      st.th_stop = st.convergence;
      st.is_feasible = 0; st.was_feasible = 0;
      st.xreg = P.reg_init; st.recalc = 1; st.iter = 0;
      return;
    }
    start_ddp = !st.is_feasible;
  }
  if (start_ddp) {
    st.phase = PHASE_DDP;
    st.xreg = P.reg_init; st.was_feasible = 0; st.recalc = 1; st.iter = 0;
    return;
  }
  st.phase = PHASE_DONE;
  st.iters_out = st.total_iters - 1;
}

// One block per OCP.  The step lengths of the stage are visited in the reference's order (src/sbfddp.cpp:260-290,
// :348-368); for each one whose rollout succeeded the block evaluates the node costs of that trial trajectory in parallel
// (thread per node), thread 0 adds them in node order (cost_try_) and applies the acceptance test — and the loop stops at
// the first accepted step, so the costs of the later, speculative rollouts are never computed.
template <class D, bool CONTACT = false>
__global__ void __launch_bounds__(256, 2) decide_kernel(Buffers bf, DecideParams dp, const __grid_constant__ DevModel M) {
  constexpr int NX = D::NX, NU = D::NU;
  constexpr int CH = 512;  // nodes per chunk of the ordered cost sum
  const int b = bf.b0 + blockIdx.x;
  __shared__ int s_acc, s_last, s_go;
  __shared__ int s_stop[2];  // one slot per parity of the step-length index: thread 0 may already write the flag of step n + 1
                             // (a failed rollout has no barrier of its own) while slower threads still read the one of step n
  __shared__ double s_smooth;
  __shared__ double s_cost[CH];
  const empc_solver_params_t& P = dp.P;
  const int T = bf.T, T1 = T + 1;
  const int tid = threadIdx.x;
  const int n_begin = dp.stage == 0 ? 0 : dp.width_a, n_end = dp.stage == 0 ? dp.width_a : EMPC_N_ALPHAS;
  OcpState st;  // live in thread 0 only
  int acc = -1, last = -1;
  bool mine = false;
  if (tid == 0) {
    st = bf.st[b];
    mine = st.phase != PHASE_DONE && (dp.stage == 0 || st.pending);
    int go = 0;
    if (mine) {
      if (dp.stage == 1) { st.pending = 0; atomicSub(bf.n_active, 1); }  // re-counted below if still active
      if (st.bw_fail) {
        st.bw_fail = 0;
        end_inner_solve(st, P);  // computeDirection gave up at reg_max: the inner solve returns false
      } else {
        go = 1;
      }
    }
    s_go = go; s_smooth = st.smooth; s_acc = -1; s_last = -1;
  }
  __syncthreads();
  if (s_go) {
    const double smooth = s_smooth;
    const int* costsets = bf.node_costset + (size_t)bf.ocp_map[b] * T1;
    for (int n = n_begin; n < n_end; ++n) {
      const size_t tn = (size_t)b * EMPC_N_ALPHAS + n;
      const int okn = bf.ok[tn];  // uniform over the block
      double cost_try = 0.0;
      if (okn) {
        // node costs of trial n, chunk by chunk; thread 0 accumulates in node order
        const size_t trial = (size_t)n * bf.B + b;
        for (int base = 0; base < T1; base += CH) {
          const int cnt = min(CH, T1 - base);
          for (int tt = tid; tt < cnt; tt += blockDim.x) {
            const int t = base + tt;
            double x[NX], u[NU];
            const double* xg = bf.xs_try + (trial * T1 + t) * NX;
#pragma unroll
            for (int i = 0; i < NX; ++i) x[i] = xg[i];
            if (t < T) {
              const double* ug = bf.us_try + (trial * T + t) * NU;
#pragma unroll
              for (int i = 0; i < NU; ++i) u[i] = ug[i];
            } else {
#pragma unroll
              for (int i = 0; i < NU; ++i) u[i] = 0.0;
            }
            s_cost[tt] = node_cost_value<D, CONTACT>(M, bf.ct, costsets[t], smooth, x, u);
          }
          __syncthreads();
          if (tid == 0) for (int i = 0; i < cnt; ++i) cost_try += s_cost[i];
          __syncthreads();
        }
      }
      if (tid == 0) {
        st.steplength = 1.0 / (double)(1 << n);
        last = n;
        int stop = 0;
        if (okn) {
          bf.cost_try[tn] = cost_try;
          if (raise_if_nan(cost_try)) bf.ok[tn] = 0;  // raiseIfNaN(cost_try_): "forward_error", try the next step length
          else {
            const int ddp = st.phase == PHASE_DDP;
            const double dV = st.cost - cost_try;
            double d0, d1;
            if (ddp) { d0 = st.dg0; d1 = st.dq0; }
            else {
              const double dv = st.is_feasible ? 0.0 : bf.dv[tn];
              d0 = st.dg + dv; d1 = st.dq - 2 * dv;
            }
            st.d0_last = d0; st.d1_last = d1;
            const double dVexp = st.steplength * (d0 + 0.5 * st.steplength * d1);
            bool accept = false;
            if (dVexp >= 0) {
              if (ddp) accept = (d0 < P.th_grad || !st.is_feasible || dV > P.th_acceptstep * dVexp);
              else accept = (d0 < P.th_grad || dV > P.th_acceptstep * dVexp);
            } else if (!ddp) {
              accept = dV > P.th_acceptnegstep * dVexp;
            }
            if (accept) {
              st.was_feasible = st.is_feasible;
              st.is_feasible = ddp ? 1 : ((st.was_feasible || n == 0) ? 1 : 0);
              st.cost_prev = st.cost; st.cost = cost_try;
              acc = n;
              stop = 1;
            }
          }
        }
        s_stop[n & 1] = stop;
      }
      __syncthreads();
      if (s_stop[n & 1]) break;
    }
  }
  if (tid == 0 && mine) {
    bool finish = true;
    if (s_go) {
      const int ddp = st.phase == PHASE_DDP;
      if (acc < 0 && n_end < EMPC_N_ALPHAS) {
        // none of the stage-A step lengths passed: the smaller ones are evaluated by rollout stage B, then decide again
        st.pending = 1;
        atomicAdd(bf.n_active, 1);
        bf.st[b] = st;
        finish = false;
        last = -1;
      } else {
        st.recalc = acc >= 0 ? 1 : 0;
        st.accepted = acc;
        bool ended = false;
        if (st.steplength > P.th_stepdec) decrease_reg(st, P);
        if (st.steplength <= P.th_stepinc) {
          increase_reg(st, P);
          if (st.xreg == P.reg_max) { end_inner_solve(st, P); ended = true; }
        }
        if (!ended) {
          // stoppingCriteria(): the fork's StopCriteriaCostReduction (inferred) or upstream's sum ||Qu||^2
          st.stop = (P.stop_criteria == EMPC_STOP_CRITERIA_QU_NORM) ? st.qu2 : fabs(st.cost_prev - st.cost);
          // callbacks run here in the reference (src/sbfddp.cpp:303-307, :381-385): one record per iteration
          if (bf.iter_log && bf.log_cap > 0) {
            empc_iter_record_t r;
            r.iter = st.iter; r.total_iter = st.total_iters + st.iter; r.phase = st.phase; r.accepted = acc;
            r.is_feasible = st.is_feasible; r.reserved = 0;
            r.cost = st.cost; r.stop = st.stop; r.steplength = st.steplength; r.xreg = st.xreg;
            r.d0 = st.d0_last; r.d1 = st.d1_last; r.smooth = st.smooth;
            bf.iter_log[(size_t)b * bf.log_cap + (st.log_count % bf.log_cap)] = r;
            st.log_count += 1;
          }
          // stoppingTest() of the FDDP passes: the fork's StopTestGaps (inferred) or upstream's feasibility rule;
          // stoppingTestFeasible() of the DDP clean-up (src/sbfddp.cpp:387)
          bool stop_now;
          if (ddp || P.stop_test == EMPC_STOP_TEST_FEASIBLE) stop_now = st.was_feasible && st.stop < st.th_stop;
          else {
            const double gn = st.is_feasible ? 0.0 : (P.stop_gap_norm == 0 ? st.gap_inf : st.gap_l1);
            stop_now = st.stop < st.th_stop && gn < P.th_stop_gaps;
          }
          if (stop_now) end_inner_solve(st, P);
          else {
            st.iter += 1;
            if (st.iter >= P.maxiter) { st.iter = P.maxiter - 1; end_inner_solve(st, P); }
          }
        }
      }
    }
    if (finish) {
      if (st.phase != PHASE_DONE) { atomicAdd(bf.n_active, 1); if (st.recalc) atomicAdd(bf.n_active + 1, 1); }
      bf.st[b] = st;
    }
    s_acc = acc; s_last = last;
  }
  __syncthreads();
  acc = s_acc; last = s_last;
  if (last >= 0) {
    // xs_try_[0] persists into the DDP phase from the last forwardPass executed (src/sbfddp.cpp:430)
    const size_t trial = (size_t)last * bf.B + b;
    for (int i = tid; i < NX; i += blockDim.x) bf.xs_try0[(size_t)b * NX + i] = bf.xs_try[trial * T1 * NX + i];
  }
  if (acc >= 0) {
    const size_t trial = (size_t)acc * bf.B + b;
    // candidate copy (setCandidate): 4 independent loads in flight per thread
    auto copy4 = [&](double* __restrict__ dst, const double* __restrict__ src, int cnt) {
      int i = tid;
      for (; i + 3 * (int)blockDim.x < cnt; i += 4 * blockDim.x) {
        const double a = src[i], b2 = src[i + blockDim.x], c = src[i + 2 * blockDim.x], d = src[i + 3 * blockDim.x];
        dst[i] = a; dst[i + blockDim.x] = b2; dst[i + 2 * blockDim.x] = c; dst[i + 3 * blockDim.x] = d;
      }
      for (; i < cnt; i += blockDim.x) dst[i] = src[i];
    };
    copy4(bf.xs + (size_t)b * T1 * NX, bf.xs_try + trial * T1 * NX, T1 * NX);
    copy4(bf.us + (size_t)b * T * NU, bf.us_try + trial * T * NU, T * NU);
  }
}

// state of an OCP at the start of solve() (src/sbfddp.cpp:198-210)
__device__ __forceinline__ void init_ocp_state(OcpState& st, const empc_solver_params_t& P, int is_feasible_arg, double cost_prev) {
  st.smooth = P.smooth_init; st.smooth_next = P.smooth_init;
  st.convergence = P.convergence_init; st.th_stop = P.convergence_init;
  st.xreg = P.reg_init; st.cost = 0; st.cost_prev = cost_prev; st.stop = 0; st.steplength = 1;
  st.dg = st.dq = st.dg0 = st.dq0 = 0; st.gap_inf = 0; st.gap_l1 = 0;
  st.iter = 0; st.total_iters = 0; st.is_feasible = 0; st.was_feasible = 0; st.recalc = 1; st.bw_fail = 0;
  st.iters_out = 0; st.accepted = -1; st.pending = 0;
  st.qu2 = 0; st.d0_last = 0; st.d1_last = 0; st.log_count = 0; st.pad_ = 0;
  if (P.solver_type != EMPC_SOLVER_SBFDDP) {
    // crocoddyl::SolverBoxFDDP / SolverBoxDDP (src/mpc-controllers/carrot-mpc.cpp:236-241): one SolverFDDP::solve or
    // SolverDDP::solve pass with th_stop_ = 5e-5 that honours the caller's feasibility flag (setCandidate)
    st.th_stop = P.th_stop; st.is_feasible = is_feasible_arg;
    st.phase = (P.solver_type == EMPC_SOLVER_BOXDDP) ? PHASE_DDP : PHASE_FDDP;
    return;
  }
  // solveFDDP(maxiter, false, ...) overrides the caller's flag (src/sbfddp.cpp:210,230)
  if (P.convergence_init >= P.convergence_stop) st.phase = PHASE_FDDP;
  else { st.phase = is_feasible_arg ? PHASE_DONE : PHASE_DDP; st.is_feasible = is_feasible_arg; }
  if (st.phase == PHASE_DONE) st.iters_out = -1;
}
__global__ void init_state_kernel(Buffers bf, empc_solver_params_t P, int is_feasible_arg, int nx) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bf.B) return;
  OcpState st;
  init_ocp_state(st, P, is_feasible_arg, bf.st[b].cost_prev);
  bf.st[b] = st;
  for (int i = 0; i < nx; ++i) bf.xs_try0[(size_t)b * nx + i] = bf.x0[(size_t)b * nx + i];  // xs_try_[0] = x0 (:198)
}

// ---- streaming solve (empc_solve_stream): the handle's OCP slots are refilled from a queue of jobs between
// batch-iterations, so a batch with very different iteration counts per OCP does not run at the pace of its slowest member.
struct StreamBuffers {
  int n_jobs;
  const double* job_x0;   // n_jobs x nx
  double* out_xs;         // n_jobs x (T+1) x nx, or nullptr
  double* out_us;         // n_jobs x T x nu, or nullptr
  double* out_us_squash;  // n_jobs x T x nu, or nullptr
  double* out_cost; double* out_stop; int* out_iters; int* out_feasible;  // n_jobs
  int* slot_job;          // per slot: job being solved, -1: none
  int* queue_next;        // next job to hand out
};
// One block per slot.  start = 1: hand job b to slot b (the first `batch` jobs).  Otherwise: a slot whose OCP has finished
// writes its result to the job's rows and takes the next job from the queue (x0 -> slot, zero candidate, fresh state:
// exactly solve([], [], maxiter) of a new solver, src/sbfddp.cpp:192-210).
template <class D>
__global__ void __launch_bounds__(128) stream_refill_kernel(Buffers bf, StreamBuffers sb, empc_solver_params_t P, int start) {
  constexpr int NX = D::NX, NU = D::NU;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int T = bf.T, T1 = T + 1;
  __shared__ int s_job;
  __shared__ double s_smooth;
  int job = -1;
  if (start) {
    job = b < sb.n_jobs ? b : -1;
  } else {
    const int cur = sb.slot_job[b];
    if (cur < 0) return;
    if (tid == 0) { const OcpState st = bf.st[b]; s_job = (st.phase == PHASE_DONE) ? 1 : 0; s_smooth = st.smooth; }
    __syncthreads();
    if (!s_job) return;  // still solving
    // ---- harvest ----
    if (sb.out_xs) for (int i = tid; i < T1 * NX; i += blockDim.x) sb.out_xs[(size_t)cur * T1 * NX + i] = bf.xs[(size_t)b * T1 * NX + i];
    if (sb.out_us) for (int i = tid; i < T * NU; i += blockDim.x) sb.out_us[(size_t)cur * T * NU + i] = bf.us[(size_t)b * T * NU + i];
    if (sb.out_us_squash) {
      const DevModel& M = *bf.model;
      for (int t = tid; t < T; t += blockDim.x) {
        double u[NU], s[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) u[i] = bf.us[((size_t)b * T + t) * NU + i];
        squash<D>(M, s_smooth, u, s);
#pragma unroll
        for (int i = 0; i < NU; ++i) sb.out_us_squash[((size_t)cur * T + t) * NU + i] = s[i];
      }
    }
    if (tid == 0) {
      const OcpState st = bf.st[b];
      sb.out_cost[cur] = st.cost; sb.out_stop[cur] = st.stop; sb.out_iters[cur] = st.iters_out; sb.out_feasible[cur] = st.is_feasible;
      const int nxt = atomicAdd(sb.queue_next, 1);
      s_job = nxt < sb.n_jobs ? nxt : -1;
    }
    __syncthreads();   // (also: every thread is done reading xs / us of the finished OCP)
    job = s_job;
  }
  if (tid == 0) sb.slot_job[b] = job;
  if (job < 0) {
    if (start && tid == 0) { OcpState st; init_ocp_state(st, P, 0, 0.0); st.phase = PHASE_DONE; st.recalc = 0; bf.st[b] = st; }
    return;
  }
  // ---- load the job: x0, xs[t] = state.zero(), us[t] = 0, fresh solver state ----
  for (int i = tid; i < NX; i += blockDim.x) {
    const double v = sb.job_x0[(size_t)job * NX + i];
    const_cast<double*>(bf.x0)[(size_t)b * NX + i] = v; bf.xs_try0[(size_t)b * NX + i] = v;  // (x0 is read-only for every other kernel)
  }
  for (int i = tid; i < T1 * NX; i += blockDim.x) bf.xs[(size_t)b * T1 * NX + i] = ((i % NX) == 6) ? 1.0 : 0.0;
  for (int i = tid; i < T * NU; i += blockDim.x) bf.us[(size_t)b * T * NU + i] = 0.0;
  // Box solvers: k_[t] is the warm start of the node's box QP; a job is a fresh solver (k_ = 0), whichever slot it lands in
  if (P.solver_type != EMPC_SOLVER_SBFDDP)
    for (int i = tid; i < T * NU; i += blockDim.x) bf.k[(size_t)b * T * NU + i] = 0.0;
  if (tid == 0) {
    OcpState st;
    init_ocp_state(st, P, 0, 0.0);
    bf.st[b] = st;
    if (!start && st.phase != PHASE_DONE) { atomicAdd(bf.n_active, 1); atomicAdd(bf.n_active + 1, 1); }
  }
}

// fillSquashedOutputs (src/sbfddp.cpp:479-486): us_squash[t] = s(us[t]) with the smoothing of the last pass
template <class D>
__global__ void squash_out_kernel(Buffers bf) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= bf.B * bf.T) return;
  const int b = n / bf.T;
  const DevModel& M = *bf.model;
  double u[D::NU], s[D::NU];
#pragma unroll
  for (int i = 0; i < D::NU; ++i) u[i] = bf.us[(size_t)n * D::NU + i];
  squash<D>(M, bf.st[b].smooth, u, s);
#pragma unroll
  for (int i = 0; i < D::NU; ++i) bf.us_squash[(size_t)n * D::NU + i] = s[i];
}

// RK4 plant of the closed-loop drivers (bindings/python/eagle_mpc/utils/simulator.py:7-29: IntegratedActionModelRK4 over
// DifferentialActionModelFreeFwdDynamics with the plain multicopter actuation, no costs).  One thread per instance.
template <class D>
__global__ void plant_rk4_kernel(const DevModel* Mp, const double* xin, const double* __restrict__ uin, double dt, double* xout, int n,
                                 size_t u_stride) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  const DevModel& M = *Mp;
  double x[D::NX], u[D::NU], tau[D::NV];
#pragma unroll
  for (int i = 0; i < D::NX; ++i) x[i] = xin[(size_t)b * D::NX + i];
#pragma unroll
  for (int i = 0; i < D::NU; ++i) u[i] = uin[(size_t)b * u_stride + i];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double t = 0;
#pragma unroll
    for (int j = 0; j < D::NR; ++j) t += M.tau_f[i * D::NR + j] * u[j];
    tau[i] = t;
  }
#pragma unroll
  for (int i = 0; i < D::NA; ++i) tau[6 + i] = u[D::NR + i];
  const double c[4] = {0.0, 0.5, 0.5, 1.0}, wgt[4] = {1.0, 2.0, 2.0, 1.0};
  double ksum[D::NDX], kprev[D::NDX];
#pragma unroll
  for (int i = 0; i < D::NDX; ++i) { ksum[i] = 0; kprev[i] = 0; }
#pragma unroll 1
  for (int s = 0; s < 4; ++s) {
    double y[D::NX], dxs[D::NDX];
#pragma unroll
    for (int i = 0; i < D::NDX; ++i) dxs[i] = dt * c[s] * kprev[i];
    if (s == 0) {
#pragma unroll
      for (int i = 0; i < D::NX; ++i) y[i] = x[i];
    } else {
      state_integrate<D>(x, dxs, y);
    }
    NodeData<D> nd;
    aba<D>(M, y, tau, nd);
#pragma unroll
    for (int i = 0; i < D::NV; ++i) { kprev[i] = y[D::NQ + i]; kprev[D::NV + i] = nd.a[i]; }
#pragma unroll
    for (int i = 0; i < D::NDX; ++i) ksum[i] += wgt[s] * kprev[i];
  }
  double dx[D::NDX], xn[D::NX];
#pragma unroll
  for (int i = 0; i < D::NDX; ++i) dx[i] = ksum[i] * (dt / 6.0);
  state_integrate<D>(x, dx, xn);
#pragma unroll
  for (int i = 0; i < D::NX; ++i) xout[(size_t)b * D::NX + i] = xn[i];
}

// empc_get_solution for small batches: xs | us | us_squash | per-OCP (cost, stop, iter, is_feasible) packed into one
// buffer, so that an MPC step reads its result with a single device-to-host copy
__global__ void pack_solution_kernel(Buffers bf, double* __restrict__ out, size_t nxs, size_t nus) {
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = gridDim.x * (size_t)blockDim.x;
  for (size_t i = tid; i < nxs; i += nth) out[i] = bf.xs[i];
  for (size_t i = tid; i < nus; i += nth) { out[nxs + i] = bf.us[i]; out[nxs + nus + i] = bf.us_squash[i]; }
  double* sc = out + nxs + 2 * nus;
  for (size_t b = tid; b < (size_t)bf.B; b += nth) {
    const OcpState& s = bf.st[b];
    sc[4 * b] = s.cost; sc[4 * b + 1] = s.stop; sc[4 * b + 2] = (double)s.iters_out; sc[4 * b + 3] = (double)s.is_feasible;
  }
}

// SolverAbstract::setCandidate with empty warm starts: xs[t] = state.zero(), us[t] = 0
__global__ void zero_candidate_kernel(double* xs, size_t n_states, int nx) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_states * (size_t)nx) return;
  xs[i] = ((int)(i % nx) == 6) ? 1.0 : 0.0;
}

__global__ void override_state_kernel(Buffers bf, double xreg, int is_feasible, int set_xreg) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bf.B) return;
  if (set_xreg) bf.st[b].xreg = xreg;
  bf.st[b].is_feasible = is_feasible;
  bf.st[b].bw_fail = 0;
  bf.st[b].pending = 0;
}

}  // namespace empc
