// rollout.cuh — nonlinear forward rollouts of the line search, all EMPC_N_ALPHAS step lengths of an OCP concurrently
// (included from kernels.cuh inside namespace empc).
//
// Replaces crocoddyl::SolverFDDP::forwardPass / SolverSbFDDP::forwardPassDDP as called by tryStep / tryStepDDP
// (src/sbfddp.cpp:264, :410-460).  The reference tries the step lengths one after the other and stops at the first
// accepted one; the trials are independent, so the device evaluates all of them at once and decide_kernel picks the
// first accepted in the reference's order.
//
// Speculation width: a launch evaluates W consecutive step lengths per OCP.  Stage A (W = 4: alpha = 1 .. 1/8) runs for
// every active OCP; stage B (the remaining 1/16 .. 1/512) only for the OCPs whose stage-A trials were all rejected
// (decide_kernel flags them `pending`), which is rare — so the common case costs 4 rollouts per OCP instead of 10 and one
// warp per SM sub-partition with the full register file, not 10 trials squeezed into 168 registers.
//
// Mapping: one warp = 32/W OCPs x W step lengths, one lane per trial, sequential in t.  The chain
// x_t -> u_t -> ABA -> x_{t+1} is latency-bound, so the kernel is organised around per-node latency:
//   * the per-node inputs shared by the 10 trials of an OCP (xs, us, k, fs, Vxx.fs, K: 235 doubles for flying_arm_3) are
//     fetched one node ahead with cp.async into a double-buffered shared-memory stage (coalesced, read once per OCP
//     instead of once per trial) and read back as shared-memory broadcasts;
//   * the robot model comes in as a __grid_constant__ kernel parameter: with the joint loops fully unrolled every model
//     access is a constant-bank operand and the kinematic chain stays in registers (no local-memory NodeData);
//   * each lane streams its own xs_try / us_try rows straight from registers (L2 merges the partial sectors);
//   * the chain carries the DYNAMICS only.  The running cost of a trial does not feed back into the states, so the node
//     costs of all (trial, node) pairs are evaluated afterwards by trial_cost_kernel at full occupancy and summed in node
//     order by trial_sum_kernel (the reference's cost_try_ += cost, forwardPass): ~25 % fewer instructions on the chain.
#pragma once

struct RoParams {
  int force, force_feasible, force_ddp;
  double force_smooth;
  int a_begin;  // first step-length index of this launch (0: stage A, the width of stage A: stage B)
  int box;      // SolverBoxFDDP / SolverBoxDDP::forwardPass: trial controls are clamped to [u_lb, u_ub] (overlay instantiation only)
};

// Step lengths tried by stage A; stage B covers the rest.  4 when the batch fills the GPU (one warp per sub-partition
// already at 4096 OCPs x 4 trials); 8 for smaller batches, where the extra trials run on sub-partitions that would idle
// and stage B (alpha <= 1/256) is practically never needed.
constexpr int RO_WIDTH_A = 4, RO_WIDTH_A_SMALL = 8;

template <class D, int W>
struct RoCfg {
  static constexpr int NX = D::NX, NDX = D::NDX, NU = D::NU;
  static constexpr int OCPS = 32 / W;  // OCPs per warp
  // stage of one OCP and one node (doubles).  16-byte aligned members first (NDX and NU*NDX are even).
  static constexpr int oK = 0, oFs = oK + NU * NDX, oG = oFs + NDX, oXs = oG + NDX, oUs = oXs + NX, oKk = oUs + NU,
                       STAGE0 = oKk + NU, STAGE = STAGE0 + (STAGE0 & 1);
  static constexpr int SMEM_DOUBLES = 2 * OCPS * STAGE;
};

// CONTACT: the overlay instantiation — the problem has contact stages, whose nodes integrate the contact dynamics
// (contact.cuh), or it uses the RK4 integrator (rk4.cuh)
template <class D, int W, bool CONTACT = false>
__global__ void __launch_bounds__(32, 4) rollout_kernel(Buffers bf, RoParams P, const __grid_constant__ DevModel M) {
  constexpr int NX = D::NX, NDX = D::NDX, NU = D::NU;
  using S = RoCfg<D, W>;
  extern __shared__ __align__(16) double ro_sm[];
  const int lane = threadIdx.x;
  const int o = lane / W;                              // OCP slot of this lane
  const int j = lane - o * W;                          // position among the lanes of the slot
  const int ai = P.a_begin + j;                        // step-length index of this lane (>= EMPC_N_ALPHAS: idle lane)
  const int bl = blockIdx.x * S::OCPS + o;             // OCP of this lane, relative to the window
  const int T = bf.T, T1 = T + 1;
  const bool has_ocp = bl < bf.nb;
  const int b = bf.b0 + (has_ocp ? bl : 0);

  bool slot_on = has_ocp;                              // the OCP takes part in this launch (all its lanes agree)
  int ddp = 0, feasible = 0;
  double smooth = 0.0;
  {
    const OcpState st = bf.st[b];
    if (!P.force && (st.phase == PHASE_DONE || st.bw_fail || (P.a_begin > 0 && !st.pending))) slot_on = false;
    ddp = P.force ? P.force_ddp : (st.phase == PHASE_DDP);
    feasible = P.force ? P.force_feasible : st.is_feasible;
    smooth = P.force ? P.force_smooth : st.smooth;
  }
  if (__ballot_sync(0xffffffffu, slot_on) == 0) return;
  const bool mine = slot_on && ai < EMPC_N_ALPHAS;     // this lane owns a trial
  bool active = mine;

  const double alpha = 1.0 / (double)(1 << (ai < EMPC_N_ALPHAS ? ai : 0));
  const bool plain = ddp || feasible || ai == 0;
  const size_t trial = (size_t)(ai < EMPC_N_ALPHAS ? ai : 0) * bf.B + b;
  double* xs_try = bf.xs_try + trial * T1 * NX;
  double* us_try = bf.us_try + trial * T * NU;

  // Asynchronous fetch of node t: the W lanes of a slot copy their own OCP's record (K | fs | g | xs | us | k) in
  // interleaved 16-byte (8-byte where the row length is odd) pieces; no loop over slots, constant offsets only.
  double* my_stage = ro_sm + (size_t)o * S::STAGE;
  auto prefetch = [&](int t, int buf) {
    if (slot_on) {
      double* dst = my_stage + (size_t)buf * S::OCPS * S::STAGE;
      const size_t node = (size_t)b * T1 + t;
      const double* fs = bf.fs + node * NDX;
      const double* g = bf.g + node * NDX;
      const double* xs = bf.xs + node * NX;
#pragma unroll
      for (int c = 0; c < (NDX / 2 + W - 1) / W; ++c) {
        const int i = j + c * W;
        if (i < NDX / 2) { cp_async16(dst + S::oFs + 2 * i, fs + 2 * i); cp_async16(dst + S::oG + 2 * i, g + 2 * i); }
      }
#pragma unroll
      for (int c = 0; c < (NX + W - 1) / W; ++c) {
        const int i = j + c * W;
        if (i < NX) cp_async8(dst + S::oXs + i, xs + i);
      }
      if (t < T) {
        const size_t nodeu = (size_t)b * T + t;
        const double* Kg = bf.K + nodeu * NU * NDX;
        const double* us = bf.us + nodeu * NU;
        const double* kg = bf.k + nodeu * NU;
#pragma unroll
        for (int c = 0; c < (NU * NDX / 2 + W - 1) / W; ++c) {
          const int i = j + c * W;
          if (i < NU * NDX / 2) cp_async16(dst + S::oK + 2 * i, Kg + 2 * i);
        }
#pragma unroll
        for (int c = 0; c < (NU + W - 1) / W; ++c) {
          const int i = j + c * W;
          if (i < NU) { cp_async8(dst + S::oUs + i, us + i); cp_async8(dst + S::oKk + i, kg + i); }
        }
      }
    }
    cp_async_commit();
  };

  double xn[NX];  // running state (xnext of the previous node)
  {
    const double* src = ddp ? (bf.xs_try0 + (size_t)b * NX) : (bf.x0 + (size_t)b * NX);
#pragma unroll
    for (int i = 0; i < NX; ++i) xn[i] = src[i];
  }
  double dv = 0;
  int ok = 1;

  prefetch(0, 0);
  for (int t = 0; t <= T; ++t) {
    if (t < T) { prefetch(t + 1, (t + 1) & 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncwarp();
    const double* in = my_stage + (size_t)(t & 1) * S::OCPS * S::STAGE;
    if (active) {
      double xt[NX];
      if (plain) {
#pragma unroll
        for (int i = 0; i < NX; ++i) xt[i] = xn[i];
      } else {
        double gap[NDX];
#pragma unroll
        for (int i = 0; i < NDX; ++i) gap[i] = in[S::oFs + i] * (alpha - 1);
        state_integrate<D>(xn, gap, xt);
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) __stcs(xs_try + (size_t)t * NX + i, xt[i]);  // trial rows: read once by decide_kernel, streaming
      double dx[NDX];
      {
        double x0t[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) x0t[i] = in[S::oXs + i];
        state_diff<D>(x0t, xt, dx);
      }
      if (!ddp && !feasible) {
        // dv -= fs . Vxx diff(xs_try, xs)  ==  + (Vxx fs) . diff(xs, xs_try)   (Vxx symmetric)
        double s = 0;
#pragma unroll
        for (int i = 0; i < NDX; ++i) s += in[S::oG + i] * dx[i];
        dv += s;
      }
      if (t < T) {
        double u[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          // (one accumulator per row: splitting the row products into partial sums shortens the dependent chain but costs
          //  registers this kernel does not have — measured slower, B = 1: 0.48 vs 0.42 ms per two rollouts)
          double kd = 0;
#pragma unroll
          for (int jj = 0; jj < NDX; ++jj) kd += in[S::oK + i * NDX + jj] * dx[jj];
          u[i] = in[S::oUs + i] - in[S::oKk + i] * alpha - kd;
          if (CONTACT) { if (P.box) u[i] = fmin(fmax(u[i], M.u_lb[i]), M.u_ub[i]); }  // us_try.cwiseMax(u_lb).cwiseMin(u_ub)
          __stcs(us_try + (size_t)t * NU + i, u[i]);
        }
        if (CONTACT) {
          if (M.integrator == EMPC_INTEGRATOR_RK4) {
            node_dyn_rk4<D>(M, smooth, xt, u, xn);
          } else {
            const int ci = bf.ct.costset_contact[bf.node_costset[(size_t)bf.ocp_map[b] * T1 + t]];
            if (ci >= 0) node_dyn_contact<D>(M, bf.ct.contacts + ci, smooth, xt, u, xn);
            else node_dyn<D, true>(M, smooth, xt, u, xn);
          }
        } else {
          node_dyn<D, true>(M, smooth, xt, u, xn);
        }
        int worst = -1;  // raiseIfNaN(xnext.lpNorm<Infinity>()): integer test on the high words (node.cuh: raise_bits)
#pragma unroll
        for (int i = 0; i < NX; ++i) worst = max(worst, raise_bits(xn[i]));
        if (worst >= 0) { ok = 0; active = false; }  // "forward_error": this step length is skipped by decide_kernel
      }
    }
    if (__ballot_sync(0xffffffffu, active) == 0) break;  // every trial of this warp hit a forward error
    __syncwarp();  // this node's shared-memory reads are done before the next prefetch overwrites the other buffer
  }
  cp_async_wait<0>();
  if (mine) {
    const size_t n = (size_t)b * EMPC_N_ALPHAS + ai;
    bf.dv[n] = dv;
    bf.ok[n] = ok;  // cost_try and its NaN test follow in trial_cost_kernel / trial_sum_kernel
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Node costs of the trial trajectories: one thread per (OCP, step length of this stage, node).
template <class D, bool CONTACT = false>
__global__ void __launch_bounds__(128, 4) trial_cost_kernel(Buffers bf, RoParams P, int width, const __grid_constant__ DevModel M) {
  constexpr int NX = D::NX, NU = D::NU;
  const int T = bf.T, T1 = T + 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_ocp = (long long)width * T1;
  if (idx >= (long long)bf.nb * per_ocp) return;
  const int bl = (int)(idx / per_ocp);
  const int rem = (int)(idx - (long long)bl * per_ocp);
  const int j = rem / T1, t = rem - j * T1;
  const int ai = P.a_begin + j;
  if (ai >= EMPC_N_ALPHAS) return;
  const int b = bf.b0 + bl;
  const OcpState st = bf.st[b];
  if (!P.force && (st.phase == PHASE_DONE || st.bw_fail || (P.a_begin > 0 && !st.pending))) return;
  if (!bf.ok[(size_t)b * EMPC_N_ALPHAS + ai]) return;  // the rollout stopped on a NaN state: rows beyond it are stale
  const double smooth = P.force ? P.force_smooth : st.smooth;
  const size_t trial = (size_t)ai * bf.B + b;
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
  const int costset = bf.node_costset[(size_t)bf.ocp_map[b] * T1 + t];
  bf.trial_node_cost[trial * T1 + t] = node_cost_value<D, CONTACT>(M, bf.ct, costset, smooth, x, u);
}

// cost_try = sum of the node costs in node order (one warp per trial: coalesced loads, lane 0 adds in order)
__global__ void __launch_bounds__(128) trial_sum_kernel(Buffers bf, RoParams P, int width) {
  __shared__ double sbuf[4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * 4 + warp;
  if (w >= (long long)bf.nb * width) return;
  const int bl = (int)(w / width), j = (int)(w - (long long)bl * width);
  const int ai = P.a_begin + j;
  if (ai >= EMPC_N_ALPHAS) return;
  const int b = bf.b0 + bl;
  const OcpState st = bf.st[b];
  if (!P.force && (st.phase == PHASE_DONE || st.bw_fail || (P.a_begin > 0 && !st.pending))) return;
  const size_t n = (size_t)b * EMPC_N_ALPHAS + ai;
  if (!bf.ok[n]) return;
  const int T1 = bf.T + 1;
  const double* c = bf.trial_node_cost + ((size_t)ai * bf.B + b) * T1;
  double s = 0;
  for (int base = 0; base < T1; base += 32) {
    const int cnt = min(32, T1 - base);
    if (lane < cnt) sbuf[warp][lane] = c[base + lane];
    __syncwarp();
    if (lane == 0) for (int i = 0; i < cnt; ++i) s += sbuf[warp][i];
    __syncwarp();
  }
  if (lane == 0) {
    bf.cost_try[n] = s;
    if (raise_if_nan(s)) bf.ok[n] = 0;  // raiseIfNaN(cost_try_)
  }
}
