// calcdiff.cuh — calc + calcDiff of every shooting node of every OCP, as two kernels (included from kernels.cuh inside
// namespace empc).
//
// Replaces, per node, crocoddyl::ShootingProblem::calc/calcDiff -> IntegratedActionModelEuler::calc/calcDiff ->
// DifferentialActionModelFreeFwdDynamics -> pinocchio::aba / computeABADerivatives + CostModelSum derivatives, and the
// gap computation of SolverDDP::calcDiff (reference call sites: src/sbfddp.cpp:244, :332).
//
// Why two kernels.  The work of a node splits into a SERIAL part (squashing, ABA, Lie-group maps, residuals: chains of
// small 3x3 / 6-vector operations that do not parallelise inside a node) and a COLUMN-PARALLEL part (the NV columns of
// d tau/dq, d tau/dv, M^-1, the NDX x NDX / NDX x NU Jacobian and Hessian blocks).  A thread-per-node kernel needs the
// whole ~9 KB matrix working set per thread (local memory, thrashing L1/L2: profiles/r1_baseline.md); a
// lanes-per-node kernel would repeat the serial part on every lane.  So:
//   node_calc_kernel  thread per node   serial dynamics part; writes xnext, gaps and most of a 344-double "packet" per
//                                        node (world placements / velocities / accelerations, composite-rigid-body
//                                        sweep, squashing slopes, Lie transport blocks)
//   node_cost_kernel  thread per node   cost value + state/control cost derivative summaries (rest of the packet)
//   node_diff_kernel  8/10/16 lanes per node  column-parallel part out of shared memory; writes the node tile
//                                        Fx|Fu|Lxx|Lxu|Luu|Lx|Lu.
// The packet is stored AoSoA in groups of 8 nodes ([group][field][8]): the producer's warp (32 consecutive nodes) writes
// full 64-byte runs per field, and the consumer's block (8 nodes) reads one contiguous chunk.
#pragma once

#ifndef EMPC_STREAM_STORES
#define EMPC_STREAM_STORES 1
#endif
#if EMPC_STREAM_STORES
#define EMPC_ST_STREAM(ptr, v) __stcs((ptr), (v))
#else
#define EMPC_ST_STREAM(ptr, v) (*(ptr) = (v))
#endif

template <class D>
struct Pk {
  static constexpr int NJ = D::NJ, NV = D::NV, NDX = D::NDX, NU = D::NU;
  static constexpr int oOM = 0;                       // NJ x (R 9, p 3)      world placements of the joints
  static constexpr int oOV = oOM + 12 * NJ;           // NJ x 6               world velocities
  static constexpr int oOA = oOV + 6 * NJ;            // NJ x 6               world accelerations (with gravity)
  static constexpr int oCOMP = oOA + 6 * NJ;          // NJ x 31              composites after adding body k:
  static constexpr int cM = 0, cMC = 1, cIO = 4, cHF = 13, cG = 16, cF = 25, COMP = 31;  //   m | m c | I_o | hf | G | F
  static constexpr int oDS = oCOMP + COMP * NJ;       // NU                   squashing slopes ds/du
  static constexpr int oJE = oDS + NU;                // JeA 9 | JeQ 9 | E.R 9 | E.p 3
  static constexpr int oLX = oJE + 30;                // NDX                  sum over state costs of w Rx^T Ar
  static constexpr int oLXXB = oLX + NDX;             // 36                   ... of w Rx^T Arr Rx, base 6x6 block
  static constexpr int oLXXD = oLXXB + 36;            // NDX-6                ... diagonal of the remaining rows
  static constexpr int oLU = oLXXD + NDX - 6;         // NU
  static constexpr int oLUUD = oLU + NU;              // NU                   diagonal of Luu
  static constexpr int oFLAG = oLUUD + NU;            // 1                    node has frame costs (0/1)
  static constexpr int SIZE = oFLAG + 1;
  static constexpr int GROUP = 8;
  static constexpr int STRIDE = SIZE | 1;             // per-node stride in shared memory (odd: spreads banks)
};

template <class D>
EMPC_DI size_t pk_index(size_t n, int f) { return (n / Pk<D>::GROUP) * (size_t)(Pk<D>::SIZE * Pk<D>::GROUP) + (size_t)f * Pk<D>::GROUP + (n % Pk<D>::GROUP); }

// ---------------------------------------------------------------------------------------------------------------------
// Kernel A: serial part, one thread per node.
#ifndef EMPC_NC_FULL
#define EMPC_NC_FULL true
#endif
// The kernel is ~19 000 straight-line instructions (300 KB): far beyond the instruction caches.  All warps of an SM
// therefore run as ONE block and re-converge at block barriers between the phases, so that they stream the same code
// window together instead of each thrashing the instruction cache at its own program counter.  Barriers need every thread,
// so threads without a node to process do not return: they shadow a valid node with their stores switched off.
#ifndef EMPC_NC_THREADS
#define EMPC_NC_THREADS 128
#endif
// 2 blocks of 128 threads per SM (255 registers): with the packet written by streaming stores the smaller spill frame of
// the 255-register build beats the third resident block (calc_diff 86.5 -> 80.8 ms per step, gpurun_out/variants_b.txt)
#ifndef EMPC_NC_BLOCKS
#define EMPC_NC_BLOCKS 2
#endif
constexpr int NC_THREADS = EMPC_NC_THREADS;
#define EMPC_NC_PHASE() __syncthreads()
template <class D>
__global__ void __launch_bounds__(NC_THREADS, EMPC_NC_BLOCKS) node_calc_kernel(Buffers bf, int force, double force_smooth, const __grid_constant__ DevModel M) {
  constexpr int NJ = D::NJ, NV = D::NV, NDX = D::NDX, NU = D::NU, NX = D::NX, NR = D::NR;
  using P = Pk<D>;
  const long long nl0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = bf.T + 1;
  const long long n_nodes = (long long)bf.nb * T1;
  bool on = nl0 < n_nodes;
  const size_t n = (size_t)bf.b0 * T1 + (on ? nl0 : n_nodes - 1);
  const int b = (int)(n / T1), t = (int)(n - (size_t)b * T1);
  const OcpState st = bf.st[b];
  if (!force && (st.phase == PHASE_DONE || !st.recalc)) on = false;
  if (__syncthreads_or(on) == 0) return;  // nothing to do for the whole block
  const double smooth = force ? force_smooth : st.smooth;
  double* pk = bf.packets + pk_index<D>(n, 0);  // field f of this node: pk[f * GROUP]
  // packet fields are written once and read once by a later kernel, far beyond the reach of L2: streaming stores
  // (evict-first) keep them from displacing the local-memory lines of the resident threads
  auto put = [&](int f, double v) { if (on) EMPC_ST_STREAM(pk + (size_t)f * P::GROUP, v); };

  double x[NX], u[NU];
  const double* xg = bf.xs + n * NX;
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = xg[i];
  if (t < bf.T) {
    const double* ug = bf.us + ((size_t)b * bf.T + t) * NU;
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = ug[i];
  } else {
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = 0.0;  // calc(data,x) == calc(data,x,0), SURVEY B.7
  }
  const int costset = bf.node_costset[bf.ocp_map[b] * T1 + t];

  // squashing + thrust map (ActuationSquashingModel)
  NodeData<D> nd;
  squash<D>(M, smooth, u, nd.s);
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    double ds = 1.0;
    if (M.use_squash) {
      const double lbv = M.u_lb[i], ubv = M.u_ub[i];
      const double dd = (ubv - lbv) * smooth, a = dd * dd;
      const double l = u[i] - lbv, h = u[i] - ubv;
      ds = 0.5 * (rsqrt_nr(a + l * l) * l - rsqrt_nr(a + h * h) * h);
    }
    put(P::oDS + i, ds);
  }
  double tau[NV];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = 0;
#pragma unroll
    for (int j = 0; j < NR; ++j) s += M.tau_f[i * NR + j] * nd.s[j];
    tau[i] = s;
  }
#pragma unroll
  for (int i = 0; i < D::NA; ++i) tau[6 + i] = nd.s[NR + i];

  EMPC_NC_PHASE();
  aba_kinematics<D, EMPC_NC_FULL>(M, x, nd);
  EMPC_NC_PHASE();
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
#pragma unroll
    for (int k = 0; k < 9; ++k) put(P::oOM + 12 * i + k, nd.oM[i].R[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) put(P::oOM + 12 * i + 9 + k, nd.oM[i].p[k]);
  }

  // forward dynamics + semi-implicit Euler
  EMPC_NC_PHASE();
  aba_dynamics<D, EMPC_NC_FULL>(M, tau, nd);
  EMPC_NC_PHASE();
  double xn[NX];
  {
    const double dt = M.dt, dt2 = dt * dt;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      nd.dx[i] = x[D::NQ + i] * dt + nd.a[i] * dt2;
      nd.dx[NV + i] = nd.a[i] * dt;
    }
    state_integrate<D>(x, nd.dx, xn);
    double* xng = bf.xnext + n * NX;
    if (on) {
#pragma unroll
      for (int i = 0; i < NX; ++i) EMPC_ST_STREAM(xng + i, xn[i]);
    }
    EMPC_NC_PHASE();
    // Lie-group transport pieces of Fx / Fu
    double JeA[9], JeQ[9]; Jexp6_blocks(nd.dx, JeA, JeQ);
    SE3 E; exp6(nd.dx, E);
#pragma unroll
    for (int k = 0; k < 9; ++k) { put(P::oJE + k, JeA[k]); put(P::oJE + 9 + k, JeQ[k]); put(P::oJE + 18 + k, E.R[k]); }
#pragma unroll
    for (int k = 0; k < 3; ++k) put(P::oJE + 27 + k, E.p[k]);
  }

  // world-frame velocities / accelerations and the tip-to-base composite sweep (DESIGN.md "ABA derivatives")
  {
    double cm = 0, cmc[3] = {0, 0, 0}, cIo[9], chf[3] = {0, 0, 0}, cG[9], cF[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 9; ++i) { cIo[i] = 0; cG[i] = 0; }
#pragma unroll
    for (int k = NJ - 1; k >= 0; --k) {
      EMPC_NC_PHASE();
      const SE3& oMk = nd.oM[k];
      double ov[6], oa[6];
      act_motion(oMk, nd.v[k], ov);
      act_motion(oMk, nd.agf[k], oa);
#pragma unroll
      for (int i = 0; i < 6; ++i) { put(P::oOV + 6 * k + i, ov[i]); put(P::oOA + 6 * k + i, oa[i]); }
      double cw[3]; matvec3(oMk.R, M.com[k], cw);
#pragma unroll
      for (int i = 0; i < 3; ++i) cw[i] += oMk.p[i];
      double RI[9], Iw[9];
      matmul3(oMk.R, M.Ic[k], RI);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Iw[3 * i + j] = RI[3 * i] * oMk.R[3 * j] + RI[3 * i + 1] * oMk.R[3 * j + 1] + RI[3 * i + 2] * oMk.R[3 * j + 2];
      const double m = M.mass[k];
      double h[6], Ya[6], vh[6];
      inertia_apply(m, cw, Iw, ov, h); inertia_apply(m, cw, Iw, oa, Ya); cross_mf(ov, h, vh);
      const double c2 = dot3(cw, cw);
      double Io[9], mc[3] = {m * cw[0], m * cw[1], m * cw[2]};
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Io[3 * i + j] = Iw[3 * i + j] + m * (((i == j) ? c2 : 0.0) - cw[i] * cw[j]);
      const double* vl = ov; const double* w = ov + 3;
      double Sw[9], A1[9], Shn[9]; skew3(w, Sw); matmul3(Sw, Io, A1); skew3(h + 3, Shn);
      const double vm = dot3(vl, mc);
      cm += m;
#pragma unroll
      for (int i = 0; i < 3; ++i) { cmc[i] += mc[i]; chf[i] += h[i]; }
#pragma unroll
      for (int i = 0; i < 6; ++i) cF[i] += Ya[i] + vh[i];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          cIo[3 * i + j] += Io[3 * i + j];
          cG[3 * i + j] += A1[3 * i + j] + A1[3 * j + i] - (mc[i] * vl[j] + vl[i] * mc[j]) + ((i == j) ? 2.0 * vm : 0.0) - Shn[3 * i + j];
        }
      const int o = P::oCOMP + P::COMP * k;
      put(o + P::cM, cm);
#pragma unroll
      for (int i = 0; i < 3; ++i) { put(o + P::cMC + i, cmc[i]); put(o + P::cHF + i, chf[i]); }
#pragma unroll
      for (int i = 0; i < 9; ++i) { put(o + P::cIO + i, cIo[i]); put(o + P::cG + i, cG[i]); }
#pragma unroll
      for (int i = 0; i < 6; ++i) put(o + P::cF + i, cF[i]);
    }
  }

  // gaps (SolverDDP::calcDiff): fs[0] = x0 (-) xs[0], fs[t+1] = xnext_t (-) xs[t+1]
  EMPC_NC_PHASE();
  if (!on) return;
  if (!st.is_feasible) {
    if (t < bf.T) {
      double x1[NX], f[NDX];
#pragma unroll
      for (int i = 0; i < NX; ++i) x1[i] = xg[NX + i];
      state_diff<D>(x1, xn, f);
      double gi = 0, g1 = 0;
#pragma unroll
      for (int i = 0; i < NDX; ++i) { bf.fs[(n + 1) * NDX + i] = f[i]; const double a = fabs(f[i]); gi = fmax(gi, a); g1 += a; if (isnan(a)) gi = a; }
      bf.gap_inf[n + 1] = gi; bf.gap_l1[n + 1] = g1;
    }
    if (t == 0) {
      double xx[NX], f[NDX];
#pragma unroll
      for (int i = 0; i < NX; ++i) xx[i] = bf.x0[(size_t)b * NX + i];
      state_diff<D>(x, xx, f);
      double gi = 0, g1 = 0;
#pragma unroll
      for (int i = 0; i < NDX; ++i) { bf.fs[n * NDX + i] = f[i]; const double a = fabs(f[i]); gi = fmax(gi, a); g1 += a; if (isnan(a)) gi = a; }
      bf.gap_inf[n] = gi; bf.gap_l1[n] = g1;
    }
  } else if (!st.was_feasible) {
    if (t < bf.T) {
#pragma unroll
      for (int i = 0; i < NDX; ++i) bf.fs[(n + 1) * NDX + i] = 0.0;
      bf.gap_inf[n + 1] = 0; bf.gap_l1[n + 1] = 0;
    }
    if (t == 0) {
#pragma unroll
      for (int i = 0; i < NDX; ++i) bf.fs[n * NDX + i] = 0.0;
      bf.gap_inf[n] = 0; bf.gap_l1[n] = 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Kernel A2: the cost half of calc / calcDiff, one thread per node: cost value (node_cost) and the derivative summaries of
// the state / control type costs (packet fields oLX .. oFLAG).  Split from node_calc_kernel so that neither kernel has to
// keep the other's working set live (the dynamics + composite sweep spill less, and this one needs no dynamics at all);
// the kinematics are recomputed only for nodes whose cost set holds frame costs.
#ifndef EMPC_NCOST_BLOCKS
#define EMPC_NCOST_BLOCKS 4
#endif
template <class D>
__global__ void __launch_bounds__(128, EMPC_NCOST_BLOCKS) node_cost_kernel(Buffers bf, int force, double force_smooth, const __grid_constant__ DevModel M) {
  constexpr int NDX = D::NDX, NU = D::NU, NX = D::NX;
  using P = Pk<D>;
  const long long nl0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = bf.T + 1;
  if (nl0 >= (long long)bf.nb * T1) return;
  const size_t n = (size_t)bf.b0 * T1 + nl0;
  const int b = (int)(n / T1), t = (int)(n - (size_t)b * T1);
  const OcpState st = bf.st[b];
  if (!force && (st.phase == PHASE_DONE || !st.recalc)) return;
  const double smooth = force ? force_smooth : st.smooth;
  double* pk = bf.packets + pk_index<D>(n, 0);
  auto put = [&](int f, double v) { EMPC_ST_STREAM(pk + (size_t)f * P::GROUP, v); };
  double x[NX], u[NU];
  const double* xg = bf.xs + n * NX;
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = xg[i];
  if (t < bf.T) {
    const double* ug = bf.us + ((size_t)b * bf.T + t) * NU;
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = ug[i];
  } else {
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = 0.0;  // calc(data,x) == calc(data,x,0), SURVEY B.7
  }
  const int costset = bf.node_costset[bf.ocp_map[b] * T1 + t];
  double csum = 0;
  {
    // State costs that share a reference (typically a regularisation and a limits barrier around the same state) share
    // the residual r = x (-) ref and its Jacobian [[A, B], [0, A]] = Jlog6(ref^-1 x): they are grouped, their activation
    // derivatives are summed first (a6 = sum w Ar[0:6], d6 = sum w Arr[0:6]) and Jl^T a6, Jl^T diag(d6) Jl are formed
    // once per group, straight into the packet — no 36-entry Hessian accumulator has to stay live across the cost loop.
    double LxT[NDX - 6], LxxD[NDX - 6], Lu[NU], Luud[NU];
#pragma unroll
    for (int i = 0; i < NDX - 6; ++i) { LxT[i] = 0; LxxD[i] = 0; }
#pragma unroll
    for (int i = 0; i < NU; ++i) { Lu[i] = 0; Luud[i] = 0; }
    double flag = 0.0;
    double gref[NX], gr[NDX], gA[9], gB[9], a6[6], d6[6];
    bool g_open = false, g_first = true;
    auto flush_group = [&]() {
      // J = [[A, B], [0, A]]:  J(k, i) for k, i in 0..5
      auto J = [&](int k, int i) -> double { return (k < 3) ? ((i < 3) ? gA[3 * k + i] : gB[3 * k + i - 3]) : ((i < 3) ? 0.0 : gA[3 * (k - 3) + i - 3]); };
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        double sacc = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) sacc += J(k, i) * a6[k];
        if (g_first) put(P::oLX + i, sacc); else pk[(size_t)(P::oLX + i) * P::GROUP] += sacc;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          double h = 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) h += J(k, i) * (d6[k] * J(k, j));
          if (g_first) put(P::oLXXB + 6 * i + j, h); else pk[(size_t)(P::oLXXB + 6 * i + j) * P::GROUP] += h;
        }
      }
      g_first = false; g_open = false;
    };
    const int c0 = bf.ct.costset_begin[costset], c1 = bf.ct.costset_begin[costset + 1];
    for (int c = c0; c < c1; ++c) {
      const empc_cost_t cs = bf.ct.costs[c];
      if (!cs.active || cs.type == EMPC_COST_CONTACT_FRICTION_CONE) continue;  // (contact nodes: contact_node_kernel)
      const double wt = cs.weight;
      if (is_frame_cost(cs.type)) {  // value here, derivatives in node_diff_kernel
        csum += wt * frame_cost_value<D>(M, bf.ct, cs, smooth, x);
        flag = 1.0;
        continue;
      }
      double Ar[NDX], Arr[NDX];
      if (cs.type == EMPC_COST_STATE) {
        const double* ref = bf.ct.pool + cs.ref_off;
        bool same = g_open;
        if (same) {
#pragma unroll
          for (int i = 0; i < NX; ++i) same &= (ref[i] == gref[i]);
        }
        if (!same) {
          if (g_open) flush_group();
#pragma unroll
          for (int i = 0; i < NX; ++i) gref[i] = ref[i];
          state_diff<D>(gref, x, gr);
          SE3 Mref, Mx, Dm;
          q_to_se3(gref, Mref); q_to_se3(x, Mx); se3_inv_mul(Mref, Mx, Dm);
          Jlog6_blocks(Dm, gA, gB);
#pragma unroll
          for (int i = 0; i < 6; ++i) { a6[i] = 0; d6[i] = 0; }
          g_open = true;
        }
        csum += wt * activation<NDX>(cs.activation, gr, bf.ct.pool + cs.w_off, bf.ct.pool + cs.lb_off, bf.ct.pool + cs.ub_off, Ar, Arr);
#pragma unroll
        for (int i = 0; i < 6; ++i) { a6[i] += wt * Ar[i]; d6[i] += wt * Arr[i]; }
#pragma unroll
        for (int i = 6; i < NDX; ++i) { LxT[i - 6] += wt * Ar[i]; LxxD[i - 6] += wt * Arr[i]; }
      } else {  // EMPC_COST_CONTROL, EMPC_COST_SQUASH_BARRIER
        double r[NDX];
        SE3 rMf;
        NodeData<D>* no_kinematics = nullptr;  // control residuals never touch the kinematics
        csum += wt * cost_eval<D>(M, bf.ct, cs, smooth, x, u, *no_kinematics, r, Ar, Arr, rMf);
#pragma unroll
        for (int i = 0; i < NU; ++i) { Lu[i] += wt * Ar[i]; Luud[i] += wt * Arr[i]; }
      }
    }
    if (g_open) flush_group();
    if (g_first) {  // no state cost at all
#pragma unroll
      for (int i = 0; i < 6; ++i) put(P::oLX + i, 0.0);
#pragma unroll
      for (int i = 0; i < 36; ++i) put(P::oLXXB + i, 0.0);
    }
#pragma unroll
    for (int i = 6; i < NDX; ++i) { put(P::oLX + i, LxT[i - 6]); put(P::oLXXD + i - 6, LxxD[i - 6]); }
#pragma unroll
    for (int i = 0; i < NU; ++i) { put(P::oLU + i, Lu[i]); put(P::oLUUD + i, Luud[i]); }
    put(P::oFLAG, flag);
  }
  bf.node_cost[n] = M.dt * csum;
}

// ---------------------------------------------------------------------------------------------------------------------
// Kernel B: column-parallel part, DiffCfg::LANES lanes per node, 8 nodes per block.
template <class D>
struct DiffCfg {
  static constexpr int NJ = D::NJ, NV = D::NV, NDX = D::NDX, NU = D::NU, NA = D::NA;
  // LANES lanes work on one node (one lane per column of the NV x NV matrices); a node never straddles a warp, so the
  // hand-overs between the phases are __syncwarp's over the node's lanes.  10 lanes/node puts 3 nodes of the nv=9
  // flying arm into a warp instead of 2.
#ifdef EMPC_ND_LANES
  static constexpr int LANES = EMPC_ND_LANES;
#else
  static constexpr int LANES = (NV <= 8 && NU <= 8) ? 8 : (NV <= 10 && NU <= 10) ? 10 : 16;
#endif
  static_assert(NV <= LANES && NU <= LANES, "one lane per column");
  static constexpr int NODES = Pk<D>::GROUP, NPW = 32 / LANES, WARPS = (NODES + NPW - 1) / NPW, THREADS = 32 * WARPS;
  // per-node work area (doubles)
  static constexpr int wJc = 0;                          // NV x 6   world motion axes
  static constexpr int wYJa = wJc + 6 * NV;              // NA x 6   Ycrb_j J_j of the arm joints
  static constexpr int wBtJa = wYJa + 6 * (NA > 0 ? NA : 1);  // NA x 3
  static constexpr int wDq = wBtJa + 3 * (NA > 0 ? NA : 1);   // NV x NV  d tau/dq  -> a_q
  static constexpr int wDv = wDq + NV * NV;              // NV x NV  d tau/dv  -> a_v
  static constexpr int wMm = wDv + NV * NV;              // NV x NV  joint-space inertia -> its Cholesky factor
  static constexpr int wLinv = wMm + NV * NV;            // NV       reciprocal Cholesky pivots
  // M^-1 (NV x NV) and, later, the residual Jacobian of a frame cost (6 x NDX) live in the packet's composite area,
  // which is dead after phase B2; short arms whose composite area is too small get them in the work area instead
  static constexpr int MINV_IN_PACKET = (NV * NV <= Pk<D>::COMP * NJ) ? 1 : 0;
  static constexpr int RX_IN_PACKET = (6 * NDX <= Pk<D>::COMP * NJ) ? 1 : 0;
  static constexpr int wMinv = wLinv + NV;
  static constexpr int wRx = wMinv + (MINV_IN_PACKET ? 0 : NV * NV);
  static constexpr int WORK0 = wRx + (RX_IN_PACKET ? 0 : 6 * NDX);
  static constexpr int WORK_TOTAL = WORK0 | 1;
  // frame costs only (phase B7, when a_q and a_v are dead): LOCAL frame Jacobian 6 x NV and the Lx accumulator (NDX)
  static constexpr int wFJ = wDq, wVec = wDv;
  static_assert(6 * NV <= NV * NV && NDX <= NV * NV, "frame-cost scratch fits the a_q / a_v areas");
  // resident blocks per SM the kernel is built for (shared memory: 5 x 43.4 KB for the 3-joint flying arm)
  static constexpr int MINB = THREADS <= 96 ? 5 : 4;
  static constexpr int SMEM_DOUBLES = NODES * (Pk<D>::STRIDE + WORK_TOTAL);
};

template <class D>
__global__ void __launch_bounds__(DiffCfg<D>::THREADS, DiffCfg<D>::MINB) node_diff_kernel(Buffers bf, int force, double force_smooth, const __grid_constant__ DevModel M) {
  constexpr int NJ = D::NJ, NV = D::NV, NDX = D::NDX, NU = D::NU, NR = D::NR;
  using P = Pk<D>;
  using W = DiffCfg<D>;
  extern __shared__ __align__(16) double cd_sm[];
  const int tid = threadIdx.x;
  const int T1 = bf.T + 1;
  const size_t n_first = (size_t)bf.b0 * T1, n_end = n_first + (size_t)bf.nb * T1;
  // groups are aligned on multiples of 8 nodes in the global node index
  const size_t g0 = n_first / P::GROUP;
  const size_t grp = g0 + blockIdx.x;
  const size_t nbase = grp * P::GROUP;

  // ---- cooperative load of the 8 packets (one contiguous chunk), transposed to node-major in shared memory.  Thread tid
  // always handles node tid % 8 (THREADS is a multiple of 8), so the activity test is done once; the copies are 8-byte
  // cp.async's fired back to back and awaited together. ----
  {
    const double* src = bf.packets + grp * (size_t)(P::SIZE * P::GROUP);
    const int i = tid % P::GROUP;
    const size_t nn = nbase + i;
    bool on = nn >= n_first && nn < n_end;
    if (on) {
      const OcpState* sp = bf.st + (int)(nn / T1);
      on = force || (sp->phase != PHASE_DONE && sp->recalc);
    }
    if (on) {
      double* dst = cd_sm + i * P::STRIDE;
#pragma unroll 4
      for (int e = tid; e < P::SIZE * P::GROUP; e += W::THREADS) cp_async8(dst + e / P::GROUP, src + e);
    }
    cp_async_commit();
    // Blocks are dispatched in index order, MINB per SM: the block that will run on this SM slot after this one is
    // about PF_AHEAD groups further on.  Pull its packet from HBM into L2 now, so that its own load above hits L2.
    {
      constexpr int PF_AHEAD = 148 * W::MINB;
      constexpr int LINES = (P::SIZE * P::GROUP * 8 + 127) / 128;
      if (blockIdx.x + PF_AHEAD < gridDim.x) {
        const char* nxt = reinterpret_cast<const char*>(src + (size_t)PF_AHEAD * (P::SIZE * P::GROUP));
        for (int ln = tid; ln < LINES; ln += W::THREADS) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (size_t)ln * 128));
      }
    }
    cp_async_wait<0>();
  }
  __syncthreads();

  const int slot = (tid & 31) / W::LANES, l = (tid & 31) - slot * W::LANES;
  const int node = (tid >> 5) * W::NPW + slot;
  if (slot >= W::NPW || node >= W::NODES) return;
  const unsigned hm = (0xFFFFFFFFu >> (32 - W::LANES)) << (W::LANES * slot);  // the lanes working on this node
  const size_t n = nbase + node;
  if (n < n_first || n >= n_end) return;
  const int b = (int)(n / T1), t = (int)(n - (size_t)b * T1);
  const OcpState st = bf.st[b];
  if (!force && (st.phase == PHASE_DONE || !st.recalc)) return;
  const double smooth = force ? force_smooth : st.smooth;
  const double* pk = cd_sm + node * P::STRIDE;
  double* wk = cd_sm + W::NODES * P::STRIDE + node * W::WORK_TOTAL;
  double* tile = bf.tiles + n * D::TILE;
  const double dt = M.dt, dt2 = dt * dt;

  // ---- B1: world motion axes Jc (lane = column) ----
  if (l < NV) {
    double s[6];
    if (l < 6) {
      const double* R = pk + P::oOM; const double* p = R + 9;
      if (l < 3) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { s[a] = R[3 * a + l]; s[3 + a] = 0.0; }
      } else {
        const int c = l - 3;
        const double r0 = R[c], r1 = R[3 + c], r2 = R[6 + c];  // column c of R
        s[0] = p[1] * r2 - p[2] * r1; s[1] = p[2] * r0 - p[0] * r2; s[2] = p[0] * r1 - p[1] * r0;
        s[3] = r0; s[4] = r1; s[5] = r2;
      }
    } else {
      const int i = l - 5;
      SE3 oMi;
#pragma unroll
      for (int k = 0; k < 9; ++k) oMi.R[k] = pk[P::oOM + 12 * i + k];
#pragma unroll
      for (int k = 0; k < 3; ++k) oMi.p[k] = pk[P::oOM + 12 * i + 9 + k];
      const double S[6] = {0, 0, 0, M.axis[i][0], M.axis[i][1], M.axis[i][2]};
      act_motion(oMi, S, s);
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) wk[W::wJc + 6 * l + a] = s[a];
  }
  __syncwarp(hm);

  // ---- B2: columns of d tau/dq, d tau/dv and of the joint-space inertia (lane = column ck) ----
  {
    const int ck = l < NV ? l : NV - 1;  // surplus lanes shadow the last column (results discarded)
    const int k = ck < 6 ? 0 : ck - 5;
    const int c_begin = (k == 0) ? 0 : 5 + k, c_end = (k == 0) ? 6 : 6 + k;
    const double* cp = pk + P::oCOMP + P::COMP * k;
    const double cm = cp[P::cM];
    double cmc[3], chf[3], cIo[9], cG[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) { cmc[i] = cp[P::cMC + i]; chf[i] = cp[P::cHF + i]; }
#pragma unroll
    for (int i = 0; i < 9; ++i) { cIo[i] = cp[P::cIO + i]; cG[i] = cp[P::cG + i]; }
    auto Yc = [&](const double* uu, double* o) {
      double t1[3], t2[3], t3[3];
      cross3(cmc, uu + 3, t1); cross3(cmc, uu, t2); matvec3(cIo, uu + 3, t3);
#pragma unroll
      for (int i = 0; i < 3; ++i) { o[i] = cm * uu[i] - t1[i]; o[3 + i] = t2[i] + t3[i]; }
    };
    auto Bc = [&](const double* uu, double* o) {
      double t1[3], t2[3];
      cross3(chf, uu + 3, t1); matvec3(cG, uu + 3, t2);
#pragma unroll
      for (int i = 0; i < 3; ++i) { o[i] = -2.0 * t1[i]; o[3 + i] = t2[i]; }
    };
    double s[6], vp[6], ap[6], ovk[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      s[a] = wk[W::wJc + 6 * ck + a];
      vp[a] = (k > 0) ? pk[P::oOV + 6 * (k > 0 ? k - 1 : 0) + a] : 0.0;
      ap[a] = (k > 0) ? pk[P::oOA + 6 * (k > 0 ? k - 1 : 0) + a] : M.a0[a];
      ovk[a] = pk[P::oOV + 6 * k + a];
    }
    double dVdq[6], dAdq[6], dAdv[6], t6[6], vsum[6];
    cross_mm(vp, s, dVdq);
    cross_mm(ap, s, dAdq); cross_mm(vp, dVdq, t6);
#pragma unroll
    for (int a = 0; a < 6; ++a) { dAdq[a] += t6[a]; vsum[a] = vp[a] + ovk[a]; }
    cross_mm(vsum, s, dAdv);
    double Pq[6], Fq[6], Fv[6], YJ[6], t1[6], t2[6], cF[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) cF[a] = cp[P::cF + a];
    Yc(dAdq, t1); Bc(dVdq, t2);
#pragma unroll
    for (int a = 0; a < 6; ++a) Pq[a] = t1[a] + t2[a];
    cross_mf(s, cF, t1);
#pragma unroll
    for (int a = 0; a < 6; ++a) Fq[a] = Pq[a] + t1[a];
    Yc(dAdv, t1); Bc(s, t2);
#pragma unroll
    for (int a = 0; a < 6; ++a) Fv[a] = t1[a] + t2[a];
    Yc(s, YJ);
    if (k > 0 && l < NV) {
      double c1[3], c2v[3];
      cross3(chf, s, c1);  // 2 hf x s_v + G^T s_w
#pragma unroll
      for (int i = 0; i < 3; ++i) c2v[i] = cG[i] * s[3] + cG[3 + i] * s[4] + cG[6 + i] * s[5];
#pragma unroll
      for (int i = 0; i < 6; ++i) wk[W::wYJa + 6 * (k - 1) + i] = YJ[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) wk[W::wBtJa + 3 * (k - 1) + i] = 2.0 * c1[i] + c2v[i];
    }
    __syncwarp(hm);
    if (l < NV) {
#pragma unroll 1
      for (int cj = 0; cj < NV; ++cj) {
        double vq_, vv_;
        if (cj < c_end) {
          double jc[6];
#pragma unroll
          for (int a = 0; a < 6; ++a) jc[a] = wk[W::wJc + 6 * cj + a];
          vq_ = (cj < c_begin) ? dot6(jc, Fq) : dot6(jc, Pq);
          vv_ = dot6(jc, Fv);
          if (cj <= ck) { const double mv = dot6(jc, YJ); wk[W::wMm + cj * NV + ck] = mv; wk[W::wMm + ck * NV + cj] = mv; }
        } else {
          const double* yj = wk + W::wYJa + 6 * (cj - 6); const double* bj = wk + W::wBtJa + 3 * (cj - 6);
          vq_ = dot6(yj, dAdq) + (bj[0] * dVdq[3] + bj[1] * dVdq[4] + bj[2] * dVdq[5]);
          vv_ = dot6(yj, dAdv) + (bj[0] * s[3] + bj[1] * s[4] + bj[2] * s[5]);
        }
        wk[W::wDq + cj * NV + ck] = vq_; wk[W::wDv + cj * NV + ck] = vv_;
      }
    }
  }
  __syncwarp(hm);

  double* Minv = W::MINV_IN_PACKET ? const_cast<double*>(pk) + P::oCOMP : wk + W::wMinv;
  // ---- B3: Cholesky of the joint-space inertia in shared memory (lane = row, left-looking, same subtraction order as
  // llt_inplace_inv) and one column of M^-1 per lane ----
  {
    double* Lm = wk + W::wMm;
    double* Linv = wk + W::wLinv;
    double row[NV];  // this lane's row of L (entries k < j are final when column j is processed)
    const int i = l < NV ? l : NV - 1;
#pragma unroll
    for (int k = 0; k < NV; ++k) row[k] = Lm[i * NV + k];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      // pivot: every lane computes it from row j (broadcast reads), so no extra hand-over is needed
      double d = Lm[j * NV + j];
#pragma unroll
      for (int k = 0; k < j; ++k) { const double ljk = Lm[j * NV + k]; d -= ljk * ljk; }
      const double inv = rsqrt_nr(d);
      double sij = row[j];
#pragma unroll
      for (int k = 0; k < j; ++k) sij -= row[k] * Lm[j * NV + k];
      sij = (i == j) ? d * inv : sij * inv;
      row[j] = sij;
      __syncwarp(hm);
      if (l < NV && i >= j) Lm[i * NV + j] = sij;
      if (l == j) Linv[j] = inv;
      __syncwarp(hm);
    }
    // column l of M^-1: L y = e_l, L^T x = y
    double e[NV];
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      double sacc = (r == l) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < r; ++k) sacc -= Lm[r * NV + k] * e[k];
      e[r] = sacc * Linv[r];
    }
#pragma unroll
    for (int r = NV - 1; r >= 0; --r) {
      double sacc = e[r];
#pragma unroll
      for (int k = r + 1; k < NV; ++k) sacc -= Lm[k * NV + r] * e[k];
      e[r] = sacc * Linv[r];
    }
    if (l < NV) {
#pragma unroll
      for (int r = 0; r < NV; ++r) Minv[r * NV + l] = e[r];
    }
  }
  __syncwarp(hm);

  // ---- B4 + B5: a_q = -M^-1 dtau/dq, a_v = -M^-1 dtau/dv (lane = column j), and straight from those registers the two
  // columns j and NV + j of Fx: rows 6.. are scaled copies, rows 0..5 go through the Lie-group transport ----
  const double* JeA = pk + P::oJE; const double* JeQ = JeA + 9;
  if (l < NV) {
    double cq[NV], cv[NV], rq[NV], rv[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) { cq[k] = wk[W::wDq + k * NV + l]; cv[k] = wk[W::wDv + k * NV + l]; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double sq = 0, sv = 0;
#pragma unroll
      for (int k = 0; k < NV; ++k) { const double mi = Minv[i * NV + k]; sq += mi * cq[k]; sv += mi * cv[k]; }
      rq[i] = -sq; rv[i] = -sv;
    }
    double* Fx = tile + D::oFx;
    // rows 6 .. NV-1 (arm joint positions) and NV .. 2 NV-1 (velocities)
#pragma unroll
    for (int i = 6; i < NV; ++i) {
      EMPC_ST_STREAM(Fx + i * NDX + l, rq[i] * dt2 + ((i == l) ? 1.0 : 0.0));
      EMPC_ST_STREAM(Fx + i * NDX + NV + l, rv[i] * dt2 + ((i == l) ? dt : 0.0));
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      EMPC_ST_STREAM(Fx + (NV + i) * NDX + l, rq[i] * dt);
      EMPC_ST_STREAM(Fx + (NV + i) * NDX + NV + l, rv[i] * dt + ((i == l) ? 1.0 : 0.0));
    }
    // rows 0..5: Je [dt^2 a_q | dt^2 a_v + dt I] + [Ad(E^-1) | 0];  Ad(E^-1) = (X*)^T of E = exp(dx), entry (a, c) = Xs[6 c + a]
    SE3 E;
#pragma unroll
    for (int k = 0; k < 9; ++k) E.R[k] = pk[P::oJE + 18 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) E.p[k] = pk[P::oJE + 27 + k];
    double Xs[36]; force_action_matrix(E, Xs);
    double tq[6], tv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { tq[k] = rq[k] * dt2; tv[k] = rv[k] * dt2 + ((k == l) ? dt : 0.0); }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      double sq = 0, sv = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double je;  // Je = [[A,Q],[0,A]]
        if (a < 3) je = (k < 3) ? JeA[3 * a + k] : JeQ[3 * a + (k - 3)];
        else je = (k < 3) ? 0.0 : JeA[3 * (a - 3) + (k - 3)];
        sq += je * tq[k]; sv += je * tv[k];
      }
      if (l < 6) {
        double xs_ = 0.0;
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) xs_ = (cc == l) ? Xs[6 * cc + a] : xs_;
        sq += xs_;
      }
      EMPC_ST_STREAM(Fx + a * NDX + l, sq);
      EMPC_ST_STREAM(Fx + a * NDX + NV + l, sv);
    }
  }

  // ---- B6: Fu = [dt^2; dt] M^-1 A diag(ds), rows 0..5 transported; lane = column ----
  {
    double* Fu = tile + D::oFu;
    __syncwarp(hm);
    static_assert(NU <= W::LANES, "one lane per control");
    if (l < NU) {
      const int j = l;
      const double dsj = pk[P::oDS + j];
      // column j of d tau / d u = (thrust map | identity) diag(ds): the same NV-term product for every lane, no
      // divergence between rotor and arm columns (the zero terms add exactly)
      double acol[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) acol[k] = 0.0;
      if (j < NR) {
#pragma unroll
        for (int k = 0; k < 6; ++k) acol[k] = M.tau_f[k * NR + j] * dsj;
      } else {
#pragma unroll
        for (int k = 6; k < NV; ++k) acol[k] = (k == 6 + (j - NR)) ? dsj : 0.0;
      }
      double top[6];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < NV; ++k) s += Minv[i * NV + k] * acol[k];
        if (i < 6) top[i] = dt2 * s; else EMPC_ST_STREAM(Fu + i * NU + j, dt2 * s);
        EMPC_ST_STREAM(Fu + (NV + i) * NU + j, dt * s);
      }
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          double je;
          if (a < 3) je = (k < 3) ? JeA[3 * a + k] : JeQ[3 * a + (k - 3)];
          else je = (k < 3) ? 0.0 : JeA[3 * (a - 3) + (k - 3)];
          s += je * top[k];
        }
        EMPC_ST_STREAM(Fu + a * NU + j, s);
      }
    }
  }
  __syncwarp(hm);  // Minv, a_q, a_v are dead from here on: their areas become the frame-cost scratch

  // ---- B7: cost derivatives ----
  {
    double* gLxx = tile + D::oLxx; double* gLuu = tile + D::oLuu;
    double* gLx = tile + D::oLx; double* gLu = tile + D::oLu;
    static_assert(D::oLxx % 2 == 0 && D::oLxu % 2 == 0 && NDX % 2 == 0, "double2 stores need even offsets");
    // Lxu == 0 and the off-diagonal part of Luu == 0 for every cost the factories build: those tile entries keep the
    // zeros of the allocation (empc_create) and are never written, nor read by backward_kernel
    for (int i = l; i < NU; i += W::LANES) EMPC_ST_STREAM(gLuu + i * NU + i, pk[P::oLUUD + i] * dt);
    for (int i = l; i < NU; i += W::LANES) EMPC_ST_STREAM(gLu + i, pk[P::oLU + i] * dt);
    if (D::TILE != D::TILE0 && l == 0) tile[D::TILE0] = 0.0;
    auto lxx_state = [&](int i, int j) -> double {  // state-cost part: 6x6 block + diagonal
      if (i < 6 && j < 6) return pk[P::oLXXB + 6 * i + j];
      if (i == j) return pk[P::oLXXD + i - 6];
      return 0.0;
    };
    // Without frame costs Lxx is a 6x6 block plus a diagonal.  Its zeros are already in the tile (zero-initialised by
    // empc_create) unless an earlier pass wrote a dense Lxx for this node (frame costs come and go when an MPC
    // controller retargets its cost sets): bf.node_dense remembers that, and only then the full block is rewritten.
    const bool was_dense = bf.node_dense[n] != 0;
    if (pk[P::oFLAG] == 0.0) {
      if (was_dense) {
        for (int e2 = l; e2 < NDX * NDX / 2; e2 += W::LANES) {
          const int e = 2 * e2, i = e / NDX, j = e - i * NDX;
          reinterpret_cast<double2*>(gLxx)[e2] = make_double2(lxx_state(i, j) * dt, lxx_state(i, j + 1) * dt);
        }
        if (l == 0) bf.node_dense[n] = 0;
      } else {
        for (int e = l; e < 36; e += W::LANES) { const int i = e / 6, j = e - 6 * i; EMPC_ST_STREAM(gLxx + i * NDX + j, pk[P::oLXXB + e] * dt); }
        for (int i = 6 + l; i < NDX; i += W::LANES) EMPC_ST_STREAM(gLxx + i * NDX + i, pk[P::oLXXD + i - 6] * dt);
      }
      for (int i = l; i < NDX; i += W::LANES) EMPC_ST_STREAM(gLx + i, pk[P::oLX + i] * dt);
    } else {
      // The dense Lxx of a node with frame costs is accumulated in place in the tile: lane l owns the columns l,
      // l + LANES, ... from the state-cost initial value to the final scaling, so no hand-over is involved.
      double* Lxx = gLxx;
      double* Lxv = wk + W::wVec;
      for (int j = l; j < NDX; j += W::LANES)
        for (int i = 0; i < NDX; ++i) Lxx[i * NDX + j] = lxx_state(i, j);
      for (int i = l; i < NDX; i += W::LANES) Lxv[i] = pk[P::oLX + i];
      __syncwarp(hm);
      // frame costs: residual / activation on every lane (serial), Jacobian columns and Hessian entries across lanes
      NodeData<D> nd;  // only oM and v are used by the frame costs
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
#pragma unroll
        for (int k = 0; k < 9; ++k) nd.oM[i].R[k] = pk[P::oOM + 12 * i + k];
#pragma unroll
        for (int k = 0; k < 3; ++k) nd.oM[i].p[k] = pk[P::oOM + 12 * i + 9 + k];
        double ov[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) ov[k] = pk[P::oOV + 6 * i + k];
        actinv_motion(nd.oM[i], ov, nd.v[i]);
      }
      const int costset = bf.node_costset[bf.ocp_map[b] * T1 + t];
      const int c0 = bf.ct.costset_begin[costset], c1 = bf.ct.costset_begin[costset + 1];
      double* fJ = wk + W::wFJ;   // 6 x NV
      double* Rx = W::RX_IN_PACKET ? const_cast<double*>(pk) + P::oCOMP : wk + W::wRx;  // 6 x NDX
      for (int c = c0; c < c1; ++c) {
        const empc_cost_t cs = bf.ct.costs[c];
        if (!cs.active) continue;
        if (!is_frame_cost(cs.type)) continue;
        double r[NDX], Ar[NDX], Arr[NDX];
        SE3 rMf;
        cost_eval<D>(M, bf.ct, cs, smooth, nullptr, nullptr, nd, r, Ar, Arr, rMf);
        const double wt = cs.weight;
        const int f = cs.frame, jf = M.frame_joint[f];
        SE3 oMf; frame_placement<D>(M, nd, f, oMf);
        for (int e = l; e < 6 * NDX; e += W::LANES) Rx[e] = 0.0;
        if (l < NV) {  // frame Jacobian column (LOCAL)
          const int jc = (l < 6) ? 0 : l - 5;
          double jcw[6], o[6];
#pragma unroll
          for (int a = 0; a < 6; ++a) jcw[a] = wk[W::wJc + 6 * l + a];
          actinv_motion(oMf, jcw, o);
#pragma unroll
          for (int a = 0; a < 6; ++a) fJ[a * NV + l] = (jc <= jf) ? o[a] : 0.0;
        }
        __syncwarp(hm);
        int nres = 3;
        bool full = false;
        if (cs.type == EMPC_COST_FRAME_PLACEMENT) {
          nres = 6;
          double Jl[36]; Jlog6(rMf, Jl);
          if (l < NV) {
#pragma unroll
            for (int a = 0; a < 6; ++a) {
              double sacc = 0;
#pragma unroll
              for (int k = 0; k < 6; ++k) sacc += Jl[6 * a + k] * fJ[k * NV + l];
              Rx[a * NDX + l] = sacc;
            }
          }
        } else if (cs.type == EMPC_COST_FRAME_ROTATION) {
          double wv[3], th, Jl[9]; log3(rMf.R, wv, th); Jlog3(th, wv, Jl);
          if (l < NV) {
#pragma unroll
            for (int a = 0; a < 3; ++a) Rx[a * NDX + l] = Jl[3 * a] * fJ[3 * NV + l] + Jl[3 * a + 1] * fJ[4 * NV + l] + Jl[3 * a + 2] * fJ[5 * NV + l];
          }
        } else if (cs.type == EMPC_COST_FRAME_TRANSLATION) {
          if (l < NV) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
              Rx[a * NDX + l] = oMf.R[3 * a] * fJ[0 * NV + l] + oMf.R[3 * a + 1] * fJ[1 * NV + l] + oMf.R[3 * a + 2] * fJ[2 * NV + l];
          }
        } else {  // FRAME_VELOCITY
          nres = 6; full = true;
          if (l < NV) {
            if (l >= 6) {  // base columns have no parent body: zero
              const int kk = l - 5;
              if (kk <= jf) {
                double ovp[6], jcw[6], cr[6], o[6];
#pragma unroll
                for (int a = 0; a < 6; ++a) { ovp[a] = pk[P::oOV + 6 * (kk - 1) + a]; jcw[a] = wk[W::wJc + 6 * l + a]; }
                cross_mm(ovp, jcw, cr); actinv_motion(oMf, cr, o);
#pragma unroll
                for (int a = 0; a < 6; ++a) Rx[a * NDX + l] = o[a];
              }
            }
#pragma unroll
            for (int a = 0; a < 6; ++a) Rx[a * NDX + NV + l] = fJ[a * NV + l];
          }
        }
        __syncwarp(hm);
        const int ncols = full ? NDX : NV;
        double wA[6], wAr[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { wA[k] = (k < nres) ? Arr[k] : 0.0; wAr[k] = (k < nres) ? Ar[k] : 0.0; }
        for (int j = l; j < ncols; j += W::LANES) {
          double rj[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) rj[k] = Rx[k * NDX + j];
          double sacc = 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) sacc += rj[k] * wAr[k];
          Lxv[j] += wt * sacc;
#pragma unroll 1
          for (int i = 0; i < ncols; ++i) {
            double h = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) h += Rx[k * NDX + i] * (wA[k] * rj[k]);
            Lxx[i * NDX + j] += wt * h;
          }
        }
        __syncwarp(hm);
      }
      for (int j = l; j < NDX; j += W::LANES)
        for (int i = 0; i < NDX; ++i) Lxx[i * NDX + j] *= dt;
      for (int i = l; i < NDX; i += W::LANES) gLx[i] = Lxv[i] * dt;
      if (l == 0 && !was_dense) bf.node_dense[n] = 1;
    }
  }
}
