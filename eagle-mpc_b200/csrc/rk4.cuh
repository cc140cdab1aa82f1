// rk4.cuh — IntegratedActionModelRK4 inside the OCP (included from kernels.cuh inside namespace empc).
//
// Replaces crocoddyl::IntegratedActionModelRK4::calc / calcDiff (crocoddyl/core/integrator/rk4.hxx) when createProblem is
// called with integration_method = "IntegratedActionModelRK4" (src/factory/int-action.cpp:29-31; the solver handles it at
// src/sbfddp.cpp:43-44,107-108,145):
//   y_0 = x,  y_i = x (+) c_i dt k_{i-1},  k_i = [v(y_i); a(y_i, u)],  c = (0, 1/2, 1/2, 1)
//   xnext = x (+) dt/6 (k_0 + 2 k_1 + 2 k_2 + k_3),   cost = dt/6 (l_0 + 2 l_1 + 2 l_2 + l_3),  l_i = l(y_i, u)
//   calcDiff: chain rule through the stages (dyi_dx, dyi_du, dki_dx, dki_du) and the Gauss-Newton pull-back of the stage
//   cost blocks, which makes Lxu and the whole Luu non-zero (backward_kernel<D, true>).
//
// Like the contact nodes (contact.cuh) this is an overlay on the Euler path, not a tuned pipeline: no YAML of the corpus
// selects RK4 (every mpc.yaml and the examples use Euler), so the node model runs as one thread per node out of local
// memory with rolled loops — four forward-dynamics evaluations with their world-frame RNEA partials and a dense chain rule
// — and the free-node kernels are not launched at all for an RK4 problem.  The rollout chain and the trial costs call the
// non-inlined node_dyn_rk4 / rk4_node_cost.
#pragma once

__device__ __constant__ double kRk4C[4] = {0.0, 0.5, 0.5, 1.0};
__device__ __constant__ double kRk4W[4] = {1.0, 2.0, 2.0, 1.0};

template <class D>
EMPC_DI void rk4_tau(const DevModel& M, double smooth, const double* u, double* s, double* tau) {
  squash<D>(M, smooth, u, s);
  EMPC_ROLLED for (int i = 0; i < 6; ++i) {
    double t = 0;
    EMPC_ROLLED for (int j = 0; j < D::NR; ++j) t += M.tau_f[i * D::NR + j] * s[j];
    tau[i] = t;
  }
  EMPC_ROLLED for (int i = 0; i < D::NA; ++i) tau[6 + i] = s[D::NR + i];
}

// Dynamics half of calc: the rollout chain's replacement of node_dyn (node.cuh) for an RK4 problem
template <class D>
__device__ __noinline__ void node_dyn_rk4(const DevModel& M, double smooth, const double* x, const double* u, double* xnext) {
  constexpr int NV = D::NV, NDX = D::NDX, NX = D::NX;
  NodeData<D> nd;
  double tau[NV], k[NDX], ksum[NDX], y[NX], dxi[NDX];
  rk4_tau<D>(M, smooth, u, nd.s, tau);
  EMPC_ROLLED for (int i = 0; i < NX; ++i) y[i] = x[i];
  EMPC_ROLLED for (int st = 0; st < 4; ++st) {
    if (st > 0) {
      EMPC_ROLLED for (int i = 0; i < NDX; ++i) dxi[i] = kRk4C[st] * M.dt * k[i];
      state_integrate<D>(x, dxi, y);
    }
    aba<D, true>(M, y, tau, nd);
    EMPC_ROLLED for (int i = 0; i < NV; ++i) { k[i] = y[D::NQ + i]; k[NV + i] = nd.a[i]; }
    EMPC_ROLLED for (int i = 0; i < NDX; ++i) ksum[i] = (st == 0 ? 0.0 : ksum[i]) + kRk4W[st] * k[i];
  }
  EMPC_ROLLED for (int i = 0; i < NDX; ++i) dxi[i] = ksum[i] * (M.dt / 6.0);
  state_integrate<D>(x, dxi, xnext);
}

// Cost half of calc for a trial node (decide_kernel, trial_cost_kernel): dt/6 sum_i w_i l(y_i, u)
template <class D>
__device__ __noinline__ double rk4_node_cost(const DevModel& M, const CostTables& C, int costset, double smooth, const double* x,
                                             const double* u) {
  constexpr int NV = D::NV, NDX = D::NDX, NX = D::NX;
  NodeData<D> nd;
  double tau[NV], k[NDX], y[NX], dxi[NDX];
  rk4_tau<D>(M, smooth, u, nd.s, tau);
  EMPC_ROLLED for (int i = 0; i < NX; ++i) y[i] = x[i];
  double csum = 0;
  EMPC_ROLLED for (int st = 0; st < 4; ++st) {
    if (st > 0) {
      EMPC_ROLLED for (int i = 0; i < NDX; ++i) dxi[i] = kRk4C[st] * M.dt * k[i];
      state_integrate<D>(x, dxi, y);
    }
    csum += kRk4W[st] * node_cost_value<D, true>(M, C, costset, smooth, y, u, /*raw=*/true);
    if (st < 3) {
      aba<D, true>(M, y, tau, nd);
      EMPC_ROLLED for (int i = 0; i < NV; ++i) { k[i] = y[D::NQ + i]; k[NV + i] = nd.a[i]; }
    }
  }
  return csum * (M.dt / 6.0);
}

// ---------------------------------------------------------------------------------------------------------------------
// calc + calcDiff of every node of an RK4 problem, one thread per node.
// (latency-bound on its local-memory operands: 32 resident warps per SM at 64 registers ran 1.8x faster than 8 warps at
//  255 registers, 785 vs 1415 ms for 1024 move_arm solves; more than that does not help)
#ifndef EMPC_RK4_THREADS
#define EMPC_RK4_THREADS 128
#define EMPC_RK4_MINB 8
#endif
template <class D>
__global__ void __launch_bounds__(EMPC_RK4_THREADS, EMPC_RK4_MINB) rk4_node_kernel(Buffers bf, int force, double force_smooth, const __grid_constant__ DevModel M) {
  constexpr int NV = D::NV, NDX = D::NDX, NU = D::NU, NX = D::NX, NR = D::NR;
  const long long nl0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = bf.T + 1;
  if (nl0 >= (long long)bf.nb * T1) return;
  const size_t n = (size_t)bf.b0 * T1 + nl0;
  const int b = (int)(n / T1), t = (int)(n - (size_t)b * T1);
  const OcpState st = bf.st[b];
  if (!force && (st.phase == PHASE_DONE || !st.recalc)) return;
  const double smooth = force ? force_smooth : st.smooth;
  const int costset = bf.node_costset[bf.ocp_map[b] * T1 + t];
  const double dt = M.dt;

  double x[NX], u[NU];
  const double* xg = bf.xs + n * NX;
  EMPC_ROLLED for (int i = 0; i < NX; ++i) x[i] = xg[i];
  EMPC_ROLLED for (int i = 0; i < NU; ++i) u[i] = (t < bf.T) ? bf.us[((size_t)b * bf.T + t) * NU + i] : 0.0;

  double* tile = bf.tiles + n * D::TILE;
  double* Fx = tile + D::oFx; double* Fu = tile + D::oFu; double* Lxx = tile + D::oLxx; double* Lxu = tile + D::oLxu;
  double* Luu = tile + D::oLuu; double* Lx = tile + D::oLx; double* Lu = tile + D::oLu;
  EMPC_ROLLED for (int i = 0; i < D::TILE; ++i) tile[i] = 0.0;

  NodeData<D> nd;
  ContactWork<D> cw;  // (world kinematics of the stage; no contact)
  double tau[NV], ds[NU];
  rk4_tau<D>(M, smooth, u, nd.s, tau);
  EMPC_ROLLED for (int i = 0; i < NU; ++i) {
    ds[i] = 1.0;
    if (M.use_squash) {
      const double dd = (M.u_ub[i] - M.u_lb[i]) * smooth, a = dd * dd;
      const double l = u[i] - M.u_lb[i], h = u[i] - M.u_ub[i];
      ds[i] = 0.5 * (rsqrt_nr(a + l * l) * l - rsqrt_nr(a + h * h) * h);
    }
  }
  double k[NDX], ksum[NDX], y[NX], dxi[NDX];
  double dyx[NDX * NDX], dyu[NDX * NU], dkx[NDX * NDX], dku[NDX * NU];
  EMPC_ROLLED for (int i = 0; i < NX; ++i) y[i] = x[i];
  double cost = 0;
  EMPC_ROLLED for (int sg = 0; sg < 4; ++sg) {
    if (sg == 0) {
      EMPC_ROLLED for (int i = 0; i < NDX * NDX; ++i) dyx[i] = 0.0;
      EMPC_ROLLED for (int i = 0; i < NDX; ++i) dyx[i * NDX + i] = 1.0;
      EMPC_ROLLED for (int i = 0; i < NDX * NU; ++i) dyu[i] = 0.0;
    } else {
      const double cdt = kRk4C[sg] * dt;
      EMPC_ROLLED for (int i = 0; i < NDX; ++i) dxi[i] = cdt * k[i];
      state_integrate<D>(x, dxi, y);
      EMPC_ROLLED for (int i = 0; i < NDX * NDX; ++i) dyx[i] = cdt * dkx[i];
      EMPC_ROLLED for (int i = 0; i < NDX * NU; ++i) dyu[i] = cdt * dku[i];
      jintegrate_apply_dev<D>(dxi, dyx, NDX, true);
      jintegrate_apply_dev<D>(dxi, dyu, NU, false);
    }
    // differential model at (y, u): forward dynamics and its partials a_q, a_v (world-frame RNEA partials, M^-1), a_u
    aba<D, true>(M, y, tau, nd);
    EMPC_ROLLED for (int i = 0; i < NV; ++i) { k[i] = y[D::NQ + i]; k[NV + i] = nd.a[i]; }
    EMPC_ROLLED for (int i = 0; i < NDX; ++i) ksum[i] = (sg == 0 ? 0.0 : ksum[i]) + kRk4W[sg] * k[i];
    cw.nc = 0; cw.jf = 0;
    cw_world_kinematics<D>(M, nd, cw);
    {
      double Mjs[NV * NV];
      cw_crba<D>(M, nd, cw, Mjs);
      spd_inverse(Mjs, NV, cw.Minv);
    }
    {
      double oa[D::NJ][6], dq[NV * NV], dv[NV * NV];
      EMPC_ROLLED for (int i = 0; i < D::NJ; ++i) act_motion(nd.oM[i], nd.agf[i], oa[i]);
      cw_rnea_partials<D>(M, nd, cw, oa, nullptr, -1, dq, dv);
      // dki_dx = dki_dy dyi_dx, dki_du = dki_dy dyi_du + [0; a_u];  dki_dy = [[0, I], [a_q, a_v]], a_q = -Minv dq, a_v = -Minv dv
      EMPC_ROLLED for (int i = 0; i < NV; ++i) {
        double aq[NV], av[NV];
        EMPC_ROLLED for (int j = 0; j < NV; ++j) {
          double sq = 0, sv = 0;
          EMPC_ROLLED for (int m = 0; m < NV; ++m) { sq += cw.Minv[i * NV + m] * dq[m * NV + j]; sv += cw.Minv[i * NV + m] * dv[m * NV + j]; }
          aq[j] = -sq; av[j] = -sv;
        }
        EMPC_ROLLED for (int c = 0; c < NDX; ++c) {
          dkx[i * NDX + c] = dyx[(NV + i) * NDX + c];
          double s = 0;
          EMPC_ROLLED for (int j = 0; j < NV; ++j) s += aq[j] * dyx[j * NDX + c] + av[j] * dyx[(NV + j) * NDX + c];
          dkx[(NV + i) * NDX + c] = s;
        }
        EMPC_ROLLED for (int c = 0; c < NU; ++c) {
          dku[i * NU + c] = dyu[(NV + i) * NU + c];
          double s = 0;  // a_u(i, c) = (Minv A diag(ds))(i, c), A = [tau_f 0; 0 I]
          if (c < NR) { EMPC_ROLLED for (int m = 0; m < 6; ++m) s += cw.Minv[i * NV + m] * (M.tau_f[m * NR + c] * ds[c]); }
          else s = cw.Minv[i * NV + 6 + (c - NR)] * ds[c];
          EMPC_ROLLED for (int j = 0; j < NV; ++j) s += aq[j] * dyu[j * NU + c] + av[j] * dyu[(NV + j) * NU + c];
          dku[(NV + i) * NU + c] = s;
        }
      }
    }
    const double wg = kRk4W[sg] * dt / 6.0;
    EMPC_ROLLED for (int i = 0; i < NDX * NDX; ++i) Fx[i] += wg * dkx[i];
    EMPC_ROLLED for (int i = 0; i < NDX * NU; ++i) Fu[i] += wg * dku[i];
    // stage cost blocks (unscaled), then their pull-back through dyi_dx / dyi_du (rk4.hxx: ddli_ddx, ddli_ddu, ddli_dxdu)
    double sLx[NDX], sLu[NU], sLxx[NDX * NDX], sLuu[NU * NU], sLxu[NDX * NU];
    EMPC_ROLLED for (int i = 0; i < NDX * NDX; ++i) sLxx[i] = 0.0;
    EMPC_ROLLED for (int i = 0; i < NDX * NU; ++i) sLxu[i] = 0.0;
    EMPC_ROLLED for (int i = 0; i < NU * NU; ++i) sLuu[i] = 0.0;
    EMPC_ROLLED for (int i = 0; i < NDX; ++i) sLx[i] = 0.0;
    EMPC_ROLLED for (int i = 0; i < NU; ++i) sLu[i] = 0.0;
    cost += kRk4W[sg] * node_cost_derivs<D>(M, bf.ct, costset, smooth, y, u, nd, cw, nullptr, nullptr, nullptr, sLx, sLu, sLxx, sLxu, sLuu);
    EMPC_ROLLED for (int c = 0; c < NDX; ++c) { double s = 0; EMPC_ROLLED for (int i = 0; i < NDX; ++i) s += sLx[i] * dyx[i * NDX + c]; Lx[c] += wg * s; }
    EMPC_ROLLED for (int c = 0; c < NU; ++c) { double s = sLu[c]; EMPC_ROLLED for (int i = 0; i < NDX; ++i) s += sLx[i] * dyu[i * NU + c]; Lu[c] += wg * s; }
    // row r of dyi_dx^T [Lxx_i dyi_dx | Lxx_i dyi_du + Lxu_i]: one row of the products at a time, no NDX x NDX temporary
    {
      double tB[NDX * NU];  // Lxx_i dyi_du + Lxu_i
      EMPC_ROLLED for (int i = 0; i < NDX; ++i)
        EMPC_ROLLED for (int c = 0; c < NU; ++c) {
          double s = sLxu[i * NU + c];
          EMPC_ROLLED for (int j = 0; j < NDX; ++j) s += sLxx[i * NDX + j] * dyu[j * NU + c];
          tB[i * NU + c] = s;
        }
      EMPC_ROLLED for (int r = 0; r < NDX; ++r) {
        double zr[NDX];  // (dyi_dx^T Lxx_i)(r, :)
        EMPC_ROLLED for (int j = 0; j < NDX; ++j) { double s = 0; EMPC_ROLLED for (int i = 0; i < NDX; ++i) s += dyx[i * NDX + r] * sLxx[i * NDX + j]; zr[j] = s; }
        EMPC_ROLLED for (int c = 0; c < NDX; ++c) { double s = 0; EMPC_ROLLED for (int j = 0; j < NDX; ++j) s += zr[j] * dyx[j * NDX + c]; Lxx[r * NDX + c] += wg * s; }
        EMPC_ROLLED for (int c = 0; c < NU; ++c) { double s = 0; EMPC_ROLLED for (int i = 0; i < NDX; ++i) s += dyx[i * NDX + r] * tB[i * NU + c]; Lxu[r * NU + c] += wg * s; }
      }
      EMPC_ROLLED for (int r = 0; r < NU; ++r)
        EMPC_ROLLED for (int c = 0; c < NU; ++c) {
          double s = sLuu[r * NU + c];
          EMPC_ROLLED for (int i = 0; i < NDX; ++i) s += dyu[i * NU + r] * tB[i * NU + c] + sLxu[i * NU + c] * dyu[i * NU + r];
          Luu[r * NU + c] += wg * s;
        }
    }
  }
  EMPC_ROLLED for (int i = 0; i < NDX; ++i) nd.dx[i] = ksum[i] * (dt / 6.0);
  double xn[NX];
  state_integrate<D>(x, nd.dx, xn);
  EMPC_ROLLED for (int i = 0; i < NX; ++i) bf.xnext[n * NX + i] = xn[i];
  bf.node_cost[n] = cost * (dt / 6.0);
  jintegrate_apply_dev<D>(nd.dx, Fx, NDX, true);
  jintegrate_apply_dev<D>(nd.dx, Fu, NU, false);
  bf.node_dense[n] = 1;
  // gaps (SolverDDP::calcDiff): fs[0] = x0 (-) xs[0], fs[t+1] = xnext_t (-) xs[t+1]
  if (!st.is_feasible) {
    if (t < bf.T) {
      double x1[NX], f[NDX];
      EMPC_ROLLED for (int i = 0; i < NX; ++i) x1[i] = xg[NX + i];
      state_diff<D>(x1, xn, f);
      double gi = 0, g1 = 0;
      EMPC_ROLLED for (int i = 0; i < NDX; ++i) { bf.fs[(n + 1) * NDX + i] = f[i]; const double a = fabs(f[i]); gi = fmax(gi, a); g1 += a; if (isnan(a)) gi = a; }
      bf.gap_inf[n + 1] = gi; bf.gap_l1[n + 1] = g1;
    }
    if (t == 0) {
      double xx[NX], f[NDX];
      EMPC_ROLLED for (int i = 0; i < NX; ++i) xx[i] = bf.x0[(size_t)b * NX + i];
      state_diff<D>(x, xx, f);
      double gi = 0, g1 = 0;
      EMPC_ROLLED for (int i = 0; i < NDX; ++i) { bf.fs[n * NDX + i] = f[i]; const double a = fabs(f[i]); gi = fmax(gi, a); g1 += a; if (isnan(a)) gi = a; }
      bf.gap_inf[n] = gi; bf.gap_l1[n] = g1;
    }
  } else if (!st.was_feasible) {
    if (t < bf.T) {
      EMPC_ROLLED for (int i = 0; i < NDX; ++i) bf.fs[(n + 1) * NDX + i] = 0.0;
      bf.gap_inf[n + 1] = 0; bf.gap_l1[n + 1] = 0;
    }
    if (t == 0) {
      EMPC_ROLLED for (int i = 0; i < NDX; ++i) bf.fs[n * NDX + i] = 0.0;
      bf.gap_inf[n] = 0; bf.gap_l1[n] = 0;
    }
  }
}
