// kernels.cuh — the four kernel families of the batched SbFDDP iteration (FP64, sm_100a).
//
//   calc_diff_kernel  thread per (OCP, node): calc + calcDiff + gaps            (src/sbfddp.cpp:244,332 -> SolverDDP::calcDiff)
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
  int accepted, pad_;
};

struct Buffers {
  // problem
  const DevModel* model;
  CostTables ct;
  const int* node_costset;
  const int* ocp_map;
  int B, T;
  // per-OCP
  OcpState* st;
  const double* x0;
  double* xs; double* us;
  double* xs_try0;
  // per node
  double* tiles; double* xnext; double* node_cost; double* fs; double* gap_inf; double* gap_l1;
  double* K; double* k; double* Vx; double* g; double* nodesc;
  // trials
  double* xs_try; double* us_try; double* cost_try; double* dv; int* ok;
  double* us_squash;
  int* n_active;
};

// ---------------------------------------------------------------------------------------------------------------------
template <class D>
__global__ void __launch_bounds__(128) calc_diff_kernel(Buffers bf, int force, double force_smooth) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = bf.T + 1;
  if (n >= bf.B * T1) return;
  const int b = n / T1, t = n - b * T1;
  const OcpState st = bf.st[b];
  if (!force && (st.phase == PHASE_DONE || !st.recalc)) return;
  const double smooth = force ? force_smooth : st.smooth;
  const DevModel& M = *bf.model;
  double x[D::NX], u[D::NU];
  const double* xg = bf.xs + (size_t)n * D::NX;
#pragma unroll
  for (int i = 0; i < D::NX; ++i) x[i] = xg[i];
  if (t < bf.T) {
    const double* ug = bf.us + ((size_t)b * bf.T + t) * D::NU;
#pragma unroll
    for (int i = 0; i < D::NU; ++i) u[i] = ug[i];
  } else {
#pragma unroll
    for (int i = 0; i < D::NU; ++i) u[i] = 0.0;  // calc(data,x) == calc(data,x,0), SURVEY B.7
  }
  const int costset = bf.node_costset[bf.ocp_map[b] * T1 + t];
  NodeData<D> nd;
  double xn[D::NX], cost;
  node_calc<D>(M, bf.ct, costset, smooth, x, u, nd, xn, cost);
  double* xng = bf.xnext + (size_t)n * D::NX;
#pragma unroll
  for (int i = 0; i < D::NX; ++i) xng[i] = xn[i];
  bf.node_cost[n] = cost;
  node_calc_diff<D>(M, bf.ct, costset, smooth, x, u, nd, bf.tiles + (size_t)n * D::TILE);
  // gaps (SolverDDP::calcDiff): fs[0] = x0 (-) xs[0], fs[t+1] = xnext_t (-) xs[t+1]
  if (!st.is_feasible) {
    if (t < bf.T) {
      double x1[D::NX], f[D::NDX];
#pragma unroll
      for (int i = 0; i < D::NX; ++i) x1[i] = xg[D::NX + i];
      state_diff<D>(x1, xn, f);
      double gi = 0, g1 = 0;
#pragma unroll
      for (int i = 0; i < D::NDX; ++i) { bf.fs[(size_t)(n + 1) * D::NDX + i] = f[i]; const double a = fabs(f[i]); gi = fmax(gi, a); g1 += a; if (isnan(a)) gi = a; }
      bf.gap_inf[n + 1] = gi; bf.gap_l1[n + 1] = g1;
    }
    if (t == 0) {
      double xx[D::NX], f[D::NDX];
#pragma unroll
      for (int i = 0; i < D::NX; ++i) xx[i] = bf.x0[(size_t)b * D::NX + i];
      state_diff<D>(x, xx, f);
      double gi = 0, g1 = 0;
#pragma unroll
      for (int i = 0; i < D::NDX; ++i) { bf.fs[(size_t)n * D::NDX + i] = f[i]; const double a = fabs(f[i]); gi = fmax(gi, a); g1 += a; if (isnan(a)) gi = a; }
      bf.gap_inf[n] = gi; bf.gap_l1[n] = g1;
    }
  } else if (!st.was_feasible) {
    if (t < bf.T) {
#pragma unroll
      for (int i = 0; i < D::NDX; ++i) bf.fs[(size_t)(n + 1) * D::NDX + i] = 0.0;
      bf.gap_inf[n + 1] = 0; bf.gap_l1[n + 1] = 0;
    }
    if (t == 0) {
#pragma unroll
      for (int i = 0; i < D::NDX; ++i) bf.fs[(size_t)n * D::NDX + i] = 0.0;
      bf.gap_inf[n] = 0; bf.gap_l1[n] = 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Warp-cooperative small dense products on shared memory.  C (M x N, ld ldc) (+)= opA(A) (M x K) * B (K x N).
// TA: A is stored K x M (use A^T).  Each lane owns RT x CT register tiles, accumulating over k in ascending order
// (the same summation order as the scalar reference loops).
template <int M_, int N_, int K_, int RT, int CT, bool TA, bool ACC, bool NEG>
EMPC_DI void warp_mm(double* __restrict__ C, int ldc, const double* __restrict__ A, int lda,
                     const double* __restrict__ Bm, int ldb, int lane) {
  constexpr int TM = (M_ + RT - 1) / RT, TN = (N_ + CT - 1) / CT;
  for (int tile = lane; tile < TM * TN; tile += 32) {
    const int i0 = (tile / TN) * RT, j0 = (tile % TN) * CT;
    double acc[RT][CT];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < CT; ++c) acc[r][c] = 0.0;
#pragma unroll 2
    for (int k = 0; k < K_; ++k) {
      double a[RT], bv[CT];
#pragma unroll
      for (int r = 0; r < RT; ++r) a[r] = (i0 + r < M_) ? (TA ? A[k * lda + i0 + r] : A[(i0 + r) * lda + k]) : 0.0;
#pragma unroll
      for (int c = 0; c < CT; ++c) bv[c] = (j0 + c < N_) ? Bm[k * ldb + j0 + c] : 0.0;
#pragma unroll
      for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int c = 0; c < CT; ++c) acc[r][c] += a[r] * bv[c];
    }
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < CT; ++c)
        if (i0 + r < M_ && j0 + c < N_) {
          double* p = &C[(i0 + r) * ldc + j0 + c];
          if (ACC) *p = NEG ? (*p - acc[r][c]) : (*p + acc[r][c]);
          else *p = NEG ? -acc[r][c] : acc[r][c];
        }
  }
}

template <class D>
struct BwSmem {
  static constexpr int NDX = D::NDX, NU = D::NU;
  // per-warp layout (doubles)
  static constexpr int oTile = 0;
  static constexpr int oV = oTile + D::TILE;         // Vxx' (NDX x NDX)
  static constexpr int oFxTV = oV + NDX * NDX;        // Fx^T Vxx'  (later: Vxx_t scratch)
  static constexpr int oFuTV = oFxTV + NDX * NDX;     // Fu^T Vxx'
  static constexpr int oK = oFuTV + NU * NDX;         // K (NU x NDX)
  static constexpr int oVx = oK + NU * NDX;           // Vx' (NDX)
  static constexpr int oVec = oVx + NDX;              // k(NU) Quuk(NU) fs(NDX) g(NDX) tmp(NDX)
  static constexpr int TOTAL0 = oVec + 2 * NU + 3 * NDX;
  static constexpr int TOTAL = TOTAL0 + (TOTAL0 & 1);
};

struct BwParams {
  double reg_max, reg_factor, th_gaptol;
  int force;          // phase hook: single attempt, take xreg / is_feasible from state as they are, no prologue
};

template <class D>
__global__ void __launch_bounds__(128) backward_kernel(Buffers bf, BwParams P) {
  constexpr int NDX = D::NDX, NU = D::NU, NX = D::NX;
  using S = BwSmem<D>;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int b = blockIdx.x * (blockDim.x >> 5) + wib;
  if (b >= bf.B) return;
  double* sm = smem + (size_t)wib * S::TOTAL;
  OcpState st = bf.st[b];
  if (!P.force && st.phase == PHASE_DONE) return;
  const int T = bf.T, T1 = T + 1;
  const size_t nb = (size_t)b * T1;

  // ---- prologue: SolverDDP::calcDiff tail — cost_ = sum of node costs (in node order), feasibility from the gaps ----
  if (!P.force && st.recalc) {
    double* tmp = sm;  // reuse the per-warp region as staging, in chunks
    double c = 0;
    for (int base = 0; base < T1; base += S::TOTAL) {
      const int cnt = min(S::TOTAL, T1 - base);
      for (int t = lane; t < cnt; t += 32) tmp[t] = bf.node_cost[nb + base + t];
      __syncwarp();
      if (lane == 0) for (int t = 0; t < cnt; ++t) c += tmp[t];
      __syncwarp();
    }
    st.cost = __shfl_sync(0xffffffffu, c, 0);
    if (!st.is_feasible) {
      double gi = 0, g1 = 0; int has_nan = 0;
      for (int t = lane; t < T1; t += 32) { const double a = bf.gap_inf[nb + t]; if (isnan(a)) has_nan = 1; gi = fmax(gi, a); g1 += bf.gap_l1[nb + t]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { gi = fmax(gi, __shfl_xor_sync(0xffffffffu, gi, o)); g1 += __shfl_xor_sync(0xffffffffu, g1, o); has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, o); }
      st.gap_inf = has_nan ? nan("") : gi; st.gap_l1 = g1;
      st.is_feasible = (!has_nan && gi < P.th_gaptol) ? 1 : 0;
    } else if (!st.was_feasible) {
      st.gap_inf = 0; st.gap_l1 = 0;
    }
    __syncwarp();
  }
  const int feasible = st.is_feasible;

  double* tile = sm + S::oTile;
  double* V = sm + S::oV;
  double* FxTV = sm + S::oFxTV;
  double* FuTV = sm + S::oFuTV;
  double* Kt = sm + S::oK;
  double* Vxp = sm + S::oVx;
  double* kv = sm + S::oVec;
  double* Quuk = kv + NU;
  double* fsv = Quuk + NU;
  double* gv = fsv + NDX;
  double* tmpv = gv + NDX;
  double* Fx = tile + D::oFx; double* Fu = tile + D::oFu; double* Qxx = tile + D::oLxx; double* Qxu = tile + D::oLxu;
  double* Quu = tile + D::oLuu; double* Qx = tile + D::oLx; double* Qu = tile + D::oLu;

  int failed;
  while (true) {
    failed = 0;
    const double xreg = st.xreg;
    // terminal node: Vxx = Lxx + xreg I ; Vx = Lx (+ Vxx fs)
    {
      const double* tg = bf.tiles + (nb + T) * D::TILE;
      for (int i = lane; i < NDX * NDX; i += 32) V[i] = tg[D::oLxx + i];
      for (int i = lane; i < NDX; i += 32) { Vxp[i] = tg[D::oLx + i]; fsv[i] = bf.fs[(nb + T) * NDX + i]; }
      __syncwarp();
      for (int i = lane; i < NDX; i += 32) V[i * NDX + i] += xreg;
      __syncwarp();
      for (int i = lane; i < NDX; i += 32) {
        double s = 0;
        for (int j = 0; j < NDX; ++j) s += V[i * NDX + j] * fsv[j];
        gv[i] = s;
      }
      __syncwarp();
      if (!feasible) for (int i = lane; i < NDX; i += 32) Vxp[i] += gv[i];
      __syncwarp();
      // per-node scalars: [Qu.k, k.Quuk, Vx.fs, fs.Vxx fs]
      double a = 0, c = 0;
      for (int i = lane; i < NDX; i += 32) { a += Vxp[i] * fsv[i]; c += fsv[i] * gv[i]; }
      // ordered (sequential) dot products keep the reference summation order
      if (lane == 0) {
        double s0 = 0, s1 = 0;
        for (int i = 0; i < NDX; ++i) { s0 += Vxp[i] * fsv[i]; s1 += fsv[i] * gv[i]; }
        double* ns = bf.nodesc + (nb + T) * 4;
        ns[0] = 0; ns[1] = 0; ns[2] = s0; ns[3] = s1;
      }
      (void)a; (void)c;
      for (int i = lane; i < NDX; i += 32) { bf.Vx[(nb + T) * NDX + i] = Vxp[i]; bf.g[(nb + T) * NDX + i] = gv[i]; }
      __syncwarp();
    }
    for (int t = T - 1; t >= 0; --t) {
      // stage the node tile in shared memory (coalesced)
      {
        const double2* tg = reinterpret_cast<const double2*>(bf.tiles + (nb + t) * D::TILE);
        double2* ts = reinterpret_cast<double2*>(tile);
        for (int i = lane; i < D::TILE / 2; i += 32) ts[i] = tg[i];
        for (int i = lane; i < NDX; i += 32) fsv[i] = bf.fs[(nb + t) * NDX + i];
      }
      __syncwarp();
      // FxTV = Fx^T V ; FuTV = Fu^T V
      warp_mm<NDX, NDX, NDX, 3, 3, true, false, false>(FxTV, NDX, Fx, NDX, V, NDX, lane);
      warp_mm<NU, NDX, NDX, 3, 3, true, false, false>(FuTV, NDX, Fu, NU, V, NDX, lane);
      // Qx += Fx^T Vx' ; Qu += Fu^T Vx'
      for (int i = lane; i < NDX + NU; i += 32) {
        double s = 0;
        if (i < NDX) { for (int l = 0; l < NDX; ++l) s += Fx[l * NDX + i] * Vxp[l]; Qx[i] += s; }
        else { const int ii = i - NDX; for (int l = 0; l < NDX; ++l) s += Fu[l * NU + ii] * Vxp[l]; Qu[ii] += s; }
      }
      __syncwarp();
      // Qxx += FxTV Fx ; Qxu += FxTV Fu ; Quu += FuTV Fu (+ ureg)
      warp_mm<NDX, NDX, NDX, 3, 3, false, true, false>(Qxx, NDX, FxTV, NDX, Fx, NDX, lane);
      warp_mm<NDX, NU, NDX, 3, 3, false, true, false>(Qxu, NU, FxTV, NDX, Fu, NU, lane);
      warp_mm<NU, NU, NDX, 3, 3, false, true, false>(Quu, NU, FuTV, NDX, Fu, NU, lane);
      __syncwarp();
      for (int i = lane; i < NU; i += 32) Quu[i * NU + i] += xreg;
      __syncwarp();
      // Quuk needs the un-factorised Quu: keep a copy of Quu in FuTV (free from here on)
      double* L = FuTV;
      for (int i = lane; i < NU * NU; i += 32) L[i] = Quu[i];
      __syncwarp();
      // Cholesky (right-looking; subtraction order equals the scalar left-looking loop)
      for (int j = 0; j < NU; ++j) {
        const double djj = L[j * NU + j];
        if (!(djj > 0.0)) failed = 1;
        const double d = sqrt(djj);
        __syncwarp();
        if (lane == 0) L[j * NU + j] = d;
        for (int i = j + 1 + lane; i < NU; i += 32) L[i * NU + j] = L[i * NU + j] / d;
        __syncwarp();
        const int rem = NU - j - 1;
        for (int idx = lane; idx < rem * rem; idx += 32) {
          const int i = j + 1 + idx / rem, kk = j + 1 + idx % rem;
          if (kk <= i) L[i * NU + kk] -= L[i * NU + j] * L[kk * NU + j];
        }
        __syncwarp();
      }
      if (failed) break;
      // K = Quu^-1 Qxu^T (one right-hand side per lane), k = Quu^-1 Qu
      for (int c = lane; c < NDX + 1; c += 32) {
        double rhs[NU];
        if (c < NDX) {
#pragma unroll
          for (int i = 0; i < NU; ++i) rhs[i] = Qxu[c * NU + i];
        } else {
#pragma unroll
          for (int i = 0; i < NU; ++i) rhs[i] = Qu[i];
        }
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          double s = rhs[i];
#pragma unroll
          for (int kk = 0; kk < i; ++kk) s -= L[i * NU + kk] * rhs[kk];
          rhs[i] = s / L[i * NU + i];
        }
#pragma unroll
        for (int i = NU - 1; i >= 0; --i) {
          double s = rhs[i];
#pragma unroll
          for (int kk = i + 1; kk < NU; ++kk) s -= L[kk * NU + i] * rhs[kk];
          rhs[i] = s / L[i * NU + i];
        }
        if (c < NDX) {
#pragma unroll
          for (int i = 0; i < NU; ++i) Kt[i * NDX + c] = rhs[i];
        } else {
#pragma unroll
          for (int i = 0; i < NU; ++i) kv[i] = rhs[i];
        }
      }
      __syncwarp();
      // Quuk = Quu k
      for (int i = lane; i < NU; i += 32) {
        double s = 0;
        for (int j = 0; j < NU; ++j) s += Quu[i * NU + j] * kv[j];
        Quuk[i] = s;
      }
      __syncwarp();
      // Vx = Qx + K^T Quuk - 2 K^T Qu ; Vxx = Qxx - Qxu K
      for (int i = lane; i < NDX; i += 32) {
        double s1 = 0, s2 = 0;
        for (int j = 0; j < NU; ++j) { s1 += Kt[j * NDX + i] * Quuk[j]; s2 += Kt[j * NDX + i] * Qu[j]; }
        tmpv[i] = Qx[i] + s1 - 2 * s2;
      }
      warp_mm<NDX, NDX, NU, 3, 3, false, true, true>(Qxx, NDX, Qxu, NU, Kt, NDX, lane);
      __syncwarp();
      // symmetrise + xreg -> V
      for (int idx = lane; idx < NDX * NDX; idx += 32) {
        const int i = idx / NDX, j = idx - i * NDX;
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        double a = 0.5 * (Qxx[lo * NDX + hi] + Qxx[hi * NDX + lo]);
        if (i == j) a += xreg;
        V[idx] = a;
      }
      __syncwarp();
      for (int i = lane; i < NDX; i += 32) {
        double s = 0;
        for (int j = 0; j < NDX; ++j) s += V[i * NDX + j] * fsv[j];
        gv[i] = s;
      }
      __syncwarp();
      for (int i = lane; i < NDX; i += 32) Vxp[i] = feasible ? tmpv[i] : (tmpv[i] + gv[i]);
      __syncwarp();
      // NaN guard ("backward_error")
      {
        int bad = 0;
        for (int i = lane; i < NDX * NDX; i += 32) if (isnan(V[i])) bad = 1;
        for (int i = lane; i < NDX; i += 32) if (isnan(Vxp[i])) bad = 1;
        if (__any_sync(0xffffffffu, bad)) { failed = 1; break; }
      }
      // outputs
      {
        double* Kg = bf.K + ((size_t)b * T + t) * NU * NDX;
        for (int i = lane; i < NU * NDX; i += 32) Kg[i] = Kt[i];
        double* kg = bf.k + ((size_t)b * T + t) * NU;
        for (int i = lane; i < NU; i += 32) kg[i] = kv[i];
        for (int i = lane; i < NDX; i += 32) { bf.Vx[(nb + t) * NDX + i] = Vxp[i]; bf.g[(nb + t) * NDX + i] = gv[i]; }
        if (lane == 0) {
          double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
          for (int i = 0; i < NU; ++i) { s0 += Qu[i] * kv[i]; s1 += kv[i] * Quuk[i]; }
          for (int i = 0; i < NDX; ++i) { s2 += Vxp[i] * fsv[i]; s3 += fsv[i] * gv[i]; }
          double* ns = bf.nodesc + (nb + t) * 4;
          ns[0] = s0; ns[1] = s1; ns[2] = s2; ns[3] = s3;
        }
      }
      __syncwarp();
    }
    failed = __any_sync(0xffffffffu, failed);
    if (!failed || P.force) break;
    // computeDirection threw: recalcDiff = false; increaseRegularization(); give up at reg_max (src/sbfddp.cpp:245-253)
    st.xreg *= P.reg_factor;
    if (st.xreg > P.reg_max) st.xreg = P.reg_max;
    if (st.xreg == P.reg_max) break;
  }
  st.bw_fail = failed ? 1 : 0;
  // SolverFDDP::updateExpectedImprovement / expectedImprovementDDP: ordered sums over the nodes
  if (!failed) {
    double* tmp = sm;
    double dg = 0, dq = 0, dg0 = 0, dq0 = 0;
    if (!feasible && lane == 0) { dg -= bf.nodesc[(nb + T) * 4 + 2]; dq += bf.nodesc[(nb + T) * 4 + 3]; }
    constexpr int CH = S::TOTAL / 4;
    for (int base = 0; base < T; base += CH) {
      const int cnt = min(CH, T - base);
      for (int i = lane; i < cnt * 4; i += 32) tmp[i] = bf.nodesc[(nb + base) * 4 + i];
      __syncwarp();
      if (lane == 0) {
        for (int t = 0; t < cnt; ++t) {
          dg += tmp[t * 4 + 0]; dq -= tmp[t * 4 + 1];
          dg0 += tmp[t * 4 + 0]; dq0 -= tmp[t * 4 + 1];
          if (!feasible) { dg -= tmp[t * 4 + 2]; dq += tmp[t * 4 + 3]; }
        }
      }
      __syncwarp();
    }
    if (lane == 0) { st.dg = dg; st.dq = dq; st.dg0 = dg0; st.dq0 = dq0; }
  }
  if (lane == 0) bf.st[b] = st;
}

// ---------------------------------------------------------------------------------------------------------------------
struct RoParams {
  int force, force_feasible, force_ddp;
  double force_smooth;
};

template <class D>
__global__ void __launch_bounds__(128) rollout_kernel(Buffers bf, RoParams P) {
  constexpr int NX = D::NX, NDX = D::NDX, NU = D::NU;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= bf.B * EMPC_N_ALPHAS) return;
  const int b = n / EMPC_N_ALPHAS, ai = n - b * EMPC_N_ALPHAS;
  const OcpState st = bf.st[b];
  if (!P.force && (st.phase == PHASE_DONE || st.bw_fail)) return;
  const int ddp = P.force ? P.force_ddp : (st.phase == PHASE_DDP);
  const int feasible = P.force ? P.force_feasible : st.is_feasible;
  const double smooth = P.force ? P.force_smooth : st.smooth;
  const double alpha = 1.0 / (double)(1 << ai);
  const DevModel& M = *bf.model;
  const int T = bf.T, T1 = T + 1;
  const size_t nb = (size_t)b * T1;
  const size_t trial = (size_t)ai * bf.B + b;
  double* xs_try = bf.xs_try + trial * T1 * NX;
  double* us_try = bf.us_try + trial * T * NU;
  const bool plain = ddp || feasible || ai == 0;

  double xn[NX];  // running state (xnext)
  if (ddp) {
#pragma unroll
    for (int i = 0; i < NX; ++i) xn[i] = bf.xs_try0[(size_t)b * NX + i];
  } else {
#pragma unroll
    for (int i = 0; i < NX; ++i) xn[i] = bf.x0[(size_t)b * NX + i];
  }
  double cost_try = 0, dv = 0;
  int ok = 1;
  for (int t = 0; t <= T; ++t) {
    double xt[NX];
    if (plain) {
#pragma unroll
      for (int i = 0; i < NX; ++i) xt[i] = xn[i];
    } else {
      double gap[NDX];
#pragma unroll
      for (int i = 0; i < NDX; ++i) gap[i] = bf.fs[(nb + t) * NDX + i] * (alpha - 1);
      state_integrate<D>(xn, gap, xt);
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) xs_try[(size_t)t * NX + i] = xt[i];
    double x0t[NX], dx[NDX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x0t[i] = bf.xs[(nb + t) * NX + i];
    state_diff<D>(x0t, xt, dx);
    if (!ddp && !feasible) {
      // dv -= fs . Vxx diff(xs_try, xs)  ==  + (Vxx fs) . diff(xs, xs_try)   (Vxx symmetric)
      double s = 0;
#pragma unroll
      for (int i = 0; i < NDX; ++i) s += bf.g[(nb + t) * NDX + i] * dx[i];
      dv += s;
    }
    double u[NU];
    const int costset = bf.node_costset[bf.ocp_map[b] * T1 + t];
    NodeData<D> nd;
    double c;
    if (t < T) {
      const double* Kg = bf.K + ((size_t)b * T + t) * NU * NDX;
      const double* kg = bf.k + ((size_t)b * T + t) * NU;
      const double* ug = bf.us + ((size_t)b * T + t) * NU;
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        double kd = 0;
#pragma unroll
        for (int j = 0; j < NDX; ++j) kd += Kg[i * NDX + j] * dx[j];
        u[i] = ug[i] - kg[i] * alpha - kd;
        us_try[(size_t)t * NU + i] = u[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < NU; ++i) u[i] = 0.0;
    }
    node_calc<D>(M, bf.ct, costset, smooth, xt, u, nd, xn, c);
    cost_try += c;
    if (isnan(cost_try)) { ok = 0; break; }
    if (t < T) {
      bool bad = false;
#pragma unroll
      for (int i = 0; i < NX; ++i) bad |= isnan(xn[i]);
      if (bad) { ok = 0; break; }
    }
  }
  bf.cost_try[n] = cost_try;
  bf.dv[n] = dv;
  bf.ok[n] = ok;
}

// ---------------------------------------------------------------------------------------------------------------------
struct DecideParams {
  empc_solver_params_t P;
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

template <class D>
__global__ void __launch_bounds__(128) decide_kernel(Buffers bf, DecideParams dp) {
  constexpr int NX = D::NX, NU = D::NU;
  const int b = blockIdx.x;
  __shared__ int s_acc, s_last;
  const empc_solver_params_t& P = dp.P;
  const int T = bf.T, T1 = T + 1;
  if (threadIdx.x == 0) {
    OcpState st = bf.st[b];
    int acc = -1, last = -1;
    if (st.phase != PHASE_DONE) {
      if (st.bw_fail) {
        st.bw_fail = 0;
        end_inner_solve(st, P);  // computeDirection gave up at reg_max: the inner solve returns false
      } else {
        const int ddp = st.phase == PHASE_DDP;
        for (int n = 0; n < EMPC_N_ALPHAS; ++n) {
          st.steplength = 1.0 / (double)(1 << n);
          last = n;
          if (!bf.ok[b * EMPC_N_ALPHAS + n]) continue;  // "forward_error": try the next step length
          const double cost_try = bf.cost_try[b * EMPC_N_ALPHAS + n];
          const double dV = st.cost - cost_try;
          double d0, d1;
          if (ddp) { d0 = st.dg0; d1 = st.dq0; }
          else {
            const double dv = st.is_feasible ? 0.0 : bf.dv[b * EMPC_N_ALPHAS + n];
            d0 = st.dg + dv; d1 = st.dq - 2 * dv;
          }
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
            break;
          }
        }
        st.recalc = acc >= 0 ? 1 : 0;
        st.accepted = acc;
        bool ended = false;
        if (st.steplength > P.th_stepdec) decrease_reg(st, P);
        if (st.steplength <= P.th_stepinc) {
          increase_reg(st, P);
          if (st.xreg == P.reg_max) { end_inner_solve(st, P); ended = true; }
        }
        if (!ended) {
          // fork stop rules, inferred: StopCriteriaCostReduction / StopTestGaps
          st.stop = fabs(st.cost_prev - st.cost);
          bool stop_now;
          if (ddp) stop_now = st.was_feasible && st.stop < st.th_stop;
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
      if (st.phase != PHASE_DONE) { atomicAdd(bf.n_active, 1); if (st.recalc) atomicAdd(bf.n_active + 1, 1); }
      bf.st[b] = st;
    }
    s_acc = acc; s_last = last;
  }
  __syncthreads();
  const int acc = s_acc, last = s_last;
  if (last >= 0) {
    // xs_try_[0] persists into the DDP phase from the last forwardPass executed (src/sbfddp.cpp:430)
    const size_t trial = (size_t)last * bf.B + b;
    for (int i = threadIdx.x; i < NX; i += blockDim.x) bf.xs_try0[(size_t)b * NX + i] = bf.xs_try[trial * T1 * NX + i];
  }
  if (acc >= 0) {
    const size_t trial = (size_t)acc * bf.B + b;
    const double* xsrc = bf.xs_try + trial * T1 * NX;
    double* xdst = bf.xs + (size_t)b * T1 * NX;
    for (int i = threadIdx.x; i < T1 * NX; i += blockDim.x) xdst[i] = xsrc[i];
    const double* usrc = bf.us_try + trial * T * NU;
    double* udst = bf.us + (size_t)b * T * NU;
    for (int i = threadIdx.x; i < T * NU; i += blockDim.x) udst[i] = usrc[i];
  }
}

// state initialisation at the start of solve() (src/sbfddp.cpp:198-210)
__global__ void init_state_kernel(Buffers bf, empc_solver_params_t P, int is_feasible_arg, int nx) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bf.B) return;
  OcpState st;
  st.smooth = P.smooth_init; st.smooth_next = P.smooth_init;
  st.convergence = P.convergence_init; st.th_stop = P.convergence_init;
  st.xreg = P.reg_init; st.cost = 0; st.cost_prev = bf.st[b].cost_prev; st.stop = 0; st.steplength = 1;
  st.dg = st.dq = st.dg0 = st.dq0 = 0; st.gap_inf = 0; st.gap_l1 = 0;
  st.iter = 0; st.total_iters = 0; st.is_feasible = 0; st.was_feasible = 0; st.recalc = 1; st.bw_fail = 0;
  st.iters_out = 0; st.accepted = -1; st.pad_ = 0;
  (void)is_feasible_arg;  // solveFDDP(maxiter, false, ...) overrides the caller's flag (src/sbfddp.cpp:210,230)
  if (P.convergence_init >= P.convergence_stop) st.phase = PHASE_FDDP;
  else { st.phase = is_feasible_arg ? PHASE_DONE : PHASE_DDP; st.is_feasible = is_feasible_arg; }
  if (st.phase == PHASE_DONE) st.iters_out = -1;
  bf.st[b] = st;
  for (int i = 0; i < nx; ++i) bf.xs_try0[(size_t)b * nx + i] = bf.x0[(size_t)b * nx + i];  // xs_try_[0] = x0 (:198)
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
}

}  // namespace empc
