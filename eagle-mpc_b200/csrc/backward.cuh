// backward.cuh — Riccati backward pass, one thread block per OCP (included from kernels.cuh inside namespace empc).
//
// Replaces crocoddyl::SolverDDP::backwardPass + computeGains and the regularisation retry loop of
// SolverSbFDDP::solveFDDP/solveDDP (src/sbfddp.cpp:242-255, :330-343), plus SolverFDDP::updateExpectedImprovement /
// SolverSbFDDP::expectedImprovementDDP (src/sbfddp.cpp:256, :395-408).
//
// Layout: the node tile (Fx|Fu|Lxx|Lxu|Luu|Lx|Lu) is staged in shared memory and turned in place into
// Qxx|Qxu|Quu|Qx|Qu; Vxx' (the value Hessian of the next node) stays resident in shared memory for the whole sweep.
// The next node's tile is prefetched from HBM into registers while the current node is processed.  Dense products
// use 2x3 register tiles per thread and accumulate over k in ascending order (the reference's summation order).
#pragma once

template <class D>
struct BwCfg {
  static constexpr int NDX = D::NDX, NU = D::NU;
  static constexpr int THREADS = (NDX <= 18) ? 64 : 128;
#ifndef EMPC_BW_MINB
#define EMPC_BW_MINB 8
#endif
  static constexpr int MIN_BLOCKS = (NDX <= 18) ? EMPC_BW_MINB : EMPC_BW_MINB / 2;  // caps registers at 128/thread so shared memory, not registers, bounds occupancy
  static constexpr int oTile = 0;
  static constexpr int oV = oTile + D::TILE;          // Vxx' (NDX x NDX)
  static constexpr int oFxTV = oV + NDX * NDX;         // Fx^T Vxx'
  static constexpr int oFuTV = oFxTV + NDX * NDX;      // Fu^T Vxx'  (later: Cholesky factor of Quu)
  static constexpr int oK = oFuTV + NU * NDX;          // K (NU x NDX)
  static constexpr int oVx = oK + NU * NDX;            // Vx' (NDX)
  static constexpr int oVec = oVx + NDX;               // k(NU) Quuk(NU) fs(NDX) g(NDX) tmp(NDX)
  static constexpr int TOTAL0 = oVec + 2 * NU + 3 * NDX;
  static constexpr int TOTAL = TOTAL0 + (TOTAL0 & 1);
  static constexpr int PREF = (D::TILE / 2 + THREADS - 1) / THREADS;  // double2 registers per thread for the prefetch
};

struct BwParams {
  double reg_max, reg_factor, th_gaptol;
  int force;  // phase hook: single attempt, xreg / is_feasible taken from the state as they are, no prologue
};

// C (M x N, ld ldc) (+/-)= opA(A) (M x K) * B (K x N) over threads [tid, tid+nth, ...]; TA: A stored K x M.
template <int M_, int N_, int K_, int RT, int CT, bool TA, bool ACC, bool NEG>
EMPC_DI void cta_mm(double* __restrict__ C, int ldc, const double* __restrict__ A, int lda, const double* __restrict__ Bm,
                    int ldb, int tid, int nth) {
  constexpr int TM = (M_ + RT - 1) / RT, TN = (N_ + CT - 1) / CT;
  for (int tile = tid; tile < TM * TN; tile += nth) {
    const int i0 = (tile / TN) * RT, j0 = (tile % TN) * CT;
    double acc[RT][CT];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < CT; ++c) acc[r][c] = 0.0;
    // clamp the ragged edge instead of branching inside the k loop (the clamped lanes' results are discarded)
    int ia[RT], jb[CT];
#pragma unroll
    for (int r = 0; r < RT; ++r) ia[r] = (i0 + r < M_) ? i0 + r : M_ - 1;
#pragma unroll
    for (int c = 0; c < CT; ++c) jb[c] = (j0 + c < N_) ? j0 + c : N_ - 1;
#pragma unroll 6
    for (int k = 0; k < K_; ++k) {
      double a[RT], bv[CT];
#pragma unroll
      for (int r = 0; r < RT; ++r) a[r] = TA ? A[k * lda + ia[r]] : A[ia[r] * lda + k];
#pragma unroll
      for (int c = 0; c < CT; ++c) bv[c] = Bm[k * ldb + jb[c]];
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
__global__ void __launch_bounds__(BwCfg<D>::THREADS, BwCfg<D>::MIN_BLOCKS) backward_kernel(Buffers bf, BwParams P) {
  constexpr int NDX = D::NDX, NU = D::NU;
  using S = BwCfg<D>;
  constexpr int NT = S::THREADS;
  extern __shared__ double sm[];
  __shared__ int s_flag;
  __shared__ double s_red[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = bf.b0 + blockIdx.x;
  OcpState st = bf.st[b];
  if (!P.force && st.phase == PHASE_DONE) return;
  const int T = bf.T, T1 = T + 1;
  const size_t nb = (size_t)b * T1;

  // ---- prologue: tail of SolverDDP::calcDiff — cost_ = sum of node costs in node order, feasibility from the gaps ----
  if (!P.force && st.recalc) {
    double c = 0;
    for (int base = 0; base < T1; base += S::TOTAL) {
      const int cnt = min(S::TOTAL, T1 - base);
      for (int t = tid; t < cnt; t += NT) sm[t] = bf.node_cost[nb + base + t];
      __syncthreads();
      if (tid == 0) for (int t = 0; t < cnt; ++t) c += sm[t];
      __syncthreads();
    }
    if (tid == 0) s_red[0] = c;
    if (!st.is_feasible) {
      double gi = 0, g1 = 0; int has_nan = 0;
      for (int t = tid; t < T1; t += NT) { const double a = bf.gap_inf[nb + t]; if (isnan(a)) has_nan = 1; gi = fmax(gi, a); g1 += bf.gap_l1[nb + t]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { gi = fmax(gi, __shfl_xor_sync(0xffffffffu, gi, o)); g1 += __shfl_xor_sync(0xffffffffu, g1, o); has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, o); }
      double* red = sm;  // [warp][3]
      if (lane == 0) { red[warp * 3] = gi; red[warp * 3 + 1] = g1; red[warp * 3 + 2] = has_nan ? 1.0 : 0.0; }
      __syncthreads();
      if (tid == 0) {
        double a = 0, l = 0, n = 0;
        for (int w = 0; w < NT / 32; ++w) { a = fmax(a, red[w * 3]); l += red[w * 3 + 1]; n += red[w * 3 + 2]; }
        s_red[1] = a; s_red[2] = l; s_red[3] = n;
      }
    }
    __syncthreads();
    st.cost = s_red[0];
    if (!st.is_feasible) {
      const bool has_nan = s_red[3] != 0.0;
      st.gap_inf = has_nan ? nan("") : s_red[1]; st.gap_l1 = s_red[2];
      st.is_feasible = (!has_nan && s_red[1] < P.th_gaptol) ? 1 : 0;
    } else if (!st.was_feasible) {
      st.gap_inf = 0; st.gap_l1 = 0;
    }
    __syncthreads();
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
  double2* tile2 = reinterpret_cast<double2*>(tile);

  int failed;
  while (true) {
    failed = 0;
    if (tid == 0) s_flag = 0;
    const double xreg = st.xreg;
    // ---- terminal node: Vxx = Lxx + xreg I ; Vx = Lx (+ Vxx fs) ----
    {
      const double* tg = bf.tiles + (nb + T) * D::TILE;
      for (int i = tid; i < NDX * NDX; i += NT) V[i] = tg[D::oLxx + i] + ((i / NDX == i % NDX) ? xreg : 0.0);
      for (int i = tid; i < NDX; i += NT) { Vxp[i] = tg[D::oLx + i]; fsv[i] = bf.fs[(nb + T) * NDX + i]; }
      __syncthreads();
      for (int i = tid; i < NDX; i += NT) {
        double s = 0;
        for (int j = 0; j < NDX; ++j) s += V[i * NDX + j] * fsv[j];
        gv[i] = s;
      }
      __syncthreads();
      if (!feasible) for (int i = tid; i < NDX; i += NT) Vxp[i] += gv[i];
      __syncthreads();
      if (tid == 0) {
        double s0 = 0, s1 = 0;
        for (int i = 0; i < NDX; ++i) { s0 += Vxp[i] * fsv[i]; s1 += fsv[i] * gv[i]; }
        double* ns = bf.nodesc + (nb + T) * 4;
        ns[0] = 0; ns[1] = 0; ns[2] = s0; ns[3] = s1;
      }
      for (int i = tid; i < NDX; i += NT) { bf.Vx[(nb + T) * NDX + i] = Vxp[i]; bf.g[(nb + T) * NDX + i] = gv[i]; }
      __syncthreads();
      // first running tile straight into shared memory
      const double2* tg2 = reinterpret_cast<const double2*>(bf.tiles + (nb + T - 1) * D::TILE);
      for (int i = tid; i < D::TILE / 2; i += NT) tile2[i] = tg2[i];
      for (int i = tid; i < NDX; i += NT) fsv[i] = bf.fs[(nb + T - 1) * NDX + i];
      __syncthreads();
    }
    for (int t = T - 1; t >= 0; --t) {
      // prefetch the next node's tile (t-1) into registers; it lands in shared memory at the end of this step
      double2 pre[S::PREF];
      double pre_fs = 0.0;
      if (t > 0) {
        const double2* tg2 = reinterpret_cast<const double2*>(bf.tiles + (nb + t - 1) * D::TILE);
#pragma unroll
        for (int r = 0; r < S::PREF; ++r) { const int i = tid + r * NT; if (i < D::TILE / 2) pre[r] = tg2[i]; }
        if (tid < NDX) pre_fs = bf.fs[(nb + t - 1) * NDX + tid];
      }
      // FxTV = Fx^T V ; FuTV = Fu^T V ; Qx += Fx^T Vx' ; Qu += Fu^T Vx'
      cta_mm<NDX, NDX, NDX, 2, 3, true, false, false>(FxTV, NDX, Fx, NDX, V, NDX, tid, NT);
      cta_mm<NU, NDX, NDX, 2, 3, true, false, false>(FuTV, NDX, Fu, NU, V, NDX, tid, NT);
      for (int i = tid; i < NDX + NU; i += NT) {
        double s = 0;
        if (i < NDX) { for (int l = 0; l < NDX; ++l) s += Fx[l * NDX + i] * Vxp[l]; Qx[i] += s; }
        else { const int ii = i - NDX; for (int l = 0; l < NDX; ++l) s += Fu[l * NU + ii] * Vxp[l]; Qu[ii] += s; }
      }
      __syncthreads();
      // Warp 0: Qxu += FxTV Fu ; Quu += FuTV Fu (+ ureg) ; Cholesky ; gains.   Other warps: Qxx += FxTV Fu meanwhile
      // (Qxx is not needed by the factorisation).
      double* L = FuTV + NU * NDX - NU * NU - NU;  // tail of the FuTV area; FuTV itself is consumed before L is written
      double* Linv = L + NU * NU;
      if (warp != 0) {
        cta_mm<NDX, NDX, NDX, 2, 3, false, true, false>(Qxx, NDX, FxTV, NDX, Fx, NDX, tid - 32, NT - 32);
      } else {
        cta_mm<NDX, NU, NDX, 2, 3, false, true, false>(Qxu, NU, FxTV, NDX, Fu, NU, lane, 32);
        cta_mm<NU, NU, NDX, 2, 3, false, true, false>(Quu, NU, FuTV, NDX, Fu, NU, lane, 32);
        __syncwarp();
        for (int i = lane; i < NU; i += 32) Quu[i * NU + i] += xreg;
        __syncwarp();
        for (int i = lane; i < NU * NU; i += 32) L[i] = Quu[i];
        __syncwarp();
        int bad = 0;
        // right-looking Cholesky; the subtraction order equals the scalar left-looking loop of the reference LLT
#pragma unroll 1
        for (int j = 0; j < NU; ++j) {
          const double djj = L[j * NU + j];
          if (!(djj > 0.0)) bad = 1;
          const double d = sqrt(djj);
          const double dinv = 1.0 / d;
          __syncwarp();
          if (lane == 0) { L[j * NU + j] = d; Linv[j] = dinv; }
          for (int i = j + 1 + lane; i < NU; i += 32) L[i * NU + j] = L[i * NU + j] * dinv;
          __syncwarp();
          for (int idx = lane; idx < NU * NU; idx += 32) {
            const int i = idx / NU, kk = idx - i * NU;
            if (kk > j && kk <= i) L[idx] -= L[i * NU + j] * L[kk * NU + j];
          }
          __syncwarp();
        }
        if (bad) { if (lane == 0) s_flag = 1; }
        else {
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
              rhs[i] = s * Linv[i];
            }
#pragma unroll
            for (int i = NU - 1; i >= 0; --i) {
              double s = rhs[i];
#pragma unroll
              for (int kk = i + 1; kk < NU; ++kk) s -= L[kk * NU + i] * rhs[kk];
              rhs[i] = s * Linv[i];
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
          for (int i = lane; i < NU; i += 32) {  // Quuk = Quu k
            double s = 0;
            for (int j = 0; j < NU; ++j) s += Quu[i * NU + j] * kv[j];
            Quuk[i] = s;
          }
        }
      }
      __syncthreads();
      if (s_flag) { failed = 1; break; }
      // Vx = Qx + K^T Quuk - 2 K^T Qu ; Vxx = Qxx - Qxu K
      for (int i = tid; i < NDX; i += NT) {
        double s1 = 0, s2 = 0;
        for (int j = 0; j < NU; ++j) { s1 += Kt[j * NDX + i] * Quuk[j]; s2 += Kt[j * NDX + i] * Qu[j]; }
        tmpv[i] = Qx[i] + s1 - 2 * s2;
      }
      cta_mm<NDX, NDX, NU, 2, 3, false, true, true>(Qxx, NDX, Qxu, NU, Kt, NDX, tid, NT);
      __syncthreads();
      // symmetrise + xreg -> V, NaN guard ("backward_error")
      {
        int bad = 0;
        for (int idx = tid; idx < NDX * NDX; idx += NT) {
          const int i = idx / NDX, j = idx - i * NDX;
          const int lo = i < j ? i : j, hi = i < j ? j : i;
          double a = 0.5 * (Qxx[lo * NDX + hi] + Qxx[hi * NDX + lo]);
          if (i == j) a += xreg;
          if (isnan(a)) bad = 1;
          V[idx] = a;
        }
        if (bad) s_flag = 1;
      }
      __syncthreads();
      for (int i = tid; i < NDX; i += NT) {
        double s = 0;
        for (int j = 0; j < NDX; ++j) s += V[i * NDX + j] * fsv[j];
        gv[i] = s;
        const double vx = feasible ? tmpv[i] : (tmpv[i] + s);
        if (isnan(vx)) s_flag = 1;
        Vxp[i] = vx;
      }
      __syncthreads();
      if (s_flag) { failed = 1; break; }
      // outputs
      {
        double* Kg = bf.K + ((size_t)b * T + t) * NU * NDX;
        for (int i = tid; i < NU * NDX; i += NT) Kg[i] = Kt[i];
        double* kg = bf.k + ((size_t)b * T + t) * NU;
        for (int i = tid; i < NU; i += NT) kg[i] = kv[i];
        for (int i = tid; i < NDX; i += NT) { bf.Vx[(nb + t) * NDX + i] = Vxp[i]; bf.g[(nb + t) * NDX + i] = gv[i]; }
        if (tid == NT - 1) {
          double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
          for (int i = 0; i < NU; ++i) { s0 += Qu[i] * kv[i]; s1 += kv[i] * Quuk[i]; }
          for (int i = 0; i < NDX; ++i) { s2 += Vxp[i] * fsv[i]; s3 += fsv[i] * gv[i]; }
          double* ns = bf.nodesc + (nb + t) * 4;
          ns[0] = s0; ns[1] = s1; ns[2] = s2; ns[3] = s3;
        }
      }
      __syncthreads();
      if (t > 0) {
#pragma unroll
        for (int r = 0; r < S::PREF; ++r) { const int i = tid + r * NT; if (i < D::TILE / 2) tile2[i] = pre[r]; }
        if (tid < NDX) fsv[tid] = pre_fs;
      }
      __syncthreads();
    }
    if (!failed || P.force) break;
    // computeDirection threw: recalcDiff = false; increaseRegularization(); give up at reg_max (src/sbfddp.cpp:245-253)
    st.xreg *= P.reg_factor;
    if (st.xreg > P.reg_max) st.xreg = P.reg_max;
    if (st.xreg == P.reg_max) break;
    __syncthreads();
  }
  st.bw_fail = failed ? 1 : 0;
  // SolverFDDP::updateExpectedImprovement / expectedImprovementDDP: ordered sums over the nodes
  if (!failed) {
    double dg = 0, dq = 0, dg0 = 0, dq0 = 0;
    __syncthreads();
    if (!feasible && tid == 0) { dg -= bf.nodesc[(nb + T) * 4 + 2]; dq += bf.nodesc[(nb + T) * 4 + 3]; }
    constexpr int CH = S::TOTAL / 4;
    for (int base = 0; base < T; base += CH) {
      const int cnt = min(CH, T - base);
      for (int i = tid; i < cnt * 4; i += NT) sm[i] = bf.nodesc[(nb + base) * 4 + i];
      __syncthreads();
      if (tid == 0) {
        for (int t = 0; t < cnt; ++t) {
          dg += sm[t * 4 + 0]; dq -= sm[t * 4 + 1];
          dg0 += sm[t * 4 + 0]; dq0 -= sm[t * 4 + 1];
          if (!feasible) { dg -= sm[t * 4 + 2]; dq += sm[t * 4 + 3]; }
        }
      }
      __syncthreads();
    }
    if (tid == 0) { st.dg = dg; st.dq = dq; st.dg0 = dg0; st.dq0 = dq0; }
  }
  if (tid == 0) bf.st[b] = st;
}
