// backward.cuh — Riccati backward pass, ONE WARP per OCP, dense products on the FP64 tensor cores (DMMA).
// (included from kernels.cuh inside namespace empc)
//
// Replaces crocoddyl::SolverDDP::backwardPass + computeGains and the regularisation retry loop of
// SolverSbFDDP::solveFDDP/solveDDP (src/sbfddp.cpp:242-255, :330-343), plus SolverFDDP::updateExpectedImprovement /
// SolverSbFDDP::expectedImprovementDDP (src/sbfddp.cpp:256, :395-408).
//
// Per node (n = ndx, m = nu):  FxTV = Fx^T V', FuTV = Fu^T V', Qxx = Lxx + FxTV Fx, Qxu = Lxu + FxTV Fu,
// Quu = Luu + FuTV Fu + ureg I, Qx = Lx + Fx^T Vx', Qu = Lu + Fu^T Vx', LLT(Quu), K = Quu^-1 Qxu^T, k = Quu^-1 Qu,
// Vx = Qx + K^T Quu k - 2 K^T Qu (+ Vxx fs), Vxx = sym(Qxx - Qxu K) + xreg I.
//
// The dense products are the only GEMM-shaped work on the SbFDDP path (18x18x18 for flying_arm_3).  They run as
// mma.sync.m8n8k4.f64 (SASS DMMA) on 8x8 tiles with zero-padded operands: one DMMA replaces 8 DFMA warp-instructions and
// needs two 8-byte fragment elements per lane instead of ~7 (the FP64 tensor rate equals the vector rate on B200, so the
// gain is issue slots and shared-memory traffic, not FLOPs).  Fx and Fu are treated as one packed operand F = [Fx | Fu]
// (n x (n + m)), so that F^T V = [FxTV ; FuTV] and the symmetric [[Qxx, Qxu], [Qux, Quu]] = (F^T V) F each take one k-loop
// and only the upper tiles of the latter are computed: 128 DMMAs per node instead of 197 for six separate padded products.
//
// What bounds the kernel (ncu, round 2: profiles/r2_backward.md): the LSU data pipe — shared-memory wavefronts, 26 % of
// them bank-conflict replays in the round-1 layout — not occupancy (14 resident warps per SM were SLOWER than 8) and not
// HBM or the FP64 pipes.  The kernel is therefore organised around shared-memory wavefronts per node:
//   * F never touches shared memory.  In both products it is read with the same fragment pattern (element
//     (4 ks + c, 8 i + r)), so each lane loads its KN x PT fragment elements of [Fx | Fu] straight from HBM into
//     registers — issued as soon as the previous node's F is dead, in flight during the factorisation — and the same
//     registers feed F^T V (as the transposed A operand) and (F^T V) F (as the B operand).
//   * Vx' rides along as column n of the V' operand: F^T [V' | Vx'] gives F^T V and the Qx / Qu updates in the same
//     DMMAs (the tile holding column n is computed anyway).
//   * every matrix in shared memory has leading dimension 24 with an XOR swizzle of the column index by row bit 1
//     (element (r, c) at r * 24 + (c ^ ((r & 2) << 1))): conflict-free for the C-layout 16-byte stores, for the B /
//     transposed-A fragment loads (4 rows x 8 columns) AND for the row-major A fragment loads (8 rows x 4 columns); the
//     plain leading dimensions of round 1 (24 / 20) made half of those two-way conflicts.
//   * Qxu is kept transposed (Qux, m x n): the gain solves read their right-hand sides and the product Qxu K its A
//     operand without conflicts from the same array.
//   * the Cholesky factor is read row-wise only (row-oriented substitutions), as 16-byte broadcasts.
//   * Qxx lives in accumulator registers from Lxx (HBM -> fragments) to the symmetrised Vxx.
// Everything is warp-cooperative with __syncwarp only: no block-level barrier anywhere on the sweep.
#pragma once

template <class D>
struct BwCfg {
  static constexpr int n = D::NDX, m = D::NU;
  static constexpr int NT = (n + 7) / 8, MT = (m + 7) / 8;      // 8-wide tiles over n, m
  static constexpr int NTV = (n + 8) / 8;                        // ... over the n + 1 columns of [V' | Vx']
  static constexpr int KN = (n + 3) / 4, KM = (m + 3) / 4;      // 4-deep k-steps over n, m
  static constexpr int NP = 8 * NT, MP = 8 * MT, KNP = 4 * KN, KMP = 4 * KM;
  static constexpr int LD = 24;                                  // leading dimension of every swizzled matrix
  static_assert(8 * NTV <= LD && KNP <= LD && NP <= LD, "operand wider than the leading dimension");
  // packed operand F = [Fx | Fu] (n x (n + m))
  static constexpr int PW = n + m, PT = (PW + 7) / 8, PP = 8 * PT;
  static constexpr int ROWS_V = KNP;                             // rows of V read as B fragments (stores beyond are masked)
  static constexpr int oV = 0;                                   // ROWS_V x LD    [Vxx' | Vx'] (Vxx' symmetric)
  // (Qxx never touches shared memory: Lxx is loaded from HBM straight into the accumulator fragments and
  //  Qxx - Qxu K is symmetrised in registers)
  static constexpr int oFTV = oV + ROWS_V * LD;                  // PP x LD        F^T V = [FxTV ; FuTV]
  // (F^T V keeps its PW live rows; the A-fragment loads of the last tile row clamp their row index)
  static constexpr int oQux = oFTV + PW * LD;                    // KMP x LD       Qux = Qxu^T, zero padded
  static constexpr int oK = oQux + KMP * LD;                     // KMP x LD       gains K (m x n), zero padded
  static constexpr int LM = m + (m & 1);                         // row stride of Quu and of its Cholesky factor (even)
  static constexpr int oQuu = oK + KMP * LD;                     // m x LM         Quu
  static constexpr int oL = oQuu + m * LM;                       // m x LM         Cholesky factor L (rows), L^T (rows), LM reciprocal pivots
  static constexpr int oVec = oL + 2 * m * LM + LM;
  // vectors: slots of NP + 2 doubles (zero beyond the vector's length: they are DMMA operands of the dot products at the
  // end of a node), skewed so that the same index of different vectors falls into different banks
  static constexpr int SLOT = (KNP > n ? KNP : n) + 2;
  static constexpr int vQx = 0, vQu = SLOT, vVx = 2 * SLOT, vFs = 3 * SLOT, vG = 4 * SLOT, vKv = 5 * SLOT,
                       vQuuk = 6 * SLOT, vTmp = 7 * SLOT, vLuu = 8 * SLOT, VEC = 9 * SLOT;
  static constexpr int TOTAL0 = oVec + VEC;
  static constexpr int TOTAL = TOTAL0 + (TOTAL0 & 1);
  // register prefetch (one node ahead) of the small cost blocks: diag(Luu) (m), Lx | Lu (n + m, contiguous) + fs.  Lxu is
  // identically zero and Luu diagonal for every cost the reference's factories build (state / control / frame
  // residuals never couple x and u; control residuals are u - ref), so those entries are neither written by
  // node_diff_kernel (the tile buffer is zero-initialised) nor read here.
  static constexpr int LBLK = m + n + m;
  static constexpr int PREF = (LBLK + 31) / 32;
  // resident warps (= OCPs) per SM; the register budget follows from it (65536 / (32 WARPS))
#ifndef EMPC_BW_WARPS
#define EMPC_BW_WARPS 12
#endif
  static constexpr int WARPS = ((TOTAL * 8 + 1024) * EMPC_BW_WARPS <= 227 * 1024 && KN * PT <= 20) ? EMPC_BW_WARPS : 8;  // the wide platforms keep 8
  // (a warp lives on one of the four sub-partitions and takes its registers from that sub-partition's 16384: WARPS / 4
  //  warps per sub-partition leave 16384 / (WARPS / 4) / 32 registers per thread — 255 for 8 warps per SM, 168 for 12,
  //  128 for 16; other values of WARPS only round down to one of these)
  static constexpr int WPS = (WARPS + 3) / 4;
  static constexpr int MAXREG = (16384 / WPS / 32) / 8 * 8 > 255 ? 255 : (16384 / WPS / 32) / 8 * 8;
  // tile rows of F^T V per pass: all at once when the registers allow it, otherwise in halves
  static constexpr int IH = (PT <= 3 || (MAXREG >= 200 && KN * PT <= 20)) ? PT : (PT + 1) / 2;
  static constexpr int HALVES = (PT + IH - 1) / IH;
  static_assert((TOTAL * 8 + 1024) * WARPS <= 228 * 1024, "shared memory of the resident warps exceeds the SM");
};

struct BwParams {
  double reg_max, reg_factor, th_gaptol;
  int force;  // phase hook: single attempt, xreg / is_feasible taken from the state as they are, no prologue
  int stop_qu_norm;  // EMPC_STOP_CRITERIA_QU_NORM: also leave sum_t ||Qu_t||^2 in the OCP state
  // SolverBoxFDDP / SolverBoxDDP (backward_kernel<D, true, true>): crocoddyl::BoxQP(nu, maxiter, th_acceptstep, th_grad, reg)
  int qp_maxiter;
  double qp_th_acceptstep, qp_th_grad, qp_reg;
};

// crocoddyl::BoxQP::solve for one node, run by the WARP (the problem is nu x nu: 4 .. 11 unknowns; lane i owns unknown i):
// projected Newton on  min 1/2 x' H x + q' x,  u_lb - u <= x <= u_ub - u,  from the clamped warm start (the k_[t] of the
// previous sweep).  Mirrors the oracle's box_qp (oracle/oracle.cpp) decision by decision: clamped = on a bound with the
// gradient pushing outwards, converged when the gradient's infinity norm <= th_grad or nothing is free, Newton step on the
// free block, projected line search alpha = 1 .. 1/512 with the Armijo test.  Two exits the reference does not have, both at
// iterates it would keep (bit for bit, or to the last few ulp) for the rest of its maxiter iterations: no step length was
// accepted (the loop is deterministic: every further iteration repeats this one), and a Newton step below 1e-15 relative
// (the iterate is the minimiser on its free set; the reference random-walks on the last bit from here).
// The free block is never compacted: the factorisation runs on the full matrix with the rows / columns of the clamped
// unknowns replaced by the identity (their factor entries are exact zeros, so the free entries are those of the compacted
// factorisation).  It is the sweep's own right-looking L D L^T by shuffles (lane i holds row i), the substitutions and the
// matrix-vector products travel by shuffles too, sums are butterfly reductions (identical on every lane, so every decision
// is warp-uniform).  A first version ran the QP on lane 0 alone: 21 k cycles per QP iteration, 49 k per node (clock64
// counters) — one dependent FP64 chain with the sweep's registers spilled around it.
// Leaves in shared memory what the gain solves of the sweep expect from a factorisation — L row-wise (sL), column-wise
// (sLT), reciprocal pivots (sLinv) — for the masked matrix of the FINAL active set.  Returns the bit mask of the clamped
// unknowns, or -1 when a free block is not positive definite (the reference's "backward_error"); xout = this lane's x.
template <int m, int LM>
__device__ __noinline__ int bw_box_qp(const double* H, const double* q, const double* u, const double* u_lb, const double* u_ub, const double* xinit,
                                       double* sL, double* sLT, double* sLinv, const BwParams& P, double& xout) {
  constexpr unsigned FULL = 0xffffffffu, MM = (1u << m) - 1u;
  const int lane = threadIdx.x, i = lane < m ? lane : m - 1;
  const bool on = lane < m;
  auto wmax = [&](double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
  };
  double h[m];  // row i of H
#pragma unroll
  for (int k = 0; k < m; ++k) h[k] = H[i * LM + k];
  const double qi = q[i], lo = u_lb[i] - u[i], hi = u_ub[i] - u[i];
  double xi = fmax(fmin(xinit[i], hi), lo), dinv_i = 1.0;
  double a[m], lt[m];  // row i of the unit-lower factor (entries k < i) / column i of it (entries k > i)
  unsigned cl = 0u, cl_fact = ~0u;
  bool cli = false;
  auto factor = [&]() -> bool {
    __syncwarp();  // the column loads of the previous factorisation (lt) are done before its shared-memory arrays are rewritten
#pragma unroll
    for (int k = 0; k < m; ++k) {
      const bool masked = cli || ((cl >> k) & 1u);
      a[k] = masked ? ((k == i) ? 1.0 : 0.0) : (h[k] + ((k == i) ? P.qp_reg : 0.0));
    }
    int bad = 0;
#pragma unroll
    for (int j = 0; j < m; ++j) {
      const double d = __shfl_sync(FULL, a[j], j);
      double acj[m];
#pragma unroll
      for (int c = j + 1; c < m; ++c) acj[c] = __shfl_sync(FULL, a[j], c);  // A(c, j) = d L(c, j)
      if (!(d > 0.0)) bad = 1;
      const double dinv = rcp_h(d);
      const double lij = a[j] * dinv;
      a[j] = lij;
#pragma unroll
      for (int c = j + 1; c < m; ++c) a[c] = fma(-lij, acj[c], a[c]);
      if (on) sLT[j * LM + i] = lij;  // column j of L = row j of L^T (entries i <= j are never read)
      if (j == i) dinv_i = dinv;
      if (lane == j) sLinv[j] = dinv;
    }
    if (on) {
#pragma unroll
      for (int k = 0; k < m; ++k) sL[i * LM + k] = a[k];  // (entries k >= i are never read)
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < m; ++k) lt[k] = sLT[i * LM + k];
    cl_fact = cl;
    return !bad;  // uniform: every lane evaluates every pivot
  };
  for (int it = 0; it < P.qp_maxiter; ++it) {
    double gi = qi;
#pragma unroll
    for (int j = 0; j < m; ++j) gi = fma(h[j], __shfl_sync(FULL, xi, j), gi);
    cli = (xi == lo && gi > 0.0) || (xi == hi && gi < 0.0);
    cl = __ballot_sync(FULL, cli && on) & MM;
    const double gmax = wmax(on ? fabs(gi) : 0.0);
    if (gmax <= P.qp_th_grad || cl == MM) {
      if (cl_fact != cl && !factor()) return -1;
      break;
    }
    if (!factor()) return -1;
    // dxf = Hff^-1 (-qf - Hfc xc) - xf through the masked factor (clamped rows: right-hand side 0, pivot 1)
    double r = -qi;
#pragma unroll
    for (int c = 0; c < m; ++c) {
      const double xc = __shfl_sync(FULL, xi, c);
      if ((cl >> c) & 1u) r = fma(-h[c], xc, r);
    }
    double y = cli ? 0.0 : r;
#pragma unroll
    for (int k = 0; k < m; ++k) {  // L y = r
      const double yk = __shfl_sync(FULL, y, k);
      if (i > k) y = fma(-a[k], yk, y);
    }
    y *= dinv_i;
#pragma unroll
    for (int k = m - 1; k >= 0; --k) {  // L' w = D^-1 y
      const double wk = __shfl_sync(FULL, y, k);
      if (i < k) y = fma(-lt[k], wk, y);
    }
    const double dxi = (cli || !on) ? 0.0 : y - xi;
    const bool big = fabs(dxi) > 1e-15 * fmax(1.0, fabs(xi));
    if (__ballot_sync(FULL, big) == 0u) break;  // the Newton step no longer moves the iterate
    auto wsum = [&](double v) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      return v;
    };
    const double fold = 0.5 * wsum(on ? xi * (gi + qi) : 0.0);  // f(x) = 1/2 x'Hx + q'x = 1/2 x'(g + q)
    bool moved = false;
    for (int n = 0; n < EMPC_N_ALPHAS; ++n) {
      const double al = 1.0 / (double)(1 << n);
      const double xn = fmax(fmin(xi + al * dxi, hi), lo);
      double hx = 0.0;
#pragma unroll
      for (int j = 0; j < m; ++j) hx = fma(h[j], __shfl_sync(FULL, xn, j), hx);
      const double fnew = wsum(on ? xn * (0.5 * hx + qi) : 0.0);
      const double gd = wsum(on ? gi * (xi - xn) : 0.0);
      if (fold - fnew > P.qp_th_acceptstep * gd) { xi = xn; moved = true; break; }  // uniform: the sums are identical on every lane
    }
    if (!moved) break;
  }
  xout = xi;
  return (int)cl;
}

// optional phase timing (-DEMPC_BW_PROFILE): lane 0 of block 0 accumulates clock64() differences between the marks of a
// node into bf.nodesc of the last OCP... (diagnostic builds only; scripts/diag/backward_phases.py)
#ifdef EMPC_BW_PROFILE
__device__ unsigned long long g_bw_prof[16];
#define EMPC_BW_MARK(k) do { if (lane == 0 && blockIdx.x == 0) { const long long c_ = clock64(); if ((k) > 0) atomicAdd(&g_bw_prof[(k)], (unsigned long long)(c_ - prof_t)); else if (t < T - 1) atomicAdd(&g_bw_prof[8], (unsigned long long)(c_ - prof_t)); prof_t = c_; } } while (0)
#else
#define EMPC_BW_MARK(k) do { } while (0)
#endif

EMPC_DI void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MT_, int NT_>
EMPC_DI void acc_zero(double (&acc)[MT_][NT_][2]) {
#pragma unroll
  for (int i = 0; i < MT_; ++i)
#pragma unroll
    for (int j = 0; j < NT_; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
}

// column swizzle of the LD = 24 matrices: element (r, c) lives at r * 24 + (c ^ bw_swz(r))
EMPC_DI constexpr int bw_swz(int r) { return (r & 2) << 1; }

// COUPLED: some cost set holds a contact-force cost, the only kind that couples x and u (contact.cuh): its nodes add the
// Lxu block and the off-diagonal part of Luu that every other node leaves at zero
// BOX: SolverBoxFDDP / SolverBoxDDP::computeGains — when the candidate is feasible the gains of a node come from the box QP
// on its Quu / Qu (bw_box_qp) instead of the LDL^T solves
template <class D, bool COUPLED = false, bool BOX = false>
__global__ void __launch_bounds__(32) __maxnreg__(BwCfg<D>::MAXREG) backward_kernel(Buffers bf, BwParams P) {
  using S = BwCfg<D>;
  constexpr int n = S::n, m = S::m, LD = S::LD, LM = S::LM, PW = S::PW;
  extern __shared__ __align__(16) double bw_sm[];
  double* sm = bw_sm;
  const int lane = threadIdx.x, fr = lane >> 2, fc = lane & 3;
  // fragment coordinates under the swizzle.  B / transposed-A fragments: element (4 ks + fc, 8 j + fr) -> column 8 j + frs;
  // C fragments and row-major A fragments: row 8 i + fr -> column offset XOR swr
  const int swr = bw_swz(fr), frs = fr ^ bw_swz(fc), c2s = (2 * fc) ^ swr;
  const int b = bf.b0 + blockIdx.x;
  OcpState st = bf.st[b];
  if (!P.force && st.phase == PHASE_DONE) return;
  const int T = bf.T, T1 = T + 1;
  const size_t nb = (size_t)b * T1;

  // ---- prologue: tail of SolverDDP::calcDiff — cost_ = sum of node costs in node order, feasibility from the gaps ----
  if (!P.force && st.recalc) {
    double cst = 0;
    for (int base = 0; base < T1; base += S::TOTAL) {
      const int cnt = min(S::TOTAL, T1 - base);
      for (int t = lane; t < cnt; t += 32) sm[t] = bf.node_cost[nb + base + t];
      __syncwarp();
      if (lane == 0) for (int t = 0; t < cnt; ++t) cst += sm[t];
      __syncwarp();
    }
    st.cost = __shfl_sync(0xffffffffu, cst, 0);
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
  }
  const int feasible = st.is_feasible;

  double* sV = sm + S::oV; double* sFTV = sm + S::oFTV; double* sQux = sm + S::oQux; double* sK = sm + S::oK;
  double* sQuu = sm + S::oQuu; double* sL = sm + S::oL; double* sLT = sL + m * LM; double* sLinv = sLT + m * LM;
  double* vec = sm + S::oVec;
  double* Qx = vec + S::vQx; double* Qu = vec + S::vQu; double* Vxp = vec + S::vVx; double* fsv = vec + S::vFs;
  double* gv = vec + S::vG; double* kv = vec + S::vKv; double* Quuk = vec + S::vQuuk; double* tmpv = vec + S::vTmp;
  double* Luud = vec + S::vLuu;

  // F = [Fx | Fu] of a node as DMMA fragments, HBM -> registers: f[ks][i] = F(4 ks + fc, 8 i + fr), zero beyond (n, PW).
  // The same element is this lane's share of the transposed-A fragment of F^T V and of the B fragment of (F^T V) F.
  int f_off[S::PT], f_str[S::PT];  // per tile column: offset of (row fc, this lane's column) in the node tile, row stride
#pragma unroll
  for (int i = 0; i < S::PT; ++i) {
    const int col = 8 * i + fr;
    f_str[i] = col < n ? n : m;
    f_off[i] = (col < n ? D::oFx + col : D::oFu + (col - n)) + fc * f_str[i];
    if (col >= PW) { f_off[i] = 0; f_str[i] = 0; }
  }
  double f[S::KN][S::PT];
  auto load_F = [&](int t) {
    const double* tg = bf.tiles + (nb + t) * D::TILE;
#pragma unroll
    for (int ks = 0; ks < S::KN; ++ks)
#pragma unroll
      for (int i = 0; i < S::PT; ++i) {
        const bool on = (4 * ks + fc < n) && (8 * i + fr < PW);
        f[ks][i] = on ? __ldg(tg + f_off[i] + 4 * ks * f_str[i]) : 0.0;
      }
  };
  // cost blocks of node t: HBM -> registers (issued early) -> shared memory (at the end of the previous node).
  // Source / destination offsets of this lane's elements are fixed: computed once, packed in 32-bit registers.
  unsigned pre_off[S::PREF];  // (src offset in the tile, relative to oLxx) << 16 | (dst offset in sm)
#pragma unroll
  for (int q = 0; q < S::PREF; ++q) {
    int e = lane + 32 * q, src = 0, dst = 0xffff;
    if (e < m) { src = (D::oLuu - D::oLxx) + e * (m + 1); dst = S::oVec + S::vLuu + e; }
    else if (e < m + n) { const int i = e - m; src = (D::oLx - D::oLxx) + i; dst = S::oVec + S::vQx + i; }
    else if (e < S::LBLK) { const int i = e - m - n; src = (D::oLu - D::oLxx) + i; dst = S::oVec + S::vQu + i; }
    pre_off[q] = ((unsigned)src << 16) | (unsigned)dst;
  }
  auto load_L = [&](int t, double (&pre)[S::PREF], double& pre_fs) {
    const double* tg = bf.tiles + (nb + t) * D::TILE + D::oLxx;
#pragma unroll
    for (int q = 0; q < S::PREF; ++q) pre[q] = tg[pre_off[q] >> 16];
    pre_fs = (lane < n) ? bf.fs[(nb + t) * n + lane] : 0.0;
  };
  auto store_L = [&](const double (&pre)[S::PREF], double pre_fs) {
#pragma unroll
    for (int q = 0; q < S::PREF; ++q) { const unsigned d = pre_off[q] & 0xffffu; if (d != 0xffffu) sm[d] = pre[q]; }
    if (lane < n) fsv[lane] = pre_fs;
  };
  // Lxx of a node as accumulator fragments (row 8 i + fr, columns 8 j + 2 fc, +1; upper tiles), HBM -> registers one node
  // ahead, issued together with F: a load issued at the start of a node would stall its first DMMA for the whole HBM
  // latency (ncu round 2: 10 % of the kernel's time sat on that one instruction)
  double qn[S::NT][S::NT][2];
  auto load_Lxx = [&](int t) {
    const double* lg = bf.tiles + (nb + t) * D::TILE + D::oLxx;
#pragma unroll
    for (int i = 0; i < S::NT; ++i)
#pragma unroll
      for (int j = i; j < S::NT; ++j) {
        const int row = 8 * i + fr, col = 8 * j + 2 * fc;
        double2 v = make_double2(0.0, 0.0);
        if (row < n && col < n) v = __ldg(reinterpret_cast<const double2*>(lg + row * n + col));
        qn[i][j][0] = v.x; qn[i][j][1] = v.y;
      }
  };
  // operands of the dot-product DMMA at the end of a node (see there): this lane's row of A / column of B
  const double* dot_a = vec + ((fr == 0) ? S::vQu : (fr == 1) ? S::vKv : (fr == 2) ? S::vVx : (fr == 3) ? S::vFs : S::vQu);
  const double* dot_b = vec + ((fr == 0) ? S::vKv : (fr == 1) ? S::vQuuk : (fr == 2) ? S::vFs : (fr == 3) ? S::vG : S::vQu);
  // column n of the V operand carries Vx': tile JV, fragment column FCV (n is even: slot 0 of the pair)
  constexpr int JV = n / 8, FCV = (n % 8) / 2;
  // this lane as a column index under the swizzle of rows with bit 1 set: (row j, column `lane`)
  const int lane_sw = lane ^ 4;

#ifdef EMPC_BW_PROFILE
  long long prof_t = clock64();
#endif
  double pre[S::PREF], pre_fs = 0.0;
  int failed;
  while (true) {
    failed = 0;
    const double xreg = st.xreg;
    // zero everything: the padding of every DMMA operand must be (and then stays) zero
    for (int i = lane; i < S::TOTAL; i += 32) sm[i] = 0.0;
    __syncwarp();
    // ---- terminal node: Vxx = Lxx + xreg I ; Vx = Lx (+ Vxx fs) ----
    {
      const double* tg = bf.tiles + (nb + T) * D::TILE;
      for (int e = lane; e < n * n; e += 32) { const int i = e / n, j = e - i * n; sV[i * LD + (j ^ bw_swz(i))] = tg[D::oLxx + e] + ((i == j) ? xreg : 0.0); }
      if (lane < n) { Vxp[lane] = tg[D::oLx + lane]; fsv[lane] = bf.fs[(nb + T) * n + lane]; }
      __syncwarp();
      if (lane < n) {
        double s = 0;
        const int sw = bw_swz(lane);
#pragma unroll 6
        for (int j = 0; j < n; ++j) s += sV[lane * LD + (j ^ sw)] * fsv[j];
        gv[lane] = s;
        if (!feasible) Vxp[lane] += s;
      }
      __syncwarp();
      if (lane == 0) {
        double s0 = 0, s1 = 0;
        for (int i = 0; i < n; ++i) { s0 += Vxp[i] * fsv[i]; s1 += fsv[i] * gv[i]; }
        double* ns = bf.nodesc + (nb + T) * 4;
        ns[0] = 0; ns[1] = 0; ns[2] = s0; ns[3] = s1;
      }
      if (lane < n) { bf.Vx[(nb + T) * n + lane] = Vxp[lane]; bf.g[(nb + T) * n + lane] = gv[lane]; sV[lane * LD + (n ^ bw_swz(lane))] = Vxp[lane]; }
      load_F(T - 1);
      load_Lxx(T - 1);
      load_L(T - 1, pre, pre_fs);
      __syncwarp();
      store_L(pre, pre_fs);
    }
    for (int t = T - 1; t >= 0; --t) {
      __syncwarp();
      EMPC_BW_MARK(0);
      // packed accumulator q of [[Qxx, Qxu], [Qux, Quu]] (upper tiles only): its top-left tiles start from Lxx
      double q[S::PT][S::PT][2];
#pragma unroll
      for (int i = 0; i < S::PT; ++i)
#pragma unroll
        for (int j = i; j < S::PT; ++j) {
          if (j < S::NT) { q[i][j][0] = qn[i][j][0]; q[i][j][1] = qn[i][j][1]; }
          else { q[i][j][0] = 0.0; q[i][j][1] = 0.0; }
        }
      // ---- F^T [V | Vx'] = [FxTV ; FuTV | Fx^T Vx' ; Fu^T Vx'].  Column n of the product is the Qx / Qu update (each
      // entry owned by one lane); it is not part of F^T V and is stored as zero. ----
#pragma unroll
      for (int hf = 0; hf < S::HALVES; ++hf) {
        double ftv[S::IH][S::NTV][2];
        acc_zero(ftv);
#pragma unroll
        for (int ks = 0; ks < S::KN; ++ks) {
          double vv[S::NTV];
#pragma unroll
          for (int j = 0; j < S::NTV; ++j) vv[j] = sV[(4 * ks + fc) * LD + 8 * j + frs];
#pragma unroll
          for (int ii = 0; ii < S::IH; ++ii)
            if (hf * S::IH + ii < S::PT) {
#pragma unroll
              for (int j = 0; j < S::NTV; ++j) dmma884(ftv[ii][j][0], ftv[ii][j][1], f[ks][hf * S::IH + ii], vv[j]);
            }
        }
#pragma unroll
        for (int ii = 0; ii < S::IH; ++ii)
          if (hf * S::IH + ii < S::PT) {
            const int row = 8 * (hf * S::IH + ii) + fr;
            if (fc == FCV) {
              const double v = ftv[ii][JV][0];
              if (row < n) Qx[row] += v;
              else if (row < PW) Qu[row - n] += v;
              ftv[ii][JV][0] = 0.0;
            }
#pragma unroll
            for (int j = 0; j < S::NT; ++j)
              if (row < PW) *reinterpret_cast<double2*>(sFTV + row * LD + 8 * j + c2s) = make_double2(ftv[ii][j][0], ftv[ii][j][1]);
          }
      }
      __syncwarp();
      EMPC_BW_MARK(1);
      // ---- [[Qxx, Qxu], [., Quu]] = [[Lxx, 0], [0, Luu + ureg I]] + (F^T V) F, upper tiles of the packed symmetric matrix.
      // Qxx (the top-left tiles) stays in registers until Qxu K has been subtracted from it; Qux and Quu go to shared memory. ----
      {
#pragma unroll
        for (int i = 0; i < S::PT; ++i)
#pragma unroll
          for (int j = i; j < S::PT; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int row = 8 * i + fr, col = 8 * j + 2 * fc + h;
              if (row == col && row >= n && row < PW) q[i][j][h] = Luud[row - n] + xreg;
            }
#pragma unroll
        for (int ks = 0; ks < S::KN; ++ks) {
          double fa[S::PT];
#pragma unroll
          for (int i = 0; i < S::PT; ++i) {
            // F^T V keeps only its PW live rows: the lanes of the last tile row that fall beyond read row PW - 1 again
            // (finite values that only reach accumulator rows nobody uses)
            const int row = (8 * i + 7 < PW) ? 8 * i + fr : min(8 * i + fr, PW - 1);
            fa[i] = sFTV[row * LD + ((4 * ks + fc) ^ bw_swz(row))];
          }
#pragma unroll
          for (int i = 0; i < S::PT; ++i)
#pragma unroll
            for (int j = i; j < S::PT; ++j) dmma884(q[i][j][0], q[i][j][1], fa[i], f[ks][j]);
        }
#pragma unroll
        for (int i = 0; i < S::PT; ++i)
#pragma unroll
          for (int j = i; j < S::PT; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int row = 8 * i + fr, col = 8 * j + 2 * fc + h;
              if (col >= n && col < PW) {
                const int cu = col - n;
                if (row < n) sQux[cu * LD + (row ^ bw_swz(cu))] = q[i][j][h];
                else if (row < PW) {
                  sQuu[(row - n) * LM + cu] = q[i][j][h];
                  if (i != j) sQuu[cu * LM + (row - n)] = q[i][j][h];  // the lower tiles are not computed: mirror
                }
              }
            }
      }
      __syncwarp();
      if (COUPLED) {
        if (bf.ct.costset_coupled[bf.node_costset[(size_t)bf.ocp_map[b] * T1 + t]]) {  // warp-uniform
          const double* tg = bf.tiles + (nb + t) * D::TILE;
          for (int e = lane; e < n * m; e += 32) { const int i = e / m, cu = e - i * m; sQux[cu * LD + (i ^ bw_swz(cu))] += tg[D::oLxu + e]; }
          for (int e = lane; e < m * m; e += 32) { const int r = e / m, c = e - r * m; if (r != c) sQuu[r * LM + c] += tg[D::oLuu + e]; }
          __syncwarp();
        }
      }
      EMPC_BW_MARK(2);
      // Fx, Fu of this node are dead: the next node's fragments travel HBM -> registers while Quu is factorised
      if (t > 0) { load_F(t - 1); load_Lxx(t - 1); load_L(t - 1, pre, pre_fs); }
      // ---- Quu = L D L^T (unit lower L, no square roots), right-looking in registers: lane i holds row i.  Per pivot the
      // dependent chain is shuffle (pivot) -> reciprocal -> multiply -> FMA; the UNSCALED column entries A(c, j) the update
      // needs travel by shuffles that do not wait for the reciprocal (an FP64 operation has ~20 cycles of latency on B200
      // and co-resident warps' DMMAs stretch every dependent step: the LL^T form with its rsqrt -> multiply -> shuffle ->
      // FMA chain per pivot and a multiply + FMA per substitution step cost ~0.7 k more cycles per node).  Positive pivots
      // <=> Quu positive definite, the same failure test as the reference's LLT.  L goes to shared memory row-wise and
      // column-wise for the substitutions, the reciprocal pivots beside it. ----
      int bad = 0;
      bool boxed = false;
      double box_x = 0.0;
      if constexpr (BOX) {
        if (feasible) {  // uniform over the warp (one OCP); an infeasible candidate takes the plain gains below
          boxed = true;
          const size_t nodeu = (size_t)b * T + t;
          const int qp = bw_box_qp<m, LM>(sQuu, Qu, bf.us + nodeu * m, bf.model->u_lb, bf.model->u_ub, bf.k + nodeu * m, sL, sLT, sLinv, P, box_x);
          if (qp < 0) { failed = 1; break; }
          // K = Quu_inv Qxu^T with Quu_inv = the inverse of the free block (zero rows / columns for the clamped controls): the
          // gain solves below run on the masked factor the QP left behind, so the clamped rows of their right-hand sides (of
          // Qxu^T) are zeroed; the clamped entries of Qu are zeroed for good ("important for accounting the algorithm
          // advancement": expected improvement and stopping criterion see the projected gradient)
          __syncwarp();
          for (int c = lane; c < LD; c += 32) {
#pragma unroll
            for (int cu = 0; cu < m; ++cu) if ((qp >> cu) & 1) sQux[cu * LD + c] = 0.0;
          }
          if (lane < m && ((qp >> lane) & 1)) Qu[lane] = 0.0;
          __syncwarp();
        }
      }
      if (!boxed) {
      {
        const int i = lane < m ? lane : m - 1;
        double a[LM];
#pragma unroll
        for (int k = 0; k < LM; k += 2) { const double2 v = *reinterpret_cast<const double2*>(sQuu + i * LM + k); a[k] = v.x; a[k + 1] = v.y; }
#pragma unroll
        for (int j = 0; j < m; ++j) {
          const double d = __shfl_sync(0xffffffffu, a[j], j);
          double acj[m];
#pragma unroll
          for (int c = j + 1; c < m; ++c) acj[c] = __shfl_sync(0xffffffffu, a[j], c);  // A(c, j) = d L(c, j): independent of the reciprocal
          if (!(d > 0.0)) bad = 1;
          const double dinv = rcp_h(d);
          const double lij = a[j] * dinv;
          a[j] = lij;
#pragma unroll
          for (int c = j + 1; c < m; ++c) a[c] = fma(-lij, acj[c], a[c]);
          if (lane < m) sLT[j * LM + i] = lij;  // column j of L = row j of L^T (entries i <= j are never read)
          if (lane == j) sLinv[j] = dinv;
        }
        if (lane < m) {
#pragma unroll
          for (int k = 0; k < LM; k += 2) *reinterpret_cast<double2*>(sL + i * LM + k) = make_double2(a[k], a[k + 1]);  // (entries k >= i are never read)
        }
        __syncwarp();
      EMPC_BW_MARK(3);
      }
      if (bad) { failed = 1; break; }  // uniform: every lane evaluates every pivot
      }  // !boxed
      // ---- gains: K = Quu^-1 Qxu^T (one right-hand side per lane), k = Quu^-1 Qu ----
      for (int c = lane; c < n + 1; c += 32) {
        double rhs[m];
        if (c < n) {
#pragma unroll
          for (int i = 0; i < m; ++i) rhs[i] = sQux[i * LD + (c ^ bw_swz(i))];
        } else {
#pragma unroll
          for (int i = 0; i < m; ++i) rhs[i] = Qu[i];
        }
        double linv[LM];
#pragma unroll
        for (int k = 0; k < LM; k += 2) { const double2 v = *reinterpret_cast<const double2*>(sLinv + k); linv[k] = v.x; linv[k + 1] = v.y; }
        // column-oriented substitutions with the unit-lower factor: as soon as an entry is final it is subtracted from all
        // the others, so the dependent chain is ONE FMA per column (L y = b, then z = D^-1 y, then L^T x = z); column kk of
        // L is row kk of L^T (16-byte broadcasts)
#pragma unroll
        for (int kk = 0; kk < m; ++kk) {
#pragma unroll
          for (int i = (kk + 1) & ~1; i < m; i += 2) {
            const double2 v = *reinterpret_cast<const double2*>(sLT + kk * LM + i);
            if (i > kk) rhs[i] -= v.x * rhs[kk];
            if (i + 1 < m) rhs[i + 1] -= v.y * rhs[kk];
          }
        }
#pragma unroll
        for (int kk = 0; kk < m; ++kk) rhs[kk] *= linv[kk];
#pragma unroll
        for (int kk = m - 1; kk >= 0; --kk) {
#pragma unroll
          for (int i = 0; i < kk; i += 2) {
            const double2 v = *reinterpret_cast<const double2*>(sL + kk * LM + i);
            rhs[i] -= v.x * rhs[kk];
            if (i + 1 < kk) rhs[i + 1] -= v.y * rhs[kk];
          }
        }
        if (c < n) {
#pragma unroll
          for (int i = 0; i < m; ++i) sK[i * LD + (c ^ bw_swz(i))] = rhs[i];
        } else {
#pragma unroll
          for (int i = 0; i < m; ++i) kv[i] = rhs[i];
        }
      }
      if constexpr (BOX) {
        if (boxed) {  // k = -du of the box QP (the column solved above for k belongs to the unconstrained step)
          __syncwarp();
          if (lane < m) kv[lane] = -box_x;
        }
      }
      __syncwarp();
      EMPC_BW_MARK(4);
      if (lane < m) {  // Quuk = Quu k
        double s[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < m; ++j) s[j % 3] = fma(sQuu[lane * LM + j], kv[j], s[j % 3]);
        const double quuk = (s[0] + s[1]) + s[2];
        Quuk[lane] = quuk;
        tmpv[lane] = quuk - 2.0 * Qu[lane];  // w = Quu k - 2 Qu
      }
      __syncwarp();
      // Vx = Qx + K^T Quuk - 2 K^T Qu = Qx + K^T (Quuk - 2 Qu)
      double vx_q = 0.0;
      const double* wv = tmpv;
      if (lane < n) {
        double s1[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < m; ++j) s1[j % 3] = fma(sK[j * LD + ((j & 2) ? lane_sw : lane)], wv[j], s1[j % 3]);
        vx_q = Qx[lane] + ((s1[0] + s1[1]) + s1[2]);
      }
      // ---- Vxx = sym(Qxx - Qxu K) + xreg I, all in the accumulator registers.  Qxx and Qxu K are symmetric up to
      // rounding, so only the upper tiles (j >= i) are computed: diagonal tiles are averaged with their own transpose
      // (the reference's 0.5 (Vxx + Vxx^T)), the lower tiles are the mirror images of the upper ones.  Entries of the
      // top-left tiles beyond row / column n belong to Qxu / Quu and are masked: V's padding has to stay zero. ----
#pragma unroll
      for (int ks = 0; ks < S::KM; ++ks) {
        double a[S::NT], bq[S::NT];
#pragma unroll
        for (int i = 0; i < S::NT; ++i) { a[i] = -sQux[(4 * ks + fc) * LD + 8 * i + frs]; bq[i] = sK[(4 * ks + fc) * LD + 8 * i + frs]; }
#pragma unroll
        for (int i = 0; i < S::NT; ++i)
#pragma unroll
          for (int j = i; j < S::NT; ++j) dmma884(q[i][j][0], q[i][j][1], a[i], bq[j]);
      }
      {
        // the mirror image of this lane's pair (row fr, columns 2 fc + h) of tile (i, j) is element (2 fc + h, fr) of
        // tile (j, i): it sits in lane (2 fc + h) * 4 + (fr >> 1), slot fr & 1
        double vsym[S::NT][S::NT][2];
        const int src0 = (2 * fc) * 4 + (fr >> 1), src1 = (2 * fc + 1) * 4 + (fr >> 1);
        const bool odd = fr & 1;
#pragma unroll
        for (int i = 0; i < S::NT; ++i)
#pragma unroll
          for (int j = i; j < S::NT; ++j) {
            // transposed element of the upper tile (i, j), as seen from this lane
            const double a0 = __shfl_sync(0xffffffffu, q[i][j][0], src0), a1 = __shfl_sync(0xffffffffu, q[i][j][1], src0);
            const double b0 = __shfl_sync(0xffffffffu, q[i][j][0], src1), b1 = __shfl_sync(0xffffffffu, q[i][j][1], src1);
            const double t0 = odd ? a1 : a0, t1 = odd ? b1 : b0;
            const int row = 8 * i + fr, col = 8 * j + 2 * fc;   // this lane's entries of tile (i, j): (row, col), (row, col + 1)
            const int rowT = 8 * j + fr, colT = 8 * i + 2 * fc;  // ... and of the mirrored tile (j, i)
            if (i == j) {
              double v0 = 0.5 * (q[i][i][0] + t0), v1 = 0.5 * (q[i][i][1] + t1);
              if (row == col) v0 += xreg;
              if (row == col + 1) v1 += xreg;
              if (row >= n || col >= n) v0 = 0.0;
              if (row >= n || col + 1 >= n) v1 = 0.0;
              if (raise_if_nan_abs(v0) || raise_if_nan_abs(v1)) bad = 1;  // "backward_error": raiseIfNaN(Vxx.lpNorm<Infinity>())
              vsym[i][i][0] = v0; vsym[i][i][1] = v1;
            } else {
              const double u0 = (row < n && col < n) ? q[i][j][0] : 0.0, u1 = (row < n && col + 1 < n) ? q[i][j][1] : 0.0;
              if (raise_if_nan_abs(u0) || raise_if_nan_abs(u1)) bad = 1;
              vsym[i][j][0] = u0; vsym[i][j][1] = u1;
              vsym[j][i][0] = (rowT < n && colT < n) ? t0 : 0.0;
              vsym[j][i][1] = (rowT < n && colT + 1 < n) ? t1 : 0.0;
            }
          }
#pragma unroll
        for (int i = 0; i < S::NT; ++i)
#pragma unroll
          for (int j = 0; j < S::NT; ++j)
            if (8 * i + fr < S::ROWS_V) *reinterpret_cast<double2*>(sV + (8 * i + fr) * LD + 8 * j + c2s) = make_double2(vsym[i][j][0], vsym[i][j][1]);
      }
      __syncwarp();
      EMPC_BW_MARK(5);
      if (lane < n) {
        double s0 = 0, s1 = 0, s2 = 0;  // row `lane` of the symmetric V read as a column: conflict-free
#pragma unroll
        for (int j = 0; j + 2 < n; j += 3) {
          s0 += sV[j * LD + ((j & 2) ? lane_sw : lane)] * fsv[j];
          s1 += sV[(j + 1) * LD + (((j + 1) & 2) ? lane_sw : lane)] * fsv[j + 1];
          s2 += sV[(j + 2) * LD + (((j + 2) & 2) ? lane_sw : lane)] * fsv[j + 2];
        }
#pragma unroll
        for (int j = n - n % 3; j < n; ++j) s0 += sV[j * LD + ((j & 2) ? lane_sw : lane)] * fsv[j];
        const double s = (s0 + s1) + s2;
        gv[lane] = s;
        const double vx = feasible ? vx_q : (vx_q + s);
        if (raise_if_nan_abs(vx)) bad = 1;  // raiseIfNaN(Vx.lpNorm<Infinity>())
        Vxp[lane] = vx;
        sV[lane * LD + (n ^ bw_swz(lane))] = vx;  // column n of the V operand of the next node (the V store above left zero there)
      }
      bad = __any_sync(0xffffffffu, bad);
      if (bad) { failed = 1; break; }
      __syncwarp();
      EMPC_BW_MARK(6);
      // outputs
      {
        double* Kg = bf.K + ((size_t)b * T + t) * m * n;
        for (int e2 = lane; e2 < m * n / 2; e2 += 32) {
          const int e = 2 * e2, i = e / n, j = e - i * n;  // n even: a pair never straddles rows
          __stcs(reinterpret_cast<double2*>(Kg) + e2, *reinterpret_cast<const double2*>(sK + i * LD + (j ^ bw_swz(i))));  // written once, read by the rollout much later: streaming
        }
        double* kg = bf.k + ((size_t)b * T + t) * m;
        if (lane < m) __stcs(kg + lane, kv[lane]);
        if (lane < n) { __stcs(bf.Vx + (nb + t) * n + lane, Vxp[lane]); __stcs(bf.g + (nb + t) * n + lane, gv[lane]); }
        {
          // five dot products as ONE 8x8 tensor-core product: row w of A and column w of B are the two vectors of
          // product w (Qu.k, k.Quuk, Vx.fs, fs.(Vxx fs), Qu.Qu), KN k-steps over the zero-padded vector slots; the
          // result is the diagonal: entry (w, w) sits in lane 4 w + (w >> 1), slot w & 1.  All lanes take part, nothing
          // diverges, and the dependent chain is KN DMMAs instead of n FMAs on five lanes.
          double c0 = 0.0, c1 = 0.0;
#pragma unroll
          for (int ks = 0; ks < S::KN; ++ks) {
            const double av = (fr < 5) ? dot_a[4 * ks + fc] : 0.0, bv = (fr < 5) ? dot_b[4 * ks + fc] : 0.0;
            dmma884(c0, c1, av, bv);
          }
          if (fr < 5 && fc == (fr >> 1)) {
            const double v = (fr & 1) ? c1 : c0;
            if (fr < 4) bf.nodesc[(nb + t) * 4 + fr] = v;
            else bf.qu2[nb + t] = v;  // ||Qu_t||^2 (upstream stoppingCriteria)
          }
        }
      }
      __syncwarp();
      EMPC_BW_MARK(7);
      if (t > 0) store_L(pre, pre_fs);
    }
    __syncwarp();
    if (!failed || P.force) break;
    // computeDirection threw: recalcDiff = false; increaseRegularization(); give up at reg_max (src/sbfddp.cpp:245-253)
    st.xreg *= P.reg_factor;
    if (st.xreg > P.reg_max) st.xreg = P.reg_max;
    if (st.xreg == P.reg_max) break;
  }
  st.bw_fail = failed ? 1 : 0;
  // SolverFDDP::updateExpectedImprovement / expectedImprovementDDP: ordered sums over the nodes
  if (!failed) {
    double dg = 0, dq = 0, dg0 = 0, dq0 = 0;
    __syncwarp();
    if (!feasible && lane == 0) { dg -= bf.nodesc[(nb + T) * 4 + 2]; dq += bf.nodesc[(nb + T) * 4 + 3]; }
    constexpr int CH = S::TOTAL / 4;
    for (int base = 0; base < T; base += CH) {
      const int cnt = min(CH, T - base);
      for (int i = lane; i < cnt * 4; i += 32) sm[i] = bf.nodesc[(nb + base) * 4 + i];
      __syncwarp();
      if (lane == 0) {
        for (int t = 0; t < cnt; ++t) {
          dg += sm[t * 4 + 0]; dq -= sm[t * 4 + 1];
          dg0 += sm[t * 4 + 0]; dq0 -= sm[t * 4 + 1];
          if (!feasible) { dg -= sm[t * 4 + 2]; dq += sm[t * 4 + 3]; }
        }
      }
      __syncwarp();
    }
    if (P.stop_qu_norm) {  // SolverDDP::stoppingCriteria of upstream: sum over the running nodes, node order
      double s = 0;
      for (int base = 0; base < T; base += S::TOTAL) {
        const int cnt = min(S::TOTAL, T - base);
        for (int i = lane; i < cnt; i += 32) sm[i] = bf.qu2[nb + base + i];
        __syncwarp();
        if (lane == 0) for (int t = 0; t < cnt; ++t) s += sm[t];
        __syncwarp();
      }
      if (lane == 0) st.qu2 = s;
    }
    if (lane == 0) { st.dg = dg; st.dq = dq; st.dg0 = dg0; st.dq0 = dq0; }
  }
  if (lane == 0) bf.st[b] = st;
}
