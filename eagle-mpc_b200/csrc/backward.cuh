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
// The six dense products are the only GEMM-shaped work on the SbFDDP path (18x18x18 for flying_arm_3).  They run as
// mma.sync.m8n8k4.f64 (SASS DMMA) on 8x8 tiles with zero-padded operands in shared memory: one DMMA replaces 8 DFMA
// warp-instructions and needs two 8-byte fragment loads instead of ~7 (measured: the FP64 tensor rate equals the vector
// rate on B200, so the gain is issue slots and shared-memory bandwidth, not FLOPs; profiles/r1_baseline.md).
// Leading dimensions are chosen per operand role so that every fragment load is bank-conflict free:
//   LDB = 24 / LDF = 40  (= 8 mod 16)  operands read as B (row k) or as transposed A:   [Fx | Fu] (packed), V', K
//   LDA = 20  (= 4 mod 16)  operands read as row-major A (row i, 4 consecutive k):   F^T V, Qxu
// Fx and Fu are packed side by side, so that F^T V = [FxTV ; FuTV] and the symmetric [[Qxx, Qxu], [Qux, Quu]] =
// (F^T V) F each take one k-loop and only the upper tiles of the latter are computed: 128 DMMAs per node instead of
// 197 for the six separate padded products.
// Everything else (Cholesky of Quu, triangular solves, vector updates, ordered reductions) is warp-cooperative out of
// shared memory with __syncwarp only: no block-level barrier anywhere on the sweep.  The next node's Fx / Fu arrive by
// cp.async while the current node is being factorised; its cost blocks are prefetched into registers.
#pragma once

template <class D>
struct BwCfg {
  static constexpr int n = D::NDX, m = D::NU;
  static constexpr int NT = (n + 7) / 8, MT = (m + 7) / 8;      // 8-wide tiles over n, m
  static constexpr int KN = (n + 3) / 4, KM = (m + 3) / 4;      // 4-deep k-steps over n, m
  static constexpr int NP = 8 * NT, MP = 8 * MT, KNP = 4 * KN, KMP = 4 * KM;
  static constexpr int LDB = 24;
  // packed operand F = [Fx | Fu] (n x (n + m)): one product F^T V gives FxTV and FuTV, one product (F^T V) F gives the
  // whole symmetric matrix [[Qxx, Qxu], [Qux, Quu]], of which only the upper tiles are computed
  static constexpr int PW = n + m, PT = (PW + 7) / 8, PP = 8 * PT;
  static constexpr int LDF = (PP <= 24) ? 24 : 40;               // = 8 mod 16, >= PP
  static_assert(PP <= LDF, "packed operand wider than its leading dimension");
  static constexpr int lda_for(int k) { return k <= 20 ? 20 : 36; }
  static constexpr int LDA = lda_for(KNP);                       // F^T V (row-major A operand)
  static constexpr int LDQ = 20;                                 // Qxu (k extent KMP <= 16)
  static_assert(NP <= LDB && MP <= LDB && KMP <= LDQ, "operand wider than its leading dimension");
  static constexpr int ROWS_K = (KNP > NP ? KNP : NP);
  static constexpr int oF = 0;                                   // KNP x LDF      [Fx | Fu]
  static constexpr int oV = oF + KNP * LDF;                      // ROWS_K x LDB   Vxx' (symmetric)
  // (Qxx never touches shared memory: Lxx is loaded from HBM straight into the accumulator fragments and
  //  Qxx - Qxu K is symmetrised in registers)
  static constexpr int oQxu = oV + ROWS_K * LDB;                 // NP x LDQ       Qxu
  static constexpr int oQuu = oQxu + NP * LDQ;                   // MP x LDQ       Quu
  static constexpr int oFTV = oQuu + MP * LDQ;                   // PP x LDA       F^T V = [FxTV ; FuTV]
  static constexpr int oK = oFTV + PP * LDA;                     // KMP x LDB      gains K (m x n), zero padded
  // the Cholesky factor lives in the F^T V area, which is dead once the Q blocks are formed
  static constexpr int oL = oFTV;                                // m x m Cholesky factor of Quu, then m reciprocal pivots
  static_assert(m * m + m <= PP * LDA, "L does not fit the F^T V area");
  static constexpr int oVec = oK + KMP * LDB;
  static constexpr int vQx = 0, vQu = vQx + NP, vVx = vQu + MP, vFs = vVx + NP, vG = vFs + NP, vKv = vG + NP,
                       vQuuk = vKv + MP, vTmp = vQuuk + MP, vLuu = vTmp + NP, VEC = vLuu + MP;
  static constexpr int TOTAL0 = oVec + VEC;
  static constexpr int TOTAL = TOTAL0 + (TOTAL0 & 1);
  // register prefetch (one node ahead) of the small cost blocks: diag(Luu) (m), Lx | Lu (n + m, contiguous) + fs.  Lxu is
  // identically zero and Luu diagonal for every cost the reference's factories build (state / control / frame
  // residuals never couple x and u; control residuals are u - ref), so those entries are neither written by
  // node_diff_kernel (the tile buffer is zero-initialised) nor read here.
  static constexpr int LBLK = m + n + m;
  static constexpr int PREF = (LBLK + 31) / 32;
#ifndef EMPC_BW_WARPS_PER_SM
#define EMPC_BW_WARPS_PER_SM 8
#endif
};

struct BwParams {
  double reg_max, reg_factor, th_gaptol;
  int force;  // phase hook: single attempt, xreg / is_feasible taken from the state as they are, no prologue
  int stop_qu_norm;  // EMPC_STOP_CRITERIA_QU_NORM: also leave sum_t ||Qu_t||^2 in the OCP state
};

EMPC_DI void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// acc[i][j] (+)= op(A) B over KS k-steps.  Fragment coordinates: r = lane >> 2, c = lane & 3.
//   AT  : A is stored k-major (A(i,k) = Ab[k * lda + i]), else row-major (A(i,k) = Ab[i * lda + k])
//   NEG : accumulate -A B
template <int MT_, int NT_, int KS_, bool AT, bool NEG>
EMPC_DI void warp_mm(double (&acc)[MT_][NT_][2], const double* __restrict__ Ab, int lda, const double* __restrict__ Bb, int ldb, int r, int c) {
#pragma unroll
  for (int ks = 0; ks < KS_; ++ks) {
    double a[MT_], b[NT_];
#pragma unroll
    for (int i = 0; i < MT_; ++i) {
      const double v = AT ? Ab[(4 * ks + c) * lda + 8 * i + r] : Ab[(8 * i + r) * lda + 4 * ks + c];
      a[i] = NEG ? -v : v;
    }
#pragma unroll
    for (int j = 0; j < NT_; ++j) b[j] = Bb[(4 * ks + c) * ldb + 8 * j + r];
#pragma unroll
    for (int i = 0; i < MT_; ++i)
#pragma unroll
      for (int j = 0; j < NT_; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}
template <int MT_, int NT_>
EMPC_DI void acc_zero(double (&acc)[MT_][NT_][2]) {
#pragma unroll
  for (int i = 0; i < MT_; ++i)
#pragma unroll
    for (int j = 0; j < NT_; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
}
// C tile (i, j): lane holds C[8 i + r][8 j + 2 c], C[8 i + r][8 j + 2 c + 1]
template <int MT_, int NT_>
EMPC_DI void acc_load(double (&acc)[MT_][NT_][2], const double* Cb, int ldc, int r, int c) {
#pragma unroll
  for (int i = 0; i < MT_; ++i)
#pragma unroll
    for (int j = 0; j < NT_; ++j) {
      const double2 v = *reinterpret_cast<const double2*>(Cb + (8 * i + r) * ldc + 8 * j + 2 * c);
      acc[i][j][0] = v.x; acc[i][j][1] = v.y;
    }
}
// columns >= ncols are not stored (the padded tail of a narrow leading dimension)
template <int MT_, int NT_>
EMPC_DI void acc_store(const double (&acc)[MT_][NT_][2], double* Cb, int ldc, int ncols, int r, int c) {
#pragma unroll
  for (int i = 0; i < MT_; ++i)
#pragma unroll
    for (int j = 0; j < NT_; ++j)
      if (8 * j + 2 * c < ncols) *reinterpret_cast<double2*>(Cb + (8 * i + r) * ldc + 8 * j + 2 * c) = make_double2(acc[i][j][0], acc[i][j][1]);
}

template <class D>
__global__ void __launch_bounds__(32, EMPC_BW_WARPS_PER_SM) backward_kernel(Buffers bf, BwParams P) {
  using S = BwCfg<D>;
  constexpr int n = S::n, m = S::m, LDB = S::LDB, LDA = S::LDA, LDQ = S::LDQ;
  extern __shared__ __align__(16) double bw_sm[];
  double* sm = bw_sm;
  const int lane = threadIdx.x, fr = lane >> 2, fc = lane & 3;
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

  constexpr int LDF = S::LDF, PW = S::PW;
  double* sF = sm + S::oF; double* sV = sm + S::oV;
  double* sQxu = sm + S::oQxu; double* sQuu = sm + S::oQuu; double* sFTV = sm + S::oFTV;
  double* sK = sm + S::oK; double* sL = sm + S::oL; double* sLinv = sL + m * m;
  double* vec = sm + S::oVec;
  double* Qx = vec + S::vQx; double* Qu = vec + S::vQu; double* Vxp = vec + S::vVx; double* fsv = vec + S::vFs;
  double* gv = vec + S::vG; double* kv = vec + S::vKv; double* Quuk = vec + S::vQuuk; double* tmpv = vec + S::vTmp;
  double* Luud = vec + S::vLuu;

  // asynchronous fetch of Fx (16-byte pieces, n even) and Fu (8-byte pieces) of node t into their padded layouts
  auto fetch_F = [&](int t) {
    const double* tg = bf.tiles + (nb + t) * D::TILE;
    {  // Fx: n rows of n/2 16-byte pieces; piece e = lane + 32 q sits in row e / (n/2); indices advance incrementally
      constexpr int H = n / 2, DI = 32 / H, DC = 32 % H;
      int i = lane / H, cc = lane - i * H;
#pragma unroll
      for (int q = 0; q < (n * H + 31) / 32; ++q) {
        if (lane + 32 * q < n * H) cp_async16(sF + i * LDF + 2 * cc, tg + D::oFx + 2 * (lane + 32 * q));
        i += DI; cc += DC;
        if (cc >= H) { cc -= H; i += 1; }
      }
    }
    {  // Fu: n rows of m 8-byte pieces
      constexpr int DI = 32 / m, DC = 32 % m;
      int i = lane / m, j = lane - i * m;
#pragma unroll
      for (int q = 0; q < (n * m + 31) / 32; ++q) {
        if (lane + 32 * q < n * m) cp_async8(sF + i * LDF + n + j, tg + D::oFu + lane + 32 * q);
        i += DI; j += DC;
        if (j >= m) { j -= m; i += 1; }
      }
    }
    cp_async_commit();
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
      for (int e = lane; e < n * n; e += 32) { const int i = e / n, j = e - i * n; sV[i * LDB + j] = tg[D::oLxx + e] + ((i == j) ? xreg : 0.0); }
      if (lane < n) { Vxp[lane] = tg[D::oLx + lane]; fsv[lane] = bf.fs[(nb + T) * n + lane]; }
      __syncwarp();
      if (lane < n) {
        double s = 0;
#pragma unroll 6
        for (int j = 0; j < n; ++j) s += sV[lane * LDB + j] * fsv[j];
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
      if (lane < n) { bf.Vx[(nb + T) * n + lane] = Vxp[lane]; bf.g[(nb + T) * n + lane] = gv[lane]; }
      fetch_F(T - 1);
      double pre[S::PREF], pre_fs;
      load_L(T - 1, pre, pre_fs);
      __syncwarp();
      store_L(pre, pre_fs);
    }
    for (int t = T - 1; t >= 0; --t) {
      cp_async_wait<0>();
      __syncwarp();
      double pre[S::PREF], pre_fs = 0.0;
      if (t > 0) load_L(t - 1, pre, pre_fs);
      // Lxx of this node: HBM -> accumulator fragments (row 8 i + fr, columns 8 j + 2 fc, +1); in flight during the
      // first two products
      // (they are the top-left tiles of the packed accumulator q of [[Qxx, Qxu], [Qux, Quu]]; upper tiles only)
      double q[S::PT][S::PT][2];
      {
        const double* lg = bf.tiles + (nb + t) * D::TILE + D::oLxx;
#pragma unroll
        for (int i = 0; i < S::PT; ++i)
#pragma unroll
          for (int j = i; j < S::PT; ++j) {
            const int row = 8 * i + fr, col = 8 * j + 2 * fc;
            double2 v = make_double2(0.0, 0.0);
            if (row < n && col < n) v = *reinterpret_cast<const double2*>(lg + row * n + col);
            q[i][j][0] = v.x; q[i][j][1] = v.y;
          }
      }
      // ---- F^T V = [FxTV ; FuTV]  (one k-loop over the packed operand) ----
      {
        double ftv[S::PT][S::NT][2];
        acc_zero(ftv);
#pragma unroll
        for (int ks = 0; ks < S::KN; ++ks) {
          double fa[S::PT], vv[S::NT];
#pragma unroll
          for (int i = 0; i < S::PT; ++i) fa[i] = sF[(4 * ks + fc) * LDF + 8 * i + fr];
#pragma unroll
          for (int i = 0; i < S::NT; ++i) vv[i] = sV[(4 * ks + fc) * LDB + 8 * i + fr];
#pragma unroll
          for (int i = 0; i < S::PT; ++i)
#pragma unroll
            for (int j = 0; j < S::NT; ++j) dmma884(ftv[i][j][0], ftv[i][j][1], fa[i], vv[j]);
        }
        acc_store(ftv, sFTV, LDA, LDA, fr, fc);
      }
      // Qx += Fx^T Vx' ; Qu += Fu^T Vx'   (lane = column of the packed operand; three interleaved partial sums)
#pragma unroll
      for (int c0 = 0; c0 < PW; c0 += 32) {
        const int c = c0 + lane;
        const int cc = c < PW ? c : 0;
        double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
        for (int l = 0; l + 2 < n; l += 3) {
          s0 += sF[l * LDF + cc] * Vxp[l]; s1 += sF[(l + 1) * LDF + cc] * Vxp[l + 1]; s2 += sF[(l + 2) * LDF + cc] * Vxp[l + 2];
        }
#pragma unroll
        for (int l = n - n % 3; l < n; ++l) s0 += sF[l * LDF + cc] * Vxp[l];
        const double sacc = (s0 + s1) + s2;
        if (c < n) Qx[c] += sacc;
        else if (c < PW) Qu[c - n] += sacc;
      }
      __syncwarp();
      // ---- [[Qxx, Qxu], [., Quu]] = [[Lxx, 0], [0, Luu + ureg I]] + (F^T V) F, upper tiles of the packed symmetric matrix.
      // Qxx (the top-left tiles) stays in registers until Qxu K has been subtracted from it; Qxu and Quu go to shared memory. ----
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
          double fa[S::PT], fb[S::PT];
#pragma unroll
          for (int i = 0; i < S::PT; ++i) { fa[i] = sFTV[(8 * i + fr) * LDA + 4 * ks + fc]; fb[i] = sF[(4 * ks + fc) * LDF + 8 * i + fr]; }
#pragma unroll
          for (int i = 0; i < S::PT; ++i)
#pragma unroll
            for (int j = i; j < S::PT; ++j) dmma884(q[i][j][0], q[i][j][1], fa[i], fb[j]);
        }
        __syncwarp();  // every lane is done reading F^T V before the Quu factor (same storage) is written further down
#pragma unroll
        for (int i = 0; i < S::PT; ++i)
#pragma unroll
          for (int j = i; j < S::PT; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int row = 8 * i + fr, col = 8 * j + 2 * fc + h;
              if (col >= n && col < PW) {
                if (row < n) sQxu[row * LDQ + (col - n)] = q[i][j][h];
                else if (row < PW) {
                  sQuu[(row - n) * LDQ + (col - n)] = q[i][j][h];
                  if (i != j) sQuu[(col - n) * LDQ + (row - n)] = q[i][j][h];  // the lower tiles are not computed: mirror
                }
              }
            }
      }
      __syncwarp();
      // Fx, Fu of this node are dead: start fetching the next node's while Quu is factorised
      if (t > 0) fetch_F(t - 1);
      // ---- Cholesky of Quu (lane = row, left-looking: the reference LLT's subtraction order) ----
      int bad = 0;
      {
        const int i = lane < m ? lane : m - 1;
        double row[m];
#pragma unroll
        for (int k = 0; k < m; ++k) row[k] = sQuu[i * LDQ + k];
#pragma unroll
        for (int j = 0; j < m; ++j) {
          double d = sQuu[j * LDQ + j];
#pragma unroll
          for (int k = 0; k < j; ++k) { const double ljk = sL[j * m + k]; d -= ljk * ljk; }
          if (!(d > 0.0)) bad = 1;
          const double dinv = rsqrt_nr(d);
          double sij = row[j];
#pragma unroll
          for (int k = 0; k < j; ++k) sij -= row[k] * sL[j * m + k];
          sij = (i == j) ? d * dinv : sij * dinv;
          row[j] = sij;
          if (lane < m && i >= j) sL[i * m + j] = sij;
          if (lane == j) sLinv[j] = dinv;
          __syncwarp();
        }
      }
      if (bad) { failed = 1; break; }  // uniform: every lane evaluates every pivot
      // ---- gains: K = Quu^-1 Qxu^T (one right-hand side per lane), k = Quu^-1 Qu ----
      for (int c = lane; c < n + 1; c += 32) {
        double rhs[m];
        if (c < n) {
#pragma unroll
          for (int i = 0; i < m; ++i) rhs[i] = sQxu[c * LDQ + i];
        } else {
#pragma unroll
          for (int i = 0; i < m; ++i) rhs[i] = Qu[i];
        }
        // column-oriented substitution: entry i receives its subtractions in the same order (k ascending / descending)
        // as the row-oriented reference loops, but the dependent chain is one multiply + one FMA per column
#pragma unroll
        for (int kk = 0; kk < m; ++kk) {
          rhs[kk] *= sLinv[kk];
#pragma unroll
          for (int i = kk + 1; i < m; ++i) rhs[i] -= sL[i * m + kk] * rhs[kk];
        }
#pragma unroll
        for (int kk = m - 1; kk >= 0; --kk) {
          rhs[kk] *= sLinv[kk];
#pragma unroll
          for (int i = kk - 1; i >= 0; --i) rhs[i] -= sL[kk * m + i] * rhs[kk];
        }
        if (c < n) {
#pragma unroll
          for (int i = 0; i < m; ++i) sK[i * LDB + c] = rhs[i];
        } else {
#pragma unroll
          for (int i = 0; i < m; ++i) kv[i] = rhs[i];
        }
      }
      __syncwarp();
      if (lane < m) {  // Quuk = Quu k
        double s = 0;
#pragma unroll
        for (int j = 0; j < m; ++j) s += sQuu[lane * LDQ + j] * kv[j];
        Quuk[lane] = s;
      }
      __syncwarp();
      // Vx = Qx + K^T Quuk - 2 K^T Qu
      if (lane < n) {
        double s1 = 0, s2 = 0;
#pragma unroll
        for (int j = 0; j < m; ++j) { const double kji = sK[j * LDB + lane]; s1 += kji * Quuk[j]; s2 += kji * Qu[j]; }
        tmpv[lane] = Qx[lane] + s1 - 2 * s2;
      }
      // ---- Vxx = sym(Qxx - Qxu K) + xreg I, all in the accumulator registers.  Qxx and Qxu K are symmetric up to
      // rounding, so only the upper tiles (j >= i) are computed: diagonal tiles are averaged with their own transpose
      // (the reference's 0.5 (Vxx + Vxx^T)), the lower tiles are the mirror images of the upper ones.  Entries of the
      // top-left tiles beyond row / column n belong to Qxu / Quu and are masked: V's padding has to stay zero. ----
#pragma unroll
      for (int ks = 0; ks < S::KM; ++ks) {
        double a[S::NT], bq[S::NT];
#pragma unroll
        for (int i = 0; i < S::NT; ++i) { a[i] = -sQxu[(8 * i + fr) * LDQ + 4 * ks + fc]; bq[i] = sK[(4 * ks + fc) * LDB + 8 * i + fr]; }
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
        acc_store(vsym, sV, LDB, LDB, fr, fc);
      }
      __syncwarp();
      if (lane < n) {
        double s0 = 0, s1 = 0, s2 = 0;  // row `lane` of the symmetric V read as a column: conflict-free
#pragma unroll
        for (int j = 0; j + 2 < n; j += 3) {
          s0 += sV[j * LDB + lane] * fsv[j]; s1 += sV[(j + 1) * LDB + lane] * fsv[j + 1]; s2 += sV[(j + 2) * LDB + lane] * fsv[j + 2];
        }
#pragma unroll
        for (int j = n - n % 3; j < n; ++j) s0 += sV[j * LDB + lane] * fsv[j];
        const double s = (s0 + s1) + s2;
        gv[lane] = s;
        const double vx = feasible ? tmpv[lane] : (tmpv[lane] + s);
        if (raise_if_nan_abs(vx)) bad = 1;  // raiseIfNaN(Vx.lpNorm<Infinity>())
        Vxp[lane] = vx;
      }
      bad = __any_sync(0xffffffffu, bad);
      if (bad) { failed = 1; break; }
      __syncwarp();
      // outputs
      {
        double* Kg = bf.K + ((size_t)b * T + t) * m * n;
        for (int e2 = lane; e2 < m * n / 2; e2 += 32) {
          const int e = 2 * e2, i = e / n, j = e - i * n;  // n even: a pair never straddles rows
          reinterpret_cast<double2*>(Kg)[e2] = *reinterpret_cast<const double2*>(sK + i * LDB + j);
        }
        double* kg = bf.k + ((size_t)b * T + t) * m;
        if (lane < m) kg[lane] = kv[lane];
        if (lane < n) { bf.Vx[(nb + t) * n + lane] = Vxp[lane]; bf.g[(nb + t) * n + lane] = gv[lane]; }
        {  // five ordered dot products on lanes 24..28: same unrolled, predicated code for all of them (loads hoisted)
          const int w = lane & 7;
          const double* pa = (w == 0) ? Qu : (w == 1) ? kv : (w == 2) ? Vxp : (w == 3) ? fsv : Qu;
          const double* pb = (w == 0) ? kv : (w == 1) ? Quuk : (w == 2) ? fsv : (w == 3) ? gv : Qu;
          const int cnt = (w == 2 || w == 3) ? n : m;
          double sacc = 0;
#pragma unroll
          for (int i = 0; i < n; ++i) { const double av = (i < cnt) ? pa[i] : 0.0, bv = (i < cnt) ? pb[i] : 0.0; sacc += av * bv; }
          if (lane >= 24 && w < 4) bf.nodesc[(nb + t) * 4 + w] = sacc;
          if (lane == 28) bf.qu2[nb + t] = sacc;  // ||Qu_t||^2 (upstream stoppingCriteria)
        }
      }
      __syncwarp();
      if (t > 0) store_L(pre, pre_fs);
    }
    cp_async_wait<0>();
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
