// contact.cuh — shooting nodes under contact dynamics (included from kernels.cuh inside namespace empc).
//
// Replaces, for the nodes of a stage that declares a contact, crocoddyl::DifferentialActionModelContactFwdDynamics
// (src/factory/diff-action.cpp:30-32) with one ContactModel3D / ContactModel6D (src/factory/contacts.cpp:32-81, zero
// Baumgarte gains as in the corpus) under IntegratedActionModelEuler, and the friction-cone cost on the contact force
// (ResidualModelContactFrictionCone, src/factory/cost.cpp:149-167):
//   calc      pinocchio::forwardDynamics:  [M Jc^T; Jc 0] [a; -lambda] = [tau - h; -a0]
//   calcDiff  implicit differentiation of that system (getKKTContactDynamicMatrixInverse, computeRNEADerivatives with
//             the contact force as external force, getJointAccelerationDerivatives for the constrained acceleration).
//
// Design.  Contact nodes are a minority of a trajectory (10 of 160 knots in eagle_catch.yaml) and their working set —
// the joint-space inertia, its inverse and the KKT blocks next to everything a free node needs — does not fit the
// lanes-per-node layout the free nodes are tuned for.  So the free-node kernels (calcdiff.cuh) stay untouched and treat a
// contact node like a free one; contact_node_kernel then runs over the contact nodes only, one thread per node out of
// local memory, and overwrites xnext, the gap, the node cost and the whole node tile — including Lxu and the full Luu,
// which only a contact-force cost makes non-zero (backward_kernel<D, true> adds them).  The rollout chain calls
// node_dyn_contact for these nodes, decide / trial_cost evaluate the friction-cone residual through contact_force.
// Everything here is written with rolled loops over the (compile-time) dimensions: compact code, no register pressure
// on the callers (every entry point is __noinline__).
#pragma once

#define EMPC_ROLLED _Pragma("unroll 1")

template <class D>
struct ContactWork {
  double J[D::NV][6];       // world motion axis of velocity column c
  double ov[D::NJ][6];      // world spatial velocities of the bodies
  double Minv[D::NV * D::NV];
  SE3 oMf;                  // world placement of the contact frame
  double fJ[D::NV][6];      // LOCAL frame Jacobian, column c (zero for joints below the frame's joint)
  double vf[6];             // LOCAL frame velocity
  double lam[6];            // contact force in contact-frame coordinates (pinocchio lambda_c)
  double Bc[D::NV][6];      // Minv Jc^T
  double Ginv[36];          // (Jc Minv Jc^T)^-1, nc x nc
  int nc, jf;               // constraint rows (3 / 6), joint carrying the contact frame
};

EMPC_DI int cw_joint(int c) { return c < 6 ? 0 : c - 5; }  // joint owning velocity column c (serial chain)

// inverse of a symmetric positive definite n x n matrix (n <= 13) by Cholesky; rolled loops
__device__ __noinline__ bool spd_inverse(const double* A, int n, double* Ainv) {
  double L[13 * 13], b[13];
  bool ok = true;
  EMPC_ROLLED for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    EMPC_ROLLED for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
    if (!(d > 0.0)) ok = false;
    d = sqrt(d);
    L[j * n + j] = d;
    EMPC_ROLLED for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      EMPC_ROLLED for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = s / d;
    }
  }
  EMPC_ROLLED for (int c = 0; c < n; ++c) {
    EMPC_ROLLED for (int i = 0; i < n; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      EMPC_ROLLED for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k];
      b[i] = s / L[i * n + i];
    }
    EMPC_ROLLED for (int i = n - 1; i >= 0; --i) {
      double s = b[i];
      EMPC_ROLLED for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k];
      b[i] = s / L[i * n + i];
    }
    EMPC_ROLLED for (int i = 0; i < n; ++i) Ainv[i * n + c] = b[i];
  }
  return ok;
}

template <class D>
__device__ __noinline__ void cw_world_kinematics(const DevModel& M, const NodeData<D>& nd, ContactWork<D>& cw) {
  EMPC_ROLLED for (int b = 0; b < 6; ++b) {
    double e[6] = {0, 0, 0, 0, 0, 0};
    e[b] = 1.0;
    act_motion(nd.oM[0], e, cw.J[b]);
  }
  EMPC_ROLLED for (int i = 1; i < D::NJ; ++i) {
    const double S[6] = {0, 0, 0, M.axis[i][0], M.axis[i][1], M.axis[i][2]};
    act_motion(nd.oM[i], S, cw.J[5 + i]);
  }
  EMPC_ROLLED for (int i = 0; i < D::NJ; ++i) act_motion(nd.oM[i], nd.v[i], cw.ov[i]);
}

// world spatial accelerations for joint velocities vq / accelerations aq on top of the base acceleration a_base (the
// gravity field M.a0 for inverse dynamics, zero for the kinematic acceleration): A_i = A_(i-1) + S_i aq_i + V_(i-1) x S_i vq_i
template <class D>
__device__ __noinline__ void cw_world_accel(const ContactWork<D>& cw, const double* vq, const double* aq, const double* a_base,
                                            double (*oa)[6]) {
  EMPC_ROLLED for (int k = 0; k < 6; ++k) {
    double s = a_base[k];
    EMPC_ROLLED for (int c = 0; c < 6; ++c) s += cw.J[c][k] * aq[c];
    oa[0][k] = s;
  }
  EMPC_ROLLED for (int i = 1; i < D::NJ; ++i) {
    const int c = 5 + i;
    double cr[6];
    cross_mm(cw.ov[i - 1], cw.J[c], cr);
    EMPC_ROLLED for (int k = 0; k < 6; ++k) oa[i][k] = oa[i - 1][k] + cw.J[c][k] * aq[c] + cr[k] * vq[c];
  }
}

// joint-space inertia by world-frame composite rigid bodies: M(cj, ck) = J_cj . Ycrb_k J_ck for joint(cj) <= joint(ck)
template <class D>
__device__ __noinline__ void cw_crba(const DevModel& M, const NodeData<D>& nd, const ContactWork<D>& cw, double* Mjs) {
  constexpr int NJ = D::NJ, NV = D::NV;
  double oY[NJ][36];
  EMPC_ROLLED for (int i = 0; i < NJ; ++i) {
    double X[36];
    force_action_matrix(nd.oM[i], X);
    congruence6(X, M.Y[i], oY[i]);
  }
  EMPC_ROLLED for (int i = NJ - 1; i > 0; --i)
    EMPC_ROLLED for (int k = 0; k < 36; ++k) oY[i - 1][k] += oY[i][k];
  EMPC_ROLLED for (int ck = 0; ck < NV; ++ck) {
    double YJ[6];
    mat6_vec(oY[cw_joint(ck)], cw.J[ck], YJ);
    EMPC_ROLLED for (int cj = 0; cj < NV; ++cj)
      if (cw_joint(cj) <= cw_joint(ck)) {
        const double v = dot6(cw.J[cj], YJ);
        Mjs[cj * NV + ck] = v;
        Mjs[ck * NV + cj] = v;
      }
  }
}

// d tau/dq, d tau/dv of inverse dynamics at the world accelerations oa (gravity field included), with the external
// force Fext (world spatial force, fixed in the local frame of joint jext, acting on the robot: tau = RNEA - J^T F);
// pinocchio::computeRNEADerivatives(q, v, a, fext).  Same recursion as DESIGN.md "ABA derivatives", one thread.
template <class D>
__device__ __noinline__ void cw_rnea_partials(const DevModel& M, const NodeData<D>& nd, const ContactWork<D>& cw, const double (*oa)[6],
                                              const double* Fext, int jext, double* dq, double* dv) {
  constexpr int NJ = D::NJ, NV = D::NV;
  double oY[NJ][36], Bm[NJ][36], F[NJ][6];
  EMPC_ROLLED for (int i = 0; i < NJ; ++i) {
    double X[36];
    force_action_matrix(nd.oM[i], X);
    congruence6(X, M.Y[i], oY[i]);
    double h[6], Ya[6], vh[6];
    mat6_vec(oY[i], cw.ov[i], h); mat6_vec(oY[i], oa[i], Ya); cross_mf(cw.ov[i], h, vh);
    EMPC_ROLLED for (int k = 0; k < 6; ++k) F[i][k] = Ya[k] + vh[k] - ((i == jext) ? Fext[k] : 0.0);
    // B_i = crf(v) Y - Y crm(v) + Hx(h): column b of crm(v) is v x e_b, row a of crf(v) Y is (v x* Y e_b)_a
    EMPC_ROLLED for (int bcol = 0; bcol < 6; ++bcol) {
      double e[6] = {0, 0, 0, 0, 0, 0}, Ye[6], t1[6], ve[6], t2[6];
      e[bcol] = 1.0;
      mat6_vec(oY[i], e, Ye); cross_mf(cw.ov[i], Ye, t1);    // crf(v) Y e_b
      cross_mm(cw.ov[i], e, ve); mat6_vec(oY[i], ve, t2);    // Y crm(v) e_b
      EMPC_ROLLED for (int a = 0; a < 6; ++a) Bm[i][6 * a + bcol] = t1[a] - t2[a];
    }
    double Shf[9], Shn[9];
    skew3(h, Shf); skew3(h + 3, Shn);
    EMPC_ROLLED for (int a = 0; a < 3; ++a)
      EMPC_ROLLED for (int b2 = 0; b2 < 3; ++b2) {
        Bm[i][6 * a + 3 + b2] -= Shf[3 * a + b2];
        Bm[i][6 * (3 + a) + b2] -= Shf[3 * a + b2];
        Bm[i][6 * (3 + a) + 3 + b2] -= Shn[3 * a + b2];
      }
  }
  EMPC_ROLLED for (int i = NJ - 1; i > 0; --i) {
    EMPC_ROLLED for (int k = 0; k < 36; ++k) { oY[i - 1][k] += oY[i][k]; Bm[i - 1][k] += Bm[i][k]; }
    EMPC_ROLLED for (int k = 0; k < 6; ++k) F[i - 1][k] += F[i][k];
  }
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  EMPC_ROLLED for (int ck = 0; ck < NV; ++ck) {
    const int k = cw_joint(ck);
    const double* s = cw.J[ck];
    const double* vp = k > 0 ? cw.ov[k - 1] : zero6;
    const double* ap = k > 0 ? oa[k - 1] : M.a0;
    double dVdq[6], dAdq[6], dAdv[6], t6[6], vsum[6];
    cross_mm(vp, s, dVdq);
    cross_mm(ap, s, dAdq); cross_mm(vp, dVdq, t6);
    EMPC_ROLLED for (int a = 0; a < 6; ++a) { dAdq[a] += t6[a]; vsum[a] = vp[a] + cw.ov[k][a]; }
    cross_mm(vsum, s, dAdv);
    double P[6], Fq[6], Fv[6], t1[6], t2[6];
    mat6_vec(oY[k], dAdq, t1); mat6_vec(Bm[k], dVdq, t2);
    EMPC_ROLLED for (int a = 0; a < 6; ++a) P[a] = t1[a] + t2[a];
    cross_mf(s, F[k], t1);
    EMPC_ROLLED for (int a = 0; a < 6; ++a) Fq[a] = P[a] + t1[a];
    mat6_vec(oY[k], dAdv, t1); mat6_vec(Bm[k], s, t2);
    EMPC_ROLLED for (int a = 0; a < 6; ++a) Fv[a] = t1[a] + t2[a];
    EMPC_ROLLED for (int cj = 0; cj < NV; ++cj) {
      const int j = cw_joint(cj);
      double vq_, vv_;
      if (j == k) { vq_ = dot6(cw.J[cj], P); vv_ = dot6(cw.J[cj], Fv); }
      else if (j < k) { vq_ = dot6(cw.J[cj], Fq); vv_ = dot6(cw.J[cj], Fv); }
      else {
        double YJ[6], BtJ[6];
        mat6_vec(oY[j], cw.J[cj], YJ); mat6T_vec(Bm[j], cw.J[cj], BtJ);
        vq_ = dot6(YJ, dAdq) + dot6(BtJ, dVdq);
        vv_ = dot6(YJ, dAdv) + dot6(BtJ, s);
      }
      dq[cj * NV + ck] = vq_; dv[cj * NV + ck] = vv_;
    }
  }
}

template <class D>
EMPC_DI void cw_frame(const DevModel& M, const NodeData<D>& nd, int f, SE3& oMf) {
  SE3 fM;
  EMPC_ROLLED for (int i = 0; i < 9; ++i) fM.R[i] = M.fR[f][i];
  EMPC_ROLLED for (int i = 0; i < 3; ++i) fM.p[i] = M.fp[f][i];
  se3_mul(nd.oM[M.frame_joint[f]], fM, oMf);
}

// Follows aba(): nd.a holds the free acceleration on entry and the constrained one on exit; cw keeps what the
// derivatives reuse.
template <class D>
__device__ __noinline__ void contact_calc_dev(const DevModel& M, const empc_contact_t& ct, const double* x, NodeData<D>& nd,
                                              ContactWork<D>& cw) {
  constexpr int NV = D::NV;
  const double* vq = x + D::NQ;
  const int nc = (ct.type == EMPC_CONTACT_6D) ? 6 : 3;
  const int jf = M.frame_joint[ct.frame];
  cw.nc = nc; cw.jf = jf;
  cw_world_kinematics<D>(M, nd, cw);
  {
    double Mjs[NV * NV];
    cw_crba<D>(M, nd, cw, Mjs);
    spd_inverse(Mjs, NV, cw.Minv);
  }
  cw_frame<D>(M, nd, ct.frame, cw.oMf);
  EMPC_ROLLED for (int c = 0; c < NV; ++c) {
    if (cw_joint(c) <= jf) actinv_motion(cw.oMf, cw.J[c], cw.fJ[c]);
    else { EMPC_ROLLED for (int k = 0; k < 6; ++k) cw.fJ[c][k] = 0.0; }
  }
  // drift: the constrained frame acceleration at zero joint acceleration
  double a0[6];
  {
    double zero[NV], z6[6] = {0, 0, 0, 0, 0, 0}, oa[D::NJ][6], af[6];
    EMPC_ROLLED for (int i = 0; i < NV; ++i) zero[i] = 0.0;
    cw_world_accel<D>(cw, vq, zero, z6, oa);
    actinv_motion(cw.oMf, cw.ov[jf], cw.vf);
    actinv_motion(cw.oMf, oa[jf], af);
    EMPC_ROLLED for (int k = 0; k < 6; ++k) a0[k] = af[k];
    if (nc == 3) {
      double cr[3];
      cross3(cw.vf + 3, cw.vf, cr);
      EMPC_ROLLED for (int k = 0; k < 3; ++k) a0[k] += cr[k];
    }
  }
  double G[36];
  EMPC_ROLLED for (int i = 0; i < NV; ++i)
    EMPC_ROLLED for (int r = 0; r < nc; ++r) {
      double s = 0;
      EMPC_ROLLED for (int k = 0; k < NV; ++k) s += cw.Minv[i * NV + k] * cw.fJ[k][r];
      cw.Bc[i][r] = s;
    }
  EMPC_ROLLED for (int r = 0; r < nc; ++r)
    EMPC_ROLLED for (int c = 0; c < nc; ++c) {
      double s = 0;
      EMPC_ROLLED for (int i = 0; i < NV; ++i) s += cw.fJ[i][r] * cw.Bc[i][c];
      G[r * nc + c] = s;
    }
  spd_inverse(G, nc, cw.Ginv);
  double rhs[6];
  EMPC_ROLLED for (int r = 0; r < nc; ++r) {
    double s = a0[r];
    EMPC_ROLLED for (int i = 0; i < NV; ++i) s += cw.fJ[i][r] * nd.a[i];
    rhs[r] = s;
  }
  EMPC_ROLLED for (int r = 0; r < 6; ++r) cw.lam[r] = 0.0;
  EMPC_ROLLED for (int r = 0; r < nc; ++r) {
    double s = 0;
    EMPC_ROLLED for (int c = 0; c < nc; ++c) s += cw.Ginv[r * nc + c] * rhs[c];
    cw.lam[r] = -s;
  }
  EMPC_ROLLED for (int i = 0; i < NV; ++i) {
    double s = 0;
    EMPC_ROLLED for (int r = 0; r < nc; ++r) s += cw.Bc[i][r] * cw.lam[r];
    nd.a[i] += s;
  }
}

// [da; -dlambda] = -Kinv [d tau_rnea/dz ; d alpha/dz],  Kinv = [[P, B Ginv], [Ginv B^T, -Ginv]],  P = Minv - B Ginv B^T.
// Outputs a_q, a_v, P (NV x NV), lam_q, lam_v (nc x NV), GiBt = Ginv B^T (nc x NV).
template <class D>
__device__ __noinline__ void contact_derivs_dev(const DevModel& M, const double* x, const NodeData<D>& nd, const ContactWork<D>& cw,
                                                double* a_q, double* a_v, double* P, double* lam_q, double* lam_v, double* GiBt) {
  constexpr int NV = D::NV, NJ = D::NJ;
  const double* vq = x + D::NQ;
  const int nc = cw.nc, jf = cw.jf;
  double oa[NJ][6];
  cw_world_accel<D>(cw, vq, nd.a, M.a0, oa);
  double fl[6], Fw[6];
  EMPC_ROLLED for (int k = 0; k < 6; ++k) fl[k] = (k < nc) ? cw.lam[k] : 0.0;
  act_force(cw.oMf, fl, Fw);
  double dq[NV * NV], dv[NV * NV];
  cw_rnea_partials<D>(M, nd, cw, oa, Fw, jf, dq, dv);
  double dal_q[6][NV], dal_v[6][NV];
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  EMPC_ROLLED for (int c = 0; c < NV; ++c) {
    const int i = cw_joint(c);
    if (i > jf) { EMPC_ROLLED for (int k = 0; k < 6; ++k) { dal_q[k][c] = 0.0; dal_v[k][c] = 0.0; } continue; }
    const double* s = cw.J[c];
    const double* Vp = i > 0 ? cw.ov[i - 1] : zero6;
    const double* Vd = cw.ov[jf];
    double Ap[6];
    EMPC_ROLLED for (int k = 0; k < 6; ++k) Ap[k] = i > 0 ? oa[i - 1][k] - M.a0[k] : 0.0;
    double dV[6], t1[6], t2[6], t3[6], aqw[6], vs[6], avw[6];
    cross_mm(Vp, s, dV);
    cross_mm(Ap, s, t1); cross_mm(Vp, dV, t2); cross_mm(dV, Vd, t3);
    EMPC_ROLLED for (int k = 0; k < 6; ++k) { aqw[k] = t1[k] + t2[k] + t3[k]; vs[k] = cw.ov[i][k] + Vp[k] - Vd[k]; }
    cross_mm(vs, s, avw);
    double vql[6], aql[6], avl[6];
    actinv_motion(cw.oMf, dV, vql); actinv_motion(cw.oMf, aqw, aql); actinv_motion(cw.oMf, avw, avl);
    const double* fJc = cw.fJ[c];
    if (nc == 3) {
      double c1[3], c2[3], c3[3], c4[3];
      cross3(vql + 3, cw.vf, c1); cross3(cw.vf + 3, vql, c2);
      cross3(fJc + 3, cw.vf, c3); cross3(cw.vf + 3, fJc, c4);
      EMPC_ROLLED for (int k = 0; k < 3; ++k) { dal_q[k][c] = aql[k] + c1[k] + c2[k]; dal_v[k][c] = avl[k] + c3[k] + c4[k]; }
    } else {
      EMPC_ROLLED for (int k = 0; k < 6; ++k) { dal_q[k][c] = aql[k]; dal_v[k][c] = avl[k]; }
    }
  }
  EMPC_ROLLED for (int r = 0; r < nc; ++r)
    EMPC_ROLLED for (int i = 0; i < NV; ++i) {
      double s = 0;
      EMPC_ROLLED for (int c = 0; c < nc; ++c) s += cw.Ginv[r * nc + c] * cw.Bc[i][c];
      GiBt[r * NV + i] = s;
    }
  EMPC_ROLLED for (int i = 0; i < NV; ++i)
    EMPC_ROLLED for (int j = 0; j < NV; ++j) {
      double s = cw.Minv[i * NV + j];
      EMPC_ROLLED for (int r = 0; r < nc; ++r) s -= cw.Bc[i][r] * GiBt[r * NV + j];
      P[i * NV + j] = s;
    }
  EMPC_ROLLED for (int i = 0; i < NV; ++i)
    EMPC_ROLLED for (int j = 0; j < NV; ++j) {
      double sq = 0, sv = 0;
      EMPC_ROLLED for (int k = 0; k < NV; ++k) { sq += P[i * NV + k] * dq[k * NV + j]; sv += P[i * NV + k] * dv[k * NV + j]; }
      EMPC_ROLLED for (int r = 0; r < nc; ++r) { sq += GiBt[r * NV + i] * dal_q[r][j]; sv += GiBt[r * NV + i] * dal_v[r][j]; }
      a_q[i * NV + j] = -sq; a_v[i * NV + j] = -sv;
    }
  EMPC_ROLLED for (int r = 0; r < nc; ++r)
    EMPC_ROLLED for (int j = 0; j < NV; ++j) {
      double sq = 0, sv = 0;
      EMPC_ROLLED for (int k = 0; k < NV; ++k) { sq += GiBt[r * NV + k] * dq[k * NV + j]; sv += GiBt[r * NV + k] * dv[k * NV + j]; }
      EMPC_ROLLED for (int c = 0; c < nc; ++c) { sq -= cw.Ginv[r * nc + c] * dal_q[c][j]; sv -= cw.Ginv[r * nc + c] * dal_v[c][j]; }
      lam_q[r * NV + j] = sq; lam_v[r * NV + j] = sv;
    }
}

// squash -> thrust map -> aba -> contact solve; leaves the kinematics in nd, the constrained acceleration in nd.a
template <class D>
EMPC_DI void contact_forward(const DevModel& M, const empc_contact_t& ct, double smooth, const double* x, const double* u,
                             NodeData<D>& nd, ContactWork<D>& cw) {
  squash<D>(M, smooth, u, nd.s);
  double tau[D::NV];
  EMPC_ROLLED for (int i = 0; i < 6; ++i) {
    double t = 0;
    EMPC_ROLLED for (int j = 0; j < D::NR; ++j) t += M.tau_f[i * D::NR + j] * nd.s[j];
    tau[i] = t;
  }
  EMPC_ROLLED for (int i = 0; i < D::NA; ++i) tau[6 + i] = nd.s[D::NR + i];
  // (the fully unrolled variant: with the model in the kernel's parameter space the rolled joint loops of
  //  aba_dynamics<D, false> returned NaN accelerations here — static indices only, as in every other kernel that takes the
  //  model as a __grid_constant__)
  aba<D, true>(M, x, tau, nd);
  contact_calc_dev<D>(M, ct, x, nd, cw);
}

// Dynamics half of calc for a contact node: the rollout chain's replacement of node_dyn (node.cuh)
template <class D>
__device__ __noinline__ void node_dyn_contact(const DevModel& M, const empc_contact_t* ctp, double smooth, const double* x, const double* u,
                                              double* xnext) {
  const empc_contact_t ct = *ctp;
  NodeData<D> nd;
  ContactWork<D> cw;
  contact_forward<D>(M, ct, smooth, x, u, nd, cw);
  const double dt = M.dt, dt2 = dt * dt;
  EMPC_ROLLED for (int i = 0; i < D::NV; ++i) {
    nd.dx[i] = x[D::NQ + i] * dt + nd.a[i] * dt2;
    nd.dx[D::NV + i] = nd.a[i] * dt;
  }
  state_integrate<D>(x, nd.dx, xnext);
}

// Contact force at (x, u): what the friction-cone residual of a trial node needs (decide_kernel, trial_cost_kernel)
template <class D>
__device__ __noinline__ void contact_force(const DevModel& M, const empc_contact_t* ctp, double smooth, const double* x, const double* u,
                                           double* lam) {
  const empc_contact_t ct = *ctp;
  NodeData<D> nd;
  ContactWork<D> cw;
  contact_forward<D>(M, ct, smooth, x, u, nd, cw);
  EMPC_ROLLED for (int k = 0; k < 6; ++k) lam[k] = cw.lam[k];
}

// J <- d integrate(x, dx)/d(dx) J (+ d integrate(x, dx)/dx when add_first) for a row-major NDX x ncols matrix J:
// StateMultibody::JintegrateTransport(second), then Jintegrate(first, addto) — rows 0..5 <- Jexp6(dx[0:6]) rows 0..5,
// += blockdiag(Ad(exp6(dx)^-1), I)
template <class D>
__device__ __noinline__ void jintegrate_apply_dev(const double* dx, double* J, int ncols, bool add_first) {
  double JeA[9], JeQ[9];
  Jexp6_blocks(dx, JeA, JeQ);
  auto Je = [&](int a, int k) -> double { return (a < 3) ? ((k < 3) ? JeA[3 * a + k] : JeQ[3 * a + k - 3]) : ((k < 3) ? 0.0 : JeA[3 * (a - 3) + k - 3]); };
  EMPC_ROLLED for (int c = 0; c < ncols; ++c) {
    double tmp[6];
    EMPC_ROLLED for (int a = 0; a < 6; ++a) {
      double s = 0;
      EMPC_ROLLED for (int k = 0; k < 6; ++k) s += Je(a, k) * J[k * ncols + c];
      tmp[a] = s;
    }
    EMPC_ROLLED for (int a = 0; a < 6; ++a) J[a * ncols + c] = tmp[a];
  }
  if (add_first) {
    SE3 E; exp6(dx, E);
    double Xs[36]; force_action_matrix(E, Xs);  // Ad(E^-1) = (X*)^T
    EMPC_ROLLED for (int a = 0; a < 6; ++a)
      EMPC_ROLLED for (int c = 0; c < 6; ++c) J[a * ncols + c] += Xs[6 * c + a];
    EMPC_ROLLED for (int i = 6; i < D::NDX; ++i) J[i * ncols + i] += 1.0;
  }
}

// Cost value and Gauss-Newton derivative blocks of one (stage of a) node, unscaled by the integrator: ACCUMULATES into
// Lx (NDX), Lu (NU), Lxx (NDX x NDX), Lxu (NDX x NU), Luu (NU x NU) and returns sum_c w_c a_c(r_c).  nd / cw hold the
// kinematics of x (cw.J, cw.ov: cw_world_kinematics); lam / lam_x / lam_u are the contact force and its Jacobians (only
// read by a friction-cone cost).
template <class D>
__device__ __noinline__ double node_cost_derivs(const DevModel& M, const CostTables& C, int costset, double smooth, const double* x,
                                                const double* u, const NodeData<D>& nd, const ContactWork<D>& cw, const double* lam,
                                                const double* lam_x, const double* lam_u, double* Lx, double* Lu, double* Lxx,
                                                double* Lxu, double* Luu) {
  constexpr int NV = D::NV, NDX = D::NDX, NU = D::NU;
  double csum = 0;
  const int c0 = C.costset_begin[costset], c1 = C.costset_begin[costset + 1];
  EMPC_ROLLED for (int c = c0; c < c1; ++c) {
    const empc_cost_t cs = C.costs[c];
    if (!cs.active) continue;
    const double wt = cs.weight;
    double r[NDX], Ar[NDX], Arr[NDX];
    SE3 rMf;
    if (cs.type == EMPC_COST_CONTACT_FRICTION_CONE) {
      csum += wt * friction_cone_eval(C, cs, lam, r, Ar, Arr);
      const double* A = C.pool + cs.ref_off;
      // Rx = A df/dx (5 x NDX), Ru = A df/du (5 x NU)
      double Rx[5][NDX], Ru[5][NU];
      EMPC_ROLLED for (int a = 0; a < 5; ++a) {
        EMPC_ROLLED for (int j = 0; j < NDX; ++j) Rx[a][j] = A[3 * a] * lam_x[j] + A[3 * a + 1] * lam_x[NDX + j] + A[3 * a + 2] * lam_x[2 * NDX + j];
        EMPC_ROLLED for (int j = 0; j < NU; ++j) Ru[a][j] = A[3 * a] * lam_u[j] + A[3 * a + 1] * lam_u[NU + j] + A[3 * a + 2] * lam_u[2 * NU + j];
      }
      EMPC_ROLLED for (int i = 0; i < NDX; ++i) {
        double s = 0;
        EMPC_ROLLED for (int k = 0; k < 5; ++k) s += Rx[k][i] * Ar[k];
        Lx[i] += wt * s;
        EMPC_ROLLED for (int j = 0; j < NDX; ++j) {
          double h = 0;
          EMPC_ROLLED for (int k = 0; k < 5; ++k) h += Rx[k][i] * (Arr[k] * Rx[k][j]);
          Lxx[i * NDX + j] += wt * h;
        }
        EMPC_ROLLED for (int j = 0; j < NU; ++j) {
          double h = 0;
          EMPC_ROLLED for (int k = 0; k < 5; ++k) h += Rx[k][i] * (Arr[k] * Ru[k][j]);
          Lxu[i * NU + j] += wt * h;
        }
      }
      EMPC_ROLLED for (int i = 0; i < NU; ++i) {
        double s = 0;
        EMPC_ROLLED for (int k = 0; k < 5; ++k) s += Ru[k][i] * Ar[k];
        Lu[i] += wt * s;
        EMPC_ROLLED for (int j = 0; j < NU; ++j) {
          double h = 0;
          EMPC_ROLLED for (int k = 0; k < 5; ++k) h += Ru[k][i] * (Arr[k] * Ru[k][j]);
          Luu[i * NU + j] += wt * h;
        }
      }
      continue;
    }
    csum += wt * cost_eval<D>(M, C, cs, smooth, x, u, nd, r, Ar, Arr, rMf);
    if (cs.type == EMPC_COST_STATE) {
      // Rx = Jdiff(xref, x, second) = blockdiag(Jlog6(Mref^-1 M), I)
      SE3 Mref, Mx, Dm;
      q_to_se3(C.pool + cs.ref_off, Mref); q_to_se3(x, Mx); se3_inv_mul(Mref, Mx, Dm);
      double Jl[36]; Jlog6(Dm, Jl);
      EMPC_ROLLED for (int i = 0; i < 6; ++i) {
        double s = 0;
        EMPC_ROLLED for (int k = 0; k < 6; ++k) s += Jl[6 * k + i] * Ar[k];
        Lx[i] += wt * s;
        EMPC_ROLLED for (int j = 0; j < 6; ++j) {
          double h = 0;
          EMPC_ROLLED for (int k = 0; k < 6; ++k) h += Jl[6 * k + i] * (Arr[k] * Jl[6 * k + j]);
          Lxx[i * NDX + j] += wt * h;
        }
      }
      EMPC_ROLLED for (int i = 6; i < NDX; ++i) { Lx[i] += wt * Ar[i]; Lxx[i * NDX + i] += wt * Arr[i]; }
    } else if (cs.type == EMPC_COST_CONTROL || cs.type == EMPC_COST_SQUASH_BARRIER) {
      EMPC_ROLLED for (int i = 0; i < NU; ++i) { Lu[i] += wt * Ar[i]; Luu[i * NU + i] += wt * Arr[i]; }
    } else {
      // frame costs: Rx = [Rq | Rv] from the LOCAL frame Jacobian
      const int f = cs.frame, jfc = M.frame_joint[f];
      const int nres = (cs.type == EMPC_COST_FRAME_PLACEMENT || cs.type == EMPC_COST_FRAME_VELOCITY) ? 6 : 3;
      SE3 oMf; cw_frame<D>(M, nd, f, oMf);
      double Rx[6][NDX];
      EMPC_ROLLED for (int a = 0; a < 6; ++a) EMPC_ROLLED for (int j = 0; j < NDX; ++j) Rx[a][j] = 0.0;
      double Jl[36];
      if (cs.type == EMPC_COST_FRAME_PLACEMENT) Jlog6(rMf, Jl);
      else if (cs.type == EMPC_COST_FRAME_ROTATION) { double wv[3], th; log3(rMf.R, wv, th); Jlog3(th, wv, Jl); }
      int ncols = NV;
      EMPC_ROLLED for (int cc = 0; cc < NV; ++cc) {
        const int k = cw_joint(cc);
        if (k > jfc) continue;
        double fj[6]; actinv_motion(oMf, cw.J[cc], fj);
        if (cs.type == EMPC_COST_FRAME_PLACEMENT) {
          EMPC_ROLLED for (int a = 0; a < 6; ++a) { double s = 0; EMPC_ROLLED for (int q = 0; q < 6; ++q) s += Jl[6 * a + q] * fj[q]; Rx[a][cc] = s; }
        } else if (cs.type == EMPC_COST_FRAME_ROTATION) {
          EMPC_ROLLED for (int a = 0; a < 3; ++a) Rx[a][cc] = Jl[3 * a] * fj[3] + Jl[3 * a + 1] * fj[4] + Jl[3 * a + 2] * fj[5];
        } else if (cs.type == EMPC_COST_FRAME_TRANSLATION) {
          EMPC_ROLLED for (int a = 0; a < 3; ++a) Rx[a][cc] = oMf.R[3 * a] * fj[0] + oMf.R[3 * a + 1] * fj[1] + oMf.R[3 * a + 2] * fj[2];
        } else {  // FRAME_VELOCITY (LOCAL): d v_f / dq_c = oMf.actInv(V_parent(c) x J_c), d v_f / dv_c = fJ_c
          if (k > 0) {
            double cr[6], o[6];
            cross_mm(cw.ov[k - 1], cw.J[cc], cr); actinv_motion(oMf, cr, o);
            EMPC_ROLLED for (int a = 0; a < 6; ++a) Rx[a][cc] = o[a];
          }
          EMPC_ROLLED for (int a = 0; a < 6; ++a) Rx[a][NV + cc] = fj[a];
          ncols = NDX;
        }
      }
      EMPC_ROLLED for (int i = 0; i < ncols; ++i) {
        double s = 0;
        EMPC_ROLLED for (int k = 0; k < nres; ++k) s += Rx[k][i] * Ar[k];
        Lx[i] += wt * s;
        EMPC_ROLLED for (int j = 0; j < ncols; ++j) {
          double h = 0;
          EMPC_ROLLED for (int k = 0; k < nres; ++k) h += Rx[k][i] * (Arr[k] * Rx[k][j]);
          Lxx[i * NDX + j] += wt * h;
        }
      }
    }
  }
  return csum;
}

// ---------------------------------------------------------------------------------------------------------------------
// calc + calcDiff of the contact nodes, one thread per node (early exit for every other node).  Runs after
// node_diff_kernel and overwrites what the free-node kernels left for these nodes.
#ifndef EMPC_CONTACT_THREADS
#define EMPC_CONTACT_THREADS 128
#define EMPC_CONTACT_MINB 8
#endif
template <class D>
__global__ void __launch_bounds__(EMPC_CONTACT_THREADS, EMPC_CONTACT_MINB) contact_node_kernel(Buffers bf, int force, double force_smooth, const __grid_constant__ DevModel M) {
  constexpr int NV = D::NV, NDX = D::NDX, NU = D::NU, NX = D::NX, NR = D::NR;
  const long long nl0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = bf.T + 1;
  if (nl0 >= (long long)bf.nb * T1) return;
  const size_t n = (size_t)bf.b0 * T1 + nl0;
  const int b = (int)(n / T1), t = (int)(n - (size_t)b * T1);
  const int costset = bf.node_costset[bf.ocp_map[b] * T1 + t];
  const int ci = bf.ct.costset_contact[costset];
  if (ci < 0) return;
  const OcpState st = bf.st[b];
  if (!force && (st.phase == PHASE_DONE || !st.recalc)) return;
  const double smooth = force ? force_smooth : st.smooth;
  const empc_contact_t ct = bf.ct.contacts[ci];

  double x[NX], u[NU];
  const double* xg = bf.xs + n * NX;
  EMPC_ROLLED for (int i = 0; i < NX; ++i) x[i] = xg[i];
  EMPC_ROLLED for (int i = 0; i < NU; ++i) u[i] = (t < bf.T) ? bf.us[((size_t)b * bf.T + t) * NU + i] : 0.0;

  NodeData<D> nd;
  ContactWork<D> cw;
  contact_forward<D>(M, ct, smooth, x, u, nd, cw);
  const double dt = M.dt, dt2 = dt * dt;
  double xn[NX];
  EMPC_ROLLED for (int i = 0; i < NV; ++i) {
    nd.dx[i] = x[D::NQ + i] * dt + nd.a[i] * dt2;
    nd.dx[NV + i] = nd.a[i] * dt;
  }
  state_integrate<D>(x, nd.dx, xn);
  EMPC_ROLLED for (int i = 0; i < NX; ++i) bf.xnext[n * NX + i] = xn[i];
  // gap of the next node (SolverDDP::calcDiff), as node_calc_kernel leaves it
  if (t < bf.T) {
    if (!st.is_feasible) {
      double x1[NX], f[NDX];
      EMPC_ROLLED for (int i = 0; i < NX; ++i) x1[i] = xg[NX + i];
      state_diff<D>(x1, xn, f);
      double gi = 0, g1 = 0;
      EMPC_ROLLED for (int i = 0; i < NDX; ++i) { bf.fs[(n + 1) * NDX + i] = f[i]; const double a = fabs(f[i]); gi = fmax(gi, a); g1 += a; if (isnan(a)) gi = a; }
      bf.gap_inf[n + 1] = gi; bf.gap_l1[n + 1] = g1;
    } else if (!st.was_feasible) {
      EMPC_ROLLED for (int i = 0; i < NDX; ++i) bf.fs[(n + 1) * NDX + i] = 0.0;
      bf.gap_inf[n + 1] = 0; bf.gap_l1[n + 1] = 0;
    }
  }

  // ---- derivatives of the dynamics ----
  double* tile = bf.tiles + n * D::TILE;
  double* Fx = tile + D::oFx; double* Fu = tile + D::oFu; double* Lxx = tile + D::oLxx; double* Lxu = tile + D::oLxu;
  double* Luu = tile + D::oLuu; double* Lx = tile + D::oLx; double* Lu = tile + D::oLu;
  double lam_x[6 * NDX], lam_u[6 * NU];
  {
    double a_q[NV * NV], a_v[NV * NV], P[NV * NV], lam_q[6 * NV], lam_v[6 * NV], GiBt[6 * NV];
    contact_derivs_dev<D>(M, x, nd, cw, a_q, a_v, P, lam_q, lam_v, GiBt);
    EMPC_ROLLED for (int r = 0; r < cw.nc; ++r)
      EMPC_ROLLED for (int j = 0; j < NV; ++j) { lam_x[r * NDX + j] = lam_q[r * NV + j]; lam_x[r * NDX + NV + j] = lam_v[r * NV + j]; }
    // squashing slopes, a_u = P A diag(ds), lam_u = -Ginv B^T A diag(ds),  A = [tau_f 0; 0 I]
    double ds[NU];
    EMPC_ROLLED for (int i = 0; i < NU; ++i) {
      ds[i] = 1.0;
      if (M.use_squash) {
        const double dd = (M.u_ub[i] - M.u_lb[i]) * smooth, a = dd * dd;
        const double l = u[i] - M.u_lb[i], h = u[i] - M.u_ub[i];
        ds[i] = 0.5 * (rsqrt_nr(a + l * l) * l - rsqrt_nr(a + h * h) * h);
      }
    }
    // Euler (euler.hxx calcDiff): rows of d(dx)/d(q, v, u) before the Lie-group transport
    EMPC_ROLLED for (int i = 0; i < NV; ++i) {
      EMPC_ROLLED for (int j = 0; j < NV; ++j) {
        Fx[i * NDX + j] = a_q[i * NV + j] * dt2;
        Fx[i * NDX + NV + j] = a_v[i * NV + j] * dt2 + ((i == j) ? dt : 0.0);
        Fx[(NV + i) * NDX + j] = a_q[i * NV + j] * dt;
        Fx[(NV + i) * NDX + NV + j] = a_v[i * NV + j] * dt;
      }
      EMPC_ROLLED for (int j = 0; j < NU; ++j) {
        double s = 0;
        if (j < NR) { EMPC_ROLLED for (int k = 0; k < 6; ++k) s += P[i * NV + k] * (M.tau_f[k * NR + j] * ds[j]); }
        else s = P[i * NV + 6 + (j - NR)] * ds[j];
        Fu[i * NU + j] = dt2 * s; Fu[(NV + i) * NU + j] = dt * s;
      }
    }
    EMPC_ROLLED for (int r = 0; r < cw.nc; ++r)
      EMPC_ROLLED for (int j = 0; j < NU; ++j) {
        double s = 0;
        if (j < NR) { EMPC_ROLLED for (int k = 0; k < 6; ++k) s += GiBt[r * NV + k] * (M.tau_f[k * NR + j] * ds[j]); }
        else s = GiBt[r * NV + 6 + (j - NR)] * ds[j];
        lam_u[r * NU + j] = -s;
      }
  }
  jintegrate_apply_dev<D>(nd.dx, Fx, NDX, true);
  jintegrate_apply_dev<D>(nd.dx, Fu, NU, false);

  // ---- costs: value and Gauss-Newton derivatives (CostModelSum::calc / calcDiff) ----
  EMPC_ROLLED for (int i = 0; i < NDX * NDX; ++i) Lxx[i] = 0.0;
  EMPC_ROLLED for (int i = 0; i < NDX * NU; ++i) Lxu[i] = 0.0;
  EMPC_ROLLED for (int i = 0; i < NU * NU; ++i) Luu[i] = 0.0;
  EMPC_ROLLED for (int i = 0; i < NDX; ++i) Lx[i] = 0.0;
  EMPC_ROLLED for (int i = 0; i < NU; ++i) Lu[i] = 0.0;
  const double csum = node_cost_derivs<D>(M, bf.ct, costset, smooth, x, u, nd, cw, cw.lam, lam_x, lam_u, Lx, Lu, Lxx, Lxu, Luu);
  bf.node_cost[n] = dt * csum;
  EMPC_ROLLED for (int i = 0; i < NDX * NDX; ++i) Lxx[i] *= dt;
  EMPC_ROLLED for (int i = 0; i < NDX * NU; ++i) Lxu[i] *= dt;
  EMPC_ROLLED for (int i = 0; i < NU * NU; ++i) Luu[i] *= dt;
  EMPC_ROLLED for (int i = 0; i < NDX; ++i) Lx[i] *= dt;
  EMPC_ROLLED for (int i = 0; i < NU; ++i) Lu[i] *= dt;
  bf.node_dense[n] = 1;  // the tile holds a dense Lxx: node_diff_kernel clears it if this node ever becomes a free one without frame costs
}
