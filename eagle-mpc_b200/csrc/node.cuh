// node.cuh — per-node action model on the device: the pieces one thread evaluates for one shooting node (squashing,
// ABA, Euler step, residuals and activations), FP64, sm_100a.  The derivative side lives in calcdiff.cuh.
//
// Replaces, for the model chain eagle-mpc instantiates (src/factory/diff-action.cpp:31-35, src/factory/int-action.cpp:26,
// src/trajectory.cpp:47-52), the per-node work crocoddyl does inside ShootingProblem::calc / calcDiff:
//   IntegratedActionModelEuler::calc/calcDiff  ->  DifferentialActionModelFreeFwdDynamics::calc/calcDiff
//   -> ActuationSquashingModel (SquashingModelSmoothSat + ActuationModelMultiCopterBase), pinocchio::aba,
//      pinocchio::computeABADerivatives, CostModelSum over CostModelResidual{State,Control,Frame*} with
//      Activation{Quad,WeightedQuad,QuadraticBarrier,WeightedQuadraticBarrier}   
// Invoked from the reference at src/sbfddp.cpp:244,332 (computeDirection -> calcDiff) and :264,:437 (rollout calc).
//
// The robot is a free-flyer base + a serial chain of NA revolute joints (all eagle-mpc platforms); template
// parameters make every loop bound a compile-time constant.
#pragma once
#include "../../include/empc_b200.h"
#include "spatial.cuh"

namespace empc {

struct DevModel {
  int nj, na, nq, nv, nx, ndx, nu, nr, T, tile, use_squash, n_frames;
  int oFx, oFu, oLxx, oLxu, oLuu, oLx, oLu, integrator;  // integrator: EMPC_INTEGRATOR_EULER / _RK4 (rk4.cuh)
  double dt;
  double jR[EMPC_MAX_JOINTS][9], jp[EMPC_MAX_JOINTS][3], axis[EMPC_MAX_JOINTS][3];
  double Y[EMPC_MAX_JOINTS][36];  // body spatial inertia in the joint frame
  double mass[EMPC_MAX_JOINTS], com[EMPC_MAX_JOINTS][3], Ic[EMPC_MAX_JOINTS][9];  // same, as (m, c, I_c)
  double a0[6];                   // -gravity (spatial)
  int frame_joint[EMPC_MAX_FRAMES];
  double fR[EMPC_MAX_FRAMES][9], fp[EMPC_MAX_FRAMES][3];
  double tau_f[6 * EMPC_MAX_ROTORS], u_lb[EMPC_MAX_NU], u_ub[EMPC_MAX_NU];
  double bar_lb[EMPC_MAX_NU], bar_ub[EMPC_MAX_NU];  // crocoddyl::ActivationBounds(u_lb,u_ub,beta=1)
  double barrier_weight;
};

struct CostTables {
  const empc_cost_t* costs;
  const double* pool;
  const int* costset_begin;
  // contact dynamics (contact.cuh); all null for a problem without contacts
  const empc_contact_t* contacts;
  const int* costset_contact;             // per cost set: contact of the model that owns it, -1 = free dynamics
  const unsigned char* costset_coupled;   // per cost set: holds a contact-force cost (Lxu and a full Luu exist)
};

template <int NA_, int NR_>
struct Dim {
  static constexpr int NA = NA_, NR = NR_, NJ = NA + 1, NV = 6 + NA, NQ = 7 + NA, NX = NQ + NV, NDX = 2 * NV,
                       NU = NR + NA;
  static constexpr int oFx = 0, oFu = oFx + NDX * NDX, oLxx = oFu + NDX * NU, oLxu = oLxx + NDX * NDX,
                       oLuu = oLxu + NDX * NU, oLx = oLuu + NU * NU, oLu = oLx + NDX, TILE0 = oLu + NU,
                       TILE = TILE0 + (TILE0 & 1);
};

// kinematic / dynamic quantities left by calc and reused by calcDiff (crocoddyl's "data")
template <class D>
struct NodeData {
  SE3 oM[D::NJ];
  double liR[D::NJ][9];  // rotation of the joint placement in the parent frame (its translation is the constant jp)
  double v[D::NJ][6];
  double agf[D::NJ][6];
  double a[D::NV];
  double dx[D::NDX];
  double s[D::NU];
};

// crocoddyl::raiseIfNaN(value): NaN, +-inf or value >= 1e30 ("forward_error" / "backward_error" of the solvers)
EMPC_DI bool raise_if_nan(double v) { return !(v < 1e30) || isinf(v); }
// ... applied to an infinity norm: max_i |v_i| trips it as soon as one entry does
EMPC_DI bool raise_if_nan_abs(double v) { return !(fabs(v) < 1e30); }
// ... the same test on the high word of the double, as an integer (|v| >= 1e30, +-inf and NaN all have a high word
// >= that of 1e30 once the sign is masked; the low word only moves the threshold by 2^-20 relative).  Integer compares
// have a quarter of the latency of FP64 ones: this variant sits on the per-node critical path of the rollout.
EMPC_DI int raise_bits(double v) { return (__double2hiint(v) & 0x7fffffff) - 0x46293e59; }  // >= 0: raise

// ---- StateMultibody ------------------------------------------------------------------------------------------------
EMPC_DI void q_to_se3(const double* q, SE3& M) {
  quat_to_R(q + 3, M.R);
  M.p[0] = q[0]; M.p[1] = q[1]; M.p[2] = q[2];
}
template <class D>
EMPC_DI void state_integrate(const double* x, const double* dx, double* out) {
  SE3 M0, E, M1;
  q_to_se3(x, M0);
  exp6(dx, E);
  se3_mul(M0, E, M1);
  double quat[4];
  R_to_quat(M1.R, quat);
  const double dotp = quat[0] * x[3] + quat[1] * x[4] + quat[2] * x[5] + quat[3] * x[6];
  if (dotp < 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) quat[i] = -quat[i];
  }
  const double n2 = quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3];
  const double alpha = (3 - n2) / 2;
  double o[D::NX];
  o[0] = M1.p[0]; o[1] = M1.p[1]; o[2] = M1.p[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) o[3 + i] = quat[i] * alpha;
#pragma unroll
  for (int i = 0; i < D::NA; ++i) o[7 + i] = x[7 + i] + dx[6 + i];
#pragma unroll
  for (int i = 0; i < D::NV; ++i) o[D::NQ + i] = x[D::NQ + i] + dx[D::NV + i];
#pragma unroll
  for (int i = 0; i < D::NX; ++i) out[i] = o[i];
}
template <class D>
EMPC_DI void state_diff(const double* x0, const double* x1, double* dx) {
  SE3 M0, M1, Dm;
  q_to_se3(x0, M0); q_to_se3(x1, M1);
  se3_inv_mul(M0, M1, Dm);
  log6(Dm, dx);
#pragma unroll
  for (int i = 0; i < D::NA; ++i) dx[6 + i] = x1[7 + i] - x0[7 + i];
#pragma unroll
  for (int i = 0; i < D::NV; ++i) dx[D::NV + i] = x1[D::NQ + i] - x0[D::NQ + i];
}

// ---- activations ----------------------------------------------------------------------------------------------------
template <int N>
EMPC_DI double activation(int type, const double* r, const double* w, const double* lb, const double* ub, double* Ar,
                          double* Arr) {
  double val = 0;
  if (type == EMPC_ACT_QUAD) {
#pragma unroll
    for (int i = 0; i < N; ++i) { val += r[i] * r[i]; Ar[i] = r[i]; Arr[i] = 1; }
    return 0.5 * val;
  } else if (type == EMPC_ACT_WEIGHTED_QUAD) {
#pragma unroll
    for (int i = 0; i < N; ++i) { const double wr = w[i] * r[i]; val += r[i] * wr; Ar[i] = wr; Arr[i] = w[i]; }
    return 0.5 * val;
  } else if (type == EMPC_ACT_QUAD_BARRIER) {
    double sl = 0, su = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double dl = r[i] - lb[i], du = r[i] - ub[i];
      const double l = dl < 0 ? dl : 0.0, uu = du > 0 ? du : 0.0;
      sl += l * l; su += uu * uu;
      Ar[i] = l + uu;
      Arr[i] = (dl <= 0) ? 1.0 : ((du >= 0) ? 1.0 : 0.0);
    }
    return 0.5 * sl + 0.5 * su;
  } else {
    double sl = 0, su = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double dl = r[i] - lb[i], du = r[i] - ub[i];
      const double l = (dl < 0 ? dl : 0.0) * w[i], uu = (du > 0 ? du : 0.0) * w[i];
      sl += l * l; su += uu * uu;
      Ar[i] = (l + uu) * w[i];
      Arr[i] = (dl <= 0) ? w[i] : ((du >= 0) ? w[i] : 0.0);
    }
    return 0.5 * sl + 0.5 * su;
  }
}

template <class D>
EMPC_DI void frame_placement(const DevModel& M, const NodeData<D>& nd, int f, SE3& oMf) {
  SE3 fM;
#pragma unroll
  for (int i = 0; i < 9; ++i) fM.R[i] = M.fR[f][i];
#pragma unroll
  for (int i = 0; i < 3; ++i) fM.p[i] = M.fp[f][i];
  const int j = M.frame_joint[f];
  SE3 oMj = nd.oM[0];
#pragma unroll
  for (int i = 1; i < D::NJ; ++i)
    if (i == j) oMj = nd.oM[i];
  se3_mul(oMj, fM, oMf);
}

// residual + activation of one cost.  r/Ar/Arr sized NDX.  rMf receives the SE3 error of placement/rotation costs.
template <class D>
EMPC_DI double cost_eval(const DevModel& M, const CostTables& C, const empc_cost_t& c, double smooth, const double* x,
                         const double* u, const NodeData<D>& nd, double* r, double* Ar, double* Arr, SE3& rMf) {
  const double* ref = C.pool + c.ref_off;
  const double* aw = C.pool + c.w_off;
  const double* lb = C.pool + c.lb_off;
  const double* ub = C.pool + c.ub_off;
  switch (c.type) {
    case EMPC_COST_STATE: {
      double xr[D::NX];
#pragma unroll
      for (int i = 0; i < D::NX; ++i) xr[i] = ref[i];
      state_diff<D>(xr, x, r);
      return activation<D::NDX>(c.activation, r, aw, lb, ub, Ar, Arr);
    }
    case EMPC_COST_CONTROL: {
#pragma unroll
      for (int i = 0; i < D::NU; ++i) r[i] = u[i] - ref[i];
      return activation<D::NU>(c.activation, r, aw, lb, ub, Ar, Arr);
    }
    case EMPC_COST_SQUASH_BARRIER: {
      double bw[D::NU];
#pragma unroll
      for (int i = 0; i < D::NU; ++i) {
        const double aux = smooth * (M.u_ub[i] - M.u_lb[i]);
        bw[i] = 1.0 / (aux * aux);
        r[i] = u[i];
      }
      return activation<D::NU>(EMPC_ACT_WEIGHTED_QUAD_BARRIER, r, bw, M.bar_lb, M.bar_ub, Ar, Arr);
    }
    case EMPC_COST_FRAME_PLACEMENT: {
      SE3 Mref, oMf;
#pragma unroll
      for (int i = 0; i < 9; ++i) Mref.R[i] = ref[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) Mref.p[i] = ref[9 + i];
      frame_placement<D>(M, nd, c.frame, oMf);
      se3_inv_mul(Mref, oMf, rMf);
      log6(rMf, r);
      return activation<6>(c.activation, r, aw, lb, ub, Ar, Arr);
    }
    case EMPC_COST_FRAME_ROTATION: {
      SE3 oMf; frame_placement<D>(M, nd, c.frame, oMf);
      double Rr[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) Rr[i] = ref[i];
      matTmul3(Rr, oMf.R, rMf.R);
      double th; log3(rMf.R, r, th);
      return activation<3>(c.activation, r, aw, lb, ub, Ar, Arr);
    }
    case EMPC_COST_FRAME_TRANSLATION: {
      SE3 oMf; frame_placement<D>(M, nd, c.frame, oMf);
#pragma unroll
      for (int i = 0; i < 3; ++i) r[i] = oMf.p[i] - ref[i];
      return activation<3>(c.activation, r, aw, lb, ub, Ar, Arr);
    }
    default: {  // EMPC_COST_FRAME_VELOCITY (LOCAL)
      const int f = c.frame, j = M.frame_joint[f];
      SE3 fM;
#pragma unroll
      for (int i = 0; i < 9; ++i) fM.R[i] = M.fR[f][i];
#pragma unroll
      for (int i = 0; i < 3; ++i) fM.p[i] = M.fp[f][i];
      double vj[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) vj[k] = nd.v[0][k];
#pragma unroll
      for (int i = 1; i < D::NJ; ++i)
        if (i == j) {
#pragma unroll
          for (int k = 0; k < 6; ++k) vj[k] = nd.v[i][k];
        }
      double vf[6]; actinv_motion(fM, vj, vf);
#pragma unroll
      for (int i = 0; i < 6; ++i) r[i] = vf[i] - ref[i];
      return activation<6>(c.activation, r, aw, lb, ub, Ar, Arr);
    }
  }
}

// SquashingModelSmoothSat::calc
template <class D>
EMPC_DI void squash(const DevModel& M, double smooth, const double* u, double* s) {
#pragma unroll
  for (int i = 0; i < D::NU; ++i) {
    if (M.use_squash) {
      const double lbv = M.u_lb[i], ubv = M.u_ub[i];
      const double dd = (ubv - lbv) * smooth, a = dd * dd;
      const double l = u[i] - lbv, h = u[i] - ubv;
      s[i] = 0.5 * (sqrt_nr(l * l + a) - sqrt_nr(h * h + a) + lbv + ubv);
    } else {
      s[i] = u[i];
    }
  }
}

// Rigid-body inertia (m, c, I_c) applied to a motion: [f;n] = Y [v;w]  (pinocchio::InertiaTpl::__mult__)
EMPC_DI void inertia_apply(double m, const double* c, const double* Ic, const double* v, double* o) {
  double cxw[3], f[3], Iw[3], cxf[3];
  cross3(c, v + 3, cxw);
#pragma unroll
  for (int i = 0; i < 3; ++i) f[i] = m * (v[i] - cxw[i]);
  matvec3(Ic, v + 3, Iw);
  cross3(c, f, cxf);
#pragma unroll
  for (int i = 0; i < 3; ++i) { o[i] = f[i]; o[3 + i] = Iw[i] + cxf[i]; }
}

// pinocchio::aba for free-flyer + serial revolute chain (local-frame three-pass recursion), split in two:
//   aba_kinematics  pass 1: placements, velocities, bias accelerations  (everything the costs need)
//   aba_dynamics    pass 2 + 3: articulated inertias tip to base, accelerations base to tip
// so that callers can evaluate the costs in between and let the world placements die before the register-hungry
// backward sweep.  The backward pass carries ONE articulated inertia (the contribution of the subtree, expressed in the
// current joint frame) instead of one per joint.  The translation of a joint placement li[i] is the model constant
// jp[i] (the revolute joint transform has no translation), so only the rotations are kept.
// FULL: unroll the joint loops completely, so that every NodeData / model access has a static index (registers and
// constant-bank operands instead of local memory); used where the register budget allows it.
template <class D, bool FULL = false>
EMPC_DI void aba_kinematics(const DevModel& M, const double* x, NodeData<D>& nd) {
  constexpr int NJ = D::NJ;
  constexpr int UNR = FULL ? NJ : 1;
  const double* vq = x + D::NQ;
  q_to_se3(x, nd.oM[0]);
#pragma unroll
  for (int k = 0; k < 9; ++k) nd.liR[0][k] = nd.oM[0].R[k];
#pragma unroll
  for (int k = 0; k < 6; ++k) { nd.v[0][k] = vq[k]; nd.agf[0][k] = 0; }
#pragma unroll UNR
  for (int i = 1; i < NJ; ++i) {
    const double th = x[6 + i];
    double ax[3] = {M.axis[i][0] * th, M.axis[i][1] * th, M.axis[i][2] * th};
    double Rj[9]; exp3(ax, Rj);
    matmul3(M.jR[i], Rj, nd.liR[i]);
    SE3 li;
#pragma unroll
    for (int k = 0; k < 9; ++k) li.R[k] = nd.liR[i][k];
#pragma unroll
    for (int k = 0; k < 3; ++k) li.p[k] = M.jp[i][k];
    se3_mul(nd.oM[i - 1], li, nd.oM[i]);
    double vJ[6] = {0, 0, 0, M.axis[i][0] * vq[5 + i], M.axis[i][1] * vq[5 + i], M.axis[i][2] * vq[5 + i]};
    double vp[6]; actinv_motion(li, nd.v[i - 1], vp);
#pragma unroll
    for (int k = 0; k < 6; ++k) nd.v[i][k] = vJ[k] + vp[k];
    cross_mm(nd.v[i], vJ, nd.agf[i]);
  }
}

template <class D, bool FULL = false>
EMPC_DI void aba_dynamics(const DevModel& M, const double* tau, NodeData<D>& nd) {
  constexpr int NJ = D::NJ, NV = D::NV;
  constexpr int UNR = FULL ? NJ : 1;
  double uu[NV];
  double Dinv[NJ], UDinv[NJ][6];
#pragma unroll
  for (int i = 0; i < NV; ++i) uu[i] = tau[i];
  // pass 2: tip to base with ONE running articulated inertia, kept as 3x3 blocks Ia = [[A, B],[B^T, C]] and moved to
  // the parent frame with the block form of X* Ia X*^T (X* = [[R,0],[[p]x R, R]]):
  //   A' = R A R^T,  B' = R B R^T - A'[p]x,  C' = R C R^T + [p]x B' - (R B R^T)^T [p]x
  double A[9], Bk[9], C[9], pA[6];
#pragma unroll
  for (int k = 0; k < 9; ++k) { A[k] = 0.0; Bk[k] = 0.0; C[k] = 0.0; }
#pragma unroll
  for (int k = 0; k < 6; ++k) pA[k] = 0.0;
#pragma unroll UNR
  for (int i = NJ - 1; i >= 1; --i) {
    {  // Ia += Y_i ; pA += v x* (Y v)
      double h[6], vh[6];
      inertia_apply(M.mass[i], M.com[i], M.Ic[i], nd.v[i], h);
      cross_mf(nd.v[i], h, vh);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          A[3 * a + b] += M.Y[i][6 * a + b]; Bk[3 * a + b] += M.Y[i][6 * a + 3 + b]; C[3 * a + b] += M.Y[i][6 * (3 + a) + 3 + b];
        }
#pragma unroll
      for (int k = 0; k < 6; ++k) pA[k] += vh[k];
    }
    const double ax[3] = {M.axis[i][0], M.axis[i][1], M.axis[i][2]};
    const int c = 5 + i;
    uu[c] -= ax[0] * pA[3] + ax[1] * pA[4] + ax[2] * pA[5];
    double U[6];  // U = Ia S, S = [0; ax]
    matvec3(Bk, ax, U); matvec3(C, ax, U + 3);
    Dinv[i] = rcp_nr(dot3(ax, U + 3));
#pragma unroll
    for (int k = 0; k < 6; ++k) UDinv[i][k] = U[k] * Dinv[i];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        A[3 * a + b] -= UDinv[i][a] * U[b];
        Bk[3 * a + b] -= UDinv[i][a] * U[3 + b];
        C[3 * a + b] -= UDinv[i][3 + a] * U[3 + b];
      }
    double pa[6];
    {  // pa = pA + Ia c_i + UDinv u_i
      const double* cv = nd.agf[i];
      double t1[3], t2[3], t3[3], t4[3];
      matvec3(A, cv, t1); matvec3(Bk, cv + 3, t2); matTvec3(Bk, cv, t3); matvec3(C, cv + 3, t4);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        pa[k] = pA[k] + t1[k] + t2[k] + UDinv[i][k] * uu[c];
        pa[3 + k] = pA[3 + k] + t3[k] + t4[k] + UDinv[i][3 + k] * uu[c];
      }
    }
    {  // carry to the parent frame
      SE3 li;
#pragma unroll
      for (int k = 0; k < 9; ++k) li.R[k] = nd.liR[i][k];
#pragma unroll
      for (int k = 0; k < 3; ++k) li.p[k] = M.jp[i][k];
      const double* R = li.R; const double* p = li.p;
      double T1[9], An[9], Bt[9], Cn[9], P[9];
      matmul3(R, A, T1);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) An[3 * a + b] = T1[3 * a] * R[3 * b] + T1[3 * a + 1] * R[3 * b + 1] + T1[3 * a + 2] * R[3 * b + 2];
      matmul3(R, Bk, T1);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) Bt[3 * a + b] = T1[3 * a] * R[3 * b] + T1[3 * a + 1] * R[3 * b + 1] + T1[3 * a + 2] * R[3 * b + 2];
      matmul3(R, C, T1);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) Cn[3 * a + b] = T1[3 * a] * R[3 * b] + T1[3 * a + 1] * R[3 * b + 1] + T1[3 * a + 2] * R[3 * b + 2];
      skew3(p, P);
      double AP[9], Bn[9], PB[9], BtP[9];
      matmul3(An, P, AP);
#pragma unroll
      for (int k = 0; k < 9; ++k) Bn[k] = Bt[k] - AP[k];
      matmul3(P, Bn, PB);
      matTmul3(Bt, P, BtP);
#pragma unroll
      for (int k = 0; k < 9; ++k) { A[k] = An[k]; Bk[k] = Bn[k]; C[k] = Cn[k] + PB[k] - BtP[k]; }
      act_force(li, pa, pA);
    }
  }
  // root (free-flyer): Ia0 = Y_0 + children = [[A0, B0], [B0^T, C0]] in 3x3 blocks, D = Ia0.  The solve D a = u goes through
  // the Schur complement with cofactor inverses of the two SPD 3x3 blocks (A0: the mass block, S = C0 - B0^T A0^-1 B0): a
  // quarter of the dependent operations of a 6x6 factorisation, and no 36-entry factor to keep live (the rollout's chain
  // waited on its local-memory reloads, ncu round 2).
  double A0[9], B0[9], C0[9];
  {
    double h[6], vh[6];
    inertia_apply(M.mass[0], M.com[0], M.Ic[0], nd.v[0], h);
    cross_mf(nd.v[0], h, vh);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        A0[3 * a + b] = A[3 * a + b] + M.Y[0][6 * a + b];
        B0[3 * a + b] = Bk[3 * a + b] + M.Y[0][6 * a + 3 + b];
        C0[3 * a + b] = C[3 * a + b] + M.Y[0][6 * (3 + a) + 3 + b];
      }
#pragma unroll
    for (int k = 0; k < 6; ++k) uu[k] -= pA[k] + vh[k];
  }
  double Ai[9], Tm[9], Si[9];
  {
    inv3_sym(A0, Ai);
    matmul3(Ai, B0, Tm);  // T = A0^-1 B0
    double S[9];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = a; b < 3; ++b) S[3 * a + b] = C0[3 * a + b] - (B0[a] * Tm[b] + B0[3 + a] * Tm[3 + b] + B0[6 + a] * Tm[6 + b]);  // C0 - B0^T T, upper
    inv3_sym(S, Si);
  }
  // pass 3
  {
    double g[6];
    {  // gravity in the base frame: oM[0]^-1 acting on a0 = (-gravity, 0); oM[0].R is liR[0]
      SE3 o0;
#pragma unroll
      for (int k = 0; k < 9; ++k) o0.R[k] = nd.liR[0][k];
      o0.p[0] = o0.p[1] = o0.p[2] = 0.0;  // a0 has no angular part, so the translation does not enter
      actinv_motion(o0, M.a0, g);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) nd.agf[0][k] += g[k];
    double rhs[6];
    {  // y2 = S^-1 (r2 - T^T r1), y1 = A0^-1 r1 - T y2
      double w[3], y2[3], t1[3], t2[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) w[a] = uu[3 + a] - (Tm[a] * uu[0] + Tm[3 + a] * uu[1] + Tm[6 + a] * uu[2]);
      matvec3(Si, w, y2);
      matvec3(Ai, uu, t1);
      matvec3(Tm, y2, t2);
#pragma unroll
      for (int a = 0; a < 3; ++a) { rhs[a] = t1[a] - t2[a]; rhs[3 + a] = y2[a]; }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) { nd.a[k] = rhs[k] - nd.agf[0][k]; }
#pragma unroll
    for (int k = 0; k < 6; ++k) nd.agf[0][k] += nd.a[k];
  }
#pragma unroll UNR
  for (int i = 1; i < NJ; ++i) {
    SE3 li;
#pragma unroll
    for (int k = 0; k < 9; ++k) li.R[k] = nd.liR[i][k];
#pragma unroll
    for (int k = 0; k < 3; ++k) li.p[k] = M.jp[i][k];
    double ap[6]; actinv_motion(li, nd.agf[i - 1], ap);
#pragma unroll
    for (int k = 0; k < 6; ++k) nd.agf[i][k] += ap[k];
    const int c = 5 + i;
    nd.a[c] = Dinv[i] * uu[c] - dot6(UDinv[i], nd.agf[i]);
#pragma unroll
    for (int k = 0; k < 3; ++k) nd.agf[i][3 + k] += M.axis[i][k] * nd.a[c];
  }
}

template <class D, bool FULL = false>
EMPC_DI void aba(const DevModel& M, const double* x, const double* tau, NodeData<D>& nd) {
  aba_kinematics<D, FULL>(M, x, nd);
  aba_dynamics<D, FULL>(M, tau, nd);
}

// IntegratedActionModelEuler::calc.  u == nullptr => terminal convention u = 0.
template <class D, bool FULL = false>
EMPC_DI void node_calc(const DevModel& M, const CostTables& C, int costset, double smooth, const double* x,
                       const double* u, NodeData<D>& nd, double* xnext, double& cost) {
  squash<D>(M, smooth, u, nd.s);
  double tau[D::NV];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double t = 0;
#pragma unroll
    for (int j = 0; j < D::NR; ++j) t += M.tau_f[i * D::NR + j] * nd.s[j];
    tau[i] = t;
  }
#pragma unroll
  for (int i = 0; i < D::NA; ++i) tau[6 + i] = nd.s[D::NR + i];
  aba_kinematics<D, FULL>(M, x, nd);
  // costs depend on (x, u) and the kinematics only: evaluated before the dynamics sweep so that the world placements
  // need not stay live across it
  double csum = 0;
  const int c0 = C.costset_begin[costset], c1 = C.costset_begin[costset + 1];
  for (int c = c0; c < c1; ++c) {
    const empc_cost_t cs = C.costs[c];
    if (!cs.active || cs.type == EMPC_COST_CONTACT_FRICTION_CONE) continue;
    double r[D::NDX], Ar[D::NDX], Arr[D::NDX];
    SE3 rMf;
    csum += cs.weight * cost_eval<D>(M, C, cs, smooth, x, u, nd, r, Ar, Arr, rMf);
  }
  aba_dynamics<D, FULL>(M, tau, nd);
  const double dt = M.dt, dt2 = dt * dt;
#pragma unroll
  for (int i = 0; i < D::NV; ++i) {
    nd.dx[i] = x[D::NQ + i] * dt + nd.a[i] * dt2;
    nd.dx[D::NV + i] = nd.a[i] * dt;
  }
  state_integrate<D>(x, nd.dx, xnext);
  cost = dt * csum;
}

// Dynamics half of IntegratedActionModelEuler::calc (squash -> thrust map -> ABA -> semi-implicit Euler), without the
// costs: the rollout's sequential chain (rollout.cuh) evaluates only this; the trial costs are computed in parallel over
// all (trial, node) pairs afterwards (trial_cost_kernel).
template <class D, bool FULL = false>
EMPC_DI void node_dyn(const DevModel& M, double smooth, const double* x, const double* u, double* xnext) {
  NodeData<D> nd;
  squash<D>(M, smooth, u, nd.s);
  double tau[D::NV];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double t = 0;
#pragma unroll
    for (int j = 0; j < D::NR; ++j) t += M.tau_f[i * D::NR + j] * nd.s[j];
    tau[i] = t;
  }
#pragma unroll
  for (int i = 0; i < D::NA; ++i) tau[6 + i] = nd.s[D::NR + i];
  aba_kinematics<D, FULL>(M, x, nd);
  aba_dynamics<D, FULL>(M, tau, nd);
  const double dt = M.dt, dt2 = dt * dt;
#pragma unroll
  for (int i = 0; i < D::NV; ++i) {
    nd.dx[i] = x[D::NQ + i] * dt + nd.a[i] * dt2;
    nd.dx[D::NV + i] = nd.a[i] * dt;
  }
  state_integrate<D>(x, nd.dx, xnext);
}

// Value of one FRAME cost at state x (kinematics recomputed inside).  Deliberately not inlined: the world placements and
// joint velocities (NodeData, ~1.3 KB) then live in this function's frame only, and only on the rare nodes with frame costs.
// (OVERLAY: the contact / RK4 code paths get their own instance, so that the register allocation and frame of the one
//  the tuned kernels call do not depend on them)
template <class D, bool OVERLAY = false>
__device__ __noinline__ double frame_cost_value(const DevModel& M, const CostTables& C, const empc_cost_t& cs, double smooth, const double* x) {
  NodeData<D> nd;
  aba_kinematics<D, false>(M, x, nd);
  double r[D::NDX], Ar[D::NDX], Arr[D::NDX];
  SE3 rMf;
  return cost_eval<D>(M, C, cs, smooth, x, nullptr, nd, r, Ar, Arr, rMf);
}
EMPC_DI bool is_frame_cost(int type) { return type >= EMPC_COST_FRAME_PLACEMENT && type <= EMPC_COST_FRAME_TRANSLATION; }

// contact.cuh: contact force at (x, u) of a node under contact dynamics
template <class D>
__device__ __noinline__ void contact_force(const DevModel& M, const empc_contact_t* ctp, double smooth, const double* x, const double* u,
                                           double* lam);
// value (and activation derivatives) of the friction-cone cost: activation(A f) with the quadratic barrier the factory
// attaches (src/factory/cost.cpp:149-167); f = contact force in contact-frame coordinates
EMPC_DI double friction_cone_eval(const CostTables& C, const empc_cost_t& cs, const double* lam, double* r, double* Ar, double* Arr) {
  const double* A = C.pool + cs.ref_off;
#pragma unroll
  for (int i = 0; i < 5; ++i) r[i] = A[3 * i] * lam[0] + A[3 * i + 1] * lam[1] + A[3 * i + 2] * lam[2];
  return activation<5>(cs.activation, r, C.pool + cs.w_off, C.pool + cs.lb_off, C.pool + cs.ub_off, Ar, Arr);
}

// Cost half: dt * sum_c w_c a_c(r_c(x, u)) in the reference's cost order; needs the kinematics only when the cost set
// holds frame costs.
// rk4.cuh: cost of a node under IntegratedActionModelRK4 (the four stage costs)
template <class D>
__device__ __noinline__ double rk4_node_cost(const DevModel& M, const CostTables& C, int costset, double smooth, const double* x,
                                             const double* u);
// CONTACT: the overlay instantiation — the problem has contact stages, so a cost set may hold a friction-cone cost (its
// residual needs the contact solve at (x, u)), or it integrates with RK4 (the node cost is the weighted sum of four stage
// costs): calls the other instantiation does not carry.  raw: the cost sum without the integrator's dt.
template <class D, bool CONTACT = false>
EMPC_DI double node_cost_value(const DevModel& M, const CostTables& C, int costset, double smooth, const double* x, const double* u,
                               bool raw = false) {
  if (CONTACT && !raw && M.integrator == EMPC_INTEGRATOR_RK4) return rk4_node_cost<D>(M, C, costset, smooth, x, u);
  const int c0 = C.costset_begin[costset], c1 = C.costset_begin[costset + 1];
  double csum = 0;
  double gref[D::NX], gr[D::NDX];  // state costs that share a reference share the residual x (-) ref
  bool g_open = false;
  for (int c = c0; c < c1; ++c) {
    const empc_cost_t cs = C.costs[c];
    if (!cs.active) continue;
    if (is_frame_cost(cs.type)) { csum += cs.weight * frame_cost_value<D, CONTACT>(M, C, cs, smooth, x); continue; }
    double Ar[D::NDX], Arr[D::NDX];
    if (!CONTACT && cs.type == EMPC_COST_CONTACT_FRICTION_CONE) continue;  // (empc_create refuses it without a contact)
    if (CONTACT && cs.type == EMPC_COST_CONTACT_FRICTION_CONE) {  // needs the contact force: the contact solve at (x, u)
      const int ci = C.costset_contact ? C.costset_contact[costset] : -1;
      if (ci >= 0) {
        double lam[6], r5[5];
        contact_force<D>(M, C.contacts + ci, smooth, x, u, lam);
        csum += cs.weight * friction_cone_eval(C, cs, lam, r5, Ar, Arr);
      }
      continue;
    }
    if (cs.type == EMPC_COST_STATE) {
      const double* ref = C.pool + cs.ref_off;
      bool same = g_open;
      if (same) {
#pragma unroll
        for (int i = 0; i < D::NX; ++i) same &= (ref[i] == gref[i]);
      }
      if (!same) {
#pragma unroll
        for (int i = 0; i < D::NX; ++i) gref[i] = ref[i];
        state_diff<D>(gref, x, gr);
        g_open = true;
      }
      csum += cs.weight * activation<D::NDX>(cs.activation, gr, C.pool + cs.w_off, C.pool + cs.lb_off, C.pool + cs.ub_off, Ar, Arr);
    } else {
      double r[D::NDX];
      SE3 rMf;
      NodeData<D>* no_kinematics = nullptr;  // control residuals never touch the kinematics
      csum += cs.weight * cost_eval<D>(M, C, cs, smooth, x, u, *no_kinematics, r, Ar, Arr, rMf);
    }
  }
  return raw ? csum : M.dt * csum;
}

}  // namespace empc
