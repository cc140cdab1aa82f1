// oracle_model.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle; parity unpinned, see oracle_math.hpp header).
//
// Per-node action model of eagle-mpc's shooting problems, restated from the Crocoddyl 1.x / Pinocchio 2.x chain the
// reference instantiates (src/factory/diff-action.cpp:31-35, src/factory/int-action.cpp:26, src/trajectory.cpp:47-52):
//   IntegratedActionModelEuler            crocoddyl/core/integrator/euler.hxx           (calc / calcDiff)
//   DifferentialActionModelFreeFwdDynamics crocoddyl/multibody/actions/free-fwddyn.hxx
//   DifferentialActionModelContactFwdDynamics crocoddyl/multibody/actions/contact-fwddyn.hxx, ContactModel3D / 6D,
//     pinocchio::forwardDynamics / getKKTContactDynamicMatrixInverse, ResidualModelContactFrictionCone
//     (src/factory/diff-action.cpp:30-32, src/factory/contacts.cpp:32-81, src/factory/cost.cpp:149-167)
//   ActuationSquashingModel + SquashingModelSmoothSat + ActuationModelMultiCopterBase
//   CostModelSum / CostModelResidual / Activation* / Residual{State,Control,Frame*}
//   StateMultibody (integrate / diff / Jdiff / Jintegrate / JintegrateTransport)
//   pinocchio::aba, pinocchio::computeABADerivatives, frame kinematics
#pragma once
#include <vector>

#include "../include/empc_b200.h"
#include "oracle_math.hpp"

namespace orc {

constexpr int MAXJ = EMPC_MAX_JOINTS;
constexpr int MAXV = 6 + MAXJ - 1;
constexpr int MAXQ = MAXV + 1;
constexpr int MAXX = MAXQ + MAXV;
constexpr int MAXDX = 2 * MAXV;
constexpr int MAXU = EMPC_MAX_NU;

struct Model {
  const empc_problem_desc_t* d = nullptr;
  int nj = 0, na = 0, nq = 0, nv = 0, nx = 0, ndx = 0, nu = 0, nr = 0, tile = 0;
  int col0[MAXJ];          // first velocity column of joint i
  int ncol[MAXJ];          // 6 for the free-flyer, 1 for revolute
  int joint_of_col[MAXV];  // joint owning velocity column c
  bool anc[MAXJ][MAXJ];    // anc[j][k]: joint j is an ancestor of k, or k itself
  SE3 jplace[MAXJ];
  double Y[MAXJ][36];
  SE3 fplace[EMPC_MAX_FRAMES];
  double a0[6];            // -gravity as a spatial acceleration
  double dt = 0;
  // tile offsets
  int oFx, oFu, oLxx, oLxu, oLuu, oLx, oLu;

  void init(const empc_problem_desc_t* desc) {
    d = desc;
    const empc_robot_t& r = d->robot;
    nj = r.n_joints; na = nj - 1; nq = 7 + na; nv = 6 + na; nx = nq + nv; ndx = 2 * nv;
    nr = d->n_rotors; nu = nr + na; dt = d->dt;
    for (int i = 0; i < nj; ++i) {
      col0[i] = (i == 0) ? 0 : 5 + i;
      ncol[i] = (i == 0) ? 6 : 1;
      std::memcpy(jplace[i].R, r.jplace_R[i], sizeof(double) * 9);
      std::memcpy(jplace[i].p, r.jplace_p[i], sizeof(double) * 3);
      inertia_matrix(r.mass[i], r.com[i], r.inertia[i], Y[i]);
    }
    for (int c = 0; c < nv; ++c) joint_of_col[c] = (c < 6) ? 0 : c - 5;
    for (int j = 0; j < nj; ++j)
      for (int k = 0; k < nj; ++k) {
        bool a = false;
        for (int m = k; m >= 0; m = r.parent[m]) if (m == j) { a = true; break; }
        anc[j][k] = a;
      }
    for (int f = 0; f < r.n_frames; ++f) {
      std::memcpy(fplace[f].R, r.frame_R[f], sizeof(double) * 9);
      std::memcpy(fplace[f].p, r.frame_p[f], sizeof(double) * 3);
    }
    for (int i = 0; i < 3; ++i) { a0[i] = -r.gravity[i]; a0[3 + i] = 0; }
    oFx = 0; oFu = oFx + ndx * ndx; oLxx = oFu + ndx * nu; oLxu = oLxx + ndx * ndx; oLuu = oLxu + ndx * nu;
    oLx = oLuu + nu * nu; oLu = oLx + ndx;
    tile = oLu + nu; tile += tile & 1;
  }
};

// ---- StateMultibody (crocoddyl/multibody/states/multibody.hxx) -----------------------------------------------------
inline void state_zero(const Model& m, double* x) {
  for (int i = 0; i < m.nx; ++i) x[i] = 0;
  x[6] = 1;
}
inline void q_to_se3(const double* q, SE3& M) {
  quat_to_R(q + 3, M.R);
  M.p[0] = q[0]; M.p[1] = q[1]; M.p[2] = q[2];
}
// pinocchio::integrate on SE3 x R^na (special-euclidean.hpp integrate_impl) ; v += dv
inline void state_integrate(const Model& m, const double* x, const double* dx, double* out) {
  SE3 M0, E, M1;
  q_to_se3(x, M0);
  exp6(dx, E);
  se3_mul(M0, E, M1);
  double quat[4];
  R_to_quat(M1.R, quat);
  const double dotp = quat[0] * x[3] + quat[1] * x[4] + quat[2] * x[5] + quat[3] * x[6];
  if (dotp < 0) for (int i = 0; i < 4; ++i) quat[i] = -quat[i];
  const double n2 = quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3];
  const double alpha = (3 - n2) / 2;  // quaternion::firstOrderNormalize
  double o[MAXX];
  o[0] = M1.p[0]; o[1] = M1.p[1]; o[2] = M1.p[2];
  for (int i = 0; i < 4; ++i) o[3 + i] = quat[i] * alpha;
  for (int i = 0; i < m.na; ++i) o[7 + i] = x[7 + i] + dx[6 + i];
  for (int i = 0; i < m.nv; ++i) o[m.nq + i] = x[m.nq + i] + dx[m.nv + i];
  std::memcpy(out, o, sizeof(double) * m.nx);
}
// dx = x1 (-) x0
inline void state_diff(const Model& m, const double* x0, const double* x1, double* dx) {
  SE3 M0, M1, D;
  q_to_se3(x0, M0); q_to_se3(x1, M1);
  se3_inv_mul(M0, M1, D);
  log6(D, dx);
  for (int i = 0; i < m.na; ++i) dx[6 + i] = x1[7 + i] - x0[7 + i];
  for (int i = 0; i < m.nv; ++i) dx[m.nv + i] = x1[m.nq + i] - x0[m.nq + i];
}

// ---- per-node scratch (the crocoddyl "data" objects) ------------------------------------------------------------------
struct Work {
  double x[MAXX], u[MAXU];
  double s[MAXU], ds[MAXU];   // squashed control and ds/du
  double tau[MAXV];
  SE3 liMi[MAXJ], oMi[MAXJ];
  double v[MAXJ][6];          // local joint spatial velocities
  double agf[MAXJ][6];        // local spatial accelerations incl. gravity field (data.a_gf)
  double a[MAXV];             // joint accelerations
  double dx[MAXDX];
  double xnext[MAXX];
  double cost;
  // world-frame quantities of calcDiff
  double J[6][MAXV];
  double ov[MAXJ][6], oa[MAXJ][6];
  double Minv[MAXV * MAXV];
  // contact dynamics (nc = 0: free dynamics)
  int nc = 0, cframe = -1;
  SE3 oMf;                    // world placement of the contact frame
  double fJ[6][MAXV];         // its LOCAL Jacobian; the constraint rows Jc are rows 0 .. nc-1
  double vf[6];               // its LOCAL velocity
  double lam[6];              // contact force, contact-frame coordinates (pinocchio lambda_c)
  double Bc[MAXV][6];         // Minv Jc^T
  double Ginv[36];            // (Jc Minv Jc^T)^-1, nc x nc
};

struct SolverCtx {  // the pieces of solver state the node model depends on
  double smooth;         // SquashingModelSmoothSat::smooth_
  double barrier_weight; // weight of the "barrier" cost (1e-3)
};

// pinocchio::aba (algorithm/aba.hxx), local-frame three-pass recursion, SURVEY.md B.8
inline void aba(const Model& m, const double* q, const double* vq, const double* tau, Work& w) {
  const empc_robot_t& r = m.d->robot;
  double Ia[MAXJ][36], pA[MAXJ][6], uu[MAXV];
  double U[MAXJ][6], Dinv[MAXJ], UDinv[MAXJ][6];
  for (int i = 0; i < m.nv; ++i) uu[i] = tau[i];
  // pass 1
  for (int i = 0; i < m.nj; ++i) {
    if (i == 0) {
      q_to_se3(q, w.liMi[0]);
      w.oMi[0] = w.liMi[0];
      for (int k = 0; k < 6; ++k) w.v[0][k] = vq[k];
      for (int k = 0; k < 6; ++k) w.agf[0][k] = 0;  // c = v x vJ = 0 for the root
    } else {
      const double th = q[6 + i];  // q[7 + (i-1)]
      double ax[3] = {r.axis[i][0] * th, r.axis[i][1] * th, r.axis[i][2] * th};
      SE3 Mj; exp3(ax, Mj.R); Mj.p[0] = Mj.p[1] = Mj.p[2] = 0;
      se3_mul(m.jplace[i], Mj, w.liMi[i]);
      se3_mul(w.oMi[r.parent[i]], w.liMi[i], w.oMi[i]);
      double vJ[6] = {0, 0, 0, r.axis[i][0] * vq[5 + i], r.axis[i][1] * vq[5 + i], r.axis[i][2] * vq[5 + i]};
      double vp[6]; actinv_motion(w.liMi[i], w.v[r.parent[i]], vp);
      for (int k = 0; k < 6; ++k) w.v[i][k] = vJ[k] + vp[k];
      cross_mm(w.v[i], vJ, w.agf[i]);
    }
    std::memcpy(Ia[i], m.Y[i], sizeof(double) * 36);
    double h[6]; mat6_vec(m.Y[i], w.v[i], h);
    cross_mf(w.v[i], h, pA[i]);
  }
  // pass 2
  double LL[36];  // Cholesky factor of the root articulated inertia
  for (int i = m.nj - 1; i >= 0; --i) {
    if (i == 0) {
      for (int k = 0; k < 6; ++k) uu[k] -= pA[0][k];
      std::memcpy(LL, Ia[0], sizeof(LL));
      llt_inplace(LL, 6);
    } else {
      const double S[6] = {0, 0, 0, r.axis[i][0], r.axis[i][1], r.axis[i][2]};
      const int c = m.col0[i];
      uu[c] -= dot6(S, pA[i]);
      mat6_vec(Ia[i], S, U[i]);
      Dinv[i] = 1.0 / dot6(S, U[i]);
      for (int k = 0; k < 6; ++k) UDinv[i][k] = U[i][k] * Dinv[i];
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) Ia[i][6 * a + b] -= UDinv[i][a] * U[i][b];
      double pa[6], Iac[6];
      mat6_vec(Ia[i], w.agf[i], Iac);
      for (int k = 0; k < 6; ++k) pa[k] = pA[i][k] + Iac[k] + UDinv[i][k] * uu[c];
      double X[36], Ip[36], fp[6];
      force_action_matrix(w.liMi[i], X);
      congruence6(X, Ia[i], Ip);
      const int p = r.parent[i];
      for (int k = 0; k < 36; ++k) Ia[p][k] += Ip[k];
      act_force(w.liMi[i], pa, fp);
      for (int k = 0; k < 6; ++k) pA[p][k] += fp[k];
    }
  }
  // pass 3
  for (int i = 0; i < m.nj; ++i) {
    if (i == 0) {
      double g[6]; actinv_motion(w.oMi[0], m.a0, g);
      for (int k = 0; k < 6; ++k) w.agf[0][k] += g[k];
      double rhs[6];
      for (int k = 0; k < 6; ++k) rhs[k] = uu[k];
      llt_solve(LL, 6, rhs, 1);  // Dinv * u
      for (int k = 0; k < 6; ++k) { w.a[k] = rhs[k] - w.agf[0][k]; }  // UDinv = I for the free-flyer
      for (int k = 0; k < 6; ++k) w.agf[0][k] += w.a[k];
    } else {
      double ap[6]; actinv_motion(w.liMi[i], w.agf[r.parent[i]], ap);
      for (int k = 0; k < 6; ++k) w.agf[i][k] += ap[k];
      const int c = m.col0[i];
      w.a[c] = Dinv[i] * uu[c] - dot6(UDinv[i], w.agf[i]);
      for (int k = 0; k < 3; ++k) w.agf[i][3 + k] += r.axis[i][k] * w.a[c];
    }
  }
}

// Recursive Newton-Euler (local frame), used only by tests as an independent check of aba / derivatives.
inline void rnea(const Model& m, const double* q, const double* vq, const double* aq, double* tau) {
  const empc_robot_t& r = m.d->robot;
  SE3 liMi[MAXJ], oMi[MAXJ];
  double v[MAXJ][6], a[MAXJ][6], f[MAXJ][6];
  for (int i = 0; i < m.nj; ++i) {
    if (i == 0) {
      q_to_se3(q, liMi[0]); oMi[0] = liMi[0];
      for (int k = 0; k < 6; ++k) v[0][k] = vq[k];
      double g[6]; actinv_motion(oMi[0], m.a0, g);
      for (int k = 0; k < 6; ++k) a[0][k] = g[k] + aq[k];
    } else {
      const double th = q[6 + i];
      double ax[3] = {r.axis[i][0] * th, r.axis[i][1] * th, r.axis[i][2] * th};
      SE3 Mj; exp3(ax, Mj.R); Mj.p[0] = Mj.p[1] = Mj.p[2] = 0;
      se3_mul(m.jplace[i], Mj, liMi[i]);
      se3_mul(oMi[r.parent[i]], liMi[i], oMi[i]);
      double vJ[6] = {0, 0, 0, r.axis[i][0] * vq[5 + i], r.axis[i][1] * vq[5 + i], r.axis[i][2] * vq[5 + i]};
      double vp[6], ap[6], c[6];
      actinv_motion(liMi[i], v[r.parent[i]], vp);
      for (int k = 0; k < 6; ++k) v[i][k] = vJ[k] + vp[k];
      cross_mm(v[i], vJ, c);
      actinv_motion(liMi[i], a[r.parent[i]], ap);
      for (int k = 0; k < 6; ++k) a[i][k] = ap[k] + c[k];
      for (int k = 0; k < 3; ++k) a[i][3 + k] += r.axis[i][k] * aq[5 + i];
    }
    double h[6], Ya[6], vh[6];
    mat6_vec(m.Y[i], v[i], h); mat6_vec(m.Y[i], a[i], Ya); cross_mf(v[i], h, vh);
    for (int k = 0; k < 6; ++k) f[i][k] = Ya[k] + vh[k];
  }
  for (int i = m.nj - 1; i >= 0; --i) {
    if (i == 0) {
      for (int k = 0; k < 6; ++k) tau[k] = f[0][k];
    } else {
      tau[5 + i] = r.axis[i][0] * f[i][3] + r.axis[i][1] * f[i][4] + r.axis[i][2] * f[i][5];
      double fp[6]; act_force(liMi[i], f[i], fp);
      for (int k = 0; k < 6; ++k) f[r.parent[i]][k] += fp[k];
    }
  }
}

// ---- activations (crocoddyl/core/activations/*.hpp) ------------------------------------------------------------------
// returns a_value; fills Ar, Arr (diagonal).  w/lb/ub may be null depending on the type.
inline double activation(int type, int n, const double* r, const double* w, const double* lb, const double* ub,
                         double* Ar, double* Arr) {
  double val = 0;
  switch (type) {
    case EMPC_ACT_QUAD:
      for (int i = 0; i < n; ++i) { val += r[i] * r[i]; Ar[i] = r[i]; Arr[i] = 1; }
      return 0.5 * val;
    case EMPC_ACT_WEIGHTED_QUAD:
      for (int i = 0; i < n; ++i) { const double wr = w[i] * r[i]; val += r[i] * wr; Ar[i] = wr; Arr[i] = w[i]; }
      return 0.5 * val;
    case EMPC_ACT_QUAD_BARRIER: {
      double sl = 0, su = 0;
      for (int i = 0; i < n; ++i) {
        const double dl = r[i] - lb[i], du = r[i] - ub[i];
        const double l = dl < 0 ? dl : 0.0, uu = du > 0 ? du : 0.0;
        sl += l * l; su += uu * uu;
        Ar[i] = l + uu;
        Arr[i] = (dl <= 0) ? 1.0 : ((du >= 0) ? 1.0 : 0.0);
      }
      return 0.5 * sl + 0.5 * su;
    }
    case EMPC_ACT_WEIGHTED_QUAD_BARRIER: {
      // weighted-quadratic-barrier.hpp: value and gradient use (w r)^2, the Hessian uses w (as upstream)
      double sl = 0, su = 0;
      for (int i = 0; i < n; ++i) {
        const double dl = r[i] - lb[i], du = r[i] - ub[i];
        const double l = (dl < 0 ? dl : 0.0) * w[i], uu = (du > 0 ? du : 0.0) * w[i];
        sl += l * l; su += uu * uu;
        Ar[i] = (l + uu) * w[i];
        Arr[i] = (dl <= 0) ? w[i] : ((du >= 0) ? w[i] : 0.0);
      }
      return 0.5 * sl + 0.5 * su;
    }
  }
  return 0;
}

struct CostEval {  // residual + activation of one cost, kept between calc and calcDiff
  double r[MAXDX], Ar[MAXDX], Arr[MAXDX];
  SE3 rMf;  // frame placement / rotation error
};

inline int residual_dim(const Model& m, int type) {
  switch (type) {
    case EMPC_COST_STATE: return m.ndx;
    case EMPC_COST_CONTROL: return m.nu;
    case EMPC_COST_SQUASH_BARRIER: return m.nu;
    case EMPC_COST_FRAME_PLACEMENT: return 6;
    case EMPC_COST_FRAME_VELOCITY: return 6;
    case EMPC_COST_CONTACT_FRICTION_CONE: return 5;
    default: return 3;
  }
}

inline void frame_placement(const Model& m, const Work& w, int f, SE3& oMf) {
  se3_mul(w.oMi[m.d->robot.frame_joint[f]], m.fplace[f], oMf);
}

// residual value + activation; returns the activation value
inline double cost_calc(const Model& m, const SolverCtx& ctx, const empc_cost_t& c, const Work& w, CostEval& e) {
  const double* pool = m.d->pool;
  const double* ref = c.ref_off >= 0 ? pool + c.ref_off : nullptr;
  const double* aw = c.w_off >= 0 ? pool + c.w_off : nullptr;
  const double* lb = c.lb_off >= 0 ? pool + c.lb_off : nullptr;
  const double* ub = c.ub_off >= 0 ? pool + c.ub_off : nullptr;
  const int n = residual_dim(m, c.type);
  switch (c.type) {
    case EMPC_COST_STATE: state_diff(m, ref, w.x, e.r); break;
    case EMPC_COST_CONTROL: for (int i = 0; i < n; ++i) e.r[i] = w.u[i] - ref[i]; break;
    case EMPC_COST_SQUASH_BARRIER: {
      // src/sbfddp.cpp:22-24,169-190,464-477: WeightedQuadraticBarrier(bounds(s_lb,s_ub,beta=1), 1/(smooth (ub-lb))^2)
      double bw[MAXU], blb[MAXU], bub[MAXU];
      for (int i = 0; i < n; ++i) {
        const double aux = ctx.smooth * (m.d->u_ub[i] - m.d->u_lb[i]);
        bw[i] = 1.0 / (aux * aux);
        const double mid = 0.5 * (m.d->u_lb[i] + m.d->u_ub[i]), dd = 0.5 * (m.d->u_ub[i] - m.d->u_lb[i]);
        blb[i] = mid - 1.0 * dd; bub[i] = mid + 1.0 * dd;  // crocoddyl::ActivationBounds ctor with beta = 1
        e.r[i] = w.u[i];
      }
      return activation(EMPC_ACT_WEIGHTED_QUAD_BARRIER, n, e.r, bw, blb, bub, e.Ar, e.Arr);
    }
    case EMPC_COST_FRAME_PLACEMENT: {
      SE3 Mref, oMf; std::memcpy(Mref.R, ref, 72); std::memcpy(Mref.p, ref + 9, 24);
      frame_placement(m, w, c.frame, oMf);
      se3_inv_mul(Mref, oMf, e.rMf);
      log6(e.rMf, e.r);
    } break;
    case EMPC_COST_FRAME_ROTATION: {
      SE3 oMf; frame_placement(m, w, c.frame, oMf);
      matTmul3(ref, oMf.R, e.rMf.R);
      double th; log3(e.rMf.R, e.r, th);
    } break;
    case EMPC_COST_FRAME_TRANSLATION: {
      SE3 oMf; frame_placement(m, w, c.frame, oMf);
      for (int i = 0; i < 3; ++i) e.r[i] = oMf.p[i] - ref[i];
    } break;
    case EMPC_COST_FRAME_VELOCITY: {
      double vf[6]; actinv_motion(m.fplace[c.frame], w.v[m.d->robot.frame_joint[c.frame]], vf);
      for (int i = 0; i < 6; ++i) e.r[i] = vf[i] - ref[i];
    } break;
    case EMPC_COST_CONTACT_FRICTION_CONE:  // ResidualModelContactFrictionCone: r = A f, f in the contact frame
      for (int i = 0; i < 5; ++i) e.r[i] = (w.nc > 0) ? ref[3 * i] * w.lam[0] + ref[3 * i + 1] * w.lam[1] + ref[3 * i + 2] * w.lam[2] : 0.0;
      break;
  }
  return activation(c.activation, n, e.r, aw, lb, ub, e.Ar, e.Arr);
}

inline void contact_calc(const Model& m, const empc_contact_t& ct, const double* vq, Work& w);

// Differential action model at (x, u) — DifferentialActionModelFreeFwdDynamics / ContactFwdDynamics::calc: squashing,
// thrust map, forward dynamics, CostModelSum.  Fills w (kinematics, accelerations w.a) and returns the cost (not yet
// weighted by the integrator).
inline double diff_calc(const Model& m, const SolverCtx& ctx, int costset, const double* x, const double* u, Work& w,
                        std::vector<CostEval>* evals = nullptr) {
  const empc_problem_desc_t& d = *m.d;
  std::memcpy(w.x, x, sizeof(double) * m.nx);
  for (int i = 0; i < m.nu; ++i) w.u[i] = u ? u[i] : 0.0;  // calc(data,x) == calc(data,x,unone_=0), SURVEY B.7
  // squashing (crocoddyl/core/actuation/squashing/smooth-sat.hpp)
  for (int i = 0; i < m.nu; ++i) {
    if (d.use_squash) {
      const double dd = (d.u_ub[i] - d.u_lb[i]) * ctx.smooth, a = dd * dd;
      const double l = w.u[i] - d.u_lb[i], h = w.u[i] - d.u_ub[i];
      w.s[i] = 0.5 * (std::sqrt(l * l + a) - std::sqrt(h * h + a) + d.u_lb[i] + d.u_ub[i]);
    } else {
      w.s[i] = w.u[i];
    }
  }
  // ActuationModelMultiCopterBase: tau = [tau_f s_rotors ; s_arm]
  for (int i = 0; i < 6; ++i) {
    double t = 0;
    for (int j = 0; j < m.nr; ++j) t += d.tau_f[i * m.nr + j] * w.s[j];
    w.tau[i] = t;
  }
  for (int i = 0; i < m.na; ++i) w.tau[6 + i] = w.s[m.nr + i];
  aba(m, x, x + m.nq, w.tau, w);
  // a model with a contact: DifferentialActionModelContactFwdDynamics (src/factory/diff-action.cpp:30-32); a model of a
  // contact trajectory whose ContactModelMultiple is empty reduces to the free dynamics
  w.nc = 0;
  if (d.n_contacts > 0 && d.costset_contact && d.costset_contact[costset] >= 0)
    contact_calc(m, d.contacts[d.costset_contact[costset]], x + m.nq, w);
  // CostModelSum::calc, costs in name order
  double cost = 0;
  const int c0 = d.costset_begin[costset], c1 = d.costset_begin[costset + 1];
  if (evals) evals->resize(c1 - c0);
  CostEval tmp;
  for (int c = c0; c < c1; ++c) {
    if (!d.costs[c].active) continue;
    CostEval& e = evals ? (*evals)[c - c0] : tmp;
    cost += d.costs[c].weight * cost_calc(m, ctx, d.costs[c], w, e);
  }
  return cost;
}

// IntegratedActionModelRK4 (crocoddyl/core/integrator/rk4.hxx): stage coefficients and weights
static const double kRk4C[4] = {0.0, 0.5, 0.5, 1.0}, kRk4W[4] = {1.0, 2.0, 2.0, 1.0};

// IntegratedActionModelEuler::calc / IntegratedActionModelRK4::calc — fills w (xnext, cost, and for Euler everything
// calcDiff reuses; w keeps the first stage's data, which is where the solver reads the squashed control, src/sbfddp.cpp:145)
inline void node_calc(const Model& m, const SolverCtx& ctx, int costset, const double* x, const double* u, Work& w,
                      std::vector<CostEval>* evals = nullptr) {
  const double dt = m.dt, dt2 = dt * dt;
  const double cost0 = diff_calc(m, ctx, costset, x, u, w, evals);
  if (m.d->integrator != EMPC_INTEGRATOR_RK4) {
    // semi-implicit Euler (euler.hxx calc)
    for (int i = 0; i < m.nv; ++i) {
      w.dx[i] = x[m.nq + i] * dt + w.a[i] * dt2;
      w.dx[m.nv + i] = w.a[i] * dt;
    }
    state_integrate(m, x, w.dx, w.xnext);
    w.cost = dt * cost0;
    return;
  }
  // y_i = x (+) c_i dt k_{i-1}, k_i = [v(y_i); a(y_i, u)]; dx = dt/6 sum w_i k_i; cost = dt/6 sum w_i l(y_i, u)
  double k[MAXDX], ksum[MAXDX], y[MAXX], dxi[MAXDX];
  for (int i = 0; i < m.nv; ++i) { k[i] = x[m.nq + i]; k[m.nv + i] = w.a[i]; }
  for (int i = 0; i < m.ndx; ++i) ksum[i] = k[i];
  double csum = cost0;
  static thread_local Work ws;  // (a stage's data is not kept: calcDiff re-evaluates the stages)
  for (int st = 1; st < 4; ++st) {
    for (int i = 0; i < m.ndx; ++i) dxi[i] = kRk4C[st] * dt * k[i];
    state_integrate(m, x, dxi, y);
    const double c = diff_calc(m, ctx, costset, y, u, ws, nullptr);
    for (int i = 0; i < m.nv; ++i) { k[i] = y[m.nq + i]; k[m.nv + i] = ws.a[i]; }
    for (int i = 0; i < m.ndx; ++i) ksum[i] += kRk4W[st] * k[i];
    csum += kRk4W[st] * c;
  }
  for (int i = 0; i < m.ndx; ++i) w.dx[i] = ksum[i] * (dt / 6.0);
  state_integrate(m, x, w.dx, w.xnext);
  w.cost = csum * (dt / 6.0);
}

// world Jacobian columns (motion axes of the velocity columns) and world body velocities
inline void world_kinematics(const Model& m, Work& w) {
  const empc_robot_t& r = m.d->robot;
  for (int i = 0; i < m.nj; ++i) {
    if (i == 0) {
      double X[36]; motion_action_matrix(w.oMi[0], X);
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) w.J[a][b] = X[6 * a + b];
    } else {
      const double S[6] = {0, 0, 0, r.axis[i][0], r.axis[i][1], r.axis[i][2]};
      double Jc[6]; act_motion(w.oMi[i], S, Jc);
      for (int a = 0; a < 6; ++a) w.J[a][m.col0[i]] = Jc[a];
    }
    act_motion(w.oMi[i], w.v[i], w.ov[i]);
  }
}

// world spatial accelerations of the bodies for joint velocities vq and accelerations aq, on top of the base acceleration
// a_base (m.a0 = the gravity field of RNEA, or zero for the kinematic acceleration): A_i = A_p + S_i aq_i + V_p x S_i vq_i
inline void world_accelerations(const Model& m, const Work& w, const double* vq, const double* aq, const double* a_base,
                                double oa[][6]) {
  const empc_robot_t& r = m.d->robot;
  for (int i = 0; i < m.nj; ++i) {
    if (i == 0) {
      for (int k = 0; k < 6; ++k) {
        double s = a_base[k];
        for (int c = 0; c < 6; ++c) s += w.J[k][c] * aq[c];
        oa[0][k] = s;  // V_0 x (S vq) = V_0 x V_0 = 0
      }
    } else {
      const int c = m.col0[i], p = r.parent[i];
      double col[6], cr[6];
      for (int k = 0; k < 6; ++k) col[k] = w.J[k][c];
      cross_mm(w.ov[p], col, cr);
      for (int k = 0; k < 6; ++k) oa[i][k] = oa[p][k] + col[k] * aq[c] + cr[k] * vq[c];
    }
  }
}

// RNEA partial derivatives d tau/dq, d tau/dv (nv x nv) and the joint-space inertia M at the accelerations w.oa (world,
// gravity field included), with an optional external force Fext_w (world spatial force, constant in the local frame of
// joint jext, acting ON the robot: tau = RNEA - J^T F): pinocchio::computeRNEADerivatives(q, v, a, fext).
// w.J, w.ov, w.oa must be set.
inline void rnea_partials(const Model& m, Work& w, const double* Fext_w, int jext, double* M, double* dq, double* dv) {
  const empc_robot_t& r = m.d->robot;
  const int nv = m.nv;
  double oY[MAXJ][36], Bm[MAXJ][36], F[MAXJ][6];
  for (int i = 0; i < m.nj; ++i) {
    double X[36]; force_action_matrix(w.oMi[i], X);
    congruence6(X, m.Y[i], oY[i]);
    double h[6], Ya[6], vh[6];
    mat6_vec(oY[i], w.ov[i], h); mat6_vec(oY[i], w.oa[i], Ya); cross_mf(w.ov[i], h, vh);
    for (int k = 0; k < 6; ++k) F[i][k] = Ya[k] + vh[k];
    if (Fext_w && i == jext) for (int k = 0; k < 6; ++k) F[i][k] -= Fext_w[k];
    // B_i = crf(v) Y - Y crm(v) + Hx(h)   (DESIGN.md "RNEA derivatives")
    double Sv[9], Sw[9]; skew3(w.ov[i], Sv); skew3(w.ov[i] + 3, Sw);
    double crf[36], crm[36];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        crf[6 * a + b] = Sw[3 * a + b]; crf[6 * a + 3 + b] = 0;
        crf[6 * (3 + a) + b] = Sv[3 * a + b]; crf[6 * (3 + a) + 3 + b] = Sw[3 * a + b];
        crm[6 * a + b] = Sw[3 * a + b]; crm[6 * a + 3 + b] = Sv[3 * a + b];
        crm[6 * (3 + a) + b] = 0; crm[6 * (3 + a) + 3 + b] = Sw[3 * a + b];
      }
    double Shf[9], Shn[9]; skew3(h, Shf); skew3(h + 3, Shn);
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) {
        double s1 = 0, s2 = 0;
        for (int k = 0; k < 6; ++k) { s1 += crf[6 * a + k] * oY[i][6 * k + b]; s2 += oY[i][6 * a + k] * crm[6 * k + b]; }
        Bm[i][6 * a + b] = s1 - s2;
      }
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        Bm[i][6 * a + 3 + b] -= Shf[3 * a + b];
        Bm[i][6 * (3 + a) + b] -= Shf[3 * a + b];
        Bm[i][6 * (3 + a) + 3 + b] -= Shn[3 * a + b];
      }
  }
  // backward accumulation: composite inertia, composite B, subtree force
  for (int i = m.nj - 1; i > 0; --i) {
    const int p = r.parent[i];
    for (int k = 0; k < 36; ++k) { oY[p][k] += oY[i][k]; Bm[p][k] += Bm[i][k]; }
    for (int k = 0; k < 6; ++k) F[p][k] += F[i][k];
  }
  // YJ_c = Ycrb_j J_c ; BtJ_c = Bcrb_j^T J_c
  double YJ[MAXV][6], BtJ[MAXV][6], Jc[MAXV][6];
  for (int c = 0; c < nv; ++c) {
    const int j = m.joint_of_col[c];
    for (int a = 0; a < 6; ++a) Jc[c][a] = w.J[a][c];
    mat6_vec(oY[j], Jc[c], YJ[c]);
    mat6T_vec(Bm[j], Jc[c], BtJ[c]);
  }
  // joint-space inertia
  for (int cj = 0; cj < nv; ++cj)
    for (int ck = 0; ck < nv; ++ck) {
      const int j = m.joint_of_col[cj], k = m.joint_of_col[ck];
      double val = 0;
      if (m.anc[j][k]) val = dot6(Jc[cj], YJ[ck]);
      else if (m.anc[k][j]) val = dot6(Jc[ck], YJ[cj]);
      M[cj * nv + ck] = val;
    }
  // RNEA partial derivatives
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  for (int ck = 0; ck < nv; ++ck) {
    const int k = m.joint_of_col[ck], pk = r.parent[k];
    const double* s = Jc[ck];
    const double* vp = pk >= 0 ? w.ov[pk] : zero6;
    const double* ap = pk >= 0 ? w.oa[pk] : m.a0;
    double dVdq[6], dAdq[6], dAdv[6], t6[6], vsum[6];
    cross_mm(vp, s, dVdq);
    cross_mm(ap, s, dAdq); cross_mm(vp, dVdq, t6);
    for (int a = 0; a < 6; ++a) { dAdq[a] += t6[a]; vsum[a] = vp[a] + w.ov[k][a]; }
    cross_mm(vsum, s, dAdv);
    double P[6], Fq[6], Fv[6], t1[6], t2[6];
    mat6_vec(oY[k], dAdq, t1); mat6_vec(Bm[k], dVdq, t2);
    for (int a = 0; a < 6; ++a) P[a] = t1[a] + t2[a];
    cross_mf(s, F[k], t1);
    for (int a = 0; a < 6; ++a) Fq[a] = P[a] + t1[a];
    mat6_vec(oY[k], dAdv, t1); mat6_vec(Bm[k], s, t2);
    for (int a = 0; a < 6; ++a) Fv[a] = t1[a] + t2[a];
    for (int cj = 0; cj < nv; ++cj) {
      const int j = m.joint_of_col[cj];
      double vq_ = 0, vv_ = 0;
      if (j == k) { vq_ = dot6(Jc[cj], P); vv_ = dot6(Jc[cj], Fv); }
      else if (m.anc[j][k]) { vq_ = dot6(Jc[cj], Fq); vv_ = dot6(Jc[cj], Fv); }
      else if (m.anc[k][j]) {
        vq_ = dot6(YJ[cj], dAdq) + dot6(BtJ[cj], dVdq);
        vv_ = dot6(YJ[cj], dAdv) + dot6(BtJ[cj], s);
      }
      dq[cj * nv + ck] = vq_; dv[cj * nv + ck] = vv_;
    }
  }
}

inline void invert_spd(const double* M, int n, double* Minv) {
  double L[MAXV * MAXV];
  std::memcpy(L, M, sizeof(double) * n * n);
  llt_inplace(L, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Minv[i * n + j] = (i == j) ? 1.0 : 0.0;
  llt_solve(L, n, Minv, n);
}

// world Jacobian columns, velocities/accelerations, M^-1, d(ddq)/d(q,v,tau).  aq (nv x nv), av, Minv in w.Minv.
inline void aba_derivatives(const Model& m, Work& w, double* a_q, double* a_v) {
  const int nv = m.nv;
  world_kinematics(m, w);
  for (int i = 0; i < m.nj; ++i) act_motion(w.oMi[i], w.agf[i], w.oa[i]);
  double M[MAXV * MAXV], dq[MAXV * MAXV], dv[MAXV * MAXV];
  rnea_partials(m, w, nullptr, -1, M, dq, dv);
  // Minv by Cholesky; ddq_dq = -Minv dtau_dq etc. (computeABADerivatives)
  invert_spd(M, nv, w.Minv);
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < nv; ++j) {
      double sq = 0, sv = 0;
      for (int k = 0; k < nv; ++k) { sq += w.Minv[i * nv + k] * dq[k * nv + j]; sv += w.Minv[i * nv + k] * dv[k * nv + j]; }
      a_q[i * nv + j] = -sq; a_v[i * nv + j] = -sv;
    }
}

// local frame Jacobian (6 x nv) of frame f: fJ = Ad(oMf^-1) J restricted to supporting columns
inline void frame_jacobian(const Model& m, const Work& w, int f, const SE3& oMf, double fJ[6][MAXV]) {
  const int j = m.d->robot.frame_joint[f];
  for (int c = 0; c < m.nv; ++c) {
    if (m.anc[m.joint_of_col[c]][j]) {
      double col[6] = {w.J[0][c], w.J[1][c], w.J[2][c], w.J[3][c], w.J[4][c], w.J[5][c]}, o[6];
      actinv_motion(oMf, col, o);
      for (int a = 0; a < 6; ++a) fJ[a][c] = o[a];
    } else {
      for (int a = 0; a < 6; ++a) fJ[a][c] = 0;
    }
  }
}

// ---- DifferentialActionModelContactFwdDynamics (one ContactModel3D / 6D, zero Baumgarte gains) ------------------------
// calc: pinocchio::forwardDynamics(q, v, tau, Jc, a0, 0):  [M Jc^T; Jc 0] [a; -lambda] = [tau - h; -a0], with
// Jc = the constrained rows of the LOCAL frame Jacobian and a0 = the constrained frame acceleration at zero joint
// acceleration (3D: classical acceleration a.linear + w x v of the frame origin; 6D: the spatial acceleration).
// Follows aba() on the same state: w.a holds the free acceleration on entry, the constrained one on exit.
inline void contact_calc(const Model& m, const empc_contact_t& ct, const double* vq, Work& w) {
  const int nv = m.nv, nc = (ct.type == EMPC_CONTACT_6D) ? 6 : 3;
  const int jf = m.d->robot.frame_joint[ct.frame];
  w.nc = nc; w.cframe = ct.frame;
  world_kinematics(m, w);
  // joint-space inertia and its inverse (world-frame composite rigid bodies)
  {
    double zero[MAXV] = {0};
    world_accelerations(m, w, vq, zero, m.a0, w.oa);
    double M[MAXV * MAXV], dq[MAXV * MAXV], dv[MAXV * MAXV];
    rnea_partials(m, w, nullptr, -1, M, dq, dv);
    invert_spd(M, nv, w.Minv);
  }
  frame_placement(m, w, ct.frame, w.oMf);
  frame_jacobian(m, w, ct.frame, w.oMf, w.fJ);
  // drift a0
  double a0[6];
  {
    double zero[MAXV] = {0}, z6[6] = {0, 0, 0, 0, 0, 0}, oa[MAXJ][6], af[6];
    world_accelerations(m, w, vq, zero, z6, oa);
    actinv_motion(w.oMf, w.ov[jf], w.vf);
    actinv_motion(w.oMf, oa[jf], af);
    for (int k = 0; k < 6; ++k) a0[k] = af[k];
    if (nc == 3) { double cr[3]; cross3(w.vf + 3, w.vf, cr); for (int k = 0; k < 3; ++k) a0[k] += cr[k]; }
  }
  // B = Minv Jc^T, G = Jc B
  double G[36];
  for (int i = 0; i < nv; ++i)
    for (int r = 0; r < nc; ++r) {
      double s = 0;
      for (int k = 0; k < nv; ++k) s += w.Minv[i * nv + k] * w.fJ[r][k];
      w.Bc[i][r] = s;
    }
  for (int r = 0; r < nc; ++r)
    for (int c = 0; c < nc; ++c) {
      double s = 0;
      for (int i = 0; i < nv; ++i) s += w.fJ[r][i] * w.Bc[i][c];
      G[r * nc + c] = s;
    }
  invert_spd(G, nc, w.Ginv);
  // lambda = -G^-1 (Jc a_free + a0),  a = a_free + B lambda
  double rhs[6];
  for (int r = 0; r < nc; ++r) {
    double s = a0[r];
    for (int i = 0; i < nv; ++i) s += w.fJ[r][i] * w.a[i];
    rhs[r] = s;
  }
  for (int r = 0; r < 6; ++r) w.lam[r] = 0;
  for (int r = 0; r < nc; ++r) {
    double s = 0;
    for (int c = 0; c < nc; ++c) s += w.Ginv[r * nc + c] * rhs[c];
    w.lam[r] = -s;
  }
  for (int i = 0; i < nv; ++i) {
    double s = 0;
    for (int r = 0; r < nc; ++r) s += w.Bc[i][r] * w.lam[r];
    w.a[i] += s;
  }
}

// calcDiff: implicit differentiation of the KKT system,
//   [da; -dlambda] = -Kinv [d tau_rnea/dz (a and the local contact force held fixed); d alpha/dz (a held fixed)],
//   Kinv = [[P, B Ginv], [Ginv B^T, -Ginv]],  P = Minv - B Ginv B^T  (getKKTContactDynamicMatrixInverse),
// alpha = the constrained frame acceleration (ContactModel3D/6D::calcDiff: getJointAccelerationDerivatives, LOCAL).
// Outputs a_q, a_v (nv x nv), P (nv x nv), lam_q, lam_v (nc x nv), GiBt = Ginv B^T (nc x nv).
inline void contact_derivatives(const Model& m, Work& w, const double* vq, double* a_q, double* a_v, double* P,
                                double* lam_q, double* lam_v, double* GiBt) {
  const empc_robot_t& r = m.d->robot;
  const int nv = m.nv, nc = w.nc, jf = r.frame_joint[w.cframe];
  world_accelerations(m, w, vq, w.a, m.a0, w.oa);
  double fl[6] = {w.lam[0], w.lam[1], w.lam[2], nc == 6 ? w.lam[3] : 0.0, nc == 6 ? w.lam[4] : 0.0, nc == 6 ? w.lam[5] : 0.0};
  double Fw[6]; act_force(w.oMf, fl, Fw);
  double M[MAXV * MAXV], dq[MAXV * MAXV], dv[MAXV * MAXV];
  rnea_partials(m, w, Fw, jf, M, dq, dv);
  // d alpha / dq, d alpha / dv
  double dal_q[6][MAXV], dal_v[6][MAXV];
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  for (int c = 0; c < nv; ++c) {
    const int i = m.joint_of_col[c], pi = r.parent[i];
    if (!m.anc[i][jf]) { for (int k = 0; k < 6; ++k) { dal_q[k][c] = 0; dal_v[k][c] = 0; } continue; }
    double s[6], Ap[6];
    for (int k = 0; k < 6; ++k) { s[k] = w.J[k][c]; Ap[k] = pi >= 0 ? w.oa[pi][k] - m.a0[k] : 0.0; }
    const double* Vp = pi >= 0 ? w.ov[pi] : zero6;
    const double* Vd = w.ov[jf];
    double dV[6], t1[6], t2[6], t3[6], aqw[6], vs[6], avw[6];
    cross_mm(Vp, s, dV);
    cross_mm(Ap, s, t1); cross_mm(Vp, dV, t2); cross_mm(dV, Vd, t3);
    for (int k = 0; k < 6; ++k) { aqw[k] = t1[k] + t2[k] + t3[k]; vs[k] = w.ov[i][k] + Vp[k] - Vd[k]; }
    cross_mm(vs, s, avw);
    double vql[6], aql[6], avl[6], fJc[6];
    actinv_motion(w.oMf, dV, vql); actinv_motion(w.oMf, aqw, aql); actinv_motion(w.oMf, avw, avl);
    for (int k = 0; k < 6; ++k) fJc[k] = w.fJ[k][c];
    if (nc == 3) {
      double c1[3], c2[3], c3[3], c4[3];
      cross3(vql + 3, w.vf, c1); cross3(w.vf + 3, vql, c2);
      cross3(fJc + 3, w.vf, c3); cross3(w.vf + 3, fJc, c4);
      for (int k = 0; k < 3; ++k) { dal_q[k][c] = aql[k] + c1[k] + c2[k]; dal_v[k][c] = avl[k] + c3[k] + c4[k]; }
    } else {
      for (int k = 0; k < 6; ++k) { dal_q[k][c] = aql[k]; dal_v[k][c] = avl[k]; }
    }
  }
  for (int rr = 0; rr < nc; ++rr)
    for (int i = 0; i < nv; ++i) {
      double sacc = 0;
      for (int c = 0; c < nc; ++c) sacc += w.Ginv[rr * nc + c] * w.Bc[i][c];
      GiBt[rr * nv + i] = sacc;
    }
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < nv; ++j) {
      double sacc = w.Minv[i * nv + j];
      for (int rr = 0; rr < nc; ++rr) sacc -= w.Bc[i][rr] * GiBt[rr * nv + j];
      P[i * nv + j] = sacc;
    }
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < nv; ++j) {
      double sq = 0, sv = 0;
      for (int k = 0; k < nv; ++k) { sq += P[i * nv + k] * dq[k * nv + j]; sv += P[i * nv + k] * dv[k * nv + j]; }
      for (int rr = 0; rr < nc; ++rr) { sq += GiBt[rr * nv + i] * dal_q[rr][j]; sv += GiBt[rr * nv + i] * dal_v[rr][j]; }
      a_q[i * nv + j] = -sq; a_v[i * nv + j] = -sv;
    }
  for (int rr = 0; rr < nc; ++rr)
    for (int j = 0; j < nv; ++j) {
      double sq = 0, sv = 0;
      for (int k = 0; k < nv; ++k) { sq += GiBt[rr * nv + k] * dq[k * nv + j]; sv += GiBt[rr * nv + k] * dv[k * nv + j]; }
      for (int c = 0; c < nc; ++c) { sq -= w.Ginv[rr * nc + c] * dal_q[c][j]; sv -= w.Ginv[rr * nc + c] * dal_v[c][j]; }
      lam_q[rr * nv + j] = sq; lam_v[rr * nv + j] = sv;
    }
}

// IntegratedActionModelEuler::calcDiff.  `tile` receives Fx|Fu|Lxx|Lxu|Luu|Lx|Lu (Model::o* offsets).
// Must follow diff_calc on the same (x,u) — crocoddyl's convention.
// Differential mode (out_aq != nullptr): the derivatives of the DIFFERENTIAL model only — a_q, a_v (nv x nv), a_u (nv x nu)
// go to out_*, the tile's L blocks receive the cost derivatives unscaled and its F blocks stay zero (what an integrator
// other than Euler builds on).
inline void node_calc_diff_impl(const Model& m, const SolverCtx& ctx, int costset, Work& w, std::vector<CostEval>& evals,
                                double* tile, double* out_aq = nullptr, double* out_av = nullptr, double* out_au = nullptr) {
  const empc_problem_desc_t& d = *m.d;
  const int nv = m.nv, ndx = m.ndx, nu = m.nu, nr = m.nr;
  const double dt = m.dt, dt2 = dt * dt;
  for (int i = 0; i < m.tile; ++i) tile[i] = 0;
  double* Fx = tile + m.oFx; double* Fu = tile + m.oFu; double* Lxx = tile + m.oLxx; double* Lxu = tile + m.oLxu;
  double* Luu = tile + m.oLuu; double* Lx = tile + m.oLx; double* Lu = tile + m.oLu;
  // squashing derivative
  for (int i = 0; i < nu; ++i) {
    if (d.use_squash) {
      const double dd = (d.u_ub[i] - d.u_lb[i]) * ctx.smooth, a = dd * dd;
      const double l = w.u[i] - d.u_lb[i], h = w.u[i] - d.u_ub[i];
      w.ds[i] = 0.5 * ((1.0 / std::sqrt(a + l * l)) * l - (1.0 / std::sqrt(a + h * h)) * h);
    } else {
      w.ds[i] = 1.0;
    }
  }
  double a_q[MAXV * MAXV], a_v[MAXV * MAXV], a_u[MAXV * MAXU];
  double lam_x[6 * MAXDX], lam_u[6 * MAXU];  // d lambda / d(q, v) (nc x ndx) and d lambda / du (nc x nu)
  const double* Pm = w.Minv;                 // d a / d tau
  double Pc[MAXV * MAXV], GiBt[6 * MAXV];
  if (w.nc > 0) {
    double lam_q[6 * MAXV], lam_v[6 * MAXV];
    contact_derivatives(m, w, w.x + m.nq, a_q, a_v, Pc, lam_q, lam_v, GiBt);
    Pm = Pc;
    for (int r = 0; r < w.nc; ++r)
      for (int j = 0; j < nv; ++j) { lam_x[r * ndx + j] = lam_q[r * nv + j]; lam_x[r * ndx + nv + j] = lam_v[r * nv + j]; }
  } else {
    aba_derivatives(m, w, a_q, a_v);
  }
  // Fu_cont = (d a / d tau) * (A diag(ds)),  A = [tau_f 0; 0 I];  d lambda / du = -Ginv B^T A diag(ds)
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < nu; ++j) {
      double s = 0;
      if (j < nr) { for (int k = 0; k < 6; ++k) s += Pm[i * nv + k] * (d.tau_f[k * nr + j] * w.ds[j]); }
      else s = Pm[i * nv + 6 + (j - nr)] * w.ds[j];
      a_u[i * nu + j] = s;
    }
  for (int r = 0; r < w.nc; ++r)
    for (int j = 0; j < nu; ++j) {
      double s = 0;
      if (j < nr) { for (int k = 0; k < 6; ++k) s += GiBt[r * nv + k] * (d.tau_f[k * nr + j] * w.ds[j]); }
      else s = GiBt[r * nv + 6 + (j - nr)] * w.ds[j];
      lam_u[r * nu + j] = -s;
    }
  const bool differential_only = out_aq != nullptr;
  if (differential_only) {
    std::memcpy(out_aq, a_q, sizeof(double) * nv * nv);
    std::memcpy(out_av, a_v, sizeof(double) * nv * nv);
    std::memcpy(out_au, a_u, sizeof(double) * nv * nu);
  }
  // Euler: discrete Jacobians before the Lie-group transport
  for (int i = 0; i < nv && !differential_only; ++i) {
    for (int j = 0; j < nv; ++j) {
      Fx[i * ndx + j] = a_q[i * nv + j] * dt2;
      Fx[i * ndx + nv + j] = a_v[i * nv + j] * dt2;
      Fx[(nv + i) * ndx + j] = a_q[i * nv + j] * dt;
      Fx[(nv + i) * ndx + nv + j] = a_v[i * nv + j] * dt;
    }
    Fx[i * ndx + nv + i] += dt;
    for (int j = 0; j < nu; ++j) { Fu[i * nu + j] = dt2 * a_u[i * nu + j]; Fu[(nv + i) * nu + j] = dt * a_u[i * nu + j]; }
  }
  // JintegrateTransport(x,dx,.,second): rows 0..5 <- Jexp6(dx[0:6]) * rows 0..5
  double Je[36]; Jexp6(w.dx, Je);
  if (!differential_only) {
    double tmp[6][MAXDX];
    for (int a = 0; a < 6; ++a)
      for (int c = 0; c < ndx; ++c) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += Je[6 * a + k] * Fx[k * ndx + c];
        tmp[a][c] = s;
      }
    for (int a = 0; a < 6; ++a) for (int c = 0; c < ndx; ++c) Fx[a * ndx + c] = tmp[a][c];
    for (int a = 0; a < 6; ++a)
      for (int c = 0; c < nu; ++c) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += Je[6 * a + k] * Fu[k * nu + c];
        tmp[a][c] = s;
      }
    for (int a = 0; a < 6; ++a) for (int c = 0; c < nu; ++c) Fu[a * nu + c] = tmp[a][c];
  }
  // Jintegrate(x,dx,first,addto): += blockdiag(Ad(exp6(dx)^-1), I)
  if (!differential_only) {
    SE3 E; exp6(w.dx, E);
    double Xs[36]; force_action_matrix(E, Xs);  // Ad(E^-1) = (X*)^T
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) Fx[a * ndx + b] += Xs[6 * b + a];
    for (int i = 6; i < ndx; ++i) Fx[i * ndx + i] += 1.0;
  }
  // ---- cost derivatives (CostModelSum::calcDiff, Gauss-Newton) ----
  const int c0 = d.costset_begin[costset], c1 = d.costset_begin[costset + 1];
  for (int c = c0; c < c1; ++c) {
    const empc_cost_t& cs = d.costs[c];
    if (!cs.active) continue;
    const CostEval& e = evals[c - c0];
    const double wt = cs.weight;
    switch (cs.type) {
      case EMPC_COST_STATE: {
        // Rx = Jdiff(xref, x, second) = blockdiag(Jlog6(Mref^-1 M), I)
        SE3 Mref, Mx, D; q_to_se3(d.pool + cs.ref_off, Mref); q_to_se3(w.x, Mx); se3_inv_mul(Mref, Mx, D);
        double Jl[36]; Jlog6(D, Jl);
        for (int i = 0; i < 6; ++i) {
          double s = 0;
          for (int k = 0; k < 6; ++k) s += Jl[6 * k + i] * e.Ar[k];
          Lx[i] += wt * s;
          for (int j = 0; j < 6; ++j) {
            double h = 0;
            for (int k = 0; k < 6; ++k) h += Jl[6 * k + i] * (e.Arr[k] * Jl[6 * k + j]);
            Lxx[i * ndx + j] += wt * h;
          }
        }
        for (int i = 6; i < ndx; ++i) { Lx[i] += wt * e.Ar[i]; Lxx[i * ndx + i] += wt * e.Arr[i]; }
      } break;
      case EMPC_COST_CONTROL:
      case EMPC_COST_SQUASH_BARRIER:
        for (int i = 0; i < nu; ++i) { Lu[i] += wt * e.Ar[i]; Luu[i * nu + i] += wt * e.Arr[i]; }
        break;
      case EMPC_COST_CONTACT_FRICTION_CONE: {
        // Rx = A df/dx, Ru = A df/du (ResidualModelContactFrictionCone::calcDiff); the only cost that couples x and u
        if (w.nc == 0) break;
        const double* A = d.pool + cs.ref_off;
        double Rx[5][MAXDX], Ru[5][MAXU];
        for (int a = 0; a < 5; ++a) {
          for (int b = 0; b < ndx; ++b) Rx[a][b] = A[3 * a] * lam_x[b] + A[3 * a + 1] * lam_x[ndx + b] + A[3 * a + 2] * lam_x[2 * ndx + b];
          for (int b = 0; b < nu; ++b) Ru[a][b] = A[3 * a] * lam_u[b] + A[3 * a + 1] * lam_u[nu + b] + A[3 * a + 2] * lam_u[2 * nu + b];
        }
        for (int i = 0; i < ndx; ++i) {
          double s = 0;
          for (int k = 0; k < 5; ++k) s += Rx[k][i] * e.Ar[k];
          Lx[i] += wt * s;
          for (int j = 0; j < ndx; ++j) {
            double h = 0;
            for (int k = 0; k < 5; ++k) h += Rx[k][i] * (e.Arr[k] * Rx[k][j]);
            Lxx[i * ndx + j] += wt * h;
          }
          for (int j = 0; j < nu; ++j) {
            double h = 0;
            for (int k = 0; k < 5; ++k) h += Rx[k][i] * (e.Arr[k] * Ru[k][j]);
            Lxu[i * nu + j] += wt * h;
          }
        }
        for (int i = 0; i < nu; ++i) {
          double s = 0;
          for (int k = 0; k < 5; ++k) s += Ru[k][i] * e.Ar[k];
          Lu[i] += wt * s;
          for (int j = 0; j < nu; ++j) {
            double h = 0;
            for (int k = 0; k < 5; ++k) h += Ru[k][i] * (e.Arr[k] * Ru[k][j]);
            Luu[i * nu + j] += wt * h;
          }
        }
      } break;
      default: {
        // frame costs: Rx = [Rq (nr x nv), Rv (nr x nv)]
        const int f = cs.frame, n = residual_dim(m, cs.type);
        SE3 oMf; frame_placement(m, w, f, oMf);
        double fJ[6][MAXV]; frame_jacobian(m, w, f, oMf, fJ);
        double Rx[6][MAXDX];
        for (int a = 0; a < 6; ++a) for (int b = 0; b < ndx; ++b) Rx[a][b] = 0;
        int ncols = nv;
        if (cs.type == EMPC_COST_FRAME_PLACEMENT) {
          double Jl[36]; Jlog6(e.rMf, Jl);
          for (int a = 0; a < 6; ++a)
            for (int b = 0; b < nv; ++b) {
              double s = 0;
              for (int k = 0; k < 6; ++k) s += Jl[6 * a + k] * fJ[k][b];
              Rx[a][b] = s;
            }
        } else if (cs.type == EMPC_COST_FRAME_ROTATION) {
          double wv[3], th, Jl[9]; log3(e.rMf.R, wv, th); Jlog3(th, wv, Jl);
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < nv; ++b) Rx[a][b] = Jl[3 * a] * fJ[3][b] + Jl[3 * a + 1] * fJ[4][b] + Jl[3 * a + 2] * fJ[5][b];
        } else if (cs.type == EMPC_COST_FRAME_TRANSLATION) {
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < nv; ++b)
              Rx[a][b] = oMf.R[3 * a] * fJ[0][b] + oMf.R[3 * a + 1] * fJ[1][b] + oMf.R[3 * a + 2] * fJ[2][b];
        } else {  // FRAME_VELOCITY, LOCAL: dv_f/dq_c = oMf.actInv(ov_parent(c) x J_c), dv_f/dv = fJ
          const int jf = d.robot.frame_joint[f];
          for (int b = 0; b < nv; ++b) {
            const int k = m.joint_of_col[b], pk = d.robot.parent[k];
            if (m.anc[k][jf] && pk >= 0) {
              double col[6] = {w.J[0][b], w.J[1][b], w.J[2][b], w.J[3][b], w.J[4][b], w.J[5][b]}, cr[6], o[6];
              cross_mm(w.ov[pk], col, cr); actinv_motion(oMf, cr, o);
              for (int a = 0; a < 6; ++a) Rx[a][b] = o[a];
            }
            for (int a = 0; a < 6; ++a) Rx[a][nv + b] = fJ[a][b];
          }
          ncols = ndx;
        }
        for (int i = 0; i < ncols; ++i) {
          double s = 0;
          for (int k = 0; k < n; ++k) s += Rx[k][i] * e.Ar[k];
          Lx[i] += wt * s;
          for (int j = 0; j < ncols; ++j) {
            double h = 0;
            for (int k = 0; k < n; ++k) h += Rx[k][i] * (e.Arr[k] * Rx[k][j]);
            Lxx[i * ndx + j] += wt * h;
          }
        }
      } break;
    }
  }
  if (differential_only) return;
  // Euler scales the cost derivatives by dt (Lxu is zero unless the node has a contact-force cost)
  for (int i = 0; i < ndx * ndx; ++i) Lxx[i] *= dt;
  for (int i = 0; i < ndx * nu; ++i) Lxu[i] *= dt;
  for (int i = 0; i < nu * nu; ++i) Luu[i] *= dt;
  for (int i = 0; i < ndx; ++i) Lx[i] *= dt;
  for (int i = 0; i < nu; ++i) Lu[i] *= dt;
}


// J <- d integrate(x, dx)/d(dx) * J  (+ d integrate/dx when add_first): JintegrateTransport(second) then Jintegrate(first,
// addto) of StateMultibody, for an ndx x ncols row-major matrix J
inline void jintegrate_apply(const Model& m, const double* dx, double* J, int ncols, bool add_first) {
  double Je[36]; Jexp6(dx, Je);
  double tmp[6];
  for (int c = 0; c < ncols; ++c) {
    for (int a = 0; a < 6; ++a) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += Je[6 * a + k] * J[k * ncols + c];
      tmp[a] = s;
    }
    for (int a = 0; a < 6; ++a) J[a * ncols + c] = tmp[a];
  }
  if (add_first) {
    SE3 E; exp6(dx, E);
    double Xs[36]; force_action_matrix(E, Xs);  // Ad(E^-1) = (X*)^T
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) J[a * ncols + b] += Xs[6 * b + a];
    for (int i = 6; i < m.ndx; ++i) J[i * ncols + i] += 1.0;
  }
}

// IntegratedActionModelRK4::calcDiff (crocoddyl/core/integrator/rk4.hxx): chain rule through the four stages,
//   dki_dx = dki_dy dyi_dx,  dki_du = dki_dy dyi_du + [0; a_u],  dki_dy = [[0, I], [a_q, a_v]],
//   dyi_dx = Jint_2(x, c_i dt k_{i-1}) c_i dt dk(i-1)_dx + Jint_1,  dyi_du = Jint_2 c_i dt dk(i-1)_du,
//   Fx = Jint_2(x, dx) dt/6 sum w_i dki_dx + Jint_1,  Fu = Jint_2 dt/6 sum w_i dki_du,
// and the stage costs pulled back through (dyi_dx, dyi_du) with the Gauss-Newton products of rk4.hxx.
inline void node_calc_diff_rk4(const Model& m, const SolverCtx& ctx, int costset, Work& w, double* tile) {
  const int nv = m.nv, ndx = m.ndx, nu = m.nu;
  const double dt = m.dt;
  for (int i = 0; i < m.tile; ++i) tile[i] = 0;
  double* Fx = tile + m.oFx; double* Fu = tile + m.oFu; double* Lxx = tile + m.oLxx; double* Lxu = tile + m.oLxu;
  double* Luu = tile + m.oLuu; double* Lx = tile + m.oLx; double* Lu = tile + m.oLu;
  std::vector<double> st_tile(m.tile);
  std::vector<CostEval> ev;
  static thread_local Work ws;
  double x[MAXX], u[MAXU];
  std::memcpy(x, w.x, sizeof(double) * m.nx); std::memcpy(u, w.u, sizeof(double) * nu);
  std::vector<double> dyx(ndx * ndx, 0.0), dyu(ndx * nu, 0.0), dkx(ndx * ndx), dku(ndx * nu), tA(ndx * ndx), tB(ndx * nu);
  for (int i = 0; i < ndx; ++i) dyx[i * ndx + i] = 1.0;  // stage 0: y_0 = x
  double k[MAXDX], y[MAXX], dxi[MAXDX], a_q[MAXV * MAXV], a_v[MAXV * MAXV], a_u[MAXV * MAXU];
  std::memcpy(y, x, sizeof(double) * m.nx);
  for (int st = 0; st < 4; ++st) {
    if (st > 0) {
      // y_st and its Jacobians from the previous stage's k
      for (int i = 0; i < ndx; ++i) dxi[i] = kRk4C[st] * dt * k[i];
      state_integrate(m, x, dxi, y);
      for (int i = 0; i < ndx * ndx; ++i) dyx[i] = kRk4C[st] * dt * dkx[i];
      for (int i = 0; i < ndx * nu; ++i) dyu[i] = kRk4C[st] * dt * dku[i];
      jintegrate_apply(m, dxi, dyx.data(), ndx, true);
      jintegrate_apply(m, dxi, dyu.data(), nu, false);
    }
    diff_calc(m, ctx, costset, y, u, ws, &ev);
    node_calc_diff_impl(m, ctx, costset, ws, ev, st_tile.data(), a_q, a_v, a_u);
    for (int i = 0; i < nv; ++i) { k[i] = y[m.nq + i]; k[nv + i] = ws.a[i]; }
    // dki_dx = dki_dy dyi_dx ; dki_du = dki_dy dyi_du + [0; a_u]
    for (int i = 0; i < nv; ++i) {
      for (int c = 0; c < ndx; ++c) {
        dkx[i * ndx + c] = dyx[(nv + i) * ndx + c];
        double s = 0;
        for (int j = 0; j < nv; ++j) s += a_q[i * nv + j] * dyx[j * ndx + c] + a_v[i * nv + j] * dyx[(nv + j) * ndx + c];
        dkx[(nv + i) * ndx + c] = s;
      }
      for (int c = 0; c < nu; ++c) {
        dku[i * nu + c] = dyu[(nv + i) * nu + c];
        double s = a_u[i * nu + c];
        for (int j = 0; j < nv; ++j) s += a_q[i * nv + j] * dyu[j * nu + c] + a_v[i * nv + j] * dyu[(nv + j) * nu + c];
        dku[(nv + i) * nu + c] = s;
      }
    }
    const double wg = kRk4W[st] * dt / 6.0;
    for (int i = 0; i < ndx * ndx; ++i) Fx[i] += wg * dkx[i];
    for (int i = 0; i < ndx * nu; ++i) Fu[i] += wg * dku[i];
    // stage cost derivatives (unscaled) pulled back through dyi_dx, dyi_du
    const double* sLxx = st_tile.data() + m.oLxx; const double* sLxu = st_tile.data() + m.oLxu; const double* sLuu = st_tile.data() + m.oLuu;
    const double* sLx = st_tile.data() + m.oLx; const double* sLu = st_tile.data() + m.oLu;
    for (int c = 0; c < ndx; ++c) { double s = 0; for (int i = 0; i < ndx; ++i) s += sLx[i] * dyx[i * ndx + c]; Lx[c] += wg * s; }
    for (int c = 0; c < nu; ++c) { double s = sLu[c]; for (int i = 0; i < ndx; ++i) s += sLx[i] * dyu[i * nu + c]; Lu[c] += wg * s; }
    // tA = Lxx_i dyi_dx (ndx x ndx), tB = Lxx_i dyi_du + Lxu_i (ndx x nu)
    for (int i = 0; i < ndx; ++i) {
      for (int c = 0; c < ndx; ++c) { double s = 0; for (int j = 0; j < ndx; ++j) s += sLxx[i * ndx + j] * dyx[j * ndx + c]; tA[i * ndx + c] = s; }
      for (int c = 0; c < nu; ++c) { double s = sLxu[i * nu + c]; for (int j = 0; j < ndx; ++j) s += sLxx[i * ndx + j] * dyu[j * nu + c]; tB[i * nu + c] = s; }
    }
    for (int r = 0; r < ndx; ++r) {
      for (int c = 0; c < ndx; ++c) { double s = 0; for (int i = 0; i < ndx; ++i) s += dyx[i * ndx + r] * tA[i * ndx + c]; Lxx[r * ndx + c] += wg * s; }
      for (int c = 0; c < nu; ++c) { double s = 0; for (int i = 0; i < ndx; ++i) s += dyx[i * ndx + r] * tB[i * nu + c]; Lxu[r * nu + c] += wg * s; }
    }
    // Luu_i + Lxu_i^T dyi_du + (Lxu_i^T dyi_du)^T + dyi_du^T Lxx_i dyi_du  =  Luu_i + dyi_du^T tB + (Lxu_i^T dyi_du)^T
    for (int r = 0; r < nu; ++r)
      for (int c = 0; c < nu; ++c) {
        double s = sLuu[r * nu + c];
        for (int i = 0; i < ndx; ++i) s += dyu[i * nu + r] * tB[i * nu + c] + sLxu[i * nu + c] * dyu[i * nu + r];
        Luu[r * nu + c] += wg * s;
      }
  }
  jintegrate_apply(m, w.dx, Fx, ndx, true);
  jintegrate_apply(m, w.dx, Fu, nu, false);
}

inline void node_calc_diff(const Model& m, const SolverCtx& ctx, int costset, Work& w, std::vector<CostEval>& evals,
                           double* tile) {
  if (m.d->integrator == EMPC_INTEGRATOR_RK4) node_calc_diff_rk4(m, ctx, costset, w, tile);
  else node_calc_diff_impl(m, ctx, costset, w, evals, tile);
}

}  // namespace orc
