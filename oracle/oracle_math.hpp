// oracle_math.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle).  Parity unpinned: the reference's arithmetic lives in
// Crocoddyl (PepMS fork) + Pinocchio, which are absent from /root/reference and from this machine; this file restates
// their published algorithms from the maths and is validated by finite differences and
// algebraic identities in tests/, not by reference goldens.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use anything in oracle/.
//
// Lie-group and spatial-algebra helpers.  Conventions follow Pinocchio 2.x:
//   Motion = [linear ; angular], Force = [force ; torque], SE3 M=(R,p): x_A = R x_B + p.
//   pinocchio/spatial/explog.hpp (exp3/exp6/Jexp3), pinocchio/spatial/log.hxx (log3/log6/Jlog3/Jlog6),
//   pinocchio/multibody/liegroup/special-euclidean.hpp (integrate/difference/dIntegrate/dDifference).
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

static const double kTaylorPrec = 1.220703125e-4;  // eps^(1/4), TaylorSeriesExpansion<double>::precision<3>()


// ---- numerically stable trigonometric coefficient functions --------------------------------------------------------
// Pinocchio switches to 2-term Taylor series only below eps^(1/4) ~ 1.2e-4 and uses closed forms above, which lose
// up to ~8 digits to cancellation just above the switch (e.g. 1/t^2 - sin t/(2t(1-cos t))).  The reference result is
// the same function; to keep the CPU oracle and the CUDA kernels within 1e-9 of each other regardless of 1-ulp libm
// differences, both evaluate these coefficients by series for t < kSeriesBelow and by closed form above.
static const double kSeriesBelow = 0.2;

// A = sin t / t, B = (1 - cos t)/t^2, C = (t - sin t)/t^3
inline void so3_coef(double t2, double t, double& A, double& B, double& C) {
  if (t < kSeriesBelow) {
    A = 1 + t2 * (-1.0 / 6 + t2 * (1.0 / 120 + t2 * (-1.0 / 5040 + t2 * (1.0 / 362880 - t2 / 39916800))));
    B = 0.5 + t2 * (-1.0 / 24 + t2 * (1.0 / 720 + t2 * (-1.0 / 40320 + t2 * (1.0 / 3628800 - t2 / 479001600))));
    C = 1.0 / 6 + t2 * (-1.0 / 120 + t2 * (1.0 / 5040 + t2 * (-1.0 / 362880 + t2 * (1.0 / 39916800 - t2 / 6227020800.0))));
  } else {
    const double st = std::sin(t), ct = std::cos(t);
    A = st / t; B = (1 - ct) / t2; C = (t - st) / (t2 * t);
  }
}
// alpha = (t/2) cot(t/2), beta = (1 - alpha)/t^2, bdot = (d beta/dt)/t
inline void log_coef(double t, double& alpha, double& beta, double& bdot) {
  const double t2 = t * t;
  if (t < kSeriesBelow) {
    beta = 1.0 / 12 + t2 * (1.0 / 720 + t2 * (1.0 / 30240 + t2 * (1.0 / 1209600 + t2 * (1.0 / 47900160 + t2 * (691.0 / 1307674368000.0)))));
    bdot = 1.0 / 360 + t2 * (1.0 / 7560 + t2 * (1.0 / 201600 + t2 * (1.0 / 5987520 + t2 * (691.0 / 130767436800.0))));
    alpha = 1 - t2 * beta;
  } else {
    const double st = std::sin(t), ct = std::cos(t), inv_2_2ct = 1 / (2 * (1 - ct)), tinv = 1 / t, t2inv = tinv * tinv;
    alpha = t * st * inv_2_2ct;
    beta = t2inv - st * tinv * inv_2_2ct;
    bdot = -2 * t2inv * t2inv + (1 + st * tinv) * t2inv * inv_2_2ct;
  }
}
// c2 = (t^2 + 2 cos t - 2)/(2 t^4), c3 = (2t - 3 sin t + t cos t)/(2 t^5)
inline void q_coef(double t2, double t, double& c2, double& c3) {
  if (t < kSeriesBelow) {
    c2 = 1.0 / 24 + t2 * (-1.0 / 720 + t2 * (1.0 / 40320 + t2 * (-1.0 / 3628800 + t2 / 479001600)));
    c3 = 1.0 / 120 + t2 * (-1.0 / 2520 + t2 * (1.0 / 120960 + t2 * (-1.0 / 9979200 + t2 / 1245404160.0)));
  } else {
    const double st = std::sin(t), ct = std::cos(t), t4 = t2 * t2;
    c2 = (t2 + 2 * ct - 2) / (2 * t4);
    c3 = (2 * t - 3 * st + t * ct) / (2 * t4 * t);
  }
}

struct SE3 {
  double R[9];
  double p[3];
};

inline void cross3(const double* a, const double* b, double* o) {
  double x = a[1] * b[2] - a[2] * b[1];
  double y = a[2] * b[0] - a[0] * b[2];
  double z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void matvec3(const double* R, const double* v, double* o) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
inline void matTvec3(const double* R, const double* v, double* o) {
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
inline void matmul3(const double* A, const double* B, double* C) {  // C = A B (C may not alias)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
inline void matTmul3(const double* A, const double* B, double* C) {  // C = A^T B
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
inline void skew3(const double* v, double* S) {
  S[0] = 0; S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2]; S[4] = 0; S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}

// Eigen::Quaternion::toRotationMatrix (assumes unit norm), q = (x,y,z,w)
inline void quat_to_R(const double* q, double* R) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// Eigen's rotation-matrix -> quaternion (used by pinocchio::quaternion::assignQuaternion)
inline void R_to_quat(const double* R, double* q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
}

inline void exp3(const double* w, double* R) {
  const double t2 = dot3(w, w), t = std::sqrt(t2);
  double A, B, C; so3_coef(t2, t, A, B, C);
  const double dg = 1 - t2 * B;  // cos t
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = B * w[i] * w[j];
  R[1] -= A * w[2]; R[3] += A * w[2];
  R[2] += A * w[1]; R[6] -= A * w[1];
  R[5] -= A * w[0]; R[7] += A * w[0];
  R[0] += dg; R[4] += dg; R[8] += dg;
}

inline void exp6(const double* nu, SE3& M) {
  const double* v = nu; const double* w = nu + 3;
  const double t2 = dot3(w, w), t = std::sqrt(t2);
  double A, B, C; so3_coef(t2, t, A, B, C);
  const double dg = 1 - t2 * B, a_w = C * dot3(w, v);
  double wxv[3]; cross3(w, v, wxv);
  for (int i = 0; i < 3; ++i) M.p[i] = A * v[i] + a_w * w[i] + B * wxv[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M.R[3 * i + j] = B * w[i] * w[j];
  M.R[1] -= A * w[2]; M.R[3] += A * w[2];
  M.R[2] += A * w[1]; M.R[6] -= A * w[1];
  M.R[5] -= A * w[0]; M.R[7] += A * w[0];
  M.R[0] += dg; M.R[4] += dg; M.R[8] += dg;
}

inline void log3(const double* R, double* w, double& theta) {
  double tr = R[0] + R[4] + R[8];
  if (tr >= 3) { tr = 3; theta = 0; }
  else if (tr <= -1) { tr = -1; theta = M_PI; }
  else theta = std::acos((tr - 1) / 2);
  if (theta >= M_PI - 1e-2) {
    const double cphi = -(tr - 1) / 2;
    const double beta = theta * theta / (1 + cphi);
    const double t0 = (R[0] + cphi) * beta, t1 = (R[4] + cphi) * beta, t2 = (R[8] + cphi) * beta;
    w[0] = (R[7] > R[5] ? 1.0 : -1.0) * (t0 > 0 ? std::sqrt(t0) : 0.0);
    w[1] = (R[2] > R[6] ? 1.0 : -1.0) * (t1 > 0 ? std::sqrt(t1) : 0.0);
    w[2] = (R[3] > R[1] ? 1.0 : -1.0) * (t2 > 0 ? std::sqrt(t2) : 0.0);
  } else {
    double A, B, C; so3_coef(theta * theta, theta, A, B, C);
    const double t = (1.0 / A) / 2;
    w[0] = t * (R[7] - R[5]); w[1] = t * (R[2] - R[6]); w[2] = t * (R[3] - R[1]);
  }
}

inline void Jlog3(double theta, const double* w, double* J) {
  double alpha, beta, bdot; log_coef(theta, alpha, beta, bdot);  // Jlog3 = beta w w^T + alpha I + [w]x / 2
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) J[3 * i + j] = beta * w[i] * w[j];
  J[0] += alpha; J[4] += alpha; J[8] += alpha;
  J[1] -= 0.5 * w[2]; J[3] += 0.5 * w[2];
  J[2] += 0.5 * w[1]; J[6] -= 0.5 * w[1];
  J[5] -= 0.5 * w[0]; J[7] += 0.5 * w[0];
}

inline void log6(const SE3& M, double* nu) {
  double w[3], t; log3(M.R, w, t);
  double alpha, beta, bdot; log_coef(t, alpha, beta, bdot);
  double wxp[3]; cross3(w, M.p, wxp);
  const double wTp = dot3(w, M.p);
  for (int i = 0; i < 3; ++i) nu[i] = alpha * M.p[i] - 0.5 * wxp[i] + (beta * wTp) * w[i];
  nu[3] = w[0]; nu[4] = w[1]; nu[5] = w[2];
}

// d log6(M exp6(d)) / d d at d=0 (6x6, [lin;ang] ordering), pinocchio::Jlog6
inline void Jlog6(const SE3& M, double* J) {
  double w[3], t; log3(M.R, w, t);
  double A[9]; Jlog3(t, w, A);
  const double t2 = t * t;
  double alpha, beta, bdot; log_coef(t, alpha, beta, bdot);
  const double* p = M.p;
  const double wTp = dot3(w, p);
  double v3[3];
  for (int i = 0; i < 3; ++i) v3[i] = (bdot * wTp) * w[i] - (t2 * bdot + 2 * beta) * p[i];
  double C[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = v3[i] * w[j] + beta * w[i] * p[j];
  C[0] += wTp * beta; C[4] += wTp * beta; C[8] += wTp * beta;
  C[1] -= 0.5 * p[2]; C[3] += 0.5 * p[2];
  C[2] += 0.5 * p[1]; C[6] -= 0.5 * p[1];
  C[5] -= 0.5 * p[0]; C[7] += 0.5 * p[0];
  double B[9]; matmul3(C, A, B);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      J[6 * i + j] = A[3 * i + j];
      J[6 * i + 3 + j] = B[3 * i + j];
      J[6 * (3 + i) + j] = 0;
      J[6 * (3 + i) + 3 + j] = A[3 * i + j];
    }
}

// Right Jacobian of exp3: exp3(w+dw) ~ exp3(w) exp3(Jexp3 dw), pinocchio::Jexp3
inline void Jexp3(const double* r, double* J) {
  const double n2 = dot3(r, r), n = std::sqrt(n2);
  double a, b, c; so3_coef(n2, n, a, b, c);
  b = -b;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) J[3 * i + j] = c * r[i] * r[j];
  J[0] += a; J[4] += a; J[8] += a;
  J[1] += -b * r[2]; J[3] += b * r[2];
  J[2] += b * r[1]; J[6] += -b * r[1];
  J[5] += -b * r[0]; J[7] += b * r[0];
}

// Right Jacobian of exp6: exp6(nu+d) ~ exp6(nu) exp6(Jexp6 d).  Closed form: Jr(nu) = Jl(-nu) with Barfoot's
// Q block (State Estimation for Robotics, eq. 7.86), [lin;ang] ordering.  Equals pinocchio::Jexp6.
inline void Jexp6(const double* nu, double* J) {
  const double* v = nu; const double* w = nu + 3;
  double A[9]; Jexp3(w, A);
  const double t2 = dot3(w, w), t = std::sqrt(t2);
  double c0a, c0b, c1, c2, c3;  // c1=(t-sin t)/t^3, c2=(t^2+2cos t-2)/(2 t^4), c3=(2t-3 sin t+t cos t)/(2 t^5)
  so3_coef(t2, t, c0a, c0b, c1);
  q_coef(t2, t, c2, c3);
  // Q_r = Q_l(-v,-w) = -1/2 V + c1 (WV + VW - WVW) - c2 (WWV + VWW - 3 WVW) + c3 (WVWW + WWVW)
  double V[9], W[9], WV[9], VW[9], WVW[9], WW[9], WWV[9], VWW[9], WVWW[9], WWVW[9];
  skew3(v, V); skew3(w, W);
  matmul3(W, V, WV); matmul3(V, W, VW); matmul3(WV, W, WVW); matmul3(W, W, WW);
  matmul3(WW, V, WWV); matmul3(V, WW, VWW); matmul3(WVW, W, WVWW); matmul3(W, WVW, WWVW);
  double Q[9];
  for (int i = 0; i < 9; ++i)
    Q[i] = -0.5 * V[i] + c1 * (WV[i] + VW[i] - WVW[i]) - c2 * (WWV[i] + VWW[i] - 3 * WVW[i]) +
           c3 * (WVWW[i] + WWVW[i]);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      J[6 * i + j] = A[3 * i + j];
      J[6 * i + 3 + j] = Q[3 * i + j];
      J[6 * (3 + i) + j] = 0;
      J[6 * (3 + i) + 3 + j] = A[3 * i + j];
    }
}

// ---- SE3 actions on spatial vectors -------------------------------------------------------------------------------
inline void se3_mul(const SE3& A, const SE3& B, SE3& C) {  // C = A*B
  double R[9], p[3];
  matmul3(A.R, B.R, R);
  matvec3(A.R, B.p, p);
  for (int i = 0; i < 3; ++i) p[i] += A.p[i];
  std::memcpy(C.R, R, sizeof(R)); std::memcpy(C.p, p, sizeof(p));
}
inline void se3_inv_mul(const SE3& A, const SE3& B, SE3& C) {  // C = A^-1 * B
  double R[9], d[3], p[3];
  matTmul3(A.R, B.R, R);
  for (int i = 0; i < 3; ++i) d[i] = B.p[i] - A.p[i];
  matTvec3(A.R, d, p);
  std::memcpy(C.R, R, sizeof(R)); std::memcpy(C.p, p, sizeof(p));
}
inline void act_motion(const SE3& M, const double* m, double* o) {  // B -> A
  double Rv[3], Rw[3], pxRw[3];
  matvec3(M.R, m, Rv); matvec3(M.R, m + 3, Rw); cross3(M.p, Rw, pxRw);
  for (int i = 0; i < 3; ++i) { o[i] = Rv[i] + pxRw[i]; o[3 + i] = Rw[i]; }
}
inline void actinv_motion(const SE3& M, const double* m, double* o) {  // A -> B
  double pxw[3], t[3], v[3], w[3];
  cross3(M.p, m + 3, pxw);
  for (int i = 0; i < 3; ++i) t[i] = m[i] - pxw[i];
  matTvec3(M.R, t, v); matTvec3(M.R, m + 3, w);
  for (int i = 0; i < 3; ++i) { o[i] = v[i]; o[3 + i] = w[i]; }
}
inline void act_force(const SE3& M, const double* f, double* o) {
  double Rf[3], Rn[3], pxRf[3];
  matvec3(M.R, f, Rf); matvec3(M.R, f + 3, Rn); cross3(M.p, Rf, pxRf);
  for (int i = 0; i < 3; ++i) { o[i] = Rf[i]; o[3 + i] = Rn[i] + pxRf[i]; }
}
// 6x6 force action matrix X* = [[R,0],[px R, R]]; motion action X = [[R, px R],[0,R]]; X^-1 = X*^T
inline void force_action_matrix(const SE3& M, double* X) {
  double S[9], SR[9]; skew3(M.p, S); matmul3(S, M.R, SR);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      X[6 * i + j] = M.R[3 * i + j]; X[6 * i + 3 + j] = 0;
      X[6 * (3 + i) + j] = SR[3 * i + j]; X[6 * (3 + i) + 3 + j] = M.R[3 * i + j];
    }
}
inline void motion_action_matrix(const SE3& M, double* X) {
  double S[9], SR[9]; skew3(M.p, S); matmul3(S, M.R, SR);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      X[6 * i + j] = M.R[3 * i + j]; X[6 * i + 3 + j] = SR[3 * i + j];
      X[6 * (3 + i) + j] = 0; X[6 * (3 + i) + 3 + j] = M.R[3 * i + j];
    }
}
inline void cross_mm(const double* a, const double* b, double* o) {  // motion x motion
  double t1[3], t2[3], t3[3];
  cross3(a + 3, b, t1); cross3(a, b + 3, t2); cross3(a + 3, b + 3, t3);
  for (int i = 0; i < 3; ++i) { o[i] = t1[i] + t2[i]; o[3 + i] = t3[i]; }
}
inline void cross_mf(const double* a, const double* f, double* o) {  // motion x* force
  double t1[3], t2[3], t3[3];
  cross3(a + 3, f, t1); cross3(a + 3, f + 3, t2); cross3(a, f, t3);
  for (int i = 0; i < 3; ++i) { o[i] = t1[i]; o[3 + i] = t2[i] + t3[i]; }
}
inline void mat6_vec(const double* A, const double* v, double* o) {
  double t[6];
  for (int i = 0; i < 6; ++i) {
    double s = 0;
    for (int j = 0; j < 6; ++j) s += A[6 * i + j] * v[j];
    t[i] = s;
  }
  std::memcpy(o, t, sizeof(t));
}
inline void mat6T_vec(const double* A, const double* v, double* o) {
  double t[6];
  for (int i = 0; i < 6; ++i) {
    double s = 0;
    for (int j = 0; j < 6; ++j) s += A[6 * j + i] * v[j];
    t[i] = s;
  }
  std::memcpy(o, t, sizeof(t));
}
inline double dot6(const double* a, const double* b) {
  double s = 0;
  for (int i = 0; i < 6; ++i) s += a[i] * b[i];
  return s;
}
// spatial inertia matrix from (m, c, Ic)
inline void inertia_matrix(double m, const double* c, const double* Ic, double* Y) {
  double S[9], SS[9]; skew3(c, S); matmul3(S, S, SS);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Y[6 * i + j] = (i == j) ? m : 0.0;
      Y[6 * i + 3 + j] = -m * S[3 * i + j];
      Y[6 * (3 + i) + j] = m * S[3 * i + j];
      Y[6 * (3 + i) + 3 + j] = Ic[3 * i + j] - m * SS[3 * i + j];
    }
}
// Y' = X Y X^T for 6x6
inline void congruence6(const double* X, const double* Y, double* O) {
  double T[36];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += X[6 * i + k] * Y[6 * k + j];
      T[6 * i + j] = s;
    }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += T[6 * i + k] * X[6 * j + k];
      O[6 * i + j] = s;
    }
}
// Cholesky LL^T of n x n (ld = n) in place (lower); returns false if not SPD / NaN
inline bool llt_inplace(double* A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  return true;
}
// solve L L^T x = b in place, nrhs columns stored row-major in Bm (n x nrhs)
inline void llt_solve(const double* L, int n, double* Bm, int nrhs) {
  for (int c = 0; c < nrhs; ++c) {
    for (int i = 0; i < n; ++i) {
      double s = Bm[i * nrhs + c];
      for (int k = 0; k < i; ++k) s -= L[i * n + k] * Bm[k * nrhs + c];
      Bm[i * nrhs + c] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = Bm[i * nrhs + c];
      for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * Bm[k * nrhs + c];
      Bm[i * nrhs + c] = s / L[i * n + i];
    }
  }
}

}  // namespace orc
