// spatial.cuh — per-thread FP64 Lie-group / spatial-algebra device functions (sm_100a).
//
// Conventions follow Pinocchio 2.x (Motion=[lin;ang], Force=[f;n], SE3 (R,p): x_A = R x_B + p), i.e. the maths behind
// pinocchio/spatial/{explog.hpp,log.hxx} and multibody/liegroup/special-euclidean.hpp that the reference reaches through
// crocoddyl::StateMultibody.  Trigonometric coefficient functions are evaluated by
// series below |theta| < 0.2 and in closed form above, so that results are insensitive to 1-ulp libm differences
// (DESIGN.md "Numerics").
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace empc {

#define EMPC_DI __device__ __forceinline__

struct SE3 {
  double R[9];
  double p[3];
};

// ---- cp.async (LDGSTS): asynchronous global -> shared copies, awaited per commit group ---------------------------------
EMPC_DI void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
}
EMPC_DI void cp_async16(double* smem_dst, const double* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
EMPC_DI void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
EMPC_DI void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- branch-free reciprocal square root / square root / reciprocal --------------------------------------------------
// The library sqrt() and 1/x expand to a MUFU seed, Newton steps AND a branch to a slow path for denormal / huge
// operands (BSSY/BSYNC + a call): ~25 instructions and a fetch bubble each, ~80 times per node.  The quantities that
// go through these helpers (squashing radicands, Cholesky pivots, articulated joint inertias) are positive normal
// numbers, so the MUFU seed (2^-22 relative) plus two Newton steps (-> rounding level, <= 2 ulp) is enough; the
// deviation from the correctly rounded value is ~1e-16 relative, far below the 1e-9 parity bar.
EMPC_DI double rsqrt_nr(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double h = 0.5 * x;
  r = r * fma(-h * r, r, 1.5);
  r = r * fma(-h * r, r, 1.5);
  return r;
}
// ... with one third-order (Halley) step instead of two Newton steps: r (1 + e/2 + 3 e^2 / 8), e = 1 - x r^2.  The seed's
// 2^-22 gives e^3 ~ 2^-64; four dependent FP64 operations instead of six (the Cholesky pivots of the Riccati sweep sit on
// the critical path of every node, and an FP64 operation has ~20 cycles of latency on B200).
EMPC_DI double rsqrt_h(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-(x * r), r, 1.0);
  return fma(r * e, fma(e, 0.375, 0.5), r);
}
// sqrt(x) for x >= 0 (0 -> 0, negative -> NaN like sqrt)
#ifndef EMPC_SQRT_HALLEY
#define EMPC_SQRT_HALLEY 1
#endif
EMPC_DI double sqrt_nr(double x) {
  const double r = EMPC_SQRT_HALLEY ? rsqrt_h(x) : rsqrt_nr(x);
  double s = x * r;
  s = fma(fma(-s, s, x), 0.5 * r, s);
  return x == 0.0 ? 0.0 : s;
}
EMPC_DI double rcp_nr(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}

// 1 / x for positive normal x with one third-order step, r (1 + e + e^2), e = 1 - x r: three dependent FP64 operations after
// the MUFU seed (2^-20 relative, so e^3 ~ 2^-60) instead of the four of two Newton steps
EMPC_DI double rcp_h(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
}

// inverse of a symmetric positive definite 3x3 matrix by cofactors (row-major 9, only the upper entries of A are read;
// the result is written symmetric).  Five dependent FP64 operations plus one reciprocal instead of a three-pivot
// factorisation: used for the free-flyer's 6x6 articulated inertia, split in 3x3 blocks (node.cuh: aba_dynamics).
EMPC_DI void inv3_sym(const double* A, double* Ai) {
  const double a00 = A[0], a01 = A[1], a02 = A[2], a11 = A[4], a12 = A[5], a22 = A[8];
  const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
  const double c11 = a00 * a22 - a02 * a02, c12 = a01 * a02 - a00 * a12, c22 = a00 * a11 - a01 * a01;
  const double r = rcp_nr(a00 * c00 + a01 * c01 + a02 * c02);
  Ai[0] = c00 * r; Ai[1] = c01 * r; Ai[2] = c02 * r;
  Ai[3] = Ai[1];   Ai[4] = c11 * r; Ai[5] = c12 * r;
  Ai[6] = Ai[2];   Ai[7] = Ai[5];   Ai[8] = c22 * r;
}

EMPC_DI void cross3(const double* a, const double* b, double* o) {
  const double x = a[1] * b[2] - a[2] * b[1];
  const double y = a[2] * b[0] - a[0] * b[2];
  const double z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
EMPC_DI double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
EMPC_DI void matvec3(const double* R, const double* v, double* o) {
  const double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  const double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  const double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
EMPC_DI void matTvec3(const double* R, const double* v, double* o) {
  const double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  const double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  const double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
EMPC_DI void matmul3(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
EMPC_DI void matTmul3(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
EMPC_DI void skew3(const double* v, double* S) {
  S[0] = 0; S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2]; S[4] = 0; S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}

// ---- stable coefficient functions (same thresholds/series as documented in DESIGN.md) ------------------------------
#define EMPC_SERIES_BELOW 0.2
// A = sin t/t, B = (1-cos t)/t^2, C = (t-sin t)/t^3
EMPC_DI void so3_coef(double t2, double t, double& A, double& B, double& C) {
  if (t < EMPC_SERIES_BELOW) {
    A = 1 + t2 * (-1.0 / 6 + t2 * (1.0 / 120 + t2 * (-1.0 / 5040 + t2 * (1.0 / 362880 - t2 * (1.0 / 39916800)))));
    B = 0.5 + t2 * (-1.0 / 24 + t2 * (1.0 / 720 + t2 * (-1.0 / 40320 + t2 * (1.0 / 3628800 - t2 * (1.0 / 479001600)))));
    C = 1.0 / 6 + t2 * (-1.0 / 120 + t2 * (1.0 / 5040 + t2 * (-1.0 / 362880 + t2 * (1.0 / 39916800 - t2 * (1.0 / 6227020800.0)))));
  } else {
    double st, ct; sincos(t, &st, &ct);
    A = st / t; B = (1 - ct) / t2; C = (t - st) / (t2 * t);
  }
}
// alpha = (t/2)cot(t/2), beta = (1-alpha)/t^2, bdot = beta'(t)/t
EMPC_DI void log_coef(double t, double& alpha, double& beta, double& bdot) {
  const double t2 = t * t;
  if (t < EMPC_SERIES_BELOW) {
    beta = 1.0 / 12 + t2 * (1.0 / 720 + t2 * (1.0 / 30240 + t2 * (1.0 / 1209600 + t2 * (1.0 / 47900160 + t2 * (691.0 / 1307674368000.0)))));
    bdot = 1.0 / 360 + t2 * (1.0 / 7560 + t2 * (1.0 / 201600 + t2 * (1.0 / 5987520 + t2 * (691.0 / 130767436800.0))));
    alpha = 1 - t2 * beta;
  } else {
    double st, ct; sincos(t, &st, &ct);
    const double inv_2_2ct = 1 / (2 * (1 - ct)), tinv = 1 / t, t2inv = tinv * tinv;
    alpha = t * st * inv_2_2ct;
    beta = t2inv - st * tinv * inv_2_2ct;
    bdot = -2 * t2inv * t2inv + (1 + st * tinv) * t2inv * inv_2_2ct;
  }
}
// c2 = (t^2+2cos t-2)/(2t^4), c3 = (2t-3sin t+t cos t)/(2t^5)
EMPC_DI void q_coef(double t2, double t, double& c2, double& c3) {
  if (t < EMPC_SERIES_BELOW) {
    c2 = 1.0 / 24 + t2 * (-1.0 / 720 + t2 * (1.0 / 40320 + t2 * (-1.0 / 3628800 + t2 * (1.0 / 479001600))));
    c3 = 1.0 / 120 + t2 * (-1.0 / 2520 + t2 * (1.0 / 120960 + t2 * (-1.0 / 9979200 + t2 * (1.0 / 1245404160.0))));
  } else {
    double st, ct; sincos(t, &st, &ct);
    const double t4 = t2 * t2;
    c2 = (t2 + 2 * ct - 2) / (2 * t4);
    c3 = (2 * t - 3 * st + t * ct) / (2 * t4 * t);
  }
}

// Eigen::Quaternion::toRotationMatrix, q = (x,y,z,w)
EMPC_DI void quat_to_R(const double* q, double* R) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// Eigen rotation matrix -> quaternion (pinocchio::quaternion::assignQuaternion)
EMPC_DI void R_to_quat(const double* R, double* q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else if (R[0] >= R[4] && R[0] >= R[8]) {  // i=0, j=1, k=2
    t = sqrt(R[0] - R[4] - R[8] + 1.0);
    q[0] = 0.5 * t; t = 0.5 / t;
    q[3] = (R[7] - R[5]) * t; q[1] = (R[3] + R[1]) * t; q[2] = (R[6] + R[2]) * t;
  } else if (R[4] > R[0] && R[4] >= R[8]) {   // i=1, j=2, k=0
    t = sqrt(R[4] - R[8] - R[0] + 1.0);
    q[1] = 0.5 * t; t = 0.5 / t;
    q[3] = (R[2] - R[6]) * t; q[2] = (R[7] + R[5]) * t; q[0] = (R[1] + R[3]) * t;
  } else {                                     // i=2, j=0, k=1
    t = sqrt(R[8] - R[0] - R[4] + 1.0);
    q[2] = 0.5 * t; t = 0.5 / t;
    q[3] = (R[3] - R[1]) * t; q[0] = (R[2] + R[6]) * t; q[1] = (R[5] + R[7]) * t;
  }
}

EMPC_DI void exp3(const double* w, double* R) {
  const double t2 = dot3(w, w), t = sqrt(t2);
  double A, B, C; so3_coef(t2, t, A, B, C);
  const double dg = 1 - t2 * B;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) R[3 * i + j] = B * w[i] * w[j];
  R[1] -= A * w[2]; R[3] += A * w[2];
  R[2] += A * w[1]; R[6] -= A * w[1];
  R[5] -= A * w[0]; R[7] += A * w[0];
  R[0] += dg; R[4] += dg; R[8] += dg;
}

EMPC_DI void exp6(const double* nu, SE3& M) {
  const double* v = nu; const double* w = nu + 3;
  const double t2 = dot3(w, w), t = sqrt(t2);
  double A, B, C; so3_coef(t2, t, A, B, C);
  const double dg = 1 - t2 * B, a_w = C * dot3(w, v);
  double wxv[3]; cross3(w, v, wxv);
#pragma unroll
  for (int i = 0; i < 3; ++i) M.p[i] = A * v[i] + a_w * w[i] + B * wxv[i];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) M.R[3 * i + j] = B * w[i] * w[j];
  M.R[1] -= A * w[2]; M.R[3] += A * w[2];
  M.R[2] += A * w[1]; M.R[6] -= A * w[1];
  M.R[5] -= A * w[0]; M.R[7] += A * w[0];
  M.R[0] += dg; M.R[4] += dg; M.R[8] += dg;
}

#define EMPC_PI 3.14159265358979323846
EMPC_DI void log3(const double* R, double* w, double& theta) {
  double tr = R[0] + R[4] + R[8];
  if (tr >= 3) { tr = 3; theta = 0; }
  else if (tr <= -1) { tr = -1; theta = EMPC_PI; }
  else theta = acos((tr - 1) / 2);
  if (theta >= EMPC_PI - 1e-2) {
    const double cphi = -(tr - 1) / 2;
    const double beta = theta * theta / (1 + cphi);
    const double t0 = (R[0] + cphi) * beta, t1 = (R[4] + cphi) * beta, t2 = (R[8] + cphi) * beta;
    w[0] = (R[7] > R[5] ? 1.0 : -1.0) * (t0 > 0 ? sqrt(t0) : 0.0);
    w[1] = (R[2] > R[6] ? 1.0 : -1.0) * (t1 > 0 ? sqrt(t1) : 0.0);
    w[2] = (R[3] > R[1] ? 1.0 : -1.0) * (t2 > 0 ? sqrt(t2) : 0.0);
  } else {
    double A, B, C; so3_coef(theta * theta, theta, A, B, C);
    const double t = (1.0 / A) / 2;
    w[0] = t * (R[7] - R[5]); w[1] = t * (R[2] - R[6]); w[2] = t * (R[3] - R[1]);
  }
}

EMPC_DI void Jlog3(double theta, const double* w, double* J) {
  double alpha, beta, bdot; log_coef(theta, alpha, beta, bdot);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) J[3 * i + j] = beta * w[i] * w[j];
  J[0] += alpha; J[4] += alpha; J[8] += alpha;
  J[1] -= 0.5 * w[2]; J[3] += 0.5 * w[2];
  J[2] += 0.5 * w[1]; J[6] -= 0.5 * w[1];
  J[5] -= 0.5 * w[0]; J[7] += 0.5 * w[0];
}

EMPC_DI void log6(const SE3& M, double* nu) {
  double w[3], t; log3(M.R, w, t);
  double alpha, beta, bdot; log_coef(t, alpha, beta, bdot);
  double wxp[3]; cross3(w, M.p, wxp);
  const double wTp = dot3(w, M.p);
#pragma unroll
  for (int i = 0; i < 3; ++i) nu[i] = alpha * M.p[i] - 0.5 * wxp[i] + (beta * wTp) * w[i];
  nu[3] = w[0]; nu[4] = w[1]; nu[5] = w[2];
}

// 6x6 Jlog6 as blocks: J = [[A, B],[0, A]]
EMPC_DI void Jlog6_blocks(const SE3& M, double* A, double* B) {
  double w[3], t; log3(M.R, w, t);
  Jlog3(t, w, A);
  const double t2 = t * t;
  double alpha, beta, bdot; log_coef(t, alpha, beta, bdot);
  const double* p = M.p;
  const double wTp = dot3(w, p);
  double v3[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) v3[i] = (bdot * wTp) * w[i] - (t2 * bdot + 2 * beta) * p[i];
  double C[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = v3[i] * w[j] + beta * w[i] * p[j];
  C[0] += wTp * beta; C[4] += wTp * beta; C[8] += wTp * beta;
  C[1] -= 0.5 * p[2]; C[3] += 0.5 * p[2];
  C[2] += 0.5 * p[1]; C[6] -= 0.5 * p[1];
  C[5] -= 0.5 * p[0]; C[7] += 0.5 * p[0];
  matmul3(C, A, B);
}
EMPC_DI void Jlog6(const SE3& M, double* J) {
  double A[9], B[9]; Jlog6_blocks(M, A, B);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      J[6 * i + j] = A[3 * i + j]; J[6 * i + 3 + j] = B[3 * i + j];
      J[6 * (3 + i) + j] = 0; J[6 * (3 + i) + 3 + j] = A[3 * i + j];
    }
}

EMPC_DI void Jexp3(const double* r, double* J) {
  const double n2 = dot3(r, r), n = sqrt(n2);
  double a, b, c; so3_coef(n2, n, a, b, c);
  b = -b;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) J[3 * i + j] = c * r[i] * r[j];
  J[0] += a; J[4] += a; J[8] += a;
  J[1] += -b * r[2]; J[3] += b * r[2];
  J[2] += b * r[1]; J[6] += -b * r[1];
  J[5] += -b * r[0]; J[7] += b * r[0];
}

// Right Jacobian of exp6 as blocks [[A, Q],[0, A]] (Barfoot eq. 7.86 evaluated at -nu)
EMPC_DI void Jexp6_blocks(const double* nu, double* A, double* Q) {
  const double* v = nu; const double* w = nu + 3;
  Jexp3(w, A);
  const double t2 = dot3(w, w), t = sqrt(t2);
  double c0a, c0b, c1, c2, c3;
  so3_coef(t2, t, c0a, c0b, c1);
  q_coef(t2, t, c2, c3);
  double V[9], W[9], WV[9], VW[9], WVW[9], WW[9], WWV[9], VWW[9], WVWW[9], WWVW[9];
  skew3(v, V); skew3(w, W);
  matmul3(W, V, WV); matmul3(V, W, VW); matmul3(WV, W, WVW); matmul3(W, W, WW);
  matmul3(WW, V, WWV); matmul3(V, WW, VWW); matmul3(WVW, W, WVWW); matmul3(W, WVW, WWVW);
#pragma unroll
  for (int i = 0; i < 9; ++i)
    Q[i] = -0.5 * V[i] + c1 * (WV[i] + VW[i] - WVW[i]) - c2 * (WWV[i] + VWW[i] - 3 * WVW[i]) +
           c3 * (WVWW[i] + WWVW[i]);
}

// ---- SE3 actions ---------------------------------------------------------------------------------------------------
EMPC_DI void se3_mul(const SE3& A, const SE3& B, SE3& C) {
  double R[9], p[3];
  matmul3(A.R, B.R, R);
  matvec3(A.R, B.p, p);
#pragma unroll
  for (int i = 0; i < 3; ++i) C.p[i] = p[i] + A.p[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) C.R[i] = R[i];
}
EMPC_DI void se3_inv_mul(const SE3& A, const SE3& B, SE3& C) {
  double R[9], d[3], p[3];
  matTmul3(A.R, B.R, R);
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = B.p[i] - A.p[i];
  matTvec3(A.R, d, p);
#pragma unroll
  for (int i = 0; i < 3; ++i) C.p[i] = p[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) C.R[i] = R[i];
}
EMPC_DI void act_motion(const SE3& M, const double* m, double* o) {
  double Rv[3], Rw[3], pxRw[3];
  matvec3(M.R, m, Rv); matvec3(M.R, m + 3, Rw); cross3(M.p, Rw, pxRw);
#pragma unroll
  for (int i = 0; i < 3; ++i) { o[i] = Rv[i] + pxRw[i]; o[3 + i] = Rw[i]; }
}
EMPC_DI void actinv_motion(const SE3& M, const double* m, double* o) {
  double pxw[3], t[3], v[3], w[3];
  cross3(M.p, m + 3, pxw);
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = m[i] - pxw[i];
  matTvec3(M.R, t, v); matTvec3(M.R, m + 3, w);
#pragma unroll
  for (int i = 0; i < 3; ++i) { o[i] = v[i]; o[3 + i] = w[i]; }
}
EMPC_DI void act_force(const SE3& M, const double* f, double* o) {
  double Rf[3], Rn[3], pxRf[3];
  matvec3(M.R, f, Rf); matvec3(M.R, f + 3, Rn); cross3(M.p, Rf, pxRf);
#pragma unroll
  for (int i = 0; i < 3; ++i) { o[i] = Rf[i]; o[3 + i] = Rn[i] + pxRf[i]; }
}
EMPC_DI void cross_mm(const double* a, const double* b, double* o) {
  double t1[3], t2[3], t3[3];
  cross3(a + 3, b, t1); cross3(a, b + 3, t2); cross3(a + 3, b + 3, t3);
#pragma unroll
  for (int i = 0; i < 3; ++i) { o[i] = t1[i] + t2[i]; o[3 + i] = t3[i]; }
}
EMPC_DI void cross_mf(const double* a, const double* f, double* o) {
  double t1[3], t2[3], t3[3];
  cross3(a + 3, f, t1); cross3(a + 3, f + 3, t2); cross3(a, f, t3);
#pragma unroll
  for (int i = 0; i < 3; ++i) { o[i] = t1[i]; o[3 + i] = t2[i] + t3[i]; }
}
EMPC_DI double dot6(const double* a, const double* b) {
  double s = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) s += a[i] * b[i];
  return s;
}
EMPC_DI void mat6_vec(const double* A, const double* v, double* o) {
  double t[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = 0;
#pragma unroll
    for (int j = 0; j < 6; ++j) s += A[6 * i + j] * v[j];
    t[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) o[i] = t[i];
}
EMPC_DI void mat6T_vec(const double* A, const double* v, double* o) {
  double t[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = 0;
#pragma unroll
    for (int j = 0; j < 6; ++j) s += A[6 * j + i] * v[j];
    t[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) o[i] = t[i];
}
// force action matrix X* = [[R,0],[px R,R]] of M;  motion action X = [[R, px R],[0,R]];  X^-1 = (X*)^T
EMPC_DI void force_action_matrix(const SE3& M, double* X) {
  double S[9], SR[9]; skew3(M.p, S); matmul3(S, M.R, SR);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      X[6 * i + j] = M.R[3 * i + j]; X[6 * i + 3 + j] = 0;
      X[6 * (3 + i) + j] = SR[3 * i + j]; X[6 * (3 + i) + 3 + j] = M.R[3 * i + j];
    }
}
// O = X Y X^T (6x6)
EMPC_DI void congruence6(const double* X, const double* Y, double* O) {
  double T[36];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s += X[6 * i + k] * Y[6 * k + j];
      T[6 * i + j] = s;
    }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s += T[6 * i + k] * X[6 * j + k];
      O[6 * i + j] = s;
    }
}
// in-place lower Cholesky of an N x N matrix (row-major, ld = N); false if not positive definite / NaN
template <int N>
EMPC_DI bool llt_inplace(double* A) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double d = A[j * N + j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= A[j * N + k] * A[j * N + k];
    if (!(d > 0.0)) ok = false;
    d = sqrt(d);
    A[j * N + j] = d;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double s = A[i * N + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= A[i * N + k] * A[j * N + k];
      A[i * N + j] = s / d;
    }
  }
  return ok;
}
// Cholesky that also returns the reciprocals of the diagonal, and substitution that multiplies by them: keeps the long
// FP64 division sequences (~30 dependent instructions each) out of the substitution chains.
template <int N>
EMPC_DI bool llt_inplace_inv(double* A, double* dinv) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double d = A[j * N + j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= A[j * N + k] * A[j * N + k];
    if (!(d > 0.0)) ok = false;
    const double inv = rsqrt_nr(d);  // NaN for d < 0, inf for d == 0: both poison the factor like sqrt()/division would
    A[j * N + j] = d * inv;
    dinv[j] = inv;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double s = A[i * N + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= A[i * N + k] * A[j * N + k];
      A[i * N + j] = s * inv;
    }
  }
  return ok;
}
template <int N>
EMPC_DI void llt_solve_vec_inv(const double* L, const double* dinv, double* b, int stride) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = b[i * stride];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i * N + k] * b[k * stride];
    b[i * stride] = s * dinv[i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double s = b[i * stride];
#pragma unroll
    for (int k = i + 1; k < N; ++k) s -= L[k * N + i] * b[k * stride];
    b[i * stride] = s * dinv[i];
  }
}
// solve L L^T x = b for one right-hand side (strided), in place
template <int N>
EMPC_DI void llt_solve_vec(const double* L, double* b, int stride) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = b[i * stride];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i * N + k] * b[k * stride];
    b[i * stride] = s / L[i * N + i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double s = b[i * stride];
#pragma unroll
    for (int k = i + 1; k < N; ++k) s -= L[k * N + i] * b[k * stride];
    b[i * stride] = s / L[i * N + i];
  }
}

}  // namespace empc
