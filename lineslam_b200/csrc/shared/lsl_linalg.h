// lsl_linalg.h — small fixed-size dense routines shared by the CPU oracle and the
// sm_100a kernels so both sides execute the same IEEE operation sequence
// (SURVEY.md §7.1 "shared/", Appendix A.3). They stand in for the third-party
// decompositions the reference calls (cv::SVD on symmetric PSD 3x3 / 4x4,
// cv::Mat::inv on 3x3 / 6x6): only sign/permutation-invariant quantities are
// consumed downstream, so a cyclic Jacobi eigen-solver is an admissible
// replacement (Tier-T against the real OpenCV, Tier-E between oracle and GPU).
#pragma once
#include "lsl_math.h"

namespace lslm {

// Cyclic Jacobi for a symmetric NxN matrix (row-major, destroyed). On return
// w[] holds eigenvalues sorted descending and V (row-major) the matching
// eigenvectors in its COLUMNS.
template <int N>
LSL_HD void jacobi_sym(double* A, double* w, double* V) {
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) V[i * N + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < N; ++i) {
      dg += A[i * N + i] * A[i * N + i];
      for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
    }
    if (off <= 1e-34 * dg || off == 0.0) break;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) {
        double apq = A[p * N + q];
        if (apq == 0.0) continue;
        double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < N; ++k) {  // A <- A * J
          double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {  // A <- J^T * A
          double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < N; ++k) {
          double vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = c * vkp - s * vkq;
          V[k * N + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < N; ++i) w[i] = A[i * N + i];
  for (int i = 0; i < N - 1; ++i) {  // selection sort, descending, stable on ties
    int m = i;
    for (int j = i + 1; j < N; ++j)
      if (w[j] > w[m]) m = j;
    if (m != i) {
      double tw = w[i]; w[i] = w[m]; w[m] = tw;
      for (int k = 0; k < N; ++k) {
        double tv = V[k * N + i]; V[k * N + i] = V[k * N + m]; V[k * N + m] = tv;
      }
    }
  }
}

// RandomPoint3d(pos, cov) constructor (src/line/lineslam.h:59-81): cov = U diag(W) U^T,
// W_sqrt = sqrt(W), DU = diag(1/W_sqrt) U^T.
LSL_HD void cov_to_DU(const double cov[9], double DU[9], double W_sqrt[3]) {
  double A[9], w[3], V[9];
  for (int i = 0; i < 9; ++i) A[i] = cov[i];
  jacobi_sym<3>(A, w, V);
  for (int i = 0; i < 3; ++i) {
    W_sqrt[i] = sqrt(w[i]);
    double inv = 1.0 / W_sqrt[i];
    for (int j = 0; j < 3; ++j) DU[i * 3 + j] = inv * V[j * 3 + i];
  }
}

// 3x3 inverse by cofactors (what cv::invert does for n == 3). Returns det.
LSL_HD double inv3(const double a[9], double r[9]) {
  double c00 = a[4] * a[8] - a[5] * a[7];
  double c01 = a[5] * a[6] - a[3] * a[8];
  double c02 = a[3] * a[7] - a[4] * a[6];
  double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  double id = 1.0 / det;
  r[0] = c00 * id; r[1] = (a[2] * a[7] - a[1] * a[8]) * id; r[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  r[3] = c01 * id; r[4] = (a[0] * a[8] - a[2] * a[6]) * id; r[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  r[6] = c02 * id; r[7] = (a[1] * a[6] - a[0] * a[7]) * id; r[8] = (a[0] * a[4] - a[1] * a[3]) * id;
  return det;
}

// In-place NxN inverse by Gauss-Jordan LU with partial pivoting (cv::Mat::inv
// DECOMP_LU analogue for the 6x6 Hessian, src/line/utils.cpp:1044). Returns 0 if singular.
template <int N>
LSL_HD int inv_lu(double* A, double* R) {
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) R[i * N + j] = (i == j) ? 1.0 : 0.0;
  for (int i = 0; i < N; ++i) {
    int k = i;
    for (int j = i + 1; j < N; ++j)
      if (fabs(A[j * N + i]) > fabs(A[k * N + i])) k = j;
    if (fabs(A[k * N + i]) < 2.2250738585072014e-308) return 0;
    if (k != i)
      for (int j = 0; j < N; ++j) {
        double t = A[i * N + j]; A[i * N + j] = A[k * N + j]; A[k * N + j] = t;
        t = R[i * N + j]; R[i * N + j] = R[k * N + j]; R[k * N + j] = t;
      }
    double d = -1.0 / A[i * N + i];
    for (int j = i + 1; j < N; ++j) {
      double alpha = A[j * N + i] * d;
      for (int c = i + 1; c < N; ++c) A[j * N + c] += alpha * A[i * N + c];
      for (int c = 0; c < N; ++c) R[j * N + c] += alpha * R[i * N + c];
    }
    A[i * N + i] = -d;
  }
  for (int i = N - 1; i >= 0; --i)
    for (int c = 0; c < N; ++c) {
      double s = R[i * N + c];
      for (int k = i + 1; k < N; ++k) s -= A[i * N + k] * R[k * N + c];
      R[i * N + c] = s * A[i * N + i];
    }
  return 1;
}

// Mahalanobis point-to-line distance, expanded form of src/line/utils.cpp:796-809.
// pos/DU describe the random point; (q1,q2) the line.
LSL_HD double mah_dist3d_pt_line(const double pos[3], const double DU[9], const double q1[3],
                                 const double q2[3]) {
  double xa = q1[0], ya = q1[1], za = q1[2], xb = q2[0], yb = q2[1], zb = q2[2];
  double c1 = DU[0], c2 = DU[1], c3 = DU[2], c4 = DU[3], c5 = DU[4], c6 = DU[5], c7 = DU[6],
         c8 = DU[7], c9 = DU[8];
  double x1 = pos[0], x2 = pos[1], x3 = pos[2];
  double a1 = c1 * (x1 - xa) + c2 * (x2 - ya) + c3 * (x3 - za);
  double a2 = c4 * (x1 - xa) + c5 * (x2 - ya) + c6 * (x3 - za);
  double a3 = c7 * (x1 - xa) + c8 * (x2 - ya) + c9 * (x3 - za);
  double b1 = c1 * (x1 - xb) + c2 * (x2 - yb) + c3 * (x3 - zb);
  double b2 = c4 * (x1 - xb) + c5 * (x2 - yb) + c6 * (x3 - zb);
  double b3 = c7 * (x1 - xb) + c8 * (x2 - yb) + c9 * (x3 - zb);
  double term1 = a1 * b2 - a2 * b1;
  double term2 = a1 * b3 - a3 * b1;
  double term3 = a2 * b3 - a3 * b2;
  double term4 = c1 * (x1 - xa) - c1 * (x1 - xb) + c2 * (x2 - ya) - c2 * (x2 - yb) + c3 * (x3 - za) - c3 * (x3 - zb);
  double term5 = c4 * (x1 - xa) - c4 * (x1 - xb) + c5 * (x2 - ya) - c5 * (x2 - yb) + c6 * (x3 - za) - c6 * (x3 - zb);
  double term6 = c7 * (x1 - xa) - c7 * (x1 - xb) + c8 * (x2 - ya) - c8 * (x2 - yb) + c9 * (x3 - za) - c9 * (x3 - zb);
  return sqrt((term1 * term1 + term2 * term2 + term3 * term3) /
              (term4 * term4 + term5 * term5 + term6 * term6));
}

// Decision form of the same distance: mah_dist3d_pt_line(pos, DU, q1, q2) < thr, bit for bit, without paying for the
// division and the square root in the clear cases. num and den are the very sums above; with q = RN(num / den) and
// s = RN(sqrt(q)) the computed s differs from the real sqrt(num / den) by less than 2^-52 relative, so
// num < den * thr^2 * (1 - 1e-9) implies s < thr and num > den * thr^2 * (1 + 1e-9) implies s > thr (margins seven orders
// of magnitude above the rounding); everything else — including NaN and den == 0 — takes the exact path.
LSL_HD bool mah_dist3d_pt_line_lt(const double pos[3], const double DU[9], const double q1[3], const double q2[3], double thr) {
  double xa = q1[0], ya = q1[1], za = q1[2], xb = q2[0], yb = q2[1], zb = q2[2];
  double c1 = DU[0], c2 = DU[1], c3 = DU[2], c4 = DU[3], c5 = DU[4], c6 = DU[5], c7 = DU[6],
         c8 = DU[7], c9 = DU[8];
  double x1 = pos[0], x2 = pos[1], x3 = pos[2];
  double a1 = c1 * (x1 - xa) + c2 * (x2 - ya) + c3 * (x3 - za);
  double a2 = c4 * (x1 - xa) + c5 * (x2 - ya) + c6 * (x3 - za);
  double a3 = c7 * (x1 - xa) + c8 * (x2 - ya) + c9 * (x3 - za);
  double b1 = c1 * (x1 - xb) + c2 * (x2 - yb) + c3 * (x3 - zb);
  double b2 = c4 * (x1 - xb) + c5 * (x2 - yb) + c6 * (x3 - zb);
  double b3 = c7 * (x1 - xb) + c8 * (x2 - yb) + c9 * (x3 - zb);
  double term1 = a1 * b2 - a2 * b1;
  double term2 = a1 * b3 - a3 * b1;
  double term3 = a2 * b3 - a3 * b2;
  double term4 = c1 * (x1 - xa) - c1 * (x1 - xb) + c2 * (x2 - ya) - c2 * (x2 - yb) + c3 * (x3 - za) - c3 * (x3 - zb);
  double term5 = c4 * (x1 - xa) - c4 * (x1 - xb) + c5 * (x2 - ya) - c5 * (x2 - yb) + c6 * (x3 - za) - c6 * (x3 - zb);
  double term6 = c7 * (x1 - xa) - c7 * (x1 - xb) + c8 * (x2 - ya) - c8 * (x2 - yb) + c9 * (x3 - za) - c9 * (x3 - zb);
  const double num = term1 * term1 + term2 * term2 + term3 * term3;
  const double den = term4 * term4 + term5 * term5 + term6 * term6;
  const double t2 = thr * thr;
  if (thr > 0.0 && num < den * (t2 * (1.0 - 1e-9))) return true;
  if (thr > 0.0 && num > den * (t2 * (1.0 + 1e-9))) return false;
  return sqrt(num / den) < thr;
}

// compPt3dCov (src/line/utils.cpp:690-722): cov = J diag(s^2, s^2, sz^2) J^T with
// J = [[z/f,0,x/z],[0,z/f,y/z],[0,0,1]], products evaluated as (J*S)*J^T, zeros included
// as in the dense 3x3 Armadillo products (adding +0.0 terms is exact).
LSL_HD void pt3d_cov(const double pt[3], double f, double sigma_impt, double c1, double c2,
                     double c3, double time_diff, double cov[9]) {
  double c2e = c2 + (time_diff - 0.005 > 0.0 ? time_diff - 0.005 : 0.0) * 0.5;
  double d = pt[2];
  double sz = c1 * d * d + c2e * d + c3;
  double J[9] = {pt[2] / f, 0, pt[0] / pt[2], 0, pt[2] / f, pt[1] / pt[2], 0, 0, 1};
  double S[3] = {sigma_impt * sigma_impt, sigma_impt * sigma_impt, sz * sz};
  double JS[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) JS[i * 3 + j] = J[i * 3 + j] * S[j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += JS[i * 3 + k] * J[j * 3 + k];
      cov[i * 3 + j] = s;
    }
}

// jac_rpt2ln_mahvec_wrt_ln (src/line/utils.cpp:1086-1115), structured form of the generated
// expressions: u = DU(x-a), e = DU(x-a) - DU(x-b), s = u.e, n = e.e, cu_j = DU(:,j).u, ce_j = DU(:,j).e
//   d/da_j [k] = DU_kj - DU_kj*s/n - e_k*(cu_j + ce_j)/n + (1/n^2)*s*(2 ce_j)*e_k
//   d/db_j [k] = cu_j*e_k/n + DU_kj*s/n - (1/n^2)*s*(2 ce_j)*e_k
// J is 3x6 row-major.
LSL_HD void jac_rpt2ln(const double pos[3], const double c[9], const double l[6], double J[18]) {
  double da[3] = {pos[0] - l[0], pos[1] - l[1], pos[2] - l[2]};
  double db[3] = {pos[0] - l[3], pos[1] - l[4], pos[2] - l[5]};
  double u[3], e[3];
  for (int k = 0; k < 3; ++k) {
    u[k] = c[k * 3] * da[0] + c[k * 3 + 1] * da[1] + c[k * 3 + 2] * da[2];
    e[k] = c[k * 3] * da[0] - c[k * 3] * db[0] + c[k * 3 + 1] * da[1] - c[k * 3 + 1] * db[1] + c[k * 3 + 2] * da[2] -
           c[k * 3 + 2] * db[2];
  }
  double s = u[0] * e[0] + u[1] * e[1] + u[2] * e[2];
  double n = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
  double inv_n2 = 1.0 / (n * n);
  for (int j = 0; j < 3; ++j) {
    double cu = c[j] * u[0] + c[3 + j] * u[1] + c[6 + j] * u[2];
    double ce = c[j] * e[0] + c[3 + j] * e[1] + c[6 + j] * e[2];
    double ce2 = c[j] * e[0] * 2.0 + c[3 + j] * e[1] * 2.0 + c[6 + j] * e[2] * 2.0;
    for (int k = 0; k < 3; ++k) {
      double ckj = c[k * 3 + j];
      J[k * 6 + j] = ckj - (ckj * s) / n - (e[k] * (cu + ce)) / n + inv_n2 * s * ce2 * e[k];
      J[k * 6 + 3 + j] = (cu * e[k]) / n + (ckj * s) / n - inv_n2 * s * ce2 * e[k];
    }
  }
}

// (v - pt)^T Cinv (v - pt): end-point residual of costFun_MLEstimateLine3d (utils.cpp:966-971),
// Eigen evaluation order (row vector times matrix first).
LSL_HD double mah_sq_pt(const double e[3], const double pt[3], const double C[9]) {
  double v[3] = {e[0] - pt[0], e[1] - pt[1], e[2] - pt[2]};
  double r0 = v[0] * C[0] + v[1] * C[3] + v[2] * C[6];
  double r1 = v[0] * C[1] + v[1] * C[4] + v[2] * C[7];
  double r2 = v[0] * C[2] + v[1] * C[5] + v[2] * C[8];
  return r0 * v[0] + r1 * v[1] + r2 * v[2];
}

}  // namespace lslm
