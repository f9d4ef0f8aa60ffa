// pair_common.cuh — device functions shared by the pair-registration kernels (k_pair.cu: line-only fast path,
// k_hybrid.cu: point + line path): minimal solver, float pose application, g2o-style edge errors, Huber, the
// LM scratch view, ordered chain sums and the column-wise 6x6 / 3x3 inverse.
#pragma once
#include "lsl_internal.h"
#include "shared/lsl_linalg.h"
#include "shared/lsl_math.h"
#include "shared/lsl_rand.h"
#include <float.h>

using namespace lslm;

#ifndef FULL
#define FULL 0xffffffffu
#endif
#ifndef POSE_THREADS
#define POSE_THREADS 256
#endif
#ifndef POSE_MINB
#define POSE_MINB 2
#endif
#define MD_STRIDE 72  // doubles of gathered data per line match

// ------------------------------------------------------- minimal solver ----
struct Iso { double R[9], t[3]; };

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static __device__ void q2r(const double* q, double* R) {  // utils.cpp:1659-1694
  double a = q[0], b = q[1], c = q[2], d = q[3];
  double nm = sqrt(a * a + b * b + c * c + d * d);
  a = a / nm; b = b / nm; c = c / nm; d = d / nm;
  R[0] = a * a + b * b - c * c - d * d; R[1] = 2 * b * c - 2 * a * d; R[2] = 2 * b * d + 2 * a * c;
  R[3] = 2 * b * c + 2 * a * d; R[4] = a * a - b * b + c * c - d * d; R[5] = 2 * c * d - 2 * a * b;
  R[6] = 2 * b * d - 2 * a * c; R[7] = 2 * c * d + 2 * a * b; R[8] = a * a - b * b - c * c + d * d;
}
__device__ __forceinline__ void skew(const double* v, double* m) {  // vec2SkewMat, utils.cpp:1649
  m[0] = 0; m[1] = -v[2]; m[2] = v[1]; m[3] = v[2]; m[4] = 0; m[5] = -v[0]; m[6] = -v[1]; m[7] = v[0]; m[8] = 0;
}
// computeRelativeMotion_svd (motion.cpp:315-365) for exactly three line pairs; a = query, b = train.
// md[k] points at the gathered data of sampled match k: qA(3) qB(3) tA(3) tB(3).
static __device__ void relmotion_svd3(const double* const md[3], double* R, double* t) {
  double au[9], ad[9], bu[9], bd[9];
  for (int i = 0; i < 3; ++i)
    for (int s = 0; s < 2; ++s) {
      const double* A = md[i] + (s ? 6 : 0);
      const double* B = A + 3;
      double* u = s ? bu + 3 * i : au + 3 * i;
      double* d = s ? bd + 3 * i : ad + 3 * i;
      double l[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
      double m[3] = {(A[0] + B[0]) * 0.5, (A[1] + B[1]) * 0.5, (A[2] + B[2]) * 0.5};
      double inv = 1 / sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
      u[0] = l[0] * inv; u[1] = l[1] * inv; u[2] = l[2] * inv;
      cross3(u, m, d);
    }
  double A[16];
  for (int i = 0; i < 16; ++i) A[i] = 0;
  for (int i = 0; i < 3; ++i) {
    double Ai[16];
    for (int k = 0; k < 16; ++k) Ai[k] = 0;
    double dm[3] = {au[3 * i] - bu[3 * i], au[3 * i + 1] - bu[3 * i + 1], au[3 * i + 2] - bu[3 * i + 2]};
    double dp[3] = {au[3 * i] + bu[3 * i], au[3 * i + 1] + bu[3 * i + 1], au[3 * i + 2] + bu[3 * i + 2]};
    double dn[3] = {bu[3 * i] - au[3 * i], bu[3 * i + 1] - au[3 * i + 1], bu[3 * i + 2] - au[3 * i + 2]};
    Ai[1] = dm[0]; Ai[2] = dm[1]; Ai[3] = dm[2];
    Ai[4] = dn[0]; Ai[8] = dn[1]; Ai[12] = dn[2];
    double S[9]; skew(dp, S);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Ai[(r + 1) * 4 + c + 1] = S[r * 3 + c];
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        double s = 0;
        for (int k = 0; k < 4; ++k) s += Ai[k * 4 + r] * Ai[k * 4 + c];
        A[r * 4 + c] = A[r * 4 + c] + s;
      }
  }
  double w[4], V[16];
  jacobi_sym<4>(A, w, V);
  double q[4] = {V[3], V[7], V[11], V[15]};
  q2r(q, R);
  double uu[9], udr[3] = {0, 0, 0};
  for (int i = 0; i < 9; ++i) uu[i] = 0;
  for (int i = 0; i < 3; ++i) {
    double S[9]; skew(bu + 3 * i, S);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += S[r * 3 + k] * S[c * 3 + k];
        uu[r * 3 + c] = uu[r * 3 + c] + s;
      }
    double Rad[3], v[3];
    for (int r = 0; r < 3; ++r) Rad[r] = R[r * 3] * ad[3 * i] + R[r * 3 + 1] * ad[3 * i + 1] + R[r * 3 + 2] * ad[3 * i + 2];
    for (int r = 0; r < 3; ++r) v[r] = bd[3 * i + r] - Rad[r];
    for (int r = 0; r < 3; ++r) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += S[k * 3 + r] * v[k];
      udr[r] = udr[r] + s;
    }
  }
  double ui[9];
  inv3(uu, ui);
  for (int r = 0; r < 3; ++r) t[r] = ui[r * 3] * udr[0] + ui[r * 3 + 1] * udr[1] + ui[r * 3 + 2] * udr[2];
}

// Eigen Matrix4f * Vector4f: per row (((a0 x) + a1 y) + a2 z) + a3 w in float; tf = 12 floats
__device__ __forceinline__ void tf_apply_f(const float* tf, const float* v, double* out) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float acc = __fmul_rn(tf[r * 4 + 0], v[0]);
    acc = __fadd_rn(__fmul_rn(tf[r * 4 + 1], v[1]), acc);
    acc = __fadd_rn(__fmul_rn(tf[r * 4 + 2], v[2]), acc);
    acc = __fadd_rn(__fmul_rn(tf[r * 4 + 3], 1.0f), acc);
    out[r] = (double)acc;
  }
}
// both Mahalanobis distances of match data md under tf (motion.cpp:688-693)
__device__ __forceinline__ void score_match(const double* md, const float* tf, double* da, double* db) {
  float qa[3] = {(float)md[0], (float)md[1], (float)md[2]}, qb[3] = {(float)md[3], (float)md[4], (float)md[5]};
  double qA[3], qB[3];
  tf_apply_f(tf, qa, qA);
  tf_apply_f(tf, qb, qB);
  *da = mah_dist3d_pt_line(md + 6, md + 12, qA, qB);
  *db = mah_dist3d_pt_line(md + 9, md + 21, qA, qB);
}

// inlier decision of a match under tf (motion.cpp:688-699): da < thr && db < thr, with the division / square root only near
// the threshold (mah_dist3d_pt_line_lt); identical decisions to score_match + compare
__device__ __forceinline__ bool score_match_inlier(const double* md, const float* tf, double thr) {
  float qa[3] = {(float)md[0], (float)md[1], (float)md[2]}, qb[3] = {(float)md[3], (float)md[4], (float)md[5]};
  double qA[3], qB[3];
  tf_apply_f(tf, qa, qA);
  tf_apply_f(tf, qb, qB);
  return mah_dist3d_pt_line_lt(md + 6, md + 12, qA, qB, thr) && mah_dist3d_pt_line_lt(md + 9, md + 21, qA, qB, thr);
}

// ------------------------------------------------- g2o-style refinement ----
__device__ __forceinline__ void iso_mul(const Iso& a, const Iso& b, Iso& c) {
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k) c.R[r * 3 + k] = a.R[r * 3] * b.R[k] + a.R[r * 3 + 1] * b.R[3 + k] + a.R[r * 3 + 2] * b.R[6 + k];
    c.t[r] = a.R[r * 3] * b.t[0] + a.R[r * 3 + 1] * b.t[1] + a.R[r * 3 + 2] * b.t[2] + a.t[r];
  }
}
__device__ __forceinline__ void iso_inv(const Iso& a, Iso& c) {
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) c.R[r * 3 + k] = a.R[k * 3 + r];
  for (int r = 0; r < 3; ++r) c.t[r] = -(c.R[r * 3] * a.t[0] + c.R[r * 3 + 1] * a.t[1] + c.R[r * 3 + 2] * a.t[2]);
}
static __device__ void iso_oplus(const Iso& est, const double* u, Iso& out) {  // VertexSE3::oplusImpl / fromVectorMQT
  double w2 = 1. - (u[3] * u[3] + u[4] * u[4] + u[5] * u[5]);
  double q[4] = {w2 > 0 ? sqrt(w2) : 0.0, u[3], u[4], u[5]};
  Iso inc;
  q2r(q, inc.R);
  inc.t[0] = u[0]; inc.t[1] = u[1]; inc.t[2] = u[2];
  iso_mul(est, inc, out);
}
// EdgeSE3LineEndpts::computeError (edge_se3_lineendpts.cpp:146-189); w2n = pose^-1
static __device__ void edge_error(const Iso& w2n, const double* L, const double* meas, const double* AffA, const double* AffB, double* e) {
  double ptA[3], ptB[3];
  for (int r = 0; r < 3; ++r) {
    ptA[r] = w2n.R[r * 3] * L[0] + w2n.R[r * 3 + 1] * L[1] + w2n.R[r * 3 + 2] * L[2] + w2n.t[r];
    ptB[r] = w2n.R[r * 3] * L[3] + w2n.R[r * 3 + 1] * L[4] + w2n.R[r * 3 + 2] * L[5] + w2n.t[r];
  }
  for (int h = 0; h < 2; ++h) {
    const double* Af = h ? AffB : AffA;
    const double* mp = meas + 3 * h;
    double dA[3] = {ptA[0] - mp[0], ptA[1] - mp[1], ptA[2] - mp[2]}, dB[3] = {ptB[0] - mp[0], ptB[1] - mp[1], ptB[2] - mp[2]};
    double Ap[3], Bp[3], BA[3];
    for (int r = 0; r < 3; ++r) {
      Ap[r] = Af[r * 3] * dA[0] + Af[r * 3 + 1] * dA[1] + Af[r * 3 + 2] * dA[2];
      Bp[r] = Af[r * 3] * dB[0] + Af[r * 3 + 1] * dB[1] + Af[r * 3 + 2] * dB[2];
    }
    for (int r = 0; r < 3; ++r) BA[r] = Bp[r] - Ap[r];
    double tt = -(Ap[0] * BA[0] + Ap[1] * BA[1] + Ap[2] * BA[2]) / (BA[0] * BA[0] + BA[1] * BA[1] + BA[2] * BA[2]);
    for (int r = 0; r < 3; ++r) e[3 * h + r] = Ap[r] + tt * BA[r];
  }
}
static __device__ void affn(const double* cov, double* Af) {  // endpt_AffnMat = D^-1/2 U^T (transformation_estimation.cpp:349-372)
  double A[9], w[3], V[9];
  for (int i = 0; i < 9; ++i) A[i] = cov[i];
  jacobi_sym<3>(A, w, V);
  for (int i = 0; i < 3; ++i) {
    double d = sqrt(1 / w[i]);
    for (int j = 0; j < 3; ++j) Af[i * 3 + j] = d * V[j * 3 + i];
  }
}
__device__ __forceinline__ void huber(double e2, double delta, double* rho) {  // g2o RobustKernelHuber::robustify
  double dsqr = delta * delta;
  if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.; rho[2] = 0.; }
  else { double sqrte = sqrt(e2); rho[0] = 2 * sqrte * delta - dsqr; rho[1] = delta / sqrte; rho[2] = -0.5 * rho[1] / e2; }
}

struct PoseParams {
  double thr, line_weight_g2o, huber_delta;
  int robust, max_iter, min_matches, min_loopclose, line_weight;
};

// Per-match scratch of the LM (doubles), laid out [field][match] blocks inside the pair's slice
#define LM_STRIDE 306
struct LmView {
  double *L, *Lnew, *Hll, *Hpl, *bl, *HllInv, *contrib, *dl, *terms, *chi;  // 6,6,36,36,6,36,42,6,6,2 per match
  size_t cap;     // matches per pair slice; `contrib` is stored [42][cap] (the ordered chain sums read contiguous memory)
  double *J;      // 124 per match: Jl(newer) 36 | Jl(older) 36 | Jp 36 | e(newer) 6 | e(older) 6 | wgt 2 | pad 2
  int32_t* sel;   // [n] index into the pair's match list
  int32_t* okf;   // [n]
};

// ordered sum s = init (+|-) v[0] (+|-) v[stride] ... (n terms): the adds form one dependent chain in the
// reference's order; the loads are issued 16 at a time ahead of it.
template <bool SUB>
__device__ __forceinline__ double chain_sum(double init, const double* __restrict__ v, int stride, int n) {
  double s = init;
  int i = 0;
  for (; i + 16 <= n; i += 16) {
    double t[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) t[k] = v[(size_t)(i + k) * stride];
#pragma unroll
    for (int k = 0; k < 16; ++k) s = SUB ? s - t[k] : s + t[k];
  }
  for (; i < n; ++i) s = SUB ? s - v[(size_t)i * stride] : s + v[(size_t)i * stride];
  return s;
}

// chi2 of all edges (SparseOptimizer::activeRobustChi2), terms in edge order: per match side 0 (newer) then 1
// (older); one thread per (match, side)
static __device__ void chi2_terms(const LmView& V, const double* md_all, int n, const Iso& w2n, const Iso& ident, const double* Lv,
                           const PoseParams& PP) {
  for (int t = threadIdx.x; t < 2 * n; t += blockDim.x) {
    const int i = t >> 1, side = t & 1;
    const double* md = md_all + (size_t)V.sel[i] * MD_STRIDE;
    double e[6];
    edge_error(side ? w2n : ident, Lv + 6 * i, side ? md + 6 : md, side ? md + 54 : md + 36, side ? md + 63 : md + 45, e);
    double c2 = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) c2 += e[k] * PP.line_weight_g2o * e[k];
    if (PP.robust) { double rho[3]; huber(c2, PP.huber_delta, rho); c2 = rho[0]; }
    V.chi[t] = c2;
  }
}

// Column c of inv_lu<6> (shared/lsl_linalg.h: Gauss-Jordan LU with partial pivoting, cv::Mat::inv analogue):
// the columns of the inverse evolve independently, so six threads each run the elimination on a register
// copy of A and keep one column. Row swaps are compare-and-select over the unrolled rows (static indices).
// M = H (row-major, stride 6) with lambda added to the diagonal. Returns 0 if singular.
template <int N>
__device__ __forceinline__ int invN_column(const double* __restrict__ H, double lambda, int c, double* Rc) {
  double A[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < N; ++j) A[i][j] = H[i * N + j];
    A[i][i] += lambda;
    Rc[i] = (i == c) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    int k = i;
    double best = fabs(A[i][i]);
#pragma unroll
    for (int j = i + 1; j < N; ++j)
      if (fabs(A[j][i]) > best) { best = fabs(A[j][i]); k = j; }
    if (best < 2.2250738585072014e-308) return 0;
#pragma unroll
    for (int r = i + 1; r < N; ++r)
      if (r == k) {
#pragma unroll
        for (int j = 0; j < N; ++j) { double t = A[i][j]; A[i][j] = A[r][j]; A[r][j] = t; }
        double t = Rc[i]; Rc[i] = Rc[r]; Rc[r] = t;
      }
    double d = -1.0 / A[i][i];
#pragma unroll
    for (int j = i + 1; j < N; ++j) {
      double alpha = A[j][i] * d;
#pragma unroll
      for (int q = i + 1; q < N; ++q) A[j][q] += alpha * A[i][q];
      Rc[j] += alpha * Rc[i];
    }
    A[i][i] = -d;
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double s = Rc[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) s -= A[i][k] * Rc[k];
    Rc[i] = s * A[i][i];
  }
  return 1;
}


__device__ __forceinline__ int inv6_column(const double* __restrict__ H, double lambda, int c, double* Rc) {
  return invN_column<6>(H, lambda, c, Rc);
}
