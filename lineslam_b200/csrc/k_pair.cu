// k_pair.cu — pair registration on the device, one CTA per (query, train) pair (sm_100a):
//   K9  match_lines_kernel  Node::lineMatching (src/node.cpp:1619-1694): gated 72-D descriptor distance
//                           matrix (f64, cv::norm summation order), mutual first-minimum + 0.7 ratio test.
//   K10 pose_kernel         getTransform_PtsLines_ransac, line matches (src/line/motion.cpp:605-849):
//                           500 minimal solves (getTransform_Line_svd / computeRelativeMotion_svd,
//                           motion.cpp:315-365, 581-603), scoring of every hypothesis against every match
//                           (float transform, f64 Mahalanobis, motion.cpp:688-699), first-best selection,
//                           then the iterated refinement with the native g2o-style LM
//                           (getTransformFromHybridMatchesG2O, src/transformation_estimation.cpp:218-461;
//                           EdgeSE3LineEndpts::computeError, src/line/edge_se3_lineendpts.cpp:146-189).
// Pairs are independent (GraphManager::nodeComparisons maps over them, src/graph_manager.cpp:555), so a
// batch of pairs is a grid of CTAs. Sums that feed decisions run in the reference's order.
#include "pair_common.cuh"
#include <stdio.h>
#ifdef POSE_PROFILE
__device__ unsigned long long g_pose_prof[16];
#define PT(k) do { __syncthreads(); if (threadIdx.x == 0) { long long now_ = clock64(); atomicAdd(&g_pose_prof[k], (unsigned long long)(now_ - t_last_)); t_last_ = now_; } } while (0)
#else
#define PT(k)
#endif

// ------------------------------------------------------------ lineMatching ----
__device__ __forceinline__ double cvnorm_diff72(const double* a, const double* b) {
  double result = 0;
  for (int i = 0; i < 72; i += 4) {
    double v0 = a[i] - b[i], v1 = a[i + 1] - b[i + 1];
    result += v0 * v0 + v1 * v1;
    v0 = a[i + 2] - b[i + 2]; v1 = a[i + 3] - b[i + 3];
    result += v0 * v0 + v1 * v1;
  }
  return sqrt(result);
}
__device__ __forceinline__ double pt_to_line_dist2d(const double* p, const double* l) {  // utils.cpp:1250-1264
  double a = l[0], b = l[1], c = l[2], x = p[0], y = p[1];
  return fabs((a * x + b * y + c)) / sqrt(a * a + b * b);
}
__device__ __forceinline__ double norm2d(double x, double y) { return sqrt(x * x + y * y); }
__device__ __forceinline__ double project2d(const double* X, const double* A, const double* B) {  // utils.cpp:1612-1618
  double BX[2] = {X[0] - B[0], X[1] - B[1]}, BA[2] = {A[0] - B[0], A[1] - B[1]};
  double n = norm2d(BA[0], BA[1]);
  return (BX[0] * BA[0] + BX[1] * BA[1]) / n / n;
}
__device__ double lineSegmentOverlap(const lsl_line_rec& a, const lsl_line_rec& b) {  // utils.cpp:1620-1638
  double la = norm2d(a.p[0] - a.q[0], a.p[1] - a.q[1]), lb = norm2d(b.p[0] - b.q[0], b.p[1] - b.q[1]);
  if (la < lb) {
    double lp = project2d(a.p, b.p, b.q), lq = project2d(a.q, b.p, b.q);
    if ((lp < 0 && lq < 0) || (lp > 1 && lq > 1)) return -1;
    return fabs(lp - lq) * lb;
  } else {
    double lp = project2d(b.p, a.p, a.q), lq = project2d(b.q, a.p, a.q);
    if ((lp < 0 && lq < 0) || (lp > 1 && lq > 1)) return -1;
    return fabs(lp - lq) * la;
  }
}

// first-minimum (cv::minMaxLoc) of D[base + k*stride], k < n, over the warp; returns value, *arg = index
__device__ __forceinline__ double warp_first_min(const double* D, size_t stride, int n, int* arg) {
  const int lane = threadIdx.x & 31;
  double best = 1e300;
  int bi = 1 << 30;
  for (int k = lane; k < n; k += 32) {
    double v = D[(size_t)k * stride];
    if (v < best) { best = v; bi = k; }
  }
  for (int o = 16; o; o >>= 1) {
    double t = __shfl_xor_sync(FULL, best, o);
    int k = __shfl_xor_sync(FULL, bi, o);
    if (t < best || (t == best && k < bi)) { best = t; bi = k; }
  }
  *arg = bi;
  return best;
}
// min over k != skip, starting from 100 (node.cpp:1670-1679)
__device__ __forceinline__ double warp_min_except(const double* D, size_t stride, int n, int skip) {
  const int lane = threadIdx.x & 31;
  double best = 100;
  for (int k = lane; k < n; k += 32) {
    if (k == skip) continue;
    double v = D[(size_t)k * stride];
    if (best > v) best = v;
  }
  for (int o = 16; o; o >>= 1) {
    double t = __shfl_xor_sync(FULL, best, o);
    if (best > t) best = t;
  }
  return best;
}

__global__ void __launch_bounds__(256) match_lines_kernel(const LslPairDesc* __restrict__ pairs, double* __restrict__ Dall,
                                                          lsl_match* __restrict__ matches_all, int32_t* __restrict__ nmatch,
                                                          double cosT) {
  // one 64-byte block: [0..7] per-warp hit flags, [8] running output position, rest padding (the compiler reads
  // neighbouring ints with LDS.128; the padding keeps those reads inside the allocation)
  __shared__ __align__(16) int s_sh[16];
  int* s_warp_cnt = s_sh;
  int& s_base = s_sh[8];
  const LslPairDesc pd = pairs[blockIdx.x];
  const int n1 = pd.nq, n2 = pd.nt, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  lsl_match* out = matches_all + pd.m_off;
  if (tid < 16) s_sh[tid] = 0;
  if (n1 == 0 || n2 == 0) { if (tid == 0) nmatch[blockIdx.x] = 0; return; }
  double* D = Dall + pd.d_off;
  double lineDistThresh, descDiffThresh, lineOverlapThresh;
  const double ratio = 0.7;
  if (pd.adjacent) { lineDistThresh = 45; descDiffThresh = 0.85; lineOverlapThresh = 0; }
  else { lineDistThresh = 80; descDiffThresh = 0.7; lineOverlapThresh = -1; }
  for (int e = tid; e < n1 * n2; e += blockDim.x) {
    int i = e / n2, j = e - i * n2;
    const lsl_line_rec& a = pd.q[i];
    const lsl_line_rec& b = pd.t[j];
    double v = 100.0;
    if (a.r[0] * b.r[0] + a.r[1] * b.r[1] > cosT) {
      double dd = 0.25 * pt_to_line_dist2d(a.p, b.lineEq2d) + 0.25 * pt_to_line_dist2d(a.q, b.lineEq2d) +
                  0.25 * pt_to_line_dist2d(b.p, a.lineEq2d) + 0.25 * pt_to_line_dist2d(b.q, a.lineEq2d);
      if (dd < lineDistThresh && lineSegmentOverlap(a, b) > lineOverlapThresh) v = cvnorm_diff72(a.des, b.des);
    }
    D[e] = v;
  }
  __syncthreads();
  if (tid == 0) s_base = 0;
  // rows in chunks of 8 (one per warp); matches are appended in row order
  for (int i0 = 0; i0 < n1; i0 += 8) {
    int i = i0 + warp;
    bool hit = false;
    int minX = 0;
    double minVal = 0;
    if (i < n1) {
      minVal = warp_first_min(D + (size_t)i * n2, 1, n2, &minX);
      if (minVal < descDiffThresh) {
        int minY;
        warp_first_min(D + minX, n2, n1, &minY);
        if (i == minY) {
          double rowmin2 = warp_min_except(D + (size_t)i * n2, 1, n2, minX);
          double colmin2 = warp_min_except(D + minX, n2, n1, minY);
          if (rowmin2 * ratio > minVal && colmin2 * ratio > minVal) hit = pd.q[i].haveDepth && pd.t[minX].haveDepth;
        }
      }
    }
    if (lane == 0) s_warp_cnt[warp] = hit ? 1 : 0;
    __syncthreads();
    if (hit && lane == 0) {
      int pos = s_base;
      for (int k = 0; k < warp; ++k) pos += s_warp_cnt[k];
      if (pos < pd.cap_m) { out[pos].queryIdx = i; out[pos].trainIdx = minX; out[pos].distance = (float)minVal; }
    }
    __syncthreads();
    if (tid == 0) { int c = 0; for (int k = 0; k < 8; ++k) c += s_warp_cnt[k]; s_base += c; }
    __syncthreads();
  }
  if (tid == 0) nmatch[blockIdx.x] = s_base;
}

// getTransformFromHybridMatchesG2O restated (see oracle/oracle_pair.cpp:refine_pose_lines for the derivation):
// pose vertex + one free 6-vector per line match, numeric central-difference Jacobians, Huber, g2o's LM
// damping policy, landmark blocks eliminated exactly. Whole CTA, six threads per match (one per Jacobian
// column / block row); sums over the matches run as ordered chains on dedicated threads.
// tf (12 floats, shared memory) in/out.
__device__ void refine_pose(const LmView& V, const double* md_all, int n, float* tf, int iterations, const PoseParams& PP,
                            double* s_red /* >= 64 doubles shared */, double* s_S /* 144 doubles shared */, Iso* s_ci /* 12, shared */, long long& t_last_) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  if (n == 0) return;
  Iso tfd, cam1, ident;
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) tfd.R[r * 3 + c] = (double)tf[r * 4 + c]; tfd.t[r] = (double)tf[r * 4 + 3]; }
  iso_inv(tfd, cam1);
  for (int i = 0; i < 9; ++i) ident.R[i] = (i % 4 == 0) ? 1 : 0;
  ident.t[0] = ident.t[1] = ident.t[2] = 0;
  for (int t = tid; t < 6 * n; t += nthr) {
    const int i = t / 6, k = t - 6 * i;
    V.L[t] = md_all[(size_t)V.sel[i] * MD_STRIDE + k];
  }
  __syncthreads();
  const double w = PP.line_weight_g2o;
  double lambda = 0, ni = 2, currentChi = 0;
  const double tau = 1e-5, lowS = 1. / 3., upS = 2. / 3.;
  const double del = 1e-9, scalar = 1 / (2 * del);
  for (int it = 0; it < iterations; ++it) {
    Iso w2n;
    iso_inv(cam1, w2n);
    // chi2 of the current estimate: iteration 0 computes it; every later iteration follows an accepted step, whose trial
    // chi2 (tempChi) IS this sum — same L, same pose, same edge order, same adds — so it is carried over, not recomputed
    if (it == 0) chi2_terms(V, md_all, n, w2n, ident, V.L, PP);
    PT(4);
    // the twelve perturbed camera poses (cam1 (+) +-delta e_d)^-1 are the same for every match: once per iteration
    if (tid < 12) {
      double u[6] = {0, 0, 0, 0, 0, 0};
      const int d = tid >> 1;
#pragma unroll
      for (int k = 0; k < 6; ++k) if (k == d) u[k] = (tid & 1) ? -del : del;
      Iso c;
      iso_oplus(cam1, u, c);
      iso_inv(c, s_ci[tid]);
    }
    __syncthreads();
    // ---- numeric Jacobian columns: thread (match i, column d)
    for (int t = tid; t < 6 * n; t += nthr) {
      const int i = t / 6, d = t - 6 * i;
      const double* md = md_all + (size_t)V.sel[i] * MD_STRIDE;
      const double* Li = V.L + 6 * i;
      double* Jm = V.J + (size_t)124 * i;
      double Lp[6], e1[6], e2[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) Lp[k] = Li[k];
      const double lplus = Li[d] + del, lminus = Li[d] + (-del);
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const double* meas = side ? md + 6 : md;
        const double* A1 = side ? md + 54 : md + 36;
        const double* A2 = A1 + 9;
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k == d) Lp[k] = lplus;
        edge_error(side ? w2n : ident, Lp, meas, A1, A2, e1);
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k == d) Lp[k] = lminus;
        edge_error(side ? w2n : ident, Lp, meas, A1, A2, e2);
#pragma unroll
        for (int k = 0; k < 6; ++k) Jm[side * 36 + k * 6 + d] = scalar * (e1[k] - e2[k]);
        if (side) {
#pragma unroll
          for (int k = 0; k < 6; ++k) Lp[k] = Li[k];
          edge_error(s_ci[2 * d], Lp, meas, A1, A2, e1);
          edge_error(s_ci[2 * d + 1], Lp, meas, A1, A2, e2);
#pragma unroll
          for (int k = 0; k < 6; ++k) Jm[72 + k * 6 + d] = scalar * (e1[k] - e2[k]);
        }
        if (d == side) {  // residual and robust weight of this edge (threads d = 0 / 1 of the group)
#pragma unroll
          for (int k = 0; k < 6; ++k) Lp[k] = Li[k];
          double e[6];
          edge_error(side ? w2n : ident, Lp, meas, A1, A2, e);
          double c2 = 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) c2 += e[k] * w * e[k];
          double wgt = w;
          if (PP.robust) { double rho[3]; huber(c2, PP.huber_delta, rho); wgt = rho[1] * w; }
#pragma unroll
          for (int k = 0; k < 6; ++k) Jm[108 + 6 * side + k] = e[k];
          Jm[120 + side] = wgt;
        }
      }
    }
    __syncthreads();
    PT(5);
    // ---- block rows: thread (match i, row a)
    for (int t = tid; t < 6 * n; t += nthr) {
      const int i = t / 6, a = t - 6 * i;
      const double* Jm = V.J + (size_t)124 * i;
      double* hll = V.Hll + 36 * i; double* hpl = V.Hpl + 36 * i; double* cp = V.contrib + i; const size_t cs = V.cap;
      double b = 0, hrow[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const double* Jl = Jm + 36 * side;
        const double* e = Jm + 108 + 6 * side;
        const double wgt = Jm[120 + side];
        double s = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s += Jl[k * 6 + a] * (wgt * e[k]);
        b -= s;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double h = 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) h += Jl[k * 6 + a] * wgt * Jl[k * 6 + c];
          hrow[c] += h;
        }
        if (side) {
          const double* Jp = Jm + 72;
          double sp = 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) sp += Jp[k * 6 + a] * (wgt * e[k]);
          cp[(36 + a) * cs] = sp;
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            double h = 0, g = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) { h += Jp[k * 6 + a] * wgt * Jp[k * 6 + c]; g += Jp[k * 6 + a] * wgt * Jl[k * 6 + c]; }
            cp[(a * 6 + c) * cs] = h;
            hpl[a * 6 + c] = 0.0 + g;
          }
        }
      }
      V.bl[6 * i + a] = b;
#pragma unroll
      for (int c = 0; c < 6; ++c) hll[a * 6 + c] = hrow[c];
    }
    __syncthreads();
    PT(6);
    // ordered sums over the matches: Hpp (36), bp (6) on threads 0..41; chi2 on thread 64
    if (tid < 36) s_S[tid] = chain_sum<false>(0.0, V.contrib + (size_t)tid * V.cap, 1, n);
    else if (tid < 42) s_S[tid] = chain_sum<true>(0.0, V.contrib + (size_t)tid * V.cap, 1, n);
    else if (tid == 64 && it == 0) s_red[0] = chain_sum<false>(0.0, V.chi, 1, 2 * n);
    __syncthreads();
    // Hpp | bp stay in shared memory (s_S[96..137]) for the damping retries: a per-thread copy indexed by tid lived in
    // local memory (42 doubles of stack per thread, read back through L1 / L2 on every retry)
    double* const s_H = s_S + 96;
    if (tid < 42) s_H[tid] = s_S[tid];
    double bp[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) bp[k] = s_S[36 + k];
    if (it == 0) currentChi = s_red[0];
    __syncthreads();
    if (it == 0) {  // computeLambdaInit: tau * max |diagonal entry|
      double md_ = 0;
      for (int t = tid; t < 6 * n; t += nthr) { const int i = t / 6, a = t - 6 * i; md_ = fmax(fabs(V.Hll[36 * i + a * 6 + a]), md_); }
      for (int o = 16; o; o >>= 1) md_ = fmax(md_, __shfl_xor_sync(FULL, md_, o));
      if ((tid & 31) == 0) s_red[1 + (tid >> 5)] = md_;
      __syncthreads();
      double maxDiag = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) maxDiag = fmax(fabs(s_H[a * 6 + a]), maxDiag);
      for (int k = 0; k < nthr / 32; ++k) maxDiag = fmax(maxDiag, s_red[1 + k]);
      __syncthreads();
      lambda = tau * maxDiag;
      ni = 2;
    }
    PT(7);
    double rho = 0;
    int qmax = 0;
    do {
      // (Hll + lambda I)^-1, one column per thread
      for (int t = tid; t < 6 * n; t += nthr) {
        const int i = t / 6, c = t - 6 * i;
        double Rc[6];
        int ok = inv6_column(V.Hll + 36 * i, lambda, c, Rc);
        if (c == 0) V.okf[i] = ok;
#pragma unroll
        for (int k = 0; k < 6; ++k) V.HllInv[36 * i + k * 6 + c] = Rc[k];
      }
      __syncthreads();
      PT(8);
      // Schur terms: thread (match i, row a)
      for (int t = tid; t < 6 * n; t += nthr) {
        const int i = t / 6, a = t - 6 * i;
        const double* hi = V.HllInv + 36 * i; const double* hpl = V.Hpl + 36 * i;
        double* cp = V.contrib + i; const size_t cs = V.cap;
        double T[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double s = 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) s += hpl[a * 6 + k] * hi[k * 6 + c];
          T[c] = s;
        }
        double s = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s += T[k] * V.bl[6 * i + k];
        cp[(36 + a) * cs] = s;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double h = 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) h += T[k] * hpl[c * 6 + k];
          cp[(a * 6 + c) * cs] = h;
        }
      }
      __syncthreads();
      PT(9);
      if (tid < 42) {
        double s0 = s_H[tid];
        if (tid < 36 && (tid / 6 == tid % 6)) s0 += lambda;
        s_S[tid] = chain_sum<true>(s0, V.contrib + (size_t)tid * V.cap, 1, n);
      } else if (tid == 64) {
        int ok = 1;
        for (int i = 0; i < n; ++i) ok &= V.okf[i];
        s_red[2] = (double)ok;
      }
      __syncthreads();
      // 6x6 pose system: threads 0..5 invert S column-wise, then dp
      if (tid < 6) {
        double Rc[6];
        int ok = inv6_column(s_S, 0.0, tid, Rc);
        if (tid == 0 && !ok) s_red[2] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s_S[48 + k * 6 + tid] = Rc[k];
      }
      __syncthreads();
      PT(10);
      double dp[6];
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s += s_S[48 + a * 6 + k] * s_S[36 + k];
        dp[a] = s;
      }
      const bool ok = s_red[2] != 0.0;
      double scale = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) scale += dp[a] * (lambda * dp[a] + bp[a]);
      for (int t = tid; t < 6 * n; t += nthr) {
        const int i = t / 6, a = t - 6 * i;
        const double* hi = V.HllInv + 36 * i; const double* hpl = V.Hpl + 36 * i;
        double s = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          double q = 0;
#pragma unroll
          for (int m = 0; m < 6; ++m) q += hpl[m * 6 + k] * dp[m];
          s += hi[a * 6 + k] * (V.bl[6 * i + k] - q);
        }
        V.dl[t] = s;
        V.terms[t] = s * (lambda * s + V.bl[t]);
        V.Lnew[t] = V.L[t] + s;
      }
      Iso camNew, w2nNew;
      iso_oplus(cam1, dp, camNew);
      iso_inv(camNew, w2nNew);
      __syncthreads();
      PT(11);
      chi2_terms(V, md_all, n, w2nNew, ident, V.Lnew, PP);
      if (tid == nthr - 1) s_red[3] = chain_sum<false>(scale, V.terms, 1, 6 * n);   // overlaps the chi2 terms of the other warps
      __syncthreads();
      if (tid == 64) s_red[4] = chain_sum<false>(0.0, V.chi, 1, 2 * n);
      __syncthreads();
      scale = s_red[3];
      double tempChi = s_red[4];
      __syncthreads();
      PT(12);
      if (!ok) tempChi = DBL_MAX;
      rho = (currentChi - tempChi);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - (2 * rho - 1) * (2 * rho - 1) * (2 * rho - 1);
        alpha = fmin(alpha, upS);
        double scaleFactor = fmax(lowS, alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
        cam1 = camNew;
        for (int i = tid; i < 6 * n; i += nthr) V.L[i] = V.Lnew[i];
        __syncthreads();
      } else {
        lambda *= ni;
        ni *= 2;
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    if (qmax == 10 || rho == 0) break;
  }
  Iso out;
  iso_inv(cam1, out);
  __syncthreads();
  if (tid == 0) {
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) tf[r * 4 + c] = (float)out.R[r * 3 + c]; tf[r * 4 + 3] = (float)out.t[r]; }
  }
  __syncthreads();
}

// score all matches under tf; ordered inlier list -> sel; returns count; *sse_out (double or float accumulation)
__device__ int score_all(const double* md_all, int nm, const float* tf, double thr, double* da_s, double* db_s, int32_t* sel,
                         bool float_sse, double* sse_out, int* s_i /* 2 ints shared */, double* s_d /* 1 double shared */) {
  const int tid = threadIdx.x;
  for (int i = tid; i < nm; i += blockDim.x) {
    double da, db;
    score_match(md_all + (size_t)i * MD_STRIDE, tf, &da, &db);
    da_s[i] = da; db_s[i] = db;
  }
  __syncthreads();
  if (tid == 0) {
    int c = 0;
    float sf = 0; double sd = 0;
    for (int i = 0; i < nm; ++i) {
      double da = da_s[i], db = db_s[i];
      if (da < thr && db < thr) {
        sel[c++] = i;
        if (float_sse) sf += da * da + db * db; else sd += da * da + db * db;
      }
    }
    s_i[0] = c;
    s_d[0] = float_sse ? (double)sf : sd;
  }
  __syncthreads();
  int c = s_i[0];
  *sse_out = s_d[0];
  __syncthreads();
  return c;
}

__global__ void __launch_bounds__(POSE_THREADS, POSE_MINB) pose_kernel(const LslPairDesc* __restrict__ pairs, const lsl_match* __restrict__ matches_all,
                                                            const int32_t* __restrict__ nmatch, LslPairScratch sc, PoseParams PP,
                                                            lsl_pose_rec* __restrict__ out) {
  __shared__ float s_tf[16];
  __shared__ double s_red[64];
  __shared__ double s_S[144];   // S | bp (42) .. inverse (48..83) .. Hpp | bp of the iteration (96..137)
  __shared__ Iso s_ci[12];
  __shared__ int s_i[4];
  __shared__ double s_d[2];
  __shared__ GRand s_rng;
  __shared__ uint16_t s_idx[LSL_MAX_MATCH];
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = POSE_THREADS / 32;
  const LslPairDesc pd = pairs[pair];
  const int nm = min(nmatch[pair], pd.cap_m);
  const lsl_match* ms = matches_all + pd.m_off;
  lsl_pose_rec* rec = out + pair;
  long long t_last_ = clock64(); (void)t_last_;
  // per-pair slices of the scratch
  double* md_all = sc.md + pd.m_off * MD_STRIDE;
  double* da_s = sc.dab + pd.m_off * 2;
  double* db_s = da_s + pd.cap_m;
  int32_t* sel_r = sc.sel + pd.m_off * 3;      // RANSAC inlier set
  int32_t* sel_f = sel_r + pd.cap_m;            // refined inlier set
  int32_t* sel_t = sel_f + pd.cap_m;            // trial set
  float* tfs = sc.tfs + (size_t)pair * sc.max_iter * 12;
  int32_t* cnts = sc.cnts + (size_t)pair * sc.max_iter;
  uint16_t* trip = sc.trip + (size_t)pair * sc.max_iter * 3;
  LmView V;
  {
    double* lm = sc.lm + pd.m_off * LM_STRIDE;
    const size_t c = pd.cap_m;
    V.L = lm; V.Lnew = V.L + 6 * c; V.Hll = V.Lnew + 6 * c; V.Hpl = V.Hll + 36 * c; V.bl = V.Hpl + 36 * c;
    V.HllInv = V.bl + 6 * c; V.contrib = V.HllInv + 36 * c; V.dl = V.contrib + 42 * c; V.terms = V.dl + 6 * c; V.chi = V.terms + 6 * c;
    V.J = V.chi + 2 * c;
    V.cap = c;
    V.okf = sc.okf + pd.m_off;
    V.sel = sel_r;
  }
  if (tid == 0) {
    rec->id_train = pd.id_t; rec->id_query = pd.id_q; rec->found = 0; rec->n_line_matches = nm;
    rec->n_ransac_inliers = 0; rec->n_inliers = 0; rec->rmse = 1e9f; rec->best_iter = -1;
    for (int i = 0; i < 16; ++i) rec->tf[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int i = 0; i < 8; ++i) rec->pad[i] = 0;
    sc.n_inl[pair] = 0; sc.n_rinl[pair] = 0;
  }
  const int nPt = 0, nLn = nm, line_weight = PP.line_weight;
  int min_inlier_nmb = PP.min_matches;
  if (nPt + nLn * line_weight < min_inlier_nmb) return;  // motion.cpp:621-624 (uniform)
  if (min_inlier_nmb > 0.7 * (nPt + nLn * line_weight)) min_inlier_nmb = (int)(0.7 * (nPt + nLn * line_weight));
  if (abs(pd.id_t - pd.id_q) > 50) min_inlier_nmb = PP.min_loopclose;
  const int maxIter = PP.max_iter;
  // ---- gather: qA qB tA tB | t.DU_A t.DU_B | affn(q.covA) affn(q.covB) affn(t.covA) affn(t.covB)
  for (int i = tid; i < nm; i += blockDim.x) {
    const lsl_line_rec& q = pd.q[ms[i].queryIdx];
    const lsl_line_rec& t = pd.t[ms[i].trainIdx];
    double* md = md_all + (size_t)i * MD_STRIDE;
    for (int k = 0; k < 3; ++k) { md[k] = q.A[k]; md[3 + k] = q.B[k]; md[6 + k] = t.A[k]; md[9 + k] = t.B[k]; }
    for (int k = 0; k < 9; ++k) { md[12 + k] = t.DU_A[k]; md[21 + k] = t.DU_B[k]; }
    affn(q.covA, md + 36); affn(q.covB, md + 45); affn(t.covA, md + 54); affn(t.covB, md + 63);
  }
  // ---- the 500 sample triples: one rand() stream, cumulative shuffle (motion.cpp:635-658)
  if (tid == 0) {
    grand_seed(&s_rng, pd.seed);
    uint16_t* idx = s_idx;  // the `indexes` vector
    for (int i = 0; i < nm; ++i) idx[i] = (uint16_t)i;
    for (int it = 0; it < maxIter; ++it) {
      int left = nm;
      for (int k = 0; k < 3; ++k) {
        int r = grand_next(&s_rng) % left;
        uint16_t t = idx[k]; idx[k] = idx[k + r]; idx[k + r] = t;
        --left;
      }
      trip[3 * it] = (uint16_t)idx[0]; trip[3 * it + 1] = (uint16_t)idx[1]; trip[3 * it + 2] = (uint16_t)idx[2];
    }
  }
  __syncthreads();
  PT(0);
  // ---- minimal solutions (getTransform_Line_svd, motion.cpp:581-603)
  for (int h = tid; h < maxIter; h += blockDim.x) {
    const double* mdp[3] = {md_all + (size_t)trip[3 * h] * MD_STRIDE, md_all + (size_t)trip[3 * h + 1] * MD_STRIDE,
                            md_all + (size_t)trip[3 * h + 2] * MD_STRIDE};
    double R[9], t[3];
    relmotion_svd3(mdp, R, t);
    float* tf = tfs + (size_t)h * 12;
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) tf[r * 4 + c] = (float)R[r * 3 + c]; tf[r * 4 + 3] = (float)t[r]; }
  }
  __syncthreads();
  PT(1);
  // ---- scoring: warp per hypothesis, lanes over matches
  for (int h = warp; h < maxIter; h += nwarp) {
    const float* tfh = tfs + (size_t)h * 12;
    float tf[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) tf[k] = tfh[k];
    int c = 0;
    for (int i0 = 0; i0 < nm; i0 += 32) {
      int i = i0 + lane;
      bool in = false;
      if (i < nm) {
        in = score_match_inlier(md_all + (size_t)i * MD_STRIDE, tf, PP.thr);
      }
      c += __popc(__ballot_sync(FULL, in));
    }
    if (lane == 0) cnts[h] = c;
  }
  __syncthreads();
  PT(2);
  // ---- first best (strict >, motion.cpp:714-720)
  if (warp == 0) {
    int best = 0, bh = 1 << 30;
    for (int h = lane; h < maxIter; h += 32) {
      int c = line_weight * cnts[h];
      if (c > best) { best = c; bh = h; }
    }
    for (int o = 16; o; o >>= 1) {
      int b2 = __shfl_xor_sync(FULL, best, o), h2 = __shfl_xor_sync(FULL, bh, o);
      if (b2 > best || (b2 == best && h2 < bh)) { best = b2; bh = h2; }
    }
    if (lane == 0) { s_i[2] = best > 0 ? bh : -1; }
  }
  __syncthreads();
  const int bh = s_i[2];
  if (bh < 0) return;
  if (tid < 12) s_tf[tid] = tfs[(size_t)bh * 12 + tid];
  if (tid >= 12 && tid < 16) s_tf[tid] = tid == 15 ? 1.f : 0.f;
  __syncthreads();
  double sse;
  const int best_cnt = score_all(md_all, nm, s_tf, PP.thr, da_s, db_s, sel_r, true, &sse, s_i, s_d);
  if (tid == 0) rec->best_iter = bh;
  if (best_cnt < 3) return;  // motion.cpp:722-725
  if (tid == 0) {
    rec->n_ransac_inliers = best_cnt;
    sc.n_rinl[pair] = best_cnt;
    float* tr = sc.tf_ransac + (size_t)pair * 16;
    for (int i = 0; i < 16; ++i) tr[i] = s_tf[i];
  }
  const float sum_squared_error = (float)sse;
  PT(3);
  // ---- refinement (motion.cpp:726-839)
  V.sel = sel_r;
  refine_pose(V, md_all, best_cnt, s_tf, 25, PP, s_red, s_S, s_ci, t_last_);
  double refined_rmse = (double)sqrtf(sum_squared_error / (float)best_cnt);   // float division + std::sqrt(float), motion.cpp:731
  int refined_cnt = 0;
  for (int it = 0; it < 20; ++it) {
    double tmp_sse;
    PT(13);
    int c = score_all(md_all, nm, s_tf, PP.thr, da_s, db_s, sel_t, false, &tmp_sse, s_i, s_d);
    PT(14);
    if (c * line_weight > refined_cnt * line_weight) {
      for (int i = tid; i < c; i += blockDim.x) sel_f[i] = sel_t[i];
      refined_cnt = c;
      refined_rmse = sqrt(tmp_sse / (double)c);
      __syncthreads();
      V.sel = sel_f;
      refine_pose(V, md_all, refined_cnt, s_tf, 20, PP, s_red, s_S, s_ci, t_last_);
    } else break;
  }
  __syncthreads();
  if (tid == 0) {
    rec->n_inliers = refined_cnt;
    rec->rmse = (float)refined_rmse;
    for (int i = 0; i < 12; ++i) rec->tf[i] = s_tf[i];
    rec->found = (line_weight * refined_cnt) >= min_inlier_nmb ? 1 : 0;
    sc.n_inl[pair] = refined_cnt;
  }
}

// ------------------------------------------------------------- launchers ----
int lsl_launch_match(lsl_ctx* ctx, int npairs) {
  const double PI_T = 3.14159265;  // lineslam.h:38
  double cosT = lsl_cos(30 * PI_T / 180);
  LSL_KSTART(ctx, LSL_K_MATCH);
  match_lines_kernel<<<npairs, 256, 0, ctx->stream>>>(ctx->pw.d_pairs, ctx->pw.D, ctx->pw.matches, ctx->pw.nmatch, cosT);
  LSL_KSTOP(ctx, LSL_K_MATCH);
  LSL_CUDA(cudaGetLastError());
  return LSL_OK;
}

int lsl_launch_pose(lsl_ctx* ctx, int npairs) {
  const lsl_params& P = ctx->P;
  PoseParams PP;
  PP.thr = P.max_mah_dist_for_inliers; PP.line_weight_g2o = P.g2o_line_error_weight; PP.huber_delta = P.g2o_BA_kernel_delta;
  PP.robust = P.g2o_BA_use_kernel; PP.max_iter = P.ransac_iters_line_motion; PP.min_matches = P.min_feature_matches;
  PP.min_loopclose = P.min_matches_loopclose; PP.line_weight = P.line_match_number_weight;
  LSL_KSTART(ctx, LSL_K_POSE);
  pose_kernel<<<npairs, POSE_THREADS, 0, ctx->stream>>>(ctx->pw.d_pairs, ctx->pw.matches, ctx->pw.nmatch, ctx->pw.sc, PP, ctx->pw.recs);
  LSL_KSTOP(ctx, LSL_K_POSE);
  LSL_CUDA(cudaGetLastError());
#ifdef POSE_PROFILE
  {
    cudaStreamSynchronize(ctx->stream);
    unsigned long long h[16];
    cudaMemcpyFromSymbol(h, g_pose_prof, sizeof(h));
    static const char* nm_[16] = {"gather+samples", "minimal solves", "scoring", "best+score_all", "lm chi2", "lm jacobians", "lm block rows",
                                  "lm chain Hpp", "lm inverse", "lm schur", "lm chain S+solve", "lm update", "lm trial chi2", "loop misc", "score_all", "-"};
    double tot = 0; for (int i = 0; i < 15; ++i) tot += (double)h[i];
    for (int i = 0; i < 15; ++i) fprintf(stderr, "pose phase %-18s %6.2f %%  %.3f ms/pair\n", nm_[i], 100.0 * h[i] / tot, h[i] / 1.965e6 / npairs);
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_pose_prof, z, sizeof(z));
  }
#endif
  return LSL_OK;
}
