// k_hybrid.cu — pair registration with point AND line features (sm_100a), one CTA per pair:
//   match_points_kernel / match_points_accept_kernel
//        Node::featureMatching, BRUTEFORCE branch (src/node.cpp:606-641): exact float L2 (OpenCV 2.4 normL2Sqr_
//        order) k = 2 nearest neighbours per query row, then the serial ratio / unique-train / rand()-jitter pass.
//   pose_hybrid_kernel
//        getTransform_PtsLines_ransac (src/line/motion.cpp:605-849) with both modalities: 3-samples from one
//        shuffle stream, all-line samples -> getTransform_Line_svd, mixed samples -> getTransform_Lns_Pts_pcl
//        (PCL weighted Kabsch in float, motion.cpp:530-579), points scored with errorFunction2
//        (src/misc.cpp:699-786), lines with the two Mahalanobis end-point distances, first best, then the
//        iterated refinement getTransformFromHybridMatchesG2O (src/transformation_estimation.cpp:218-461) with
//        EdgeSE3PointXYZ (src/line/edge_se3_ptxyz.cpp:84-90) and EdgeSE3LineEndpts edges.
// The operation order follows oracle/oracle_pair.cpp (refine_pose_hybrid, getTransform_PtsLines_ransac); with no
// point matches the result is identical to k_pair.cu's line-only pose_kernel.
#include "pair_common.cuh"
#include <stdlib.h>
#include "shared/lsl_points.h"

#define PMD_STRIDE 22   // doubles per point match: q xyz1 + t xyz1 (8 floats = 4 doubles) | Omega_q 9 | Omega_t 9
#define PLM_STRIDE 160  // doubles of LM scratch per point match

struct HybParams {
  double sigma_depth, nn_ratio, fx, sigma_impt, c1, c2, c3, dt;
};

// ------------------------------------------------------------ featureMatching ----
struct Knn2 { float d1, d2; int i1; };
__device__ __forceinline__ Knn2 knn_merge(Knn2 a, Knn2 b) {
  if (b.d1 < a.d1 || (b.d1 == a.d1 && b.i1 < a.i1)) { Knn2 t = a; a = b; b = t; }
  a.d2 = fminf(a.d2, b.d1);
  return a;
}

__global__ void __launch_bounds__(256) match_points_kernel(const LslPairPts* __restrict__ pp, Knn2* __restrict__ knn_all) {
  extern __shared__ float s_q[];   // [8][dim]
  const LslPairPts pd = pp[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  if (i >= pd.nqp) return;
  float* q = s_q + warp * pd.dim;
  for (int k = lane; k < pd.dim; k += 32) q[k] = pd.qd[(size_t)i * pd.dim + k];
  __syncwarp();
  Knn2 best; best.d1 = FLT_MAX; best.d2 = FLT_MAX; best.i1 = 1 << 30;
  for (int j = lane; j < pd.ntp; j += 32) {
    float d = sqrtf(l2sqr_f(q, pd.td + (size_t)j * pd.dim, pd.dim));
    if (d < best.d1) { best.d2 = best.d1; best.d1 = d; best.i1 = j; }
    else if (d < best.d2) best.d2 = d;
  }
  for (int o = 16; o; o >>= 1) {
    Knn2 other;
    other.d1 = __shfl_xor_sync(FULL, best.d1, o); other.d2 = __shfl_xor_sync(FULL, best.d2, o);
    other.i1 = __shfl_xor_sync(FULL, best.i1, o);
    best = knn_merge(best, other);
  }
  if (lane == 0) knn_all[pd.knn_off + i] = best;
}

// "BruteForce-HammingLUT" (ORB rows, src/node.cpp:609-613): the distance is the popcount of the XOR over the row, handed to
// the ratio test as a float; same k = 2 rule (strict <, earlier train row first on ties). Rows are dim bytes (multiple of 4).
__global__ void __launch_bounds__(256) match_points_hamming_kernel(const LslPairPts* __restrict__ pp, Knn2* __restrict__ knn_all) {
  extern __shared__ float s_q[];   // [8][dim / 4] words
  const LslPairPts pd = pp[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  if (i >= pd.nqp) return;
  const int nw = pd.dim >> 2;
  uint32_t* q = reinterpret_cast<uint32_t*>(s_q) + warp * nw;
  const uint32_t* qd = reinterpret_cast<const uint32_t*>(pd.qd);
  const uint32_t* td = reinterpret_cast<const uint32_t*>(pd.td);
  for (int k = lane; k < nw; k += 32) q[k] = qd[(size_t)i * nw + k];
  __syncwarp();
  Knn2 best; best.d1 = FLT_MAX; best.d2 = FLT_MAX; best.i1 = 1 << 30;
  for (int j = lane; j < pd.ntp; j += 32) {
    int c = 0;
    for (int k = 0; k < nw; ++k) c += __popc(q[k] ^ td[(size_t)j * nw + k]);
    const float d = (float)c;
    if (d < best.d1) { best.d2 = best.d1; best.d1 = d; best.i1 = j; }
    else if (d < best.d2) best.d2 = d;
  }
  for (int o = 16; o; o >>= 1) {
    Knn2 other;
    other.d1 = __shfl_xor_sync(FULL, best.d1, o); other.d2 = __shfl_xor_sync(FULL, best.d2, o);
    other.i1 = __shfl_xor_sync(FULL, best.i1, o);
    best = knn_merge(best, other);
  }
  if (lane == 0) knn_all[pd.knn_off + i] = best;
}

// squareroot_descriptor_space (src/node.cpp:1823-1837) in place: thread per row (the L1 sum is a sequential float chain)
// up to 32 rows per CTA staged through shared memory (coalesced 128-byte loads and stores; row stride dim + 1 keeps the
// per-thread walks conflict-free); the row sum is the sequential float sum of cv::reduce(SUM), one thread per row.
__global__ void __launch_bounds__(128) rootsift_kernel(float* __restrict__ desc, int n, int dim, int rpc) {
  extern __shared__ float s_rows[];                 // [rpc][dim + 1], rpc <= 32 rows per CTA
  const int r0 = blockIdx.x * rpc, nr = min(rpc, n - r0), ld = dim + 1;
  float* base = desc + (size_t)r0 * dim;
  for (int e = threadIdx.x; e < nr * dim; e += blockDim.x) s_rows[(e / dim) * ld + e % dim] = fabsf(base[e]);
  __syncthreads();
  if (threadIdx.x < nr) {
    float* d = s_rows + threadIdx.x * ld;
    float sum = 0.f;
    for (int c = 0; c < dim; ++c) sum = __fadd_rn(sum, d[c]);
    if (sum != 0.f)
      for (int c = 0; c < dim; ++c) d[c] = __fsqrt_rn(__fdiv_rn(d[c], sum));
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nr * dim; e += blockDim.x) base[e] = s_rows[(e / dim) * ld + e % dim];
}

// serial acceptance pass (row order): ratio test, unique trainIdx, distance jitter from rand()
__global__ void __launch_bounds__(32) match_points_accept_kernel(const LslPairPts* __restrict__ pp, const LslPairDesc* __restrict__ pairs,
                                                                 const Knn2* __restrict__ knn_all, lsl_match* __restrict__ pm_all,
                                                                 int32_t* __restrict__ npm, int32_t* __restrict__ rng_out, double nn_ratio) {
  __shared__ uint32_t s_used[LSL_MAX_POINTS / 32];
  __shared__ GRand s_rng;
  const LslPairPts pd = pp[blockIdx.x];
  for (int k = threadIdx.x; k < LSL_MAX_POINTS / 32; k += 32) s_used[k] = 0u;
  __syncwarp();
  if (threadIdx.x != 0) return;
  grand_seed(&s_rng, pairs[blockIdx.x].seed);
  int cnt = 0;
  lsl_match* out = pm_all + pd.pm_off;
  if (pd.nqp > 0 && pd.ntp >= 2) {
    const Knn2* knn = knn_all + pd.knn_off;
    for (int i = 0; i < pd.nqp; ++i) {
      const Knn2 k = knn[i];
      const float dist_ratio_fac = k.d1 / k.d2;
      if ((double)dist_ratio_fac < nn_ratio) {
        if (s_used[k.i1 >> 5] & (1u << (k.i1 & 31))) continue;
        s_used[k.i1 >> 5] |= 1u << (k.i1 & 31);
        lsl_match m; m.queryIdx = i; m.trainIdx = k.i1;
        m.distance = (float)((double)dist_ratio_fac + (double)(float)grand_next(&s_rng) / (1000.0 * 2147483647));
        if (cnt < pd.cap_pm) out[cnt] = m;
        ++cnt;
      }
    }
  }
  npm[blockIdx.x] = cnt < pd.cap_pm ? cnt : pd.cap_pm;
  int32_t* st = rng_out + (size_t)blockIdx.x * 33;
  for (int k = 0; k < 31; ++k) st[k] = s_rng.r[k];
  st[31] = s_rng.f; st[32] = s_rng.b;
}

// ------------------------------------------------------------------- refinement ----
// EdgeSE3PointXYZ::computeError: e = w2n * X - measurement
__device__ __forceinline__ void pt_edge_error(const Iso& w2n, const double* X, const double* meas, double* e) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
    e[r] = (w2n.R[r * 3] * X[0] + w2n.R[r * 3 + 1] * X[1] + w2n.R[r * 3 + 2] * X[2] + w2n.t[r]) - meas[r];
}
__device__ __forceinline__ double pt_chi2(const double* e, const double* Om, double* Oe) {
#pragma unroll
  for (int r = 0; r < 3; ++r) Oe[r] = (Om[r * 3] * e[0] + Om[r * 3 + 1] * e[1]) + Om[r * 3 + 2] * e[2];
  return (e[0] * Oe[0] + e[1] * Oe[1]) + e[2] * Oe[2];
}

// LM scratch of the point landmarks ([field][match] blocks in the pair's slice)
struct PtView {
  double *X, *Xnew, *Hxx, *Hpx, *bx, *Inv, *contrib, *dx, *terms, *chi;   // 3,3,9,18,3,9,42,3,3,2
  size_t cap;    // contrib is stored [42][cap]
  double *J;     // 64 per match: Jl(newer) 9 | Jl(older) 9 | Jp 18 | e 2x3 | WOe 2x3 | r1 2 | pad
  int32_t* sel;  // [np] index into the pair's point match list
  int32_t* okf;
  const double* pmd;  // gathered data of all point matches of the pair
};
__device__ __forceinline__ const double* pt_meas(const double* pmd, int side, double* buf) {
  const float* f = reinterpret_cast<const float*>(pmd) + (side ? 4 : 0);
  buf[0] = (double)f[0]; buf[1] = (double)f[1]; buf[2] = (double)f[2];
  return buf;
}

__device__ void chi2_terms_pts(const PtView& P, int np, const Iso& w2n, const Iso& ident, const double* Xv, const PoseParams& PP) {
  for (int t = threadIdx.x; t < 2 * np; t += blockDim.x) {
    const int i = t >> 1, side = t & 1;
    const double* pmd = P.pmd + (size_t)P.sel[i] * PMD_STRIDE;
    double mb[3], e[3], Oe[3];
    pt_edge_error(side ? w2n : ident, Xv + 3 * i, pt_meas(pmd, side, mb), e);
    double c2 = pt_chi2(e, pmd + 4 + 9 * side, Oe);
    if (PP.robust) { double rho[3]; huber(c2, PP.huber_delta, rho); c2 = rho[0]; }
    P.chi[t] = c2;
  }
}

// getTransformFromHybridMatchesG2O restated for point + line landmarks (oracle/oracle_pair.cpp:refine_pose_hybrid).
// Six threads per landmark; ordered sums over the landmarks (points first, then lines) on dedicated threads.
__device__ void refine_pose_hybrid(const PtView& P, int np, const LmView& V, const double* md_all, int n, float* tf, int iterations,
                                   const PoseParams& PP, double* s_red, double* s_S, Iso* s_ci) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  if (n + np == 0) return;
  Iso tfd, cam1, ident;
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) tfd.R[r * 3 + c] = (double)tf[r * 4 + c]; tfd.t[r] = (double)tf[r * 4 + 3]; }
  iso_inv(tfd, cam1);
  for (int i = 0; i < 9; ++i) ident.R[i] = (i % 4 == 0) ? 1 : 0;
  ident.t[0] = ident.t[1] = ident.t[2] = 0;
  for (int t = tid; t < 6 * n; t += nthr) {
    const int i = t / 6, k = t - 6 * i;
    V.L[t] = md_all[(size_t)V.sel[i] * MD_STRIDE + k];
  }
  for (int t = tid; t < 3 * np; t += nthr) {
    const int i = t / 3, k = t - 3 * i;
    P.X[t] = (double)reinterpret_cast<const float*>(P.pmd + (size_t)P.sel[i] * PMD_STRIDE)[k];
  }
  __syncthreads();
  const double w = PP.line_weight_g2o;
  double lambda = 0, ni = 2;
  const double tau = 1e-5, lowS = 1. / 3., upS = 2. / 3.;
  const double del = 1e-9, scalar = 1 / (2 * del);
  for (int it = 0; it < iterations; ++it) {
    Iso w2n;
    iso_inv(cam1, w2n);
    chi2_terms_pts(P, np, w2n, ident, P.X, PP);
    chi2_terms(V, md_all, n, w2n, ident, V.L, PP);
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
    // ---- points: numeric Jacobian columns, thread (match i, column d)
    for (int t = tid; t < 6 * np; t += nthr) {
      const int i = t / 6, d = t - 6 * i;
      const double* pmd = P.pmd + (size_t)P.sel[i] * PMD_STRIDE;
      const double* Xi = P.X + 3 * i;
      double* Jm = P.J + (size_t)64 * i;
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        double mb[3];
        const double* meas = pt_meas(pmd, side, mb);
        if (d < 3) {
          double Xp[3] = {Xi[0], Xi[1], Xi[2]}, e1[3], e2[3];
          const double xplus = Xi[d] + del, xminus = Xi[d] + (-del);
#pragma unroll
          for (int k = 0; k < 3; ++k) if (k == d) Xp[k] = xplus;
          pt_edge_error(side ? w2n : ident, Xp, meas, e1);
#pragma unroll
          for (int k = 0; k < 3; ++k) if (k == d) Xp[k] = xminus;
          pt_edge_error(side ? w2n : ident, Xp, meas, e2);
#pragma unroll
          for (int k = 0; k < 3; ++k) Jm[side * 9 + k * 3 + d] = scalar * (e1[k] - e2[k]);
        }
        if (side) {
          double e1[3], e2[3];
          pt_edge_error(s_ci[2 * d], Xi, meas, e1);
          pt_edge_error(s_ci[2 * d + 1], Xi, meas, e2);
#pragma unroll
          for (int k = 0; k < 3; ++k) Jm[18 + k * 6 + d] = scalar * (e1[k] - e2[k]);
        }
        if (d == side) {  // residual, Omega e and robust weight of this edge
          double e[3], Oe[3];
          pt_edge_error(side ? w2n : ident, Xi, meas, e);
          double c2 = pt_chi2(e, pmd + 4 + 9 * side, Oe);
          double r1 = 1.0;
          if (PP.robust) { double rho[3]; huber(c2, PP.huber_delta, rho); r1 = rho[1]; }
#pragma unroll
          for (int k = 0; k < 3; ++k) { Jm[36 + 3 * side + k] = e[k]; Jm[42 + 3 * side + k] = r1 * Oe[k]; }
          Jm[48 + side] = r1;
        }
      }
    }
    // ---- lines: numeric Jacobian columns, thread (match i, column d)
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
        if (d == side) {
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
    // ---- points: block rows, thread (match i, row a)
    for (int t = tid; t < 6 * np; t += nthr) {
      const int i = t / 6, a = t - 6 * i;
      const double* pmd = P.pmd + (size_t)P.sel[i] * PMD_STRIDE;
      const double* Jm = P.J + (size_t)64 * i;
      double* cp = P.contrib + i; const size_t cs = P.cap;
      if (a < 3) {
        double b = 0, hrow[3] = {0, 0, 0};
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const double* Jl = Jm + 9 * side;
          const double* WOe = Jm + 42 + 3 * side;
          const double* Om = pmd + 4 + 9 * side;
          const double r1 = Jm[48 + side];
          double AtO[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            double s = 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) s += Jl[j * 3 + a] * (r1 * Om[j * 3 + k]);
            AtO[k] = s;
          }
          double s = 0;
#pragma unroll
          for (int k = 0; k < 3; ++k) s += Jl[k * 3 + a] * WOe[k];
          b -= s;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            double h = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) h += AtO[k] * Jl[k * 3 + c];
            hrow[c] += h;
          }
        }
        P.bx[3 * i + a] = b;
#pragma unroll
        for (int c = 0; c < 3; ++c) P.Hxx[9 * i + a * 3 + c] = hrow[c];
      }
      {  // older edge: pose blocks
        const double* Jl = Jm + 9;
        const double* Jp = Jm + 18;
        const double* WOe = Jm + 45;
        const double* Om = pmd + 13;
        const double r1 = Jm[49];
        double PtO[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double s = 0;
#pragma unroll
          for (int j = 0; j < 3; ++j) s += Jp[j * 6 + a] * (r1 * Om[j * 3 + k]);
          PtO[k] = s;
        }
        double sp = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) sp += Jp[k * 6 + a] * WOe[k];
        cp[(36 + a) * cs] = sp;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double h = 0;
#pragma unroll
          for (int k = 0; k < 3; ++k) h += PtO[k] * Jp[k * 6 + c];
          cp[(a * 6 + c) * cs] = h;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double g = 0;
#pragma unroll
          for (int k = 0; k < 3; ++k) g += PtO[k] * Jl[k * 3 + c];
          P.Hpx[18 * i + a * 3 + c] = 0.0 + g;
        }
      }
    }
    // ---- lines: block rows, thread (match i, row a)
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
    // ordered sums over the landmarks (points, then lines): Hpp (36), bp (6) on threads 0..41; chi2 on thread 64
    if (tid < 36) s_S[tid] = chain_sum<false>(chain_sum<false>(0.0, P.contrib + (size_t)tid * P.cap, 1, np), V.contrib + (size_t)tid * V.cap, 1, n);
    else if (tid < 42) s_S[tid] = chain_sum<true>(chain_sum<true>(0.0, P.contrib + (size_t)tid * P.cap, 1, np), V.contrib + (size_t)tid * V.cap, 1, n);
    else if (tid == 64) s_red[0] = chain_sum<false>(chain_sum<false>(0.0, P.chi, 1, 2 * np), V.chi, 1, 2 * n);
    __syncthreads();
    double Hpp[36], bp[6];
#pragma unroll
    for (int k = 0; k < 36; ++k) Hpp[k] = s_S[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) bp[k] = s_S[36 + k];
    double currentChi = s_red[0];
    __syncthreads();
    if (it == 0) {  // computeLambdaInit: tau * max |diagonal entry|
      double md_ = 0;
      for (int t = tid; t < 3 * np; t += nthr) { const int i = t / 3, a = t - 3 * i; md_ = fmax(fabs(P.Hxx[9 * i + a * 3 + a]), md_); }
      for (int t = tid; t < 6 * n; t += nthr) { const int i = t / 6, a = t - 6 * i; md_ = fmax(fabs(V.Hll[36 * i + a * 6 + a]), md_); }
      for (int o = 16; o; o >>= 1) md_ = fmax(md_, __shfl_xor_sync(FULL, md_, o));
      if ((tid & 31) == 0) s_red[1 + (tid >> 5)] = md_;
      __syncthreads();
      double maxDiag = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) maxDiag = fmax(fabs(Hpp[a * 6 + a]), maxDiag);
      for (int k = 0; k < nthr / 32; ++k) maxDiag = fmax(maxDiag, s_red[1 + k]);
      __syncthreads();
      lambda = tau * maxDiag;
      ni = 2;
    }
    double rho = 0;
    int qmax = 0;
    do {
      // landmark block inverses, one column per thread
      for (int t = tid; t < 3 * np; t += nthr) {
        const int i = t / 3, c = t - 3 * i;
        double Rc[3];
        int ok = invN_column<3>(P.Hxx + 9 * i, lambda, c, Rc);
        if (c == 0) P.okf[i] = ok;
#pragma unroll
        for (int k = 0; k < 3; ++k) P.Inv[9 * i + k * 3 + c] = Rc[k];
      }
      for (int t = tid; t < 6 * n; t += nthr) {
        const int i = t / 6, c = t - 6 * i;
        double Rc[6];
        int ok = inv6_column(V.Hll + 36 * i, lambda, c, Rc);
        if (c == 0) V.okf[i] = ok;
#pragma unroll
        for (int k = 0; k < 6; ++k) V.HllInv[36 * i + k * 6 + c] = Rc[k];
      }
      __syncthreads();
      // Schur terms: thread (landmark i, pose row a)
      for (int t = tid; t < 6 * np; t += nthr) {
        const int i = t / 6, a = t - 6 * i;
        const double* hi = P.Inv + 9 * i; const double* hpx = P.Hpx + 18 * i;
        double* cp = P.contrib + i; const size_t cs = P.cap;
        double T[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double s = 0;
#pragma unroll
          for (int k = 0; k < 3; ++k) s += hpx[a * 3 + k] * hi[k * 3 + c];
          T[c] = s;
        }
        double s = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) s += T[k] * P.bx[3 * i + k];
        cp[(36 + a) * cs] = s;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double h = 0;
#pragma unroll
          for (int k = 0; k < 3; ++k) h += T[k] * hpx[c * 3 + k];
          cp[(a * 6 + c) * cs] = h;
        }
      }
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
      if (tid < 42) {
        double s0 = tid < 36 ? Hpp[tid] : bp[tid - 36];
        if (tid < 36 && (tid / 6 == tid % 6)) s0 += lambda;
        s_S[tid] = chain_sum<true>(chain_sum<true>(s0, P.contrib + (size_t)tid * P.cap, 1, np), V.contrib + (size_t)tid * V.cap, 1, n);
      } else if (tid == 64) {
        int ok = 1;
        for (int i = 0; i < np; ++i) ok &= P.okf[i];
        for (int i = 0; i < n; ++i) ok &= V.okf[i];
        s_red[2] = (double)ok;
      }
      __syncthreads();
      if (tid < 6) {
        double Rc[6];
        int ok = inv6_column(s_S, 0.0, tid, Rc);
        if (tid == 0 && !ok) s_red[2] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s_S[48 + k * 6 + tid] = Rc[k];
      }
      __syncthreads();
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
      for (int t = tid; t < 3 * np; t += nthr) {
        const int i = t / 3, a = t - 3 * i;
        const double* hi = P.Inv + 9 * i; const double* hpx = P.Hpx + 18 * i;
        double s = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double q = 0;
#pragma unroll
          for (int m = 0; m < 6; ++m) q += hpx[m * 3 + k] * dp[m];
          s += hi[a * 3 + k] * (P.bx[3 * i + k] - q);
        }
        P.dx[t] = s;
        P.terms[t] = s * (lambda * s + P.bx[t]);
        P.Xnew[t] = P.X[t] + s;
      }
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
      chi2_terms_pts(P, np, w2nNew, ident, P.Xnew, PP);
      chi2_terms(V, md_all, n, w2nNew, ident, V.Lnew, PP);
      if (tid == nthr - 1) s_red[3] = chain_sum<false>(chain_sum<false>(scale, P.terms, 1, 3 * np), V.terms, 1, 6 * n);
      __syncthreads();
      if (tid == 64) s_red[4] = chain_sum<false>(chain_sum<false>(0.0, P.chi, 1, 2 * np), V.chi, 1, 2 * n);
      __syncthreads();
      scale = s_red[3];
      double tempChi = s_red[4];
      __syncthreads();
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
        for (int i = tid; i < 3 * np; i += nthr) P.X[i] = P.Xnew[i];
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

// ---------------------------------------------------------------------- scoring ----
__device__ __forceinline__ double score_point(const double* pmd, const float* tf, double sigma_depth) {
  const float* f = reinterpret_cast<const float*>(pmd);
  double tfd[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) tfd[k] = (double)tf[k];
  return error_function2(f, f + 4, tfd, sigma_depth);
}

// scores every match under tf; ordered inlier lists (points, lines) and the SSE in the reference's accumulation order
__device__ void score_all_hybrid(const double* pmd_all, int npm, const double* md_all, int nm, const float* tf, double thr,
                                 double sigma_depth, double* pd2, double* da_s, double* db_s, int32_t* psel, int32_t* sel,
                                 bool float_sse, int* cp_out, int* cl_out, double* sse_out, int* s_i, double* s_d) {
  const int tid = threadIdx.x;
  for (int i = tid; i < npm; i += blockDim.x) pd2[i] = score_point(pmd_all + (size_t)i * PMD_STRIDE, tf, sigma_depth);
  for (int i = tid; i < nm; i += blockDim.x) {
    double da, db;
    score_match(md_all + (size_t)i * MD_STRIDE, tf, &da, &db);
    da_s[i] = da; db_s[i] = db;
  }
  __syncthreads();
  if (tid == 0) {
    int cp = 0, cl = 0;
    float sf = 0; double sd = 0;
    const double thr2 = thr * thr;
    for (int i = 0; i < npm; ++i) {
      double d2 = pd2[i];
      if (d2 < thr2) { psel[cp++] = i; if (float_sse) sf += d2; else sd += d2; }
    }
    for (int i = 0; i < nm; ++i) {
      double da = da_s[i], db = db_s[i];
      if (da < thr && db < thr) {
        sel[cl++] = i;
        if (float_sse) sf += da * da + db * db; else sd += da * da + db * db;
      }
    }
    s_i[0] = cp; s_i[1] = cl;
    s_d[0] = float_sse ? (double)sf : sd;
  }
  __syncthreads();
  *cp_out = s_i[0]; *cl_out = s_i[1];
  *sse_out = s_d[0];
  __syncthreads();
}

#ifndef HYB_MINB
#define HYB_MINB 1
#endif
__global__ void __launch_bounds__(POSE_THREADS, HYB_MINB) pose_hybrid_kernel(const LslPairDesc* __restrict__ pairs, const LslPairPts* __restrict__ ppairs,
                                                                   const lsl_match* __restrict__ matches_all, const int32_t* __restrict__ nmatch,
                                                                   const lsl_match* __restrict__ pm_all, const int32_t* __restrict__ npmatch,
                                                                   LslPairScratch sc, LslHybScratch hs, PoseParams PP, HybParams HP,
                                                                   lsl_pose_rec* __restrict__ out) {
  __shared__ float s_tf[16];
  __shared__ double s_red[64];
  __shared__ double s_S[96];
  __shared__ Iso s_ci[12];
  __shared__ int s_i[4];
  __shared__ double s_d[2];
  __shared__ GRand s_rng;
  __shared__ uint16_t s_idx[LSL_MAX_MATCH + LSL_MAX_POINTS];
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = POSE_THREADS / 32;
  const LslPairDesc pd = pairs[pair];
  const LslPairPts pq = ppairs[pair];
  const int nm = min(nmatch[pair], pd.cap_m);
  const int npm = min(npmatch[pair], pq.cap_pm);
  const lsl_match* ms = matches_all + pd.m_off;
  const lsl_match* pms = pm_all + pq.pm_off;
  lsl_pose_rec* rec = out + pair;
  double* md_all = sc.md + pd.m_off * MD_STRIDE;
  double* da_s = sc.dab + pd.m_off * 2;
  double* db_s = da_s + pd.cap_m;
  int32_t* sel_r = sc.sel + pd.m_off * 3;
  int32_t* sel_f = sel_r + pd.cap_m;
  int32_t* sel_t = sel_f + pd.cap_m;
  double* pmd_all = hs.pmd + pq.pm_off * PMD_STRIDE;
  double* pd2 = hs.pd2 + pq.pm_off;
  int32_t* psel_r = hs.psel + pq.pm_off * 3;
  int32_t* psel_f = psel_r + pq.cap_pm;
  int32_t* psel_t = psel_f + pq.cap_pm;
  float* tfs = sc.tfs + (size_t)pair * sc.max_iter * 12;
  int32_t* cnts = sc.cnts + (size_t)pair * sc.max_iter;
  uint16_t* trip = sc.trip + (size_t)pair * sc.max_iter * 3;
  uint8_t* ptix = hs.ptidx + (size_t)pair * sc.max_iter * 2;
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
  PtView P;
  {
    double* lm = hs.plm + pq.pm_off * PLM_STRIDE;
    const size_t c = pq.cap_pm;
    P.X = lm; P.Xnew = P.X + 3 * c; P.Hxx = P.Xnew + 3 * c; P.Hpx = P.Hxx + 9 * c; P.bx = P.Hpx + 18 * c; P.Inv = P.bx + 3 * c;
    P.cap = c;
    P.contrib = P.Inv + 9 * c; P.dx = P.contrib + 42 * c; P.terms = P.dx + 3 * c; P.chi = P.terms + 3 * c; P.J = P.chi + 2 * c;
    P.okf = hs.pokf + pq.pm_off;
    P.sel = psel_r;
    P.pmd = pmd_all;
  }
  if (tid == 0) {
    rec->id_train = pd.id_t; rec->id_query = pd.id_q; rec->found = 0; rec->n_line_matches = nm;
    rec->n_ransac_inliers = 0; rec->n_inliers = 0; rec->rmse = 1e9f; rec->best_iter = -1;
    for (int i = 0; i < 16; ++i) rec->tf[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int i = 0; i < 8; ++i) rec->pad[i] = 0;
    rec->pad[0] = npm;
    sc.n_inl[pair] = 0; sc.n_rinl[pair] = 0; hs.n_pinl[pair] = 0; hs.n_prinl[pair] = 0;
  }
  const int nPt = npm, nLn = nm, line_weight = PP.line_weight;
  int min_inlier_nmb = PP.min_matches;
  if (nPt + nLn * line_weight < min_inlier_nmb) return;  // motion.cpp:621-624 (uniform)
  if (min_inlier_nmb > 0.7 * (nPt + nLn * line_weight)) min_inlier_nmb = (int)(0.7 * (nPt + nLn * line_weight));
  if (abs(pd.id_t - pd.id_q) > 50) min_inlier_nmb = PP.min_loopclose;
  const int maxIter = PP.max_iter;
  // ---- gather
  for (int i = tid; i < nm; i += blockDim.x) {
    const lsl_line_rec& q = pd.q[ms[i].queryIdx];
    const lsl_line_rec& t = pd.t[ms[i].trainIdx];
    double* md = md_all + (size_t)i * MD_STRIDE;
    for (int k = 0; k < 3; ++k) { md[k] = q.A[k]; md[3 + k] = q.B[k]; md[6 + k] = t.A[k]; md[9 + k] = t.B[k]; }
    for (int k = 0; k < 9; ++k) { md[12 + k] = t.DU_A[k]; md[21 + k] = t.DU_B[k]; }
    affn(q.covA, md + 36); affn(q.covB, md + 45); affn(t.covA, md + 54); affn(t.covB, md + 63);
  }
  for (int i = tid; i < npm; i += blockDim.x) {
    const float* qx = pq.qx + 4 * (size_t)pms[i].queryIdx;
    const float* tx = pq.tx + 4 * (size_t)pms[i].trainIdx;
    double* pmd = pmd_all + (size_t)i * PMD_STRIDE;
    float* f = reinterpret_cast<float*>(pmd);
    for (int k = 0; k < 4; ++k) { f[k] = qx[k]; f[4 + k] = tx[k]; }
    pt_info_f(qx, HP.fx, HP.sigma_impt, HP.c1, HP.c2, HP.c3, HP.dt, pmd + 4);
    pt_info_f(tx, HP.fx, HP.sigma_impt, HP.c1, HP.c2, HP.c3, HP.dt, pmd + 13);
  }
  // ---- samples: one rand() stream continuing after featureMatching's draws (motion.cpp:635-658, 544)
  if (tid == 0) {
    const int32_t* st = hs.rng + (size_t)pair * 33;
    for (int k = 0; k < 31; ++k) s_rng.r[k] = st[k];
    s_rng.f = st[31]; s_rng.b = st[32];
    uint16_t* idx = s_idx;
    const int tot = nPt + nLn;
    for (int i = 0; i < tot; ++i) idx[i] = (uint16_t)i;
    for (int it = 0; it < maxIter; ++it) {
      int left = tot;
      for (int k = 0; k < 3; ++k) {
        int r = grand_next(&s_rng) % left;
        uint16_t t = idx[k]; idx[k] = idx[k + r]; idx[k + r] = t;
        --left;
      }
      int nl = 0;
      for (int k = 0; k < 3; ++k) { trip[3 * it + k] = idx[k]; nl += idx[k] >= nPt; }
      ptix[2 * it] = 0; ptix[2 * it + 1] = 0;
      if (nl < 3) {
        const int nps = 3 - nl;
        for (int k = 0; k < nl; ++k) ptix[2 * it + k] = (uint8_t)(grand_next(&s_rng) % nps);
      }
    }
  }
  __syncthreads();
  // ---- minimal solutions
  for (int h = tid; h < maxIter; h += blockDim.x) {
    int pti[3], lni[3], nps = 0, nls = 0;
    for (int k = 0; k < 3; ++k) { int v = trip[3 * h + k]; if (v < nPt) pti[nps++] = v; else lni[nls++] = v - nPt; }
    float* tfo = tfs + (size_t)h * 12;
    bool valid = true;
    if (nls == 3) {
      const double* mdp[3] = {md_all + (size_t)lni[0] * MD_STRIDE, md_all + (size_t)lni[1] * MD_STRIDE, md_all + (size_t)lni[2] * MD_STRIDE};
      double R[9], t[3];
      relmotion_svd3(mdp, R, t);
      for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) tfo[r * 4 + c] = (float)R[r * 3 + c]; tfo[r * 4 + 3] = (float)t[r]; }
    } else {  // getTransform_Lns_Pts_pcl
      Tfc tfc; tfc_reset(&tfc);
      for (int i = 0; i < nls; ++i) {
        const float* pf = reinterpret_cast<const float*>(pmd_all + (size_t)pti[ptix[2 * h + i]] * PMD_STRIDE);
        const double* md = md_all + (size_t)lni[i] * MD_STRIDE;
        double query_pt[3] = {(double)pf[0], (double)pf[1], (double)pf[2]}, train_pt[3] = {(double)pf[4], (double)pf[5], (double)pf[6]};
        double train_prj[3], query_prj[3];
        project_pt_ln(train_pt, md + 6, md + 9, train_prj);
        project_pt_ln(query_pt, md, md + 3, query_prj);
        float from[3] = {(float)query_prj[0], (float)query_prj[1], (float)query_prj[2]}, to[3] = {(float)train_prj[0], (float)train_prj[1], (float)train_prj[2]};
        if (from[2] != from[2] || to[2] != to[2]) continue;
        float weight = 1 / (fabsf(to[2]) + fabsf(from[2]));
        tfc_add(&tfc, from, to, weight);
      }
      for (int i = 0; i < nps; ++i) {
        const float* pf = reinterpret_cast<const float*>(pmd_all + (size_t)pti[i] * PMD_STRIDE);
        const float* from = pf; const float* to = pf + 4;
        if (from[2] != from[2] || to[2] != to[2]) continue;
        float weight = 1 / (fabsf(to[2]) + fabsf(from[2]));
        tfc_add(&tfc, from, to, weight);
      }
      valid = tfc.n >= 3;
      float tf16[16];
      tfc_get(&tfc, tf16);
      for (int k = 0; k < 12; ++k) tfo[k] = tf16[k];
    }
    cnts[h] = valid ? 0 : -1;
  }
  __syncthreads();
  // ---- scoring: warp per hypothesis, lanes over matches
  for (int h = warp; h < maxIter; h += nwarp) {
    if (cnts[h] < 0) continue;   // invalid sample: `continue` in the reference
    const float* tfh = tfs + (size_t)h * 12;
    float tf[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) tf[k] = tfh[k];
    int c = 0;
    const double thr2 = PP.thr * PP.thr;
    for (int i0 = 0; i0 < npm; i0 += 32) {
      int i = i0 + lane;
      bool in = false;
      if (i < npm) in = score_point(pmd_all + (size_t)i * PMD_STRIDE, tf, HP.sigma_depth) < thr2;
      c += __popc(__ballot_sync(FULL, in));
    }
    for (int i0 = 0; i0 < nm; i0 += 32) {
      int i = i0 + lane;
      bool in = false;
      if (i < nm) {
        double da, db;
        score_match(md_all + (size_t)i * MD_STRIDE, tf, &da, &db);
        in = da < PP.thr && db < PP.thr;
      }
      c += line_weight * __popc(__ballot_sync(FULL, in));
    }
    __syncwarp();
    if (lane == 0) cnts[h] = c;
  }
  __syncthreads();
  // ---- first best (strict >, motion.cpp:714-720)
  if (warp == 0) {
    int best = 0, bh = 1 << 30;
    for (int h = lane; h < maxIter; h += 32) {
      int c = cnts[h];
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
  int bp_cnt, bl_cnt;
  score_all_hybrid(pmd_all, npm, md_all, nm, s_tf, PP.thr, HP.sigma_depth, pd2, da_s, db_s, psel_r, sel_r, true, &bp_cnt, &bl_cnt, &sse, s_i, s_d);
  if (tid == 0) rec->best_iter = bh;
  if (bp_cnt + bl_cnt < 3) return;  // motion.cpp:722-725
  if (tid == 0) {
    rec->n_ransac_inliers = bl_cnt; rec->pad[1] = bp_cnt;
    sc.n_rinl[pair] = bl_cnt; hs.n_prinl[pair] = bp_cnt;
    float* tr = sc.tf_ransac + (size_t)pair * 16;
    for (int i = 0; i < 16; ++i) tr[i] = s_tf[i];
  }
  const float sum_squared_error = (float)sse;
  // ---- refinement (motion.cpp:726-839)
  V.sel = sel_r; P.sel = psel_r;
  refine_pose_hybrid(P, bp_cnt, V, md_all, bl_cnt, s_tf, 25, PP, s_red, s_S, s_ci);
  double refined_rmse = (double)sqrtf(sum_squared_error / (float)(bp_cnt + bl_cnt));
  int rp_cnt = 0, rl_cnt = 0;
  for (int it = 0; it < 20; ++it) {
    double tmp_sse;
    int cp, cl;
    score_all_hybrid(pmd_all, npm, md_all, nm, s_tf, PP.thr, HP.sigma_depth, pd2, da_s, db_s, psel_t, sel_t, false, &cp, &cl, &tmp_sse, s_i, s_d);
    if (cp + cl * line_weight > rp_cnt + rl_cnt * line_weight) {
      for (int i = tid; i < cp; i += blockDim.x) psel_f[i] = psel_t[i];
      for (int i = tid; i < cl; i += blockDim.x) sel_f[i] = sel_t[i];
      rp_cnt = cp; rl_cnt = cl;
      refined_rmse = sqrt(tmp_sse / (double)(cp + cl));
      __syncthreads();
      V.sel = sel_f; P.sel = psel_f;
      refine_pose_hybrid(P, rp_cnt, V, md_all, rl_cnt, s_tf, 20, PP, s_red, s_S, s_ci);
    } else break;
  }
  __syncthreads();
  if (tid == 0) {
    rec->n_inliers = rl_cnt; rec->pad[2] = rp_cnt;
    rec->rmse = (float)refined_rmse;
    for (int i = 0; i < 12; ++i) rec->tf[i] = s_tf[i];
    rec->found = (rp_cnt + line_weight * rl_cnt) >= min_inlier_nmb ? 1 : 0;
    sc.n_inl[pair] = rl_cnt; hs.n_pinl[pair] = rp_cnt;
  }
}

// ------------------------------------------------------------- launchers ----
int lsl_launch_match_points(lsl_ctx* ctx, int npairs, int max_nq, int dim, int kind) {
  LslHybWork& h = ctx->hw;
  LSL_KSTART(ctx, LSL_K_MATCHPTS);
  if (max_nq > 0) {
    dim3 g((max_nq + 7) / 8, npairs);
    int tc_used = 0;
    if (kind == 1) match_points_hamming_kernel<<<g, 256, 8 * dim, ctx->stream>>>(h.d_ppairs, (Knn2*)h.knn);
    else {
      // f32 rows: the distance matrix goes through the tensor cores (tcgen05 tf32 pre-filter + exact re-evaluation of the
      // candidates, k_match_tc.cu) unless LSL_MATCH_TC=0 or the row length does not fit that path
      static int want_tc = -1;
      if (want_tc < 0) { const char* e = getenv("LSL_MATCH_TC"); want_tc = (e && e[0] == '0') ? 0 : 1; }
      if (want_tc) { int rc = lsl_launch_match_points_tc(ctx, npairs, max_nq, dim, &tc_used); if (rc) return rc; }
      if (!tc_used) match_points_kernel<<<g, 256, 8 * dim * sizeof(float), ctx->stream>>>(h.d_ppairs, (Knn2*)h.knn);
    }
    if (!tc_used) ctx->stats.kernel_launches += 1;
  }
  match_points_accept_kernel<<<npairs, 32, 0, ctx->stream>>>(h.d_ppairs, ctx->pw.d_pairs, (const Knn2*)h.knn, h.pmatches, h.npmatch,
                                                            h.hs.rng, ctx->P.nn_distance_ratio);
  LSL_KSTOP(ctx, LSL_K_MATCHPTS);
  LSL_CUDA(cudaGetLastError());
  return LSL_OK;
}

// Node::computeInliersAndError (src/node.cpp:1019-1080): squared Mahalanobis distance (errorFunction2) of every point match under
// `tf`, matches with a zero depth on either side skipped, outliers (> squared_max) and non-finite values dropped; the inlier
// list and the sum keep the match order (thread 0 walks the flags: the sum is one ordered chain of <= 2048 adds).
__global__ void __launch_bounds__(256) inliers_error_kernel(const float* __restrict__ qx, const float* __restrict__ tx, const lsl_match* __restrict__ ms,
                                                            int n, const float* __restrict__ tf16, double sigma_depth, double squared_max,
                                                            double* __restrict__ dist, lsl_match* __restrict__ out, int32_t* __restrict__ n_out,
                                                            double* __restrict__ rmse) {
  __shared__ double s_tf[16];
  if (threadIdx.x < 16) s_tf[threadIdx.x] = (double)tf16[threadIdx.x];      // transformation4f.cast<double>()
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float* o = qx + 4 * (size_t)ms[i].queryIdx;
    const float* t = tx + 4 * (size_t)ms[i].trainIdx;
    double d = -1.0;                                                          // -1: skipped
    if (!(o[2] == 0.0f || t[2] == 0.0f)) {                                    // does NOT trigger on NaN (node.cpp:1046)
      const double m = error_function2(o, t, s_tf, sigma_depth);
      if (!(m > squared_max) && m >= 0.0) d = m;
    }
    dist[i] = d;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double mean = 0.0;
    int k = 0;
    for (int i = 0; i < n; ++i)
      if (dist[i] >= 0.0) { mean += dist[i]; out[k++] = ms[i]; }
    *n_out = k;
    *rmse = k < 3 ? 1e9 : sqrt(mean / (double)k);
  }
}
int lsl_launch_inliers_error(lsl_ctx* ctx, const float* qx, const float* tx, const lsl_match* d_ms, int n, const float* d_tf, double squared_max,
                             double* d_dist, lsl_match* d_out, int32_t* d_n, double* d_rmse) {
  inliers_error_kernel<<<1, 256, 0, ctx->stream>>>(qx, tx, d_ms, n, d_tf, ctx->P.sigma_depth, squared_max, d_dist, d_out, d_n, d_rmse);
  ctx->stats.kernel_launches += 1;
  LSL_CUDA(cudaGetLastError());
  return LSL_OK;
}

int lsl_launch_rootsift(lsl_ctx* ctx, float* d_desc, int n, int dim) {
  if (n <= 0) return LSL_OK;
  const int rpc = dim <= 256 ? 32 : 8;              // <= 33 KB of shared memory either way (dim <= 512)
  rootsift_kernel<<<(n + rpc - 1) / rpc, 128, (size_t)rpc * (dim + 1) * sizeof(float), ctx->stream>>>(d_desc, n, dim, rpc);
  ctx->stats.kernel_launches += 1;
  LSL_CUDA(cudaGetLastError());
  return LSL_OK;
}

int lsl_launch_pose_hybrid(lsl_ctx* ctx, int npairs, double fx, double dt) {
  const lsl_params& P = ctx->P;
  PoseParams PP;
  PP.thr = P.max_mah_dist_for_inliers; PP.line_weight_g2o = P.g2o_line_error_weight; PP.huber_delta = P.g2o_BA_kernel_delta;
  PP.robust = P.g2o_BA_use_kernel; PP.max_iter = P.ransac_iters_line_motion; PP.min_matches = P.min_feature_matches;
  PP.min_loopclose = P.min_matches_loopclose; PP.line_weight = P.line_match_number_weight;
  HybParams HP;
  HP.sigma_depth = P.sigma_depth; HP.nn_ratio = P.nn_distance_ratio; HP.fx = fx; HP.sigma_impt = P.stdev_sample_pt_imgline;
  HP.c1 = P.depth_stdev_coeff_c1; HP.c2 = P.depth_stdev_coeff_c2; HP.c3 = P.depth_stdev_coeff_c3; HP.dt = dt;
  LSL_KSTART(ctx, LSL_K_POSEHYB);
  pose_hybrid_kernel<<<npairs, POSE_THREADS, 0, ctx->stream>>>(ctx->pw.d_pairs, ctx->hw.d_ppairs, ctx->pw.matches, ctx->pw.nmatch,
                                                               ctx->hw.pmatches, ctx->hw.npmatch, ctx->pw.sc, ctx->hw.hs, PP, HP, ctx->pw.recs);
  LSL_KSTOP(ctx, LSL_K_POSEHYB);
  LSL_CUDA(cudaGetLastError());
  return LSL_OK;
}

// =====================================================================================================
// computeRelativeMotion_Ransac (src/line/motion.cpp:367-526): line-only RANSAC with Euclidean consensus and the
// levmar refinement optimizeRelmotion (motion.cpp:98-139; dlevmar_dif m = 7, cost motion.cpp:60-96).
// One CTA per problem. The LM follows oracle/oracle_extract.cpp:dlevmar_dif_restated (pinned against the
// reference's levmar build): residuals / Jacobian rows over the threads, ordered sums on dedicated threads,
// the 7x7 Crout LU (Axb_core.c:1140-1277) and the 4-accumulator norm (misc_core.c:721-809) on thread 0.
#define RM_THREADS 128


__device__ __forceinline__ double dist3d_pt_line_d(const double* X, const double* A, const double* B) {  // utils.cpp:626-636
  double AB[3] = {A[0] - B[0], A[1] - B[1], A[2] - B[2]};
  double nab = sqrt(AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2]);
  if (nab < 1e-10) return -1;
  double XA[3] = {X[0] - A[0], X[1] - A[1], X[2] - A[2]};
  double ax = sqrt(XA[0] * XA[0] + XA[1] * XA[1] + XA[2] * XA[2]);
  double inv = 1 / nab;
  double nv[3] = {(B[0] - A[0]) * inv, (B[1] - A[1]) * inv, (B[2] - A[2]) * inv};
  double d = XA[0] * nv[0] + XA[1] * nv[1] + XA[2] * nv[2];
  return sqrt(fabs(ax * ax - d * d));
}
__device__ __forceinline__ bool rm_consistent(const double* g, const double* R, const double* t, double distThresh, double angThresh) {
  const double PI_T = 3.14159265;
  const double *aAp = g, *aBp = g + 3, *bAp = g + 6, *bBp = g + 9;
  double aA[3], aB[3], RaAB[3];
  double aAB[3] = {aAp[0] - aBp[0], aAp[1] - aBp[1], aAp[2] - aBp[2]};
  double bAB[3] = {bAp[0] - bBp[0], bAp[1] - bBp[1], bAp[2] - bBp[2]};
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    aA[r] = ((R[r * 3] * aAp[0] + R[r * 3 + 1] * aAp[1]) + R[r * 3 + 2] * aAp[2]) + t[r];
    aB[r] = ((R[r * 3] * aBp[0] + R[r * 3 + 1] * aBp[1]) + R[r * 3 + 2] * aBp[2]) + t[r];
    RaAB[r] = (R[r * 3] * aAB[0] + R[r * 3 + 1] * aAB[1]) + R[r * 3 + 2] * aAB[2];
  }
  double dist = 0.5 * dist3d_pt_line_d(aA, bAp, bBp) + 0.5 * dist3d_pt_line_d(aB, bAp, bBp);
  double dt = (RaAB[0] * bAB[0] + RaAB[1] * bAB[1]) + RaAB[2] * bAB[2];
  double na = sqrt(aAB[0] * aAB[0] + aAB[1] * aAB[1] + aAB[2] * aAB[2]);
  double nb = sqrt(bAB[0] * bAB[0] + bAB[1] * bAB[1] + bAB[2] * bAB[2]);
  double angle = 180 * lsl_acos(fabs(dt / na / nb)) / PI_T;
  return dist < distThresh && angle < angThresh;
}
// costFun_optimizeRelmotion for one pair (motion.cpp:60-96)
__device__ double rm_cost_one(const double* g, const double* R, const double* t) {
  const double *aAp = g, *aBp = g + 3, *bAp = g + 6, *bBp = g + 9;
  double aA[3], aB[3], bA[3], bB[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    aA[r] = (R[r * 3] * aAp[0] + R[r * 3 + 1] * aAp[1] + R[r * 3 + 2] * aAp[2]) + t[r];
    aB[r] = (R[r * 3] * aBp[0] + R[r * 3 + 1] * aBp[1] + R[r * 3 + 2] * aBp[2]) + t[r];
  }
  double dA[3] = {bAp[0] - t[0], bAp[1] - t[1], bAp[2] - t[2]}, dB[3] = {bBp[0] - t[0], bBp[1] - t[1], bBp[2] - t[2]};
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    bA[r] = R[r] * dA[0] + R[3 + r] * dA[1] + R[6 + r] * dA[2];
    bB[r] = R[r] * dB[0] + R[3 + r] * dB[1] + R[6 + r] * dB[2];
  }
  return 0.25 * (mah_dist3d_pt_line(bAp, g + 30, aA, aB) + mah_dist3d_pt_line(bBp, g + 39, aA, aB) +
                 mah_dist3d_pt_line(aAp, g + 12, bA, bB) + mah_dist3d_pt_line(aBp, g + 21, bA, bB));
}
__device__ void rm_cost(const double* g_all, const int32_t* sub, int n, const double* p, double* out) {
  double R[9];
  q2r(p, R);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = rm_cost_one(g_all + (size_t)sub[i] * RM_STRIDE, R, p + 4);
}
// LEVMAR_L2NRMXMY with x = 0 (misc_core.c:721-809), serial
__device__ double rm_l2nrm(double* e, const double* y, int n) {
  double sum0 = 0.0, sum1 = 0.0, sum2 = 0.0, sum3 = 0.0;
  const int blockn = (n >> 3) << 3;
#define EI(j) (e[j] = 0.0 - y[j], e[j] * e[j])
  for (int i = blockn - 1; i > 0; i -= 8) {
    sum0 += EI(i); sum1 += EI(i - 1); sum2 += EI(i - 2); sum3 += EI(i - 3);
    sum0 += EI(i - 4); sum1 += EI(i - 5); sum2 += EI(i - 6); sum3 += EI(i - 7);
  }
  int i = blockn;
  if (i < n) {
    switch (n - i) {
      case 7: sum0 += EI(i); ++i;
      case 6: sum1 += EI(i); ++i;
      case 5: sum2 += EI(i); ++i;
      case 4: sum3 += EI(i); ++i;
      case 3: sum0 += EI(i); ++i;
      case 2: sum1 += EI(i); ++i;
      case 1: sum2 += EI(i);
    }
  }
#undef EI
  return sum0 + sum1 + sum2 + sum3;
}
// AX_EQ_B_LU (Axb_core.c:1140-1277) for m = 7, serial; A is JtJ with mu already on the diagonal
__device__ int rm_lu7(const double* A, const double* B, double* x) {
  const int m = RM_M;
  double a[RM_M * RM_M], work[RM_M];
  int idx[RM_M];
  int maxi = -1;
  for (int i = 0; i < m * m; ++i) a[i] = A[i];
  for (int i = 0; i < m; ++i) x[i] = B[i];
  for (int i = 0; i < m; ++i) {
    double max = 0.0, tmp;
    for (int j = 0; j < m; ++j)
      if ((tmp = fabs(a[i * m + j])) > max) max = tmp;
    if (max == 0.0) return 0;
    work[i] = 1.0 / max;
  }
  for (int j = 0; j < m; ++j) {
    for (int i = 0; i < j; ++i) {
      double sum = a[i * m + j];
      for (int k = 0; k < i; ++k) sum -= a[i * m + k] * a[k * m + j];
      a[i * m + j] = sum;
    }
    double max = 0.0, tmp;
    for (int i = j; i < m; ++i) {
      double sum = a[i * m + j];
      for (int k = 0; k < j; ++k) sum -= a[i * m + k] * a[k * m + j];
      a[i * m + j] = sum;
      if ((tmp = work[i] * fabs(sum)) >= max) { max = tmp; maxi = i; }
    }
    if (j != maxi) {
      for (int k = 0; k < m; ++k) { double t = a[maxi * m + k]; a[maxi * m + k] = a[j * m + k]; a[j * m + k] = t; }
      work[maxi] = work[j];
    }
    idx[j] = maxi;
    if (a[j * m + j] == 0.0) a[j * m + j] = DBL_EPSILON;
    if (j != m - 1) {
      double tmp2 = 1.0 / (a[j * m + j]);
      for (int i = j + 1; i < m; ++i) a[i * m + j] *= tmp2;
    }
  }
  int k = 0;
  for (int i = 0; i < m; ++i) {
    int j = idx[i];
    double sum = x[j];
    x[j] = x[i];
    if (k != 0)
      for (j = k - 1; j < i; ++j) sum -= a[i * m + j] * x[j];
    else if (sum != 0.0) k = i + 1;
    x[i] = sum;
  }
  for (int i = m - 1; i >= 0; --i) {
    double sum = x[i];
    for (int j = i + 1; j < m; ++j) sum -= a[i * m + j] * x[j];
    x[i] = sum / a[i * m + i];
  }
  return 1;
}

struct RmLmShared {
  double p[RM_M], pDp[RM_M], Dp[RM_M], jacTe[RM_M], diag[RM_M], JtJ[RM_M * RM_M], ptmp[RM_M];
  double p_eL2, pDp_eL2, Dp_L2, fd;
  int issolved;
};
// optimizeRelmotion: dlevmar_dif(m = 7, n, itmax 50, opts {1e-3, 1e-10, 1e-20, 1e-20, 1e-6}) on the subset `sub`;
// R, t (shared memory) in / out. Whole CTA.
__device__ void rm_optimize(const double* g_all, const int32_t* sub, int n, double* R, double* t, double* hx, double* wrk, double* e,
                            double* wrk2, double* jac, RmLmShared& S) {
  const int tid = threadIdx.x, m = RM_M;
  if (tid == 0) {  // r2q (utils.cpp:1696-1707)
    double tr = R[0] + R[4] + R[8];
    double r = sqrt(1 + tr);
    double s = 0.5 / r;
    S.p[0] = 0.5 * r; S.p[1] = (R[7] - R[5]) * s; S.p[2] = (R[2] - R[6]) * s; S.p[3] = (R[3] - R[1]) * s;
    S.p[4] = t[0]; S.p[5] = t[1]; S.p[6] = t[2];
  }
  __syncthreads();
  const double tau = 1E-03, eps1 = 1E-10, eps2 = 1E-20, eps2_sq = 1E-20 * 1E-20, eps3 = 1E-20, delta = 1E-06;
  const int itmax = 50;
  if (n >= m) {
    double mu = 0, jacTe_inf = 0, p_L2 = 0, tmp, p_eL2, dF, dL;
    int nu = 20, nu2, stop = 0, K = 10, updjac = 0, updp = 1, newjac = 0, k;
    rm_cost(g_all, sub, n, S.p, hx);
    __syncthreads();
    if (tid == 0) S.p_eL2 = rm_l2nrm(e, hx, n);
    __syncthreads();
    p_eL2 = S.p_eL2;
    if (!isfinite(p_eL2)) stop = 7;
    for (k = 0; k < itmax && !stop; ++k) {
      if (p_eL2 <= eps3) { stop = 6; break; }
      if ((updp && nu > 16) || updjac == K) {
        for (int j = 0; j < m; ++j) {  // LEVMAR_FDIF_FORW_JAC_APPROX (misc_core.c:137-172)
          if (tid == 0) {
            double d = 1E-04 * S.p[j];
            d = fabs(d);
            if (d < delta) d = delta;
            for (int c = 0; c < m; ++c) S.ptmp[c] = S.p[c];
            S.ptmp[j] = S.p[j] + d;
            S.fd = 1.0 / d;
          }
          __syncthreads();
          rm_cost(g_all, sub, n, S.ptmp, wrk);
          __syncthreads();
          const double dinv = S.fd;
          for (int i = tid; i < n; i += blockDim.x) jac[(size_t)i * m + j] = (wrk[i] - hx[i]) * dinv;
          __syncthreads();
        }
        nu = 2; updjac = 0; updp = 0; newjac = 1;
      }
      if (newjac) {
        newjac = 0;
        // J^T J (lower triangle) and J^T e, accumulated for l = n-1 .. 0 (lm_core.c:618-639): one accumulator per thread
        if (tid < 35) {
          double acc = 0.0;
          if (tid < 28) {
            int i = 0, r = tid;
            while (r > i) { r -= i + 1; ++i; }  // tid -> (i, j = r), j <= i
            for (int l = n - 1; l >= 0; --l) acc += jac[(size_t)l * m + r] * jac[(size_t)l * m + i];
            S.JtJ[i * m + r] = acc; S.JtJ[r * m + i] = acc;
          } else {
            const int i = tid - 28;
            for (int l = n - 1; l >= 0; --l) acc += jac[(size_t)l * m + i] * e[l];
            S.jacTe[i] = acc;
          }
        }
        __syncthreads();
        p_L2 = jacTe_inf = 0.0;
        for (int i = 0; i < m; ++i) {
          if (jacTe_inf < (tmp = fabs(S.jacTe[i]))) jacTe_inf = tmp;
          p_L2 += S.p[i] * S.p[i];
        }
        if (tid < m) S.diag[tid] = S.JtJ[tid * m + tid];
        __syncthreads();
      }
      if (jacTe_inf <= eps1) { stop = 1; break; }
      if (k == 0) {
        tmp = DBL_MIN;
        for (int i = 0; i < m; ++i)
          if (S.diag[i] > tmp) tmp = S.diag[i];
        mu = tau * tmp;
      }
      __syncthreads();
      if (tid == 0) {
        for (int i = 0; i < m; ++i) S.JtJ[i * m + i] += mu;
        S.issolved = rm_lu7(S.JtJ, S.jacTe, S.Dp);
        if (S.issolved) {
          double d2 = 0.0;
          for (int i = 0; i < m; ++i) { double v = S.Dp[i]; S.pDp[i] = S.p[i] + v; d2 += v * v; }
          S.Dp_L2 = d2;
        }
      }
      __syncthreads();
      if (S.issolved) {
        const double Dp_L2 = S.Dp_L2;
        if (Dp_L2 <= eps2_sq * p_L2) { stop = 2; break; }
        if (Dp_L2 >= (p_L2 + eps2) / (1E-12 * 1E-12)) { stop = 4; break; }
        rm_cost(g_all, sub, n, S.pDp, wrk);
        __syncthreads();
        if (tid == 0) S.pDp_eL2 = rm_l2nrm(wrk2, wrk, n);
        __syncthreads();
        const double pDp_eL2 = S.pDp_eL2;
        if (!isfinite(pDp_eL2)) { stop = 7; break; }
        dF = p_eL2 - pDp_eL2;
        if (updp || dF > 0) {  // Broyden rank-one update
          for (int i = tid; i < n; i += blockDim.x) {
            double tt = 0.0;
            for (int l = 0; l < m; ++l) tt += jac[(size_t)i * m + l] * S.Dp[l];
            tt = (wrk[i] - hx[i] - tt) / Dp_L2;
            for (int j = 0; j < m; ++j) jac[(size_t)i * m + j] += tt * S.Dp[j];
          }
          ++updjac;
          newjac = 1;
        }
        dL = 0.0;
        for (int i = 0; i < m; ++i) dL += S.Dp[i] * (mu * S.Dp[i] + S.jacTe[i]);
        if (dL > 0.0 && dF > 0.0) {
          tmp = (2.0 * dF / dL - 1.0);
          tmp = 1.0 - tmp * tmp * tmp;
          mu = mu * ((tmp >= 0.3333333334) ? tmp : 0.3333333334);
          nu = 2;
          __syncthreads();
          if (tid < m) S.p[tid] = S.pDp[tid];
          for (int i = tid; i < n; i += blockDim.x) { e[i] = wrk2[i]; hx[i] = wrk[i]; }
          p_eL2 = pDp_eL2;
          updp = 1;
          __syncthreads();
          continue;
        }
      }
      mu *= nu;
      nu2 = nu << 1;
      if (nu2 <= nu) { stop = 5; break; }
      nu = nu2;
      __syncthreads();
      if (tid < m) S.JtJ[tid * m + tid] = S.diag[tid];
      __syncthreads();
    }
  }
  __syncthreads();
  if (tid == 0) {
    q2r(S.p, R);
    t[0] = S.p[4]; t[1] = S.p[5]; t[2] = S.p[6];
  }
  __syncthreads();
}

// flags -> ordered index list (thread 0), returns the count
__device__ int rm_compact(const int32_t* flag, int n, int32_t* out, int* s_cnt) {
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int i = 0; i < n; ++i) if (flag[i]) out[c++] = i;
    *s_cnt = c;
  }
  __syncthreads();
  int c = *s_cnt;
  __syncthreads();
  return c;
}

#ifdef RM_MINB
#define RM_BOUNDS __launch_bounds__(RM_THREADS, RM_MINB)
#else
#define RM_BOUNDS __launch_bounds__(RM_THREADS)
#endif
__global__ void RM_BOUNDS relmotion_kernel(const LslPairDesc* __restrict__ pairs, const lsl_match* __restrict__ matches_all,
                                                               const int32_t* __restrict__ nmatch, RmScratch rs, double distThresh,
                                                               double angThresh, double cosDeg, int maxIters) {
  __shared__ RmLmShared S;
  __shared__ double s_R[9], s_t[3], s_Ro[9], s_to[3];
  __shared__ int s_cnt, s_best;
  __shared__ GRand s_rng;
  __shared__ uint16_t s_idx[LSL_MAX_MATCH];
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = RM_THREADS / 32;
  const LslPairDesc pd = pairs[pair];
  const int n = min(nmatch[pair], pd.cap_m);
  const lsl_match* ms = matches_all + pd.m_off;
  double* g_all = rs.g + pd.m_off * RM_STRIDE;
  double* hx = rs.hx + pd.m_off * 4;
  double* wrk = hx + pd.cap_m; double* e = wrk + pd.cap_m; double* wrk2 = e + pd.cap_m;
  double* jac = rs.jac + pd.m_off * RM_M;
  int32_t* flag = rs.flag + pd.m_off;
  int32_t* cur = rs.cur + pd.m_off * 2;
  int32_t* trial = cur + pd.cap_m;
  double* hyp = rs.hyp + (size_t)pair * rs.max_iter * 12;
  int32_t* cnts = rs.cnts + (size_t)pair * rs.max_iter;
  uint16_t* trip = rs.trip + (size_t)pair * rs.max_iter * 3;
  double* outRt = rs.outRt + (size_t)pair * 12;
  int32_t* outn = rs.outn + (size_t)pair * 4;
  if (tid == 0) { outn[0] = 0; outn[1] = 0; outn[2] = 0; outn[3] = 0; }
  if (n < 3) return;
  for (int i = tid; i < n; i += blockDim.x) {
    const lsl_line_rec& a = pd.q[ms[i].queryIdx];
    const lsl_line_rec& b = pd.t[ms[i].trainIdx];
    double* g = g_all + (size_t)i * RM_STRIDE;
    for (int k = 0; k < 3; ++k) { g[k] = a.A[k]; g[3 + k] = a.B[k]; g[6 + k] = b.A[k]; g[9 + k] = b.B[k]; }
    for (int k = 0; k < 9; ++k) { g[12 + k] = a.DU_A[k]; g[21 + k] = a.DU_B[k]; g[30 + k] = b.DU_A[k]; g[39 + k] = b.DU_B[k]; }
    double l[3] = {a.B[0] - a.A[0], a.B[1] - a.A[1], a.B[2] - a.A[2]};
    double inv = 1 / sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    for (int k = 0; k < 3; ++k) g[48 + k] = l[k] * inv;
  }
  if (tid == 0) {
    grand_seed(&s_rng, pd.seed);
    for (int i = 0; i < n; ++i) s_idx[i] = (uint16_t)i;
    for (int it = 0; it < maxIters; ++it) {
      int left = n;
      for (int k = 0; k < 3; ++k) {
        int r = grand_next(&s_rng) % left;
        uint16_t t = s_idx[k]; s_idx[k] = s_idx[k + r]; s_idx[k + r] = t;
        --left;
      }
      trip[3 * it] = s_idx[0]; trip[3 * it + 1] = s_idx[1]; trip[3 * it + 2] = s_idx[2];
    }
  }
  __syncthreads();
  for (int h = tid; h < maxIters; h += blockDim.x) {
    const double* g3[3] = {g_all + (size_t)trip[3 * h] * RM_STRIDE, g_all + (size_t)trip[3 * h + 1] * RM_STRIDE,
                           g_all + (size_t)trip[3 * h + 2] * RM_STRIDE};
    bool degenerate = true;
    for (int i = 0; i < 3 && degenerate; ++i)
      for (int j = i + 1; j < 3; ++j) {
        const double* ui = g3[i] + 48; const double* uj = g3[j] + 48;
        if (fabs(ui[0] * uj[0] + ui[1] * uj[1] + ui[2] * uj[2]) < cosDeg) { degenerate = false; break; }
      }
    if (degenerate) { cnts[h] = -1; continue; }
    double R[9], t[3];
    relmotion_svd3(g3, R, t);     // md layout qA qB tA tB == g[0..11]
    double* o = hyp + (size_t)h * 12;
    for (int k = 0; k < 9; ++k) o[k] = R[k];
    for (int k = 0; k < 3; ++k) o[9 + k] = t[k];
    cnts[h] = 0;
  }
  __syncthreads();
  for (int h = warp; h < maxIters; h += nwarp) {
    if (cnts[h] < 0) continue;
    double Rt[12];
    for (int k = 0; k < 12; ++k) Rt[k] = hyp[(size_t)h * 12 + k];
    int c = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
      int i = i0 + lane;
      bool in = i < n && rm_consistent(g_all + (size_t)i * RM_STRIDE, Rt, Rt + 9, distThresh, angThresh);
      c += __popc(__ballot_sync(FULL, in));
    }
    __syncwarp();
    if (lane == 0) cnts[h] = c;
  }
  __syncthreads();
  if (warp == 0) {
    int best = 0, bh = 1 << 30;
    for (int h = lane; h < maxIters; h += 32) { int c = cnts[h]; if (c > best) { best = c; bh = h; } }
    for (int o = 16; o; o >>= 1) {
      int b2 = __shfl_xor_sync(FULL, best, o), h2 = __shfl_xor_sync(FULL, bh, o);
      if (b2 > best || (b2 == best && h2 < bh)) { best = b2; bh = h2; }
    }
    if (lane == 0) s_best = best > 0 ? bh : -1;
  }
  __syncthreads();
  const int bh = s_best;
  if (bh < 0) return;   // maxConSet.size() < 1
  if (tid < 9) s_Ro[tid] = hyp[(size_t)bh * 12 + tid];
  if (tid < 3) s_to[tid] = hyp[(size_t)bh * 12 + 9 + tid];
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) flag[i] = rm_consistent(g_all + (size_t)i * RM_STRIDE, s_Ro, s_to, distThresh, angThresh);
  int ncur = rm_compact(flag, n, cur, &s_cnt);
  int lm_calls = 0, nout = ncur;
  if (ncur >= 4) {
    if (tid < 9) s_R[tid] = s_Ro[tid];
    if (tid < 3) s_t[tid] = s_to[tid];
    __syncthreads();
    rm_optimize(g_all, cur, ncur, s_R, s_t, hx, wrk, e, wrk2, jac, S); lm_calls++;
    if (tid < 9) s_Ro[tid] = s_R[tid];      // Ro, to = first optimisation result; R = Ro, t = to
    if (tid < 3) s_to[tid] = s_t[tid];
    __syncthreads();
    int nprev = 0;
    while (true) {
      for (int i = tid; i < n; i += blockDim.x) flag[i] = rm_consistent(g_all + (size_t)i * RM_STRIDE, s_R, s_t, distThresh, angThresh);
      int nc = rm_compact(flag, n, trial, &s_cnt);
      if (nc <= nprev) break;
      nprev = nc;
      for (int i = tid; i < nc; i += blockDim.x) cur[i] = trial[i];
      if (tid < 9) s_Ro[tid] = s_R[tid];
      if (tid < 3) s_to[tid] = s_t[tid];
      __syncthreads();
      rm_optimize(g_all, cur, nc, s_R, s_t, hx, wrk, e, wrk2, jac, S); lm_calls++;
    }
    nout = nprev;
  }
  __syncthreads();
  if (tid == 0) {
    for (int k = 0; k < 9; ++k) outRt[k] = s_Ro[k];
    for (int k = 0; k < 3; ++k) outRt[9 + k] = s_to[k];
    outn[0] = nout; outn[1] = lm_calls; outn[2] = 1;
  }
}

int lsl_launch_relmotion(lsl_ctx* ctx, int npairs, RmScratch rs) {
  const lsl_params& P = ctx->P;
  const double PI_T = 3.14159265;
  LSL_KSTART(ctx, LSL_K_RELMOTION);
  relmotion_kernel<<<npairs, RM_THREADS, 0, ctx->stream>>>(ctx->pw.d_pairs, ctx->pw.matches, ctx->pw.nmatch, rs, P.pt2line3d_dist_relmotion,
                                                          P.line3d_angle_relmotion, lsl_cos(5 * PI_T / 180), P.ransac_iters_line_motion);
  LSL_KSTOP(ctx, LSL_K_RELMOTION);
  LSL_CUDA(cudaGetLastError());
  return LSL_OK;
}
