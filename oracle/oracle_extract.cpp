// oracle_extract.cpp — CPU restatement of Node::detect3DLines and its helpers
// (TEST INFRASTRUCTURE). Reference: src/line/lineslam.cpp:200-357, 518-537;
// src/line/utils.cpp (cited per function); external/levmar-2.6/lm_core.c:438-847.
#include "oracle.h"
#include <math.h>
#include <float.h>
#include <string.h>
#include <algorithm>
#include "../lineslam_b200/csrc/shared/lsl_math.h"
#include "../lineslam_b200/csrc/shared/lsl_linalg.h"
#include "../lineslam_b200/csrc/shared/lsl_cvdraw.h"

using namespace lslm;

namespace orc {

// ---------------------------------------------------------------- rand() ----
// glibc random_r.c TYPE_3 (r[i] = r[i-31] + r[i-3]), srandom_r seeding + 310 discards.
void GlibcRand::seed(uint32_t s) {
  if (s == 0) s = 1;
  r[0] = (int32_t)s;
  for (int i = 1; i < 31; ++i) {
    long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
    long word = 16807 * lo - 2836 * hi;
    if (word < 0) word += 2147483647;
    r[i] = (int32_t)word;
  }
  f = 3; b = 0;
  for (int i = 0; i < 310; ++i) next();
}
int GlibcRand::next() {
  uint32_t v = (uint32_t)r[f] + (uint32_t)r[b];
  r[f] = (int32_t)v;
  int out = (int)(v >> 1);
  if (++f >= 31) f = 0;
  if (++b >= 31) b = 0;
  return out;
}

// ----------------------------------------------------------------- Sobel ----
// cv::Sobel(gray, CV_64F, dx, dy, 5), BORDER_REFLECT_101 (lineslam.cpp:313-314).
static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * n - 2 - i; }
  return i;
}
void sobel5(const uint8_t* gray, int W, int H, std::vector<double>& gx, std::vector<double>& gy) {
  static const int dv[5] = {-1, -2, 0, 2, 1}, sm[5] = {1, 4, 6, 4, 1};
  gx.assign((size_t)W * H, 0.0);
  gy.assign((size_t)W * H, 0.0);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int sx = 0, sy = 0;
      for (int j = 0; j < 5; ++j) {
        int yy = reflect101(y + j - 2, H);
        for (int i = 0; i < 5; ++i) {
          int v = gray[(size_t)yy * W + reflect101(x + i - 2, W)];
          sx += sm[j] * dv[i] * v;
          sy += dv[j] * sm[i] * v;
        }
      }
      gx[(size_t)y * W + x] = (double)sx;
      gy[(size_t)y * W + x] = (double)sy;
    }
}

// ---------------------------------------------------------- 3D line RANSAC ----
static double norm3(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// projectPt3d2Ln3d (utils.cpp:496-504) with mid / direction
static void project_pt(const double P[3], const double mid[3], const double drct[3], double out[3]) {
  double A[3] = {mid[0], mid[1], mid[2]};
  double B[3] = {mid[0] + drct[0], mid[1] + drct[1], mid[2] + drct[2]};
  double AB[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
  double AP[3] = {P[0] - A[0], P[1] - A[1], P[2] - A[2]};
  double s = dot3(AB, AP) / dot3(AB, AB);
  for (int k = 0; k < 3; ++k) out[k] = A[k] + s * AB[k];
}

// verify3dLine (utils.cpp:570-624)
static bool verify3dLine(const std::vector<Pt3>& pts, const std::vector<int>& idx, const double A[3],
                         const double B[3], const Params& P) {
  int nCells = P.num_cells_lineseg_range;
  std::vector<int> cells(nCells, 0);
  int nPts = (int)idx.size();
  double minv = 100, maxv = -100;
  int idx1 = 0, idx2 = 0;
  double BA[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
  for (int i = 0; i < nPts; ++i) {
    const double* p = pts[idx[i]].pos;
    double d[3] = {p[0] - A[0], p[1] - A[1], p[2] - A[2]};
    double v = dot3(d, BA);
    if (v < minv) { minv = v; idx1 = i; }
    if (v > maxv) { maxv = v; idx2 = i; }
  }
  double mid[3] = {(A[0] + B[0]) * 0.5, (A[1] + B[1]) * 0.5, (A[2] + B[2]) * 0.5};
  double C[3], D[3];
  project_pt(pts[idx[idx1]].pos, mid, BA, C);
  project_pt(pts[idx[idx2]].pos, mid, BA, D);
  double DC[3] = {D[0] - C[0], D[1] - C[1], D[2] - C[2]};
  double cd = norm3(DC);
  if (cd < 1e-10) return false;
  for (int i = 0; i < nPts; ++i) {
    const double* X = pts[idx[i]].pos;
    double XC[3] = {X[0] - C[0], X[1] - C[1], X[2] - C[2]};
    double lambda = fabs(dot3(XC, DC) / cd / cd);
    if (lambda >= 1) cells[nCells - 1] += 1;
    else cells[(unsigned)floor(lambda * 10)] += 1;
  }
  double sum = 0;
  for (int i = 0; i < nCells; ++i)
    if (cells[i] > 0) sum = sum + 1;
  return sum / nCells > P.ratio_support_pts_on_line;
}

// computeLine3d_svd (utils.cpp:471-493): PCA direction = dominant eigenvector of the scatter
// matrix (== first right singular vector of the centred n x 3 matrix; sign is irrelevant).
static void line3d_pca(const std::vector<Pt3>& pts, const std::vector<int>& idx, double mean[3], double drct[3]) {
  int n = (int)idx.size();
  mean[0] = mean[1] = mean[2] = 0;
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) mean[k] = mean[k] + pts[idx[i]].pos[k];
  for (int k = 0; k < 3; ++k) mean[k] = mean[k] * (1.0 / n);
  double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    double d[3] = {pts[idx[i]].pos[0] - mean[0], pts[idx[i]].pos[1] - mean[1], pts[idx[i]].pos[2] - mean[2]};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) S[a * 3 + b] += d[a] * d[b];
  }
  double w[3], V[9];
  jacobi_sym<3>(S, w, V);
  drct[0] = V[0]; drct[1] = V[3]; drct[2] = V[6];
}

struct Line3dFit {
  std::vector<int> inliers;
  double A[3], B[3];
};

// extract3dline_mahdist (utils.cpp:343-427)
static void extract3dline_mahdist(const std::vector<Pt3>& pts, GlibcRand& rng, const Params& P, Line3dFit& rl,
                                  int* draws) {
  int n = (int)pts.size();
  int maxIterNo = std::min(P.ransac_iters_extract_line, int(n * (n - 1) * 0.5));
  double distThresh = P.pt2line_mahdist_extractline;
  std::vector<int> indexes(n);
  for (int i = 0; i < n; ++i) indexes[i] = i;
  std::vector<int> maxInlierSet;
  int bestA = -1, bestB = -1;
  for (int iter = 0; iter < maxIterNo; iter++) {
    std::vector<int> inlierSet;
    {  // random_unique(begin, end, 2) — utils.h:49-60; swaps persist across iterations
      int left = n;
      for (int k = 0; k < 2; ++k) {
        int r = rng.next() % left;
        std::swap(indexes[k], indexes[k + r]);
        --left;
        if (draws) ++*draws;
      }
    }
    const Pt3& A = pts[indexes[0]];
    const Pt3& B = pts[indexes[1]];
    double BA[3] = {B.pos[0] - A.pos[0], B.pos[1] - A.pos[1], B.pos[2] - A.pos[2]};
    if (norm3(BA) < 1e-10) continue;
    for (int i = 0; i < n; ++i) {
      double dist = mah_dist3d_pt_line(pts[i].pos, pts[i].DU, A.pos, B.pos);
      if (dist < distThresh) inlierSet.push_back(i);
    }
    if (inlierSet.size() > maxInlierSet.size()) {
      if (verify3dLine(pts, inlierSet, A.pos, B.pos, P)) {
        maxInlierSet = inlierSet;
        bestA = indexes[0]; bestB = indexes[1];
      }
    }
    if (maxInlierSet.size() > n * 0.9) break;
  }
  rl.inliers.clear();
  for (int k = 0; k < 3; ++k) rl.A[k] = rl.B[k] = 0;
  if (maxInlierSet.size() >= 2) {
    double m[3], d[3];
    for (int k = 0; k < 3; ++k) {
      m[k] = (pts[bestA].pos[k] + pts[bestB].pos[k]) * 0.5;
      d[k] = pts[bestB].pos[k] - pts[bestA].pos[k];
    }
    while (true) {
      std::vector<int> tmpInlierSet;
      double tmp_m[3], tmp_d[3], q2[3];
      line3d_pca(pts, maxInlierSet, tmp_m, tmp_d);
      for (int k = 0; k < 3; ++k) q2[k] = tmp_m[k] + tmp_d[k];
      for (int i = 0; i < n; ++i)
        if (mah_dist3d_pt_line(pts[i].pos, pts[i].DU, tmp_m, q2) < distThresh) tmpInlierSet.push_back(i);
      if (tmpInlierSet.size() > maxInlierSet.size()) {
        maxInlierSet = tmpInlierSet;
        for (int k = 0; k < 3; ++k) { m[k] = tmp_m[k]; d[k] = tmp_d[k]; }
      } else break;
    }
    double minv = 100, maxv = -100;
    int idx_end1 = 0, idx_end2 = 0;
    for (int i = 0; i < (int)maxInlierSet.size(); ++i) {
      const double* p = pts[maxInlierSet[i]].pos;
      double pm[3] = {p[0] - m[0], p[1] - m[1], p[2] - m[2]};
      double dproduct = dot3(pm, d);
      if (dproduct < minv) { minv = dproduct; idx_end1 = i; }
      if (dproduct > maxv) { maxv = dproduct; idx_end2 = i; }
    }
    for (int k = 0; k < 3; ++k) {
      rl.A[k] = pts[maxInlierSet[idx_end1]].pos[k];
      rl.B[k] = pts[maxInlierSet[idx_end2]].pos[k];
    }
  }
  rl.inliers = maxInlierSet;
}

// ------------------------------------------------------------------ MSLD ----
// cv::norm(CV_64F vector): OpenCV 2.4 normL2_ pairs the squares two by two (UNVERIFIED,
// source not in the container; see DESIGN.md "third-party arithmetic").
double cvnorm(const double* v, int len) {
  double result = 0;
  int i = 0;
  for (; i <= len - 4; i += 4) {
    double v0 = v[i], v1 = v[i + 1];
    result += v0 * v0 + v1 * v1;
    v0 = v[i + 2]; v1 = v[i + 3];
    result += v0 * v0 + v1 * v1;
  }
  for (; i < len; i++) result += v[i] * v[i];
  return sqrt(result);
}

// computeSubPSR (utils.cpp:1510-1542)
static int computeSubPSR(const double* xG, const double* yG, double px, double py, double s, double gx_, double gy_,
                         int width, int height, double vs[4]) {
  double tl_x = floor(px - s / 2), tl_y = floor(py - s / 2);
  if (tl_x < 0 || tl_y < 0 || tl_x + s + 1 > width || tl_y + s + 1 > height) return 0;
  double v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  for (int y = (int)tl_y; y < tl_y + s; ++y)
    for (int x = (int)tl_x; x < tl_x + s; ++x) {
      double tmp1 = xG[y * width + x] * gx_ + yG[y * width + x] * gy_;
      double tmp2 = xG[y * width + x] * (-gy_) + yG[y * width + x] * gx_;
      if (tmp1 >= 0) v1 = v1 + tmp1; else v2 = v2 - tmp1;
      if (tmp2 >= 0) v3 = v3 + tmp2; else v4 = v4 - tmp2;
    }
  vs[0] = v1; vs[1] = v2; vs[2] = v3; vs[3] = v4;
  return 1;
}

// computeMSLD (utils.cpp:1544-1610)
static int computeMSLD(Line& l, const double* xG, const double* yG, int width, int height, const Params& P,
                       GlibcRand& rng, int* draws) {
  int s = 5 * width / 800.0;
  double dx = l.p[0] - l.q[0], dy = l.p[1] - l.q[1];
  double len = sqrt(dx * dx + dy * dy);
  std::vector<std::vector<double>> GDM;
  double step = P.msld_sample_interval;
  for (int i = 0; i * step < len; ++i) {
    std::vector<double> col;
    col.reserve(36);
    double f = i * step / len;
    double ptx = l.p[0] + (l.q[0] - l.p[0]) * f, pty = l.p[1] + (l.q[1] - l.p[1]) * f;
    bool fail = false;
    for (int j = -4; j <= 4; ++j) {
      double psr[4];
      int js = j * s;
      if (computeSubPSR(xG, yG, ptx + js * l.r[0], pty + js * l.r[1], s, l.r[0], l.r[1], width, height, psr)) {
        col.push_back(psr[0]); col.push_back(psr[1]); col.push_back(psr[2]); col.push_back(psr[3]);
      } else { fail = true; break; }
    }
    if (fail) continue;
    GDM.push_back(col);
  }
  if (GDM.size() == 0) {
    for (int i = 0; i < 72; ++i) { l.des[i] = rng.next(); if (draws) ++*draws; }
    return 0;
  }
  static const double gauss[9] = {0.24142, 0.30046, 0.35127, 0.38579, 0.39804, 0.38579, 0.35127, 0.30046, 0.24142};
  double MS[72];
  for (int i = 0; i < 36; ++i) {
    double sum = 0, sum2 = 0;
    for (size_t j = 0; j < GDM.size(); ++j) {
      GDM[j][i] = GDM[j][i] * gauss[i / 4];
      sum += GDM[j][i];
      sum2 += GDM[j][i] * GDM[j][i];
    }
    double mean = sum / GDM.size();
    double sd = sqrt(sum2 / GDM.size() - mean * mean);
    MS[i] = mean;
    MS[i + 36] = sd;
  }
  double a = 1. / cvnorm(MS, 36), b = 1. / cvnorm(MS + 36, 36);  // Mat / s  ==  Mat * (1./s)
  for (int i = 0; i < 36; ++i) { MS[i] = MS[i] * a; MS[i + 36] = MS[i + 36] * b; }
  for (int i = 0; i < 72; ++i)
    if (MS[i] > 0.4) MS[i] = 0.4;
  double c = 1. / cvnorm(MS, 72);
  for (int i = 0; i < 72; ++i) l.des[i] = MS[i] * c;
  return 1;
}

// FrameLine::getGradient (lineslam.cpp:527-537) over cv::LineIterator(8-connected)
static void getGradient(Line& l, const double* xG, const double* yG, int W, int H) {
  int x1 = (int)nearbyint(l.p[0]), y1 = (int)nearbyint(l.p[1]);  // cvRound: half to even
  int x2 = (int)nearbyint(l.q[0]), y2 = (int)nearbyint(l.q[1]);
  double xSum = 0, ySum = 0;
  // LineIterator ctor (OpenCV 2.4 drawing.cpp): end points outside the image -> cv::clipLine; count = 0 if nothing is left
  if (lslm::line_iter_endpoints(W, H, &x1, &y1, &x2, &y2)) {
    int dx = x2 - x1, dy = y2 - y1;
    int sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1;
    dx = dx < 0 ? -dx : dx; dy = dy < 0 ? -dy : dy;
    bool steep = dy > dx;
    int dmaj = steep ? dy : dx, dmin = steep ? dx : dy;
    int err = dmaj - (dmin + dmin), plusDelta = dmaj + dmaj, minusDelta = -(dmin + dmin);
    int count = dmaj + 1, x = x1, y = y1;
    for (int i = 0; i < count; ++i) {
      xSum += xG[(size_t)y * W + x];
      ySum += yG[(size_t)y * W + x];
      int mask = err < 0 ? -1 : 0;
      err += minusDelta + (plusDelta & mask);
      if (steep) { y += sy; if (mask) x += sx; } else { x += sx; if (mask) y += sy; }
    }
  }
  double len = sqrt(xSum * xSum + ySum * ySum);
  l.r[0] = xSum / len;
  l.r[1] = ySum / len;
}

void get_gradient_probe(const double* xG, const double* yG, int W, int H, const double* pq, double* r) {
  Line l;
  l.p[0] = pq[0]; l.p[1] = pq[1]; l.q[0] = pq[2]; l.q[1] = pq[3];
  getGradient(l, xG, yG, W, H);
  r[0] = l.r[0]; r[1] = l.r[1];
}

// --------------------------------------------------------------- levmar ----
// AX_EQ_B_LU, built-in Crout LU with implicit scaling (Axb_core.c:1140-1277)
static int ax_eq_b_lu(const double* A, const double* B, double* x, int m) {
  std::vector<double> a(A, A + m * m), work(m);
  std::vector<int> idx(m);
  int maxi = -1;
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
      for (int k = 0; k < m; ++k) std::swap(a[maxi * m + k], a[j * m + k]);
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

// LEVMAR_L2NRMXMY (misc_core.c:721-809): e = x - y (x NULL -> 0), returns ||e||^2 with the
// reference's 4-accumulator, downward, 8-unrolled summation order.
static double l2nrmxmy(double* e, const double* x, const double* y, int n) {
  const int blocksize = 8, bpwr = 3;
  double sum0 = 0.0, sum1 = 0.0, sum2 = 0.0, sum3 = 0.0;
  int blockn = (n >> bpwr) << bpwr;
#define EI(j) (e[j] = (x ? x[j] : 0.0) - y[j], e[j] * e[j])
  for (int i = blockn - 1; i > 0; i -= blocksize) {
    sum0 += EI(i); sum1 += EI(i - 1); sum2 += EI(i - 2); sum3 += EI(i - 3);
    sum0 += EI(i - 4); sum1 += EI(i - 5); sum2 += EI(i - 6); sum3 += EI(i - 7);
  }
  int i = blockn;
  if (i < n) {
    switch (n - i) {
      case 7: sum0 += EI(i); ++i;  // fallthrough
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

// dlevmar_dif (lm_core.c:438-847), forward differences only (opts[4] > 0), small-problem
// J^T J path (n*m <= 1024 always holds here), no covariance output.
int dlevmar_dif_restated(void (*func)(double*, double*, int, int, void*), double* p, double* x, int m, int n,
                         int itmax, const double opts[5], double info[10], void* adata) {
  if (n < m) return -1;
  double tau = opts[0], eps1 = opts[1], eps2 = opts[2], eps2_sq = opts[2] * opts[2], eps3 = opts[3], delta = opts[4];
  std::vector<double> e(n), hx(n), jacTe(m), jac((size_t)n * m), jacTjac((size_t)m * m), Dp(m), diag(m), pDp(m),
      wrk(n), wrk2(n);
  double mu = 0, jacTe_inf = 0, p_L2 = 0, tmp, p_eL2, pDp_eL2, Dp_L2 = DBL_MAX, dF, dL, init_p_eL2;
  int nu, nu2, stop = 0, nfev, njap = 0, nlss = 0, K = (m >= 10) ? m : 10, updjac = 0, updp = 1, newjac = 0, k;
  (*func)(p, hx.data(), m, n, adata); nfev = 1;
  p_eL2 = l2nrmxmy(e.data(), x, hx.data(), n);
  init_p_eL2 = p_eL2;
  if (!std::isfinite(p_eL2)) stop = 7;
  nu = 20;
  for (k = 0; k < itmax && !stop; ++k) {
    if (p_eL2 <= eps3) { stop = 6; break; }
    if ((updp && nu > 16) || updjac == K) {
      // LEVMAR_FDIF_FORW_JAC_APPROX (misc_core.c:137-172)
      for (int j = 0; j < m; ++j) {
        double d = 1E-04 * p[j];
        d = fabs(d);
        if (d < delta) d = delta;
        double t = p[j];
        p[j] += d;
        (*func)(p, wrk.data(), m, n, adata);
        p[j] = t;
        d = 1.0 / d;
        for (int i = 0; i < n; ++i) jac[(size_t)i * m + j] = (wrk[i] - hx[i]) * d;
      }
      ++njap; nfev += m;
      nu = 2; updjac = 0; updp = 0; newjac = 1;
    }
    if (newjac) {
      newjac = 0;
      for (int i = m * m; i-- > 0;) jacTjac[i] = 0.0;
      for (int i = m; i-- > 0;) jacTe[i] = 0.0;
      for (int l = n; l-- > 0;) {
        const double* jaclm = &jac[(size_t)l * m];
        for (int i = m; i-- > 0;) {
          double* jacTjacim = &jacTjac[(size_t)i * m];
          double alpha = jaclm[i];
          for (int j = i + 1; j-- > 0;) jacTjacim[j] += jaclm[j] * alpha;
          jacTe[i] += alpha * e[l];
        }
      }
      for (int i = m; i-- > 0;)
        for (int j = i + 1; j < m; ++j) jacTjac[i * m + j] = jacTjac[j * m + i];
      p_L2 = jacTe_inf = 0.0;
      for (int i = 0; i < m; ++i) {
        if (jacTe_inf < (tmp = fabs(jacTe[i]))) jacTe_inf = tmp;
        diag[i] = jacTjac[i * m + i];
        p_L2 += p[i] * p[i];
      }
    }
    if (jacTe_inf <= eps1) { Dp_L2 = 0.0; stop = 1; break; }
    if (k == 0) {
      tmp = DBL_MIN;
      for (int i = 0; i < m; ++i)
        if (diag[i] > tmp) tmp = diag[i];
      mu = tau * tmp;
    }
    for (int i = 0; i < m; ++i) jacTjac[i * m + i] += mu;
    int issolved = ax_eq_b_lu(jacTjac.data(), jacTe.data(), Dp.data(), m);
    ++nlss;
    if (issolved) {
      Dp_L2 = 0.0;
      for (int i = 0; i < m; ++i) { pDp[i] = p[i] + (tmp = Dp[i]); Dp_L2 += tmp * tmp; }
      if (Dp_L2 <= eps2_sq * p_L2) { stop = 2; break; }
      if (Dp_L2 >= (p_L2 + eps2) / (1E-12 * 1E-12)) { stop = 4; break; }
      (*func)(pDp.data(), wrk.data(), m, n, adata); ++nfev;
      pDp_eL2 = l2nrmxmy(wrk2.data(), x, wrk.data(), n);
      if (!std::isfinite(pDp_eL2)) { stop = 7; break; }
      dF = p_eL2 - pDp_eL2;
      if (updp || dF > 0) {
        for (int i = 0; i < n; ++i) {
          tmp = 0.0;
          for (int l = 0; l < m; ++l) tmp += jac[(size_t)i * m + l] * Dp[l];
          tmp = (wrk[i] - hx[i] - tmp) / Dp_L2;
          for (int j = 0; j < m; ++j) jac[(size_t)i * m + j] += tmp * Dp[j];
        }
        ++updjac;
        newjac = 1;
      }
      dL = 0.0;
      for (int i = 0; i < m; ++i) dL += Dp[i] * (mu * Dp[i] + jacTe[i]);
      if (dL > 0.0 && dF > 0.0) {
        tmp = (2.0 * dF / dL - 1.0);
        tmp = 1.0 - tmp * tmp * tmp;
        mu = mu * ((tmp >= 0.3333333334) ? tmp : 0.3333333334);
        nu = 2;
        for (int i = 0; i < m; ++i) p[i] = pDp[i];
        for (int i = 0; i < n; ++i) { e[i] = wrk2[i]; hx[i] = wrk[i]; }
        p_eL2 = pDp_eL2;
        updp = 1;
        continue;
      }
    }
    mu *= nu;
    nu2 = nu << 1;
    if (nu2 <= nu) { stop = 5; break; }
    nu = nu2;
    for (int i = 0; i < m; ++i) jacTjac[i * m + i] = diag[i];
  }
  if (k >= itmax) stop = 3;
  if (info) {
    info[0] = init_p_eL2; info[1] = p_eL2; info[2] = jacTe_inf; info[3] = Dp_L2;
    tmp = DBL_MIN;
    for (int i = 0; i < m; ++i) if (tmp < diag[i]) tmp = diag[i];
    info[4] = mu / tmp; info[5] = (double)k; info[6] = (double)stop; info[7] = (double)nfev;
    info[8] = (double)njap; info[9] = (double)nlss;
  }
  return (stop != 4 && stop != 7) ? k : -1;
}

// ------------------------------------------------------------------- MLE ----
struct MleData {
  int idx1, idx2;
  const std::vector<Pt3>* pts;
  double cov_inv1[9], cov_inv2[9];
};
// costFun_MLEstimateLine3d (utils.cpp:954-978)
static void mle_cost(double* p, double* error, int, int, void* adata) {
  const MleData* d = (const MleData*)adata;
  const std::vector<Pt3>& pts = *d->pts;
  for (size_t i = 0; i < pts.size(); ++i) {
    if ((int)i == d->idx1 || (int)i == d->idx2) {
      const double* C = ((int)i == d->idx1) ? d->cov_inv1 : d->cov_inv2;
      const double* e = ((int)i == d->idx1) ? p : p + 3;
      error[i] = mah_sq_pt(e, pts[i].pos, C);
    } else {
      error[i] = mah_dist3d_pt_line(pts[i].pos, pts[i].DU, p, p + 3);
    }
  }
}

// MLEstimateLine3d (utils.cpp:980-1050) + MleLine3dCov (:1138-1159)
static int MLEstimateLine3d(const std::vector<Pt3>& pts, Line& out, const double A0[3], const double B0[3],
                            const Params& P) {
  double minv = 100, maxv = -100;
  int idx_end1 = 0, idx_end2 = 0;
  double AB[3] = {A0[0] - B0[0], A0[1] - B0[1], A0[2] - B0[2]};
  for (size_t i = 0; i < pts.size(); ++i) {
    double d[3] = {pts[i].pos[0] - A0[0], pts[i].pos[1] - A0[1], pts[i].pos[2] - A0[2]};
    double dproduct = dot3(d, AB);
    if (dproduct < minv) { minv = dproduct; idx_end1 = (int)i; }
    if (dproduct > maxv) { maxv = dproduct; idx_end2 = (int)i; }
  }
  if (idx_end1 > idx_end2) std::swap(idx_end1, idx_end2);
  double opts[5] = {1E-03, 1E-10, 1E-20, 1E-20, 1E-06}, info[10];
  MleData data;
  data.pts = &pts; data.idx1 = idx_end1; data.idx2 = idx_end2;
  inv3(pts[idx_end1].cov, data.cov_inv1);
  inv3(pts[idx_end2].cov, data.cov_inv2);
  double para[6];
  std::vector<double> meas(pts.size(), 0.0);
  // paraVec is filled while scanning i: idx_end1 first, then idx_end2 (if equal: only 3 params
  // would exist in the reference and levmar would be called with m=3; never happens for >= 2 pts)
  for (int k = 0; k < 3; ++k) { para[k] = pts[idx_end1].pos[k]; para[3 + k] = pts[idx_end2].pos[k]; }
  int nit = dlevmar_dif_restated(mle_cost, para, meas.data(), 6, (int)pts.size(), P.line3d_mle_iter_num, opts, info, &data);
  for (int k = 0; k < 3; ++k) { out.A[k] = para[k]; out.B[k] = para[3 + k]; }
  // H = J^T J with J (3n x 6); rows accumulated in order
  double H[36];
  for (int i = 0; i < 36; ++i) H[i] = 0;
  for (size_t i = 0; i < pts.size(); ++i) {
    double J[18];
    for (int k = 0; k < 18; ++k) J[k] = 0;
    if ((int)i == idx_end1) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) J[r * 6 + c] = -pts[i].DU[r * 3 + c]; }
    else if ((int)i == idx_end2) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) J[r * 6 + 3 + c] = -pts[i].DU[r * 3 + c]; }
    else jac_rpt2ln(pts[i].pos, pts[i].DU, para, J);
    for (int r = 0; r < 3; ++r)
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) H[a * 6 + b] += J[r * 6 + a] * J[r * 6 + b];
  }
  double cov[36];
  inv_lu<6>(H, cov);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) { out.covA[r * 3 + c] = cov[r * 6 + c]; out.covB[r * 3 + c] = cov[(r + 3) * 6 + 3 + c]; }
  cov_to_DU(out.covA, out.DU_A, out.Wsqrt_A);
  cov_to_DU(out.covB, out.DU_B, out.Wsqrt_B);
  return nit;
}

// ------------------------------------------------------- detect3DLines ----
void detect3DLines(const uint8_t* gray, const float* depth, int W, int H, const double K[9], double asynch_dt,
                   uint32_t seed, const Params& P, std::vector<Line>& lines, ExtractDebug* dbg, int omp_threads) {
  lines.clear();
  std::vector<Segment> segs;
  lsd_detect(gray, W, H, P, segs);
  struct Cand { Line l; int seg; bool have; std::vector<Pt3> pts; std::vector<int> inl; };
  std::vector<Cand> all;
  for (size_t i = 0; i < segs.size(); ++i) {
    double a = segs[i].x1, b = segs[i].y1, c = segs[i].x2, d = segs[i].y2;
    if (sqrt((a - c) * (a - c) + (b - d) * (b - d)) > P.line_2d_len_thres) {
      Cand cd;
      memset(&cd.l, 0, sizeof(Line));
      cd.l.p[0] = a; cd.l.p[1] = b; cd.l.q[0] = c; cd.l.q[1] = d;
      cd.seg = (int)i; cd.have = false;
      all.push_back(cd);
    }
  }
  double Kinv[9];
  inv3(K, Kinv);
  GlibcRand rng;
  rng.seed(seed);
  int draws = 0;
  const bool par = omp_threads > 1;  // timing-only mode: per-line reseed (the reference itself is
                                     // non-deterministic under OpenMP: threads share rand())
#pragma omp parallel for schedule(dynamic) num_threads(omp_threads) if (par)
  for (int i = 0; i < (int)all.size(); ++i) {
    Line& L = all[i].l;
    double ddx = L.p[0] - L.q[0], ddy = L.p[1] - L.q[1];
    double len = sqrt(ddx * ddx + ddy * ddy);
    double numSmp = std::min(std::max(len / P.line_sample_interval, (double)P.line_sample_min_num),
                             (double)P.line_sample_max_num);
    std::vector<Pt3> pts3d;
    for (int j = 0; j <= numSmp; ++j) {
      double ptx = L.p[0] * (1 - j / numSmp) + L.q[0] * (j / numSmp);
      double pty = L.p[1] * (1 - j / numSmp) + L.q[1] * (j / numSmp);
      if (ptx < 0 || pty < 0 || ptx >= W || pty >= H) continue;
      int row, col;
      if ((floor(ptx) == ptx) && (floor(pty) == pty)) {
        col = std::max(int(ptx - 1), 0);
        row = std::max(int(pty - 1), 0);
      } else { col = int(ptx); row = int(pty); }
      double zval = -1;
      double depval = depth[(size_t)row * W + col];
      if (depval < 1e-10 || std::isnan((float)depval)) {}
      else zval = depval / P.depth_scaling;
      if (zval > 0) {
        double x0 = Kinv[0] * ptx + Kinv[1] * pty + Kinv[2] * 1.0;
        double x1 = Kinv[3] * ptx + Kinv[4] * pty + Kinv[5] * 1.0;
        double x2 = Kinv[6] * ptx + Kinv[7] * pty + Kinv[8] * 1.0;
        double inv = 1.0 / x2;  // Eigen 3.x: vec / scalar == vec * (1/scalar) for floating types
        x0 = x0 * inv; x1 = x1 * inv;
        Pt3 p;
        p.pos[0] = x0 * zval; p.pos[1] = x1 * zval; p.pos[2] = zval;
        pts3d.push_back(p);
      }
    }
    if (pts3d.size() < std::max(10.0, numSmp * P.collin_pts_ratio)) continue;
    for (size_t j = 0; j < pts3d.size(); ++j) {
      pt3d_cov(pts3d[j].pos, K[0], P.stdev_sample_pt_imgline, P.depth_stdev_coeff_c1, P.depth_stdev_coeff_c2,
               P.depth_stdev_coeff_c3, asynch_dt, pts3d[j].cov);
      cov_to_DU(pts3d[j].cov, pts3d[j].DU, pts3d[j].W_sqrt);
    }
    Line3dFit fit;
    if (par) { GlibcRand r2; r2.seed(seed + 1 + i); extract3dline_mahdist(pts3d, r2, P, fit, nullptr); }
    else extract3dline_mahdist(pts3d, rng, P, fit, &draws);
    double dAB[3] = {fit.A[0] - fit.B[0], fit.A[1] - fit.B[1], fit.A[2] - fit.B[2]};
    if (fit.inliers.size() / numSmp > P.collin_pts_ratio && norm3(dAB) > P.line_3d_len_thres_m) {
      all[i].have = true;
      for (int k = 0; k < 3; ++k) { L.A[k] = fit.A[k]; L.B[k] = fit.B[k]; }
      all[i].inl = fit.inliers;
      all[i].pts.resize(fit.inliers.size());
      for (size_t j = 0; j < fit.inliers.size(); ++j) all[i].pts[j] = pts3d[fit.inliers[j]];
    }
  }
  std::vector<double> gx, gy;
  sobel5(gray, W, H, gx, gy);
  std::vector<int> keep;
  for (size_t i = 0; i < all.size(); ++i)
    if (all[i].have) {
      Line& L = all[i].l;
      L.haveDepth = 1;
      L.lid = (int)keep.size();
      // complineEq2d (lineslam.h:139-150): (p,1) x (q,1), normalised by the first two entries
      double l0 = L.p[1] * 1 - 1 * L.q[1], l1 = 1 * L.q[0] - L.p[0] * 1, l2 = L.p[0] * L.q[1] - L.p[1] * L.q[0];
      double nrm = sqrt(l0 * l0 + l1 * l1);
      double inv = 1. / nrm;  // Mat / s  ==  Mat * (1./s)
      L.lineEq2d[0] = l0 * inv; L.lineEq2d[1] = l1 * inv; L.lineEq2d[2] = l2 * inv;
      getGradient(L, gx.data(), gy.data(), W, H);
      keep.push_back((int)i);
    }
  lines.resize(keep.size());
  std::vector<int> iters(keep.size(), 0);
  std::vector<std::vector<double>> dpts(keep.size());
  std::vector<double> a0b0(keep.size() * 6);
  // MSLD failure fill draws rand(): serial pass first keeps the stream order of a 1-thread run
  for (size_t i = 0; i < keep.size(); ++i) {
    lines[i] = all[keep[i]].l;
    computeMSLD(lines[i], gx.data(), gy.data(), W, H, P, rng, &draws);
  }
#pragma omp parallel for schedule(dynamic) num_threads(omp_threads) if (par)
  for (int i = 0; i < (int)keep.size(); ++i) {
    double A0[3], B0[3];
    for (int k = 0; k < 3; ++k) { A0[k] = lines[i].A[k]; B0[k] = lines[i].B[k]; a0b0[i * 6 + k] = A0[k]; a0b0[i * 6 + 3 + k] = B0[k]; }
    iters[i] = MLEstimateLine3d(all[keep[i]].pts, lines[i], A0, B0, P);
  }
  if (dbg) {
    dbg->segs = segs;
    dbg->seg_of_line.clear();
    dbg->inlier_idx.clear();
    dbg->pts.clear();
    for (size_t i = 0; i < keep.size(); ++i) {
      dbg->seg_of_line.push_back(all[keep[i]].seg);
      dbg->inlier_idx.push_back(all[keep[i]].inl);
      std::vector<double> pp;
      for (auto& q : all[keep[i]].pts) { pp.push_back(q.pos[0]); pp.push_back(q.pos[1]); pp.push_back(q.pos[2]); }
      dbg->pts.push_back(pp);
    }
    dbg->A0B0 = a0b0;
    dbg->lm_iters = iters;
    dbg->gx.swap(gx); dbg->gy.swap(gy);
    dbg->rand_draws = draws;
  }
}

}  // namespace orc
