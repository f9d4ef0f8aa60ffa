// oracle_lsd.cpp — CPU restatement of LSD 1.5 as LineSLAM calls it (TEST INFRASTRUCTURE).
// Follows external/lsd/lsd.cpp (function:line cited per routine) and callLsd
// (src/line/utils.cpp:112-135). Pinned against the unmodified upstream
// external/lsd/lsd-1.5/lsd.c built into oracle/_ref (tests/test_oracle_ref.py).
#include "oracle.h"
#include <math.h>
#include <float.h>
#include <string.h>
#include "../lineslam_b200/csrc/shared/lsl_math.h"

using namespace lslm;
#ifdef ORC_USE_LIBM  /* diagnostic build: glibc libm instead of the shared math */
#define lsl_exp exp
#define lsl_log log
#define lsl_log10 log10
#define lsl_sin sin
#define lsl_cos cos
#define lsl_atan2 atan2
#define lsl_pow pow
#define lsl_sinh sinh
#endif

namespace orc {

static const double NOTDEF = -1024.0;          // lsd.cpp:102
static const double M_3_2_PI_T = 4.71238898038; // lsd.cpp:105 (truncated on purpose)
static const double M_2__PI_T = 6.28318530718;  // lsd.cpp:108 (truncated on purpose)

void gray_from_3ch(const uint8_t* img, int W, int H, uint8_t* gray) {
  // OpenCV 2.4 RGB2Gray<uchar>: (c0*R2Y + c1*G2Y + c2*B2Y + (1<<13)) >> 14, coefficients
  // applied in MEMORY order (src/node.cpp:191-196 passes the BGR buffer with CV_RGB2GRAY).
  for (int i = 0; i < W * H; ++i) {
    const uint8_t* p = img + 3 * i;
    gray[i] = (uint8_t)((p[0] * 4899 + p[1] * 9617 + p[2] * 1868 + 8192) >> 14);
  }
}

// lsd.cpp:160-181
static bool double_equal(double a, double b) {
  if (a == b) return true;
  double abs_diff = fabs(a - b), aa = fabs(a), bb = fabs(b);
  double abs_max = aa > bb ? aa : bb;
  if (abs_max < DBL_MIN) abs_max = DBL_MIN;
  return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}
static double dist(double x1, double y1, double x2, double y2) {  // lsd.cpp:186
  return sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
}

// lsd.cpp:466-489
static void gaussian_kernel(double* k, int n, double sigma, double mean) {
  double sum = 0.0;
  for (int i = 0; i < n; ++i) {
    double val = ((double)i - mean) / sigma;
    k[i] = lsl_exp(-0.5 * val * val);
    sum += k[i];
  }
  if (sum >= 0.0)
    for (int i = 0; i < n; ++i) k[i] /= sum;
}

// lsd.cpp:529-646
static void gaussian_sampler(const std::vector<double>& in, int xs, int ys, double scale,
                             double sigma_scale, std::vector<double>& out, int& N, int& M) {
  N = (int)floor(xs * scale);
  M = (int)floor(ys * scale);
  std::vector<double> aux((size_t)N * ys);
  out.assign((size_t)N * M, 0.0);
  double sigma = scale < 1.0 ? sigma_scale / scale : sigma_scale;
  double prec = 3.0;
  int h = (int)ceil(sigma * sqrt(2.0 * prec * lsl_log(10.0)));
  int n = 1 + 2 * h;
  std::vector<double> kernel(n);
  int dxs = 2 * xs, dys = 2 * ys;
  for (int x = 0; x < N; ++x) {
    double xx = (double)x / scale;
    int xc = (int)floor(xx + 0.5);
    gaussian_kernel(kernel.data(), n, sigma, (double)h + xx - (double)xc);
    for (int y = 0; y < ys; ++y) {
      double sum = 0.0;
      for (int i = 0; i < n; ++i) {
        int j = xc - h + i;
        while (j < 0) j += dxs;
        while (j >= dxs) j -= dxs;
        if (j >= xs) j = dxs - 1 - j;
        sum += in[j + (size_t)y * xs] * kernel[i];
      }
      aux[x + (size_t)y * N] = sum;
    }
  }
  for (int y = 0; y < M; ++y) {
    double yy = (double)y / scale;
    int yc = (int)floor(yy + 0.5);
    gaussian_kernel(kernel.data(), n, sigma, (double)h + yy - (double)yc);
    for (int x = 0; x < N; ++x) {
      double sum = 0.0;
      for (int i = 0; i < n; ++i) {
        int j = yc - h + i;
        while (j < 0) j += dys;
        while (j >= dys) j -= dys;
        if (j >= ys) j = dys - 1 - j;
        sum += aux[x + (size_t)j * N] * kernel[i];
      }
      out[x + (size_t)y * N] = sum;
    }
  }
}

// lsd.cpp:670-794. The linked-list bins become per-bin vectors; concatenating bins
// n_bins-1 .. 1 reproduces list_p (bin 0 is never appended: loop at :781 stops at i>0,
// unless bin 0 is the highest non-empty one — kept for exactness).
static void ll_angle(const std::vector<double>& in, int p, int n, double threshold, int n_bins,
                     double max_grad, std::vector<double>& g, std::vector<double>& modgrad,
                     std::vector<int32_t>& seeds) {
  g.assign((size_t)p * n, 0.0);
  modgrad.assign((size_t)p * n, 0.0);
  std::vector<std::vector<int32_t>> bins(n_bins);
  for (int x = 0; x < p; ++x) g[(size_t)(n - 1) * p + x] = NOTDEF;
  for (int y = 0; y < n; ++y) g[(size_t)p * y + p - 1] = NOTDEF;
  for (int x = 0; x < p - 1; ++x)
    for (int y = 0; y < n - 1; ++y) {
      size_t adr = (size_t)y * p + x;
      double com1 = in[adr + p + 1] - in[adr];
      double com2 = in[adr + 1] - in[adr + p];
      double gx = com1 + com2, gy = com1 - com2;
      double norm2 = gx * gx + gy * gy;
      double norm = sqrt(norm2 / 4.0);
      modgrad[adr] = norm;
      if (norm <= threshold) g[adr] = NOTDEF;
      else {
        g[adr] = lsl_atan2(gx, -gy);
        unsigned i = (unsigned)(norm * (double)n_bins / max_grad);
        if (i >= (unsigned)n_bins) i = n_bins - 1;
        bins[i].push_back(x | (y << 16));
      }
    }
  seeds.clear();
  int i = n_bins - 1;
  for (; i > 0 && bins[i].empty(); --i) {}
  seeds.insert(seeds.end(), bins[i].begin(), bins[i].end());
  if (!bins[i].empty())
    for (--i; i > 0; --i) seeds.insert(seeds.end(), bins[i].begin(), bins[i].end());
}

struct Img {
  int xs, ys;
  const double* angles;
  const double* modgrad;
  uint8_t* used;
};
struct Pt { int x, y; };
struct Rect { double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p; };  // lsd.cpp:1075-1084

// lsd.cpp:799-832
static bool isaligned(int x, int y, const Img& I, double theta, double prec) {
  double a = I.angles[x + (size_t)y * I.xs];
  if (a == NOTDEF) return false;
  theta -= a;
  if (theta < 0.0) theta = -theta;
  if (theta > M_3_2_PI_T) {
    theta -= M_2__PI_T;
    if (theta < 0.0) theta = -theta;
  }
  return theta < prec;
}
static double angle_diff(double a, double b) {  // lsd.cpp:837
  a -= b;
  while (a <= -LSL_PI) a += M_2__PI_T;
  while (a > LSL_PI) a -= M_2__PI_T;
  if (a < 0.0) a = -a;
  return a;
}
static double angle_diff_signed(double a, double b) {  // lsd.cpp:849
  a -= b;
  while (a <= -LSL_PI) a += M_2__PI_T;
  while (a > LSL_PI) a -= M_2__PI_T;
  return a;
}

// lsd.cpp:886-931
static double log_gamma_lanczos(double x) {
  static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705,
                              1168.92649479, 83.8676043424, 2.50662827511};
  double a = (x + 0.5) * lsl_log(x + 5.5) - (x + 5.5);
  double b = 0.0;
  for (int n = 0; n < 7; ++n) {
    a -= lsl_log(x + (double)n);
    b += q[n] * lsl_pow(x, (double)n);
  }
  return a + lsl_log(b);
}
static double log_gamma_windschitl(double x) {
  return 0.918938533204673 + (x - 0.5) * lsl_log(x) - x +
         0.5 * x * lsl_log(x * lsl_sinh(1 / x) + 1 / (810.0 * lsl_pow(x, 6.0)));
}
static double log_gamma(double x) { return x > 15.0 ? log_gamma_windschitl(x) : log_gamma_lanczos(x); }

// lsd.cpp:980-1065 (the static inv[] table only caches 1.0/i)
static double nfa(int n, int k, double p, double logNT) {
  const double tolerance = 0.1;
  if (n == 0 || k == 0) return -logNT;
  if (n == k) return -logNT - (double)n * lsl_log10(p);
  double p_term = p / (1.0 - p);
  double log1term = log_gamma((double)n + 1.0) - log_gamma((double)k + 1.0) -
                    log_gamma((double)(n - k) + 1.0) + (double)k * lsl_log(p) +
                    (double)(n - k) * lsl_log(1.0 - p);
  double term = lsl_exp(log1term);
  if (double_equal(term, 0.0)) {
    if ((double)k > (double)n * p) return -log1term / LSL_LN10 - logNT;
    return -logNT;
  }
  double bin_tail = term;
  for (int i = k + 1; i <= n; ++i) {
    double bin_term = (double)(n - i + 1) * (1.0 / (double)i);
    double mult_term = bin_term * p_term;
    term *= mult_term;
    bin_tail += term;
    if (bin_term < 1.0) {
      double err = term * ((1.0 - lsl_pow(mult_term, (double)(n - i + 1))) / (1.0 - mult_term) - 1.0);
      if (err < tolerance * fabs(-lsl_log10(bin_tail) - logNT) * bin_tail) break;
    }
  }
  return -lsl_log10(bin_tail) - logNT;
}

// lsd.cpp:1183-1208
static double inter_low(double x, double x1, double y1, double x2, double y2) {
  if (double_equal(x1, x2) && y1 < y2) return y1;
  if (double_equal(x1, x2) && y1 > y2) return y2;
  return y1 + (x - x1) * (y2 - y1) / (x2 - x1);
}
static double inter_hi(double x, double x1, double y1, double x2, double y2) {
  if (double_equal(x1, x2) && y1 < y2) return y2;
  if (double_equal(x1, x2) && y1 > y2) return y1;
  return y1 + (x - x1) * (y2 - y1) / (x2 - x1);
}

// Rectangle iterator of lsd.cpp:1165-1383 unrolled into a column scan: the pixel set is
// { (x,y) : ceil(vx[0]) <= x <= vx[2], ceil(ys(x)) <= y <= ye(x) }.
static double rect_nfa(const Rect& r, const Img& I, double logNT) {  // lsd.cpp:1388-1410
  double vx0[4], vy0[4], vx[4], vy[4];
  vx0[0] = r.x1 - r.dy * r.width / 2.0; vy0[0] = r.y1 + r.dx * r.width / 2.0;
  vx0[1] = r.x2 - r.dy * r.width / 2.0; vy0[1] = r.y2 + r.dx * r.width / 2.0;
  vx0[2] = r.x2 + r.dy * r.width / 2.0; vy0[2] = r.y2 - r.dx * r.width / 2.0;
  vx0[3] = r.x1 + r.dy * r.width / 2.0; vy0[3] = r.y1 - r.dx * r.width / 2.0;
  int offset;
  if (r.x1 < r.x2 && r.y1 <= r.y2) offset = 0;
  else if (r.x1 >= r.x2 && r.y1 < r.y2) offset = 1;
  else if (r.x1 > r.x2 && r.y1 >= r.y2) offset = 2;
  else offset = 3;
  for (int n = 0; n < 4; ++n) { vx[n] = vx0[(offset + n) % 4]; vy[n] = vy0[(offset + n) % 4]; }
  int pts = 0, alg = 0;
  for (int x = (int)ceil(vx[0]); (double)x <= vx[2]; ++x) {
    double ys, ye;
    if ((double)x < vx[3]) ys = inter_low((double)x, vx[0], vy[0], vx[3], vy[3]);
    else ys = inter_low((double)x, vx[3], vy[3], vx[2], vy[2]);
    if ((double)x < vx[1]) ye = inter_hi((double)x, vx[0], vy[0], vx[1], vy[1]);
    else ye = inter_hi((double)x, vx[1], vy[1], vx[2], vy[2]);
    for (int y = (int)ceil(ys); (double)y <= ye; ++y)
      if (x >= 0 && y >= 0 && x < I.xs && y < I.ys) {
        ++pts;
        if (isaligned(x, y, I, r.theta, r.prec)) ++alg;
      }
  }
  return nfa(pts, alg, r.p, logNT);
}

// lsd.cpp:1474-1512
static double get_theta(const Pt* reg, int reg_size, double x, double y, const Img& I,
                        double reg_angle, double prec) {
  double Ixx = 0.0, Iyy = 0.0, Ixy = 0.0;
  for (int i = 0; i < reg_size; ++i) {
    double weight = I.modgrad[reg[i].x + (size_t)reg[i].y * I.xs];
    Ixx += ((double)reg[i].y - y) * ((double)reg[i].y - y) * weight;
    Iyy += ((double)reg[i].x - x) * ((double)reg[i].x - x) * weight;
    Ixy -= ((double)reg[i].x - x) * ((double)reg[i].y - y) * weight;
  }
  double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
  double theta = fabs(Ixx) > fabs(Iyy) ? lsl_atan2(lambda - Ixx, Ixy) : lsl_atan2(Ixy, lambda - Iyy);
  if (angle_diff(theta, reg_angle) > prec) theta += LSL_PI;
  return theta;
}

// lsd.cpp:1517-1604
static void region2rect(const Pt* reg, int reg_size, const Img& I, double reg_angle, double prec,
                        double p, Rect& rec) {
  double x = 0.0, y = 0.0, sum = 0.0;
  for (int i = 0; i < reg_size; ++i) {
    double weight = I.modgrad[reg[i].x + (size_t)reg[i].y * I.xs];
    x += (double)reg[i].x * weight;
    y += (double)reg[i].y * weight;
    sum += weight;
  }
  x /= sum;
  y /= sum;
  double theta = get_theta(reg, reg_size, x, y, I, reg_angle, prec);
  double dx = lsl_cos(theta), dy = lsl_sin(theta);
  double l_min = 0.0, l_max = 0.0, w_min = 0.0, w_max = 0.0;
  for (int i = 0; i < reg_size; ++i) {
    double l = ((double)reg[i].x - x) * dx + ((double)reg[i].y - y) * dy;
    double w = -((double)reg[i].x - x) * dy + ((double)reg[i].y - y) * dx;
    if (l > l_max) l_max = l;
    if (l < l_min) l_min = l;
    if (w > w_max) w_max = w;
    if (w < w_min) w_min = w;
  }
  rec.x1 = x + l_min * dx; rec.y1 = y + l_min * dy;
  rec.x2 = x + l_max * dx; rec.y2 = y + l_max * dy;
  rec.width = w_max - w_min;
  rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
  if (rec.width < 1.0) rec.width = 1.0;
}

// lsd.cpp:1610-1656
static void region_grow(int x, int y, const Img& I, Pt* reg, int& reg_size, double& reg_angle,
                        double prec) {
  reg_size = 1;
  reg[0].x = x; reg[0].y = y;
  reg_angle = I.angles[x + (size_t)y * I.xs];
  double sumdx = lsl_cos(reg_angle), sumdy = lsl_sin(reg_angle);
  I.used[x + (size_t)y * I.xs] = 1;
  for (int i = 0; i < reg_size; ++i)
    for (int xx = reg[i].x - 1; xx <= reg[i].x + 1; ++xx)
      for (int yy = reg[i].y - 1; yy <= reg[i].y + 1; ++yy)
        if (xx >= 0 && yy >= 0 && xx < I.xs && yy < I.ys && I.used[xx + (size_t)yy * I.xs] != 1 &&
            isaligned(xx, yy, I, reg_angle, prec)) {
          I.used[xx + (size_t)yy * I.xs] = 1;
          reg[reg_size].x = xx; reg[reg_size].y = yy;
          ++reg_size;
          double a = I.angles[xx + (size_t)yy * I.xs];
          sumdx += lsl_cos(a);
          sumdy += lsl_sin(a);
          reg_angle = lsl_atan2(sumdy, sumdx);
        }
}

// lsd.cpp:1662-1768
static double rect_improve(Rect& rec, const Img& I, double logNT, double eps) {
  Rect r;
  const double delta = 0.5, delta_2 = delta / 2.0;
  double log_nfa = rect_nfa(rec, I, logNT), log_nfa_new;
  if (log_nfa > eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n) {
    r.p /= 2.0; r.prec = r.p * LSL_PI;
    log_nfa_new = rect_nfa(r, I, logNT);
    if (log_nfa_new > log_nfa) { log_nfa = log_nfa_new; rec = r; }
  }
  if (log_nfa > eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n)
    if ((r.width - delta) >= 0.5) {
      r.width -= delta;
      log_nfa_new = rect_nfa(r, I, logNT);
      if (log_nfa_new > log_nfa) { rec = r; log_nfa = log_nfa_new; }
    }
  if (log_nfa > eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n)
    if ((r.width - delta) >= 0.5) {
      r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2;
      r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2;
      r.width -= delta;
      log_nfa_new = rect_nfa(r, I, logNT);
      if (log_nfa_new > log_nfa) { rec = r; log_nfa = log_nfa_new; }
    }
  if (log_nfa > eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n)
    if ((r.width - delta) >= 0.5) {
      r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2;
      r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2;
      r.width -= delta;
      log_nfa_new = rect_nfa(r, I, logNT);
      if (log_nfa_new > log_nfa) { rec = r; log_nfa = log_nfa_new; }
    }
  if (log_nfa > eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n) {
    r.p /= 2.0; r.prec = r.p * LSL_PI;
    log_nfa_new = rect_nfa(r, I, logNT);
    if (log_nfa_new > log_nfa) { log_nfa = log_nfa_new; rec = r; }
  }
  return log_nfa;
}

// lsd.cpp:1775-1841
static bool reduce_region_radius(Pt* reg, int& reg_size, const Img& I, double reg_angle, double prec,
                                 double p, Rect& rec, double density_th) {
  double density = (double)reg_size / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
  if (density >= density_th) return true;
  double xc = (double)reg[0].x, yc = (double)reg[0].y;
  double rad1 = dist(xc, yc, rec.x1, rec.y1), rad2 = dist(xc, yc, rec.x2, rec.y2);
  double rad = rad1 > rad2 ? rad1 : rad2;
  while (density < density_th) {
    rad *= 0.75;
    for (int i = 0; i < reg_size; ++i)
      if (dist(xc, yc, (double)reg[i].x, (double)reg[i].y) > rad) {
        I.used[reg[i].x + (size_t)reg[i].y * I.xs] = 0;
        reg[i] = reg[reg_size - 1];
        --reg_size;
        --i;
      }
    if (reg_size < 2) return false;
    region2rect(reg, reg_size, I, reg_angle, prec, p, rec);
    density = (double)reg_size / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
  }
  return true;
}

// lsd.cpp:1853-1921
static bool refine(Pt* reg, int& reg_size, const Img& I, double reg_angle, double prec, double p,
                   Rect& rec, double density_th) {
  double density = (double)reg_size / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
  if (density >= density_th) return true;
  double xc = (double)reg[0].x, yc = (double)reg[0].y;
  double ang_c = I.angles[reg[0].x + (size_t)reg[0].y * I.xs];
  double sum = 0.0, s_sum = 0.0;
  int n = 0;
  for (int i = 0; i < reg_size; ++i) {
    I.used[reg[i].x + (size_t)reg[i].y * I.xs] = 0;
    if (dist(xc, yc, (double)reg[i].x, (double)reg[i].y) < rec.width) {
      double angle = I.angles[reg[i].x + (size_t)reg[i].y * I.xs];
      double ang_d = angle_diff_signed(angle, ang_c);
      sum += ang_d;
      s_sum += ang_d * ang_d;
      ++n;
    }
  }
  double mean_angle = sum / (double)n;
  double tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)n + mean_angle * mean_angle);
  region_grow(reg[0].x, reg[0].y, I, reg, reg_size, reg_angle, tau);
  if (reg_size < 2) return false;
  region2rect(reg, reg_size, I, reg_angle, prec, p, rec);
  density = (double)reg_size / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
  if (density < density_th)
    return reduce_region_radius(reg, reg_size, I, reg_angle, prec, p, rec, density_th);
  return true;
}

void lsd_detect(const uint8_t* gray, int W, int H, const Params& P, std::vector<Segment>& out,
                LsdDebug* dbg) {
  out.clear();
  // callLsd: u8 -> double copy (utils.cpp:112-135); values are exact
  std::vector<double> image((size_t)W * H);
  for (size_t i = 0; i < image.size(); ++i) image[i] = (double)gray[i];
  // LineSegmentDetection (lsd.cpp:1931-2065)
  double prec = LSL_PI * P.lsd_ang_th / 180.0;
  double p = P.lsd_ang_th / 180.0;
  double rho = P.lsd_quant / lsl_sin(prec);
  std::vector<double> scaled, angles, modgrad;
  std::vector<int32_t> seeds;
  int xs, ys;
  if (P.lsd_scale != 1.0) gaussian_sampler(image, W, H, P.lsd_scale, P.lsd_sigma_scale, scaled, xs, ys);
  else { scaled = image; xs = W; ys = H; }
  ll_angle(scaled, xs, ys, rho, P.lsd_n_bins, P.lsd_max_grad, angles, modgrad, seeds);
  double logNT = 5.0 * (lsl_log10((double)xs) + lsl_log10((double)ys)) / 2.0;
  int min_reg_size = (int)(-logNT / lsl_log10(p));
  std::vector<uint8_t> used((size_t)xs * ys, 0);
  std::vector<Pt> reg((size_t)xs * ys);
  Img I{xs, ys, angles.data(), modgrad.data(), used.data()};
  for (size_t s = 0; s < seeds.size(); ++s) {
    int sx = seeds[s] & 0xffff, sy = seeds[s] >> 16;
    if (used[sx + (size_t)sy * xs] != 0 || angles[sx + (size_t)sy * xs] == NOTDEF) continue;
    int reg_size;
    double reg_angle;
    Rect rec;
    region_grow(sx, sy, I, reg.data(), reg_size, reg_angle, prec);
    if (reg_size < min_reg_size) continue;
    region2rect(reg.data(), reg_size, I, reg_angle, prec, p, rec);
    if (!refine(reg.data(), reg_size, I, reg_angle, prec, p, rec, P.lsd_density_th)) continue;
    double log_nfa = rect_improve(rec, I, logNT, P.lsd_eps);
    if (log_nfa <= P.lsd_eps) continue;
    rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
    if (P.lsd_scale != 1.0) {
      rec.x1 /= P.lsd_scale; rec.y1 /= P.lsd_scale;
      rec.x2 /= P.lsd_scale; rec.y2 /= P.lsd_scale;
      rec.width /= P.lsd_scale;
    }
    out.push_back(Segment{rec.x1, rec.y1, rec.x2, rec.y2, rec.width});
  }
  if (dbg) {
    dbg->sw = xs; dbg->sh = ys;
    dbg->scaled.swap(scaled); dbg->angles.swap(angles); dbg->modgrad.swap(modgrad);
    dbg->seeds.swap(seeds);
  }
}

}  // namespace orc
