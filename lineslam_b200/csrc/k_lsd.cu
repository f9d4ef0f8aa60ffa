// k_lsd.cu — K3: LSD region growing, rectangle fitting, density refinement and NFA validation
// (external/lsd/lsd.cpp:799-2065) for a batch of frames, one warp per frame (sm_100a).
//
// The reference loop is sequential by construction: seeds are consumed in list order, a pixel
// claimed by an earlier region is unavailable later, and refine()/reduce_region_radius() un-mark
// pixels. The frame-level order is therefore kept (one warp walks one frame's seed list) and the
// parallelism comes from (i) the 32 lanes inside every step and (ii) many frames per launch.
//
//  * region_grow: lanes 0..8 fetch the 3x3 neighbourhood of reg[i] in the reference's (xx, yy)
//    order; candidates are then resolved in that order against the running sums. The alignment test
//    |atan2(S,C) - a| < prec is decided from dot/cross products of (C,S) with the pixel's stored
//    (cos a, sin a) with a 1e-9 guard band; only inside the band is the reference's exact
//    atan2-based test evaluated, so the decision always equals the reference's and the expensive
//    region angle is needed once per region instead of once per pixel.
//  * region2rect / get_theta / refine: the floating-point sums run in reg[] order (lanes load 32
//    terms, the adds are chained in order); min/max extents are order-free warp reductions.
//  * rect_nfa: lanes own rectangle columns; nfa(): the independent log/pow terms of the three
//    log-gammas are spread over lanes, the binomial tail is evaluated 32 terms at a time.
#include "lsl_internal.h"
#include "shared/lsl_math.h"
#include <float.h>

using namespace lslm;

#define FULL 0xffffffffu
#define M_3_2_PI_T 4.71238898038  // lsd.cpp:105 (truncated on purpose)
#define M_2__PI_T 6.28318530718   // lsd.cpp:108

struct Rect { double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p; };

struct FrameView {
  int xs, ys;
  const double* angles;
  const double* modgrad;
  const double2* cs;
  uint8_t* used;
  int32_t* reg;
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ bool dbl_equal(double a, double b) {  // lsd.cpp:160-181
  if (a == b) return true;
  double abs_diff = fabs(a - b), aa = fabs(a), bb = fabs(b);
  double abs_max = aa > bb ? aa : bb;
  if (abs_max < DBL_MIN) abs_max = DBL_MIN;
  return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}
__device__ __forceinline__ double dist2d(double x1, double y1, double x2, double y2) {
  return sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
}
__device__ __forceinline__ bool isaligned_exact(double a, double theta, double prec) {  // lsd.cpp:799-832
  if (a == LSL_NOTDEF) return false;
  theta -= a;
  if (theta < 0.0) theta = -theta;
  if (theta > M_3_2_PI_T) {
    theta -= M_2__PI_T;
    if (theta < 0.0) theta = -theta;
  }
  return theta < prec;
}
__device__ __forceinline__ double angle_diff_signed(double a, double b) {  // lsd.cpp:849
  a -= b;
  while (a <= -LSL_PI) a += M_2__PI_T;
  while (a > LSL_PI) a -= M_2__PI_T;
  return a;
}
__device__ __forceinline__ double angle_diff(double a, double b) {
  a = angle_diff_signed(a, b);
  return a < 0.0 ? -a : a;
}

// ------------------------------------------------------------ region_grow ----
// Returns the region size; *Cs,*Ss are the exact running sums of cos/sin in reg[] order
// (== sumdx, sumdy of lsd.cpp:1638-1655). seed_angle is reg_angle while the region has one pixel.
__device__ int region_grow(const FrameView& V, int seed, double prec, double tanp, bool fast_ok, double* Cs, double* Ss,
                           double* seed_angle) {
  const int lane = threadIdx.x & 31;
  const int xs = V.xs, ys = V.ys, npx = xs * ys;
  int sx = seed & 0xffff, sy = seed >> 16;
  int sidx = sx + sy * xs;
  double2 c0 = V.cs[sidx];
  double C = c0.x, S = c0.y;
  double a0 = V.angles[sidx];
  if (lane == 0) { V.reg[0] = seed; V.used[sidx] = 1; }
  __syncwarp();
  int n = 1;
  const int dxk = lane / 3 - 1, dyk = lane % 3 - 1;  // lanes 0..8: xx outer, yy inner
  for (int i = 0; i < n; ++i) {
    int pi = V.reg[i];
    int px = pi & 0xffff, py = pi >> 16;
    int xx = px + dxk, yy = py + dyk;
    bool cand = lane < 9 && lane != 4 && xx >= 0 && yy >= 0 && xx < xs && yy < ys;
    int idx = xx + yy * xs;
    double2 c2 = make_double2(2.0, 0.0);
    if (cand) {   // both loads are issued together: one memory latency per step instead of two dependent ones
      const uint8_t u = V.used[idx];
      const double2 cc = V.cs[idx];
      cand = (u == 0) && (cc.x <= 1.5);
      if (cand) c2 = cc;
    }
    unsigned mask = __ballot_sync(FULL, cand);
    while (mask) {
      int k = __ffs(mask) - 1;
      mask &= mask - 1;
      double ck = shfl_d(c2.x, k), sk = shfl_d(c2.y, k);
      int idxk = __shfl_sync(FULL, idx, k);
      double dot = C * ck + S * sk;
      double crs = fabs(C * sk - S * ck);
      double rhs = tanp * dot;
      bool aligned;
      if (fast_ok && dot > 0.0 && crs < rhs * (1.0 - 1e-9)) aligned = true;
      else if (fast_ok && (dot <= 0.0 || crs > rhs * (1.0 + 1e-9))) aligned = false;
      else {
        double theta = (n == 1) ? a0 : lsl_atan2(S, C);
        aligned = isaligned_exact(V.angles[idxk], theta, prec);
      }
      if (aligned) {
        if (lane == k) { V.used[idx] = 1; V.reg[n] = xx | (yy << 16); }
        ++n;
        C += ck; S += sk;
        // the 3 x 3 neighbourhood of this pixel is examined when the walk reaches reg[n - 1]: pull its rows of the
        // `cs` (lanes 0..2) and `used` (lanes 3..5) planes into L1 now (these two dependent loads were 40 % of
        // the kernel's stall samples)
        if (lane < 6) {
          const int row = lane < 3 ? lane - 1 : lane - 4;
          int pi = idxk + row * xs - 1;
          pi = pi < 0 ? 0 : (pi > npx - 1 ? npx - 1 : pi);
          const void* ptr = lane < 3 ? (const void*)(V.cs + pi) : (const void*)(V.used + pi);
          asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
          if (lane < 3) asm volatile("prefetch.global.L1 [%0];" ::"l"((const void*)(V.cs + min(pi + 2, npx - 1))));
        }
      }
    }
    __syncwarp();
  }
  *Cs = C; *Ss = S; *seed_angle = a0;
  return n;
}

// ------------------------------------------------------------ region2rect ----
// Ordered sums: every lane holds one term of a 32-chunk; all lanes chain the adds in order.
__device__ void region2rect(const FrameView& V, int n, double reg_angle, double prec, double p, Rect* rec) {
  const int lane = threadIdx.x & 31;
  double x = 0.0, y = 0.0, sum = 0.0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    int i = i0 + lane;
    double w = 0.0, xw = 0.0, yw = 0.0;
    if (i < n) {
      int pi = V.reg[i];
      int px = pi & 0xffff, py = pi >> 16;
      w = V.modgrad[px + py * V.xs];
      xw = (double)px * w; yw = (double)py * w;
    }
    int cnt = min(32, n - i0);
    for (int k = 0; k < cnt; ++k) { x += shfl_d(xw, k); y += shfl_d(yw, k); sum += shfl_d(w, k); }
  }
  x /= sum; y /= sum;
  // get_theta (lsd.cpp:1474-1512)
  double Ixx = 0.0, Iyy = 0.0, Ixy = 0.0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    int i = i0 + lane;
    double a = 0.0, b = 0.0, c = 0.0;
    if (i < n) {
      int pi = V.reg[i];
      int px = pi & 0xffff, py = pi >> 16;
      double w = V.modgrad[px + py * V.xs];
      a = ((double)py - y) * ((double)py - y) * w;
      b = ((double)px - x) * ((double)px - x) * w;
      c = ((double)px - x) * ((double)py - y) * w;
    }
    int cnt = min(32, n - i0);
    for (int k = 0; k < cnt; ++k) { Ixx += shfl_d(a, k); Iyy += shfl_d(b, k); Ixy -= shfl_d(c, k); }
  }
  double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
  double theta = fabs(Ixx) > fabs(Iyy) ? lsl_atan2(lambda - Ixx, Ixy) : lsl_atan2(Ixy, lambda - Iyy);
  if (angle_diff(theta, reg_angle) > prec) theta += LSL_PI;
  double dx, dy;
  lsl_sincos(theta, &dy, &dx);
  double l_min = 0.0, l_max = 0.0, w_min = 0.0, w_max = 0.0;
  for (int i = lane; i < n; i += 32) {
    int pi = V.reg[i];
    int px = pi & 0xffff, py = pi >> 16;
    double l = ((double)px - x) * dx + ((double)py - y) * dy;
    double w = -((double)px - x) * dy + ((double)py - y) * dx;
    if (l > l_max) l_max = l;
    if (l < l_min) l_min = l;
    if (w > w_max) w_max = w;
    if (w < w_min) w_min = w;
  }
  for (int o = 16; o; o >>= 1) {
    double t;
    t = __shfl_xor_sync(FULL, l_max, o); if (t > l_max) l_max = t;
    t = __shfl_xor_sync(FULL, l_min, o); if (t < l_min) l_min = t;
    t = __shfl_xor_sync(FULL, w_max, o); if (t > w_max) w_max = t;
    t = __shfl_xor_sync(FULL, w_min, o); if (t < w_min) w_min = t;
  }
  rec->x1 = x + l_min * dx; rec->y1 = y + l_min * dy;
  rec->x2 = x + l_max * dx; rec->y2 = y + l_max * dy;
  rec->width = w_max - w_min;
  rec->x = x; rec->y = y; rec->theta = theta; rec->dx = dx; rec->dy = dy; rec->prec = prec; rec->p = p;
  if (rec->width < 1.0) rec->width = 1.0;
}

// ---------------------------------------------------- reduce_region_radius ----
// lsd.cpp:1775-1841. The swap-with-last compaction decides the order of reg[] for the following
// ordered sums; lane 0 replays it verbatim (rare path: a few hundred calls per frame).
__device__ bool reduce_region_radius(const FrameView& V, int* np, double reg_angle, double prec, double p, Rect* rec,
                                     double density_th) {
  const int lane = threadIdx.x & 31;
  int n = *np;
  double density = (double)n / (dist2d(rec->x1, rec->y1, rec->x2, rec->y2) * rec->width);
  if (density >= density_th) return true;
  int p0 = V.reg[0];
  double xc = (double)(p0 & 0xffff), yc = (double)(p0 >> 16);
  double rad1 = dist2d(xc, yc, rec->x1, rec->y1), rad2 = dist2d(xc, yc, rec->x2, rec->y2);
  double rad = rad1 > rad2 ? rad1 : rad2;
  while (density < density_th) {
    rad *= 0.75;
    if (lane == 0) {
      for (int i = 0; i < n; ++i) {
        int pi = V.reg[i];
        int px = pi & 0xffff, py = pi >> 16;
        if (dist2d(xc, yc, (double)px, (double)py) > rad) {
          V.used[px + py * V.xs] = 0;
          V.reg[i] = V.reg[n - 1];
          --n;
          --i;
        }
      }
    }
    n = __shfl_sync(FULL, n, 0);
    __syncwarp();
    if (n < 2) { *np = n; return false; }
    region2rect(V, n, reg_angle, prec, p, rec);
    density = (double)n / (dist2d(rec->x1, rec->y1, rec->x2, rec->y2) * rec->width);
  }
  *np = n;
  return true;
}

// ------------------------------------------------------------------ refine ----
// lsd.cpp:1853-1921
__device__ bool refine(const FrameView& V, int* np, double prec, double p, Rect* rec, double density_th) {
  const int lane = threadIdx.x & 31;
  int n = *np;
  double density = (double)n / (dist2d(rec->x1, rec->y1, rec->x2, rec->y2) * rec->width);
  if (density >= density_th) return true;
  int p0 = V.reg[0];
  int x0 = p0 & 0xffff, y0 = p0 >> 16;
  double xc = (double)x0, yc = (double)y0;
  double ang_c = V.angles[x0 + y0 * V.xs];
  double sum = 0.0, s_sum = 0.0;
  int cntn = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    int i = i0 + lane;
    bool in = false;
    double ang_d = 0.0;
    if (i < n) {
      int pi = V.reg[i];
      int px = pi & 0xffff, py = pi >> 16;
      V.used[px + py * V.xs] = 0;
      if (dist2d(xc, yc, (double)px, (double)py) < rec->width) {
        in = true;
        ang_d = angle_diff_signed(V.angles[px + py * V.xs], ang_c);
      }
    }
    unsigned m = __ballot_sync(FULL, in);
    cntn += __popc(m);
    while (m) {
      int k = __ffs(m) - 1;
      m &= m - 1;
      double d = shfl_d(ang_d, k);
      sum += d;
      s_sum += d * d;
    }
  }
  __syncwarp();
  double mean_angle = sum / (double)cntn;
  double tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)cntn + mean_angle * mean_angle);
  double C, S, a0;
  // guard-banded fast test is valid while the tolerance stays well inside (0, pi/2)
  bool fast_ok = tau > 1e-6 && tau < 1.4;
  double st, ct;
  lsl_sincos(fast_ok ? tau : 0.5, &st, &ct);
  n = region_grow(V, p0, tau, st / ct, fast_ok, &C, &S, &a0);
  *np = n;
  if (n < 2) return false;
  double reg_angle = lsl_atan2(S, C);
  region2rect(V, n, reg_angle, prec, p, rec);
  density = (double)n / (dist2d(rec->x1, rec->y1, rec->x2, rec->y2) * rec->width);
  if (density < density_th) return reduce_region_radius(V, np, reg_angle, prec, p, rec, density_th);
  return true;
}

// --------------------------------------------------------------------- nfa ----
// log_gamma of three arguments at once (lsd.cpp:886-931). Group g = lane/8 (g < 3) owns argument
// X[g]; lane j = lane%8 evaluates the j-th log / pow term, sums are chained in the reference order.
__device__ double log_gamma3(double X0, double X1, double X2, double* lg1, double* lg2) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 3, j = lane & 7;
  double X = g == 0 ? X0 : (g == 1 ? X1 : X2);
  const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705,
                       1168.92649479, 83.8676043424, 2.50662827511};
  bool big = X > 15.0;
  double t1 = 0.0, t2 = 0.0;
  if (g < 3) {
    if (big) {
      if (j == 0) t1 = lsl_log(X);
      else if (j == 1) t1 = lsl_sinh(1 / X);
      else if (j == 2) t1 = lsl_pow(X, 6.0);
    } else {
      if (j < 7) { t1 = lsl_log(X + (double)j); t2 = q[j] * lsl_pow(X, (double)j); }
      else t1 = lsl_log(X + 5.5);
    }
  }
  __syncwarp();
  // gather the group's terms (uniform shuffles, every lane assembles its own group's value)
  int b0 = (g < 3 ? g : 0) * 8;
  double r;
  double u0 = shfl_d(t1, b0 + 0), u1 = shfl_d(t1, b0 + 1), u2 = shfl_d(t1, b0 + 2), u3 = shfl_d(t1, b0 + 3),
         u4 = shfl_d(t1, b0 + 4), u5 = shfl_d(t1, b0 + 5), u6 = shfl_d(t1, b0 + 6), u7 = shfl_d(t1, b0 + 7);
  double v0 = shfl_d(t2, b0 + 0), v1 = shfl_d(t2, b0 + 1), v2 = shfl_d(t2, b0 + 2), v3 = shfl_d(t2, b0 + 3),
         v4 = shfl_d(t2, b0 + 4), v5 = shfl_d(t2, b0 + 5), v6 = shfl_d(t2, b0 + 6);
  double inner;
  if (big) inner = X * u1 + 1 / (810.0 * u2);
  else {
    double b = 0.0;
    b += v0; b += v1; b += v2; b += v3; b += v4; b += v5; b += v6;
    inner = b;
  }
  double li = lsl_log(inner);
  if (big) r = 0.918938533204673 + (X - 0.5) * u0 - X + 0.5 * X * li;
  else {
    double a = (X + 0.5) * u7 - (X + 5.5);
    a -= u0; a -= u1; a -= u2; a -= u3; a -= u4; a -= u5; a -= u6;
    r = a + li;
  }
  __syncwarp();
  double r0 = shfl_d(r, 0);
  *lg1 = shfl_d(r, 8);
  *lg2 = shfl_d(r, 16);
  return r0;
}

// lsd.cpp:980-1065
__device__ double nfa(int n, int k, double p, double logNT) {
  const int lane = threadIdx.x & 31;
  const double tolerance = 0.1;
  if (n == 0 || k == 0) return -logNT;
  if (n == k) return -logNT - (double)n * lsl_log10(p);
  double p_term = p / (1.0 - p);
  double lgk, lgnk;
  double lgn = log_gamma3((double)n + 1.0, (double)k + 1.0, (double)(n - k) + 1.0, &lgk, &lgnk);
  double lp = lane == 0 ? lsl_log(p) : (lane == 1 ? lsl_log(1.0 - p) : 0.0);
  double logp = shfl_d(lp, 0), log1p_ = shfl_d(lp, 1);
  double log1term = lgn - lgk - lgnk + (double)k * logp + (double)(n - k) * log1p_;
  double term = lsl_exp(log1term);
  if (dbl_equal(term, 0.0)) {
    if ((double)k > (double)n * p) return -log1term / LSL_LN10 - logNT;
    return -logNT;
  }
  double bin_tail = term;
  for (int i0 = k + 1; i0 <= n; i0 += 32) {
    double my_term = 0.0, my_tail = 0.0, my_mult = 0.0, my_bin = 2.0;
    int cnt = min(32, n - i0 + 1);
    for (int jj = 0; jj < cnt; ++jj) {
      int i = i0 + jj;
      double bin_term = (double)(n - i + 1) * (1.0 / (double)i);
      double mult_term = bin_term * p_term;
      term *= mult_term;
      bin_tail += term;
      if (lane == jj) { my_term = term; my_tail = bin_tail; my_mult = mult_term; my_bin = bin_term; }
    }
    bool brk = false;
    if (lane < cnt && my_bin < 1.0) {
      int i = i0 + lane;
      double err = my_term * ((1.0 - lsl_pow(my_mult, (double)(n - i + 1))) / (1.0 - my_mult) - 1.0);
      brk = err < tolerance * fabs(-lsl_log10(my_tail) - logNT) * my_tail;
    }
    unsigned m = __ballot_sync(FULL, brk);
    if (m) {
      bin_tail = shfl_d(my_tail, __ffs(m) - 1);
      break;
    }
  }
  return -lsl_log10(bin_tail) - logNT;
}

// --------------------------------------------------------------- rect_nfa ----
__device__ __forceinline__ double inter_low(double x, double x1, double y1, double x2, double y2) {
  if (dbl_equal(x1, x2) && y1 < y2) return y1;
  if (dbl_equal(x1, x2) && y1 > y2) return y2;
  return y1 + (x - x1) * (y2 - y1) / (x2 - x1);
}
__device__ __forceinline__ double inter_hi(double x, double x1, double y1, double x2, double y2) {
  if (dbl_equal(x1, x2) && y1 < y2) return y2;
  if (dbl_equal(x1, x2) && y1 > y2) return y1;
  return y1 + (x - x1) * (y2 - y1) / (x2 - x1);
}
// lsd.cpp:1388-1410 with the rectangle iterator (:1165-1383) unrolled into per-lane columns.
// NOT inlined: rect_improve calls it from six places (21 after unrolling) and the kernel was instruction-fetch bound
// (ncu: 40 % of the stall samples `no_instructions` at 19 288 SASS instructions).
__device__ __noinline__ double rect_nfa(const FrameView& V, const Rect& r, double logNT) {
  const int lane = threadIdx.x & 31;
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
#pragma unroll
  for (int q = 0; q < 4; ++q) { vx[q] = vx0[(offset + q) & 3]; vy[q] = vy0[(offset + q) & 3]; }
  int pts = 0, alg = 0;
  int xbeg = (int)ceil(vx[0]);
  for (int x = xbeg + lane; (double)x <= vx[2]; x += 32) {
    double ys, ye;
    if ((double)x < vx[3]) ys = inter_low((double)x, vx[0], vy[0], vx[3], vy[3]);
    else ys = inter_low((double)x, vx[3], vy[3], vx[2], vy[2]);
    if ((double)x < vx[1]) ye = inter_hi((double)x, vx[0], vy[0], vx[1], vy[1]);
    else ye = inter_hi((double)x, vx[1], vy[1], vx[2], vy[2]);
    if (x < 0 || x >= V.xs) continue;
    for (int y = (int)ceil(ys); (double)y <= ye; ++y)
      if (y >= 0 && y < V.ys) {
        ++pts;
        if (isaligned_exact(V.angles[x + y * V.xs], r.theta, r.prec)) ++alg;
      }
  }
  for (int o = 16; o; o >>= 1) { pts += __shfl_xor_sync(FULL, pts, o); alg += __shfl_xor_sync(FULL, alg, o); }
  return nfa(pts, alg, r.p, logNT);
}

// lsd.cpp:1662-1768
__device__ double rect_improve(const FrameView& V, Rect* rec, double logNT, double eps) {
  Rect r;
  const double delta = 0.5, delta_2 = delta / 2.0;
  double log_nfa = rect_nfa(V, *rec, logNT), log_nfa_new;
  if (log_nfa > eps) return log_nfa;
  r = *rec;
#pragma unroll 1
  for (int n = 0; n < 5; ++n) {
    r.p /= 2.0; r.prec = r.p * LSL_PI;
    log_nfa_new = rect_nfa(V, r, logNT);
    if (log_nfa_new > log_nfa) { log_nfa = log_nfa_new; *rec = r; }
  }
  if (log_nfa > eps) return log_nfa;
  r = *rec;
#pragma unroll 1
  for (int n = 0; n < 5; ++n)
    if ((r.width - delta) >= 0.5) {
      r.width -= delta;
      log_nfa_new = rect_nfa(V, r, logNT);
      if (log_nfa_new > log_nfa) { *rec = r; log_nfa = log_nfa_new; }
    }
  if (log_nfa > eps) return log_nfa;
  r = *rec;
#pragma unroll 1
  for (int n = 0; n < 5; ++n)
    if ((r.width - delta) >= 0.5) {
      r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2;
      r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2;
      r.width -= delta;
      log_nfa_new = rect_nfa(V, r, logNT);
      if (log_nfa_new > log_nfa) { *rec = r; log_nfa = log_nfa_new; }
    }
  if (log_nfa > eps) return log_nfa;
  r = *rec;
#pragma unroll 1
  for (int n = 0; n < 5; ++n)
    if ((r.width - delta) >= 0.5) {
      r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2;
      r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2;
      r.width -= delta;
      log_nfa_new = rect_nfa(V, r, logNT);
      if (log_nfa_new > log_nfa) { *rec = r; log_nfa = log_nfa_new; }
    }
  if (log_nfa > eps) return log_nfa;
  r = *rec;
#pragma unroll 1
  for (int n = 0; n < 5; ++n) {
    r.p /= 2.0; r.prec = r.p * LSL_PI;
    log_nfa_new = rect_nfa(V, r, logNT);
    if (log_nfa_new > log_nfa) { log_nfa = log_nfa_new; *rec = r; }
  }
  return log_nfa;
}

// ------------------------------------------------------------- frame loop ----
// LineSegmentDetection main loop (lsd.cpp:1995-2054), split in three kernels:
//   lsd_region_kernel   the order-dependent part (seed walk, region_grow, region2rect, refine: these read and
//                       write `used`), one warp per frame; every region that survives refine() leaves its
//                       rectangle in rects[] in sequence order;
//   lsd_nfa_kernel      rect_improve + NFA validation (lsd.cpp:2024-2031) — no side effects on `used`, so all
//                       candidate rectangles of all frames are validated in parallel, one warp each;
//   lsd_compact_kernel  accepted rectangles -> output rows in sequence order (add_5tuple order, lsd.cpp:2033-2046).
__global__ void __launch_bounds__(32) lsd_region_kernel(LslWork w, int xs, int ys, double ang_th, double density_th, int min_reg_size) {
  const int f = blockIdx.x, lane = threadIdx.x;
  const size_t po = (size_t)f * xs * ys;
  FrameView V;
  V.xs = xs; V.ys = ys;
  V.angles = w.angles + po; V.modgrad = w.modgrad + po; V.cs = w.cs + po;
  V.used = w.used + po; V.reg = w.reg + po;
  const int32_t* seeds = w.seeds + po;
  const int nseeds = w.nseeds[f];
  double* rects = w.rects + (size_t)f * LSL_MAX_RECTS * 12;
  const double prec = LSL_PI * ang_th / 180.0;
  const double p = ang_th / 180.0;
  const bool fast_ok = prec > 1e-6 && prec < 1.4;
  double sp_, cp_;
  lsl_sincos(fast_ok ? prec : 0.5, &sp_, &cp_);
  const double tanp = sp_ / cp_;
  int nout = 0;
  for (int s0 = 0; s0 < nseeds; s0 += 32) {
    int my = (s0 + lane < nseeds) ? seeds[s0 + lane] : -1;
    bool mu = true;
    if (my >= 0) mu = V.used[(my & 0xffff) + (my >> 16) * xs] != 0;  // a stale "used" stays true; "free" is re-read below
    unsigned freem = __ballot_sync(FULL, !mu);
    while (freem) {
      int j = __ffs(freem) - 1;
      freem &= freem - 1;
      int seed = __shfl_sync(FULL, my, j);
      int sidx = (seed & 0xffff) + (seed >> 16) * xs;
      if (V.used[sidx] != 0) continue;
      double C, S, a0;
      int n = region_grow(V, seed, prec, tanp, fast_ok, &C, &S, &a0);
      if (n < min_reg_size) continue;
      double reg_angle = lsl_atan2(S, C);
      Rect rec;
      region2rect(V, n, reg_angle, prec, p, &rec);
      if (!refine(V, &n, prec, p, &rec, density_th)) continue;
      if (nout < LSL_MAX_RECTS) {
        const double v[12] = {rec.x1, rec.y1, rec.x2, rec.y2, rec.width, rec.x, rec.y, rec.theta, rec.dx, rec.dy, rec.prec, rec.p};
        double mine = v[0];
#pragma unroll
        for (int k = 1; k < 12; ++k) if (lane == k) mine = v[k];
        if (lane < 12) rects[(size_t)nout * 12 + lane] = mine;
      }
      ++nout;
    }
  }
  if (lane == 0) w.nrects[f] = nout;
}

#ifndef NFA_MINB
#define NFA_MINB 8    // 64 registers instead of 126: 32 instead of 16 warps per SM (per 592 frames: 4.66 -> 4.49 / 4.21 ms at 6 / 8 CTAs per SM)
#endif
#ifndef NFA_WARPS
#define NFA_WARPS 4   // warps (= candidate rectangles in flight) per CTA
#endif
#define NFA_BOUNDS __launch_bounds__(32 * NFA_WARPS, NFA_MINB * 4 / NFA_WARPS)
__global__ void NFA_BOUNDS lsd_nfa_kernel(LslWork w, int xs, int ys, double eps, double scale, double logNT) {
  const int f = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t po = (size_t)f * xs * ys;
  FrameView V;
  V.xs = xs; V.ys = ys;
  V.angles = w.angles + po; V.modgrad = w.modgrad + po; V.cs = w.cs + po;
  V.used = w.used + po; V.reg = w.reg + po;
  const int nr = min(w.nrects[f], LSL_MAX_RECTS);
  double* rects = w.rects + (size_t)f * LSL_MAX_RECTS * 12;
  uint8_t* ok = w.rect_ok + (size_t)f * LSL_MAX_RECTS;
  for (int c = blockIdx.x * NFA_WARPS + warp; c < nr; c += gridDim.x * NFA_WARPS) {
    double* r = rects + (size_t)c * 12;
    Rect rec;
    rec.x1 = r[0]; rec.y1 = r[1]; rec.x2 = r[2]; rec.y2 = r[3]; rec.width = r[4]; rec.x = r[5]; rec.y = r[6];
    rec.theta = r[7]; rec.dx = r[8]; rec.dy = r[9]; rec.prec = r[10]; rec.p = r[11];
    __syncwarp();
    double log_nfa = rect_improve(V, &rec, logNT, eps);
    const bool acc = log_nfa > eps;
    if (acc) {
      rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
      if (scale != 1.0) {
        rec.x1 /= scale; rec.y1 /= scale; rec.x2 /= scale; rec.y2 /= scale;
        rec.width /= scale;
      }
      if (lane == 0) { r[0] = rec.x1; r[1] = rec.y1; r[2] = rec.x2; r[3] = rec.y2; r[4] = rec.width; }
    }
    if (lane == 0) ok[c] = acc ? 1 : 0;
  }
}

__global__ void __launch_bounds__(32) lsd_compact_kernel(LslWork w) {
  const int f = blockIdx.x, lane = threadIdx.x;
  const int nraw = w.nrects[f];
  const int nr = min(nraw, LSL_MAX_RECTS);
  const double* rects = w.rects + (size_t)f * LSL_MAX_RECTS * 12;
  const uint8_t* ok = w.rect_ok + (size_t)f * LSL_MAX_RECTS;
  double* out = w.segs + (size_t)f * LSL_MAX_SEGS * 5;
  int nout = 0;
  for (int c0 = 0; c0 < nr; c0 += 32) {
    const int c = c0 + lane;
    const bool a = c < nr && ok[c] != 0;
    const unsigned m = __ballot_sync(FULL, a);
    const int pos = nout + __popc(m & ((1u << lane) - 1u));
    if (a && pos < LSL_MAX_SEGS) {
#pragma unroll
      for (int k = 0; k < 5; ++k) out[(size_t)pos * 5 + k] = rects[(size_t)c * 12 + k];
    }
    nout += __popc(m);
  }
  if (lane == 0) w.nsegs[f] = nraw > LSL_MAX_RECTS ? LSL_MAX_SEGS + 1 : nout;   // overflow -> capacity error on the host
}

int lsl_launch_lsd(lsl_ctx* ctx, int n) {
  const LslDims& d = ctx->dims;
  const lsl_params& P = ctx->P;
  LslWork& w = ctx->wk;
  LSL_CUDA(cudaMemsetAsync(w.used, 0, (size_t)n * d.sw * d.sh, ctx->stream));
  double p = P.lsd_ang_th / 180.0;
  double logNT = 5.0 * (lsl_log10((double)d.sw) + lsl_log10((double)d.sh)) / 2.0;
  int min_reg_size = (int)(-logNT / lsl_log10(p));
  LSL_KSTART(ctx, LSL_K_REGION);
  lsd_region_kernel<<<n, 32, 0, ctx->stream>>>(w, d.sw, d.sh, P.lsd_ang_th, P.lsd_density_th, min_reg_size);
  LSL_KSTOP(ctx, LSL_K_REGION);
  LSL_KSTART(ctx, LSL_K_NFA);
  dim3 g(256 / NFA_WARPS, n);
  lsd_nfa_kernel<<<g, 32 * NFA_WARPS, 0, ctx->stream>>>(w, d.sw, d.sh, P.lsd_eps, P.lsd_scale, logNT);
  lsd_compact_kernel<<<n, 32, 0, ctx->stream>>>(w);
  ctx->stats.kernel_launches += 1;   // two launches inside one timing bracket
  LSL_KSTOP(ctx, LSL_K_NFA);
  LSL_CUDA(cudaGetLastError());
  return LSL_OK;
}
