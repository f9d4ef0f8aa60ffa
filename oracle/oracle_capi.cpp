// oracle_capi.cpp — extern "C" surface of the CPU oracle for ctypes (TEST INFRASTRUCTURE).
// Shares the POD layouts of include/lsl.h so tests can compare records byte for byte.
#include "oracle.h"
#include "../include/lsl.h"
#include <string.h>
#include <chrono>
#include <omp.h>
#include "../lineslam_b200/csrc/shared/lsl_math.h"
#include "../lineslam_b200/csrc/shared/lsl_params_default.h"
#include "../lineslam_b200/csrc/shared/lsl_points.h"
#include "../lineslam_b200/csrc/shared/lsl_linalg.h"
#include "../lineslam_b200/csrc/shared/lsl_cvdraw.h"

static orc::Params toP(const lsl_params* p) {
  orc::Params P;
  if (!p) return P;
  P.lsd_scale = p->lsd_scale; P.lsd_sigma_scale = p->lsd_sigma_scale; P.lsd_quant = p->lsd_quant;
  P.lsd_ang_th = p->lsd_ang_th; P.lsd_eps = p->lsd_eps; P.lsd_density_th = p->lsd_density_th;
  P.lsd_max_grad = p->lsd_max_grad; P.lsd_n_bins = p->lsd_n_bins;
  P.line_2d_len_thres = p->line_2d_len_thres; P.msld_sample_interval = p->msld_sample_interval;
  P.line_3d_len_thres_m = p->line_3d_len_thres_m; P.collin_pts_ratio = p->collin_pts_ratio;
  P.line_sample_interval = p->line_sample_interval; P.line_sample_max_num = p->line_sample_max_num;
  P.line_sample_min_num = p->line_sample_min_num; P.line3d_mle_iter_num = p->line3d_mle_iter_num;
  P.pt2line_mahdist_extractline = p->pt2line_mahdist_extractline;
  P.ransac_iters_extract_line = p->ransac_iters_extract_line;
  P.num_cells_lineseg_range = p->num_cells_lineseg_range;
  P.ratio_support_pts_on_line = p->ratio_support_pts_on_line;
  P.stdev_sample_pt_imgline = p->stdev_sample_pt_imgline;
  P.depth_stdev_coeff_c1 = p->depth_stdev_coeff_c1; P.depth_stdev_coeff_c2 = p->depth_stdev_coeff_c2;
  P.depth_stdev_coeff_c3 = p->depth_stdev_coeff_c3; P.depth_scaling = p->depth_scaling;
  P.ransac_iters_line_motion = p->ransac_iters_line_motion;
  P.adjacent_linematch_window = p->adjacent_linematch_window;
  P.line_match_number_weight = p->line_match_number_weight;
  P.min_feature_matches = p->min_feature_matches; P.min_matches_loopclose = p->min_matches_loopclose;
  P.max_mah_dist_for_inliers = p->max_mah_dist_for_inliers; P.g2o_line_error_weight = p->g2o_line_error_weight;
  P.g2o_BA_kernel_delta = p->g2o_BA_kernel_delta; P.g2o_BA_use_kernel = p->g2o_BA_use_kernel;
  P.pt2line3d_dist_relmotion = p->pt2line3d_dist_relmotion; P.line3d_angle_relmotion = p->line3d_angle_relmotion;
  P.sigma_depth = p->sigma_depth; P.nn_distance_ratio = p->nn_distance_ratio;
  return P;
}
static_assert(sizeof(orc::Line) == sizeof(lsl_line_rec), "record layouts must agree");
static_assert(sizeof(orc::Match) == sizeof(lsl_match), "match layouts must agree");

extern "C" {

int orc_nproc() { return omp_get_max_threads(); }
void orc_params_default(lsl_params* p) { lsl_params_default_impl(p); }

// math probes (tests/test_math.py)
double orc_m_exp(double x) { return lslm::lsl_exp(x); }
double orc_m_log(double x) { return lslm::lsl_log(x); }
double orc_m_log10(double x) { return lslm::lsl_log10(x); }
double orc_m_sin(double x) { return lslm::lsl_sin(x); }
double orc_m_cos(double x) { return lslm::lsl_cos(x); }
double orc_m_atan2(double y, double x) { return lslm::lsl_atan2(y, x); }
double orc_m_pow(double x, double y) { return lslm::lsl_pow(x, y); }
double orc_m_sinh(double x) { return lslm::lsl_sinh(x); }

void orc_rand(uint32_t seed, int n, int32_t* out) {
  orc::GlibcRand r; r.seed(seed);
  for (int i = 0; i < n; ++i) out[i] = r.next();
}

void orc_gray(const uint8_t* img, int W, int H, uint8_t* gray) { orc::gray_from_3ch(img, W, H, gray); }

// LSD; optional intermediates (pass NULL to skip). Returns number of segments.
int orc_lsd(const uint8_t* gray, int W, int H, const lsl_params* p, double* segs, int cap, double* scaled,
            double* angles, double* modgrad, int32_t* seeds, int* nseeds, int* sw, int* sh) {
  orc::Params P = toP(p);
  std::vector<orc::Segment> s;
  orc::LsdDebug dbg;
  orc::lsd_detect(gray, W, H, P, s, &dbg);
  int n = (int)s.size();
  for (int i = 0; i < n && i < cap; ++i) memcpy(segs + 5 * i, &s[i], 40);
  if (scaled) memcpy(scaled, dbg.scaled.data(), dbg.scaled.size() * 8);
  if (angles) memcpy(angles, dbg.angles.data(), dbg.angles.size() * 8);
  if (modgrad) memcpy(modgrad, dbg.modgrad.data(), dbg.modgrad.size() * 8);
  if (seeds) memcpy(seeds, dbg.seeds.data(), dbg.seeds.size() * 4);
  if (nseeds) *nseeds = (int)dbg.seeds.size();
  if (sw) *sw = dbg.sw;
  if (sh) *sh = dbg.sh;
  return n;
}

void orc_sobel5(const uint8_t* gray, int W, int H, double* gx, double* gy) {
  std::vector<double> a, b;
  orc::sobel5(gray, W, H, a, b);
  memcpy(gx, a.data(), a.size() * 8);
  memcpy(gy, b.data(), b.size() * 8);
}

// detect3DLines. Debug outputs optional: seg_of_line[cap], n_inl[cap], inl_idx[cap*101], a0b0[cap*6], lm_iters[cap]
int orc_detect3DLines(const uint8_t* img, int channels, const float* depth, int W, int H, const double* K, double dt,
                      uint32_t seed, const lsl_params* p, lsl_line_rec* out, int cap, int omp_threads,
                      int* seg_of_line, int* n_inl, int* inl_idx, double* a0b0, int* lm_iters, int* nsegs,
                      double* segs, int segcap) {
  orc::Params P = toP(p);
  std::vector<uint8_t> gray;
  const uint8_t* g = img;
  if (channels == 3) { gray.resize((size_t)W * H); orc::gray_from_3ch(img, W, H, gray.data()); g = gray.data(); }
  std::vector<orc::Line> lines;
  orc::ExtractDebug dbg;
  orc::detect3DLines(g, depth, W, H, K, dt, seed, P, lines, &dbg, omp_threads);
  int n = (int)lines.size();
  for (int i = 0; i < n && i < cap; ++i) {
    memcpy(&out[i], &lines[i], sizeof(lsl_line_rec));
    if (seg_of_line) seg_of_line[i] = dbg.seg_of_line[i];
    if (n_inl) n_inl[i] = (int)dbg.inlier_idx[i].size();
    if (inl_idx) for (size_t k = 0; k < dbg.inlier_idx[i].size() && k < 101; ++k) inl_idx[i * 101 + k] = dbg.inlier_idx[i][k];
    if (a0b0) memcpy(a0b0 + 6 * i, &dbg.A0B0[6 * i], 48);
    if (lm_iters) lm_iters[i] = dbg.lm_iters[i];
  }
  if (nsegs) *nsegs = (int)dbg.segs.size();
  if (segs) for (int i = 0; i < (int)dbg.segs.size() && i < segcap; ++i) memcpy(segs + 5 * i, &dbg.segs[i], 40);
  return n;
}

int orc_lineMatching(const lsl_line_rec* f1, int n1, const lsl_line_rec* f2, int n2, int adjacent, lsl_match* out,
                     int cap, int omp_threads) {
  std::vector<orc::Line> a(n1), b(n2);
  if (n1) memcpy(a.data(), f1, sizeof(lsl_line_rec) * n1);
  if (n2) memcpy(b.data(), f2, sizeof(lsl_line_rec) * n2);
  std::vector<orc::Match> m;
  orc::lineMatching(a, b, adjacent != 0, m, omp_threads);
  for (int i = 0; i < (int)m.size() && i < cap; ++i) memcpy(&out[i], &m[i], sizeof(lsl_match));
  return (int)m.size();
}

int orc_pose_ransac(const lsl_line_rec* train, int ntrain, const lsl_line_rec* query, int nquery, int id_train,
                    int id_query, const lsl_match* ms, int nm, uint32_t seed, const lsl_params* p, lsl_pose_rec* rec,
                    lsl_match* inl, int cap, int* n_inl, lsl_match* rinl, int cap2, int* n_rinl, float* tf_ransac) {
  orc::Params P = toP(p);
  std::vector<orc::Line> t(ntrain), q(nquery);
  if (ntrain) memcpy(t.data(), train, sizeof(lsl_line_rec) * ntrain);
  if (nquery) memcpy(q.data(), query, sizeof(lsl_line_rec) * nquery);
  std::vector<orc::Match> m(nm);
  if (nm) memcpy(m.data(), ms, sizeof(lsl_match) * nm);
  orc::PoseResult r;
  orc::getTransform_Lines_ransac(t, q, id_train, id_query, m, seed, P, r);
  memset(rec, 0, sizeof(*rec));
  rec->id_train = id_train; rec->id_query = id_query; rec->found = r.found ? 1 : 0;
  rec->n_line_matches = nm; rec->n_ransac_inliers = (int)r.ransac_inliers.size(); rec->n_inliers = (int)r.inliers.size();
  rec->rmse = r.rmse; rec->best_iter = r.best_iter;
  memcpy(rec->tf, r.tf, 64);
  if (tf_ransac) memcpy(tf_ransac, r.tf_ransac, 64);
  if (n_inl) *n_inl = (int)r.inliers.size();
  if (n_rinl) *n_rinl = (int)r.ransac_inliers.size();
  for (int i = 0; i < (int)r.inliers.size() && i < cap; ++i) memcpy(&inl[i], &r.inliers[i], sizeof(lsl_match));
  for (int i = 0; i < (int)r.ransac_inliers.size() && i < cap2; ++i) memcpy(&rinl[i], &r.ransac_inliers[i], sizeof(lsl_match));
  return 0;
}

// Node::featureMatching (BRUTEFORCE). Returns the number of matches; *draws = rand() calls consumed.
int orc_featureMatching(const float* qdesc, int nq, const float* tdesc, int nt, int dim, double nn_ratio, uint32_t seed,
                        lsl_match* out, int cap) {
  orc::Points q, t;
  q.n = nq; q.dim = dim; q.desc = qdesc; t.n = nt; t.dim = dim; t.desc = tdesc;
  orc::GlibcRand rng; rng.seed(seed);
  std::vector<orc::Match> m;
  orc::featureMatching(q, t, nn_ratio, rng, m);
  for (int i = 0; i < (int)m.size() && i < cap; ++i) memcpy(&out[i], &m[i], sizeof(lsl_match));
  return (int)m.size();
}
void orc_rootsift(float* desc, int n, int dim) { orc::rootsift(desc, n, dim); }
int orc_featureMatching_hamming(const uint8_t* qd, int nq, const uint8_t* td, int nt, int nbytes, double nn_ratio, uint32_t seed,
                                lsl_match* out, int cap) {
  orc::GlibcRand r; r.seed(seed);
  std::vector<orc::Match> m;
  orc::featureMatching_hamming(qd, nq, td, nt, nbytes, nn_ratio, r, m);
  for (int i = 0; i < (int)m.size() && i < cap; ++i) memcpy(&out[i], &m[i], sizeof(lsl_match));
  return (int)m.size();
}

// getTransform_PtsLines_ransac with point + line matches. skip_draws: rand() calls already consumed from the
// seed's stream (the featureMatching jitter of the same matchNodePair call). pad[0] = #point matches,
// pad[1] = #point inliers of the best hypothesis, pad[2] = #refined point inliers.
int orc_pose_ransac_hybrid(const lsl_line_rec* train, int ntrain, const lsl_line_rec* query, int nquery,
                           const float* train_xyz1, int ntp, const float* query_xyz1, int nqp, int id_train, int id_query,
                           const lsl_match* pms, int npm, const lsl_match* ms, int nm, uint32_t seed, int skip_draws,
                           double fx, double dt, const lsl_params* p, lsl_pose_rec* rec, lsl_match* inl, int* n_inl,
                           lsl_match* rinl, int* n_rinl, lsl_match* pinl, int* n_pinl, lsl_match* prinl, int* n_prinl,
                           float* tf_ransac) {
  orc::Params P = toP(p);
  std::vector<orc::Line> t(ntrain), q(nquery);
  if (ntrain) memcpy(t.data(), train, sizeof(lsl_line_rec) * ntrain);
  if (nquery) memcpy(q.data(), query, sizeof(lsl_line_rec) * nquery);
  std::vector<orc::Match> m(nm), pm(npm);
  if (nm) memcpy(m.data(), ms, sizeof(lsl_match) * nm);
  if (npm) memcpy(pm.data(), pms, sizeof(lsl_match) * npm);
  orc::Points tp, qp;
  tp.n = ntp; tp.xyz1 = train_xyz1; qp.n = nqp; qp.xyz1 = query_xyz1;
  orc::GlibcRand rng; rng.seed(seed);
  for (int i = 0; i < skip_draws; ++i) rng.next();
  orc::PoseResult r;
  orc::getTransform_PtsLines_ransac(t, q, tp, qp, id_train, id_query, pm, m, rng, fx, dt, P, r);
  memset(rec, 0, sizeof(*rec));
  rec->id_train = id_train; rec->id_query = id_query; rec->found = r.found ? 1 : 0;
  rec->n_line_matches = nm; rec->n_ransac_inliers = (int)r.ransac_inliers.size(); rec->n_inliers = (int)r.inliers.size();
  rec->rmse = r.rmse; rec->best_iter = r.best_iter;
  rec->pad[0] = npm; rec->pad[1] = (int)r.pt_ransac_inliers.size(); rec->pad[2] = (int)r.pt_inliers.size();
  memcpy(rec->tf, r.tf, 64);
  if (tf_ransac) memcpy(tf_ransac, r.tf_ransac, 64);
  *n_inl = (int)r.inliers.size(); *n_rinl = (int)r.ransac_inliers.size();
  *n_pinl = (int)r.pt_inliers.size(); *n_prinl = (int)r.pt_ransac_inliers.size();
  for (size_t i = 0; i < r.inliers.size(); ++i) memcpy(&inl[i], &r.inliers[i], sizeof(lsl_match));
  for (size_t i = 0; i < r.ransac_inliers.size(); ++i) memcpy(&rinl[i], &r.ransac_inliers[i], sizeof(lsl_match));
  for (size_t i = 0; i < r.pt_inliers.size(); ++i) memcpy(&pinl[i], &r.pt_inliers[i], sizeof(lsl_match));
  for (size_t i = 0; i < r.pt_ransac_inliers.size(); ++i) memcpy(&prinl[i], &r.pt_ransac_inliers[i], sizeof(lsl_match));
  return 0;
}

// point math probes (tests/test_oracle_points.py)
double orc_error_function2(const float* x1, const float* x2, const float* tf, double sigma_depth) {
  double tfd[16];
  for (int i = 0; i < 16; ++i) tfd[i] = (double)tf[i];
  return lslm::error_function2(x1, x2, tfd, sigma_depth);
}
void orc_kabsch(const float* from, const float* to, const float* w, int n, float* tf) {
  lslm::Tfc t; lslm::tfc_reset(&t);
  for (int i = 0; i < n; ++i) lslm::tfc_add(&t, from + 3 * i, to + 3 * i, w[i]);
  lslm::tfc_get(&t, tf);
}
void orc_ldlt3_solve(const double* A, const double* b, double* x) { lslm::ldlt3_solve(A, b, x); }

// computeRelativeMotion_Ransac (motion.cpp:367-526) on paired lines a[i] <-> b[i]; Rt = R (9) then t (3).
int orc_relmotion_ransac(const lsl_line_rec* a, const lsl_line_rec* b, int n, uint32_t seed, const lsl_params* p,
                         double* Rt, int32_t* conset, int* lm_calls, int* have) {
  orc::Params P = toP(p);
  std::vector<orc::Line> va(n), vb(n);
  if (n) { memcpy(va.data(), a, sizeof(lsl_line_rec) * n); memcpy(vb.data(), b, sizeof(lsl_line_rec) * n); }
  orc::RelMotion r;
  orc::computeRelativeMotion_Ransac(va, vb, seed, P, r);
  if (r.have) { memcpy(Rt, r.R, 72); memcpy(Rt + 9, r.t, 24); }
  for (size_t i = 0; i < r.conset.size(); ++i) conset[i] = r.conset[i];
  if (lm_calls) *lm_calls = r.lm_calls;
  if (have) *have = r.have ? 1 : 0;
  return (int)r.conset.size();
}
void orc_optimizeRelmotion(const lsl_line_rec* a, const lsl_line_rec* b, int n, double* Rt) {
  std::vector<orc::Line> va(n), vb(n);
  if (n) { memcpy(va.data(), a, sizeof(lsl_line_rec) * n); memcpy(vb.data(), b, sizeof(lsl_line_rec) * n); }
  orc::optimizeRelmotion(va, vb, Rt, Rt + 9);
}
double orc_m_acos(double x) { return lslm::lsl_acos(x); }
// error-free product probe (Dekker splitting on the host; the device uses one FMA — tests/test_oracle_units.py)
void orc_m_two_prod(double a, double b, double* pe) { lslm::dd r = lslm::two_prod(a, b); pe[0] = r.h; pe[1] = r.l; }

// levmar restatement probe: Rosenbrock-like known-answer problems are driven from tests through this.
typedef void (*orc_lm_fn)(double*, double*, int, int, void*);
int orc_dlevmar_dif(orc_lm_fn f, double* p, double* x, int m, int n, int itmax, const double* opts, double* info) {
  return orc::dlevmar_dif_restated(f, p, x, m, n, itmax, opts, info, nullptr);
}

// The reference CPU path of one stream step, timed (bench.py cpu_baseline / --impl reference):
// extract frame `cur`, match against cached `prev` lines, RANSAC pose. Returns seconds.
double orc_stream_step(const uint8_t* img, int channels, const float* depth, int W, int H, const double* K,
                       uint32_t seed, const lsl_params* p, const lsl_line_rec* prev, int nprev, lsl_line_rec* cur,
                       int cap, int* ncur, lsl_pose_rec* rec, int omp_threads) {
  auto t0 = std::chrono::steady_clock::now();
  orc::Params P = toP(p);
  std::vector<uint8_t> gray;
  const uint8_t* g = img;
  if (channels == 3) { gray.resize((size_t)W * H); orc::gray_from_3ch(img, W, H, gray.data()); g = gray.data(); }
  std::vector<orc::Line> lines;
  orc::detect3DLines(g, depth, W, H, K, 0.0, seed, P, lines, nullptr, omp_threads);
  memset(rec, 0, sizeof(*rec));
  if (nprev > 0) {
    std::vector<orc::Line> pv(nprev);
    memcpy(pv.data(), prev, sizeof(lsl_line_rec) * nprev);
    std::vector<orc::Match> m;
    orc::lineMatching(lines, pv, true, m, omp_threads);
    orc::PoseResult r;
    if ((int)m.size() * P.line_match_number_weight >= P.min_feature_matches)
      orc::getTransform_Lines_ransac(pv, lines, 0, 1, m, seed, P, r);
    rec->found = r.found; rec->n_line_matches = (int)m.size(); rec->n_inliers = (int)r.inliers.size();
    rec->rmse = r.rmse; memcpy(rec->tf, r.tf, 64);
  }
  *ncur = (int)lines.size();
  for (int i = 0; i < (int)lines.size() && i < cap; ++i) memcpy(&cur[i], &lines[i], sizeof(lsl_line_rec));
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---- third-party restatement probes (tests/test_oracle_cv2.py compares them with cv2 4.13) ----
double orc_cvnorm(const double* v, int n) { return orc::cvnorm(v, n); }
double orc_cvnorm_diff72(const double* a, const double* b) { return orc::cvnorm_diff72(a, b); }
float orc_l2sqr_f(const float* a, const float* b, int n) { return lslm::l2sqr_f(a, b, n); }
// RandomPoint3d ctor (cv::SVD of the 3x3 covariance, src/line/lineslam.h:59-81)
void orc_cov_to_DU(const double* cov, double* DU, double* W_sqrt) { lslm::cov_to_DU(cov, DU, W_sqrt); }
// cv::Mat::inv() on 3x3 (src/line/motion.cpp:363, utils.cpp:1012) and 6x6 (utils.cpp:1044); returns 0 if singular
double orc_inv3(const double* a, double* r) { return lslm::inv3(a, r); }
int orc_inv6(const double* a, double* r) { double A[36]; memcpy(A, a, sizeof(A)); return lslm::inv_lu<6>(A, r); }
// cv::SVD of a symmetric 4x4 (Zhang's A in computeRelativeMotion_svd, motion.cpp:353): eigenvalues descending + vectors in columns
void orc_jacobi4(const double* a, double* w, double* V) { double A[16]; memcpy(A, a, sizeof(A)); lslm::jacobi_sym<4>(A, w, V); }
void orc_jacobi3(const double* a, double* w, double* V) { double A[9]; memcpy(A, a, sizeof(A)); lslm::jacobi_sym<3>(A, w, V); }

// cv::clipLine restatement probe; pts = x1 y1 x2 y2 in/out, returns the bool
int orc_clip_line(int W, int H, int* pts) { return lslm::clip_line(W, H, pts, pts + 1, pts + 2, pts + 3) ? 1 : 0; }
// FrameLine::getGradient probe: r (2) for the segment p -> q over the given f64 gradient planes
void orc_get_gradient(const double* xG, const double* yG, int W, int H, const double* pq, double* r) {
  orc::get_gradient_probe(xG, yG, W, H, pq, r);
}

// decision form of the Mahalanobis point-to-line test (shared/lsl_linalg.h) against the plain formula: both results
void orc_mah_lt(const double* pos, const double* DU, const double* q1, const double* q2, double thr, int* fast, int* plain) {
  *fast = lslm::mah_dist3d_pt_line_lt(pos, DU, q1, q2, thr) ? 1 : 0;
  *plain = (lslm::mah_dist3d_pt_line(pos, DU, q1, q2) < thr) ? 1 : 0;
}
double orc_mah_dist(const double* pos, const double* DU, const double* q1, const double* q2) { return lslm::mah_dist3d_pt_line(pos, DU, q1, q2); }

}  // extern "C"
