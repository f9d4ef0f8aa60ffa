// lsl_internal.h — context / frame / workspace layout of liblsl_b200 (product code, CUDA only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/lsl.h"

#define LSL_MAX_SEGS 2048      // LSD segments per frame (ntuple_list rows)
#define LSL_MAX_LINES 1024     // 3D lines kept per frame
#define LSL_MAX_RECTS 4096     // regions per frame that reach the NFA validation
#define LSL_MAX_SMP 101        // samples per line (line_sample_max_num + 1)
#define LSL_NOTDEF (-1024.0)   // external/lsd/lsd.cpp:102
#define LSL_MAX_MATCH 1024     // line matches per pair
#define LSL_MAX_POINTS 2048    // point features per frame (max_keypoints is 600 in the reference)

struct LslDims {
  int W, H;      // input image
  int sw, sh;    // LSD scaled image (floor(W*scale), floor(H*scale))
  int msld_s;    // int(5*W/800.0)
};

// Gaussian sampler tap tables (one set per context, depend on W,H,scale,sigma_scale only)
struct LslTaps {
  double* kx;  // [8][sw]  7 normalised taps (+pad) per output column, tap-major (lsd.cpp:581-585)
  int* xc;     // [sw]     centre pixel per output column
  double* ky;  // [sh][8]
  int* yc;     // [sh]
  int h, n;    // half width, taps (3, 7 for sigma 0.75)
};

// Per-frame device scratch; frame f lives at base + f * stride for each array.
struct LslWork {
  uint8_t* img;     // [B][H*W*3]  (only used by the host-buffer path)
  float* depth;     // [B][H*W]
  uint8_t* gray;    // [B][H*W]
  double* aux;      // [B][H*sw]   x-pass of the sampler
  double* scaled;   // [B][sh*sw]
  double* angles;   // [B][sh*sw]
  double* modgrad;  // [B][sh*sw]
  double2* cs;      // [B][sh*sw]  (cos, sin) of the level-line angle; x = 2.0 marks NOTDEF
  uint16_t* binT;   // [B][sw*sh]  gradient bin, column-major (x*sh+y); 0xFFFF = not a seed
  uint8_t* used;    // [B][sh*sw]
  int32_t* seeds;   // [B][sh*sw]  x | y<<16 in list_p order
  int32_t* nseeds;  // [B]
  int32_t* reg;     // [B][sh*sw]  region pixel list x | y<<16
  double* rects;    // [B][LSL_MAX_RECTS][12] candidate rectangles (struct rect of lsd.cpp:1075-1084) in sequence order
  uint8_t* rect_ok; // [B][LSL_MAX_RECTS]     1 = passed the NFA validation
  int32_t* nrects;  // [B]
  double* segs;     // [B][LSL_MAX_SEGS*5]
  int32_t* nsegs;   // [B]
  int16_t* gx;      // [B][H*W]  Sobel 5x5 d/dx (exact integers, |v| <= 6570)
  int16_t* gy;      // [B][H*W]
  // per-frame line tables
  int32_t* cand_seg;   // [B][LSL_MAX_SEGS] segment index of candidate c (len > thres)
  int32_t* ncand;      // [B]
  int32_t* keep_cand;  // [B][LSL_MAX_LINES] candidate index of kept line k
  int32_t* nlines;     // [B]
  int32_t* npts;       // [B][LSL_MAX_LINES] inlier count of kept line
  double* pts;         // [B][LSL_MAX_LINES][LSL_MAX_SMP*3] inlier points xyz
  int32_t* inl_idx;    // [B][LSL_MAX_LINES][LSL_MAX_SMP] inlier sample indices (debug / parity)
  lsl_line_rec* lines; // [B][LSL_MAX_LINES]
  uint32_t* seeds_rng; // [B]
  int32_t* rng_state;  // [B][36] glibc TYPE_3 state carried from the RANSAC stage to the MSLD stage
  int32_t* lm_iters;   // [B][LSL_MAX_LINES]
  int32_t* msld_fail;  // [B][LSL_MAX_LINES] 1 = descriptor has no valid sample (filled from rand())
};

// ---- pair registration workspace (k_pair.cu) ----
struct LslPairDesc {
  const lsl_line_rec* q;  // query (newer) frame lines, device
  const lsl_line_rec* t;  // train (older) frame lines, device
  int nq, nt, id_q, id_t;
  uint32_t seed;
  int adjacent;
  size_t d_off;  // offset of this pair's nq x nt distance matrix in LslPairWork::D (doubles)
  size_t m_off;  // offset of this pair's match-sized slices (in matches)
  int cap_m;     // min(nq, nt): upper bound of the match count
  int pad_;
};
struct LslPairScratch {
  double* md;        // [M][72]  gathered per-match data
  double* dab;       // [M][2]   Mahalanobis distances of the last scoring pass
  int32_t* sel;      // [M][3]   RANSAC / refined / trial inlier index lists
  double* lm;        // [M][306] per-match blocks of the LM refinement (k_pair.cu:LM_STRIDE)
  int32_t* okf;      // [M]
  float* tfs;        // [pairs][max_iter][12] hypotheses
  int32_t* cnts;     // [pairs][max_iter]     inlier counts
  uint16_t* trip;    // [pairs][max_iter][3]  sampled match triples
  int32_t* n_inl;    // [pairs]
  int32_t* n_rinl;   // [pairs]
  float* tf_ransac;  // [pairs][16]
  int max_iter;
};
struct LslPairWork {
  LslPairDesc* d_pairs;
  double* D;
  lsl_match* matches;  // [M]
  int32_t* nmatch;     // [pairs]
  lsl_pose_rec* recs;  // [pairs]
  LslPairScratch sc;
  size_t cap_pairs, cap_m, cap_d;
  size_t last_tot_m;                  // match slots of the last batch (sum of cap_m)
  std::vector<LslPairDesc> h_pairs;   // descriptors of the last batch (host copy)
  std::vector<int32_t> h_nmatch, h_ninl, h_nrinl;
};

// ---- point features of a pair (k_hybrid.cu) ----
struct LslPairPts {
  const float* qx;   // query feature_locations_3d_ [nqp][4]
  const float* tx;   // train
  const float* qd;   // query feature_descriptors_ [nqp][dim]
  const float* td;
  int nqp, ntp, dim, cap_pm;   // cap_pm = nqp: upper bound of the point match count
  int kind, pad_;              // 0: f32 rows, L2 (BruteForce); 1: u8 rows of dim bytes, Hamming (BruteForce-HammingLUT)
  size_t pm_off;     // offset of this pair's point-match-sized slices
  size_t knn_off;    // offset of this pair's per-query-row nearest-neighbour records
};
struct LslHybScratch {
  double* pmd;       // [PM][22]  gathered per-point-match data (k_hybrid.cu:PMD_STRIDE)
  double* pd2;       // [PM]      errorFunction2 values of the last scoring pass
  int32_t* psel;     // [PM][3]   RANSAC / refined / trial point inlier index lists
  double* plm;       // [PM][160] LM blocks of the point landmarks (k_hybrid.cu:PLM_STRIDE)
  int32_t* pokf;     // [PM]
  uint8_t* ptidx;    // [pairs][max_iter][2] rand()%nPt draws of getTransform_Lns_Pts_pcl
  int32_t* rng;      // [pairs][33] glibc rand() state after featureMatching
  int32_t* n_pinl;   // [pairs]
  int32_t* n_prinl;  // [pairs]
};
struct LslHybWork {
  LslPairPts* d_ppairs;
  void* knn;            // [sum nqp] {float d1, d2; int i1}
  lsl_match* pmatches;  // [PM]
  int32_t* npmatch;     // [pairs]
  LslHybScratch hs;
  size_t cap_pairs, cap_pm, cap_knn;
  int max_iter;
  std::vector<LslPairPts> h_ppairs;
  std::vector<int32_t> h_npmatch, h_npinl, h_nprinl;
  bool last_hybrid;     // the last pair call ran the point + line path
  // tensor-core pre-filter of featureMatching (k_match_tc.cu): candidate train rows per query row
  int32_t* tc_cand; int32_t* tc_cnt; void* tc_stats; size_t cap_tc;
};

// ---- point-feature detection workspace (k_sift.cu): pyramid of a few frames + candidate lists + dense output tables ----
struct LslSiftWork { void* block; size_t bytes; };

// ---- computeRelativeMotion_Ransac scratch (k_hybrid.cu:relmotion_kernel), allocated per call ----
#define RM_STRIDE 52   // aA aB bA bB (12) | a.DU_A a.DU_B b.DU_A b.DU_B (36) | a.u (3) | pad
#define RM_M 7
struct RmScratch {
  double* g;       // [M][52] gathered pairs
  double* hx;      // [M] x 4: hx, wrk, e, wrk2
  double* jac;     // [M][7]
  int32_t* flag;   // [M]
  int32_t* cur;    // [M] x 2: current subset / trial consensus
  double* hyp;     // [pairs][max_iter][12] R (9), t (3)
  int32_t* cnts;   // [pairs][max_iter]
  uint16_t* trip;  // [pairs][max_iter][3]
  double* outRt;   // [pairs][12]
  int32_t* outn;   // [pairs][4]: consensus size, LM calls, have, pad
  int max_iter;
};

// Kernel ids for the per-kernel device timers (CUDA events on the context stream)
enum LslKernelId {
  LSL_K_GRAY = 0, LSL_K_XPASS, LSL_K_YPASS, LSL_K_LLANGLE, LSL_K_SEEDS, LSL_K_SOBEL, LSL_K_REGION, LSL_K_NFA, LSL_K_RANSAC3D,
  LSL_K_MSLD, LSL_K_RANDFILL, LSL_K_MLE, LSL_K_GATHER, LSL_K_MATCH, LSL_K_POSE, LSL_K_MATCHPTS, LSL_K_POSEHYB, LSL_K_RELMOTION, LSL_K_PNG, LSL_K_INFLATE, LSL_K_PNG_D, LSL_K_INFLATE_D, LSL_K_SIFT, LSL_K_COUNT
};

// Line records of all frames of one extract call live in ONE device allocation (stream-ordered pool);
// frames reference-count it.
struct LslLineBlock {
  lsl_line_rec* d;
  int refs;
};

// Point features of all frames of one lsl_frames_set_points_batch call live in one device allocation; frames reference-count it.
struct LslPointBlock {
  float* d_xyz1; void* d_desc;
  float* d_kp;      // [n][6] x, y, size, angle, response, octave + 256 layer — only for points detected on the device (k_sift.cu)
  bool pooled;      // one stream-ordered allocation (cudaMallocAsync) starting at d_xyz1 holds all three tables
  int refs;
};

struct lsl_frame {
  lsl_ctx* ctx;
  int nlines, nsegs;
  lsl_line_rec* d_lines;            // device records (inside blk, or own allocation when blk == nullptr)
  LslLineBlock* blk;
  std::vector<lsl_line_rec> lines;  // host mirror, filled on first lsl_frame_lines call
  bool have_host;
  // debug mode only (lsl_ctx_set_debug): LSD output and per-line intermediates
  std::vector<double> segs;
  std::vector<int32_t> dbg_npts, dbg_inl, dbg_seg, dbg_lm;
  bool have_dbg;
  // point features handed in by the caller (lsl_frame_set_points); device copies
  int npoints, pdim, pkind;    // pkind 0: f32 descriptors, 1: u8 (ORB) rows of pdim bytes
  float* d_xyz1;   // [npoints][4]
  float* d_desc;   // [npoints][pdim]
  LslPointBlock* pblk;   // non-null: d_xyz1 / d_desc point into a shared block
  float* d_kp;           // keypoint geometry of device-detected points (inside pblk), else nullptr
};

struct lsl_ctx {
  // every entry point that touches the context takes this lock (LSL_ENTER): calls from several host threads on ONE
  // context are serialised in arrival order (recursive: lsl_extract -> lsl_extract_batch, lsl_graph_add_frame ->
  // lsl_match_pair_batch); concurrency across host threads = one context per thread
  mutable std::recursive_mutex mu;
  lsl_params P;
  int device, max_batch, max_w, max_h;
  cudaStream_t stream;       // stream every call of this context runs on
  cudaStream_t own_stream;   // created with the context; `stream` may be replaced by lsl_ctx_set_stream
  cudaEvent_t ev0, ev3;      // whole-call bracket
  cudaStream_t copy_stream;  // depth upload of the host-buffer path (overlaps the image / LSD kernels)
  cudaEvent_t ev_depth, ev_fork;
  bool depth_async;          // the 3D-line stage must wait for ev_depth
  int img_chunks;            // > 0: the RGB planes arrive in this many upload chunks (events ev_img[c]) on the copy stream
  cudaEvent_t ev_img[8];
  cudaEvent_t kev[LSL_K_COUNT][2];
  bool kran[LSL_K_COUNT];
  float kms[LSL_K_COUNT];
  int debug;
  int32_t* d_goff;           // [max_batch] gather offsets
  void* nccl_lib; void* nccl_comm; int nccl_rank, nccl_nranks; bool nccl_own;
  cudaStream_t pair_stream; cudaEvent_t ev0p, ev3p; int pair_inflight; bool pair_hybrid;   // lsl_match_pair_batch_begin / _end
  LslSiftWork sift; int sift_kind, sift_max_kp, sift_root;   // point detector run by every extract call (0: none, 1: SIFT)
  CUtensorMap tmap_gray; bool tmap_gray_ok;   // TMA tile map of the gray planes (sobel5_tma_kernel), valid for `dims`
  LslDims dims;   // dims the workspace / taps were last prepared for
  LslTaps taps;
  LslWork wk;
  void* wk_block; size_t wk_bytes;
  LslPairWork pw;
  LslHybWork hw;
  double cam_fx, cam_dt;     // focal length / asynch time used by the point-edge information (compPt3dCov)
  uint8_t* h_pin; size_t h_pin_bytes;   // pinned staging
  uint8_t* d_gather; size_t d_gather_bytes;      // pose exchange buffer of lsl_allgather_poses
  uint16_t* d_depth16; size_t d_depth16_bytes;   // raw 16-bit depth planes of lsl_extract_batch_u16 (allocated on first use)
  std::string err;
  lsl_stats stats;
  float ms_total, ms_rg;
};

#define LSL_LOCK(c) std::lock_guard<std::recursive_mutex> lsl_lock_((c)->mu)
#define LSL_ENTER(c) LSL_LOCK(c); cudaSetDevice((c)->device)

// device-time bracket of one kernel launch
#define LSL_KSTART(ctx, id) do { cudaEventRecord((ctx)->kev[id][0], (ctx)->stream); } while (0)
#define LSL_KSTOP(ctx, id) do { cudaEventRecord((ctx)->kev[id][1], (ctx)->stream); (ctx)->kran[id] = true; (ctx)->stats.kernel_launches += 1; } while (0)

#define LSL_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
      return LSL_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

// kernel launchers (one per .cu)
int lsl_launch_image(lsl_ctx* ctx, int f0, int n, const uint8_t* d_img, int channels);
int lsl_launch_seeds(lsl_ctx* ctx, int n);
int lsl_launch_lsd(lsl_ctx* ctx, int n);
int lsl_launch_lines(lsl_ctx* ctx, int n, const float* d_depth, const double K[9], double dt);
int lsl_prepare_taps(lsl_ctx* ctx);
int lsl_prepare_tmaps(lsl_ctx* ctx);
int lsl_launch_gather(lsl_ctx* ctx, int n, lsl_line_rec* dst);
int lsl_launch_depth_u16(lsl_ctx* ctx, cudaStream_t st, const uint16_t* d_in, float* d_out, size_t count, float scale);
int lsl_launch_match(lsl_ctx* ctx, int npairs);
int lsl_launch_pose(lsl_ctx* ctx, int npairs);
int lsl_launch_match_points(lsl_ctx* ctx, int npairs, int max_nq, int dim, int kind);
int lsl_launch_rootsift(lsl_ctx* ctx, float* d_desc, int n, int dim);
int lsl_launch_inliers_error(lsl_ctx* ctx, const float* qx, const float* tx, const lsl_match* d_ms, int n, const float* d_tf, double squared_max,
                             double* d_dist, lsl_match* d_out, int32_t* d_n, double* d_rmse);
int lsl_launch_match_points_tc(lsl_ctx* ctx, int npairs, int max_nq, int dim, int* used);
int lsl_launch_sift(lsl_ctx* ctx, int n, const float* d_depth, int W, int H, const double K[9], lsl_frame** frames);
int lsl_launch_pose_hybrid(lsl_ctx* ctx, int npairs, double fx, double dt);
int lsl_launch_relmotion(lsl_ctx* ctx, int npairs, RmScratch rs);
