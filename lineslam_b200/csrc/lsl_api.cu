// lsl_api.cu — C ABI of liblsl_b200 (include/lsl.h): contexts, frames, batched extraction, pair
// registration and the pose exchange. There is no CPU fallback anywhere in this library: without a
// CUDA device every entry point that computes returns LSL_ERR_NO_DEVICE.
#include "lsl_internal.h"
#include "shared/lsl_params_default.h"
#include "shared/lsl_rand.h"
#include "../../include/lsl_tum.h"
#include <dlfcn.h>
#include <string.h>
#include <stdlib.h>
#include <new>

extern "C" void lsl_params_default(lsl_params* p) { if (p) lsl_params_default_impl(p); }

extern "C" const char* lsl_strerror(int s) {
  switch (s) {
    case LSL_OK: return "ok";
    case LSL_ERR_ARG: return "invalid argument";
    case LSL_ERR_CUDA: return "CUDA error";
    case LSL_ERR_CAPACITY: return "capacity exceeded";
    case LSL_ERR_NO_DEVICE: return "no CUDA device (liblsl_b200 has no CPU fallback)";
    case LSL_ERR_NCCL: return "NCCL error";
    case LSL_ERR_BUSY: return "a pair batch is in flight";
    default: return "unknown status";
  }
}
extern "C" const char* lsl_last_error(const lsl_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }

static const char* const kKernelNames[LSL_K_COUNT] = {
    "gray_kernel", "xpass_kernel", "ypass_kernel", "ll_angle_kernel", "seed_list_kernel", "sobel5_kernel",
    "lsd_region_kernel", "lsd_nfa_kernel", "line3d_ransac_kernel", "line_msld_kernel", "msld_randfill_kernel", "line_mle_kernel",
    "gather_lines_kernel", "match_lines_kernel", "pose_kernel", "match_points_kernel", "pose_hybrid_kernel", "relmotion_kernel", "png_unfilter_kernel", "png_inflate_kernel", "png_unfilter_kernel(depth)", "png_inflate_kernel(depth)", "sift_kernels"};
extern "C" const char* lsl_kernel_name(int i) { return (i >= 0 && i < LSL_K_COUNT) ? kKernelNames[i] : ""; }

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

// carve the per-frame work arrays out of one allocation
static int carve(lsl_ctx* ctx, bool measure, size_t* total) {
  const int B = ctx->max_batch;
  const size_t npix = (size_t)ctx->max_w * ctx->max_h;
  const size_t sw = (size_t)floor(ctx->max_w * ctx->P.lsd_scale), sh = (size_t)floor(ctx->max_h * ctx->P.lsd_scale);
  const size_t spix = sw * sh;
  size_t off = 0;
  uint8_t* base = (uint8_t*)ctx->wk_block;
  LslWork& w = ctx->wk;
#define CARVE(field, type, count)                               \
  do {                                                          \
    if (!measure) w.field = (type*)(base + off);                \
    off += align_up(sizeof(type) * (size_t)(count));            \
  } while (0)
  CARVE(img, uint8_t, B * npix * 3);
  CARVE(depth, float, B * npix);
  CARVE(gray, uint8_t, B * npix);
  CARVE(aux, double, B * (size_t)ctx->max_h * sw);
  CARVE(scaled, double, B * spix);
  CARVE(angles, double, B * spix);
  CARVE(modgrad, double, B * spix);
  CARVE(cs, double2, B * spix);
  CARVE(binT, uint16_t, B * spix);
  CARVE(used, uint8_t, B * spix);
  CARVE(seeds, int32_t, B * spix);
  CARVE(nseeds, int32_t, B);
  CARVE(reg, int32_t, B * spix);
  CARVE(rects, double, (size_t)B * LSL_MAX_RECTS * 12);
  CARVE(rect_ok, uint8_t, (size_t)B * LSL_MAX_RECTS);
  CARVE(nrects, int32_t, B);
  CARVE(segs, double, (size_t)B * LSL_MAX_SEGS * 5);
  CARVE(nsegs, int32_t, B);
  CARVE(gx, int16_t, B * npix);
  CARVE(gy, int16_t, B * npix);
  CARVE(cand_seg, int32_t, (size_t)B * LSL_MAX_SEGS);
  CARVE(ncand, int32_t, B);
  CARVE(keep_cand, int32_t, (size_t)B * LSL_MAX_LINES);
  CARVE(nlines, int32_t, B);
  CARVE(npts, int32_t, (size_t)B * LSL_MAX_LINES);
  CARVE(pts, double, (size_t)B * LSL_MAX_LINES * LSL_MAX_SMP * 3);
  CARVE(inl_idx, int32_t, (size_t)B * LSL_MAX_LINES * LSL_MAX_SMP);
  CARVE(lines, lsl_line_rec, (size_t)B * LSL_MAX_LINES);
  CARVE(seeds_rng, uint32_t, B);
  CARVE(rng_state, int32_t, (size_t)B * 36);
  CARVE(lm_iters, int32_t, (size_t)B * LSL_MAX_LINES);
  CARVE(msld_fail, int32_t, (size_t)B * LSL_MAX_LINES);
  // tap tables, gather offsets
  if (!measure) ctx->taps.kx = (double*)(base + off); off += align_up(sizeof(double) * sw * 8);
  if (!measure) ctx->taps.xc = (int*)(base + off); off += align_up(sizeof(int) * sw);
  if (!measure) ctx->taps.ky = (double*)(base + off); off += align_up(sizeof(double) * sh * 8);
  if (!measure) ctx->taps.yc = (int*)(base + off); off += align_up(sizeof(int) * sh);
  if (!measure) ctx->d_goff = (int32_t*)(base + off); off += align_up(sizeof(int32_t) * B);
#undef CARVE
  *total = off;
  return LSL_OK;
}

static void free_pair_ws(lsl_ctx* ctx) {
  LslPairWork& p = ctx->pw;
  void* ptrs[] = {p.d_pairs, p.D, p.matches, p.nmatch, p.recs, p.sc.md, p.sc.dab, p.sc.sel, p.sc.lm, p.sc.okf, p.sc.tfs,
                  p.sc.cnts, p.sc.trip, p.sc.n_inl, p.sc.n_rinl, p.sc.tf_ransac};
  for (void* q : ptrs) if (q) cudaFree(q);
  memset(&p.sc, 0, sizeof(p.sc));
  p.d_pairs = nullptr; p.D = nullptr; p.matches = nullptr; p.nmatch = nullptr; p.recs = nullptr;
  p.cap_pairs = p.cap_m = p.cap_d = 0;
  LslHybWork& h = ctx->hw;
  void* hp[] = {h.d_ppairs, h.knn, h.pmatches, h.npmatch, h.hs.pmd, h.hs.pd2, h.hs.psel, h.hs.plm, h.hs.pokf, h.hs.ptidx,
                h.hs.rng, h.hs.n_pinl, h.hs.n_prinl};
  for (void* q : hp) if (q) cudaFree(q);
  memset(&h.hs, 0, sizeof(h.hs));
  h.d_ppairs = nullptr; h.knn = nullptr; h.pmatches = nullptr; h.npmatch = nullptr;
  h.cap_pairs = h.cap_pm = h.cap_knn = 0; h.max_iter = 0; h.last_hybrid = false;
  if (h.tc_cand) cudaFree(h.tc_cand);
  if (h.tc_cnt) cudaFree(h.tc_cnt);
  if (h.tc_stats) cudaFree(h.tc_stats);
  h.tc_cand = nullptr; h.tc_cnt = nullptr; h.tc_stats = nullptr; h.cap_tc = 0;
}

extern "C" int lsl_ctx_create(lsl_ctx** out, const lsl_params* params, int cuda_device, int max_batch, int max_w, int max_h) {
  if (!out || max_batch < 1 || max_w < 16 || max_h < 16 || (max_w & 3)) return LSL_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return LSL_ERR_NO_DEVICE;
  if (cuda_device < 0 || cuda_device >= ndev) return LSL_ERR_ARG;
  lsl_ctx* ctx = new (std::nothrow) lsl_ctx();
  if (!ctx) return LSL_ERR_ARG;
  if (params) ctx->P = *params; else lsl_params_default_impl(&ctx->P);
  ctx->device = cuda_device; ctx->max_batch = max_batch; ctx->max_w = max_w; ctx->max_h = max_h;
  ctx->wk_block = nullptr;
  memset(&ctx->pw.sc, 0, sizeof(ctx->pw.sc));
  ctx->pw.d_pairs = nullptr; ctx->pw.D = nullptr; ctx->pw.matches = nullptr; ctx->pw.nmatch = nullptr; ctx->pw.recs = nullptr;
  ctx->pw.cap_pairs = ctx->pw.cap_m = ctx->pw.cap_d = 0; ctx->pw.last_tot_m = 0;
  memset(&ctx->hw.hs, 0, sizeof(ctx->hw.hs));
  ctx->hw.d_ppairs = nullptr; ctx->hw.knn = nullptr; ctx->hw.pmatches = nullptr; ctx->hw.npmatch = nullptr;
  ctx->hw.cap_pairs = ctx->hw.cap_pm = ctx->hw.cap_knn = 0; ctx->hw.max_iter = 0; ctx->hw.last_hybrid = false;
  ctx->hw.tc_cand = nullptr; ctx->hw.tc_cnt = nullptr; ctx->hw.tc_stats = nullptr; ctx->hw.cap_tc = 0;
  ctx->cam_fx = 525.0; ctx->cam_dt = 0.0;   // K(0,0) of src/openni_listener.cpp:1256; replaced by the K of the last extract call
  ctx->h_pin = nullptr; ctx->h_pin_bytes = 0;
  ctx->d_depth16 = nullptr; ctx->d_depth16_bytes = 0;
  ctx->d_gather = nullptr; ctx->d_gather_bytes = 0;
  ctx->tmap_gray_ok = false;
  ctx->pair_inflight = 0; ctx->pair_hybrid = false;
  // (an extraction stream of higher priority than this one was measured: 202 instead of 198 ms per 1184-frame step)
  cudaStreamCreateWithFlags(&ctx->pair_stream, cudaStreamNonBlocking);
  cudaEventCreate(&ctx->ev0p); cudaEventCreate(&ctx->ev3p);
  ctx->sift.block = nullptr; ctx->sift.bytes = 0; ctx->sift_kind = 0; ctx->sift_max_kp = 600; ctx->sift_root = 1;
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  memset(&ctx->dims, 0, sizeof(ctx->dims));
  memset(ctx->kran, 0, sizeof(ctx->kran));
  memset(ctx->kms, 0, sizeof(ctx->kms));
  ctx->ms_total = ctx->ms_rg = 0.f;
  ctx->debug = 0;
  ctx->nccl_lib = nullptr; ctx->nccl_comm = nullptr; ctx->nccl_rank = 0; ctx->nccl_nranks = 1; ctx->nccl_own = false;
  if (ctx->P.lsd_n_bins > 4096 || ctx->P.line_sample_max_num + 1 > LSL_MAX_SMP || ctx->P.num_cells_lineseg_range > 32 ||
      ctx->P.num_cells_lineseg_range < 10) { delete ctx; return LSL_ERR_ARG; }
  cudaError_t e = cudaSetDevice(cuda_device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_depth, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  ctx->depth_async = false; ctx->img_chunks = 0;
  for (int c = 0; c < 8 && e == cudaSuccess; ++c) e = cudaEventCreateWithFlags(&ctx->ev_img[c], cudaEventDisableTiming);
  ctx->stream = ctx->own_stream;
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev3);
  for (int k = 0; k < LSL_K_COUNT && e == cudaSuccess; ++k) {
    e = cudaEventCreate(&ctx->kev[k][0]);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->kev[k][1]);
  }
  if (e == cudaSuccess) {  // keep freed line blocks in the stream-ordered pool instead of returning them to the OS
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, cuda_device) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  size_t total = 0;
  carve(ctx, true, &total);
  ctx->wk_bytes = total;
  if (e == cudaSuccess) e = cudaMalloc(&ctx->wk_block, total);
  if (e != cudaSuccess) { delete ctx; return LSL_ERR_CUDA; }
  carve(ctx, false, &total);
  *out = ctx;
  return LSL_OK;
}

static void nccl_teardown(lsl_ctx* ctx);

extern "C" void lsl_ctx_destroy(lsl_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  lsl_tum_release(ctx);   // staging buffers of the TUM loader are keyed by the context: never outlive it
  nccl_teardown(ctx);
  if (ctx->wk_block) cudaFree(ctx->wk_block);
  free_pair_ws(ctx);
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  if (ctx->d_depth16) cudaFree(ctx->d_depth16);
  if (ctx->d_gather) cudaFree(ctx->d_gather);
  if (ctx->sift.block) cudaFree(ctx->sift.block);
  cudaStreamSynchronize(ctx->pair_stream); cudaStreamDestroy(ctx->pair_stream);
  cudaEventDestroy(ctx->ev0p); cudaEventDestroy(ctx->ev3p);
  cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev3);
  for (int k = 0; k < LSL_K_COUNT; ++k) { cudaEventDestroy(ctx->kev[k][0]); cudaEventDestroy(ctx->kev[k][1]); }
  cudaStreamDestroy(ctx->own_stream);
  cudaStreamDestroy(ctx->copy_stream); cudaEventDestroy(ctx->ev_depth); cudaEventDestroy(ctx->ev_fork);
  for (int c = 0; c < 8; ++c) cudaEventDestroy(ctx->ev_img[c]);
  delete ctx;
}

extern "C" int lsl_ctx_set_stream(lsl_ctx* ctx, void* cuda_stream) {
  if (!ctx) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return LSL_OK;
}
extern "C" int lsl_ctx_set_debug(lsl_ctx* ctx, int on) {
  if (!ctx) return LSL_ERR_ARG;
  LSL_LOCK(ctx);
  ctx->debug = on ? 1 : 0;
  return LSL_OK;
}

static int set_dims(lsl_ctx* ctx, int W, int H) {
  if (W > ctx->max_w || H > ctx->max_h || W < 16 || H < 16 || (W & 3)) { ctx->err = "image size outside context limits"; return LSL_ERR_ARG; }
  if (ctx->dims.W == W && ctx->dims.H == H) return LSL_OK;
  ctx->dims.W = W; ctx->dims.H = H;
  ctx->dims.sw = (int)floor(W * ctx->P.lsd_scale);
  ctx->dims.sh = (int)floor(H * ctx->P.lsd_scale);
  ctx->dims.msld_s = (int)(5 * W / 800.0);
  if (ctx->dims.sw >= 65536 || ctx->dims.sh >= 32768) return LSL_ERR_ARG;
  int rc = lsl_prepare_taps(ctx);
  if (rc) return rc;
  return lsl_prepare_tmaps(ctx);
}

static int ensure_pinned(lsl_ctx* ctx, size_t bytes) {
  if (ctx->h_pin_bytes >= bytes) return LSL_OK;
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  ctx->h_pin = nullptr; ctx->h_pin_bytes = 0;
  LSL_CUDA(cudaMallocHost((void**)&ctx->h_pin, bytes));
  ctx->h_pin_bytes = bytes;
  return LSL_OK;
}

static void clear_ktimes(lsl_ctx* ctx, int from, int to) {
  for (int k = from; k < to; ++k) { ctx->kran[k] = false; ctx->kms[k] = 0.f; }
}
static void collect_ktimes(lsl_ctx* ctx, int from, int to) {
  for (int k = from; k < to; ++k)
    if (ctx->kran[k]) cudaEventElapsedTime(&ctx->kms[k], ctx->kev[k][0], ctx->kev[k][1]);
}

// Runs all extraction kernels for n frames whose inputs are on the device, then builds the handles.
static int extract_device(lsl_ctx* ctx, int n, const uint8_t* d_imgs, int channels, const float* d_depths, int W, int H,
                          const double K[9], double dt, const uint32_t* seeds, lsl_frame** out) {
  int rc = set_dims(ctx, W, H);
  if (rc) return rc;
  ctx->cam_fx = K[0]; ctx->cam_dt = dt;   // the global K of src/node.cpp:200-206 (read again by the g2o set-up)
  LslWork& w = ctx->wk;
  cudaStream_t st = ctx->stream;
  std::vector<uint32_t> sd(n);
  for (int i = 0; i < n; ++i) sd[i] = seeds ? seeds[i] : 1u;
  LSL_CUDA(cudaMemcpyAsync(w.seeds_rng, sd.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
  clear_ktimes(ctx, 0, LSL_K_MATCH);
  cudaEventRecord(ctx->ev0, st);
  if (ctx->img_chunks > 0) {   // host-buffer path: image stage per upload chunk (the next chunk's copy runs underneath)
    const int C = ctx->img_chunks;
    ctx->img_chunks = 0;
    for (int c = 0; c < C; ++c) {
      const int f0 = (int)((long)n * c / C), f1 = (int)((long)n * (c + 1) / C);
      if (f1 == f0) continue;
      LSL_CUDA(cudaStreamWaitEvent(st, ctx->ev_img[c], 0));
      if ((rc = lsl_launch_image(ctx, f0, f1 - f0, d_imgs, channels))) return rc;
    }
  } else if ((rc = lsl_launch_image(ctx, 0, n, d_imgs, channels))) return rc;
  if ((rc = lsl_launch_seeds(ctx, n))) return rc;
  if ((rc = lsl_launch_lsd(ctx, n))) return rc;
  if ((rc = lsl_launch_lines(ctx, n, d_depths, K, dt))) return rc;
  // ---- counts back (8 bytes per frame), then one dense block for the records of the whole batch
  std::vector<int32_t> nsegs(n), nlines(n), goff(n);
  LSL_CUDA(cudaMemcpyAsync(nsegs.data(), w.nsegs, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  LSL_CUDA(cudaMemcpyAsync(nlines.data(), w.nlines, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  LSL_CUDA(cudaStreamSynchronize(st));
  ctx->stats.d2h_bytes += 8 * n;
  size_t tot = 0;
  for (int f = 0; f < n; ++f) {
    if (nsegs[f] > LSL_MAX_SEGS || nlines[f] > LSL_MAX_LINES) { ctx->err = "per-frame segment/line table overflow"; return LSL_ERR_CAPACITY; }
    goff[f] = (int32_t)tot;
    tot += nlines[f];
  }
  LslLineBlock* blk = nullptr;
  if (tot) {
    blk = new (std::nothrow) LslLineBlock();
    if (!blk) return LSL_ERR_ARG;
    blk->refs = 0; blk->d = nullptr;
    // size classes of 16384 records (17 MB) with 1/8 headroom: consecutive batches of a stream then ask the stream-ordered
    // pool for the same block size and reuse each other's freed blocks (an exact-fit request grew the pool by a fresh
    // cudaMalloc whenever a batch carried a few more lines than any before it: sporadic 100-500 ms stalls)
    const size_t want = (size_t)tot + (size_t)tot / 8;
    const size_t recs_alloc = (want + 16383) / 16384 * 16384;
    LSL_CUDA(cudaMallocAsync((void**)&blk->d, sizeof(lsl_line_rec) * recs_alloc, st));
    LSL_CUDA(cudaMemcpyAsync(ctx->d_goff, goff.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    if ((rc = lsl_launch_gather(ctx, n, blk->d))) return rc;
  }
  cudaEventRecord(ctx->ev3, st);
  for (int f = 0; f < n; ++f) {
    lsl_frame* fr = new (std::nothrow) lsl_frame();
    if (!fr) return LSL_ERR_ARG;
    fr->ctx = ctx; fr->nlines = nlines[f]; fr->nsegs = nsegs[f]; fr->d_lines = nullptr; fr->blk = nullptr;
    fr->have_host = false; fr->have_dbg = false;
    if (fr->nlines) { fr->blk = blk; blk->refs += 1; fr->d_lines = blk->d + goff[f]; }
    if (ctx->debug) {
      fr->have_dbg = true;
      fr->segs.resize((size_t)fr->nsegs * 5);
      fr->dbg_npts.resize(fr->nlines); fr->dbg_inl.resize((size_t)fr->nlines * LSL_MAX_SMP);
      fr->dbg_seg.resize(fr->nlines); fr->dbg_lm.resize(fr->nlines);
      if (fr->nsegs)
        LSL_CUDA(cudaMemcpyAsync(fr->segs.data(), w.segs + (size_t)f * LSL_MAX_SEGS * 5, sizeof(double) * 5 * fr->nsegs, cudaMemcpyDeviceToHost, st));
      if (fr->nlines) {
        LSL_CUDA(cudaMemcpyAsync(fr->dbg_npts.data(), w.npts + (size_t)f * LSL_MAX_LINES, 4 * fr->nlines, cudaMemcpyDeviceToHost, st));
        LSL_CUDA(cudaMemcpyAsync(fr->dbg_inl.data(), w.inl_idx + (size_t)f * LSL_MAX_LINES * LSL_MAX_SMP, 4 * (size_t)fr->nlines * LSL_MAX_SMP, cudaMemcpyDeviceToHost, st));
        LSL_CUDA(cudaMemcpyAsync(fr->dbg_seg.data(), w.keep_cand + (size_t)f * LSL_MAX_LINES, 4 * fr->nlines, cudaMemcpyDeviceToHost, st));
        LSL_CUDA(cudaMemcpyAsync(fr->dbg_lm.data(), w.lm_iters + (size_t)f * LSL_MAX_LINES, 4 * fr->nlines, cudaMemcpyDeviceToHost, st));
      }
    }
    ctx->stats.segments += fr->nsegs; ctx->stats.lines3d += fr->nlines;
    out[f] = fr;
  }
  LSL_CUDA(cudaStreamSynchronize(st));
  ctx->kran[LSL_K_SIFT] = false; ctx->kms[LSL_K_SIFT] = 0.f;
  if (ctx->sift_kind == 1 && channels >= 1) {   // the other half of Node::Node: point features of the same frames (src/node.cpp:219-310)
    if ((rc = lsl_launch_sift(ctx, n, d_depths, W, H, K, out))) return rc;
    collect_ktimes(ctx, LSL_K_SIFT, LSL_K_SIFT + 1);
  }
  ctx->stats.frames += n;
  cudaEventElapsedTime(&ctx->ms_total, ctx->ev0, ctx->ev3);
  collect_ktimes(ctx, 0, LSL_K_MATCH);
  ctx->ms_rg = ctx->kms[LSL_K_REGION];
  return LSL_OK;
}

extern "C" int lsl_extract_batch_dev(lsl_ctx* ctx, int n, const uint8_t* d_imgs, int channels, const float* d_depths, int W,
                                     int H, const double K[9], double dt, const uint32_t* seeds, lsl_frame** out) {
  if (!ctx || !d_imgs || !d_depths || !K || !out || n < 1 || (channels != 1 && channels != 3)) return LSL_ERR_ARG;
  if (n > ctx->max_batch) { ctx->err = "batch larger than the context's max_batch"; return LSL_ERR_CAPACITY; }
  LSL_ENTER(ctx);
  return extract_device(ctx, n, d_imgs, channels, d_depths, W, H, K, dt, seeds, out);
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// Host-buffer path shared by lsl_extract_batch (f32 depth, depth_elem = 4) and lsl_extract_batch_u16 (raw 16-bit depth,
// depth_elem = 2: converted to metres on the device, on the copy stream right behind its upload).
static int extract_host(lsl_ctx* ctx, int n, const uint8_t* const* imgs, int channels, const void* const* depths, int depth_elem,
                        float depth_scale, int W, int H, const double K[9], double dt, const uint32_t* seeds, lsl_frame** out) {
  if (!ctx || !imgs || !depths || !K || !out || n < 1 || (channels != 1 && channels != 3)) return LSL_ERR_ARG;
  if (n > ctx->max_batch) { ctx->err = "batch larger than the context's max_batch"; return LSL_ERR_CAPACITY; }
  LSL_ENTER(ctx);
  int rc = set_dims(ctx, W, H);
  if (rc) return rc;
  const size_t npix = (size_t)W * H;
  const size_t ib = npix * channels, db = npix * (size_t)depth_elem;
  for (int i = 0; i < n; ++i)
    if (!imgs[i] || !depths[i]) return LSL_ERR_ARG;
  uint8_t* d_draw = (uint8_t*)ctx->wk.depth;   // where the uploaded depth bytes land
  if (depth_elem == 2) {
    if (ctx->d_depth16_bytes < db * n) {
      LSL_CUDA(cudaStreamSynchronize(ctx->stream));
      if (ctx->d_depth16) cudaFree(ctx->d_depth16);
      ctx->d_depth16 = nullptr; ctx->d_depth16_bytes = 0;
      LSL_CUDA(cudaMalloc((void**)&ctx->d_depth16, db * (size_t)ctx->max_batch));
      ctx->d_depth16_bytes = db * (size_t)ctx->max_batch;
    }
    d_draw = (uint8_t*)ctx->d_depth16;
  }
  // page-locked caller buffers are copied straight from where they are; pageable ones are staged through
  // the context's pinned buffer so the copies stay asynchronous
  bool pinned = is_pinned(imgs[0]) && is_pinned(depths[0]);
  cudaStream_t depth_stream = ctx->stream;
  if (pinned) {
    bool contiguous = true;
    for (int i = 1; i < n; ++i) contiguous &= (imgs[i] == imgs[0] + ib * i) && ((const uint8_t*)depths[i] == (const uint8_t*)depths[0] + db * i);
    if (contiguous) {
      // the depth planes are first read by the 3D-line stage: their upload runs on the copy stream underneath the
      // image / LSD kernels (the RANSAC launch waits for ev_depth)
      LSL_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
      LSL_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
      const int C = n >= 32 ? 4 : 1;   // RGB planes in C chunks: chunk c + 1 uploads while the image kernels of chunk c run
      for (int c = 0; c < C; ++c) {
        const size_t f0 = (size_t)n * c / C, f1 = (size_t)n * (c + 1) / C;
        if (f1 > f0) LSL_CUDA(cudaMemcpyAsync(ctx->wk.img + ib * f0, imgs[0] + ib * f0, ib * (f1 - f0), cudaMemcpyHostToDevice, ctx->copy_stream));
        LSL_CUDA(cudaEventRecord(ctx->ev_img[c], ctx->copy_stream));
      }
      ctx->img_chunks = C;
      LSL_CUDA(cudaMemcpyAsync(d_draw, depths[0], db * n, cudaMemcpyHostToDevice, ctx->copy_stream));
      depth_stream = ctx->copy_stream;
      ctx->depth_async = true;
    } else {
      for (int i = 0; i < n; ++i) {
        LSL_CUDA(cudaMemcpyAsync(ctx->wk.img + ib * i, imgs[i], ib, cudaMemcpyHostToDevice, ctx->stream));
        LSL_CUDA(cudaMemcpyAsync(d_draw + db * i, depths[i], db, cudaMemcpyHostToDevice, ctx->stream));
      }
    }
  } else {
    if ((rc = ensure_pinned(ctx, (ib + db) * n))) return rc;
    uint8_t* hp = ctx->h_pin;
    for (int i = 0; i < n; ++i) {
      memcpy(hp + ib * i, imgs[i], ib);
      memcpy(hp + ib * n + db * i, depths[i], db);
    }
    LSL_CUDA(cudaMemcpyAsync(ctx->wk.img, hp, ib * n, cudaMemcpyHostToDevice, ctx->stream));
    LSL_CUDA(cudaMemcpyAsync(d_draw, hp + ib * n, db * n, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (depth_elem == 2 && (rc = lsl_launch_depth_u16(ctx, depth_stream, ctx->d_depth16, ctx->wk.depth, npix * n, depth_scale))) return rc;
  if (ctx->depth_async) LSL_CUDA(cudaEventRecord(ctx->ev_depth, ctx->copy_stream));
  ctx->stats.h2d_bytes += (ib + db) * n;
  return extract_device(ctx, n, ctx->wk.img, channels, ctx->wk.depth, W, H, K, dt, seeds, out);
}

extern "C" int lsl_extract_batch(lsl_ctx* ctx, int n, const uint8_t* const* imgs, int channels, const float* const* depths,
                                 int W, int H, const double K[9], double dt, const uint32_t* seeds, lsl_frame** out) {
  return extract_host(ctx, n, imgs, channels, (const void* const*)depths, 4, 1.f, W, H, K, dt, seeds, out);
}

extern "C" int lsl_extract_batch_u16(lsl_ctx* ctx, int n, const uint8_t* const* imgs, int channels, const uint16_t* const* depths,
                                     int W, int H, const double K[9], double dt, const uint32_t* seeds, double depth_factor,
                                     lsl_frame** out) {
  if (!(depth_factor > 0)) return LSL_ERR_ARG;
  return extract_host(ctx, n, imgs, channels, (const void* const*)depths, 2, (float)(1.0 / depth_factor), W, H, K, dt, seeds, out);
}

extern "C" int lsl_extract(lsl_ctx* ctx, const uint8_t* img, int channels, const float* depth, int W, int H,
                           const double K[9], double dt, uint32_t seed, lsl_frame** out) {
  return lsl_extract_batch(ctx, 1, &img, channels, &depth, W, H, K, dt, &seed, out);
}

extern "C" int lsl_frame_num_lines(const lsl_frame* f) { return f ? f->nlines : LSL_ERR_ARG; }
extern "C" int lsl_frame_lines(const lsl_frame* fc, lsl_line_rec* dst, int cap, int* n) {
  lsl_frame* f = const_cast<lsl_frame*>(fc);
  if (!f || !n) return LSL_ERR_ARG;
  *n = f->nlines;
  if (cap < f->nlines || (!dst && f->nlines)) return LSL_ERR_CAPACITY;
  if (!f->nlines) return LSL_OK;
  if (!f->have_host) {  // host mirror on first use
    lsl_ctx* ctx = f->ctx;
    LSL_ENTER(ctx);
    f->lines.resize(f->nlines);
    LSL_CUDA(cudaMemcpyAsync(f->lines.data(), f->d_lines, sizeof(lsl_line_rec) * f->nlines, cudaMemcpyDeviceToHost, ctx->stream));
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += sizeof(lsl_line_rec) * f->nlines;
    f->have_host = true;
  }
  memcpy(dst, f->lines.data(), sizeof(lsl_line_rec) * f->nlines);
  return LSL_OK;
}
extern "C" int lsl_frame_segments(const lsl_frame* f, double* dst, int cap, int* n) {
  if (!f || !n) return LSL_ERR_ARG;
  *n = f->nsegs;
  if (!f->have_dbg) return f->nsegs ? LSL_ERR_ARG : LSL_OK;  // LSD rows are only kept in debug mode
  if (cap < f->nsegs || (!dst && f->nsegs)) return LSL_ERR_CAPACITY;
  if (f->nsegs) memcpy(dst, f->segs.data(), sizeof(double) * 5 * f->nsegs);
  return LSL_OK;
}
extern "C" int lsl_frame_from_lines(lsl_ctx* ctx, const lsl_line_rec* recs, int n, lsl_frame** out) {
  if (!ctx || !out || n < 0 || (n && !recs) || n > LSL_MAX_LINES) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  lsl_frame* fr = new (std::nothrow) lsl_frame();
  if (!fr) return LSL_ERR_ARG;
  fr->ctx = ctx; fr->nlines = n; fr->nsegs = 0; fr->d_lines = nullptr; fr->blk = nullptr;
  fr->have_dbg = false; fr->have_host = true;
  fr->lines.assign(recs, recs + n);
  if (n) {
    LSL_CUDA(cudaMalloc((void**)&fr->d_lines, sizeof(lsl_line_rec) * n));
    LSL_CUDA(cudaMemcpyAsync(fr->d_lines, recs, sizeof(lsl_line_rec) * n, cudaMemcpyHostToDevice, ctx->stream));
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.h2d_bytes += sizeof(lsl_line_rec) * n;
  }
  *out = fr;
  return LSL_OK;
}
static void release_points(lsl_frame* f) {
  if (f->pblk) {
    if (--f->pblk->refs == 0) {
      if (f->pblk->pooled) cudaFreeAsync(f->pblk->d_xyz1, f->ctx->stream);
      else { cudaFree(f->pblk->d_xyz1); cudaFree(f->pblk->d_desc); if (f->pblk->d_kp) cudaFree(f->pblk->d_kp); }
      delete f->pblk;
    }
    f->pblk = nullptr;
  } else {
    if (f->d_xyz1) cudaFree(f->d_xyz1);
    if (f->d_desc) cudaFree(f->d_desc);
  }
  f->d_xyz1 = nullptr; f->d_desc = nullptr; f->d_kp = nullptr; f->npoints = 0;
}

extern "C" void lsl_frame_free(lsl_frame* f) {
  if (!f) return;
  if (f->ctx && f->ctx->pair_inflight) { LSL_ENTER(f->ctx); cudaStreamSynchronize(f->ctx->pair_stream); }   // the batch in flight may read this frame
  if (f->d_lines) {
    LSL_ENTER(f->ctx);
    if (f->blk) {
      if (--f->blk->refs == 0) { cudaFreeAsync(f->blk->d, f->ctx->stream); delete f->blk; }
    } else cudaFree(f->d_lines);
  }
  { LSL_LOCK(f->ctx); cudaSetDevice(f->ctx->device); release_points(f); }
  delete f;
}
extern "C" int lsl_frame_clear_lines(lsl_frame* f) {
  if (!f) return LSL_ERR_ARG;
  if (f->ctx && f->ctx->pair_inflight) { LSL_ENTER(f->ctx); cudaStreamSynchronize(f->ctx->pair_stream); }
  if (f->d_lines) {
    LSL_ENTER(f->ctx);
    if (f->blk) {
      if (--f->blk->refs == 0) { cudaFreeAsync(f->blk->d, f->ctx->stream); delete f->blk; }
    } else cudaFree(f->d_lines);
  }
  f->d_lines = nullptr; f->blk = nullptr; f->nlines = 0;
  f->lines.clear(); f->have_host = true;
  return LSL_OK;
}
// Point features of a frame (inputs of the hot path: Node::feature_locations_3d_ and feature_descriptors_,
// src/node.h; SIFT/SURF rows after squareroot_descriptor_space). Copies to the device; replaces earlier points.
extern "C" int lsl_frame_set_points_ex(lsl_ctx* ctx, lsl_frame* f, const float* xyz1, const void* desc, int n, int dim, int desc_is_u8,
                                       int root_sift) {
  if (!ctx || !f || n < 0 || n > LSL_MAX_POINTS || (n && (!xyz1 || !desc)) || dim < 1 || dim > 512) return LSL_ERR_ARG;
  if (desc_is_u8 && (dim & 3)) { ctx->err = "binary descriptor rows must be a multiple of 4 bytes"; return LSL_ERR_ARG; }
  LSL_ENTER(ctx);
  release_points(f);
  f->npoints = n; f->pdim = dim; f->pkind = desc_is_u8 ? 1 : 0;
  if (n) {
    const size_t row = desc_is_u8 ? (size_t)dim : sizeof(float) * (size_t)dim;
    LSL_CUDA(cudaMalloc((void**)&f->d_xyz1, sizeof(float) * 4 * n));
    LSL_CUDA(cudaMalloc((void**)&f->d_desc, row * n));
    LSL_CUDA(cudaMemcpyAsync(f->d_xyz1, xyz1, sizeof(float) * 4 * n, cudaMemcpyHostToDevice, ctx->stream));
    LSL_CUDA(cudaMemcpyAsync(f->d_desc, desc, row * n, cudaMemcpyHostToDevice, ctx->stream));
    if (root_sift && !desc_is_u8) { int rc = lsl_launch_rootsift(ctx, f->d_desc, n, dim); if (rc) return rc; }
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.h2d_bytes += sizeof(float) * 4 * (size_t)n + row * n;
  }
  return LSL_OK;
}
extern "C" int lsl_frame_set_points(lsl_ctx* ctx, lsl_frame* f, const float* xyz1, const float* desc, int n, int dim) {
  return lsl_frame_set_points_ex(ctx, f, xyz1, desc, n, dim, 0, 0);
}
// Point features of n frames in one go (the Node constructors of a batch): ONE device block, two uploads, one optional
// RootSIFT launch over all rows — instead of two cudaMalloc + a synchronisation per frame.
extern "C" int lsl_frames_set_points_batch(lsl_ctx* ctx, int n, lsl_frame* const* frames, const float* xyz1, const void* desc,
                                           const int32_t* counts, int dim, int desc_is_u8, int root_sift) {
  if (!ctx || n < 1 || !frames || !counts || dim < 1 || dim > 512) return LSL_ERR_ARG;
  if (desc_is_u8 && (dim & 3)) return LSL_ERR_ARG;
  size_t tot = 0;
  for (int i = 0; i < n; ++i) {
    if (!frames[i] || counts[i] < 0 || counts[i] > LSL_MAX_POINTS) return LSL_ERR_ARG;
    tot += (size_t)counts[i];
  }
  if (tot && (!xyz1 || !desc)) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  const size_t row = desc_is_u8 ? (size_t)dim : sizeof(float) * (size_t)dim;
  LslPointBlock* blk = nullptr;
  if (tot) {
    blk = new (std::nothrow) LslPointBlock();
    if (!blk) return LSL_ERR_ARG;
    blk->refs = 0; blk->d_xyz1 = nullptr; blk->d_desc = nullptr; blk->d_kp = nullptr; blk->pooled = false;
    LSL_CUDA(cudaMalloc((void**)&blk->d_xyz1, sizeof(float) * 4 * tot));
    LSL_CUDA(cudaMalloc((void**)&blk->d_desc, row * tot));
    LSL_CUDA(cudaMemcpyAsync(blk->d_xyz1, xyz1, sizeof(float) * 4 * tot, cudaMemcpyHostToDevice, ctx->stream));
    LSL_CUDA(cudaMemcpyAsync(blk->d_desc, desc, row * tot, cudaMemcpyHostToDevice, ctx->stream));
    if (root_sift && !desc_is_u8) { int rc = lsl_launch_rootsift(ctx, (float*)blk->d_desc, (int)tot, dim); if (rc) return rc; }
    ctx->stats.h2d_bytes += sizeof(float) * 4 * tot + row * tot;
  }
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    lsl_frame* f = frames[i];
    release_points(f);
    f->npoints = counts[i]; f->pdim = dim; f->pkind = desc_is_u8 ? 1 : 0;
    if (counts[i]) {
      f->pblk = blk; blk->refs += 1;
      f->d_xyz1 = blk->d_xyz1 + 4 * off;
      f->d_desc = (float*)((uint8_t*)blk->d_desc + row * off);
    }
    off += (size_t)counts[i];
  }
  if (tot) LSL_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSL_OK;
}

// computeRelativeMotion_Ransac + optimizeRelmotion for EVERY pair of the last lsl_match_pair_batch, on the line matches that
// call left on the device (BASELINE config 5: "levmar refine per edge"): one launch, one CTA per pair.
extern "C" int lsl_relmotion_batch(lsl_ctx* ctx, int npairs, double* Rt /* [npairs][12]: R row-major, t */, int32_t* info /* [npairs][4]: consensus size, LM calls, have, 0 */) {
  if (!ctx || npairs < 1 || !Rt || !info) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  if (ctx->pair_inflight) { ctx->err = "a pair batch is in flight: call lsl_match_pair_batch_end first"; return LSL_ERR_BUSY; }
  if ((size_t)npairs > ctx->pw.h_pairs.size()) { ctx->err = "no pair batch of that size on the device"; return LSL_ERR_ARG; }
  cudaStream_t st = ctx->stream;
  const size_t M = ctx->pw.last_tot_m > 0 ? ctx->pw.last_tot_m : 1, it = ctx->P.ransac_iters_line_motion, np = (size_t)npairs;
  const size_t bytes = align_up(M * RM_STRIDE * 8) + align_up(M * 4 * 8) + align_up(M * RM_M * 8) + align_up(M * 4) + align_up(M * 2 * 4) +
                       align_up(np * it * 12 * 8) + align_up(np * it * 4) + align_up(np * it * 3 * 2) + align_up(np * 12 * 8) + align_up(np * 16);
  uint8_t* blk = nullptr;
  LSL_CUDA(cudaMallocAsync((void**)&blk, bytes, st));
  RmScratch rs;
  size_t off = 0;
  rs.g = (double*)(blk + off); off += align_up(M * RM_STRIDE * 8);
  rs.hx = (double*)(blk + off); off += align_up(M * 4 * 8);
  rs.jac = (double*)(blk + off); off += align_up(M * RM_M * 8);
  rs.flag = (int32_t*)(blk + off); off += align_up(M * 4);
  rs.cur = (int32_t*)(blk + off); off += align_up(M * 2 * 4);
  rs.hyp = (double*)(blk + off); off += align_up(np * it * 12 * 8);
  rs.cnts = (int32_t*)(blk + off); off += align_up(np * it * 4);
  rs.trip = (uint16_t*)(blk + off); off += align_up(np * it * 3 * 2);
  rs.outRt = (double*)(blk + off); off += align_up(np * 12 * 8);
  rs.outn = (int32_t*)(blk + off);
  rs.max_iter = (int)it;
  int rc = lsl_launch_relmotion(ctx, npairs, rs);
  if (rc) { cudaFreeAsync(blk, st); return rc; }
  LSL_CUDA(cudaMemcpyAsync(Rt, rs.outRt, np * 12 * 8, cudaMemcpyDeviceToHost, st));
  LSL_CUDA(cudaMemcpyAsync(info, rs.outn, np * 16, cudaMemcpyDeviceToHost, st));
  LSL_CUDA(cudaFreeAsync(blk, st));
  LSL_CUDA(cudaStreamSynchronize(st));
  collect_ktimes(ctx, LSL_K_RELMOTION, LSL_K_RELMOTION + 1);
  ctx->stats.d2h_bytes += np * (12 * 8 + 16);
  return LSL_OK;
}

extern "C" int lsl_frame_descriptors(lsl_ctx* ctx, const lsl_frame* f, void* dst, int64_t cap_bytes) {
  if (!ctx || !f || (f->npoints && !dst)) return LSL_ERR_ARG;
  const size_t bytes = (f->pkind ? (size_t)f->pdim : sizeof(float) * (size_t)f->pdim) * (size_t)f->npoints;
  if ((int64_t)bytes > cap_bytes) return LSL_ERR_CAPACITY;
  if (!bytes) return LSL_OK;
  LSL_ENTER(ctx);
  LSL_CUDA(cudaMemcpyAsync(dst, f->d_desc, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  LSL_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stats.d2h_bytes += bytes;
  return LSL_OK;
}
// Point detector of the context: kind 0 = none (point features enter through lsl_frame_set_points), 1 = SIFT on the device.
extern "C" int lsl_ctx_set_point_detector(lsl_ctx* ctx, int kind, int max_keypoints, int root_sift) {
  if (!ctx || kind < 0 || kind > 1 || max_keypoints < 1 || max_keypoints > LSL_MAX_POINTS) return LSL_ERR_ARG;
  LSL_LOCK(ctx);
  ctx->sift_kind = kind; ctx->sift_max_kp = max_keypoints; ctx->sift_root = root_sift ? 1 : 0;
  return LSL_OK;
}
// xyz1 [n][4], descriptor rows [n][dim] (f32 rows only) and, for device-detected points, kp [n][6] back to the host
extern "C" int lsl_frame_points(lsl_ctx* ctx, const lsl_frame* f, float* xyz1, float* desc, float* kp, int cap, int* n) {
  if (!ctx || !f || !n) return LSL_ERR_ARG;
  *n = f->npoints;
  if (f->npoints > cap) return LSL_ERR_CAPACITY;
  if (!f->npoints) return LSL_OK;
  if (f->pkind != 0 && desc) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  const size_t np = (size_t)f->npoints;
  if (xyz1) LSL_CUDA(cudaMemcpyAsync(xyz1, f->d_xyz1, sizeof(float) * 4 * np, cudaMemcpyDeviceToHost, ctx->stream));
  if (desc) LSL_CUDA(cudaMemcpyAsync(desc, f->d_desc, sizeof(float) * (size_t)f->pdim * np, cudaMemcpyDeviceToHost, ctx->stream));
  if (kp) {
    if (!f->d_kp) return LSL_ERR_ARG;
    LSL_CUDA(cudaMemcpyAsync(kp, f->d_kp, sizeof(float) * 6 * np, cudaMemcpyDeviceToHost, ctx->stream));
  }
  LSL_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSL_OK;
}
extern "C" int lsl_compute_inliers_and_error(lsl_ctx* ctx, const lsl_frame* query, const lsl_frame* train, const lsl_match* matches, int n,
                                             const float tf[16], double squared_max_inlier_dist, lsl_match* inliers, int cap, int* n_inliers,
                                             double* rmse) {
  if (!ctx || !query || !train || !tf || !n_inliers || !rmse || n < 0 || n > LSL_MAX_POINTS || (n && !matches)) return LSL_ERR_ARG;
  *n_inliers = 0; *rmse = 1e9;
  if (n == 0) return LSL_OK;                                   // (the reference asserts a non-empty list)
  for (int i = 0; i < n; ++i)
    if (matches[i].queryIdx < 0 || matches[i].queryIdx >= query->npoints || matches[i].trainIdx < 0 || matches[i].trainIdx >= train->npoints) {
      ctx->err = "point match index outside the frames' point tables";
      return LSL_ERR_ARG;
    }
  LSL_ENTER(ctx);
  cudaStream_t st = ctx->stream;
  const size_t mb = align_up(sizeof(lsl_match) * (size_t)n), db = align_up(sizeof(double) * (size_t)n);
  uint8_t* blk = nullptr;
  LSL_CUDA(cudaMallocAsync((void**)&blk, 2 * mb + db + 512, st));
  lsl_match* d_ms = (lsl_match*)blk; lsl_match* d_out = (lsl_match*)(blk + mb);
  double* d_dist = (double*)(blk + 2 * mb); double* d_rmse = (double*)(blk + 2 * mb + db);
  int32_t* d_n = (int32_t*)(d_rmse + 1); float* d_tf = (float*)(d_rmse + 2);
  LSL_CUDA(cudaMemcpyAsync(d_ms, matches, sizeof(lsl_match) * (size_t)n, cudaMemcpyHostToDevice, st));
  LSL_CUDA(cudaMemcpyAsync(d_tf, tf, 16 * sizeof(float), cudaMemcpyHostToDevice, st));
  int rc = lsl_launch_inliers_error(ctx, query->d_xyz1, train->d_xyz1, d_ms, n, d_tf, squared_max_inlier_dist, d_dist, d_out, d_n, d_rmse);
  if (rc) { cudaFreeAsync(blk, st); return rc; }
  int32_t k = 0;
  LSL_CUDA(cudaMemcpyAsync(&k, d_n, sizeof(k), cudaMemcpyDeviceToHost, st));
  LSL_CUDA(cudaMemcpyAsync(rmse, d_rmse, sizeof(double), cudaMemcpyDeviceToHost, st));
  LSL_CUDA(cudaStreamSynchronize(st));
  *n_inliers = k;
  int ret = LSL_OK;
  if (k > cap || (k && !inliers)) ret = LSL_ERR_CAPACITY;
  else if (k) { LSL_CUDA(cudaMemcpyAsync(inliers, d_out, sizeof(lsl_match) * (size_t)k, cudaMemcpyDeviceToHost, st)); LSL_CUDA(cudaStreamSynchronize(st)); }
  cudaFreeAsync(blk, st);
  ctx->stats.h2d_bytes += sizeof(lsl_match) * (size_t)n + 64; ctx->stats.d2h_bytes += sizeof(lsl_match) * (size_t)k + 12;
  return ret;
}
extern "C" int lsl_frame_num_points(const lsl_frame* f) { return f ? f->npoints : LSL_ERR_ARG; }
extern "C" int lsl_ctx_set_camera(lsl_ctx* ctx, double fx, double asynch_dt_s) {
  if (!ctx || !(fx > 0)) return LSL_ERR_ARG;
  LSL_LOCK(ctx);
  ctx->cam_fx = fx; ctx->cam_dt = asynch_dt_s;
  return LSL_OK;
}

// parity-test read-back of per-line intermediates (inlier sample indices of the 3D-line RANSAC etc.)
extern "C" int lsl_frame_debug(const lsl_frame* f, int32_t* npts, int32_t* inl_idx, int32_t* seg_of_line, int32_t* lm_iters) {
  if (!f || !f->have_dbg) return LSL_ERR_ARG;
  if (npts) memcpy(npts, f->dbg_npts.data(), 4 * f->dbg_npts.size());
  if (inl_idx) memcpy(inl_idx, f->dbg_inl.data(), 4 * f->dbg_inl.size());
  if (seg_of_line) memcpy(seg_of_line, f->dbg_seg.data(), 4 * f->dbg_seg.size());
  if (lm_iters) memcpy(lm_iters, f->dbg_lm.data(), 4 * f->dbg_lm.size());
  return LSL_OK;
}

// ------------------------------------------------------------- pair registration ----
template <class T>
static cudaError_t regrow(T** p, size_t count) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  return cudaMalloc((void**)p, sizeof(T) * (count ? count : 1));
}
// Sizes the pair workspace for npairs pairs with tot_m match slots and tot_d distance-matrix entries.
static int ensure_pair_ws(lsl_ctx* ctx, size_t npairs, size_t tot_m, size_t tot_d) {
  LslPairWork& p = ctx->pw;
  const int max_iter = ctx->P.ransac_iters_line_motion;
  if (max_iter < 1 || max_iter > 65535) { ctx->err = "ransac_iters_line_motion out of range"; return LSL_ERR_ARG; }
  if (npairs > p.cap_pairs || p.sc.max_iter != max_iter) {
    size_t c = npairs + npairs / 2;
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    LSL_CUDA(regrow(&p.d_pairs, c)); LSL_CUDA(regrow(&p.nmatch, c)); LSL_CUDA(regrow(&p.recs, c));
    LSL_CUDA(regrow(&p.sc.tfs, c * max_iter * 12)); LSL_CUDA(regrow(&p.sc.cnts, c * max_iter));
    LSL_CUDA(regrow(&p.sc.trip, c * max_iter * 3)); LSL_CUDA(regrow(&p.sc.n_inl, c)); LSL_CUDA(regrow(&p.sc.n_rinl, c));
    LSL_CUDA(regrow(&p.sc.tf_ransac, c * 16));
    p.cap_pairs = c; p.sc.max_iter = max_iter;
  }
  if (tot_m > p.cap_m) {
    size_t c = tot_m * 2;
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    LSL_CUDA(regrow(&p.matches, c)); LSL_CUDA(regrow(&p.sc.md, c * 72)); LSL_CUDA(regrow(&p.sc.dab, c * 2));
    LSL_CUDA(regrow(&p.sc.sel, c * 3)); LSL_CUDA(regrow(&p.sc.lm, c * 306)); LSL_CUDA(regrow(&p.sc.okf, c));
    p.cap_m = c;
  }
  if (tot_d > p.cap_d) {
    size_t c = tot_d * 2;
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    LSL_CUDA(regrow(&p.D, c));
    p.cap_d = c;
  }
  return LSL_OK;
}

// Fills the descriptors of a batch (host + device). cap_override >= 0 forces the match capacity (pose-only calls).
static int setup_pairs(lsl_ctx* ctx, int npairs, const lsl_frame* const* queries, const lsl_frame* const* trains,
                       const int32_t* id_query, const int32_t* id_train, const uint32_t* seeds, const int* adjacent,
                       int cap_override) {
  if (ctx->pair_inflight) { ctx->err = "a pair batch is in flight: call lsl_match_pair_batch_end first"; return LSL_ERR_BUSY; }
  LslPairWork& p = ctx->pw;
  p.h_pairs.resize(npairs);
  size_t m_off = 0, d_off = 0;
  for (int i = 0; i < npairs; ++i) {
    const lsl_frame* q = queries[i];
    const lsl_frame* t = trains[i];
    if (!q || !t) return LSL_ERR_ARG;
    LslPairDesc& d = p.h_pairs[i];
    d.q = q->d_lines; d.t = t->d_lines; d.nq = q->nlines; d.nt = t->nlines;
    d.id_q = id_query ? id_query[i] : 1; d.id_t = id_train ? id_train[i] : 0;
    d.seed = seeds ? seeds[i] : 1u;
    d.adjacent = adjacent ? adjacent[i] : (abs(d.id_q - d.id_t) <= ctx->P.adjacent_linematch_window);  // node.cpp:1505-1507
    d.cap_m = cap_override >= 0 ? cap_override : (d.nq < d.nt ? d.nq : d.nt);
    d.d_off = d_off; d.m_off = m_off; d.pad_ = 0;
    m_off += (size_t)(d.cap_m > 0 ? d.cap_m : 1);
    d_off += (size_t)d.nq * d.nt;
  }
  p.last_tot_m = m_off;
  int rc = ensure_pair_ws(ctx, npairs, m_off, d_off);
  if (rc) return rc;
  LSL_CUDA(cudaMemcpyAsync(p.d_pairs, p.h_pairs.data(), sizeof(LslPairDesc) * npairs, cudaMemcpyHostToDevice, ctx->stream));
  return LSL_OK;
}

// Point side of a batch: descriptors of the pairs' point sets and the point-match-sized scratch.
static int setup_ppairs(lsl_ctx* ctx, int npairs, const lsl_frame* const* queries, const lsl_frame* const* trains,
                        int cap_override, int* max_nq, int* dim_out, int* kind_out = nullptr) {
  LslHybWork& h = ctx->hw;
  h.h_ppairs.resize(npairs);
  size_t pm_off = 0, knn_off = 0;
  int mq = 0, dim = 0, kind = 0;
  for (int i = 0; i < npairs; ++i) {
    const lsl_frame* q = queries[i];
    const lsl_frame* t = trains[i];
    LslPairPts& d = h.h_ppairs[i];
    d.qx = q->d_xyz1; d.tx = t->d_xyz1; d.qd = q->d_desc; d.td = t->d_desc;
    d.nqp = q->npoints; d.ntp = t->npoints;
    if (d.nqp && d.ntp && q->pdim != t->pdim) { ctx->err = "descriptor dimensions of the two frames differ"; return LSL_ERR_ARG; }
    d.dim = d.nqp ? q->pdim : t->pdim;
    d.kind = d.nqp ? q->pkind : t->pkind; d.pad_ = 0;
    if (d.nqp && d.ntp && q->pkind != t->pkind) { ctx->err = "descriptor types of the two frames differ"; return LSL_ERR_ARG; }
    if (cap_override < 0 && d.nqp && d.ntp) {
      if (dim && (d.dim != dim || d.kind != kind)) { ctx->err = "mixed descriptor dimensions / types in one batch"; return LSL_ERR_ARG; }
      dim = d.dim; kind = d.kind;
    }
    d.cap_pm = cap_override >= 0 ? cap_override : d.nqp;
    d.pm_off = pm_off; d.knn_off = knn_off;
    pm_off += (size_t)(d.cap_pm > 0 ? d.cap_pm : 1);
    knn_off += (size_t)(d.nqp > 0 ? d.nqp : 1);
    if (d.nqp > mq && d.ntp >= 2) mq = d.nqp;
  }
  const int max_iter = ctx->P.ransac_iters_line_motion;
  if ((size_t)npairs > h.cap_pairs || h.max_iter != max_iter) {
    size_t c = (size_t)npairs + npairs / 2;
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    LSL_CUDA(regrow(&h.d_ppairs, c)); LSL_CUDA(regrow(&h.npmatch, c)); LSL_CUDA(regrow(&h.hs.ptidx, c * max_iter * 2));
    LSL_CUDA(regrow(&h.hs.rng, c * 33)); LSL_CUDA(regrow(&h.hs.n_pinl, c)); LSL_CUDA(regrow(&h.hs.n_prinl, c));
    h.cap_pairs = c; h.max_iter = max_iter;
  }
  if (pm_off > h.cap_pm) {
    size_t c = pm_off + pm_off / 2;
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    LSL_CUDA(regrow(&h.pmatches, c)); LSL_CUDA(regrow(&h.hs.pmd, c * 22)); LSL_CUDA(regrow(&h.hs.pd2, c));
    LSL_CUDA(regrow(&h.hs.psel, c * 3)); LSL_CUDA(regrow(&h.hs.plm, c * 160)); LSL_CUDA(regrow(&h.hs.pokf, c));
    h.cap_pm = c;
  }
  if (knn_off > h.cap_knn) {
    size_t c = knn_off + knn_off / 2;
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h.knn) cudaFree(h.knn);
    h.knn = nullptr;
    LSL_CUDA(cudaMalloc(&h.knn, 12 * c));
    h.cap_knn = c;
  }
  LSL_CUDA(cudaMemcpyAsync(h.d_ppairs, h.h_ppairs.data(), sizeof(LslPairPts) * npairs, cudaMemcpyHostToDevice, ctx->stream));
  if (max_nq) *max_nq = mq;
  if (dim_out) *dim_out = dim ? dim : 1;
  if (kind_out) *kind_out = kind;
  return LSL_OK;
}
// seeds the per-pair rand() state on the host (pose-only calls have no featureMatching pass before them)
static int upload_fresh_rng(lsl_ctx* ctx, uint32_t seed, int skip) {
  lslm::GRand g;
  lslm::grand_seed(&g, seed);
  for (int i = 0; i < skip; ++i) lslm::grand_next(&g);
  int32_t st[33];
  for (int k = 0; k < 31; ++k) st[k] = g.r[k];
  st[31] = g.f; st[32] = g.b;
  LSL_CUDA(cudaMemcpyAsync(ctx->hw.hs.rng, st, sizeof(st), cudaMemcpyHostToDevice, ctx->stream));
  LSL_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSL_OK;
}
static int fetch_counts_hyb(lsl_ctx* ctx, int npairs) {
  LslHybWork& h = ctx->hw;
  h.h_npmatch.resize(npairs); h.h_npinl.resize(npairs); h.h_nprinl.resize(npairs);
  LSL_CUDA(cudaMemcpyAsync(h.h_npmatch.data(), h.npmatch, 4 * npairs, cudaMemcpyDeviceToHost, ctx->stream));
  LSL_CUDA(cudaMemcpyAsync(h.h_npinl.data(), h.hs.n_pinl, 4 * npairs, cudaMemcpyDeviceToHost, ctx->stream));
  LSL_CUDA(cudaMemcpyAsync(h.h_nprinl.data(), h.hs.n_prinl, 4 * npairs, cudaMemcpyDeviceToHost, ctx->stream));
  return LSL_OK;
}

static int fetch_counts(lsl_ctx* ctx, int npairs) {
  LslPairWork& p = ctx->pw;
  p.h_nmatch.resize(npairs); p.h_ninl.resize(npairs); p.h_nrinl.resize(npairs);
  LSL_CUDA(cudaMemcpyAsync(p.h_nmatch.data(), p.nmatch, 4 * npairs, cudaMemcpyDeviceToHost, ctx->stream));
  LSL_CUDA(cudaMemcpyAsync(p.h_ninl.data(), p.sc.n_inl, 4 * npairs, cudaMemcpyDeviceToHost, ctx->stream));
  LSL_CUDA(cudaMemcpyAsync(p.h_nrinl.data(), p.sc.n_rinl, 4 * npairs, cudaMemcpyDeviceToHost, ctx->stream));
  return LSL_OK;
}

extern "C" int lsl_match_lines(lsl_ctx* ctx, const lsl_frame* query, const lsl_frame* train, int adjacent, lsl_match* out,
                               int cap, int* n) {
  if (!ctx || !query || !train || !n) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  int adj = adjacent ? 1 : 0;
  int rc = setup_pairs(ctx, 1, &query, &train, nullptr, nullptr, nullptr, &adj, -1);
  if (rc) return rc;
  clear_ktimes(ctx, LSL_K_MATCH, LSL_K_PNG);
  if ((rc = lsl_launch_match(ctx, 1))) return rc;
  int32_t nm = 0;
  LSL_CUDA(cudaMemcpyAsync(&nm, ctx->pw.nmatch, 4, cudaMemcpyDeviceToHost, ctx->stream));
  LSL_CUDA(cudaStreamSynchronize(ctx->stream));
  collect_ktimes(ctx, LSL_K_MATCH, LSL_K_PNG);
  *n = nm;
  ctx->stats.pairs += 1; ctx->stats.matches += nm;
  if (nm > cap || (nm && !out)) return LSL_ERR_CAPACITY;
  if (nm) {
    LSL_CUDA(cudaMemcpyAsync(out, ctx->pw.matches, sizeof(lsl_match) * nm, cudaMemcpyDeviceToHost, ctx->stream));
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += sizeof(lsl_match) * nm;
  }
  return LSL_OK;
}

// copies the index list `sel` (indices into the pair's match list) back as matches
static int fetch_sel(lsl_ctx* ctx, const LslPairDesc& d, int which, int count, const std::vector<lsl_match>& all, lsl_match* out) {
  if (!count) return LSL_OK;
  std::vector<int32_t> idx(count);
  LSL_CUDA(cudaMemcpyAsync(idx.data(), ctx->pw.sc.sel + d.m_off * 3 + (size_t)which * d.cap_m, 4 * count, cudaMemcpyDeviceToHost, ctx->stream));
  LSL_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < count; ++i) out[i] = all[idx[i]];
  ctx->stats.d2h_bytes += 4 * count;
  return LSL_OK;
}

// Node::featureMatching, BRUTEFORCE branch (src/node.cpp:606-641) on the frames' point features.
extern "C" int lsl_match_points(lsl_ctx* ctx, const lsl_frame* query, const lsl_frame* train, uint32_t seed, lsl_match* out,
                                int cap, int* n) {
  if (!ctx || !query || !train || !n) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  int id_q = 1, id_t = 0;
  int rc = setup_pairs(ctx, 1, &query, &train, &id_q, &id_t, &seed, nullptr, -1);
  if (rc) return rc;
  int mq = 0, dim = 1, kind = 0;
  if ((rc = setup_ppairs(ctx, 1, &query, &train, -1, &mq, &dim, &kind))) return rc;
  clear_ktimes(ctx, LSL_K_MATCH, LSL_K_PNG);
  if ((rc = lsl_launch_match_points(ctx, 1, mq, dim, kind))) return rc;
  int32_t nm = 0;
  LSL_CUDA(cudaMemcpyAsync(&nm, ctx->hw.npmatch, 4, cudaMemcpyDeviceToHost, ctx->stream));
  LSL_CUDA(cudaStreamSynchronize(ctx->stream));
  collect_ktimes(ctx, LSL_K_MATCH, LSL_K_PNG);
  *n = nm;
  ctx->stats.pairs += 1; ctx->stats.matches += nm;
  if (nm > cap || (nm && !out)) return LSL_ERR_CAPACITY;
  if (nm) {
    LSL_CUDA(cudaMemcpyAsync(out, ctx->hw.pmatches, sizeof(lsl_match) * nm, cudaMemcpyDeviceToHost, ctx->stream));
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += sizeof(lsl_match) * nm;
  }
  return LSL_OK;
}

static int fetch_psel(lsl_ctx* ctx, const LslPairPts& d, int which, int count, const std::vector<lsl_match>& all, lsl_match* out) {
  if (!count) return LSL_OK;
  std::vector<int32_t> idx(count);
  LSL_CUDA(cudaMemcpyAsync(idx.data(), ctx->hw.hs.psel + d.pm_off * 3 + (size_t)which * d.cap_pm, 4 * count, cudaMemcpyDeviceToHost, ctx->stream));
  LSL_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < count; ++i) out[i] = all[idx[i]];
  ctx->stats.d2h_bytes += 4 * count;
  return LSL_OK;
}

extern "C" int lsl_pose_ransac(lsl_ctx* ctx, const lsl_frame* train, const lsl_frame* query, int id_train, int id_query,
                               const lsl_match* pt_matches, int npt, const lsl_match* ln_matches, int nln, uint32_t seed,
                               lsl_pose_rec* rec, lsl_match* inliers_out, int cap, int* n_inl, lsl_match* ransac_inliers_out,
                               int cap2, int* n_rinl) {
  if (!ctx || !train || !query || !rec || nln < 0 || (nln && !ln_matches) || npt < 0 || (npt && !pt_matches)) return LSL_ERR_ARG;
  if (nln > LSL_MAX_MATCH || npt > LSL_MAX_POINTS) return LSL_ERR_CAPACITY;
  for (int i = 0; i < nln; ++i)
    if (ln_matches[i].queryIdx < 0 || ln_matches[i].queryIdx >= query->nlines || ln_matches[i].trainIdx < 0 ||
        ln_matches[i].trainIdx >= train->nlines) return LSL_ERR_ARG;
  for (int i = 0; i < npt; ++i)
    if (pt_matches[i].queryIdx < 0 || pt_matches[i].queryIdx >= query->npoints || pt_matches[i].trainIdx < 0 ||
        pt_matches[i].trainIdx >= train->npoints) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  int rc = setup_pairs(ctx, 1, &query, &train, &id_query, &id_train, &seed, nullptr, nln);
  if (rc) return rc;
  const bool hybrid = npt > 0;
  ctx->hw.last_hybrid = hybrid;
  if (hybrid) {
    if ((rc = setup_ppairs(ctx, 1, &query, &train, npt, nullptr, nullptr))) return rc;
    int32_t np = npt;
    LSL_CUDA(cudaMemcpyAsync(ctx->hw.npmatch, &np, 4, cudaMemcpyHostToDevice, ctx->stream));
    LSL_CUDA(cudaMemcpyAsync(ctx->hw.pmatches, pt_matches, sizeof(lsl_match) * npt, cudaMemcpyHostToDevice, ctx->stream));
    // matchNodePair order: featureMatching has consumed one rand() per point match before the RANSAC draws
    if ((rc = upload_fresh_rng(ctx, seed, npt))) return rc;
    ctx->stats.h2d_bytes += sizeof(lsl_match) * npt;
  }
  int32_t nm = nln;
  LSL_CUDA(cudaMemcpyAsync(ctx->pw.nmatch, &nm, 4, cudaMemcpyHostToDevice, ctx->stream));
  if (nln) LSL_CUDA(cudaMemcpyAsync(ctx->pw.matches, ln_matches, sizeof(lsl_match) * nln, cudaMemcpyHostToDevice, ctx->stream));
  ctx->stats.h2d_bytes += sizeof(lsl_match) * nln;
  clear_ktimes(ctx, LSL_K_MATCH, LSL_K_PNG);
  if (hybrid) { if ((rc = lsl_launch_pose_hybrid(ctx, 1, ctx->cam_fx, ctx->cam_dt))) return rc; }
  else if ((rc = lsl_launch_pose(ctx, 1))) return rc;
  if ((rc = fetch_counts(ctx, 1))) return rc;
  if (hybrid && (rc = fetch_counts_hyb(ctx, 1))) return rc;
  LSL_CUDA(cudaMemcpyAsync(rec, ctx->pw.recs, sizeof(lsl_pose_rec), cudaMemcpyDeviceToHost, ctx->stream));
  LSL_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->pw.h_nmatch[0] = nln;
  if (hybrid) ctx->hw.h_npmatch[0] = npt;
  collect_ktimes(ctx, LSL_K_MATCH, LSL_K_PNG);
  ctx->stats.pairs += 1; ctx->stats.d2h_bytes += sizeof(lsl_pose_rec);
  std::vector<lsl_match> all(ln_matches, ln_matches + nln);
  int ni = ctx->pw.h_ninl[0], nr = ctx->pw.h_nrinl[0];
  if (n_inl) *n_inl = ni;
  if (n_rinl) *n_rinl = nr;
  if (inliers_out) { if (ni > cap) return LSL_ERR_CAPACITY; if ((rc = fetch_sel(ctx, ctx->pw.h_pairs[0], 1, ni, all, inliers_out))) return rc; }
  if (ransac_inliers_out) { if (nr > cap2) return LSL_ERR_CAPACITY; if ((rc = fetch_sel(ctx, ctx->pw.h_pairs[0], 0, nr, all, ransac_inliers_out))) return rc; }
  return LSL_OK;
}

// computeRelativeMotion_Ransac (src/line/motion.cpp:367-526) + optimizeRelmotion (motion.cpp:98-139) on the
// matched line pairs of two frames: a = query lines, b = train lines, x_b = R x_a + t.
extern "C" int lsl_relmotion_ransac(lsl_ctx* ctx, const lsl_frame* train, const lsl_frame* query, const lsl_match* ln_matches,
                                    int nln, uint32_t seed, double R[9], double t[3], int32_t* conset, int cap, int* n_conset,
                                    int* lm_calls, int* have) {
  if (!ctx || !train || !query || nln < 0 || (nln && !ln_matches) || !R || !t || !n_conset) return LSL_ERR_ARG;
  if (nln > LSL_MAX_MATCH) return LSL_ERR_CAPACITY;
  for (int i = 0; i < nln; ++i)
    if (ln_matches[i].queryIdx < 0 || ln_matches[i].queryIdx >= query->nlines || ln_matches[i].trainIdx < 0 ||
        ln_matches[i].trainIdx >= train->nlines) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  int idq = 1, idt = 0;
  int rc = setup_pairs(ctx, 1, &query, &train, &idq, &idt, &seed, nullptr, nln);
  if (rc) return rc;
  int32_t nm = nln;
  cudaStream_t st = ctx->stream;
  LSL_CUDA(cudaMemcpyAsync(ctx->pw.nmatch, &nm, 4, cudaMemcpyHostToDevice, st));
  if (nln) LSL_CUDA(cudaMemcpyAsync(ctx->pw.matches, ln_matches, sizeof(lsl_match) * nln, cudaMemcpyHostToDevice, st));
  const size_t M = nln > 0 ? nln : 1, it = ctx->P.ransac_iters_line_motion;
  const size_t bytes = align_up(M * RM_STRIDE * 8) + align_up(M * 4 * 8) + align_up(M * RM_M * 8) + align_up(M * 4) + align_up(M * 2 * 4) +
                       align_up(it * 12 * 8) + align_up(it * 4) + align_up(it * 3 * 2) + align_up(12 * 8) + align_up(16);
  uint8_t* blk = nullptr;
  LSL_CUDA(cudaMallocAsync((void**)&blk, bytes, st));
  RmScratch rs;
  size_t off = 0;
  rs.g = (double*)(blk + off); off += align_up(M * RM_STRIDE * 8);
  rs.hx = (double*)(blk + off); off += align_up(M * 4 * 8);
  rs.jac = (double*)(blk + off); off += align_up(M * RM_M * 8);
  rs.flag = (int32_t*)(blk + off); off += align_up(M * 4);
  rs.cur = (int32_t*)(blk + off); off += align_up(M * 2 * 4);
  rs.hyp = (double*)(blk + off); off += align_up(it * 12 * 8);
  rs.cnts = (int32_t*)(blk + off); off += align_up(it * 4);
  rs.trip = (uint16_t*)(blk + off); off += align_up(it * 3 * 2);
  rs.outRt = (double*)(blk + off); off += align_up(12 * 8);
  rs.outn = (int32_t*)(blk + off);
  rs.max_iter = (int)it;
  clear_ktimes(ctx, LSL_K_MATCH, LSL_K_PNG);
  if ((rc = lsl_launch_relmotion(ctx, 1, rs))) { cudaFreeAsync(blk, st); return rc; }
  double Rt[12];
  int32_t on[4];
  LSL_CUDA(cudaMemcpyAsync(Rt, rs.outRt, sizeof(Rt), cudaMemcpyDeviceToHost, st));
  LSL_CUDA(cudaMemcpyAsync(on, rs.outn, sizeof(on), cudaMemcpyDeviceToHost, st));
  LSL_CUDA(cudaStreamSynchronize(st));
  collect_ktimes(ctx, LSL_K_MATCH, LSL_K_PNG);
  *n_conset = on[0];
  if (lm_calls) *lm_calls = on[1];
  if (have) *have = on[2];
  if (on[2]) { memcpy(R, Rt, 72); memcpy(t, Rt + 9, 24); }
  int rcap = LSL_OK;
  if (on[0] > cap || (on[0] && !conset)) rcap = LSL_ERR_CAPACITY;
  else if (on[0]) LSL_CUDA(cudaMemcpyAsync(conset, rs.cur, 4 * on[0], cudaMemcpyDeviceToHost, st));
  LSL_CUDA(cudaFreeAsync(blk, st));
  LSL_CUDA(cudaStreamSynchronize(st));
  ctx->stats.pairs += 1; ctx->stats.d2h_bytes += sizeof(Rt) + sizeof(on) + 4 * (size_t)on[0];
  return rcap;
}

// Runs `body` with the context's launches, copies and kernel timers directed to the pair stream.
struct PairStreamScope {
  lsl_ctx* c; cudaStream_t saved;
  explicit PairStreamScope(lsl_ctx* ctx) : c(ctx), saved(ctx->stream) { ctx->stream = ctx->pair_stream; }
  ~PairStreamScope() { c->stream = saved; }
};

extern "C" int lsl_match_pair_batch_begin(lsl_ctx* ctx, int npairs, const lsl_frame* const* queries, const lsl_frame* const* trains,
                                          const int32_t* id_query, const int32_t* id_train, const uint32_t* seeds) {
  if (!ctx || npairs < 1 || !queries || !trains) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  if (ctx->pair_inflight) { ctx->err = "a pair batch is in flight: call lsl_match_pair_batch_end first"; return LSL_ERR_BUSY; }
  // the frames' records were written on the context stream (extract calls return synchronised; uploads of
  // lsl_frame_from_lines / set_points are synchronised too), so the pair stream needs no event to see them
  PairStreamScope scope(ctx);
  int rc = setup_pairs(ctx, npairs, queries, trains, id_query, id_train, seeds, nullptr, -1);
  if (rc) return rc;
  bool hybrid = false;   // any frame with point features -> Node::matchNodePair with both modalities
  for (int i = 0; i < npairs; ++i) hybrid = hybrid || (queries[i]->npoints > 0 && trains[i]->npoints > 0);
  ctx->hw.last_hybrid = hybrid;
  int mq = 0, dim = 1, kind = 0;
  if (hybrid && (rc = setup_ppairs(ctx, npairs, queries, trains, -1, &mq, &dim, &kind))) return rc;
  clear_ktimes(ctx, LSL_K_MATCH, LSL_K_PNG);
  cudaEventRecord(ctx->ev0p, ctx->stream);
  rc = LSL_OK;
  if (hybrid) rc = lsl_launch_match_points(ctx, npairs, mq, dim, kind);   // featureMatching first (node.cpp:1504)
  if (!rc) rc = lsl_launch_match(ctx, npairs);
  if (!rc) rc = hybrid ? lsl_launch_pose_hybrid(ctx, npairs, ctx->cam_fx, ctx->cam_dt) : lsl_launch_pose(ctx, npairs);
  if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }   // nothing of a failed batch keeps running on the pair workspace
  cudaEventRecord(ctx->ev3p, ctx->stream);
  ctx->pair_inflight = npairs; ctx->pair_hybrid = hybrid;
  return LSL_OK;
}

extern "C" int lsl_match_pair_batch_end(lsl_ctx* ctx, lsl_pose_rec* out, int cap) {
  if (!ctx || !out) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  const int npairs = ctx->pair_inflight;
  if (!npairs) { ctx->err = "no pair batch in flight"; return LSL_ERR_ARG; }
  if (cap < npairs) return LSL_ERR_CAPACITY;     // the batch stays in flight: call again with room for npairs records
  ctx->pair_inflight = 0;                        // consumed whatever happens below (a CUDA error must not leave the context fenced)
  PairStreamScope scope(ctx);
  int rc;
  if ((rc = fetch_counts(ctx, npairs))) return rc;
  if (ctx->pair_hybrid && (rc = fetch_counts_hyb(ctx, npairs))) return rc;
  LSL_CUDA(cudaMemcpyAsync(out, ctx->pw.recs, sizeof(lsl_pose_rec) * npairs, cudaMemcpyDeviceToHost, ctx->stream));
  LSL_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&ctx->ms_total, ctx->ev0p, ctx->ev3p);
  collect_ktimes(ctx, LSL_K_MATCH, LSL_K_PNG);
  ctx->stats.pairs += npairs; ctx->stats.d2h_bytes += (sizeof(lsl_pose_rec) + 12) * npairs;
  for (int i = 0; i < npairs; ++i) ctx->stats.matches += ctx->pw.h_nmatch[i];
  return LSL_OK;
}

extern "C" int lsl_match_pair_batch(lsl_ctx* ctx, int npairs, const lsl_frame* const* queries, const lsl_frame* const* trains,
                                    const int32_t* id_query, const int32_t* id_train, const uint32_t* seeds, lsl_pose_rec* out) {
  if (!ctx || npairs < 1 || !queries || !trains || !out) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  int rc = lsl_match_pair_batch_begin(ctx, npairs, queries, trains, id_query, id_train, seeds);
  if (rc) return rc;
  return lsl_match_pair_batch_end(ctx, out, npairs);
}

extern "C" int lsl_pair_matches(lsl_ctx* ctx, int pair, int what, lsl_match* out, int cap, int* n) {
  if (!ctx || !n || pair < 0 || pair >= (int)ctx->pw.h_nmatch.size() || what < 0 || what > 5) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  if (ctx->pair_inflight) { ctx->err = "a pair batch is in flight: call lsl_match_pair_batch_end first"; return LSL_ERR_BUSY; }
  if (what >= 3) {   // point lists: 3 all point matches, 4 refined point inliers, 5 point inliers of the best hypothesis
    const LslHybWork& h = ctx->hw;
    if (!h.last_hybrid || pair >= (int)h.h_npmatch.size()) { *n = 0; return LSL_OK; }
    const LslPairPts& d = h.h_ppairs[pair];
    int nm = h.h_npmatch[pair];
    int cnt = what == 3 ? nm : (what == 4 ? h.h_npinl[pair] : h.h_nprinl[pair]);
    *n = cnt;
    if (cnt > cap || (cnt && !out)) return LSL_ERR_CAPACITY;
    if (!cnt) return LSL_OK;
    std::vector<lsl_match> all(nm);
    LSL_CUDA(cudaMemcpyAsync(all.data(), h.pmatches + d.pm_off, sizeof(lsl_match) * nm, cudaMemcpyDeviceToHost, ctx->stream));
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (what == 3) { memcpy(out, all.data(), sizeof(lsl_match) * nm); return LSL_OK; }
    return fetch_psel(ctx, d, what == 4 ? 1 : 0, cnt, all, out);
  }
  const LslPairDesc& d = ctx->pw.h_pairs[pair];
  int nm = ctx->pw.h_nmatch[pair];
  int cnt = what == 0 ? nm : (what == 1 ? ctx->pw.h_ninl[pair] : ctx->pw.h_nrinl[pair]);
  *n = cnt;
  if (cnt > cap || (cnt && !out)) return LSL_ERR_CAPACITY;
  if (!cnt) return LSL_OK;
  std::vector<lsl_match> all(nm);
  LSL_CUDA(cudaMemcpyAsync(all.data(), ctx->pw.matches + d.m_off, sizeof(lsl_match) * nm, cudaMemcpyDeviceToHost, ctx->stream));
  LSL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (what == 0) { memcpy(out, all.data(), sizeof(lsl_match) * nm); return LSL_OK; }
  return fetch_sel(ctx, d, what == 1 ? 1 : 0, cnt, all, out);
}

// ------------------------------------------------------------------ pose exchange (NCCL) ----
// NCCL is resolved at run time (dlopen) so the library has no link-time dependency on it; the only
// collective of the path is the all-gather of fixed-size pose records at graph-insert time.
struct lsl_nccl_uid { char internal[128]; };  // ncclUniqueId
typedef int (*nccl_get_uid_fn)(lsl_nccl_uid*);
typedef int (*nccl_init_rank_fn)(void**, int, lsl_nccl_uid, int);
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*nccl_destroy_fn)(void*);
typedef const char* (*nccl_errstr_fn)(int);

static void* nccl_open(lsl_ctx* ctx) {
  if (ctx && ctx->nccl_lib) return ctx->nccl_lib;
  const char* env = getenv("LSL_NCCL_LIB");
  void* h = nullptr;
  if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (ctx) ctx->nccl_lib = h;
  return h;
}
extern "C" int lsl_comm_unique_id(void* id128) {
  if (!id128) return LSL_ERR_ARG;
  void* h = nccl_open(nullptr);
  if (!h) return LSL_ERR_NCCL;
  nccl_get_uid_fn f = (nccl_get_uid_fn)dlsym(h, "ncclGetUniqueId");
  if (!f || f((lsl_nccl_uid*)id128) != 0) return LSL_ERR_NCCL;
  return LSL_OK;
}
extern "C" int lsl_comm_init(lsl_ctx* ctx, const void* id128, int nranks, int rank) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  void* h = nccl_open(ctx);
  if (!h) { ctx->err = "libnccl not found (set LSL_NCCL_LIB)"; return LSL_ERR_NCCL; }
  nccl_init_rank_fn f = (nccl_init_rank_fn)dlsym(h, "ncclCommInitRank");
  if (!f) return LSL_ERR_NCCL;
  lsl_nccl_uid uid;
  memcpy(&uid, id128, 128);
  void* comm = nullptr;
  int rc = f(&comm, nranks, uid, rank);
  if (rc != 0) {
    nccl_errstr_fn es = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
    ctx->err = std::string("ncclCommInitRank: ") + (es ? es(rc) : "error");
    return LSL_ERR_NCCL;
  }
  ctx->nccl_comm = comm; ctx->nccl_rank = rank; ctx->nccl_nranks = nranks; ctx->nccl_own = true;
  return LSL_OK;
}
static void nccl_teardown(lsl_ctx* ctx) {
  if (ctx->nccl_comm && ctx->nccl_own && ctx->nccl_lib) {
    nccl_destroy_fn f = (nccl_destroy_fn)dlsym(ctx->nccl_lib, "ncclCommDestroy");
    if (f) f(ctx->nccl_comm);
  }
  ctx->nccl_comm = nullptr;
}
extern "C" int lsl_allgather_poses(lsl_ctx* ctx, void* nccl_comm, int nranks, const lsl_pose_rec* local_recs, int nlocal,
                                   lsl_pose_rec* all_recs) {
  if (!ctx || !all_recs || nlocal < 1) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  void* comm = nccl_comm ? nccl_comm : ctx->nccl_comm;
  if (!nccl_comm) nranks = ctx->nccl_nranks;
  if (!comm || nranks < 1) { ctx->err = "no communicator: call lsl_comm_init or pass an ncclComm_t"; return LSL_ERR_NCCL; }
  // local_recs == NULL: the records of the last lsl_match_pair_batch are gathered straight from the device buffer the
  // pose kernel wrote (no host round trip on the send side)
  if (!local_recs && ctx->pair_inflight) { ctx->err = "a pair batch is in flight: call lsl_match_pair_batch_end first"; return LSL_ERR_BUSY; }
  if (!local_recs && (size_t)nlocal > ctx->pw.h_pairs.size()) { ctx->err = "no device-resident records of that count"; return LSL_ERR_ARG; }
  void* h = nccl_open(ctx);
  if (!h) return LSL_ERR_NCCL;
  nccl_allgather_fn ag = (nccl_allgather_fn)dlsym(h, "ncclAllGather");
  if (!ag) return LSL_ERR_NCCL;
  const size_t lb = sizeof(lsl_pose_rec) * (size_t)nlocal;
  const size_t need = lb * (size_t)(nranks + 1);
  if (ctx->d_gather_bytes < need) {   // persistent exchange buffer (grown, never shrunk)
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_gather) cudaFree(ctx->d_gather);
    ctx->d_gather = nullptr; ctx->d_gather_bytes = 0;
    LSL_CUDA(cudaMalloc((void**)&ctx->d_gather, need * 2));
    ctx->d_gather_bytes = need * 2;
  }
  uint8_t* d = ctx->d_gather;
  const void* src = ctx->pw.recs;
  if (local_recs) { LSL_CUDA(cudaMemcpyAsync(d, local_recs, lb, cudaMemcpyHostToDevice, ctx->stream)); src = d; ctx->stats.h2d_bytes += lb; }
  int rc = ag(src, d + lb, lb, /* ncclChar */ 0, comm, ctx->stream);
  if (rc != 0) {
    nccl_errstr_fn es = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
    ctx->err = std::string("ncclAllGather: ") + (es ? es(rc) : "error");
    return LSL_ERR_NCCL;
  }
  LSL_CUDA(cudaMemcpyAsync(all_recs, d + lb, lb * (size_t)nranks, cudaMemcpyDeviceToHost, ctx->stream));
  LSL_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stats.d2h_bytes += lb * nranks;
  return LSL_OK;
}

typedef int (*nccl_bcast_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
extern "C" int lsl_bcast_frame(lsl_ctx* ctx, int root, lsl_frame* frame, lsl_frame** out) {
  if (!ctx || !out) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  if (!ctx->nccl_comm) { ctx->err = "no communicator: call lsl_comm_init first"; return LSL_ERR_NCCL; }
  if (root < 0 || root >= ctx->nccl_nranks) return LSL_ERR_ARG;
  const bool is_root = ctx->nccl_rank == root;
  if (is_root && !frame) return LSL_ERR_ARG;
  void* h = nccl_open(ctx);
  nccl_bcast_fn bc = h ? (nccl_bcast_fn)dlsym(h, "ncclBroadcast") : nullptr;
  if (!bc) return LSL_ERR_NCCL;
  cudaStream_t st = ctx->stream;
  int32_t* d_n = nullptr;
  LSL_CUDA(cudaMallocAsync((void**)&d_n, sizeof(int32_t), st));
  int32_t n = is_root ? frame->nlines : 0;
  if (is_root) LSL_CUDA(cudaMemcpyAsync(d_n, &n, sizeof(n), cudaMemcpyHostToDevice, st));
  int rc = bc(d_n, d_n, sizeof(int32_t), /* ncclChar */ 0, root, ctx->nccl_comm, st);
  if (rc == 0) {
    LSL_CUDA(cudaMemcpyAsync(&n, d_n, sizeof(n), cudaMemcpyDeviceToHost, st));
    LSL_CUDA(cudaStreamSynchronize(st));
  }
  cudaFreeAsync(d_n, st);
  if (rc != 0 || n < 0 || n > LSL_MAX_LINES) { ctx->err = "ncclBroadcast (line count) failed"; return LSL_ERR_NCCL; }
  lsl_frame* fr = frame;
  if (!is_root) {
    fr = new (std::nothrow) lsl_frame();
    if (!fr) return LSL_ERR_ARG;
    fr->ctx = ctx; fr->nlines = n; fr->nsegs = 0; fr->d_lines = nullptr; fr->blk = nullptr; fr->have_dbg = false; fr->have_host = false;
    if (n) LSL_CUDA(cudaMalloc((void**)&fr->d_lines, sizeof(lsl_line_rec) * (size_t)n));
  }
  if (n) {
    rc = bc(fr->d_lines, fr->d_lines, sizeof(lsl_line_rec) * (size_t)n, 0, root, ctx->nccl_comm, st);
    if (rc != 0) { ctx->err = "ncclBroadcast (line records) failed"; if (!is_root) lsl_frame_free(fr); return LSL_ERR_NCCL; }
    LSL_CUDA(cudaStreamSynchronize(st));
  }
  *out = fr;
  return LSL_OK;
}

// Ring shift of the block tails of a stream split over the ranks: send `frame` to the next rank, receive the previous
// rank's in *out. Two grouped ncclSend / ncclRecv rounds (line count, then the records), device to device.
typedef int (*nccl_send_fn)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_recv_fn)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_group_fn)(void);
extern "C" int lsl_shift_frame(lsl_ctx* ctx, const lsl_frame* frame, lsl_frame** out) {
  if (!ctx || !frame || !out) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  if (!ctx->nccl_comm) { ctx->err = "no communicator: call lsl_comm_init first"; return LSL_ERR_NCCL; }
  void* h = nccl_open(ctx);
  nccl_send_fn snd = h ? (nccl_send_fn)dlsym(h, "ncclSend") : nullptr;
  nccl_recv_fn rcv = h ? (nccl_recv_fn)dlsym(h, "ncclRecv") : nullptr;
  nccl_group_fn gs = h ? (nccl_group_fn)dlsym(h, "ncclGroupStart") : nullptr;
  nccl_group_fn ge = h ? (nccl_group_fn)dlsym(h, "ncclGroupEnd") : nullptr;
  if (!snd || !rcv || !gs || !ge) return LSL_ERR_NCCL;
  const int nr = ctx->nccl_nranks, next = (ctx->nccl_rank + 1) % nr, prev = (ctx->nccl_rank + nr - 1) % nr;
  cudaStream_t st = ctx->stream;
  int32_t* d_n = nullptr;
  LSL_CUDA(cudaMallocAsync((void**)&d_n, 2 * sizeof(int32_t), st));
  int32_t n[2] = {frame->nlines, 0};
  LSL_CUDA(cudaMemcpyAsync(d_n, n, sizeof(n), cudaMemcpyHostToDevice, st));
  int rc = gs();
  rc |= snd(d_n, sizeof(int32_t), /* ncclChar */ 0, next, ctx->nccl_comm, st);
  rc |= rcv(d_n + 1, sizeof(int32_t), 0, prev, ctx->nccl_comm, st);
  rc |= ge();
  if (rc == 0) {
    LSL_CUDA(cudaMemcpyAsync(&n[1], d_n + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    LSL_CUDA(cudaStreamSynchronize(st));
  }
  cudaFreeAsync(d_n, st);
  if (rc != 0 || n[1] < 0 || n[1] > LSL_MAX_LINES) { ctx->err = "ncclSend / ncclRecv (line count) failed"; return LSL_ERR_NCCL; }
  lsl_frame* fr = new (std::nothrow) lsl_frame();
  if (!fr) return LSL_ERR_ARG;
  fr->ctx = ctx; fr->nlines = n[1]; fr->nsegs = 0; fr->d_lines = nullptr; fr->blk = nullptr; fr->have_dbg = false; fr->have_host = false;
  if (n[1]) LSL_CUDA(cudaMalloc((void**)&fr->d_lines, sizeof(lsl_line_rec) * (size_t)n[1]));
  rc = gs();
  if (n[0]) rc |= snd(frame->d_lines, sizeof(lsl_line_rec) * (size_t)n[0], 0, next, ctx->nccl_comm, st);
  if (n[1]) rc |= rcv(fr->d_lines, sizeof(lsl_line_rec) * (size_t)n[1], 0, prev, ctx->nccl_comm, st);
  rc |= ge();
  if (rc != 0) { ctx->err = "ncclSend / ncclRecv (line records) failed"; lsl_frame_free(fr); return LSL_ERR_NCCL; }
  LSL_CUDA(cudaStreamSynchronize(st));
  *out = fr;
  return LSL_OK;
}

// ------------------------------------------------------------------ introspection ----
extern "C" int lsl_get_stats(const lsl_ctx* ctx, lsl_stats* out) {
  if (!ctx || !out) return LSL_ERR_ARG;
  LSL_LOCK(ctx);
  *out = ctx->stats;
  return LSL_OK;
}
extern "C" int lsl_last_timing(const lsl_ctx* ctx, float* ms_total, float* ms_rg) {
  if (!ctx) return LSL_ERR_ARG;
  LSL_LOCK(ctx);
  if (ms_total) *ms_total = ctx->ms_total;
  if (ms_rg) *ms_rg = ctx->ms_rg;
  return LSL_OK;
}
extern "C" int lsl_kernel_times(const lsl_ctx* ctx, float* ms, int cap, int* n) {
  if (!ctx || !n) return LSL_ERR_ARG;
  LSL_LOCK(ctx);
  *n = LSL_K_COUNT;
  if (cap < LSL_K_COUNT || !ms) return LSL_ERR_CAPACITY;
  for (int k = 0; k < LSL_K_COUNT; ++k) ms[k] = ctx->kran[k] ? ctx->kms[k] : 0.f;
  return LSL_OK;
}

extern "C" int64_t lsl_debug_read(lsl_ctx* ctx, int what, void* dst, int64_t cap_bytes) {
  if (!ctx || !dst) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  const LslDims& d = ctx->dims;
  const LslWork& w = ctx->wk;
  const void* src = nullptr;
  int64_t count = 0, esz = 1;
  int32_t ns = 0;
  switch (what) {
    case 0: src = w.gray; count = (int64_t)d.W * d.H; esz = 1; break;
    case 1: src = w.scaled; count = (int64_t)d.sw * d.sh; esz = 8; break;
    case 2: src = w.angles; count = (int64_t)d.sw * d.sh; esz = 8; break;
    case 3: src = w.modgrad; count = (int64_t)d.sw * d.sh; esz = 8; break;
    case 4:
      if (cudaMemcpy(&ns, w.nseeds, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return LSL_ERR_CUDA;
      src = w.seeds; count = ns; esz = 4; break;
    case 5: src = w.gx; count = (int64_t)d.W * d.H; esz = 2; break;
    case 6: src = w.gy; count = (int64_t)d.W * d.H; esz = 2; break;
    default: return LSL_ERR_ARG;
  }
  if (count * esz > cap_bytes) return LSL_ERR_CAPACITY;
  if (count && cudaMemcpy(dst, src, count * esz, cudaMemcpyDeviceToHost) != cudaSuccess) return LSL_ERR_CUDA;
  return count;
}
