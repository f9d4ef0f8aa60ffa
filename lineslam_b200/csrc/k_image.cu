// k_image.cu — dense image passes of the front end for a batch of frames (sm_100a).
//   K1a gray_kernel        cvtColor CV_RGB2GRAY, OpenCV 2.4 fixed point (src/node.cpp:191-196)
//   K1b xpass/ypass        LSD gaussian_sampler (external/lsd/lsd.cpp:529-646)
//   K2  ll_angle_kernel    LSD ll_angle gradient / level-line angle / bins (lsd.cpp:670-794)
//   K2b seed_list_kernel   ordered bin concatenation == list_p (lsd.cpp:756-787)
//   K4  sobel5_kernel      cv::Sobel ksize 5, both derivatives (src/line/lineslam.cpp:313-314)
// All of these are HBM/L2-streaming stencils; arithmetic follows the reference's operation order
// exactly (compiled with --fmad=false) so every plane is bit-identical to the CPU path.
#include "lsl_internal.h"
#include <cuda/barrier>
#include "shared/lsl_math.h"

using namespace lslm;

// ---------------------------------------------------------------- taps ----
// gaussian_kernel (lsd.cpp:466-489) evaluated once per output coordinate, as the reference does.
// Tap i of output x is stored at k[x * sx + i * si]: (8, 1) = one row of taps per output (y table, read uniformly by a
// warp), (1, nout) = tap-major (x table: a warp reads consecutive doubles).
__global__ void taps_kernel(double* __restrict__ k, int* __restrict__ c, int nout, double scale, double sigma, int h, int sx, int si) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nout) return;
  double xx = (double)x / scale;
  int xc = (int)floor(xx + 0.5);
  double mean = (double)h + xx - (double)xc;
  int n = 1 + 2 * h;
  double v[8], sum = 0.0;
  for (int i = 0; i < n; ++i) {
    double val = ((double)i - mean) / sigma;
    v[i] = lsl_exp(-0.5 * val * val);
    sum += v[i];
  }
  if (sum >= 0.0)
    for (int i = 0; i < n; ++i) v[i] /= sum;
  for (int i = 0; i < 8; ++i) k[(size_t)x * sx + (size_t)i * si] = i < n ? v[i] : 0.0;
  c[x] = xc;
}

int lsl_prepare_taps(lsl_ctx* ctx) {
  const lsl_params& P = ctx->P;
  LslDims& d = ctx->dims;
  double sigma = P.lsd_scale < 1.0 ? P.lsd_sigma_scale / P.lsd_scale : P.lsd_sigma_scale;
  // h = ceil(sigma * sqrt(2 * 3 * ln 10)); ln 10 to double precision is enough for the ceil
  int h = (int)ceil(sigma * sqrt(2.0 * 3.0 * 2.302585092994046));
  if (1 + 2 * h > 8) { ctx->err = "gaussian kernel wider than 8 taps"; return LSL_ERR_ARG; }
  ctx->taps.h = h; ctx->taps.n = 1 + 2 * h;
  taps_kernel<<<(d.sw + 127) / 128, 128, 0, ctx->stream>>>(ctx->taps.kx, ctx->taps.xc, d.sw, P.lsd_scale, sigma, h, 1, d.sw);
  taps_kernel<<<(d.sh + 127) / 128, 128, 0, ctx->stream>>>(ctx->taps.ky, ctx->taps.yc, d.sh, P.lsd_scale, sigma, h, 8, 1);
  ctx->stats.kernel_launches += 2;
  LSL_CUDA(cudaGetLastError());
  return LSL_OK;
}

// ---------------------------------------------------------------- gray ----
// 4 pixels per thread: 12 interleaved bytes in (3 x 32-bit loads), one 32-bit store.
__global__ void gray_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ gray, long npix4) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix4) return;
  const uint32_t* p = reinterpret_cast<const uint32_t*>(img) + i * 3;
  uint32_t a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  uint32_t by[12] = {a & 255, (a >> 8) & 255, (a >> 16) & 255, a >> 24, b & 255, (b >> 8) & 255,
                     (b >> 16) & 255, b >> 24, c & 255, (c >> 8) & 255, (c >> 16) & 255, c >> 24};
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t g = (by[3 * k] * 4899u + by[3 * k + 1] * 9617u + by[3 * k + 2] * 1868u + 8192u) >> 14;
    out |= (g & 255u) << (8 * k);
  }
  reinterpret_cast<uint32_t*>(gray)[i] = out;
}

// ------------------------------------------------------------- sampler ----
__device__ __forceinline__ int sym_index(int j, int n) {  // lsd.cpp:596-599 symmetric boundary
  int n2 = 2 * n;
  while (j < 0) j += n2;
  while (j >= n2) j -= n2;
  if (j >= n) j = n2 - 1 - j;
  return j;
}
// x pass: 4 outputs of one row per thread (128 apart, so that a warp writes consecutive doubles), the tap
// loop unrolled for the 7-tap kernel of sigma 0.75 (NT = 0: generic tap count). Sum order i = 0..ntap-1 (lsd.cpp:600-615).
template <int NT>
__global__ void __launch_bounds__(128) xpass_kernel(const uint8_t* __restrict__ gray, double* __restrict__ aux, const double* __restrict__ kx,
                                                    const int* __restrict__ xc, int W, int H, int sw, int h, int ntap_rt) {
  const int ntap = NT ? NT : ntap_rt;
  const int x0 = blockIdx.x * 512 + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
  const uint8_t* row = gray + ((size_t)f * H + y) * W;
  double* out = aux + ((size_t)f * H + y) * sw;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int x = x0 + q * 128;           // coalesced: lanes write consecutive outputs
    if (x >= sw) break;
    const int c = xc[x] - h;
    double sum = 0.0;
    if (c >= 0 && c + ntap <= W) {          // interior: no boundary folding
#pragma unroll
      for (int i = 0; i < ntap; ++i) sum += (double)row[c + i] * kx[i * sw + x];
    } else {
#pragma unroll
      for (int i = 0; i < ntap; ++i) sum += (double)row[sym_index(c + i, W)] * kx[i * sw + x];
    }
    out[x] = sum;
  }
}
__global__ void ypass_kernel(const double* __restrict__ aux, double* __restrict__ out, const double* __restrict__ ky,
                             const int* __restrict__ yc, int H, int sw, int sh, int h, int ntap) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
  if (x >= sw) return;
  const double* a = aux + (size_t)f * H * sw;
  int c = yc[y];
  double sum = 0.0;
  for (int i = 0; i < ntap; ++i) {
    int j = sym_index(c - h + i, H);
    sum += a[(size_t)j * sw + x] * ky[y * 8 + i];
  }
  out[((size_t)f * sh + y) * sw + x] = sum;
}

// y pass with TMA staging (sm_100a): one CTA produces YP_R consecutive output rows x 128 columns. The aux rows
// those outputs tap (rows yc[y0]-h .. yc[y0+YP_R-1]+h, folded by the symmetric boundary) are fetched ONCE into
// shared memory by bulk asynchronous copies (cp.async.bulk.shared::cluster.global, one 1 KB row segment each,
// completion counted on an mbarrier) instead of seven global loads per output; the sum keeps the reference's tap
// order i = 0..6 (lsd.cpp:625-640). Needs sw even (16-byte segments); otherwise ypass_kernel is used.
#define YP_R 8
#define YP_ROWS 24
__global__ void __launch_bounds__(512) ypass_tma_kernel(const double* __restrict__ aux, double* __restrict__ out,
                                                        const double* __restrict__ ky, const int* __restrict__ yc, int H, int sw,
                                                        int sh, int h, int ntap) {
  __shared__ __align__(128) double tile[YP_ROWS][128];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ cuda::barrier<cuda::thread_scope_block> bar;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 128 + tx;
  const int x0 = blockIdx.x * 128, y0 = blockIdx.y * YP_R, f = blockIdx.z;
  const int y1 = min(y0 + YP_R, sh) - 1;
  const double* a = aux + (size_t)f * H * sw;
  const int jlo = yc[y0] - h, jhi = yc[y1] + h;          // unfolded row range of this CTA
  const int nrows = jhi - jlo + 1;                        // <= YP_ROWS (checked on the host)
  const int ncol = min(128, sw - x0);
  if (tid == 0) {
    init(&bar, 512);
    cuda::device::experimental::fence_proxy_async_shared_cta();
  }
  __syncthreads();
  cuda::barrier<cuda::thread_scope_block>::arrival_token token;
  if (tid == 0) {
    const size_t bytes = (size_t)ncol * sizeof(double);
    for (int r = 0; r < nrows; ++r)
      cuda::device::memcpy_async_tx(&tile[r][0], a + (size_t)sym_index(jlo + r, H) * sw + x0, cuda::aligned_size_t<16>(bytes), bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, bytes * nrows);
  } else token = bar.arrive();
  bar.wait(std::move(token));
  if (tx < ncol) {
    for (int yy = ty; yy < YP_R; yy += 4) {
      const int y = y0 + yy;
      if (y >= sh) break;
      const int c = yc[y];
      double sum = 0.0;
      for (int i = 0; i < ntap; ++i) sum += tile[c - h + i - jlo][tx] * ky[y * 8 + i];
      out[((size_t)f * sh + y) * sw + x0 + tx] = sum;
    }
  }
}

// ------------------------------------------------------------ ll_angle ----
// ll_angle (lsd.cpp:670-794) on 32 x 32 tiles. The gradient pass is per pixel; the pixels above the gradient
// threshold (a few percent of a frame) are compacted inside the tile so that the correctly rounded atan2 /
// sincos run on dense warps; the bin plane is written column-major (the reference's x-outer / y-inner seed
// order) through a shared-memory transpose, i.e. with coalesced stores.
#ifndef LLA_TH
#define LLA_TH 8      // tile rows per CTA (tile = 32 x LLA_TH pixels). After the gradient pass only the few warps that hold compacted
                      // above-threshold pixels keep working (double-double atan2 / sincos), but a CTA keeps its warp slots until
                      // they finish: with 32 x 32 tiles (2 CTAs per SM) 18.5 % of the warp slots were active; 32 x 8 tiles put
                      // eight independent tiles on an SM: 4.66 -> < 3 ms per 592 frames
#endif
__global__ void __launch_bounds__(32 * LLA_TH) ll_angle_kernel(const double* __restrict__ in, double* __restrict__ angles,
                                                        double* __restrict__ modgrad, double2* __restrict__ cs,
                                                        uint16_t* __restrict__ binT, int p, int n, double threshold,
                                                        int n_bins, double max_grad) {
  __shared__ uint16_t s_bin[LLA_TH][33];
  __shared__ uint16_t s_list[32 * LLA_TH];
  __shared__ double s_gx[32 * LLA_TH], s_gy[32 * LLA_TH];
  __shared__ int s_cnt;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const int x = blockIdx.x * 32 + tx, y = blockIdx.y * LLA_TH + ty, f = blockIdx.z;
  const size_t base = (size_t)f * p * n;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  uint16_t bin = 0xFFFF;
  if (x < p && y < n) {
    const size_t adr = (size_t)y * p + x;
    double norm = 0.0;
    bool valid = false;
    if (x < p - 1 && y < n - 1) {
      const double* I = in + base;
      double com1 = I[adr + p + 1] - I[adr];
      double com2 = I[adr + 1] - I[adr + p];
      double gx = com1 + com2, gy = com1 - com2;
      double norm2 = gx * gx + gy * gy;
      norm = sqrt(norm2 / 4.0);
      if (!(norm <= threshold)) {
        valid = true;
        unsigned i = (unsigned)(norm * (double)n_bins / max_grad);
        if (i >= (unsigned)n_bins) i = n_bins - 1;
        bin = (uint16_t)i;
        int k = atomicAdd(&s_cnt, 1);
        s_list[k] = (uint16_t)tid; s_gx[k] = gx; s_gy[k] = gy;
      }
    }
    modgrad[base + adr] = norm;
    if (!valid) { angles[base + adr] = LSL_NOTDEF; cs[base + adr] = make_double2(2.0, 0.0); }
  }
  s_bin[ty][tx] = bin;
  __syncthreads();
  const int cnt = s_cnt;
  for (int k = tid; k < cnt; k += 32 * LLA_TH) {
    const int t = s_list[k];
    const size_t adr = (size_t)(blockIdx.y * LLA_TH + (t >> 5)) * p + (blockIdx.x * 32 + (t & 31));
    double ang = lsl_atan2(s_gx[k], -s_gy[k]);
    double sn, c;
    lsl_sincos(ang, &sn, &c);
    angles[base + adr] = ang;
    cs[base + adr] = make_double2(c, sn);
  }
  // transposed store: thread tid writes pixel (x0 + tid / TH, y0 + tid % TH) -> consecutive y for one x
  const int lx = tid / LLA_TH, ly = tid % LLA_TH;
  const int xo = blockIdx.x * 32 + lx, yo = blockIdx.y * LLA_TH + ly;
  if (xo < p && yo < n) binT[base + (size_t)xo * n + yo] = s_bin[ly][lx];
}

// ---------------------------------------------------------- seed list ----
// One CTA (8 warps) per frame reproduces list_p of ll_angle (lsd.cpp:723-787): pixels in the reference's
// x-outer / y-inner order, bucketed by gradient bin, bins 1023 .. 1 concatenated (bin 0 would only be appended if
// it were the highest non-empty bin, which needs norm < max_grad/n_bins <= threshold and therefore never holds a
// seed). The pixel range is cut into 8 contiguous segments, one per warp: pass A counts the bins of every segment
// (shared-memory atomics, order-free), pass B turns the counts into start cursors (bin-major, then segment order),
// pass C lets every warp scatter its own segment stably (ballot + match_any keep the in-segment order).
#define SEED_WARPS 8
__global__ void __launch_bounds__(32 * SEED_WARPS) seed_list_kernel(const uint16_t* __restrict__ binT, int32_t* __restrict__ seeds,
                                                                   int32_t* __restrict__ nseeds, int p, int n, int n_bins) {
  extern __shared__ int s_cnt[];  // [SEED_WARPS][n_bins] counts -> cursors, then [n_bins] bin starts
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint16_t* B = binT + (size_t)f * p * n;
  int32_t* S = seeds + (size_t)f * p * n;
  const int total = p * n;
  const int seg = ((total + SEED_WARPS - 1) / SEED_WARPS + 31) & ~31;   // segment length, multiple of 32
  const int lo = min(warp * seg, total), hi = min(lo + seg, total);
  int* cnt = s_cnt + warp * n_bins;
  int* start = s_cnt + SEED_WARPS * n_bins;
  for (int i = tid; i < SEED_WARPS * n_bins; i += 32 * SEED_WARPS) s_cnt[i] = 0;
  __syncthreads();
  for (int i = lo + lane; i < hi; i += 32) {
    uint16_t b = B[i];
    if (b != 0xFFFF && b != 0) atomicAdd(&cnt[b], 1);
  }
  __syncthreads();
  // bin totals
  for (int b = tid; b < n_bins; b += 32 * SEED_WARPS) {
    int t = 0;
    for (int w = 0; w < SEED_WARPS; ++w) t += s_cnt[w * n_bins + b];
    start[b] = t;
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int b = n_bins - 1; b >= 1; --b) { int c = start[b]; start[b] = run; run += c; }
    nseeds[f] = run;
  }
  __syncthreads();
  for (int b = tid; b < n_bins; b += 32 * SEED_WARPS) {
    int run = start[b];
    for (int w = 0; w < SEED_WARPS; ++w) { int c = s_cnt[w * n_bins + b]; s_cnt[w * n_bins + b] = run; run += c; }
  }
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
  for (int i0 = lo; i0 < hi; i0 += 32) {
    int i = i0 + lane;
    uint16_t b = i < hi ? B[i] : (uint16_t)0xFFFF;
    bool valid = (b != 0xFFFF && b != 0);
    unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      unsigned grp = __match_any_sync(vm, (unsigned)b);
      int pos = cnt[b] + __popc(grp & lt);
      int x = i / n, y = i - x * n;
      S[pos] = x | (y << 16);
      __syncwarp(vm);
      if ((grp & lt) == 0) cnt[b] += __popc(grp);
    }
    __syncwarp();
  }
}

// --------------------------------------------------------------- Sobel ----
__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}
__global__ void sobel5_kernel(const uint8_t* __restrict__ gray, int16_t* __restrict__ gx, int16_t* __restrict__ gy, int W, int H) {
  __shared__ uint8_t t[20][36 + 4];
  int f = blockIdx.z;
  const uint8_t* G = gray + (size_t)f * W * H;
  int x0 = blockIdx.x * 32, y0 = blockIdx.y * 16;
  for (int i = threadIdx.y * 32 + threadIdx.x; i < 20 * 36; i += 32 * 16) {
    int ty = i / 36, tx = i - ty * 36;
    int yy = reflect101(y0 + ty - 2, H), xx = reflect101(x0 + tx - 2, W);
    t[ty][tx] = G[(size_t)yy * W + xx];
  }
  __syncthreads();
  int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= W || y >= H) return;
  const int dv[5] = {-1, -2, 0, 2, 1}, sm[5] = {1, 4, 6, 4, 1};
  int sx = 0, sy = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j)
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      int v = t[threadIdx.y + j][threadIdx.x + i];
      sx += sm[j] * dv[i] * v;
      sy += dv[j] * sm[i] * v;
    }
  gx[(size_t)f * W * H + (size_t)y * W + x] = (int16_t)sx;
  gy[(size_t)f * W * H + (size_t)y * W + x] = (int16_t)sy;
}

// cv::Sobel ksize 5 with TMA tile staging: one CTA = 128 x 32 output pixels; the (16 + 128 + 16) x 36 byte halo tile of
// the gray plane (the innermost box coordinate of a TMA tile must be 16-byte aligned: measured with tools/ubench/tma_probe.cu,
// a start at x0 - 2 raises an illegal-instruction fault) arrives in shared memory by ONE cp.async.bulk.tensor.3d (tensor map over [batch][H][W] u8, out-of-image
// bytes zero-filled by the copy engine, completion on an mbarrier; UTMALDG in SASS). BORDER_REFLECT_101 only concerns CTAs
// on the image border: the mirrored pixel always lies inside the same tile, so those CTAs remap the tile index instead
// of touching global memory. Each thread produces a 4 x 4 micro-tile from eight aligned 8-byte row segments (separable:
// horizontal [-1 -2 0 2 1] / [1 4 6 4 1] per row, then the vertical taps) and stores 8 bytes per row and plane: exact
// integer arithmetic, so the planes equal sobel5_kernel's bit for bit.
#define SB_W 128
#define SB_H 32
#define SB_TW 160   // tile row pitch in bytes, columns x0 - 16 .. x0 + 143
#define SB_X0 16    // tile column of image column x0
#define SB_TH 36
__global__ void __launch_bounds__(256) sobel5_tma_kernel(const __grid_constant__ CUtensorMap tmap, int16_t* __restrict__ gx,
                                                        int16_t* __restrict__ gy, int W, int H, int f0) {
  __shared__ __align__(128) uint8_t t[SB_TH][SB_TW];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * SB_W, y0 = blockIdx.y * SB_H, f = blockIdx.z;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), t_a = (uint32_t)__cvta_generic_to_shared(&t[0][0]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_a), "r"(SB_TH * SB_TW));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                 ::"r"(t_a), "l"(&tmap), "r"(x0 - SB_X0), "r"(y0 - 2), "r"(f0 + f), "r"(bar_a) : "memory");
  }
  __syncthreads();   // the barrier is initialised for everybody
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar_a) : "memory");
  }
  const int cx = (tid & 31) * 4, cy = (tid >> 5) * 4;      // micro-tile origin inside the CTA tile
  if (x0 + cx >= W || y0 + cy >= H) return;                // micro-tile entirely outside the image (ragged sizes)
  const bool border = x0 == 0 || y0 == 0 || x0 + SB_W + 2 > W || y0 + SB_H + 2 > H;
  int hd[8][4], hs[8][4];                                  // horizontal derivative / smoothing sums of the 8 rows
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int v[8];
    if (!border) {
      // pixels x0 + cx - 2 .. x0 + cx + 5 = tile columns cx + 14 .. cx + 21: three aligned words
      const uint32_t w0 = *reinterpret_cast<const uint32_t*>(&t[cy + r][cx + 12]), w1 = *reinterpret_cast<const uint32_t*>(&t[cy + r][cx + 16]),
                     w2 = *reinterpret_cast<const uint32_t*>(&t[cy + r][cx + 20]);
      v[0] = (w0 >> 16) & 0xff; v[1] = w0 >> 24;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[2 + k] = (w1 >> (8 * k)) & 0xff;
      v[6] = w2 & 0xff; v[7] = (w2 >> 8) & 0xff;
    } else {
      // mirrored coordinates of pixels an in-image output needs lie inside the tile; the clamps only touch inputs of
      // outputs beyond the image edge, which are never stored
      const int ty = min(max(reflect101(y0 + cy + r - 2, H) - (y0 - 2), 0), SB_TH - 1);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = t[ty][min(max(reflect101(x0 + cx + k - 2, W) - (x0 - SB_X0), 0), SB_TW - 1)];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      hd[r][c] = -v[c] - 2 * v[c + 1] + 2 * v[c + 3] + v[c + 4];
      hs[r][c] = v[c] + 4 * v[c + 1] + 6 * v[c + 2] + 4 * v[c + 3] + v[c + 4];
    }
  }
  const int x = x0 + cx;
#pragma unroll
  for (int yy = 0; yy < 4; ++yy) {
    const int y = y0 + cy + yy;
    if (y >= H) break;
    short ox[4], oy[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      ox[c] = (short)(hd[yy][c] + 4 * hd[yy + 1][c] + 6 * hd[yy + 2][c] + 4 * hd[yy + 3][c] + hd[yy + 4][c]);
      oy[c] = (short)(-hs[yy][c] - 2 * hs[yy + 1][c] + 2 * hs[yy + 3][c] + hs[yy + 4][c]);
    }
    const size_t o = (size_t)f * W * H + (size_t)y * W + x;
    *reinterpret_cast<uint2*>(gx + o) = make_uint2((uint16_t)ox[0] | ((uint32_t)(uint16_t)ox[1] << 16), (uint16_t)ox[2] | ((uint32_t)(uint16_t)ox[3] << 16));
    *reinterpret_cast<uint2*>(gy + o) = make_uint2((uint16_t)oy[0] | ((uint32_t)(uint16_t)oy[1] << 16), (uint16_t)oy[2] | ((uint32_t)(uint16_t)oy[3] << 16));
  }
}

// Tensor map of the context's gray planes: rank 3 {W, H, max_batch} u8, box {160, 36, 1}; (re)encoded when the frame size
// changes. cuTensorMapEncodeTiled is taken from the driver through the runtime (no link-time dependency on libcuda).
int lsl_prepare_tmaps(lsl_ctx* ctx) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  ctx->tmap_gray_ok = false;
  const LslDims& d = ctx->dims;
  if ((d.W & 15) || d.W < SB_TW || d.H < SB_TH) return LSL_OK;       // TMA stride rule (multiples of 16 bytes): such sizes keep the plain kernel
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return LSL_OK;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)ctx->max_batch};
  const cuuint64_t strides[2] = {(cuuint64_t)d.W, (cuuint64_t)d.W * d.H};
  const cuuint32_t box[3] = {SB_TW, SB_TH, 1}, estr[3] = {1, 1, 1};
  CUresult r = ((encode_fn)fn)(&ctx->tmap_gray, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ctx->wk.gray, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ctx->tmap_gray_ok = (r == CUDA_SUCCESS);
  return LSL_OK;
}

// Image stage for the frames [f0, f0 + n) of the batch (the host-buffer path calls it per upload chunk so that
// the RGB upload of the next chunk overlaps these kernels). channels == 1: the caller already has a gray plane.
int lsl_launch_image(lsl_ctx* ctx, int f0, int n, const uint8_t* d_img, int channels) {
  const LslDims& d = ctx->dims;
  const lsl_params& P = ctx->P;
  const LslWork& w = ctx->wk;
  cudaStream_t st = ctx->stream;
  const size_t npix = (size_t)d.W * d.H, spix = (size_t)d.sw * d.sh, apix = (size_t)d.H * d.sw;
  uint8_t* gray = w.gray + f0 * npix;
  const uint8_t* img = d_img + (size_t)f0 * npix * channels;
  if (channels == 3) {
    long n4 = (long)(npix * n / 4);
    LSL_KSTART(ctx, LSL_K_GRAY);
    gray_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(img, gray, n4);
    LSL_KSTOP(ctx, LSL_K_GRAY);
  } else {
    LSL_CUDA(cudaMemcpyAsync(gray, img, npix * n, cudaMemcpyDeviceToDevice, st));
  }
  dim3 bx(128), gxp((d.sw + 511) / 512, d.H, n), gyp((d.sw + 127) / 128, d.sh, n);
  LSL_KSTART(ctx, LSL_K_XPASS);
  if (ctx->taps.n == 7) xpass_kernel<7><<<gxp, bx, 0, st>>>(gray, w.aux + f0 * apix, ctx->taps.kx, ctx->taps.xc, d.W, d.H, d.sw, ctx->taps.h, 7);
  else xpass_kernel<0><<<gxp, bx, 0, st>>>(gray, w.aux + f0 * apix, ctx->taps.kx, ctx->taps.xc, d.W, d.H, d.sw, ctx->taps.h, ctx->taps.n);
  LSL_KSTOP(ctx, LSL_K_XPASS);
  LSL_KSTART(ctx, LSL_K_YPASS);
  // rows tapped by YP_R consecutive outputs: ceil(YP_R / scale) + 2 h + 2 must fit the staged tile
  if ((d.sw & 1) == 0 && (int)ceil(YP_R / P.lsd_scale) + 2 * ctx->taps.h + 2 <= YP_ROWS) {
    dim3 byt(128, 4), gyt((d.sw + 127) / 128, (d.sh + YP_R - 1) / YP_R, n);
    ypass_tma_kernel<<<gyt, byt, 0, st>>>(w.aux + f0 * apix, w.scaled + f0 * spix, ctx->taps.ky, ctx->taps.yc, d.H, d.sw, d.sh, ctx->taps.h, ctx->taps.n);
  } else
    ypass_kernel<<<gyp, bx, 0, st>>>(w.aux + f0 * apix, w.scaled + f0 * spix, ctx->taps.ky, ctx->taps.yc, d.H, d.sw, d.sh, ctx->taps.h, ctx->taps.n);
  LSL_KSTOP(ctx, LSL_K_YPASS);
  double prec = LSL_PI * P.lsd_ang_th / 180.0;
  double rho = P.lsd_quant / lsl_sin(prec);
  LSL_KSTART(ctx, LSL_K_LLANGLE);
  dim3 gla((d.sw + 31) / 32, (d.sh + LLA_TH - 1) / LLA_TH, n), bla(32, LLA_TH);
  ll_angle_kernel<<<gla, bla, 0, st>>>(w.scaled + f0 * spix, w.angles + f0 * spix, w.modgrad + f0 * spix, w.cs + f0 * spix,
                                      w.binT + f0 * spix, d.sw, d.sh, rho, P.lsd_n_bins, P.lsd_max_grad);
  LSL_KSTOP(ctx, LSL_K_LLANGLE);
  LSL_KSTART(ctx, LSL_K_SOBEL);
  if (ctx->tmap_gray_ok) {
    dim3 gt((d.W + SB_W - 1) / SB_W, (d.H + SB_H - 1) / SB_H, n);
    sobel5_tma_kernel<<<gt, 256, 0, st>>>(ctx->tmap_gray, w.gx + f0 * npix, w.gy + f0 * npix, d.W, d.H, f0);
  } else {
    dim3 bs(32, 16), gs((d.W + 31) / 32, (d.H + 15) / 16, n);
    sobel5_kernel<<<gs, bs, 0, st>>>(gray, w.gx + f0 * npix, w.gy + f0 * npix, d.W, d.H);
  }
  LSL_KSTOP(ctx, LSL_K_SOBEL);
  LSL_CUDA(cudaGetLastError());
  return LSL_OK;
}

// Seed lists of all n frames (one CTA per frame, latency-bound: launched once per batch, not per upload chunk)
int lsl_launch_seeds(lsl_ctx* ctx, int n) {
  const LslDims& d = ctx->dims;
  const LslWork& w = ctx->wk;
  const size_t seed_smem = (SEED_WARPS + 1) * ctx->P.lsd_n_bins * sizeof(int);
  if (seed_smem > 48 * 1024) cudaFuncSetAttribute(seed_list_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_smem);
  LSL_KSTART(ctx, LSL_K_SEEDS);
  seed_list_kernel<<<n, 32 * SEED_WARPS, seed_smem, ctx->stream>>>(w.binT, w.seeds, w.nseeds, d.sw, d.sh, ctx->P.lsd_n_bins);
  LSL_KSTOP(ctx, LSL_K_SEEDS);
  LSL_CUDA(cudaGetLastError());
  return LSL_OK;
}


// ---- raw 16-bit depth -> metres (src/openni_listener.cpp:1233-1244): convertTo(CV_32FC1), values < 1e-5 -> quiet NaN with
// the bit pattern the x86 reference leaves (0x7fc00000), then the MatExpr "/ factor" = float multiply by (float)(1 / factor).
// HBM streaming: 8 pixels (one 16-byte load, two 16-byte stores) per thread.
__global__ void __launch_bounds__(256) depth_u16_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, size_t count, float scale) {
  const size_t i8 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i8 + 8 <= count) {
    const uint4 v = *reinterpret_cast<const uint4*>(in + i8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = (float)(w[k] & 0xffffu), b = (float)(w[k] >> 16);
      o[2 * k] = ((double)a < 1e-5) ? __int_as_float(0x7fc00000) : __fmul_rn(a, scale);
      o[2 * k + 1] = ((double)b < 1e-5) ? __int_as_float(0x7fc00000) : __fmul_rn(b, scale);
    }
    *reinterpret_cast<float4*>(out + i8) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(out + i8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
  } else {
    for (size_t i = i8; i < count; ++i) {
      const float a = (float)in[i];
      out[i] = ((double)a < 1e-5) ? __int_as_float(0x7fc00000) : __fmul_rn(a, scale);
    }
  }
}

int lsl_launch_depth_u16(lsl_ctx* ctx, cudaStream_t st, const uint16_t* d_in, float* d_out, size_t count, float scale) {
  const size_t threads = (count + 7) / 8;
  depth_u16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_in, d_out, count, scale);
  ctx->stats.kernel_launches += 1;
  LSL_CUDA(cudaGetLastError());
  return LSL_OK;
}
