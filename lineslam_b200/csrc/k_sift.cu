// k_sift.cu — the point-feature half of Node::Node on the device (src/node.cpp:219-310, 952-1018):
//   detector->detect (SIFT)  ->  removeDepthless  ->  KeyPointsFilter::retainBest(max_keypoints)  ->  extractor->compute
//   ->  projectTo3D  ->  squareroot_descriptor_space
// for every frame of an extract call, on the gray / depth planes that are already in HBM, ending in the frame's point
// features (what lsl_frame_set_points would upload) without a host round trip.
//
// SIFT follows OpenCV's implementation (modules/features2d/src/sift.dispatch.cpp / sift.simd.hpp; restated and pinned
// against cv2 in oracle/oracle_sift.py): float images, first octave -1 (2 x bilinear up-sampling + blur 1.249), three
// layers per octave (six Gaussian images, incremental blurs, BORDER_REFLECT_101), DoG extrema with |v| > 1, quadratic
// refinement (cv::solve's 3 x 3 closed form in double), contrast 0.04 / edge 10 tests, 36-bin orientation histogram with
// cv::fastAtan2's polynomial, 4 x 4 x 8 descriptor with trilinear interpolation, 0.2 clamp, x 512 saturation.
// Tier-T against cv2 (the filters and histograms are summed in another order; histograms here are accumulated in 64-bit
// fixed point so that the result does not depend on the order of the atomics: run-to-run deterministic).
//
// Kernels (all HBM / L2 streaming or gather work; nothing here is reshaped into a GEMM):
//   sift_upsample_kernel   u8 gray -> 2W x 2H float base             sift_blur_row/col_kernel  separable Gaussian
//   sift_down_kernel       next octave = every second pixel          sift_extrema_kernel       DoG on the fly, 26-neighbour test,
//   sift_orient_kernel     warp per candidate: histogram, peaks          refinement -> candidate list
//   sift_flag / rank_kernel duplicates, depth, rank by response      sift_describe_kernel      CTA per kept keypoint
//   sift_pack_kernel       xyz1 + descriptor rows of the batch block
#include "lsl_internal.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <float.h>
#include <limits.h>
#include <mutex>
#include <new>
#include <vector>

#define SF_LAYERS 3
#define SF_NIMG 6
#define SF_MAX_OCT 12
#define SF_BORDER 5
#define SF_MAX_CAND 16384
#define SF_MAX_KP 16384
#ifndef SF_CHUNK
#define SF_CHUNK 148          // frames per pyramid pass at most (39 MB of pyramid per VGA frame: 5.8 GB). The stage is a chain of
#endif                        // ~110 small launches per pass (9 octaves x 11), so fewer, larger passes win: 148 frames per step
                              // took 29.6 / 22.0 / 17.4 / 15.6 / 13.8 ms with passes of 8 / 16 / 32 / 64 / 148 frames
#define SF_FIX 16777216.0f    // 2^24: fixed-point scale of the histogram accumulators

struct SiftOct { int W, H; size_t off; };                       // off: floats from the frame's pyramid base to layer 0
struct SiftGeom { int n_oct, W0, H0; size_t frame_floats; SiftOct o[SF_MAX_OCT]; };
struct SiftCand { int o, layer, r, c; float xi, xr, xc, contr; };
struct SiftKp { float x, y, size, angle, response; int o, layer, slot; };   // x, y, size in INPUT image pixels

__constant__ float c_taps[SF_NIMG][40];
__constant__ int c_rad[SF_NIMG];

__device__ __forceinline__ int reflect101g(int i, int n) {
  if (n == 1) return 0;
  const int p = 2 * n - 2;
  i %= p; if (i < 0) i += p;
  return i >= n ? p - i : i;
}

// cv::resize(INTER_LINEAR) by 2: dst x samples src at x / 2 - 0.25 (weights 0.75 / 0.25, clamped): x first, then y
__global__ void sift_upsample_kernel(const uint8_t* __restrict__ gray, float* __restrict__ pyr, SiftGeom G, int W, int H) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y, f = blockIdx.z;
  if (X >= 2 * W) return;
  const uint8_t* g = gray + (size_t)f * W * H;
  const int x0 = (X + 1) / 2 - 1, y0 = (Y + 1) / 2 - 1;                       // floor(X / 2 - 0.25)
  const float wx = (X & 1) ? 0.25f : 0.75f, wy = (Y & 1) ? 0.25f : 0.75f;
  const int xa = max(x0, 0), xb = min(x0 + 1, W - 1), ya = max(y0, 0), yb = min(y0 + 1, H - 1);
  const float h0 = (float)g[ya * W + xa] * (1.0f - wx) + (float)g[ya * W + xb] * wx;
  const float h1 = (float)g[yb * W + xa] * (1.0f - wx) + (float)g[yb * W + xb] * wx;
  pyr[(size_t)f * G.frame_floats + (size_t)Y * (2 * W) + X] = h0 * (1.0f - wy) + h1 * wy;   // staged in octave 0, layer 1 (scratch)
}

// horizontal / vertical pass of cv::GaussianBlur (float, taps summed in tap order from 0, BORDER_REFLECT_101).
// Row pass: a CTA stages SF_ROW_T + 2 R4 input values of one image row in shared memory (R4 = R rounded up to 4: interior
// tiles of 16-byte aligned rows are staged with float4 loads, border tiles element-wise with the reflection) and every thread
// computes EIGHT consecutive outputs from a register window read as float4 — each input value crosses L2 once instead of
// 2 R + 1 times and the staging / addressing overhead is shared by eight outputs (92 -> ~45 instructions per output at R = 5).
#define SF_ROW_T 512          // outputs per CTA: 64 threads x 8 consecutive pixels
template <int R>
__global__ void __launch_bounds__(SF_ROW_T / 8) sift_blur_row_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t fstride_src,
                                                                     size_t fstride_dst, int W, int H, int kidx) {
  constexpr int R4 = (R + 3) / 4 * 4, NT = SF_ROW_T / 8, SW = SF_ROW_T + 2 * R4;
  __shared__ __align__(16) float s_in[SW];
  const int x0 = blockIdx.x * SF_ROW_T, y = blockIdx.y, f = blockIdx.z, tid = threadIdx.x;
  const float* s = src + (size_t)f * fstride_src + (size_t)y * W;
  const bool vec = (W & 3) == 0 && (((size_t)f * fstride_src) & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (vec && x0 - R4 >= 0 && x0 + SF_ROW_T + R4 <= W) {      // interior tile of an aligned row: s_in[e] = row[x0 - R4 + e]
    const float4* s4 = reinterpret_cast<const float4*>(s + x0 - R4);
    for (int e = tid; e < SW / 4; e += NT) reinterpret_cast<float4*>(s_in)[e] = s4[e];
  } else {
    for (int e = tid; e < SW; e += NT) {
      const int x = x0 - R4 + e;
      s_in[e] = (x >= 0 && x < W) ? s[x] : s[reflect101g(x, W)];
    }
  }
  __syncthreads();
  const int x = x0 + 8 * tid;
  if (x >= W) return;
  float v[8 + 2 * R4];                                       // window of the eight outputs: v[k] = row[x - R4 + k]
#pragma unroll
  for (int k = 0; k < (8 + 2 * R4) / 4; ++k) {
    const float4 q = *reinterpret_cast<const float4*>(s_in + 8 * tid + 4 * k);
    v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
  }
  float o[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t <= 2 * R; ++t) acc += c_taps[kidx][t] * v[k + (R4 - R) + t];
    o[k] = acc;
  }
  float* d = dst + (size_t)f * fstride_dst + (size_t)y * W + x;
  if (vec && (((size_t)f * fstride_dst) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && x + 8 <= W) {
    reinterpret_cast<float4*>(d)[0] = make_float4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<float4*>(d)[1] = make_float4(o[4], o[5], o[6], o[7]);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) if (x + k < W) d[k] = o[k];
  }
}
// Column pass: a thread owns one column of a SF_COL_RUN-row strip and slides a register window down it: RUN + 2 R loads
// (each a coalesced row segment across the warp) for RUN outputs instead of (2 R + 1) RUN; strips that do not touch the image
// border walk a pointer without any bounds logic.
#define SF_COL_RUN 32
template <int R>
__global__ void __launch_bounds__(256) sift_blur_col_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t fstride_src,
                                                            size_t fstride_dst, int W, int H, int kidx) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), f = blockIdx.z;
  const int y0 = (blockIdx.y * 8 + (threadIdx.x >> 5)) * SF_COL_RUN;
  if (x >= W || y0 >= H) return;
  const float* s = src + (size_t)f * fstride_src + x;
  float v[SF_COL_RUN + 2 * R];
  if (y0 - R >= 0 && y0 + SF_COL_RUN + R <= H) {
    const float* p = s + (size_t)(y0 - R) * W;
#pragma unroll
    for (int k = 0; k < SF_COL_RUN + 2 * R; ++k) { v[k] = *p; p += W; }
  } else {
#pragma unroll
    for (int k = 0; k < SF_COL_RUN + 2 * R; ++k) {
      const int y = y0 - R + k;
      v[k] = s[(size_t)((y >= 0 && y < H) ? y : reflect101g(y, H)) * W];
    }
  }
  float* d = dst + (size_t)f * fstride_dst + (size_t)y0 * W + x;
#pragma unroll
  for (int k = 0; k < SF_COL_RUN; ++k) {
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t <= 2 * R; ++t) acc += c_taps[kidx][t] * v[k + t];
    if (y0 + k < H) *d = acc;
    d += W;
  }
}
// one separable blur: rows src -> tmp, columns tmp -> dst (the instance of the kernel's radius)
template <int R>
static void sift_blur_R(cudaStream_t st, int nf, const float* src, size_t fs_src, float* tmp, size_t fs_tmp, float* dst, size_t fs_dst, int W, int H, int kidx) {
  sift_blur_row_kernel<R><<<dim3((W + SF_ROW_T - 1) / SF_ROW_T, H, nf), SF_ROW_T / 8, 0, st>>>(src, tmp, fs_src, fs_tmp, W, H, kidx);
  sift_blur_col_kernel<R><<<dim3((W + 31) / 32, (H + 8 * SF_COL_RUN - 1) / (8 * SF_COL_RUN), nf), 256, 0, st>>>(tmp, dst, fs_tmp, fs_dst, W, H, kidx);
}
static void sift_blur(int r, cudaStream_t st, int nf, const float* src, size_t fs_src, float* tmp, size_t fs_tmp, float* dst, size_t fs_dst, int W, int H, int kidx) {
  switch (r) {
    case 5: sift_blur_R<5>(st, nf, src, fs_src, tmp, fs_tmp, dst, fs_dst, W, H, kidx); break;
    case 6: sift_blur_R<6>(st, nf, src, fs_src, tmp, fs_tmp, dst, fs_dst, W, H, kidx); break;
    case 8: sift_blur_R<8>(st, nf, src, fs_src, tmp, fs_tmp, dst, fs_dst, W, H, kidx); break;
    case 10: sift_blur_R<10>(st, nf, src, fs_src, tmp, fs_tmp, dst, fs_dst, W, H, kidx); break;
    default: sift_blur_R<13>(st, nf, src, fs_src, tmp, fs_tmp, dst, fs_dst, W, H, kidx); break;   // r == 13 (checked by the caller)
  }
}
__global__ void sift_down_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t fstride, int Ws, int Wd, int Hd) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
  if (x >= Wd) return;
  dst[(size_t)f * fstride + (size_t)y * Wd + x] = src[(size_t)f * fstride + (size_t)(2 * y) * Ws + 2 * x];
}

// DoG value of layer i at (r, c): gauss[i + 1] - gauss[i]
struct OctView {
  const float* g; int W, H;
  __device__ __forceinline__ float dog(int i, int r, int c) const {
    const size_t p = (size_t)r * W + c, L = (size_t)W * H;
    return g[(size_t)(i + 1) * L + p] - g[(size_t)i * L + p];
  }
};

// adjustLocalExtrema (sift.simd.hpp); true -> cand filled
__device__ bool sift_adjust(const OctView& V, int o, int layer, int r, int c, SiftCand* out) {
  const float img_scale = 1.f / 255.f, ds = img_scale * 0.5f, ss = img_scale, cs = img_scale * 0.25f;
  float xi = 0, xr = 0, xc = 0;
  int it = 0;
  for (; it < 5; ++it) {
    const float v = V.dog(layer, r, c);
    const float dD0 = (V.dog(layer, r, c + 1) - V.dog(layer, r, c - 1)) * ds, dD1 = (V.dog(layer, r + 1, c) - V.dog(layer, r - 1, c)) * ds,
                dD2 = (V.dog(layer + 1, r, c) - V.dog(layer - 1, r, c)) * ds;
    const float v2 = v * 2;
    const float dxx = (V.dog(layer, r, c + 1) + V.dog(layer, r, c - 1) - v2) * ss, dyy = (V.dog(layer, r + 1, c) + V.dog(layer, r - 1, c) - v2) * ss,
                dss = (V.dog(layer + 1, r, c) + V.dog(layer - 1, r, c) - v2) * ss;
    const float dxy = (V.dog(layer, r + 1, c + 1) - V.dog(layer, r + 1, c - 1) - V.dog(layer, r - 1, c + 1) + V.dog(layer, r - 1, c - 1)) * cs;
    const float dxs = (V.dog(layer + 1, r, c + 1) - V.dog(layer + 1, r, c - 1) - V.dog(layer - 1, r, c + 1) + V.dog(layer - 1, r, c - 1)) * cs;
    const float dys = (V.dog(layer + 1, r + 1, c) - V.dog(layer + 1, r - 1, c) - V.dog(layer - 1, r + 1, c) + V.dog(layer - 1, r - 1, c)) * cs;
    // cv::solve 3 x 3 closed form (Cramer, double)
    const double a00 = dxx, a01 = dxy, a02 = dxs, a10 = dxy, a11 = dyy, a12 = dys, a20 = dxs, a21 = dys, a22 = dss, b0 = dD0, b1 = dD1, b2 = dD2;
    double d = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20);
    float X0 = 0, X1 = 0, X2 = 0;
    if (d != 0.0) {
      d = 1.0 / d;
      X0 = (float)(d * (b0 * (a11 * a22 - a12 * a21) - a01 * (b1 * a22 - a12 * b2) + a02 * (b1 * a21 - a11 * b2)));
      X1 = (float)(d * (a00 * (b1 * a22 - a12 * b2) - b0 * (a10 * a22 - a12 * a20) + a02 * (a10 * b2 - b1 * a20)));
      X2 = (float)(d * (a00 * (a11 * b2 - b1 * a21) - a01 * (a10 * b2 - b1 * a20) + b0 * (a10 * a21 - a11 * a20)));
    }
    xi = -X2; xr = -X1; xc = -X0;
    if (fabsf(xi) < 0.5f && fabsf(xr) < 0.5f && fabsf(xc) < 0.5f) break;
    if (fabsf(xi) > (float)(INT_MAX / 3) || fabsf(xr) > (float)(INT_MAX / 3) || fabsf(xc) > (float)(INT_MAX / 3)) return false;
    c += (int)rintf(xc); r += (int)rintf(xr); layer += (int)rintf(xi);
    if (layer < 1 || layer > SF_LAYERS || c < SF_BORDER || c >= V.W - SF_BORDER || r < SF_BORDER || r >= V.H - SF_BORDER) return false;
  }
  if (it >= 5) return false;
  const float v = V.dog(layer, r, c);
  const float dD0 = (V.dog(layer, r, c + 1) - V.dog(layer, r, c - 1)) * ds, dD1 = (V.dog(layer, r + 1, c) - V.dog(layer, r - 1, c)) * ds,
              dD2 = (V.dog(layer + 1, r, c) - V.dog(layer - 1, r, c)) * ds;
  const float t = dD0 * xc + dD1 * xr + dD2 * xi;
  const float contr = v * img_scale + t * 0.5f;
  if (fabsf(contr) * SF_LAYERS < 0.04f) return false;
  const float v2 = v * 2;
  const float dxx = (V.dog(layer, r, c + 1) + V.dog(layer, r, c - 1) - v2) * ss, dyy = (V.dog(layer, r + 1, c) + V.dog(layer, r - 1, c) - v2) * ss;
  const float dxy = (V.dog(layer, r + 1, c + 1) - V.dog(layer, r + 1, c - 1) - V.dog(layer, r - 1, c + 1) + V.dog(layer, r - 1, c - 1)) * cs;
  const float tr = dxx + dyy, det = dxx * dyy - dxy * dxy;
  if (det <= 0 || tr * tr * 10.f >= (10.f + 1) * (10.f + 1) * det) return false;
  out->o = o; out->layer = layer; out->r = r; out->c = c; out->xi = xi; out->xr = xr; out->xc = xc; out->contr = contr;
  return true;
}

// DoG extrema of all three layers of one octave for a 32 x 8 pixel tile: the five DoG planes of the tile (+ 1 pixel of halo)
// are formed once in shared memory from the six Gaussian planes (each Gaussian value is loaded once and serves two
// differences), the 26-neighbour tests then read shared memory only. (One kernel per layer that differenced on the fly read
// every Gaussian plane 3.6 times through L2: 9 GB per 64 VGA frames, a quarter of the SIFT time.)
#define SF_ET_W 32
#define SF_ET_H 8
__global__ void __launch_bounds__(SF_ET_W * SF_ET_H) sift_extrema_kernel(const float* __restrict__ pyr, SiftGeom G, int o, SiftCand* __restrict__ cand,
                                                                         int* __restrict__ ncand) {
  __shared__ float s_dog[SF_NIMG - 1][SF_ET_H + 2][SF_ET_W + 2];
  const SiftOct O = G.o[o];
  const int f = blockIdx.z, tx = threadIdx.x & (SF_ET_W - 1), ty = threadIdx.x / SF_ET_W;
  const int c0 = SF_BORDER + blockIdx.x * SF_ET_W, r0 = SF_BORDER + blockIdx.y * SF_ET_H;
  const float* g = pyr + (size_t)f * G.frame_floats + O.off;
  const size_t L = (size_t)O.W * O.H;
  for (int e = threadIdx.x; e < (SF_ET_H + 2) * (SF_ET_W + 2); e += SF_ET_W * SF_ET_H) {
    const int yy = e / (SF_ET_W + 2), xx = e - yy * (SF_ET_W + 2);
    const int r = min(r0 - 1 + yy, O.H - 1), c = min(c0 - 1 + xx, O.W - 1);     // clamped positions are never a tested pixel's neighbour
    const float* p = g + (size_t)r * O.W + c;
    float prev = p[0];
#pragma unroll
    for (int d = 0; d < SF_NIMG - 1; ++d) { const float cur = p[(size_t)(d + 1) * L]; s_dog[d][yy][xx] = cur - prev; prev = cur; }
  }
  __syncthreads();
  const int c = c0 + tx, r = r0 + ty;
  if (c >= O.W - SF_BORDER || r >= O.H - SF_BORDER) return;
  OctView V; V.g = g; V.W = O.W; V.H = O.H;
#pragma unroll 1
  for (int layer = 1; layer <= SF_LAYERS; ++layer) {
    const float v = s_dog[layer][ty + 1][tx + 1];
    if (!(fabsf(v) > 1.0f)) continue;                  // threshold = cvFloor(0.5 * 0.04 / 3 * 255) = 1
    bool ismax = v > 0, ismin = v < 0;
#pragma unroll
    for (int di = -1; di <= 1; ++di)
#pragma unroll
      for (int dy = 0; dy <= 2; ++dy)
#pragma unroll
        for (int dx = 0; dx <= 2; ++dx) {
          const float n = s_dog[layer + di][ty + dy][tx + dx];
          ismax = ismax && v >= n; ismin = ismin && v <= n;
        }
    if (!(ismax || ismin)) continue;
    SiftCand cd;
    if (!sift_adjust(V, o, layer, r, c, &cd)) continue;
    const int slot = atomicAdd(&ncand[f], 1);
    if (slot < SF_MAX_CAND) cand[(size_t)f * SF_MAX_CAND + slot] = cd;
  }
}

// cv::fastAtan2 (degrees)
__device__ __forceinline__ float fast_atan2f(float y, float x) {
  const float p1 = 0.9997878412794807f * (float)(180 / 3.141592653589793), p3 = -0.3258083974640975f * (float)(180 / 3.141592653589793),
              p5 = 0.1555786518463281f * (float)(180 / 3.141592653589793), p7 = -0.04432655554792128f * (float)(180 / 3.141592653589793);
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) { c = ay / (ax + (float)DBL_EPSILON); c2 = c * c; a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
  else { c = ax / (ay + (float)DBL_EPSILON); c2 = c * c; a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// calcOrientationHist + the peak loop of findScaleSpaceExtrema: warp per candidate
__global__ void __launch_bounds__(128) sift_orient_kernel(const float* __restrict__ pyr, SiftGeom G, const SiftCand* __restrict__ cand,
                                                          const int* __restrict__ ncand, SiftKp* __restrict__ kps, int* __restrict__ nkp) {
  __shared__ unsigned long long s_acc[4][36];
  __shared__ float s_hist[4][40];
  const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = min(ncand[f], SF_MAX_CAND);
  for (int ci = blockIdx.x * 4 + warp; ci < n; ci += gridDim.x * 4) {
    const SiftCand cd = cand[(size_t)f * SF_MAX_CAND + ci];
    const SiftOct O = G.o[cd.o];
    const float* img = pyr + (size_t)f * G.frame_floats + O.off + (size_t)cd.layer * O.W * O.H;
    const float size = 1.6f * powf(2.f, (cd.layer + cd.xi) / SF_LAYERS) * (1 << cd.o) * 2;
    const float scl_octv = size * 0.5f / (1 << cd.o);
    const int radius = (int)rintf(4.5f * scl_octv);
    const float sigma = 1.5f * scl_octv, expf_scale = -1.f / (2.f * sigma * sigma);
    for (int k = lane; k < 36; k += 32) s_acc[warp][k] = 0ull;
    __syncwarp();
    const int side = 2 * radius + 1;
    for (int e = lane; e < side * side; e += 32) {
      const int i = e / side - radius, j = e % side - radius;
      const int y = cd.r + i, x = cd.c + j;
      if (y <= 0 || y >= O.H - 1 || x <= 0 || x >= O.W - 1) continue;
      const float dx = img[(size_t)y * O.W + x + 1] - img[(size_t)y * O.W + x - 1];
      const float dy = img[(size_t)(y - 1) * O.W + x] - img[(size_t)(y + 1) * O.W + x];
      const float w = expf((float)(i * i + j * j) * expf_scale);
      const float ori = fast_atan2f(dy, dx), mag = sqrtf(dx * dx + dy * dy);
      int bin = (int)rintf((36.f / 360.f) * ori);
      if (bin >= 36) bin -= 36;
      if (bin < 0) bin += 36;
      atomicAdd(&s_acc[warp][bin], (unsigned long long)(long long)llrintf(w * mag * SF_FIX));
    }
    __syncwarp();
    for (int k = lane; k < 36; k += 32) s_hist[warp][k] = (float)((double)(long long)s_acc[warp][k] / (double)SF_FIX);
    __syncwarp();
    float hk[2] = {0, 0};
    float omax = 0;
    for (int q = 0; q < 2; ++q) {
      const int k = lane + 32 * q;
      if (k < 36) {
        const float* t = s_hist[warp];
        hk[q] = (t[(k + 34) % 36] + t[(k + 2) % 36]) * (1.f / 16.f) + (t[(k + 35) % 36] + t[(k + 1) % 36]) * (4.f / 16.f) + t[k] * (6.f / 16.f);
        omax = fmaxf(omax, hk[q]);
      }
    }
    for (int s = 16; s; s >>= 1) omax = fmaxf(omax, __shfl_xor_sync(0xffffffffu, omax, s));
    __syncwarp();
    for (int q = 0; q < 2; ++q) { const int k = lane + 32 * q; if (k < 36) s_hist[warp][k] = hk[q]; }
    __syncwarp();
    const float mag_thr = omax * 0.8f;
    for (int q = 0; q < 2; ++q) {
      const int j = lane + 32 * q;
      if (j >= 36) continue;
      const float* h = s_hist[warp];
      const float hl = h[j > 0 ? j - 1 : 35], hr = h[j < 35 ? j + 1 : 0], hj = h[j];
      if (hj > hl && hj > hr && hj >= mag_thr) {
        float bin = j + 0.5f * (hl - hr) / (hl - 2 * hj + hr);
        bin = bin < 0 ? 36 + bin : (bin >= 36 ? bin - 36 : bin);
        float ang = 360.f - (360.f / 36) * bin;
        if (fabsf(ang - 360.f) < FLT_EPSILON) ang = 0.f;
        const int slot = atomicAdd(&nkp[f], 1);
        if (slot < SF_MAX_KP) {
          SiftKp kp;
          kp.x = (cd.c + cd.xc) * (1 << cd.o) * 0.5f; kp.y = (cd.r + cd.xr) * (1 << cd.o) * 0.5f;   // firstOctave -1: halve
          kp.size = size * 0.5f; kp.angle = ang; kp.response = fabsf(cd.contr); kp.o = cd.o; kp.layer = cd.layer; kp.slot = -1;
          kps[(size_t)f * SF_MAX_KP + slot] = kp;
        }
      }
    }
    __syncwarp();
  }
}

// removeDuplicated + removeDepthless: slot = 1 for the keypoints that stay (the first of equal pt / size / angle, with depth)
__global__ void __launch_bounds__(256) sift_flag_kernel(SiftKp* __restrict__ kps, const int* __restrict__ nkp, const float* __restrict__ depth,
                                                        int W, int H) {
  const int f = blockIdx.y;
  const int n = min(nkp[f], SF_MAX_KP);
  SiftKp* K = kps + (size_t)f * SF_MAX_KP;
  const float* D = depth + (size_t)f * W * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float ax = K[i].x, ay = K[i].y, as = K[i].size, aa = K[i].angle;
    bool alive = ax >= 0.f && ax < (float)W && ay >= 0.f && ay < (float)H;                      // removeDepthless (node.cpp:101-130)
    if (alive) { const int xi = min((int)roundf(ax), W - 1), yi = min((int)roundf(ay), H - 1); alive = !isnan(D[(size_t)yi * W + xi]); }
    for (int j = 0; j < i && alive; ++j) alive = !(K[j].x == ax && K[j].y == ay && K[j].size == as && K[j].angle == aa);
    K[i].slot = alive ? 1 : 0;
  }
}
// retainBest(max_keypoints) + resize: output slot = rank by (response desc, then x, y, size, angle) among the keypoints that stay
__global__ void __launch_bounds__(256) sift_rank_kernel(const SiftKp* __restrict__ kps, const int* __restrict__ nkp, int max_kp,
                                                        int* __restrict__ nout, int* __restrict__ slot2kp) {
  const int f = blockIdx.y;
  const int n = min(nkp[f], SF_MAX_KP);
  const SiftKp* K = kps + (size_t)f * SF_MAX_KP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const SiftKp a = K[i];
    if (!a.slot) continue;
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const SiftKp b = K[j];
      const bool better = b.response > a.response ||
                          (b.response == a.response && (b.x < a.x || (b.x == a.x && (b.y < a.y || (b.y == a.y && (b.size < a.size || (b.size == a.size && b.angle < a.angle)))))));
      rank += (b.slot && j != i && better) ? 1 : 0;
    }
    if (rank < max_kp) { slot2kp[(size_t)f * LSL_MAX_POINTS + rank] = i; atomicAdd(&nout[f], 1); }
  }
}

// calcSIFTDescriptor + projectTo3D for the keypoint of output slot blockIdx.x; rows go to the dense per-frame tables
__global__ void __launch_bounds__(128) sift_describe_kernel(const float* __restrict__ pyr, SiftGeom G, const SiftKp* __restrict__ kps,
                                                            const int* __restrict__ slot2kp, const int* __restrict__ nout,
                                                            const float* __restrict__ depth, int W, int H, float inv_fx, float inv_fy, float cx,
                                                            float cy, float* __restrict__ xyz1, float* __restrict__ desc, float* __restrict__ kpinfo) {
  __shared__ unsigned long long s_h[6 * 6 * 10];
  __shared__ float s_d[128];
  __shared__ float s_red[4];
  const int f = blockIdx.y, slot = blockIdx.x, tid = threadIdx.x;
  if (slot >= min(nout[f], LSL_MAX_POINTS)) return;
  const SiftKp kp = kps[(size_t)f * SF_MAX_KP + slot2kp[(size_t)f * LSL_MAX_POINTS + slot]];
  const SiftOct O = G.o[kp.o];
  const float* img = pyr + (size_t)f * G.frame_floats + O.off + (size_t)kp.layer * O.W * O.H;
  for (int k = tid; k < 360; k += 128) s_h[k] = 0ull;
  const int octave = kp.o - 1;
  const float scale = octave >= 0 ? 1.f / (float)(1 << octave) : (float)(1 << -octave);
  const float size = kp.size * scale, ptx = kp.x * scale, pty = kp.y * scale;
  float angle = 360.f - kp.angle;
  if (fabsf(angle - 360.f) < FLT_EPSILON) angle = 0.f;
  const float scl = size * 0.5f;
  const int px = (int)rintf(ptx), py = (int)rintf(pty);
  float cos_t = cosf(angle * (float)(3.141592653589793 / 180)), sin_t = sinf(angle * (float)(3.141592653589793 / 180));
  const float bins_per_rad = 8 / 360.f, exp_scale = -1.f / (4 * 4 * 0.5f), hist_width = 3.f * scl;
  int radius = (int)rintf(hist_width * 1.4142135623730951f * (4 + 1) * 0.5f);
  radius = min(radius, (int)sqrt((double)O.W * O.W + (double)O.H * O.H));
  cos_t /= hist_width; sin_t /= hist_width;
  __syncthreads();
  const int side = 2 * radius + 1;
  for (int e = tid; e < side * side; e += 128) {
    const int i = e / side - radius, j = e % side - radius;
    const float c_rot = j * cos_t - i * sin_t, r_rot = j * sin_t + i * cos_t;
    float rbin = r_rot + 4 / 2 - 0.5f, cbin = c_rot + 4 / 2 - 0.5f;
    const int r = py + i, c = px + j;
    if (!(rbin > -1 && rbin < 4 && cbin > -1 && cbin < 4 && r > 0 && r < O.H - 1 && c > 0 && c < O.W - 1)) continue;
    const float dx = img[(size_t)r * O.W + c + 1] - img[(size_t)r * O.W + c - 1];
    const float dy = img[(size_t)(r - 1) * O.W + c] - img[(size_t)(r + 1) * O.W + c];
    const float w = expf((c_rot * c_rot + r_rot * r_rot) * exp_scale);
    const float ori = fast_atan2f(dy, dx);
    const float mag = sqrtf(dx * dx + dy * dy) * w;
    float obin = (ori - angle) * bins_per_rad;
    const int r0 = (int)floorf(rbin), c0 = (int)floorf(cbin);
    int o0 = (int)floorf(obin);
    rbin -= r0; cbin -= c0; obin -= o0;
    if (o0 < 0) o0 += 8;
    if (o0 >= 8) o0 -= 8;
    const float v_r1 = mag * rbin, v_r0 = mag - v_r1;
    const float v_rc11 = v_r1 * cbin, v_rc10 = v_r1 - v_rc11, v_rc01 = v_r0 * cbin, v_rc00 = v_r0 - v_rc01;
    const float v_rco111 = v_rc11 * obin, v_rco110 = v_rc11 - v_rco111, v_rco101 = v_rc10 * obin, v_rco100 = v_rc10 - v_rco101;
    const float v_rco011 = v_rc01 * obin, v_rco010 = v_rc01 - v_rco011, v_rco001 = v_rc00 * obin, v_rco000 = v_rc00 - v_rco001;
    const int idx = ((r0 + 1) * 6 + c0 + 1) * 10 + o0;
#define SF_ADD(k, v) atomicAdd(&s_h[k], (unsigned long long)(long long)llrintf((v) * SF_FIX))
    SF_ADD(idx, v_rco000); SF_ADD(idx + 1, v_rco001); SF_ADD(idx + 10, v_rco010); SF_ADD(idx + 11, v_rco011);
    SF_ADD(idx + 60, v_rco100); SF_ADD(idx + 61, v_rco101); SF_ADD(idx + 70, v_rco110); SF_ADD(idx + 71, v_rco111);
#undef SF_ADD
  }
  __syncthreads();
  {  // finalise: circular orientation bins, element tid = (i * 4 + j) * 8 + k
    const int k = tid & 7, ij = tid >> 3, i = ij >> 2, j = ij & 3;
    const int idx = ((i + 1) * 6 + (j + 1)) * 10;
    double v = (double)(long long)s_h[idx + k] / (double)SF_FIX;
    if (k < 2) v += (double)(long long)s_h[idx + 8 + k] / (double)SF_FIX;
    s_d[tid] = (float)v;
  }
  __syncthreads();
  auto block_sum_sq = [&]() {
    float v = s_d[tid] * s_d[tid];
    for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    const float t = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
    __syncthreads();
    return t;
  };
  float nrm2 = block_sum_sq();
  const float thr = sqrtf(nrm2) * 0.2f;
  s_d[tid] = fminf(s_d[tid], thr);
  __syncthreads();
  nrm2 = block_sum_sq();
  const float fac = 512.f / fmaxf(sqrtf(nrm2), FLT_EPSILON);
  const float q = fminf(fmaxf(rintf(s_d[tid] * fac), 0.f), 255.f);          // saturate_cast<uchar>
  desc[((size_t)f * LSL_MAX_POINTS + slot) * 128 + tid] = q;
  if (tid == 0) {  // projectTo3D (node.cpp:952-1018): float arithmetic, depth at the rounded pixel
    const int xi = min((int)roundf(kp.x), W - 1), yi = min((int)roundf(kp.y), H - 1);
    const float Z = depth[(size_t)f * W * H + (size_t)yi * W + xi];
    float* o = xyz1 + ((size_t)f * LSL_MAX_POINTS + slot) * 4;
    o[0] = (kp.x - cx) * Z * inv_fx; o[1] = (kp.y - cy) * Z * inv_fy; o[2] = Z; o[3] = 1.f;
    float* ki = kpinfo + ((size_t)f * LSL_MAX_POINTS + slot) * 6;
    ki[0] = kp.x; ki[1] = kp.y; ki[2] = kp.size; ki[3] = kp.angle; ki[4] = kp.response; ki[5] = (float)(octave + 256 * kp.layer);
  }
}

// rows of the dense per-frame tables -> the batch block (frame f starts at goff[f])
__global__ void sift_pack_kernel(const float* __restrict__ xyz1, const float* __restrict__ desc, const float* __restrict__ kpinfo,
                                 const int* __restrict__ nout, const int* __restrict__ goff, float* __restrict__ b_xyz1, float* __restrict__ b_desc,
                                 float* __restrict__ b_kp) {
  const int f = blockIdx.y, n = min(nout[f], LSL_MAX_POINTS);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * 138; e += gridDim.x * blockDim.x) {
    const int row = e / 138, k = e % 138;
    const size_t src = (size_t)f * LSL_MAX_POINTS + row, dst = (size_t)goff[f] + row;
    if (k < 128) b_desc[dst * 128 + k] = desc[src * 128 + k];
    else if (k < 132) b_xyz1[dst * 4 + (k - 128)] = xyz1[src * 4 + (k - 128)];
    else b_kp[dst * 6 + (k - 132)] = kpinfo[src * 6 + (k - 132)];
  }
}

// ------------------------------------------------------------------------------------------------ host ----
static void sift_geometry(int W, int H, SiftGeom* G) {
  G->W0 = 2 * W; G->H0 = 2 * H;
  const int mn = G->W0 < G->H0 ? G->W0 : G->H0;
  int n = (int)rint(log((double)mn) / log(2.0) - 2) + 1;                    // cvRound(log2(min) - 2) - firstOctave
  if (n > SF_MAX_OCT) n = SF_MAX_OCT;
  if (n < 1) n = 1;
  size_t off = 0;
  int w = G->W0, h = G->H0;
  G->n_oct = 0;
  for (int o = 0; o < n && w >= 1 && h >= 1; ++o) {
    G->o[o].W = w; G->o[o].H = h; G->o[o].off = off;
    off += ((size_t)SF_NIMG * w * h + 3) & ~(size_t)3;      // octave bases (and the frame stride) stay 16-byte aligned: float4 staging
    G->n_oct = o + 1;
    w /= 2; h /= 2;
  }
  G->frame_floats = off;
}

static void sift_taps(double sigma, float* taps, int* rad) {                 // getGaussianKernel(cvRound(sigma * 8 + 1) | 1, sigma)
  const int ksize = ((int)rint(sigma * 8 + 1)) | 1;
  double k[64], sum = 0;
  for (int i = 0; i < ksize; ++i) { const double x = i - (ksize - 1) * 0.5; k[i] = exp(-0.5 * x * x / (sigma * sigma)); sum += k[i]; }
  for (int i = 0; i < 40; ++i) taps[i] = i < ksize ? (float)(k[i] / sum) : 0.f;
  *rad = ksize / 2;
}

// Detects and attaches the point features of the n frames just extracted (gray planes w.gray, depth planes d_depth).
int lsl_launch_sift(lsl_ctx* ctx, int n, const float* d_depth, int W, int H, const double K[9], lsl_frame** frames) {
  LslSiftWork& S = ctx->sift;
  cudaStream_t st = ctx->stream;
  SiftGeom G;
  sift_geometry(W, H, &G);
  const int max_kp = ctx->sift_max_kp < LSL_MAX_POINTS ? ctx->sift_max_kp : LSL_MAX_POINTS;
  const int chunk = ctx->max_batch < SF_CHUNK ? ctx->max_batch : SF_CHUNK;   // frames per pyramid pass of this context
  // ---- workspace: pyramid of `chunk` frames + one row-pass scratch plane per frame + lists + dense output tables for n frames
  const size_t pyr_floats = G.frame_floats * chunk, tmp_floats = (size_t)G.W0 * G.H0 * chunk;
  const size_t need = (pyr_floats + tmp_floats) * sizeof(float) + (size_t)chunk * (SF_MAX_CAND * sizeof(SiftCand) + SF_MAX_KP * sizeof(SiftKp)) +
                      (size_t)ctx->max_batch * ((size_t)LSL_MAX_POINTS * (4 + 128 + 6) * 4 + 8) +
                      (size_t)chunk * (LSL_MAX_POINTS * sizeof(int) + 8) + 4096;
  if (S.bytes < need) {
    LSL_CUDA(cudaStreamSynchronize(st));
    if (S.block) cudaFree(S.block);
    S.block = nullptr; S.bytes = 0;
    LSL_CUDA(cudaMalloc(&S.block, need));
    S.bytes = need;
  }
  uint8_t* base = (uint8_t*)S.block;
  float* pyr = (float*)base; base += pyr_floats * sizeof(float);
  float* tmp = (float*)base; base += tmp_floats * sizeof(float);
  SiftCand* cand = (SiftCand*)base; base += (size_t)chunk * SF_MAX_CAND * sizeof(SiftCand);
  SiftKp* kps = (SiftKp*)base; base += (size_t)chunk * SF_MAX_KP * sizeof(SiftKp);
  const size_t B = (size_t)ctx->max_batch;
  float* t_xyz1 = (float*)base; base += B * LSL_MAX_POINTS * 4 * sizeof(float);
  float* t_desc = (float*)base; base += B * LSL_MAX_POINTS * 128 * sizeof(float);
  float* t_kp = (float*)base; base += B * LSL_MAX_POINTS * 6 * sizeof(float);
  int* slot2kp = (int*)base; base += (size_t)chunk * LSL_MAX_POINTS * sizeof(int);
  int* counters = (int*)base;                            // [0..C) ncand, [C..2C) nkp, [2C..2C+B) nout, then goff [B]
  int* d_ncand = counters; int* d_nkp = counters + chunk; int* d_nout = counters + 2 * chunk; int* d_goff = d_nout + B;
  int rad[SF_NIMG];
  {
    float taps[SF_NIMG][40];
    const double sigma = 1.6, k = pow(2.0, 1.0 / SF_LAYERS);
    sift_taps((double)(float)sqrt(fmax(sigma * sigma - 0.5 * 0.5 * 4, 0.01)), taps[0], &rad[0]);       // createInitialImage
    for (int i = 1; i < SF_NIMG; ++i) { const double sp = pow(k, (double)(i - 1)) * sigma, stt = sp * k; sift_taps(sqrt(stt * stt - sp * sp), taps[i], &rad[i]); }
    static std::mutex mu;
    static bool uploaded[64] = {false};
    std::lock_guard<std::mutex> lk(mu);
    if (!uploaded[ctx->device & 63]) {   // the tap tables are constants of the algorithm: once per device
      LSL_CUDA(cudaMemcpyToSymbolAsync(c_taps, taps, sizeof(taps), 0, cudaMemcpyHostToDevice, st));
      LSL_CUDA(cudaMemcpyToSymbolAsync(c_rad, rad, sizeof(rad), 0, cudaMemcpyHostToDevice, st));
      LSL_CUDA(cudaStreamSynchronize(st));   // the host arrays go out of scope
      uploaded[ctx->device & 63] = true;
    }
  }
  for (int i = 0; i < SF_NIMG; ++i)
    if (rad[i] != 5 && rad[i] != 6 && rad[i] != 8 && rad[i] != 10 && rad[i] != 13) { ctx->err = "SIFT blur radius without a column-pass instance"; return LSL_ERR_ARG; }
  LSL_CUDA(cudaMemsetAsync(d_nout, 0, sizeof(int) * n, st));
  const float inv_fx = (float)(1.0 / K[0]), inv_fy = (float)(1.0 / K[4]), cx = (float)K[2], cy = (float)K[5];
  // LSL_SIFT_PROFILE=1: device time per phase (pyramid | extrema | orientation | flag + rank | descriptors) on stderr
  static const bool prof = getenv("LSL_SIFT_PROFILE") != nullptr;
  cudaEvent_t pe[6]; float pms[5] = {0, 0, 0, 0, 0};
  if (prof) for (int i = 0; i < 6; ++i) cudaEventCreate(&pe[i]);
#define SF_MARK(i) do { if (prof) cudaEventRecord(pe[i], st); } while (0)
  LSL_KSTART(ctx, LSL_K_SIFT);
  for (int f0 = 0; f0 < n; f0 += chunk) {
    const int nf = n - f0 < chunk ? n - f0 : chunk;
    const uint8_t* gray = ctx->wk.gray + (size_t)f0 * W * H;
    const float* dep = d_depth + (size_t)f0 * W * H;
    LSL_CUDA(cudaMemsetAsync(counters, 0, sizeof(int) * 2 * chunk, st));
    SF_MARK(0);
    // base image: up-sample into layer 1 (scratch), blur rows -> tmp, columns -> layer 0
    {
      const SiftOct O = G.o[0];
      const size_t L = (size_t)O.W * O.H;
      dim3 g((O.W + 127) / 128, O.H, nf);
      sift_upsample_kernel<<<g, 128, 0, st>>>(gray, pyr + L, G, W, H);
      sift_blur(rad[0], st, nf, pyr + L, G.frame_floats, tmp, L, pyr, G.frame_floats, O.W, O.H, 0);
      ctx->stats.kernel_launches += 3;
    }
    for (int o = 0; o < G.n_oct; ++o) {
      const SiftOct O = G.o[o];
      const size_t L = (size_t)O.W * O.H;
      dim3 g((O.W + 127) / 128, O.H, nf);
      if (o > 0) {
        const SiftOct P = G.o[o - 1];
        sift_down_kernel<<<g, 128, 0, st>>>(pyr + P.off + (size_t)SF_LAYERS * P.W * P.H - 0, pyr + O.off, G.frame_floats, P.W, O.W, O.H);
        ctx->stats.kernel_launches += 1;
      }
      for (int i = 1; i < SF_NIMG; ++i) {
        sift_blur(rad[i], st, nf, pyr + O.off + (size_t)(i - 1) * L, G.frame_floats, tmp, L, pyr + O.off + (size_t)i * L, G.frame_floats, O.W, O.H, i);
        ctx->stats.kernel_launches += 2;
      }
      if (O.W > 2 * SF_BORDER && O.H > 2 * SF_BORDER) {
        dim3 ge((O.W - 2 * SF_BORDER + SF_ET_W - 1) / SF_ET_W, (O.H - 2 * SF_BORDER + SF_ET_H - 1) / SF_ET_H, nf);
        sift_extrema_kernel<<<ge, SF_ET_W * SF_ET_H, 0, st>>>(pyr, G, o, cand, d_ncand);
        ctx->stats.kernel_launches += 1;
      }
    }
    SF_MARK(1);
    sift_orient_kernel<<<dim3(256, nf), 128, 0, st>>>(pyr, G, cand, d_ncand, kps, d_nkp);
    SF_MARK(2);
    sift_flag_kernel<<<dim3(32, nf), 256, 0, st>>>(kps, d_nkp, dep, W, H);
    sift_rank_kernel<<<dim3(32, nf), 256, 0, st>>>(kps, d_nkp, max_kp, d_nout + f0, slot2kp);
    SF_MARK(3);
    sift_describe_kernel<<<dim3(max_kp, nf), 128, 0, st>>>(pyr, G, kps, slot2kp, d_nout + f0, dep, W, H, inv_fx, inv_fy, cx, cy,
                                                           t_xyz1 + (size_t)f0 * LSL_MAX_POINTS * 4, t_desc + (size_t)f0 * LSL_MAX_POINTS * 128,
                                                           t_kp + (size_t)f0 * LSL_MAX_POINTS * 6);
    ctx->stats.kernel_launches += 4;
    SF_MARK(4);
    if (prof) {
      cudaEventSynchronize(pe[4]);
      for (int i = 0; i < 4; ++i) { float ms = 0; cudaEventElapsedTime(&ms, pe[i], pe[i + 1]); pms[i] += ms; }
    }
  }
#undef SF_MARK
  if (prof) {
    fprintf(stderr, "sift phases (%d frames): pyramid+extrema %.2f ms, orientation %.2f ms, flag+rank %.2f ms, descriptors %.2f ms\n", n, pms[0], pms[1], pms[2], pms[3]);
    for (int i = 0; i < 6; ++i) cudaEventDestroy(pe[i]);
  }
  LSL_CUDA(cudaGetLastError());
  // ---- counts back (4 bytes per frame), one block for the batch, pack, optional RootSIFT, attach
  std::vector<int> nout(n), goff(n);
  LSL_CUDA(cudaMemcpyAsync(nout.data(), d_nout, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
  LSL_CUDA(cudaStreamSynchronize(st));
  ctx->stats.d2h_bytes += 4 * n;
  size_t tot = 0;
  for (int f = 0; f < n; ++f) { if (nout[f] > max_kp) nout[f] = max_kp; goff[f] = (int)tot; tot += (size_t)nout[f]; }
  LslPointBlock* blk = nullptr;
  if (tot) {
    blk = new (std::nothrow) LslPointBlock();
    if (!blk) return LSL_ERR_ARG;
    blk->refs = 0; blk->d_xyz1 = nullptr; blk->d_desc = nullptr; blk->d_kp = nullptr; blk->pooled = true;
    // one stream-ordered allocation per batch in 4 MB size classes (consecutive batches reuse each other's blocks; a plain
    // cudaMalloc / cudaFree pair per call synchronises the device — and with it the pair stream of the previous batch)
    const size_t rows = (tot + tot / 8 + 7167) / 7168 * 7168;
    uint8_t* pb = nullptr;
    LSL_CUDA(cudaMallocAsync((void**)&pb, rows * (4 + 128 + 6) * sizeof(float), st));
    blk->d_xyz1 = (float*)pb; blk->d_desc = pb + rows * 4 * sizeof(float); blk->d_kp = (float*)(pb + rows * (4 + 128) * sizeof(float));
    LSL_CUDA(cudaMemcpyAsync(d_goff, goff.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
    sift_pack_kernel<<<dim3(64, n), 256, 0, st>>>(t_xyz1, t_desc, t_kp, d_nout, d_goff, blk->d_xyz1, (float*)blk->d_desc, blk->d_kp);
    ctx->stats.kernel_launches += 1;
    if (ctx->sift_root) { int rc = lsl_launch_rootsift(ctx, (float*)blk->d_desc, (int)tot, 128); if (rc) return rc; }
  }
  LSL_KSTOP(ctx, LSL_K_SIFT);
  ctx->stats.kernel_launches -= 1;   // LSL_KSTOP counts one launch; the launches were counted one by one above
  for (int f = 0; f < n; ++f) {
    lsl_frame* fr = frames[f];
    fr->npoints = nout[f]; fr->pdim = 128; fr->pkind = 0; fr->pblk = nullptr; fr->d_xyz1 = nullptr; fr->d_desc = nullptr; fr->d_kp = nullptr;
    if (nout[f]) {
      fr->pblk = blk; blk->refs += 1;
      fr->d_xyz1 = blk->d_xyz1 + 4 * (size_t)goff[f];
      fr->d_desc = (float*)blk->d_desc + 128 * (size_t)goff[f];
      fr->d_kp = blk->d_kp + 6 * (size_t)goff[f];
    }
  }
  LSL_CUDA(cudaStreamSynchronize(st));
  LSL_CUDA(cudaGetLastError());
  return LSL_OK;
}
