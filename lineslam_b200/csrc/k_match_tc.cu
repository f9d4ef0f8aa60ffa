// k_match_tc.cu — tensor-core pre-filter for the descriptor-distance matrix of Node::featureMatching
// (src/node.cpp:606-641: BFMatcher L2 knnMatch k = 2 over up to 600 x 600 rows of 64 / 128 floats — the one genuinely
// dense GEMM of the path; BASELINE north star: "tensor cores used only for the descriptor-distance matrix").
//
//   match_points_tc_kernel      G = Q T^T on the 5th-generation tensor cores: tcgen05.mma kind::tf32, M = 128 query rows
//                               x N = 128 train rows x K = dim per CTA step, operands staged in shared memory in the
//                               canonical K-major no-swizzle core-matrix layout, accumulator in TMEM, completion through
//                               tcgen05.commit on an mbarrier, read back with tcgen05.ld. Epilogue (thread = query row):
//                               interval [lo, hi] = (|q|^2 + |t|^2 - 2 G) -+ eps around the EXACT OpenCV-order distance,
//                               running second-smallest upper bound U2; every train row whose lower bound is <= U2 is
//                               appended to the row's candidate list.
//   match_points_refine_kernel  exact re-evaluation (l2sqr_f: OpenCV 2.4 summation order, sqrtf) of the candidates only,
//                               k = 2 merge with the sequential-scan tie rule -> the same Knn2 records the exact kernel
//                               (k_hybrid.cu: match_points_kernel) writes. A row whose list overflowed is rescanned in full.
//
// Why the result is identical to the exact kernel: with S the exact value the reference computes and |S~ - S| <= eps,
// two rows have S <= their hi <= U2 at any time, so a row with lo > U2 is strictly farther than two others and can be
// neither the nearest nor the second nearest neighbour; rows that tie after sqrtf differ by an ulp, far inside eps.
// eps_ij = TC_EPS_REL * (|q_i|^2 + |t_j|^2): tf32 operand truncation (2 * 2^-10 relative per product, Cauchy-Schwarz +
// AM-GM), tensor-core accumulation, the f32 norms and the f32 rounding of the reference sum itself (tests assert the
// measured error stays below half of it, tests/test_gpu_hybrid.py).
#include "lsl_internal.h"
#include "shared/lsl_points.h"
#include <float.h>
#include <mutex>

using namespace lslm;

#ifndef FULL
#define FULL 0xffffffffu
#endif
#define TC_M 128            // query rows per CTA (UMMA M)
#define TC_N 128            // train rows per MMA step (UMMA N, TMEM columns)
#define TC_KMAX 128         // descriptor length limit of this path
#define TC_CAP 64           // candidates kept per query row
#define TC_EPS_REL 2.5e-3f  // (2^-9 + 2^-12) + 3e-4 head-room, relative to |q|^2 + |t|^2

struct Knn2 { float d1, d2; int i1; };   // same record as k_hybrid.cu
__device__ __forceinline__ Knn2 knn_merge2(Knn2 a, Knn2 b) {
  if (b.d1 < a.d1 || (b.d1 == a.d1 && b.i1 < a.i1)) { Knn2 t = a; a = b; b = t; }
  a.d2 = fminf(a.d2, b.d1);
  return a;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle ("interleave") canonical layout, in bytes: element (r, k) of a tile with `kd` floats per row sits at
// (r % 8) * 16 + (r / 8) * SBO + (k / 4) * 128 + (k % 4) * 4 with SBO = (kd / 4) * 128: 8 x 16 B core matrices, the
// core matrices of one 8-row group contiguous along K (LBO = 128 B).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((128u >> 4) & 0x3fffu) << 16;          // leading byte offset: next core matrix along K
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;     // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
  return d;                                              // base offset 0, layout type 0 = SWIZZLE_NONE
}

// stages `rows` descriptor rows (global, row-major, dim floats) into the canonical layout, zero-filling rows >= nvalid;
// returns nothing; norms[r] = sum of squares in f32 (thread r, r < rows)
__device__ __forceinline__ void stage_rows(const float* __restrict__ g, int first, int nvalid, int dim, float* tile, float* norms) {
  const int tid = threadIdx.x;
  const int kc_n = dim >> 2;                       // 16-byte pieces per row
  const uint32_t sbo = (uint32_t)kc_n * 128u;
  for (int e = tid; e < TC_M * kc_n; e += blockDim.x) {
    const int r = e / kc_n, kc = e - r * kc_n;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (first + r < nvalid) v = *reinterpret_cast<const float4*>(g + (size_t)(first + r) * dim + 4 * kc);
    *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(tile) + (r & 7) * 16 + (r >> 3) * sbo + kc * 128) = v;
  }
  if (tid < TC_M) {
    float s = 0.f;
    if (first + tid < nvalid) {
      const float* row = g + (size_t)(first + tid) * dim;
      for (int k = 0; k < dim; ++k) s += row[k] * row[k];
    }
    norms[tid] = s;
  }
}

struct TcOut {
  int32_t* cand;    // [rows][TC_CAP] candidate train indices
  int32_t* cnt;     // [rows] number of candidates (> TC_CAP: overflow, rescan the row exactly)
};

__global__ void __launch_bounds__(128, 1) match_points_tc_kernel(const LslPairPts* __restrict__ pp, TcOut out) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  float* tileA = reinterpret_cast<float*>(tc_smem);                       // TC_M x dim
  float* tileB = tileA + TC_M * TC_KMAX;                                  // TC_N x dim
  __shared__ float s_na[TC_M], s_nb[TC_N];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const LslPairPts pd = pp[blockIdx.y];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * TC_M;
  if (q0 >= pd.nqp || pd.ntp < 2) return;                                 // uniform
  const int dim = pd.dim;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(TC_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  stage_rows(pd.qd, q0, pd.nqp, dim, tileA, s_na);
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = s_tmem;
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = TC_N, M = TC_M
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  const uint32_t sbo = (uint32_t)(dim >> 2) * 128u;
  const int row = q0 + tid;                       // this thread's query row = TMEM lane tid
  const float na = s_na[tid];
  float u1 = FLT_MAX, u2 = FLT_MAX;               // smallest / second smallest upper bound so far
  int cnt = 0;
  int32_t* my_cand = out.cand + (pd.knn_off + (size_t)row) * TC_CAP;
  uint32_t phase = 0;
  for (int t0 = 0; t0 < pd.ntp; t0 += TC_N) {
    stage_rows(pd.td, t0, pd.ntp, dim, tileB, s_nb);
    asm volatile("fence.proxy.async.shared::cta;\n");      // generic-proxy stores -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;\n");  // the previous step's tcgen05.ld are done before TMEM is overwritten
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      const uint32_t a0 = smem_u32(tileA), b0 = smem_u32(tileB);
      for (int k = 0; k < dim; k += 8) {                     // one UMMA per 8 floats of K (two core matrices = 256 B)
        const uint64_t da = umma_desc(a0 + (uint32_t)(k >> 2) * 128u, sbo), db = umma_desc(b0 + (uint32_t)(k >> 2) * 128u, sbo);
        const uint32_t acc = k ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&s_bar)));
    }
    {  // wait for the MMAs of this step
      uint32_t done = 0;
      const uint32_t bar = smem_u32(&s_bar);
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(phase));
      }
      phase ^= 1u;
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    // epilogue: thread = row (TMEM lane), 32 columns at a time
#pragma unroll 1
    for (int c0 = 0; c0 < TC_N; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
          "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      if (row < pd.nqp) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int j = t0 + c0 + c;
          if (j < pd.ntp) {
            const float nb = s_nb[c0 + c];
            const float sa = (na + nb) - 2.0f * __uint_as_float(v[c]);
            const float eps = TC_EPS_REL * (na + nb);
            const float lo = sa - eps, hi = sa + eps;
            if (!(lo > u2)) {                       // also taken when anything is NaN: such a row overflows into the exact rescan
              if (cnt < TC_CAP) my_cand[cnt] = j;
              ++cnt;
              if (!(hi == hi)) cnt = TC_CAP + 1;
            }
            if (hi < u1) { u2 = u1; u1 = hi; } else if (hi < u2) u2 = hi;
          }
        }
      }
    }
    __syncthreads();   // every thread is done with s_nb before the next step's staging overwrites it
  }
  if (row < pd.nqp) out.cnt[pd.knn_off + row] = cnt;
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(TC_N));
}

// exact k = 2 over the candidates of each query row (warp per row); identical records to match_points_kernel
__global__ void __launch_bounds__(256) match_points_refine_kernel(const LslPairPts* __restrict__ pp, TcOut out, Knn2* __restrict__ knn_all,
                                                                  unsigned long long* __restrict__ stats) {
  extern __shared__ float s_q[];   // [8][dim]
  const LslPairPts pd = pp[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  if (i >= pd.nqp || pd.ntp < 2) return;
  float* q = s_q + warp * pd.dim;
  for (int k = lane; k < pd.dim; k += 32) q[k] = pd.qd[(size_t)i * pd.dim + k];
  __syncwarp();
  const int cnt = out.cnt[pd.knn_off + i];
  const int32_t* cand = out.cand + (pd.knn_off + (size_t)i) * TC_CAP;
  Knn2 best; best.d1 = FLT_MAX; best.d2 = FLT_MAX; best.i1 = 1 << 30;
  const bool full = cnt > TC_CAP;
  const int n = full ? pd.ntp : cnt;
  for (int c = lane; c < n; c += 32) {
    const int j = full ? c : cand[c];
    const float d = sqrtf(l2sqr_f(q, pd.td + (size_t)j * pd.dim, pd.dim));
    // candidates arrive in increasing train index per lane stride: the sequential-scan rule (strict <) keeps the earlier row
    if (d < best.d1 || (d == best.d1 && j < best.i1)) { best.d2 = best.d1; best.d1 = d; best.i1 = j; }
    else if (d < best.d2) best.d2 = d;
  }
  for (int o = 16; o; o >>= 1) {
    Knn2 other;
    other.d1 = __shfl_xor_sync(FULL, best.d1, o); other.d2 = __shfl_xor_sync(FULL, best.d2, o);
    other.i1 = __shfl_xor_sync(FULL, best.i1, o);
    best = knn_merge2(best, other);
  }
  if (lane == 0) {
    knn_all[pd.knn_off + i] = best;
    if (stats) { atomicAdd(&stats[0], (unsigned long long)n); atomicAdd(&stats[1], full ? 1ull : 0ull); atomicAdd(&stats[2], 1ull); }
  }
}

// launcher: returns LSL_OK and *used = 1 when the tensor-core path ran (dim multiple of 8, <= 128)
int lsl_launch_match_points_tc(lsl_ctx* ctx, int npairs, int max_nq, int dim, int* used) {
  *used = 0;
  if (dim < 8 || dim > TC_KMAX || (dim & 7) || max_nq <= 0) return LSL_OK;
  LslHybWork& h = ctx->hw;
  if (h.cap_tc < h.cap_knn) {
    LSL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h.tc_cand) cudaFree(h.tc_cand);
    if (h.tc_cnt) cudaFree(h.tc_cnt);
    h.tc_cand = nullptr; h.tc_cnt = nullptr; h.cap_tc = 0;
    // rows are addressed by knn_off + row with row < ceil(nq / 128) * 128: pad by one tile
    LSL_CUDA(cudaMalloc((void**)&h.tc_cand, sizeof(int32_t) * (h.cap_knn + TC_M) * TC_CAP));
    LSL_CUDA(cudaMalloc((void**)&h.tc_cnt, sizeof(int32_t) * (h.cap_knn + TC_M)));
    if (!h.tc_stats) { LSL_CUDA(cudaMalloc((void**)&h.tc_stats, 3 * sizeof(unsigned long long))); LSL_CUDA(cudaMemsetAsync(h.tc_stats, 0, 24, ctx->stream)); }
    h.cap_tc = h.cap_knn;
  }
  {
    static std::mutex mu;
    static bool set[64] = {false};
    std::lock_guard<std::mutex> lk(mu);
    if (!set[ctx->device & 63]) {
      cudaFuncSetAttribute(match_points_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TC_M * TC_KMAX * sizeof(float)));
      set[ctx->device & 63] = true;
    }
  }
  TcOut o; o.cand = h.tc_cand; o.cnt = h.tc_cnt;
  dim3 g1((max_nq + TC_M - 1) / TC_M, npairs);
  match_points_tc_kernel<<<g1, 128, 2 * TC_M * TC_KMAX * sizeof(float), ctx->stream>>>(h.d_ppairs, o);
  dim3 g2((max_nq + 7) / 8, npairs);
  match_points_refine_kernel<<<g2, 256, 8 * dim * sizeof(float), ctx->stream>>>(h.d_ppairs, o, (Knn2*)h.knn, (unsigned long long*)h.tc_stats);
  ctx->stats.kernel_launches += 2;
  LSL_CUDA(cudaGetLastError());
  *used = 1;
  return LSL_OK;
}

// counters of the refine kernel since context creation: exact distance evaluations, rows rescanned in full, rows
extern "C" int lsl_match_tc_stats(lsl_ctx* ctx, int64_t out[3], int reset) {
  if (!ctx || !out) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  out[0] = out[1] = out[2] = 0;
  if (!ctx->hw.tc_stats) return LSL_OK;
  LSL_CUDA(cudaStreamSynchronize(ctx->stream));
  LSL_CUDA(cudaMemcpy(out, ctx->hw.tc_stats, 24, cudaMemcpyDeviceToHost));
  if (reset) LSL_CUDA(cudaMemset(ctx->hw.tc_stats, 0, 24));
  return LSL_OK;
}
