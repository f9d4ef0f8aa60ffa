// k_tum.cu — TUM raw-directory ingest (include/lsl_tum.h): syncidx.txt parsing and the PNG
// container on the host (chunk walk, IDAT payloads gathered into one pinned buffer); DEFLATE, scan-line unfiltering
// and layout conversion on the device. What crosses PCIe is the COMPRESSED data (≈0.4 MB per VGA image).
//
//   K18  png_inflate_kernel    one warp per zlib stream (shared/lsl_inflate.h): all lanes walk the stream convergently
//        (bit reader in registers, Huffman tables in shared memory built by lane 0), lane 0 stores literals, the warp
//        shares LZ77 copies. DEFLATE is bit-serial inside a stream; the parallelism is the 2 x batch streams in flight.
//   K19  png_unfilter_kernel   one CTA per image, one thread per scan line. PNG's filters make pixel (x, y) depend on
//        its left, upper and upper-left neighbours, so the rows advance as a wavefront: thread y reconstructs pixel
//        x = t - y at step t, reads the pixel above from the slot thread y - 1 published one step earlier (shared
//        memory, double-buffered by step parity) and keeps its own left pixel and the previous "above" in registers.
//        W + H - 1 steps per image instead of W * H serial byte operations; the output is written in the layout the
//        extraction kernels read (BGR u8 like cv::imread(…, 1), or float metres with NaN for "no reading").
//
// HBM traffic per VGA frame: 0.92 + 0.61 MB of filtered scan lines in, 0.92 MB BGR + 1.23 MB float depth out.
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <algorithm>
#include <vector>
#include "lsl_internal.h"
#include "shared/lsl_inflate.h"
#include "../../include/lsl_tum.h"

namespace {

// ---------------------------------------------------------------------------------------------- device ----
// Output side of the decoder on the device: the last LSL_INF_RING bytes of the stream live in a shared-memory ring
// (DEFLATE matches reach at most 32 KB back), literals and match copies touch only the ring, and every completed
// LSL_INF_CHUNK bytes leave for global memory as one coalesced 16-byte-per-lane store pass. Without the ring every
// match copy was a round trip to L2 for bytes this warp had just written (357 ms per 2 x 592 VGA images).
#define LSL_INF_RING 8192     // 4 chunks. Matches that reach further back than LSL_INF_NEAR bytes (rare in image data: the rows
                              // a PNG filter refers to are 1.3 - 3.8 KB back) read their source from the output in global memory,
                              // where every chunk older than the one in progress already is. 8 KB instead of the full 32 KB window
                              // + one chunk is what lets 16 instead of 5 streams be resident per SM.
#define LSL_INF_NEAR (LSL_INF_RING - 768)   // dist <= NEAR: every source byte is still in the ring when it is read (the longest
                                            // match writes 258 bytes ahead); dist > NEAR: every source byte is older than the
                                            // chunk in progress (NEAR > chunk + longest match), i.e. flushed
#define LSL_INF_CHUNK 2048
// stores output bytes [first, first + LSL_INF_CHUNK) from the ring; rare (once per 2 KB), kept out of line (and free of
// the ops object, which must stay in registers) so that the literal loop stays small
__device__ __noinline__ void inflate_flush_chunk(const uint8_t* ring, uint8_t* out, uint32_t first, int lane) {
  __syncwarp();
  const uint4* src = reinterpret_cast<const uint4*>(ring + first % LSL_INF_RING);   // chunks never wrap
  uint8_t* dst = out + first;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    for (int k = lane; k < LSL_INF_CHUNK / 16; k += 32) reinterpret_cast<uint4*>(dst)[k] = src[k];
  } else {
    const uint8_t* sb = reinterpret_cast<const uint8_t*>(src);
    for (int k = lane; k < LSL_INF_CHUNK; k += 32) dst[k] = sb[k];
  }
}

struct InflateOpsWarp {
  int lane;
  uint8_t* ring;        // shared memory, LSL_INF_RING bytes, 16-byte aligned
  uint32_t rp;          // ring index of the next output byte (= pos % LSL_INF_RING, kept incrementally)
  uint32_t next_flush;  // output position at which the next chunk is complete (multiple of LSL_INF_CHUNK)
  __device__ __forceinline__ bool leader() const { return lane == 0; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ void flush_chunk(uint8_t* out) {
    inflate_flush_chunk(ring, out, next_flush - LSL_INF_CHUNK, lane);
    next_flush += LSL_INF_CHUNK;
  }
  __device__ __forceinline__ void put(uint8_t* out, uint32_t pos, uint8_t v) {
    if (lane == 0) ring[rp] = v;
    rp = (rp + 1 == LSL_INF_RING) ? 0 : rp + 1;
    if (pos + 1 == next_flush) flush_chunk(out);
  }
  // Literal runs are 87 % of the decoder's instructions on image data (51 per literal in the general loop: an output-overflow
  // test, a chunk test, a ring wrap and five reconvergence points per symbol). Here the three events (chunk complete, output
  // full, ring end) are folded into ONE countdown computed per run, and a symbol costs a table load, the literal test, the
  // shift, a predicated store and the countdown.
  __device__ __forceinline__ uint32_t literal_run(lslm::BitIn* b, const lslm::HuffTable* h, uint8_t* out, uint32_t pos, uint32_t want) {
    // 32-bit shared-window addresses once per run (the compiler otherwise rebuilds them from generic pointers per symbol)
    const uint32_t fast_s = (uint32_t)__cvta_generic_to_shared(h->fast), ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    for (;;) {
      uint32_t left = min(min(next_flush - pos, want - pos), (uint32_t)LSL_INF_RING - rp);
      if (left == 0) return pos;                    // output full: the general loop decides what the next symbol means
      const uint32_t n0 = left;
      bool stop = false;
      // three symbols per refill test: after bits_fill there are >= 33 valid bits and a first-level code has <= 9
      while (left >= 3) {
        if (b->cnt <= 32) lslm::bits_fill(b);
        uint64_t buf = b->buf;
        uint32_t e0, e1, e2;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e0) : "r"(fast_s + 2u * ((uint32_t)buf & ((1u << LSL_INF_FAST) - 1u))));
        if (e0 - 1u >= 4095u) { stop = true; break; }   // 0: long code; >= 256 << 4: length / end-of-block symbol
        const int l0 = (int)(e0 & 15u);
        buf >>= l0;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e1) : "r"(fast_s + 2u * ((uint32_t)buf & ((1u << LSL_INF_FAST) - 1u))));
        if (lane == 0) asm volatile("st.shared.u8 [%0], %1;" :: "r"(ring_s + rp), "r"(e0 >> 4) : "memory");
        if (e1 - 1u >= 4095u) { b->buf = buf; b->cnt -= l0; rp += 1; left -= 1; stop = true; break; }
        const int l1 = (int)(e1 & 15u);
        buf >>= l1;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e2) : "r"(fast_s + 2u * ((uint32_t)buf & ((1u << LSL_INF_FAST) - 1u))));
        if (lane == 0) asm volatile("st.shared.u8 [%0], %1;" :: "r"(ring_s + rp + 1u), "r"(e1 >> 4) : "memory");
        if (e2 - 1u >= 4095u) { b->buf = buf; b->cnt -= l0 + l1; rp += 2; left -= 2; stop = true; break; }
        const int l2 = (int)(e2 & 15u);
        buf >>= l2;
        if (lane == 0) asm volatile("st.shared.u8 [%0], %1;" :: "r"(ring_s + rp + 2u), "r"(e2 >> 4) : "memory");
        b->buf = buf; b->cnt -= l0 + l1 + l2; rp += 3; left -= 3;
      }
      while (left && !stop) {
        if (b->cnt <= 32) lslm::bits_fill(b);
        uint32_t e;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(fast_s + 2u * ((uint32_t)b->buf & ((1u << LSL_INF_FAST) - 1u))));
        if (e - 1u >= 4095u) { stop = true; break; }
        const int l = (int)(e & 15u);
        b->buf >>= l; b->cnt -= l;
        if (lane == 0) asm volatile("st.shared.u8 [%0], %1;" :: "r"(ring_s + rp), "r"(e >> 4) : "memory");
        ++rp; --left;
      }
      pos += n0 - left;
      if (rp == LSL_INF_RING) rp = 0;
      if (pos == next_flush) flush_chunk(out);
      if (stop) return pos;
    }
  }
  __device__ __forceinline__ void copy(uint8_t* out, uint32_t pos, int dist, int n) {
    __syncwarp();                                   // the bytes the match refers to are visible to every lane
    if (dist <= LSL_INF_NEAR) {
      const uint32_t sbase = rp + LSL_INF_RING - (uint32_t)dist;    // < 2 * LSL_INF_RING
      for (int i = lane; i < n; i += 32) {
        uint32_t s = sbase + (uint32_t)(dist >= n ? i : i % dist);
        if (s >= LSL_INF_RING) s -= LSL_INF_RING;
        uint32_t d = rp + (uint32_t)i;
        if (d >= LSL_INF_RING) d -= LSL_INF_RING;
        ring[d] = ring[s];
      }
    } else {                                        // far match: the source left the ring, read it back from the output (L2)
      const uint8_t* src = out + (pos - (uint32_t)dist);
      for (int i = lane; i < n; i += 32) {          // dist > n here (NEAR > 258): no overlap
        uint32_t d = rp + (uint32_t)i;
        if (d >= LSL_INF_RING) d -= LSL_INF_RING;
        ring[d] = __ldcg(src + i);
      }
    }
    __syncwarp();
    rp += (uint32_t)n;
    if (rp >= LSL_INF_RING) rp -= LSL_INF_RING;
    while (pos + (uint32_t)n >= next_flush) flush_chunk(out);
  }
  __device__ __forceinline__ void stored(uint8_t* out, uint32_t pos, const uint8_t* src, uint32_t n) {
    for (uint32_t done = 0; done < n;) {             // at most one chunk at a time so that the ring never overflows
      const uint32_t m = min(n - done, (uint32_t)LSL_INF_CHUNK);
      __syncwarp();
      for (uint32_t i = lane; i < m; i += 32) ring[(rp + i) % LSL_INF_RING] = src[done + i];
      done += m;
      rp = (rp + m) % LSL_INF_RING;
      while (pos + done >= next_flush) flush_chunk(out);
    }
  }
  __device__ __forceinline__ void finish(uint8_t* out, uint32_t pos) {
    __syncwarp();
    for (uint32_t i = next_flush - LSL_INF_CHUNK + lane; i < pos; i += 32) out[i] = ring[i % LSL_INF_RING];
  }
};

// status[img]: 0 or the negative code of lslm::inflate_zlib
__global__ void __launch_bounds__(32) png_inflate_kernel(const uint8_t* __restrict__ z_all, const size_t* __restrict__ z_off,
                                                         uint8_t* filt_all, size_t filt_img_bytes, int* __restrict__ status) {
  __shared__ lslm::InflateScratch S;
  extern __shared__ __align__(16) uint8_t s_ring[];
  const int img = blockIdx.x;
  InflateOpsWarp ops;
  ops.lane = threadIdx.x; ops.ring = s_ring; ops.rp = 0; ops.next_flush = LSL_INF_CHUNK;
  const int rc = lslm::inflate_zlib(z_all + z_off[img], z_off[img + 1] - z_off[img], filt_all + (size_t)img * filt_img_bytes,
                                    filt_img_bytes, &S, ops);
  if (threadIdx.x == 0) status[img] = rc;
}

__device__ __forceinline__ unsigned paeth(unsigned a, unsigned b, unsigned c) {
  const int p = (int)a + (int)b - (int)c;
  const int pa = abs(p - (int)a), pb = abs(p - (int)b), pc = abs(p - (int)c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// KIND 0: colour / grey 8-bit -> BGR u8 [H][W][3];  KIND 1: grey 16-bit (big endian) -> f32 metres [H][W]
template <int BPP, int KIND>
__global__ void __launch_bounds__(1024) png_unfilter_kernel(const uint8_t* __restrict__ filt_all, size_t filt_img_bytes,
                                                            uint8_t* __restrict__ out_all, size_t out_img_bytes, int W, int H,
                                                            float depth_scale, int* __restrict__ status) {
  extern __shared__ uint32_t s_pub[];   // [2][H]: the pixel each row reconstructed in the previous / current step
  const int y = threadIdx.x;
  const size_t stride = 1 + (size_t)W * BPP;
  const uint8_t* row = filt_all + (size_t)blockIdx.x * filt_img_bytes + (size_t)(y < H ? y : 0) * stride;
  uint8_t* out = out_all + (size_t)blockIdx.x * out_img_bytes;
  const int ft = (y < H) ? row[0] : 0;
  if (ft > 4) status[blockIdx.x] = -9;      // not a PNG filter type (the row is then passed through unfiltered)
  uint32_t a = 0, c = 0;
  const int steps = W + H - 1;
  for (int t = 0; t < steps; ++t) {
    const int x = t - y;
    if (y < H && x >= 0 && x < W) {
      const uint32_t b = (y > 0) ? s_pub[((t - 1) & 1) * H + (y - 1)] : 0u;
      uint32_t r = 0;
#pragma unroll
      for (int j = 0; j < BPP; ++j) {
        const unsigned f = __ldg(row + 1 + (size_t)x * BPP + j);
        const unsigned aj = (a >> (8 * j)) & 255u, bj = (b >> (8 * j)) & 255u, cj = (c >> (8 * j)) & 255u;
        unsigned pred;
        switch (ft) {
          case 1: pred = aj; break;
          case 2: pred = bj; break;
          case 3: pred = (aj + bj) >> 1; break;
          case 4: pred = paeth(aj, bj, cj); break;
          default: pred = 0; break;
        }
        r |= ((f + pred) & 255u) << (8 * j);
      }
      s_pub[(t & 1) * H + y] = r;
      const size_t px = (size_t)y * W + x;
      if (KIND == 0) {
        const uint8_t c0 = (uint8_t)(r & 255u), c1 = (uint8_t)((r >> 8) & 255u), c2 = (uint8_t)((r >> 16) & 255u);
        if (BPP >= 3) { out[px * 3] = c2; out[px * 3 + 1] = c1; out[px * 3 + 2] = c0; }   // file order RGB(A) -> BGR
        else { out[px * 3] = c0; out[px * 3 + 1] = c0; out[px * 3 + 2] = c0; }            // grey -> three equal planes
      } else {
        const unsigned v = ((r & 255u) << 8) | ((r >> 8) & 255u);   // network byte order
        const float d = (float)v;                                   // convertTo(CV_32FC1): exact for 16 bits
        // values < 1e-5 -> quiet NaN (openni_listener.cpp:1238-1241), then MatExpr "/ 5000.0" = float multiply by
        // (float)(1 / 5000.0). x86 keeps the NaN's bits through the multiply (0x7fc00000); the GPU would canonicalise
        // them to 0x7fffffff, so the NaN is stored as the reference leaves it.
        reinterpret_cast<float*>(out)[px] = ((double)d < 1e-5) ? __int_as_float(0x7fc00000) : d * depth_scale;
      }
      c = b; a = r;
    }
    __syncthreads();
  }
}

// Integrity of what was just decoded (cv::imread rejects a PNG whose chunk CRC or zlib check value is wrong): one CTA
// per image. (1) CRC-32 of every IDAT chunk (type + data, PNG §5.3) over the gathered payload pieces, slice-by-4 tables
// built in shared memory, one thread per piece; (2) Adler-32 of the inflated stream (RFC 1950) as 256 contiguous
// segments: per segment A = sum d, B = sum (len - k) d_k, combined in stream order by one thread, against the trailer.
struct PngPiece { uint32_t off_lo, off_hi, len, crc; };   // offset of the piece inside the staging block, stored CRC
__global__ void __launch_bounds__(256) png_check_kernel(const uint8_t* __restrict__ z_all, const PngPiece* __restrict__ pieces,
                                                        const uint32_t* __restrict__ piece_begin, const uint32_t* __restrict__ adler_want,
                                                        const uint8_t* __restrict__ filt_all, size_t filt_img_bytes, int* __restrict__ status) {
  __shared__ uint32_t T[4][256];
  __shared__ uint32_t sA[256];
  __shared__ unsigned long long sB[256];
  const int img = blockIdx.x, tid = threadIdx.x;
  {  // CRC tables (polynomial 0xEDB88320)
    uint32_t c = (uint32_t)tid;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    T[0][tid] = c;
  }
  __syncthreads();
  for (int t = 1; t < 4; ++t) { const uint32_t c = T[t - 1][tid]; T[t][tid] = T[0][c & 255u] ^ (c >> 8); }
  __syncthreads();
  if (status[img] != 0) return;          // not decoded: the earlier error stands (uniform per CTA)
  bool bad_crc = false;
  for (uint32_t pi = piece_begin[img] + tid; pi < piece_begin[img + 1]; pi += blockDim.x) {
    const PngPiece P = pieces[pi];
    const uint8_t* d = z_all + (((size_t)P.off_hi << 32) | P.off_lo);
    uint32_t c = 0xFFFFFFFFu;
    const uint8_t ty[4] = {'I', 'D', 'A', 'T'};
    for (int k = 0; k < 4; ++k) c = T[0][(c ^ ty[k]) & 255u] ^ (c >> 8);
    uint32_t i = 0;
    for (; i < P.len && ((reinterpret_cast<uintptr_t>(d + i)) & 3); ++i) c = T[0][(c ^ d[i]) & 255u] ^ (c >> 8);
    for (; i + 4 <= P.len; i += 4) {
      c ^= __ldg(reinterpret_cast<const uint32_t*>(d + i));
      c = T[3][c & 255u] ^ T[2][(c >> 8) & 255u] ^ T[1][(c >> 16) & 255u] ^ T[0][c >> 24];
    }
    for (; i < P.len; ++i) c = T[0][(c ^ d[i]) & 255u] ^ (c >> 8);
    if ((c ^ 0xFFFFFFFFu) != P.crc) bad_crc = true;
  }
  // Adler-32 segments
  const uint8_t* f = filt_all + (size_t)img * filt_img_bytes;
  const size_t seg = (filt_img_bytes + 255) / 256;
  const size_t b0 = (size_t)tid * seg, b1 = b0 + seg < filt_img_bytes ? b0 + seg : filt_img_bytes;
  uint32_t A = 0; unsigned long long B = 0;
  if (b0 < filt_img_bytes) {
    const unsigned long long L = b1 - b0;
    for (size_t k = b0; k < b1; ++k) { const uint32_t v = f[k]; A += v; B += (L - (k - b0)) * v; }
  }
  sA[tid] = A; sB[tid] = B;
  const int any_bad = __syncthreads_or(bad_crc ? 1 : 0);
  if (tid == 0) {
    unsigned long long s1 = 1, s2 = 0;
    for (int t = 0; t < 256; ++t) {
      const size_t t0 = (size_t)t * seg;
      if (t0 >= filt_img_bytes) break;
      const unsigned long long L = (t0 + seg < filt_img_bytes ? t0 + seg : filt_img_bytes) - t0;
      s2 = (s2 + (L % 65521ull) * s1 + sB[t] % 65521ull) % 65521ull;
      s1 = (s1 + sA[t]) % 65521ull;
    }
    const uint32_t got = (uint32_t)((s2 << 16) | s1);
    if (any_bad) status[img] = -10;
    else if (got != adler_want[img]) status[img] = -11;
  }
}

// ------------------------------------------------------------------------------------------------ host ----
// CRC-32 of the small non-IDAT chunks on the host (the IDAT payload is checked on the device, png_check_kernel)
uint32_t crc32_host(const uint8_t* p, size_t n) {
  static uint32_t T[256];
  static bool init = false;
  if (!init) { for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; T[i] = c; } init = true; }
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) c = T[(c ^ p[i]) & 255u] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

struct PngView {
  int W = 0, H = 0, ch = 0, bits = 0;
  std::vector<std::pair<const uint8_t*, size_t>> idat;
  std::vector<uint32_t> idat_crc;   // stored CRC of each IDAT chunk (verified on the device)
};

bool png_parse(const uint8_t* p, size_t len, PngView* v, std::string* err) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (!p || len < 8 + 25 || std::memcmp(p, sig, 8) != 0) { *err = "not a PNG file"; return false; }
  size_t off = 8;
  bool have_ihdr = false, end = false;
  while (!end && off + 12 <= len) {
    const uint32_t clen = be32(p + off);
    const uint8_t* type = p + off + 4;
    if ((size_t)clen > len - off - 12) { *err = "truncated PNG chunk"; return false; }
    const uint8_t* data = p + off + 8;
    if (!std::memcmp(type, "IHDR", 4)) {
      if (clen != 13) { *err = "bad IHDR"; return false; }
      v->W = (int)be32(data); v->H = (int)be32(data + 4); v->bits = data[8];
      const int ctype = data[9];
      if (data[10] != 0 || data[11] != 0) { *err = "unknown PNG compression / filter method"; return false; }
      if (data[12] != 0) { *err = "interlaced PNG not supported"; return false; }
      if (ctype == 0 && (v->bits == 8 || v->bits == 16)) v->ch = 1;
      else if (ctype == 2 && v->bits == 8) v->ch = 3;
      else if (ctype == 6 && v->bits == 8) v->ch = 4;
      else { *err = "PNG colour type / bit depth not supported (grey 8/16, RGB 8, RGBA 8)"; return false; }
      have_ihdr = true;
    } else if (!std::memcmp(type, "IDAT", 4)) {
      v->idat.emplace_back(data, (size_t)clen);
      v->idat_crc.push_back(be32(data + clen));
    } else if (!std::memcmp(type, "IEND", 4)) end = true;
    if (std::memcmp(type, "IDAT", 4) != 0 && crc32_host(type, 4 + (size_t)clen) != be32(data + clen)) { *err = "PNG chunk CRC mismatch"; return false; }
    off += 12 + (size_t)clen;
  }
  if (!have_ihdr || v->idat.empty()) { *err = "PNG without IHDR / IDAT"; return false; }
  return true;
}

struct TumScratch {   // staging per list kind (0 colour, 1 depth) so that the two lists are in flight together
  uint8_t* h_z[2] = {nullptr, nullptr}; size_t h_cap[2] = {0, 0};     // pinned: [offsets (n + 1) size_t | status n int | compressed streams]
  uint8_t* d_z[2] = {nullptr, nullptr}; size_t dz_cap[2] = {0, 0};    // the same block on the device
  uint8_t* d_filt[2] = {nullptr, nullptr}; size_t d_cap[2] = {0, 0};  // inflated (still filtered) scan lines
  uint8_t* h_meta[2] = {nullptr, nullptr}; uint8_t* d_meta[2] = {nullptr, nullptr}; size_t meta_cap[2] = {0, 0};   // IDAT piece table + Adler-32 trailers
  uint8_t* d_bgr = nullptr; size_t bgr_cap = 0;  // lsl_extract_tum_batch only
  float* d_depth = nullptr; size_t depth_cap = 0;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};
std::mutex g_mu;
std::map<lsl_ctx*, TumScratch> g_scratch;

template <typename T>
bool grow_dev(T** p, size_t* cap, size_t need) {
  if (*cap >= need) return true;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  if (cudaMalloc((void**)p, need) != cudaSuccess) return false;
  *cap = need;
  return true;
}

template <int BPP, int KIND>
void launch_unfilter(cudaStream_t st, int n, const uint8_t* d_filt, size_t filt_img, uint8_t* out, size_t out_img, int W, int H,
                     int* d_status) {
  const int threads = (H + 31) & ~31;
  png_unfilter_kernel<BPP, KIND><<<n, threads, 2 * (size_t)H * sizeof(uint32_t), st>>>(d_filt, filt_img, out, out_img, W, H,
                                                                                    (float)(1.0 / 5000.0), d_status);
}

const char* inflate_error(int rc) {
  switch (rc) {
    case -1: return "bad zlib header";
    case -2: return "bad DEFLATE block type";
    case -3: return "bad stored block";
    case -4: return "bad Huffman code lengths";
    case -5: return "bad symbol or distance";
    case -6: return "corrupt data stream (more data than W x H needs)";
    case -7: return "truncated data stream";
    case -8: return "short data stream (fewer bytes than W x H needs)";
    case -9: return "bad PNG filter type";
    case -10: return "IDAT chunk CRC mismatch";
    case -11: return "zlib Adler-32 mismatch";
    default: return "corrupt data stream";
  }
}

// one list of n same-sized PNGs -> device; kind 0 colour (any supported 8-bit type), 1 depth (grey 16)
// Enqueues gather -> H2D -> inflate -> unfilter -> status D2H on `st`; the caller synchronises and calls check_list.
int decode_list(lsl_ctx* ctx, TumScratch& S, cudaStream_t st, int n, const uint8_t* const* png, const size_t* len, int W, int H, int kind,
                void* d_out) {
  std::vector<PngView> views((size_t)n);
  int ch = 0;
  size_t zbytes = 0;
  for (int i = 0; i < n; ++i) {
    std::string err;
    if (!png_parse(png[i], len[i], &views[(size_t)i], &err)) { ctx->err = "image " + std::to_string(i) + ": " + err; return LSL_ERR_ARG; }
    const PngView& v = views[(size_t)i];
    if (v.W != W || v.H != H) { ctx->err = "image " + std::to_string(i) + ": size differs from W x H"; return LSL_ERR_ARG; }
    if (kind == 1 ? !(v.ch == 1 && v.bits == 16) : v.bits != 8) {
      ctx->err = "image " + std::to_string(i) + (kind == 1 ? ": depth PNG must be 16-bit grey" : ": colour PNG must be 8 bits per sample");
      return LSL_ERR_ARG;
    }
    if (i == 0) ch = v.ch;
    else if (v.ch != ch) { ctx->err = "images of one batch must share the PNG colour type"; return LSL_ERR_ARG; }
    for (const auto& c : v.idat) zbytes += c.second;
    zbytes = (zbytes + 15) & ~(size_t)15;
  }
  const int bpp = kind == 1 ? 2 : ch;
  const size_t img_bytes = (size_t)H * (1 + (size_t)W * bpp);
  const size_t head = (((size_t)(n + 1) * sizeof(size_t) + (size_t)n * sizeof(int)) + 15) & ~(size_t)15;
  const size_t total = head + zbytes;
  if (S.h_cap[kind] < total) {
    if (S.h_z[kind]) cudaFreeHost(S.h_z[kind]);
    S.h_z[kind] = nullptr; S.h_cap[kind] = 0;
    LSL_CUDA(cudaHostAlloc((void**)&S.h_z[kind], total, cudaHostAllocDefault));
    S.h_cap[kind] = total;
  }
  uint8_t* const hz = S.h_z[kind];
  if (!grow_dev(&S.d_z[kind], &S.dz_cap[kind], total) || !grow_dev(&S.d_filt[kind], &S.d_cap[kind], img_bytes * (size_t)n)) {
    ctx->err = "cudaMalloc of the PNG staging buffers failed";
    return LSL_ERR_CUDA;
  }
  // gather: the IDAT payloads of one image form one zlib stream (PNG §10.1), streams 16-byte aligned
  uint8_t* const dz = S.d_z[kind];
  uint8_t* const dfilt = S.d_filt[kind];
  size_t* h_off = reinterpret_cast<size_t*>(hz);
  int* h_status = reinterpret_cast<int*>(hz + (size_t)(n + 1) * sizeof(size_t));
  size_t npieces = 0;
  for (int i = 0; i < n; ++i) npieces += views[(size_t)i].idat.size();
  const size_t meta_bytes = sizeof(PngPiece) * npieces + sizeof(uint32_t) * (size_t)(n + 1) + sizeof(uint32_t) * (size_t)n;
  if (S.meta_cap[kind] < meta_bytes) {
    if (S.h_meta[kind]) cudaFreeHost(S.h_meta[kind]);
    if (S.d_meta[kind]) cudaFree(S.d_meta[kind]);
    S.h_meta[kind] = nullptr; S.d_meta[kind] = nullptr; S.meta_cap[kind] = 0;
    LSL_CUDA(cudaHostAlloc((void**)&S.h_meta[kind], meta_bytes * 2, cudaHostAllocDefault));
    LSL_CUDA(cudaMalloc((void**)&S.d_meta[kind], meta_bytes * 2));
    S.meta_cap[kind] = meta_bytes * 2;
  }
  PngPiece* h_pieces = reinterpret_cast<PngPiece*>(S.h_meta[kind]);
  uint32_t* h_pbegin = reinterpret_cast<uint32_t*>(S.h_meta[kind] + sizeof(PngPiece) * npieces);
  uint32_t* h_adler = h_pbegin + (n + 1);
  size_t pi = 0;
  size_t off = head;
  for (int i = 0; i < n; ++i) {                    // layout first (serial, a few words per chunk) ...
    h_off[i] = off;
    h_pbegin[i] = (uint32_t)pi;
    const PngView& vw = views[(size_t)i];
    for (size_t c = 0; c < vw.idat.size(); ++c) {
      PngPiece& P = h_pieces[pi++];
      P.off_lo = (uint32_t)(off & 0xffffffffu); P.off_hi = (uint32_t)(off >> 32); P.len = (uint32_t)vw.idat[c].second; P.crc = vw.idat_crc[c];
      off += vw.idat[c].second;
    }
    off += (16 - (off & 15)) & 15;
    h_status[i] = 0;
  }
  {                                                // ... then the payload copies on a few host threads (0.4 - 0.5 MB per image:
    const int T = std::max(1, std::min(8, std::min(n / 16, (int)std::thread::hardware_concurrency())));   // 1 GB per 1184 frames)
    auto work = [&](int t) {
      for (int i = t; i < n; i += T) {
        const PngView& vw = views[(size_t)i];
        size_t o = h_off[i];
        for (size_t c = 0; c < vw.idat.size(); ++c) { std::memcpy(hz + o, vw.idat[c].first, vw.idat[c].second); o += vw.idat[c].second; }
        h_adler[i] = (o - h_off[i] >= 4) ? be32(hz + o - 4) : 0u;     // zlib trailer: the last four payload bytes (RFC 1950)
        std::memset(hz + o, 0, (16 - (o & 15)) & 15);                 // < 16 zero bytes behind the Adler-32 trailer: never decoded,
      }                                                               // the final block ends the stream before them
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
  }
  h_off[n] = off;
  h_pbegin[n] = (uint32_t)pi;
  LSL_CUDA(cudaMemcpyAsync(S.d_meta[kind], S.h_meta[kind], meta_bytes, cudaMemcpyHostToDevice, st));
  LSL_CUDA(cudaMemcpyAsync(dz, hz, total, cudaMemcpyHostToDevice, st));
  ctx->stats.h2d_bytes += (int64_t)total;
  const size_t* d_off = reinterpret_cast<const size_t*>(dz);
  int* d_status = reinterpret_cast<int*>(dz + (size_t)(n + 1) * sizeof(size_t));
  const int k_inf = kind == 1 ? LSL_K_INFLATE_D : LSL_K_INFLATE, k_unf = kind == 1 ? LSL_K_PNG_D : LSL_K_PNG;
  cudaEventRecord(ctx->kev[k_inf][0], st);
  {  // function attributes are per device: set once for every device this process decodes on
    static std::mutex attr_mu;
    static bool attr_set[64] = {false};
    std::lock_guard<std::mutex> lk(attr_mu);
    const int dev = ctx->device & 63;
    if (!attr_set[dev]) { cudaFuncSetAttribute(png_inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LSL_INF_RING); attr_set[dev] = true; }
  }
  png_inflate_kernel<<<n, 32, LSL_INF_RING, st>>>(dz, d_off, dfilt, img_bytes, d_status);
  cudaEventRecord(ctx->kev[k_inf][1], st);
  {
    const PngPiece* d_pieces = reinterpret_cast<const PngPiece*>(S.d_meta[kind]);
    const uint32_t* d_pbegin = reinterpret_cast<const uint32_t*>(S.d_meta[kind] + sizeof(PngPiece) * npieces);
    png_check_kernel<<<n, 256, 0, st>>>(dz, d_pieces, d_pbegin, d_pbegin + (n + 1), dfilt, img_bytes, d_status);
    ctx->stats.kernel_launches += 1;
  }
  cudaEventRecord(ctx->kev[k_unf][0], st);
  uint8_t* o = (uint8_t*)d_out;
  const size_t out_img = kind == 1 ? (size_t)W * H * sizeof(float) : (size_t)W * H * 3;
  if (kind == 1) launch_unfilter<2, 1>(st, n, dfilt, img_bytes, o, out_img, W, H, d_status);
  else if (bpp == 3) launch_unfilter<3, 0>(st, n, dfilt, img_bytes, o, out_img, W, H, d_status);
  else if (bpp == 4) launch_unfilter<4, 0>(st, n, dfilt, img_bytes, o, out_img, W, H, d_status);
  else launch_unfilter<1, 0>(st, n, dfilt, img_bytes, o, out_img, W, H, d_status);
  cudaEventRecord(ctx->kev[k_unf][1], st);
  ctx->kran[k_inf] = ctx->kran[k_unf] = true;
  ctx->stats.kernel_launches += 2;
  LSL_CUDA(cudaGetLastError());
  LSL_CUDA(cudaMemcpyAsync(h_status, d_status, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
  ctx->stats.d2h_bytes += (int64_t)n * (int64_t)sizeof(int);
  return LSL_OK;
}

// after the stream(s) have been synchronised: kernel times and the per-image status words of one list
int check_list(lsl_ctx* ctx, TumScratch& S, int n, int kind) {
  const int k_inf = kind == 1 ? LSL_K_INFLATE_D : LSL_K_INFLATE, k_unf = kind == 1 ? LSL_K_PNG_D : LSL_K_PNG;
  cudaEventElapsedTime(&ctx->kms[k_inf], ctx->kev[k_inf][0], ctx->kev[k_inf][1]);
  cudaEventElapsedTime(&ctx->kms[k_unf], ctx->kev[k_unf][0], ctx->kev[k_unf][1]);
  const int* h_status = reinterpret_cast<const int*>(S.h_z[kind] + (size_t)(n + 1) * sizeof(size_t));
  for (int i = 0; i < n; ++i)
    if (h_status[i] != 0) {
      ctx->err = std::string(kind == 1 ? "depth" : "colour") + " image " + std::to_string(i) + ": " + inflate_error(h_status[i]);
      return LSL_ERR_ARG;
    }
  return LSL_OK;
}

}  // namespace

extern "C" int lsl_tum_read_syncidx(const char* dirname, lsl_tum_entry* dst, int cap, int* n) {
  if (!dirname || !n) return LSL_ERR_ARG;
  *n = 0;
  const std::string path = std::string(dirname) + "/syncidx.txt";
  FILE* f = std::fopen(path.c_str(), "r");
  if (!f) return LSL_OK;   // like the reference: a missing list is an empty list (openni_listener.cpp:1206)
  char tok[4][256];
  int k = 0, groups = 0, rc = LSL_OK;
  char buf[256];
  while (std::fscanf(f, "%255s", buf) == 1) {     // `fsync >> tmp`: whitespace-separated tokens
    std::memcpy(tok[k], buf, sizeof buf);
    if (++k == 4) {
      k = 0;
      if (dst && groups < cap) {
        lsl_tum_entry& e = dst[groups];
        e.ts_rgb = std::atof(tok[0]); e.ts_depth = std::atof(tok[2]);
        std::memcpy(e.rgb, tok[1], 256); std::memcpy(e.depth, tok[3], 256);
      } else rc = LSL_ERR_CAPACITY;
      ++groups;
    }
  }
  std::fclose(f);
  *n = groups;
  return rc;
}

extern "C" int lsl_png_info(const uint8_t* png, size_t len, int* W, int* H, int* channels, int* bit_depth) {
  PngView v; std::string err;
  if (!png_parse(png, len, &v, &err)) return LSL_ERR_ARG;
  if (W) *W = v.W;
  if (H) *H = v.H;
  if (channels) *channels = v.ch;
  if (bit_depth) *bit_depth = v.bits;
  return LSL_OK;
}

extern "C" int lsl_tum_decode_batch(lsl_ctx* ctx, int n, const uint8_t* const* rgb_png, const size_t* rgb_len,
                                    const uint8_t* const* depth_png, const size_t* depth_len, int W, int H,
                                    uint8_t* d_bgr, float* d_depth) {
  if (!ctx || n < 0 || W <= 0 || H <= 0) return LSL_ERR_ARG;
  if (H > 1024) { ctx->err = "PNG decode: H > 1024 rows"; return LSL_ERR_CAPACITY; }
  if ((rgb_png && (!rgb_len || !d_bgr)) || (depth_png && (!depth_len || !d_depth))) return LSL_ERR_ARG;
  if (n == 0) return LSL_OK;
  LSL_ENTER(ctx);
  std::lock_guard<std::mutex> lk(g_mu);
  TumScratch& S = g_scratch[ctx];
  int rc = LSL_OK;
  // colour list on the context's stream, depth list on its copy stream: the host gathers the second list while the
  // first one is being inflated, and the two inflate kernels share the SMs (5 streams per SM fit, a list brings 4)
  const bool both = rgb_png && depth_png;
  cudaStream_t st_d = ctx->stream;
  if (both) {
    if (!S.ev_fork) { LSL_CUDA(cudaEventCreateWithFlags(&S.ev_fork, cudaEventDisableTiming)); LSL_CUDA(cudaEventCreateWithFlags(&S.ev_join, cudaEventDisableTiming)); }
    st_d = ctx->copy_stream;
    LSL_CUDA(cudaEventRecord(S.ev_fork, ctx->stream));
    LSL_CUDA(cudaStreamWaitEvent(st_d, S.ev_fork, 0));
  }
  if (rgb_png) rc = decode_list(ctx, S, ctx->stream, n, rgb_png, rgb_len, W, H, 0, d_bgr);
  if (rc == LSL_OK && depth_png) rc = decode_list(ctx, S, st_d, n, depth_png, depth_len, W, H, 1, d_depth);
  if (both) {
    cudaEventRecord(S.ev_join, st_d);
    cudaStreamWaitEvent(ctx->stream, S.ev_join, 0);
  }
  // the pinned staging blocks are reused by the next call, and the status words decide the return value
  const cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (rc != LSL_OK) return rc;
  if (e != cudaSuccess) { ctx->err = std::string("PNG decode: ") + cudaGetErrorString(e); return LSL_ERR_CUDA; }
  if (rgb_png && (rc = check_list(ctx, S, n, 0)) != LSL_OK) return rc;
  if (depth_png && (rc = check_list(ctx, S, n, 1)) != LSL_OK) return rc;
  return LSL_OK;
}

extern "C" int lsl_extract_tum_batch(lsl_ctx* ctx, int n, const uint8_t* const* rgb_png, const size_t* rgb_len,
                                     const uint8_t* const* depth_png, const size_t* depth_len, int W, int H, const double K[9],
                                     double asynch_dt_s, const uint32_t* rand_seeds, lsl_frame** out) {
  if (!ctx || n <= 0 || !rgb_png || !depth_png || !out) return LSL_ERR_ARG;
  LSL_ENTER(ctx);
  uint8_t* d_bgr; float* d_depth;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    TumScratch& S = g_scratch[ctx];
    if (!grow_dev(&S.d_bgr, &S.bgr_cap, (size_t)n * W * H * 3) || !grow_dev(&S.d_depth, &S.depth_cap, (size_t)n * W * H * sizeof(float))) {
      ctx->err = "cudaMalloc of the decoded image buffers failed";
      return LSL_ERR_CUDA;
    }
    d_bgr = S.d_bgr; d_depth = S.d_depth;
  }
  int rc = lsl_tum_decode_batch(ctx, n, rgb_png, rgb_len, depth_png, depth_len, W, H, d_bgr, d_depth);
  if (rc != LSL_OK) return rc;
  static const double Ktum[9] = {525, 0, 319.5, 0, 525, 239.5, 0, 0, 1};   // openni_listener.cpp:1256-1260
  return lsl_extract_batch_dev(ctx, n, d_bgr, 3, d_depth, W, H, K ? K : Ktum, asynch_dt_s, rand_seeds, out);
}

extern "C" void lsl_tum_release(lsl_ctx* ctx) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_scratch.find(ctx);
  if (it == g_scratch.end()) return;
  TumScratch& S = it->second;
  if (ctx) cudaSetDevice(ctx->device);
  for (int k = 0; k < 2; ++k) {
    if (S.h_z[k]) cudaFreeHost(S.h_z[k]);
    if (S.d_z[k]) cudaFree(S.d_z[k]);
    if (S.d_filt[k]) cudaFree(S.d_filt[k]);
    if (S.h_meta[k]) cudaFreeHost(S.h_meta[k]);
    if (S.d_meta[k]) cudaFree(S.d_meta[k]);
  }
  if (S.ev_fork) cudaEventDestroy(S.ev_fork);
  if (S.ev_join) cudaEventDestroy(S.ev_join);
  if (S.d_bgr) cudaFree(S.d_bgr);
  if (S.d_depth) cudaFree(S.d_depth);
  g_scratch.erase(it);
}
