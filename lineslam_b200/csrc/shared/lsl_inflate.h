// lsl_inflate.h — zlib / DEFLATE (RFC 1950, RFC 1951) decoder for one stream by one thread, host and device.
// The device runs it with one warp per PNG data stream (k_tum.cu: png_inflate_kernel, lane 0 decodes, the whole warp
// copies matches); the host build exists only so that tests can check the same code against zlib without a GPU
// (oracle/inflate_check.cpp — test infrastructure; the product never inflates on the CPU through this header).
//
// Decoding is canonical-Huffman: per code length the number of codes and the symbols in code order (RFC 1951 §3.2.2),
// with a 2^LSL_INF_FAST-entry first-level table in front (code, reversed into LSB-first bit order -> symbol | length)
// so that the common short codes cost one lookup; longer codes fall back to the length-by-length walk.
// Length / distance bases are computed (§3.2.5 tables follow base = 3 + ((4 + (s & 3)) << e), 1 + ((2 + (d & 1)) << e)).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include "lsl_math.h"

#if defined(__CUDACC__)
#define LSL_HDM __host__ __device__ __forceinline__   // member functions (LSL_HD is `static inline` on the host)
#else
#define LSL_HDM inline
#endif

namespace lslm {

#define LSL_INF_FAST 9

struct HuffTable {
  uint16_t count[16];                 // codes per length
  uint16_t symbol[288];               // symbols ordered by (length, value)
  uint16_t fast[1 << LSL_INF_FAST];   // (symbol << 4) | length for codes of length <= LSL_INF_FAST, else 0
};

struct InflateScratch {
  HuffTable lit, dist;
  uint8_t lengths[320];
  int status;
};

struct BitIn {
  const uint8_t* in;
  size_t len, pos;      // pos: next input byte that has not entered buf yet
  uint64_t buf;
  int cnt;
  int overrun;          // bytes asked for past the end (zeros are supplied; a few are normal look-ahead)
  uint32_t nextw;       // word at in + pos, loaded one refill early so that its latency hides behind decoding
  bool have_next;
};

LSL_HD uint32_t load_le32(const uint8_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(reinterpret_cast<const uint32_t*>(p));
#else
  uint32_t v; memcpy(&v, p, 4); return v;
#endif
}

// at least 33 valid bits afterwards: aligned 32-bit words while two whole words remain, single bytes otherwise
LSL_HD void bits_fill(BitIn* b) {
  while (b->cnt <= 32) {
    if ((((uintptr_t)(b->in + b->pos)) & 3) == 0 && b->pos + 8 <= b->len) {
      if (!b->have_next) { b->nextw = load_le32(b->in + b->pos); b->have_next = true; }
      const uint32_t w = b->nextw;
      b->nextw = load_le32(b->in + b->pos + 4);
      b->buf |= (uint64_t)w << b->cnt;
      b->cnt += 32; b->pos += 4;
    } else {
      b->have_next = false;
      uint64_t v = 0;
      if (b->pos < b->len) v = b->in[b->pos]; else b->overrun++;
      b->pos++;
      b->buf |= v << b->cnt;
      b->cnt += 8;
    }
  }
}
LSL_HD uint32_t bits_get(BitIn* b, int n) {   // n <= 32, after bits_fill
  const uint32_t v = (uint32_t)(b->buf & ((1ull << n) - 1ull));
  b->buf >>= n; b->cnt -= n;
  return v;
}

// builds count / symbol / fast from code lengths; returns < 0 for an over-subscribed set
LSL_HD int huff_build(HuffTable* h, const uint8_t* lengths, int n) {
  uint16_t offs[16];
  for (int l = 0; l < 16; ++l) h->count[l] = 0;
  for (int s = 0; s < n; ++s) h->count[lengths[s]]++;
  int left = 1;
  for (int l = 1; l < 16; ++l) { left <<= 1; left -= h->count[l]; if (left < 0) return -1; }
  offs[1] = 0;
  for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + h->count[l]);
  for (int s = 0; s < n; ++s)
    if (lengths[s] != 0) h->symbol[offs[lengths[s]]++] = (uint16_t)s;
  for (int i = 0; i < (1 << LSL_INF_FAST); ++i) h->fast[i] = 0;
  // canonical codes in increasing (length, symbol) order; entry index = the code's bits in stream (LSB-first) order
  int code = 0, idx = 0;
  for (int l = 1; l <= LSL_INF_FAST; ++l) {
    for (int k = 0; k < h->count[l]; ++k, ++code, ++idx) {
      int rev = 0;
      for (int q = 0; q < l; ++q) rev |= ((code >> q) & 1) << (l - 1 - q);
      const uint16_t e = (uint16_t)((h->symbol[idx] << 4) | l);
      for (int f = rev; f < (1 << LSL_INF_FAST); f += (1 << l)) h->fast[f] = e;
    }
    code <<= 1;
  }
  return left;   // > 0: incomplete code (legal for a single distance code)
}

// one symbol; < 0 when the bits match no code. Needs >= 15 valid bits in b->buf (bits_fill).
LSL_HD int huff_decode(BitIn* b, const HuffTable* h) {
  const uint16_t e = h->fast[b->buf & ((1u << LSL_INF_FAST) - 1u)];
  if (e) { const int l = e & 15; b->buf >>= l; b->cnt -= l; return e >> 4; }
  int code = 0, first = 0, index = 0;
  uint64_t bits = b->buf;
  for (int l = 1; l <= 15; ++l) {
    code |= (int)(bits & 1); bits >>= 1;
    const int count = h->count[l];
    if (code - count < first) { b->buf >>= l; b->cnt -= l; return h->symbol[index + (code - first)]; }
    index += count; first += count; first <<= 1; code <<= 1;
  }
  return -1;
}

// Error codes (negative): -1 header, -2 block type, -3 stored block, -4 code lengths, -5 bad symbol / distance,
// -6 output overflow, -7 input exhausted, -8 output short.
// Ops supplies the three ways bytes reach the output: put (one literal), copy (LZ77 match, may overlap: byte i of the
// match is out[pos - dist + i % dist]) and stored (raw bytes of a stored block). The host passes plain loops; on the
// device every lane of a warp runs this function convergently on the same stream (identical control flow, tables in
// shared memory), lane 0 writes the literals and the whole warp shares the copies.
struct InflateOpsSerial {
  LSL_HDM bool leader() const { return true; }   // the one thread that builds the shared tables
  LSL_HDM void sync() const {}                   // all threads of the group have passed this point, writes visible
  LSL_HDM void finish(uint8_t*, uint32_t) const {} // everything handed to put / copy / stored is in `out` afterwards
  LSL_HDM void put(uint8_t* out, uint32_t pos, uint8_t v) const { out[pos] = v; }
  // optional fast path: decodes a run of literals whose codes hit the first-level table and returns the new position; it
  // stops (without consuming it) at the first symbol that is not such a literal or when `want` is reached. The serial
  // variant leaves everything to the general loop.
  LSL_HDM uint32_t literal_run(BitIn*, const HuffTable*, uint8_t*, uint32_t pos, uint32_t) const { return pos; }
  LSL_HDM void copy(uint8_t* out, uint32_t pos, int dist, int n) const { for (int i = 0; i < n; ++i) out[pos + i] = out[pos + i - dist]; }
  LSL_HDM void stored(uint8_t* out, uint32_t pos, const uint8_t* src, uint32_t n) const { for (uint32_t i = 0; i < n; ++i) out[pos + i] = src[i]; }
};

template <typename Ops>
LSL_HD int inflate_zlib(const uint8_t* in, size_t len, uint8_t* out, size_t want_bytes, InflateScratch* S, Ops& ops) {
  if (len < 6 || want_bytes > 0xfffffff0u) return -1;
  const uint32_t want = (uint32_t)want_bytes;        // positions are 32-bit: cheaper on the device
  if ((in[0] & 15) != 8 || (in[0] >> 4) > 7 || (in[1] & 32) || ((in[0] << 8) | in[1]) % 31 != 0) return -1;
  BitIn b;
  b.in = in; b.len = len; b.pos = 2; b.buf = 0; b.cnt = 0; b.overrun = 0; b.nextw = 0; b.have_next = false;
  uint32_t pos = 0;
  int last = 0;
  const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  while (!last) {
    ops.sync();                                         // nobody still decodes with the previous block's tables
    bits_fill(&b);
    last = (int)bits_get(&b, 1);
    const int type = (int)bits_get(&b, 2);
    if (type == 0) {                                    // stored
      bits_get(&b, b.cnt & 7);                          // to the byte boundary
      bits_fill(&b);
      const uint32_t n = bits_get(&b, 16), nn = bits_get(&b, 16);
      if ((n ^ 0xffffu) != nn) return -3;
      size_t src = b.pos - (size_t)(b.cnt >> 3);        // bytes still in the bit buffer belong to the block
      if (src + n > len) return -7;
      if (n > want - pos) return -6;
      ops.stored(out, pos, in + src, n);
      pos += n;
      b.pos = src + n; b.buf = 0; b.cnt = 0; b.have_next = false;
      continue;
    }
    if (type == 3) return -2;
    if (type == 1) {                                    // fixed codes (§3.2.6)
      if (ops.leader()) {
        for (int s = 0; s < 144; ++s) S->lengths[s] = 8;
        for (int s = 144; s < 256; ++s) S->lengths[s] = 9;
        for (int s = 256; s < 280; ++s) S->lengths[s] = 7;
        for (int s = 280; s < 288; ++s) S->lengths[s] = 8;
        huff_build(&S->lit, S->lengths, 288);
        for (int s = 0; s < 30; ++s) S->lengths[s] = 5;
        huff_build(&S->dist, S->lengths, 30);
      }
      ops.sync();
    } else {                                            // dynamic codes (§3.2.7)
      const int nlen = (int)bits_get(&b, 5) + 257, ndist = (int)bits_get(&b, 5) + 1, ncode = (int)bits_get(&b, 4) + 4;
      if (nlen > 286 || ndist > 30) return -4;
      // every thread reads the 3-bit lengths (its bit reader must advance); the leader builds the table from them
      uint8_t cl[19];
      for (int i = 0; i < 19; ++i) cl[i] = 0;
      for (int i = 0; i < ncode; ++i) { bits_fill(&b); cl[order[i]] = (uint8_t)bits_get(&b, 3); }
      if (ops.leader()) S->status = huff_build(&S->lit, cl, 19);
      ops.sync();
      if (S->status != 0) return -4;                                 // the code-length code must be complete
      // every thread walks the code lengths (its bit reader must advance; the previous length lives in a register),
      // only the leader stores them
      int idx = 0, prev = 0, len256 = 0;
      const bool lead = ops.leader();
      while (idx < nlen + ndist) {
        bits_fill(&b);
        const int sym = huff_decode(&b, &S->lit);
        if (sym < 0) return -4;
        int rep = 1, val = sym;
        if (sym == 16) { if (idx == 0) return -4; val = prev; rep = 3 + (int)bits_get(&b, 2); }
        else if (sym == 17) { val = 0; rep = 3 + (int)bits_get(&b, 3); }
        else if (sym == 18) { val = 0; rep = 11 + (int)bits_get(&b, 7); }
        if (idx + rep > nlen + ndist) return -4;
        if (idx <= 256 && 256 < idx + rep) len256 = val;
        if (lead) for (int k = 0; k < rep; ++k) S->lengths[idx + k] = (uint8_t)val;
        idx += rep;
        prev = val;
      }
      if (len256 == 0) return -4;                                    // no end-of-block code
      // the distance lengths sit behind the literal/length ones: build dist first (lit reuses the scratch table)
      ops.sync();
      if (ops.leader()) {
        int err = huff_build(&S->dist, S->lengths + nlen, ndist);
        // RFC 1951 §3.2.7: an incomplete distance code is legal when it has one code — or none at all (an all-literal
        // block, as libdeflate / zopfli emit); a distance symbol that is then requested matches no code (-5 below)
        int bad = (err < 0 || (err > 0 && ndist - S->dist.count[0] > 1));
        err = huff_build(&S->lit, S->lengths, nlen);
        bad |= (err < 0 || (err > 0 && nlen - S->lit.count[0] != 1));
        S->status = bad;
      }
      ops.sync();
      if (S->status) return -4;
    }
    for (;;) {                                          // compressed data of the block
      pos = ops.literal_run(&b, &S->lit, out, pos, want);
      bits_fill(&b);
      int sym = huff_decode(&b, &S->lit);
      if (sym < 0) return -5;
      if (sym < 256) {
        if (pos >= want) return -6;
        ops.put(out, pos++, (uint8_t)sym);
        continue;
      }
      if (sym == 256) break;
      sym -= 257;
      if (sym >= 29) return -5;
      int mlen;
      if (sym < 8) mlen = 3 + sym;
      else if (sym == 28) mlen = 258;
      else { const int e = (sym >> 2) - 1; mlen = 3 + ((4 + (sym & 3)) << e) + (int)bits_get(&b, e); }
      bits_fill(&b);
      const int ds = huff_decode(&b, &S->dist);
      if (ds < 0 || ds >= 30) return -5;
      int dist;
      if (ds < 4) dist = 1 + ds;
      else { const int e = (ds >> 1) - 1; dist = 1 + ((2 + (ds & 1)) << e) + (int)bits_get(&b, e); }
      if ((uint32_t)dist > pos) return -5;
      if ((uint32_t)mlen > want - pos) return -6;
      ops.copy(out, pos, dist, mlen);
      pos += (uint32_t)mlen;
    }
    if (b.overrun > 8) return -7;
  }
  if (b.overrun > 8) return -7;
  ops.finish(out, pos);
  return pos == want ? 0 : -8;
}

}  // namespace lslm
