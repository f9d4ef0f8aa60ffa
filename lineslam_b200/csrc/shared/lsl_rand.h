// lsl_rand.h — glibc rand() (random_r.c TYPE_3: r[i] = r[i-31] + r[i-3], srandom_r seeding with
// 310 discarded outputs) restated for host and device, so that the explicit per-call seed of the
// C ABI replays the stream a single-threaded reference run would consume after srand(seed)
// (src/main.cpp:168, src/line/utils.h:49-60).
#pragma once
#include <stdint.h>
#include "lsl_math.h"

namespace lslm {

struct GRand {
  int32_t r[31];
  int32_t f, b;
};

LSL_HD int grand_next(GRand* g) {
  uint32_t v = (uint32_t)g->r[g->f] + (uint32_t)g->r[g->b];
  g->r[g->f] = (int32_t)v;
  if (++g->f >= 31) g->f = 0;
  if (++g->b >= 31) g->b = 0;
  return (int)(v >> 1);
}

LSL_HD void grand_seed(GRand* g, uint32_t s) {
  if (s == 0) s = 1;
  g->r[0] = (int32_t)s;
  for (int i = 1; i < 31; ++i) {
    long long hi = g->r[i - 1] / 127773, lo = g->r[i - 1] % 127773;
    long long word = 16807 * lo - 2836 * hi;
    if (word < 0) word += 2147483647;
    g->r[i] = (int32_t)word;
  }
  g->f = 3; g->b = 0;
  for (int i = 0; i < 310; ++i) grand_next(g);
}

}  // namespace lslm
