// TEST INFRASTRUCTURE — host build of the product's DEFLATE decoder (lineslam_b200/csrc/shared/lsl_inflate.h) so that
// tests/test_tum.py can check that exact code against zlib on the CPU. Never linked or loaded by the product.
#include "../lineslam_b200/csrc/shared/lsl_inflate.h"

extern "C" int lsl_inflate_host_check(const uint8_t* in, size_t len, uint8_t* out, size_t want) {
  static thread_local lslm::InflateScratch S;
  lslm::InflateOpsSerial ops;
  return lslm::inflate_zlib(in, len, out, want, &S, ops);
}
