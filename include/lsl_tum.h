/* lsl_tum.h — C ABI of the data format on the input side of the hot path ("next"): the TUM RGB-D
 * raw directory OpenNIListener::loadRawData reads (src/openni_listener.cpp:1194-1319). The trajectory writer on the
 * output side is lsl_graph_write_poses (include/lsl_graph.h).
 *
 *   syncidx.txt          whitespace-separated tokens in groups of four: ts_rgb rgb_file ts_depth depth_file
 *                        (openni_listener.cpp:1201-1217; a trailing incomplete group is dropped)
 *   rgb PNG              cv::imread(name, 1): 8-bit, 3 channels in BGR order whatever the file holds (:1233)
 *   depth PNG            16-bit grey; convertTo(CV_32FC1), values < 1e-5 -> NaN, then "/ 5000.0" (:1234-1244)
 *
 * Split of the PNG decode: the host only walks the PNG container (signature, IHDR, IDAT chunk boundaries) and gathers
 * the compressed payloads into one pinned block; the zlib / DEFLATE stream is decoded on the device, one warp per
 * image (png_inflate_kernel), followed by scan-line unfiltering (None/Sub/Up/Average/Paeth) as a row wavefront,
 * RGB->BGR / grey replication and big-endian 16-bit -> float metres with the NaN rule (png_unfilter_kernel), which
 * writes the exact buffers lsl_extract_batch_dev consumes. What crosses PCIe is the COMPRESSED data (about 0.4 MB +
 * 0.45 MB per VGA frame instead of 0.92 MB + 1.23 MB of decoded planes). Not verified: chunk CRCs and the Adler-32
 * trailer (the decoder checks the block structure, every code and distance, and the exact output length).
 *
 * "/ 5000.0" on a CV_32F cv::Mat is OpenCV 2.4's MatExpr scale: convertTo(CV_32F, alpha = 1/5000.0) whose 32f->32f
 * kernel multiplies in float by (float)alpha. OpenCV is not in this image: that reading is UNVERIFIED (the alternative,
 * a float division by 5000, differs in the last bit for some depths); parity for it is pinned to the oracle only.
 * Supported PNGs: non-interlaced, colour type 0 (grey 8/16 bit), 2 (RGB 8 bit), 6 (RGBA 8 bit, alpha dropped like
 * imread(…,1)); anything else returns LSL_ERR_ARG with the reason in lsl_last_error. H <= 1024.
 */
#ifndef LSL_TUM_H_
#define LSL_TUM_H_
#include <stddef.h>
#include "lsl.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lsl_tum_entry {
  double ts_rgb, ts_depth;          /* atof of columns 0 and 2 (openni_listener.cpp:1246-1247) */
  char rgb[256], depth[256];        /* columns 1 and 3, relative to the directory */
} lsl_tum_entry;

/* Parses <dirname>/syncidx.txt. *n = number of complete groups (also when cap is too small: LSL_ERR_CAPACITY). */
int lsl_tum_read_syncidx(const char* dirname, lsl_tum_entry* dst, int cap, int* n);

/* IHDR of a PNG file image in memory: channels = samples per pixel (1, 3, 4), bit_depth 8 or 16. */
int lsl_png_info(const uint8_t* png, size_t len, int* W, int* H, int* channels, int* bit_depth);

/* n RGB PNGs and/or n depth PNGs (file images in host memory; either list may be NULL) -> device buffers
 * d_bgr [n][H][W][3] u8 and d_depth [n][H][W] f32 (metres, NaN = no reading), on the context's stream.
 * Every image must be W x H. */
int lsl_tum_decode_batch(lsl_ctx* ctx, int n, const uint8_t* const* rgb_png, const size_t* rgb_len,
                         const uint8_t* const* depth_png, const size_t* depth_len, int W, int H,
                         uint8_t* d_bgr, float* d_depth);

/* loadRawData + Node::Node for n frames: decode as above into library-owned device buffers, then
 * lsl_extract_batch_dev (channels = 3, BGR). K defaults to the TUM intrinsics of openni_listener.cpp:1256-1260
 * (525, 525, 319.5, 239.5) when NULL. */
int lsl_extract_tum_batch(lsl_ctx* ctx, int n, const uint8_t* const* rgb_png, const size_t* rgb_len,
                          const uint8_t* const* depth_png, const size_t* depth_len, int W, int H, const double K[9],
                          double asynch_dt_s, const uint32_t* rand_seeds, lsl_frame** out);

/* Releases the decode staging buffers the library keeps per context (call before lsl_ctx_destroy). */
void lsl_tum_release(lsl_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
