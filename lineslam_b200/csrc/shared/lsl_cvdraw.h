// lsl_cvdraw.h — integer geometry of cv::LineIterator's constructor shared by the CPU oracle and the sm_100a kernels
// (FrameLine::getGradient, src/line/lineslam.cpp:527-537, iterates cv::LineIterator(*xGradient, p, q, 8)).
// OpenCV is not under /root/reference: this restates cv::clipLine (OpenCV 2.4 modules/core/src/drawing.cpp, the
// Cohen-Sutherland variant with truncating integer division); tests/test_oracle_cv2.py compares it with
// cv2.clipLine of the OpenCV in this image on random and border cases.
#pragma once
#include "lsl_math.h"

namespace lslm {

// cv::clipLine(Size(W, H), pt1, pt2): clips the segment to [0, W-1] x [0, H-1]; false when it lies outside.
LSL_HD bool clip_line(int W, int H, int* px1, int* py1, int* px2, int* py2) {
  long long x1 = *px1, y1 = *py1, x2 = *px2, y2 = *py2;
  const long long right = W - 1, bottom = H - 1;
  if (W <= 0 || H <= 0) return false;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (a - y1) * (x2 - x1) / (y2 - y1);
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (a - y2) * (x2 - x1) / (y2 - y1);
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (a - x1) * (y2 - y1) / (x2 - x1);
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (a - x2) * (y2 - y1) / (x2 - x1);
        x2 = a;
        c2 = 0;
      }
    }
    *px1 = (int)x1; *py1 = (int)y1; *px2 = (int)x2; *py2 = (int)y2;
  }
  return (c1 | c2) == 0;
}

// cv::LineIterator constructor's end-point handling: end points inside the image are used as they are, otherwise the
// segment is clipped; returns false when nothing is left (count = 0: the caller's sums stay zero).
LSL_HD bool line_iter_endpoints(int W, int H, int* x1, int* y1, int* x2, int* y2) {
  if ((unsigned)*x1 >= (unsigned)W || (unsigned)*x2 >= (unsigned)W || (unsigned)*y1 >= (unsigned)H || (unsigned)*y2 >= (unsigned)H)
    return clip_line(W, H, x1, y1, x2, y2);
  return true;
}

}  // namespace lslm
