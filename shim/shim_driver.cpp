// shim_driver.cpp — exercises the drop-in adapter the way the reference does (Node::detect3DLines on two frames,
// Node::lineMatching, getTransform_PtsLines_ransac) and dumps the results for tests/test_gpu_shim.py.
// usage: shim_driver in.bin out.bin     in.bin = int32 W, H, n; double K[9]; then n x (gray u8 W*H, depth f32 W*H)
#include <cstdio>
#include <cstdlib>
#include "stubs/shim_stubs.h"
SystemParametersStub sysPara;
extern "C" void lsl_shim_shutdown();

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  int32_t hdr[3];
  double K[9];
  if (fread(hdr, 4, 3, f) != 3 || fread(K, 8, 9, f) != 9) return 4;
  const int W = hdr[0], H = hdr[1], n = hdr[2];
  sysPara.min_feature_matches = 10;   // launch/lineslam.launch
  std::vector<Node*> nodes;
  cv::Mat Km(3, 3, CV_64F);
  for (int i = 0; i < 9; ++i) Km.at<double>(i / 3, i % 3) = K[i];
  for (int i = 0; i < n; ++i) {
    cv::Mat gray(H, W, CV_8U), depth(H, W, CV_32F);
    if (fread(gray.data, 1, (size_t)W * H, f) != (size_t)W * H || fread(depth.data, 4, (size_t)W * H, f) != (size_t)W * H) return 5;
    Node* nd = new Node();
    nd->id_ = i;
    try { nd->detect3DLines(gray, depth, sysPara.line_2d_len_thres, Km, 0.6, sysPara.line_3d_len_thres_m, 1.0, "LSD"); }
    catch (const std::exception& e) { fprintf(stderr, "shim_driver: %s\n", e.what()); return 6; }
    nodes.push_back(nd);
  }
  fclose(f);
  FILE* o = fopen(argv[2], "wb");
  if (!o) return 7;
  for (Node* nd : nodes) {   // per node: count, then p q r A B DU_A per line (2+2+2+3+3+9 doubles) + des (72)
    int32_t c = (int32_t)nd->lines.size();
    fwrite(&c, 4, 1, o);
    for (const FrameLine& L : nd->lines) {
      double v[21] = {L.p.x, L.p.y, L.q.x, L.q.y, L.r.x, L.r.y, L.line3d.A.x, L.line3d.A.y, L.line3d.A.z, L.line3d.B.x, L.line3d.B.y, L.line3d.B.z};
      for (int k = 0; k < 9; ++k) v[12 + k] = L.line3d.rndA.DU[k];
      fwrite(v, 8, 21, o);
      fwrite(L.des.data, 8, 72, o);
    }
  }
  // matchNodePair's line half: newer = nodes[1] (query), older = nodes[0] (train)
  std::vector<cv::DMatch> ms, none, pt_inl, ln_inl;
  unsigned int nm = nodes[1]->lineMatching(nodes[0], true, &ms);
  Eigen::Matrix4f tf;
  float rmse = 0;
  bool found = getTransform_PtsLines_ransac(nodes[0], nodes[1], none, ms, pt_inl, ln_inl, tf, rmse);
  int32_t c = (int32_t)nm;
  fwrite(&c, 4, 1, o);
  for (const cv::DMatch& m : ms) { int32_t q[2] = {m.queryIdx, m.trainIdx}; fwrite(q, 4, 2, o); fwrite(&m.distance, 4, 1, o); }
  int32_t fi[2] = {found ? 1 : 0, (int32_t)ln_inl.size()};
  fwrite(fi, 4, 2, o);
  fwrite(&rmse, 4, 1, o);
  float rowmajor[16];
  for (int r = 0; r < 4; ++r) for (int cc = 0; cc < 4; ++cc) rowmajor[r * 4 + cc] = tf(r, cc);
  fwrite(rowmajor, 4, 16, o);
  for (const cv::DMatch& m : ln_inl) { int32_t q[2] = {m.queryIdx, m.trainIdx}; fwrite(q, 4, 2, o); }
  fclose(o);
  for (Node* nd : nodes) delete nd;
  lsl_shim_shutdown();
  return 0;
}
