// Default parameter values: src/parameter_server.cpp:160-199 (the ParameterServer defaults, not the
// launch-file overrides), SystemParameters::init (src/line/lineslam.cpp:577-640) and the LSD
// constants of lsd_scale() (external/lsd/lsd.cpp:2070-2091). SURVEY.md Appendix B.
#pragma once
#include "../../../include/lsl.h"
static inline void lsl_params_default_impl(lsl_params* p) {
  p->lsd_scale = 0.8; p->lsd_sigma_scale = 0.6; p->lsd_quant = 2.0; p->lsd_ang_th = 22.5; p->lsd_eps = 0.0;
  p->lsd_density_th = 0.7; p->lsd_max_grad = 255.0; p->lsd_n_bins = 1024;
  p->line_2d_len_thres = 10.0; p->msld_sample_interval = 1.0; p->line_3d_len_thres_m = 0.02;
  p->collin_pts_ratio = 0.6; p->line_sample_interval = 1.0;
  p->line_sample_max_num = 100; p->line_sample_min_num = 10; p->line3d_mle_iter_num = 100;
  p->pt2line_mahdist_extractline = 1.5; p->ransac_iters_extract_line = 100; p->num_cells_lineseg_range = 10;
  p->ratio_support_pts_on_line = 0.7; p->stdev_sample_pt_imgline = 3.0;
  p->depth_stdev_coeff_c1 = 0.00273; p->depth_stdev_coeff_c2 = 0.00074; p->depth_stdev_coeff_c3 = -0.00058;
  p->depth_scaling = 1.0;
  p->ransac_iters_line_motion = 500; p->adjacent_linematch_window = 3; p->line_match_number_weight = 1;
  p->min_feature_matches = 20; p->min_matches_loopclose = 20;
  p->max_mah_dist_for_inliers = 3.0; p->g2o_line_error_weight = 1.0; p->g2o_BA_kernel_delta = 10.0;
  p->g2o_BA_use_kernel = 1; p->pt2line3d_dist_relmotion = 0.05; p->line3d_angle_relmotion = 10.0;
  p->sigma_depth = 0.01; p->nn_distance_ratio = 0.5;
}
