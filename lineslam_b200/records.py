"""POD layouts of include/lsl.h as ctypes / numpy types (shared by the product binding and the tests)."""
import ctypes as C
import numpy as np


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "lsd_scale lsd_sigma_scale lsd_quant lsd_ang_th lsd_eps lsd_density_th lsd_max_grad "
        "line_2d_len_thres msld_sample_interval line_3d_len_thres_m collin_pts_ratio line_sample_interval "
        "pt2line_mahdist_extractline ratio_support_pts_on_line stdev_sample_pt_imgline "
        "depth_stdev_coeff_c1 depth_stdev_coeff_c2 depth_stdev_coeff_c3 depth_scaling "
        "max_mah_dist_for_inliers g2o_line_error_weight g2o_BA_kernel_delta "
        "pt2line3d_dist_relmotion line3d_angle_relmotion sigma_depth nn_distance_ratio").split()] + [(n, C.c_int32) for n in (
        "lsd_n_bins line_sample_max_num line_sample_min_num line3d_mle_iter_num "
        "ransac_iters_extract_line num_cells_lineseg_range "
        "ransac_iters_line_motion adjacent_linematch_window line_match_number_weight "
        "min_feature_matches min_matches_loopclose g2o_BA_use_kernel").split()]


LINE_DTYPE = np.dtype([
    ("p", "<f8", 2), ("q", "<f8", 2), ("lineEq2d", "<f8", 3), ("r", "<f8", 2), ("des", "<f8", 72),
    ("A", "<f8", 3), ("B", "<f8", 3), ("covA", "<f8", 9), ("covB", "<f8", 9), ("DU_A", "<f8", 9), ("DU_B", "<f8", 9),
    ("Wsqrt_A", "<f8", 3), ("Wsqrt_B", "<f8", 3), ("lid", "<i4"), ("haveDepth", "<i4")])
assert LINE_DTYPE.itemsize == 1040

MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("distance", "<f4")])
assert MATCH_DTYPE.itemsize == 12

POSE_DTYPE = np.dtype([
    ("id_train", "<i4"), ("id_query", "<i4"), ("found", "<i4"), ("n_line_matches", "<i4"),
    ("n_ransac_inliers", "<i4"), ("n_inliers", "<i4"), ("rmse", "<f4"), ("tf", "<f4", 16), ("best_iter", "<i4"),
    ("pad", "<i4", 8)])
assert POSE_DTYPE.itemsize == 128


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                "kernel_launches frames segments lines3d pairs matches h2d_bytes d2h_bytes".split()]


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None
