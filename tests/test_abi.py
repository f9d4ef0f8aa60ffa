"""The C-ABI library loads without a GPU, exports every symbol include/lsl.h declares, has the record layouts
the header states, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not h.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(lsl_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol(api):
    L = api.lib()
    names = _declared()
    assert len(names) >= 40 and "lsl_graph_add_frame" in names
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/*.h but not exported"


def test_record_layouts():
    from lineslam_b200.records import LINE_DTYPE, MATCH_DTYPE, POSE_DTYPE, Params
    assert LINE_DTYPE.itemsize == 1040 and MATCH_DTYPE.itemsize == 12 and POSE_DTYPE.itemsize == 128
    assert C.sizeof(Params) == 26 * 8 + 12 * 4


def test_defaults_match_reference_parameter_server(api):
    p = api.default_params()   # src/parameter_server.cpp:160-199
    assert (p.lsd_ang_th, p.lsd_density_th, p.line_2d_len_thres) == (22.5, 0.7, 10.0)
    assert (p.ransac_iters_line_motion, p.min_feature_matches, p.adjacent_linematch_window) == (500, 20, 3)
    assert (p.ransac_iters_extract_line, p.line_sample_max_num, p.line3d_mle_iter_num) == (100, 100, 100)
    assert (p.depth_stdev_coeff_c1, p.depth_stdev_coeff_c2, p.depth_stdev_coeff_c3) == (0.00273, 0.00074, -0.00058)


def test_no_device_means_error_not_fallback(api):
    import torch
    if torch.cuda.is_available():
        return
    h = C.c_void_p()
    rc = api.lib().lsl_ctx_create(C.byref(h), None, 0, 1, 640, 480)
    assert rc == -4 and b"no CPU fallback" in api.lib().lsl_strerror(rc)   # LSL_ERR_NO_DEVICE
    import pytest
    with pytest.raises(api.LslError):
        api.Context()


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under lineslam_b200/ may include, import, link or load it."""
    pat = re.compile(r'#include\s+"[^"]*oracle|from\s+oracle|import\s+oracle|liboracle|pyoracle|-loracle|oracle/_ref')
    for d, _, files in os.walk(os.path.join(ROOT, "lineslam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")) or f == "Makefile":
                s = open(os.path.join(d, f), errors="ignore").read()
                assert not pat.search(s), os.path.join(d, f)
