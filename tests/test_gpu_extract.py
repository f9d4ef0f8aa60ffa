"""GPU parity of Node::detect3DLines through the C ABI against the oracle.

Tier-E (bit-exact): 2D end points, line equation, polarity r, MSLD descriptor, 3D-RANSAC inlier index
sets, segment-of-line map. The MLE stage (levmar restated with the same operation order on both sides)
is also compared bit for bit; the tolerance the north star allows for it is 1e-4 m on end points.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EXACT = ["p", "q", "lineEq2d", "r", "des", "lid", "haveDepth"]
MLE = ["A", "B", "covA", "covB", "DU_A", "DU_B", "Wsqrt_A", "Wsqrt_B"]


def _compare(got, ref, dbg_got, dbg_ref, tag):
    assert len(got) == len(ref), (tag, len(got), len(ref))
    assert np.array_equal(dbg_got["seg_of_line"], dbg_ref["seg_of_line"]), tag
    assert np.array_equal(dbg_got["n_inl"], dbg_ref["n_inl"]), tag
    for i in range(len(ref)):
        k = dbg_ref["n_inl"][i]
        assert np.array_equal(dbg_got["inl_idx"][i, :k], dbg_ref["inl_idx"][i, :k]), (tag, i)
    for name in EXACT:
        # equal_nan: a line with a single valid MSLD sample has std = 0 -> 0 * inf = NaN descriptor entries in the
        # reference arithmetic (utils.cpp:1596-1606); both sides must then carry the NaN in the same places
        assert np.array_equal(got[name], ref[name], equal_nan=(name == "des")), (tag, name)
    assert np.array_equal(dbg_got["lm_iters"], dbg_ref["lm_iters"]), tag
    for name in MLE:
        if not np.array_equal(got[name], ref[name]):
            # Tier-T bound of the north star: 1e-4 m on the 3D end points
            assert np.allclose(got["A"], ref["A"], atol=1e-4, rtol=0) and np.allclose(got["B"], ref["B"], atol=1e-4, rtol=0)
            d = np.abs(got[name] - ref[name]).max()
            pytest.fail(f"{tag}: field {name} not bit-exact (max abs diff {d:g}); end points within 1e-4 m")


def test_detect3DLines_vga_batch(api, oracle, stream4):
    imgs, deps, poses, K = stream4
    ctx = api.Context(max_batch=4, max_w=640, max_h=480, debug=True)
    seeds = [1, 2, 3, 4]
    frames = ctx.extract_batch(imgs, deps, K, seeds=seeds)
    for i in range(4):
        ref, dref = oracle.detect3DLines(imgs[i], deps[i], K, seed=seeds[i], debug=True)
        _compare(frames[i].lines(), ref, frames[i].debug(), dref, f"frame{i}")
    ctx.close()


def test_detect3DLines_small_and_single(api, oracle, small_frames):
    imgs, deps, poses, K = small_frames
    ctx = api.Context(max_batch=1, max_w=320, max_h=240, debug=True)
    for i in range(2):
        node = api.Node(ctx, imgs[i], deps[i], K, node_id=i, seed=7 + i)
        ref, dref = oracle.detect3DLines(imgs[i], deps[i], K, seed=7 + i, debug=True)
        _compare(node.lines, ref, node.frame.debug(), dref, f"small{i}")
    ctx.close()


def test_detect3DLines_holes_and_no_depth(api, oracle, stream4):
    """Depth with 40 % NaN holes (ragged per-line point sets) and an all-invalid depth map (no 3D line)."""
    imgs, deps, poses, K = stream4
    rng = np.random.default_rng(11)
    d1 = deps[0].copy()
    d1[rng.random(d1.shape) < 0.4] = np.nan
    d2 = np.zeros_like(deps[0])
    ctx = api.Context(max_batch=2, max_w=640, max_h=480, debug=True)
    frames = ctx.extract_batch(np.stack([imgs[0], imgs[0]]), np.stack([d1, d2]), K, seeds=[5, 6])
    ref, dref = oracle.detect3DLines(imgs[0], d1, K, seed=5, debug=True)
    _compare(frames[0].lines(), ref, frames[0].debug(), dref, "holes")
    assert frames[1].num_lines == 0
    assert len(oracle.detect3DLines(imgs[0], d2, K, seed=6)) == 0
    ctx.close()


def test_asynch_dt_and_launch_params(api, oracle, stream4):
    """MODEL_ASYNCH time offset enters the depth sigma; launch-file overrides (lsd_angle_thres 40)."""
    imgs, deps, poses, K = stream4
    p = api.default_params()
    p.lsd_ang_th = 40.0
    ctx = api.Context(params=p, max_batch=1, max_w=640, max_h=480, debug=True)
    fr = ctx.extract_batch(imgs[:1], deps[:1], K, seeds=[3], dt=0.02)[0]
    ref, dref = oracle.detect3DLines(imgs[0], deps[0], K, seed=3, params=p, dt=0.02, debug=True)
    _compare(fr.lines(), ref, fr.debug(), dref, "dt")
    ctx.close()


def test_detect3DLines_on_reference_tum_frame(api, oracle):
    """detect3DLines on the reference's real TUM frame (tests/golden/ref_tum_frame.png) with a synthetic slanted depth
    plane + holes: every record field against the oracle."""
    import os
    import cv2
    tum = cv2.imread(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_tum_frame.png"), cv2.IMREAD_COLOR)
    H, W = tum.shape[:2]
    yy, xx = np.mgrid[0:H, 0:W]
    rng = np.random.default_rng(21)
    dep = (1.2 + 0.002 * xx + 0.0015 * yy).astype(np.float32)
    dep = (np.round(dep * 5000) / 5000).astype(np.float32)
    dep[rng.random(dep.shape) < 0.05] = np.nan
    K = np.array([[525., 0, 319.5], [0, 525., 239.5], [0, 0, 1]])
    ctx = api.Context(max_batch=1, max_w=W, max_h=H, debug=True)
    fr = ctx.extract_batch(tum[None], dep[None], K, seeds=[9])[0]
    ref, dref = oracle.detect3DLines(tum, dep, K, seed=9, debug=True)
    assert len(ref) > 150
    _compare(fr.lines(), ref, fr.debug(), dref, "tum")
    ctx.close()


def test_border_segments_are_clipped_not_skipped(api, oracle):
    """FrameLine::getGradient iterates cv::LineIterator, whose constructor clips end points that round outside the
    image (src/line/lineslam.cpp:527-537): such a line has a finite polarity r and a real MSLD descriptor."""
    from lineslam_b200 import synth
    W, H = 320, 240
    K = synth.camera_K(W, H)
    yy, xx = np.mgrid[0:H, 0:W]
    dep = (1.5 + 0.003 * xx + 0.002 * yy).astype(np.float32)
    ctx = api.Context(max_batch=2, max_w=W, max_h=H, debug=True)
    imgs = np.stack([synth.border_bands(88, W, H), synth.border_bands(307, W, H)])
    frames = ctx.extract_batch(imgs, np.stack([dep, dep]), K, seeds=[1, 2])
    outside = 0
    for i in range(2):
        ref, dref = oracle.detect3DLines(imgs[i], dep, K, seed=i + 1, debug=True)
        got = frames[i].lines()
        _compare(got, ref, frames[i].debug(), dref, f"border{i}")
        for l in got:
            ends = np.rint(np.array([l["p"], l["q"]]))
            if (ends[:, 0] < 0).any() or (ends[:, 0] >= W).any() or (ends[:, 1] < 0).any() or (ends[:, 1] >= H).any():
                outside += 1
                assert np.all(np.isfinite(l["r"])) and abs(np.linalg.norm(l["des"]) - 1.0) < 1e-9
    assert outside >= 2
    ctx.close()


def test_u16_depth_entry_equals_f32_entry(api, oracle, stream4):
    """lsl_extract_batch_u16: the 16-bit depth of the sensor / TUM PNG converted on the device (0 -> NaN, x (float)(1/5000),
    src/openni_listener.cpp:1233-1244) gives the records of the f32 entry on the host-converted planes, for pageable and
    for pinned contiguous buffers (the two upload paths), and moves 2 instead of 4 depth bytes per pixel."""
    import torch
    imgs, deps, poses, K = stream4
    raw = np.nan_to_num(deps * 5000.0, nan=0.0).round().astype(np.uint16)
    raw[0, 100:140, 200:300] = 0
    conv = raw.astype(np.float32)
    conv = np.where(conv < 1e-5, np.float32(np.nan), conv * np.float32(1.0 / 5000.0)).astype(np.float32)
    ctx = api.Context(max_batch=4, max_w=640, max_h=480, debug=True)
    ref = ctx.extract_batch(imgs, conv, K, seeds=[1, 2, 3, 4])
    ref_lines = [f.lines().copy() for f in ref]
    orc = oracle.detect3DLines(imgs[0], conv[0], K, seed=1)
    assert ref_lines[0].tobytes() == orc.tobytes()
    s0 = ctx.stats().h2d_bytes
    got = ctx.extract_batch(imgs, raw, K, seeds=[1, 2, 3, 4])
    assert ctx.stats().h2d_bytes - s0 == 4 * 640 * 480 * (3 + 2)
    for g, r in zip(got, ref_lines):
        assert g.lines().tobytes() == r.tobytes()
    pi, pd = torch.from_numpy(imgs.copy()).pin_memory(), torch.from_numpy(raw.view(np.int16).copy()).pin_memory()
    got2 = ctx.extract_batch(pi.numpy(), pd.numpy().view(np.uint16), K, seeds=[1, 2, 3, 4])
    for g, r in zip(got2, ref_lines):
        assert g.lines().tobytes() == r.tobytes()
    ctx.close()
