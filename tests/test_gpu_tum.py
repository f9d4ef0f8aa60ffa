"""SURVEY.md §8f row 4 on the device: png_unfilter_kernel (wavefront unfiltering + BGR / depth conversion) against the
oracle's PNG restatement bit for bit — every filter type, every supported pixel type, ragged and full sizes — and the
whole ingest (PNG files -> lsl_extract_tum_batch) against extraction from the oracle-decoded arrays."""
import io

import numpy as np
import pytest
import torch

from oracle import oracle_png as OP
from test_tum import _img

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(api):
    c = api.Context(max_batch=8, max_w=640, max_h=480)
    yield c
    from lineslam_b200 import tum
    tum.release(c)
    c.close()


def _decode(ctx, rgb, dep, W, H):
    from lineslam_b200 import tum
    n = len(rgb if rgb is not None else dep)
    d_bgr = torch.zeros((n, H, W, 3), dtype=torch.uint8, device="cuda") if rgb is not None else None
    d_dep = torch.zeros((n, H, W), dtype=torch.float32, device="cuda") if dep is not None else None
    tum.decode_batch(ctx, rgb, dep, W, H, d_bgr.data_ptr() if rgb is not None else 0, d_dep.data_ptr() if dep is not None else 0)
    torch.cuda.synchronize()
    return (d_bgr.cpu().numpy() if rgb is not None else None), (d_dep.cpu().numpy() if dep is not None else None)


@pytest.mark.parametrize("shape", [(29, 37), (1, 5), (64, 1), (480, 640)])
def test_unfilter_all_filters_all_types(ctx, shape):
    H, W = shape
    rng = np.random.default_rng(H * 1000 + W)
    filters = [0, 1, 2, 3, 4, None] if H * W < 10000 else [None, 4]
    for kind in ["rgb", "rgba", "grey"]:
        files = [OP.png_encode(_img(rng, H, W, kind), filters=f, seed=11 + i, idat_split=1000 if i == 1 else 0)
                 for i, f in enumerate(filters)]
        got, _ = _decode(ctx, files, None, W, H)
        for i, f in enumerate(files):
            assert np.array_equal(got[i], OP.imread_bgr(f)), (kind, filters[i])
    files = [OP.png_encode(_img(rng, H, W, "depth"), filters=f, seed=3 + i) for i, f in enumerate(filters)]
    _, got = _decode(ctx, None, files, W, H)
    for i, f in enumerate(files):
        want = OP.depth_metres(f)
        assert np.array_equal(np.isnan(got[i]), np.isnan(want))
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32)), filters[i]      # bit-exact incl. the NaN pattern


def test_tall_image_960_rows(ctx):
    rng = np.random.default_rng(9)
    a = _img(rng, 960, 128, "rgb")
    f = OP.png_encode(a, filters=None, seed=1)
    got, _ = _decode(ctx, [f], None, 128, 960)
    assert np.array_equal(got[0], a[:, :, ::-1])
    d = _img(rng, 960, 128, "depth")
    _, gd = _decode(ctx, None, [OP.png_encode(d, filters=4)], 128, 960)
    ok = d != 0
    assert np.array_equal(gd[0][ok], d[ok].astype(np.float32) * np.float32(1.0 / 5000.0)) and np.isnan(gd[0][~ok]).all()


def test_rejects(api, ctx):
    from lineslam_b200 import tum
    from PIL import Image
    rng = np.random.default_rng(2)
    good = OP.png_encode(_img(rng, 24, 32, "rgb"))
    out = torch.zeros((2, 24, 32, 3), dtype=torch.uint8, device="cuda")
    with pytest.raises(api.LslError, match="size differs"):
        tum.decode_batch(ctx, [good, OP.png_encode(_img(rng, 24, 31, "rgb"))], None, 32, 24, out.data_ptr(), 0)
    with pytest.raises(api.LslError, match="16-bit grey"):
        tum.decode_batch(ctx, None, [OP.png_encode(_img(rng, 24, 32, "grey"))], 32, 24, 0, out.data_ptr())
    buf = io.BytesIO()
    Image.fromarray(_img(rng, 24, 32, "rgb")).save(buf, format="PNG", interlace=True)
    if b"IHDR" in buf.getvalue() and buf.getvalue()[28] == 1:          # Pillow wrote an Adam7 file
        with pytest.raises(api.LslError, match="interlaced"):
            tum.decode_batch(ctx, [buf.getvalue()], None, 32, 24, out.data_ptr(), 0)
    with pytest.raises(api.LslError, match="corrupt|short|truncated"):
        tum.decode_batch(ctx, [good[:-40]], None, 32, 24, out.data_ptr(), 0)
    with pytest.raises(api.LslError):
        tum.decode_batch(ctx, [OP.png_encode(_img(rng, 1100, 8, "rgb"))], None, 8, 1100, out.data_ptr(), 0)   # H > 1024
    # damaged DEFLATE data inside intact chunks: the device decoder reports it (or, rarely, decodes other pixels)
    i0 = good.index(b"IDAT") + 4
    n_err = 0
    for k in range(40):
        bad = bytearray(good)
        bad[i0 + 2 + int(rng.integers(0, 200))] ^= 1 << int(rng.integers(0, 8))
        try:
            tum.decode_batch(ctx, [good, bytes(bad)], None, 32, 24, out.data_ptr(), 0)
        except api.LslError as e:
            assert "image 1" in str(e)
            n_err += 1
    assert n_err == 40          # every flipped bit inside an IDAT chunk breaks its CRC-32 (checked on the device)
    # the zlib check value: a wrong Adler-32 trailer inside a chunk whose CRC is right -> rejected by the Adler check
    import struct
    import zlib
    ln = struct.unpack(">I", good[i0 - 8:i0 - 4])[0]
    bad = bytearray(good)
    bad[i0 + ln - 1] ^= 0x5a                                            # last byte of the zlib stream = Adler-32 low byte
    bad[i0 + ln:i0 + ln + 4] = struct.pack(">I", zlib.crc32(bytes(bad[i0 - 4:i0 + ln])) & 0xffffffff)
    with pytest.raises(api.LslError, match="Adler"):
        tum.decode_batch(ctx, [good, bytes(bad)], None, 32, 24, out.data_ptr(), 0)
    # a damaged ancillary / IEND chunk CRC is caught by the host walk
    bad = bytearray(good); bad[-1] ^= 1
    with pytest.raises(api.LslError, match="CRC"):
        tum.decode_batch(ctx, [bytes(bad)], None, 32, 24, out.data_ptr(), 0)
    tum.decode_batch(ctx, [good, good], None, 32, 24, out.data_ptr(), 0)     # the context is still usable
    torch.cuda.synchronize()
    assert np.array_equal(out[1].cpu().numpy(), OP.imread_bgr(good))


def test_raw_directory_to_frames(api, ctx, stream4, tmp_path):
    """loadRawData end to end: a TUM-shaped directory (syncidx.txt, rgb/*.png in RGB file order, depth/*.png 16-bit
    x5000) -> frames; equals extraction from the arrays the oracle decodes from the same files."""
    from PIL import Image
    from lineslam_b200 import tum
    imgs, deps, poses, K = stream4
    d = tmp_path / "rgbd_dataset_synth"
    (d / "rgb").mkdir(parents=True); (d / "depth").mkdir()
    lines = []
    for i in range(4):
        ts = 1305031102.175304 + i / 30.0
        Image.fromarray(np.ascontiguousarray(imgs[i][:, :, ::-1])).save(d / "rgb" / f"{ts:.6f}.png")     # memory order BGR -> file RGB
        z = np.nan_to_num(deps[i].astype(np.float64), nan=0.0)
        Image.fromarray(np.rint(z * 5000.0).astype(np.uint16)).save(d / "depth" / f"{ts:.6f}.png")
        lines.append(f"{ts:.6f} rgb/{ts:.6f}.png {ts - 0.01:.6f} depth/{ts:.6f}.png")
    (d / "syncidx.txt").write_text("\n".join(lines) + "\n")
    batches = list(tum.load_raw_data(ctx, str(d), batch=3, K=K))
    assert [len(b[0]) for b in batches] == [3, 1]
    frames = [f for b in batches for f in b[1]]
    stamps = [t for b in batches for t in b[0]]
    assert stamps == [float(l.split()[0]) for l in lines]
    ent = OP.read_syncidx(str(d))
    bgr = np.stack([OP.imread_bgr(open(d / e[1], "rb").read()) for e in ent])
    dep = np.stack([OP.depth_metres(open(d / e[3], "rb").read()) for e in ent])
    assert np.array_equal(bgr, imgs)                                  # lossless colour round trip
    ref = ctx.extract_batch(bgr, dep, K, seeds=[1, 2, 3, 1])          # load_raw_data seeds: 1,2,3 then batch 2 starts at 4
    ref2 = ctx.extract_batch(bgr[3:], dep[3:], K, seeds=[4])
    for i in range(3):
        assert frames[i].num_lines > 50
        assert frames[i].lines().tobytes() == ref[i].lines().tobytes()
    assert frames[3].lines().tobytes() == ref2[0].lines().tobytes()
    kt = ctx.kernel_times()
    assert kt.get("png_unfilter_kernel", 0) > 0 and kt.get("png_inflate_kernel", 0) > 0


@pytest.mark.parametrize("W,period", [(463, 16), (464, 16), (640, 24), (511, 16)])
def test_far_matches_beyond_the_shared_memory_ring(ctx, W, period):
    """Matches that reach further back than the 8 KB shared-memory ring (LSL_INF_NEAR = 7424 bytes) read their source from the
    flushed output: incompressible rows repeated with a period of 7424 (W = 463, the last distance served by the ring), 7440,
    15 384 and 8 192 bytes, unfiltered so that zlib finds the repeats at exactly that distance."""
    rng = np.random.default_rng(W + period)
    tile = rng.integers(0, 256, (period, W), dtype=np.uint8)
    H = period * 5 + 3
    a = np.tile(tile, (6, 1))[:H]
    f = OP.png_encode(a, filters=0, level=9)
    assert len(f) < 0.45 * a.size                       # the repeats were found: the stream is mostly far matches
    # the decoder contract (shared header compiled for the host) and the device agree with the pixels
    got, _ = _decode(ctx, [f, f], None, W, H)
    want = np.repeat(a[:, :, None], 3, axis=2)
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)
