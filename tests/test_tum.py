"""SURVEY.md §8f row 4, CPU side: the PNG restatement (oracle/oracle_png.py) is pinned against an independent decoder
(Pillow) on every supported pixel type and all five scan-line filters; syncidx.txt parsing of the C ABI equals the
oracle's; lsl_png_info reads the header. No device call here."""
import io
import os

import numpy as np
import pytest

from oracle import oracle_png as OP


def _img(rng, H, W, kind):
    yy, xx = np.mgrid[0:H, 0:W]
    base = (xx * 3 + yy * 2) % 256
    if kind == "rgb":
        a = np.stack([base, (base * 2 + 17) % 256, (xx * yy) % 256], -1)
        a = np.where(rng.random((H, W, 1)) < 0.1, rng.integers(0, 256, (H, W, 3)), a)
        return a.astype(np.uint8)
    if kind == "rgba":
        return np.concatenate([_img(rng, H, W, "rgb"), rng.integers(0, 256, (H, W, 1)).astype(np.uint8)], -1)
    if kind == "grey":
        return np.where(rng.random((H, W)) < 0.1, rng.integers(0, 256, (H, W)), base).astype(np.uint8)
    d = (4000 + 30 * xx + 11 * yy + rng.integers(0, 9, (H, W))).astype(np.uint16)   # 16-bit depth, holes = 0
    d[rng.random((H, W)) < 0.07] = 0
    d[0, 0] = 65535
    return d


def _pil_decode(data):
    from PIL import Image
    im = Image.open(io.BytesIO(data))
    return np.array(im)


@pytest.mark.parametrize("kind", ["rgb", "rgba", "grey", "depth"])
def test_oracle_png_equals_pillow(kind):
    rng = np.random.default_rng(3)
    a = _img(rng, 29, 37, kind)
    for filt in [0, 1, 2, 3, 4, None]:
        data = OP.png_encode(a, filters=filt, seed=5, idat_split=97 if filt == 4 else 0)
        W, H, ch, bits, rec = OP.png_decode(data)
        got = rec.reshape(H, W, -1)
        if kind == "depth":
            got = (got[:, :, 0].astype(np.uint16) << 8) | got[:, :, 1]
        elif kind == "grey":
            got = got[:, :, 0]
        assert np.array_equal(got, a)                         # the writer and the reader invert each other
        pil = _pil_decode(data)                               # and an independent decoder reads the same pixels
        assert np.array_equal(pil.astype(a.dtype), a), (kind, filt)


def test_oracle_reads_pillow_written_files():
    from PIL import Image
    rng = np.random.default_rng(4)
    for kind, mode in [("rgb", "RGB"), ("rgba", "RGBA"), ("grey", "L"), ("depth", "I;16")]:
        a = _img(rng, 48, 64, kind)
        buf = io.BytesIO()
        Image.fromarray(a).save(buf, format="PNG")              # Pillow / zlib pick their own filters
        data = buf.getvalue()
        if kind == "depth":
            d = OP.depth_metres(data)
            assert np.array_equal(np.isnan(d), a == 0)
            ok = a != 0
            assert np.array_equal(d[ok], a[ok].astype(np.float32) * np.float32(1.0 / 5000.0))
            assert abs(float(d[0, 0]) - 65535 / 5000.0) < 1e-5
        else:
            bgr = OP.imread_bgr(data)
            want = np.repeat(a[:, :, None], 3, 2) if kind == "grey" else a[:, :, 2::-1]
            assert np.array_equal(bgr, want)


def test_syncidx_matches_oracle(tmp_path, api):
    from lineslam_b200 import tum
    d = tmp_path / "rgbd_dataset_freiburg1_xyz"
    d.mkdir()
    assert tum.read_syncidx(str(d)) == [] == OP.read_syncidx(str(d))          # no list: empty (reference :1206)
    rows = [(1305031102.175304 + 0.033 * i, f"rgb/{1305031102.175304 + 0.033 * i:.6f}.png",
             1305031102.160407 + 0.033 * i, f"depth/{1305031102.160407 + 0.033 * i:.6f}.png") for i in range(40)]
    txt = "".join(f"{a:.6f} {b}\t{c:.6f}   {e}\n" + ("\n" if i % 7 == 0 else "") for i, (a, b, c, e) in enumerate(rows))
    (d / "syncidx.txt").write_text(txt + "1305031199.0 rgb/tail.png 1305031199.1")   # incomplete last group
    got, want = tum.read_syncidx(str(d)), OP.read_syncidx(str(d))
    assert got == want and len(got) == 40
    assert got[3][1] == rows[3][1] and got[3][0] == float(f"{rows[3][0]:.6f}")


def test_png_info_and_rejects(api):
    from lineslam_b200 import tum
    rng = np.random.default_rng(1)
    assert tum.png_info(OP.png_encode(_img(rng, 12, 20, "rgb"))) == (20, 12, 3, 8)
    assert tum.png_info(OP.png_encode(_img(rng, 12, 20, "depth"))) == (20, 12, 1, 16)
    assert tum.png_info(OP.png_encode(_img(rng, 12, 20, "rgba"))) == (20, 12, 4, 8)
    with pytest.raises(api.LslError):
        tum.png_info(b"not a png at all, just forty bytes of text.....")
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(_img(rng, 12, 20, "grey")).convert("P").save(buf, format="PNG")   # palette: unsupported
    with pytest.raises(api.LslError):
        tum.png_info(buf.getvalue())
