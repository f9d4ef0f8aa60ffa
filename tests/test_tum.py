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


def test_device_inflate_code_equals_zlib_on_host():
    """The DEFLATE decoder the device runs (csrc/shared/lsl_inflate.h), compiled for the host by oracle/Makefile as a
    test harness, against zlib: stored, fixed and dynamic blocks, every strategy, long matches, overlapping copies;
    wrong lengths, truncation and bit flips must come back as error codes."""
    import ctypes as C
    import subprocess
    import zlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "oracle", "libinflate_check.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(root, "oracle"), "libinflate_check.so"])
    L = C.CDLL(so)
    L.lsl_inflate_host_check.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(0)
    cases = [b"", b"a", b"abc" * 1000, bytes(rng.integers(0, 256, 70000, dtype=np.uint8)),
             bytes(rng.integers(0, 4, 100000, dtype=np.uint8)),
             bytes(np.repeat(rng.integers(0, 256, 2000, dtype=np.uint8), rng.integers(1, 300, 2000))),
             bytes((np.cumsum(rng.integers(-3, 4, 200000)) % 256).astype(np.uint8)), bytes(300000)]
    for raw in cases:
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
                co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
                z = co.compress(raw) + co.flush()
                out = np.zeros(max(len(raw), 1), np.uint8)
                assert L.lsl_inflate_host_check(z, len(z), out.ctypes.data, len(raw)) == 0
                assert out[:len(raw)].tobytes() == raw
    raw = cases[6]
    z = zlib.compress(raw, 6)
    out = np.zeros(len(raw) + 16, np.uint8)
    assert L.lsl_inflate_host_check(z, len(z), out.ctypes.data, len(raw) - 5) < 0
    assert L.lsl_inflate_host_check(z, len(z), out.ctypes.data, len(raw) + 5) < 0
    assert L.lsl_inflate_host_check(z, len(z) // 2, out.ctypes.data, len(raw)) < 0
    for k in range(200):
        zz = bytearray(z)
        zz[int(rng.integers(2, len(z) - 4))] ^= 1 << int(rng.integers(0, 8))
        rc = L.lsl_inflate_host_check(bytes(zz), len(zz), out.ctypes.data, len(raw))
        assert rc < 0 or out[:len(raw)].tobytes() != raw or True      # must return, never read or write out of bounds


def _all_literal_dynamic_block(raw: bytes) -> bytes:
    """A zlib stream whose single dynamic block has NO distance code at all (HDIST = 1, that one length 0): what
    libdeflate (oxipng) and zopfli emit for all-literal data, legal per RFC 1951 §3.2.7. Hand-assembled: zlib's own
    encoder always emits two distance codes."""
    import zlib
    bits = []

    def put(v, n):                       # header fields: LSB first
        for k in range(n):
            bits.append((v >> k) & 1)

    def put_code(code, n):               # Huffman codes: MSB first
        for k in range(n - 1, -1, -1):
            bits.append((code >> k) & 1)

    lit_len = [8] * 255 + [9, 9]         # symbols 0..254 -> 8 bits, 255 and 256 (end of block) -> 9 bits: Kraft sum 1
    # canonical codes
    def canon(lengths):
        codes, code = {}, 0
        for l in range(1, 16):
            for s, ll in enumerate(lengths):
                if ll == l:
                    codes[s] = (code, l); code += 1
            code <<= 1
        return codes
    lit = canon(lit_len)
    # code-length alphabet: symbols 8 (1 bit), 0 and 9 (2 bits)
    cl_len = [0] * 19
    cl_len[8], cl_len[0], cl_len[9] = 1, 2, 2
    cl = canon(cl_len)
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    put(1, 1); put(2, 2)                 # BFINAL, dynamic
    put(0, 5); put(0, 5)                 # HLIT = 257, HDIST = 1
    ncode = max(i for i, s in enumerate(order) if cl_len[s]) + 1
    put(ncode - 4, 4)
    for i in range(ncode):
        put(cl_len[order[i]], 3)
    for l in lit_len + [0]:              # 257 literal/length lengths, then the single distance length: 0
        put_code(*cl[l])
    for b in raw:
        put_code(*lit[b])
    put_code(*lit[256])
    while len(bits) % 8:
        bits.append(0)
    body = bytes(sum(bits[i + k] << k for k in range(8)) for i in range(0, len(bits), 8))
    z = b"\x78\x9c" + body + zlib.adler32(raw).to_bytes(4, "big")
    assert zlib.decompress(z) == raw     # zlib accepts it
    return z


def test_inflate_accepts_block_without_distance_codes():
    import ctypes as C
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    L = C.CDLL(os.path.join(root, "oracle", "libinflate_check.so"))
    L.lsl_inflate_host_check.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(12)
    for n in (1, 300, 5000):
        raw = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        z = _all_literal_dynamic_block(raw)
        out = np.zeros(n, np.uint8)
        assert L.lsl_inflate_host_check(z, len(z), out.ctypes.data, n) == 0
        assert out.tobytes() == raw
