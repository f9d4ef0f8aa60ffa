"""TEST INFRASTRUCTURE — CPU restatement of the TUM raw-directory ingest. Only tests/ may
import this; the product (lineslam_b200/csrc/k_tum.cu) never does.

  read_syncidx     OpenNIListener::loadRawData, the list part        src/openni_listener.cpp:1201-1217
  imread_bgr       cv::imread(rgbname, 1)                            src/openni_listener.cpp:1233
  depth_metres     imread(ANYDEPTH) + convertTo(CV_32FC1) + NaN rule + "/ 5000.0"   src/openni_listener.cpp:1234-1244

cv::imread is OpenCV + libpng, neither vendored in the reference tree nor present in this image. PNG decoding is a
lossless, fully specified format (ISO/IEC 15948: zlib stream, five scan-line filters), so the restatement follows the
specification and is PINNED against an independent decoder (Pillow, tests/test_tum.py) instead of the absent library.
The "/ 5000.0" step is OpenCV 2.4's MatExpr scale, convertTo(CV_32F, alpha = 1 / 5000.0), whose 32f kernel multiplies
by (float)alpha — restated from memory of modules/core/src/{matop,convert}.cpp, UNVERIFIED (parity unpinned for that
single multiply).

Also here: a minimal PNG writer with a caller-chosen filter type per row, so that tests reach all five filters.
"""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np

_SIG = b"\x89PNG\r\n\x1a\n"


def read_syncidx(dirname):
    """`while (fsync >> tmp)` in groups of four tokens; a trailing incomplete group is dropped."""
    path = os.path.join(dirname, "syncidx.txt")
    if not os.path.exists(path):
        return []
    tok = open(path).read().split()
    return [(float(tok[i]), tok[i + 1], float(tok[i + 2]), tok[i + 3]) for i in range(0, len(tok) - 3, 4)]


def _chunks(data: bytes):
    assert data[:8] == _SIG
    off = 8
    while off + 12 <= len(data):
        n, typ = struct.unpack(">I4s", data[off:off + 8])
        yield typ, data[off + 8:off + 8 + n]
        off += 12 + n
        if typ == b"IEND":
            break


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def png_decode(data: bytes):
    """-> (W, H, samples per pixel, bit depth, reconstructed bytes [H][W * bpp])."""
    W = H = bits = ctype = None
    idat = b""
    for typ, body in _chunks(data):
        if typ == b"IHDR":
            W, H, bits, ctype, comp, flt, inter = struct.unpack(">IIBBBBB", body)
            assert comp == 0 and flt == 0 and inter == 0
        elif typ == b"IDAT":
            idat += body
    ch = {0: 1, 2: 3, 6: 4}[ctype]
    bpp = ch * bits // 8
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(H, 1 + W * bpp)
    rec = np.zeros((H, W * bpp), np.int64)
    zero = np.zeros(W * bpp, np.int64)
    for y in range(H):
        ft = int(raw[y, 0])
        f = raw[y, 1:].astype(np.int64)
        up = rec[y - 1] if y > 0 else zero
        if ft == 0:
            rec[y] = f
        elif ft == 2:
            rec[y] = (f + up) & 255
        elif ft == 1:
            r = f.reshape(W, bpp)
            rec[y] = (np.cumsum(r, axis=0) & 255).reshape(-1)
        else:
            row = [0] * (W * bpp)
            upl = up.tolist()
            fl = f.tolist()
            for i in range(W * bpp):
                a = row[i - bpp] if i >= bpp else 0
                b = upl[i]
                c = upl[i - bpp] if i >= bpp else 0
                pred = ((a + b) >> 1) if ft == 3 else _paeth(a, b, c)
                row[i] = (fl[i] + pred) & 255
            rec[y] = row
    return W, H, ch, bits, rec.astype(np.uint8)


def imread_bgr(data: bytes) -> np.ndarray:
    """cv::imread(name, 1): always 3 x 8 bit, BGR; grey is replicated, alpha is dropped."""
    W, H, ch, bits, rec = png_decode(data)
    assert bits == 8
    px = rec.reshape(H, W, ch)
    if ch == 1:
        return np.repeat(px, 3, axis=2)
    return np.ascontiguousarray(px[:, :, 2::-1])


def depth_metres(data: bytes) -> np.ndarray:
    W, H, ch, bits, rec = png_decode(data)
    assert ch == 1 and bits == 16
    v = rec.reshape(H, W, 2).astype(np.uint16)
    d = ((v[:, :, 0] << 8) | v[:, :, 1]).astype(np.float32)        # convertTo(CV_32FC1)
    d[d.astype(np.float64) < 1e-5] = np.float32("nan")               # openni_listener.cpp:1238-1241
    return d * np.float32(1.0 / 5000.0)                              # MatExpr "/ 5000.0" (float multiply)


# ------------------------------------------------------------------------------------- test-side PNG writer ----
def png_encode(arr: np.ndarray, filters=None, seed: int = 0, idat_split: int = 0, level: int = 6) -> bytes:
    """arr: uint8 [H][W] / [H][W][3|4] (file order RGB(A)) or uint16 [H][W] (16-bit grey, written big endian).
    filters: None -> random type 0..4 per row (seeded); int -> that type on every row; list -> per row."""
    if arr.dtype == np.uint16:
        H, W = arr.shape
        bits, ctype, bpp = 16, 0, 2
        raw = arr.astype(">u2").view(np.uint8).reshape(H, W * 2)
    else:
        H, W = arr.shape[:2]
        ch = 1 if arr.ndim == 2 else arr.shape[2]
        bits, ctype, bpp = 8, {1: 0, 3: 2, 4: 6}[ch], ch
        raw = np.ascontiguousarray(arr, np.uint8).reshape(H, W * bpp)
    if filters is None:
        filters = np.random.default_rng(seed).integers(0, 5, H).tolist()
    elif isinstance(filters, int):
        filters = [filters] * H
    r = raw.astype(np.int64)
    left = np.zeros_like(r); left[:, bpp:] = r[:, :-bpp]
    up = np.zeros_like(r); up[1:] = r[:-1]
    ul = np.zeros_like(r); ul[1:, bpp:] = r[:-1, :-bpp]
    p = left + up - ul
    pa, pb, pc = np.abs(p - left), np.abs(p - up), np.abs(p - ul)
    pae = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, ul))
    preds = [np.zeros_like(r), left, up, (left + up) >> 1, pae]
    out = np.zeros((H, 1 + W * bpp), np.uint8)
    for y in range(H):
        out[y, 0] = filters[y]
        out[y, 1:] = (r[y] - preds[filters[y]][y]) & 255
    z = zlib.compress(out.tobytes(), level)

    def chunk(t, b):
        return struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b) & 0xFFFFFFFF)
    parts = [z] if idat_split <= 0 else [z[i:i + idat_split] for i in range(0, len(z), idat_split)]
    return (_SIG + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, bits, ctype, 0, 0, 0)) + chunk(b"tEXt", b"Comment\x00lsl test") +
            b"".join(chunk(b"IDAT", q) for q in parts) + chunk(b"IEND", b""))
