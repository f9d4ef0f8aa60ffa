"""Host-side mirror of OpenNIListener::loadRawData (src/openni_listener.cpp:1194-1319) over include/lsl_tum.h:
syncidx.txt -> (timestamps, file names), PNG file images -> device-resident BGR / depth planes -> Node frames.
Marshalling only; inflate + unfilter + conversions happen in liblsl_b200.so (host threads + png_unfilter_kernel)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .api import Frame, _check, lib
from .records import ptr


class TumEntry(C.Structure):
    _fields_ = [("ts_rgb", C.c_double), ("ts_depth", C.c_double), ("rgb", C.c_char * 256), ("depth", C.c_char * 256)]


def read_syncidx(dirname: str):
    """[(ts_rgb, rgb_file, ts_depth, depth_file)] of <dirname>/syncidx.txt."""
    n = C.c_int(0)
    lib().lsl_tum_read_syncidx(dirname.encode(), None, 0, C.byref(n))
    ent = (TumEntry * max(n.value, 1))()
    _check(lib().lsl_tum_read_syncidx(dirname.encode(), ent, len(ent), C.byref(n)))
    return [(e.ts_rgb, e.rgb.decode(), e.ts_depth, e.depth.decode()) for e in ent[:n.value]]


def png_info(data: bytes):
    w, h, ch, bits = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    _check(lib().lsl_png_info(data, C.c_size_t(len(data)), C.byref(w), C.byref(h), C.byref(ch), C.byref(bits)))
    return w.value, h.value, ch.value, bits.value


def _lists(files):
    n = len(files)
    bufs = [C.create_string_buffer(f, len(f)) for f in files]
    ptrs = (C.c_void_p * n)(*[C.addressof(b) for b in bufs])
    lens = (C.c_size_t * n)(*[len(f) for f in files])
    return bufs, ptrs, lens


def decode_batch(ctx, rgb_files, depth_files, W: int, H: int, d_bgr: int = 0, d_depth: int = 0):
    """PNG file images (bytes) -> device buffers (raw device pointers of n*H*W*3 u8 / n*H*W f32)."""
    n = len(rgb_files) if rgb_files is not None else len(depth_files)
    rb, rp, rl = _lists(rgb_files) if rgb_files is not None else (None, None, None)
    db, dp, dl = _lists(depth_files) if depth_files is not None else (None, None, None)
    _check(lib().lsl_tum_decode_batch(ctx._h, n, rp, rl, dp, dl, W, H, C.c_void_p(d_bgr or None), C.c_void_p(d_depth or None)), ctx._h)


def extract_batch(ctx, rgb_files, depth_files, W: int, H: int, K=None, seeds=None, dt: float = 0.0):
    """loadRawData + Node::Node for n frames; K = None uses the TUM intrinsics of openni_listener.cpp:1256-1260."""
    n = len(rgb_files)
    rb, rp, rl = _lists(rgb_files)
    db, dp, dl = _lists(depth_files)
    sd = np.ascontiguousarray(seeds if seeds is not None else np.arange(1, n + 1), np.uint32)
    Kc = None if K is None else np.ascontiguousarray(K, np.float64).reshape(9)
    out = (C.c_void_p * n)()
    _check(lib().lsl_extract_tum_batch(ctx._h, n, rp, rl, dp, dl, W, H, ptr(Kc), C.c_double(dt), ptr(sd), out), ctx._h)
    return [Frame(ctx, out[i]) for i in range(n)]


def load_raw_data(ctx, dirname: str, batch: int = 64, skip_first_n_frames: int = 0, data_skip_step: int = 1, K=None):
    """Generator over (timestamps, frames) batches of a TUM raw directory, in list order."""
    ent = [e for i, e in enumerate(read_syncidx(dirname)) if i >= skip_first_n_frames and i % data_skip_step == 0]
    for b in range(0, len(ent), batch):
        part = ent[b:b + batch]
        rgb = [open(os.path.join(dirname, e[1]), "rb").read() for e in part]
        dep = [open(os.path.join(dirname, e[3]), "rb").read() for e in part]
        W, H, _, _ = png_info(rgb[0])
        yield [e[0] for e in part], extract_batch(ctx, rgb, dep, W, H, K, seeds=np.arange(b + 1, b + 1 + len(part)),
                                                  dt=0.0)


def release(ctx):
    lib().lsl_tum_release(ctx._h)
