"""ctypes access to oracle/_ref: the UNMODIFIED upstream LSD 1.5 and levmar 2.6 compiled from /root/reference
by oracle/Makefile (test infrastructure; the prebuilt .so files travel with the repo snapshot)."""
import ctypes as C
import os

import numpy as np

REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


class NTuple(C.Structure):
    _fields_ = [("size", C.c_uint), ("max_size", C.c_uint), ("dim", C.c_uint), ("values", C.POINTER(C.c_double))]


class ImageDouble(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_double)), ("xsize", C.c_uint), ("ysize", C.c_uint)]


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "liblsd_ref.so")) and os.path.exists(os.path.join(REF_DIR, "liblevmar_ref.so"))


_lsd = None


def lsd_ref(gray, ang_th=22.5, density_th=0.7):
    """external/lsd/lsd-1.5/lsd.c LineSegmentDetection with the arguments lsd_scale() uses (lsd.cpp:2070-2091)."""
    global _lsd
    if _lsd is None:
        _lsd = C.CDLL(os.path.join(REF_DIR, "liblsd_ref.so"))
        _lsd.new_image_double.restype = C.POINTER(ImageDouble)
        _lsd.new_image_double.argtypes = [C.c_uint, C.c_uint]
        _lsd.LineSegmentDetection.restype = C.POINTER(NTuple)
        _lsd.LineSegmentDetection.argtypes = [C.POINTER(ImageDouble), C.c_double, C.c_double, C.c_double, C.c_double,
                                              C.c_double, C.c_double, C.c_int, C.c_double, C.c_void_p]
        _lsd.free_image_double.argtypes = [C.POINTER(ImageDouble)]
        _lsd.free_ntuple_list.argtypes = [C.POINTER(NTuple)]
    g = np.ascontiguousarray(gray, np.float64)
    H, W = g.shape
    im = _lsd.new_image_double(W, H)
    C.memmove(im.contents.data, g.ctypes.data, g.nbytes)
    out = _lsd.LineSegmentDetection(im, 0.8, 0.6, 2.0, ang_th, 0.0, density_th, 1024, 255.0, None)
    n, dim = out.contents.size, out.contents.dim
    segs = np.ctypeslib.as_array(out.contents.values, shape=(n, dim)).copy() if n else np.zeros((0, 5))
    _lsd.free_ntuple_list(out)
    _lsd.free_image_double(im)
    return segs


LMFUNC = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_void_p)
_lm = None


def levmar_ref(func, p0, x, itmax, opts):
    """external/levmar-2.6 dlevmar_dif (built-in LU, no LAPACK). Returns (ret, p, info)."""
    global _lm
    if _lm is None:
        _lm = C.CDLL(os.path.join(REF_DIR, "liblevmar_ref.so"))
    p = np.array(p0, np.float64)
    x = np.array(x, np.float64)
    info = np.zeros(10)
    o = np.array(opts, np.float64)
    ret = _lm.dlevmar_dif(func, p.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), len(p), len(x), itmax,
                          o.ctypes.data_as(C.c_void_p), info.ctypes.data_as(C.c_void_p), None, None, None)
    return ret, p, info
