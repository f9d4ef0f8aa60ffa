"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package lineslam_b200 never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

from lineslam_b200.records import Params, LINE_DTYPE, MATCH_DTYPE, POSE_DTYPE, ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/external/lsd/lsd-1.5"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        for n in "exp log log10 sin cos sinh acos".split():
            f = getattr(L, "orc_m_" + n); f.restype = C.c_double; f.argtypes = [C.c_double]
        for n in "atan2 pow".split():
            f = getattr(L, "orc_m_" + n); f.restype = C.c_double; f.argtypes = [C.c_double, C.c_double]
        L.orc_stream_step.restype = C.c_double
        L.orc_cvnorm.restype = C.c_double
        L.orc_cvnorm_diff72.restype = C.c_double
        L.orc_l2sqr_f.restype = C.c_float
        L.orc_inv3.restype = C.c_double
        L.orc_error_function2.restype = C.c_double
    return _LIB


def default_params():
    p = Params()
    lib().orc_params_default(C.byref(p))
    return p


def rand(seed, n):
    out = np.zeros(n, np.int32)
    lib().orc_rand(C.c_uint32(seed), n, ptr(out))
    return out


def gray(img):
    H, W, _ = img.shape
    g = np.zeros((H, W), np.uint8)
    lib().orc_gray(ptr(np.ascontiguousarray(img)), W, H, ptr(g))
    return g


def lsd(gray_img, params=None, debug=False):
    p = params or default_params()
    g = np.ascontiguousarray(gray_img, np.uint8)
    H, W = g.shape
    cap = 8192
    segs = np.zeros((cap, 5))
    if not debug:
        n = lib().orc_lsd(ptr(g), W, H, C.byref(p), ptr(segs), cap, None, None, None, None, None, None, None)
        return segs[:n].copy()
    sw, sh = int(np.floor(W * p.lsd_scale)), int(np.floor(H * p.lsd_scale))
    scaled = np.zeros((sh, sw)); ang = np.zeros((sh, sw)); mod = np.zeros((sh, sw))
    seeds = np.zeros(sh * sw, np.int32); ns = C.c_int(0); a = C.c_int(0); b = C.c_int(0)
    n = lib().orc_lsd(ptr(g), W, H, C.byref(p), ptr(segs), cap, ptr(scaled), ptr(ang), ptr(mod), ptr(seeds),
                      C.byref(ns), C.byref(a), C.byref(b))
    return segs[:n].copy(), dict(scaled=scaled, angles=ang, modgrad=mod, seeds=seeds[:ns.value].copy())


def sobel5(gray_img):
    g = np.ascontiguousarray(gray_img, np.uint8)
    H, W = g.shape
    gx = np.zeros((H, W)); gy = np.zeros((H, W))
    lib().orc_sobel5(ptr(g), W, H, ptr(gx), ptr(gy))
    return gx, gy


def detect3DLines(img, depth, K, seed=1, params=None, dt=0.0, omp_threads=1, debug=False):
    p = params or default_params()
    img = np.ascontiguousarray(img, np.uint8)
    depth = np.ascontiguousarray(depth, np.float32)
    H, W = depth.shape
    ch = 3 if img.ndim == 3 else 1
    cap = 4096
    out = np.zeros(cap, LINE_DTYPE)
    Kc = np.ascontiguousarray(K, np.float64)
    if not debug:
        n = lib().orc_detect3DLines(ptr(img), ch, ptr(depth), W, H, ptr(Kc), C.c_double(dt), C.c_uint32(seed),
                                    C.byref(p), ptr(out), cap, omp_threads, None, None, None, None, None, None, None, 0)
        return out[:n].copy()
    sol = np.zeros(cap, np.int32); ninl = np.zeros(cap, np.int32); idx = np.full((cap, 101), -1, np.int32)
    a0b0 = np.zeros((cap, 6)); its = np.zeros(cap, np.int32); nsegs = C.c_int(0); segs = np.zeros((8192, 5))
    n = lib().orc_detect3DLines(ptr(img), ch, ptr(depth), W, H, ptr(Kc), C.c_double(dt), C.c_uint32(seed), C.byref(p),
                                ptr(out), cap, omp_threads, ptr(sol), ptr(ninl), ptr(idx), ptr(a0b0), ptr(its),
                                C.byref(nsegs), ptr(segs), 8192)
    return out[:n].copy(), dict(seg_of_line=sol[:n].copy(), n_inl=ninl[:n].copy(), inl_idx=idx[:n].copy(),
                                a0b0=a0b0[:n].copy(), lm_iters=its[:n].copy(), segs=segs[:nsegs.value].copy())


def lineMatching(f1, f2, adjacent=True, omp_threads=1):
    f1 = np.ascontiguousarray(f1, LINE_DTYPE); f2 = np.ascontiguousarray(f2, LINE_DTYPE)
    cap = max(len(f1), 1)
    out = np.zeros(cap, MATCH_DTYPE)
    n = lib().orc_lineMatching(ptr(f1), len(f1), ptr(f2), len(f2), int(adjacent), ptr(out), cap, omp_threads)
    return out[:n].copy()


def pose_ransac(train, query, matches, id_train=0, id_query=1, seed=1, params=None):
    p = params or default_params()
    train = np.ascontiguousarray(train, LINE_DTYPE); query = np.ascontiguousarray(query, LINE_DTYPE)
    m = np.ascontiguousarray(matches, MATCH_DTYPE)
    rec = np.zeros(1, POSE_DTYPE)
    cap = max(len(m), 1)
    inl = np.zeros(cap, MATCH_DTYPE); rinl = np.zeros(cap, MATCH_DTYPE)
    n1 = C.c_int(0); n2 = C.c_int(0); tfr = np.zeros(16, np.float32)
    lib().orc_pose_ransac(ptr(train), len(train), ptr(query), len(query), id_train, id_query, ptr(m), len(m),
                          C.c_uint32(seed), C.byref(p), ptr(rec), ptr(inl), cap, C.byref(n1), ptr(rinl), cap,
                          C.byref(n2), ptr(tfr))
    return rec[0].copy(), inl[:n1.value].copy(), rinl[:n2.value].copy(), tfr.reshape(4, 4)


def stream_step(img, depth, K, seed, prev_lines, params=None, omp_threads=1):
    """One reference CPU step (extract + match against prev + RANSAC pose); returns (seconds, lines, pose)."""
    p = params or default_params()
    img = np.ascontiguousarray(img, np.uint8); depth = np.ascontiguousarray(depth, np.float32)
    H, W = depth.shape
    ch = 3 if img.ndim == 3 else 1
    cap = 4096
    cur = np.zeros(cap, LINE_DTYPE); ncur = C.c_int(0); rec = np.zeros(1, POSE_DTYPE)
    prev = np.ascontiguousarray(prev_lines, LINE_DTYPE) if prev_lines is not None else np.zeros(0, LINE_DTYPE)
    Kc = np.ascontiguousarray(K, np.float64)
    sec = lib().orc_stream_step(ptr(img), ch, ptr(depth), W, H, ptr(Kc), C.c_uint32(seed), C.byref(p), ptr(prev),
                                len(prev), ptr(cur), cap, C.byref(ncur), ptr(rec), omp_threads)
    return sec, cur[:ncur.value].copy(), rec[0].copy()


def rootsift(desc):
    d = np.ascontiguousarray(desc, np.float32).copy()
    lib().orc_rootsift(ptr(d), d.shape[0], d.shape[1])
    return d


def featureMatching(qdesc, tdesc, nn_ratio=0.5, seed=1):
    """Node::featureMatching, BRUTEFORCE branch; one rand() per returned match."""
    q = np.ascontiguousarray(qdesc, np.float32); t = np.ascontiguousarray(tdesc, np.float32)
    cap = max(len(q), 1)
    out = np.zeros(cap, MATCH_DTYPE)
    n = lib().orc_featureMatching(ptr(q), len(q), ptr(t), len(t), q.shape[1] if q.ndim == 2 else 0, C.c_double(nn_ratio),
                                  C.c_uint32(seed), ptr(out), cap)
    return out[:n].copy()


def pose_ransac_hybrid(train, query, train_xyz1, query_xyz1, pt_matches, ln_matches, id_train=0, id_query=1, seed=1,
                       skip_draws=0, fx=525.0, dt=0.0, params=None):
    p = params or default_params()
    train = np.ascontiguousarray(train, LINE_DTYPE); query = np.ascontiguousarray(query, LINE_DTYPE)
    tx = np.ascontiguousarray(train_xyz1, np.float32).reshape(-1, 4); qx = np.ascontiguousarray(query_xyz1, np.float32).reshape(-1, 4)
    pm = np.ascontiguousarray(pt_matches, MATCH_DTYPE); m = np.ascontiguousarray(ln_matches, MATCH_DTYPE)
    rec = np.zeros(1, POSE_DTYPE)
    inl = np.zeros(max(len(m), 1), MATCH_DTYPE); rinl = np.zeros(max(len(m), 1), MATCH_DTYPE)
    pinl = np.zeros(max(len(pm), 1), MATCH_DTYPE); prinl = np.zeros(max(len(pm), 1), MATCH_DTYPE)
    n = [C.c_int(0) for _ in range(4)]
    tfr = np.zeros(16, np.float32)
    lib().orc_pose_ransac_hybrid(ptr(train), len(train), ptr(query), len(query), ptr(tx), len(tx), ptr(qx), len(qx),
                                 id_train, id_query, ptr(pm), len(pm), ptr(m), len(m), C.c_uint32(seed), skip_draws,
                                 C.c_double(fx), C.c_double(dt), C.byref(p), ptr(rec), ptr(inl), C.byref(n[0]), ptr(rinl),
                                 C.byref(n[1]), ptr(pinl), C.byref(n[2]), ptr(prinl), C.byref(n[3]), ptr(tfr))
    return dict(rec=rec[0].copy(), ln_inliers=inl[:n[0].value].copy(), ln_ransac_inliers=rinl[:n[1].value].copy(),
                pt_inliers=pinl[:n[2].value].copy(), pt_ransac_inliers=prinl[:n[3].value].copy(), tf_ransac=tfr.reshape(4, 4))


def error_function2(x1, x2, tf, sigma_depth=0.01):
    a = np.ascontiguousarray(x1, np.float32); b = np.ascontiguousarray(x2, np.float32)
    t = np.ascontiguousarray(tf, np.float32).reshape(16)
    return lib().orc_error_function2(ptr(a), ptr(b), ptr(t), C.c_double(sigma_depth))


def compute_inliers_and_error(matches, origins, earlier, tf, squared_max, sigma_depth=0.01):
    """Node::computeInliersAndError (src/node.cpp:1019-1080) over errorFunction2: (indices of the inlier matches, rmse)."""
    keep, mean = [], 0.0
    for i, m in enumerate(matches):
        o, t = origins[int(m["queryIdx"])], earlier[int(m["trainIdx"])]
        if o[2] == 0.0 or t[2] == 0.0:            # does NOT trigger on NaN
            continue
        d = error_function2(o, t, tf, sigma_depth)
        if d > squared_max or not (d >= 0.0):
            continue
        mean += d
        keep.append(i)
    import math
    return keep, (1e9 if len(keep) < 3 else math.sqrt(mean / len(keep)))


def kabsch(frm, to, w):
    f = np.ascontiguousarray(frm, np.float32); t = np.ascontiguousarray(to, np.float32); ww = np.ascontiguousarray(w, np.float32)
    tf = np.zeros(16, np.float32)
    lib().orc_kabsch(ptr(f), ptr(t), ptr(ww), len(ww), ptr(tf))
    return tf.reshape(4, 4)


def ldlt3_solve(A, b):
    A = np.ascontiguousarray(A, np.float64); b = np.ascontiguousarray(b, np.float64); x = np.zeros(3)
    lib().orc_ldlt3_solve(ptr(A), ptr(b), ptr(x))
    return x


def relmotion_ransac(a, b, seed=1, params=None):
    """computeRelativeMotion_Ransac on paired line records a[i] <-> b[i] (x_b = R x_a + t)."""
    p = params or default_params()
    a = np.ascontiguousarray(a, LINE_DTYPE); b = np.ascontiguousarray(b, LINE_DTYPE)
    Rt = np.zeros(12); con = np.zeros(max(len(a), 1), np.int32); calls = C.c_int(0); have = C.c_int(0)
    n = lib().orc_relmotion_ransac(ptr(a), ptr(b), len(a), C.c_uint32(seed), C.byref(p), ptr(Rt), ptr(con), C.byref(calls),
                                   C.byref(have))
    return dict(R=Rt[:9].reshape(3, 3).copy(), t=Rt[9:].copy(), conset=con[:n].copy(), lm_calls=calls.value, have=bool(have.value))


def optimizeRelmotion(a, b, R, t):
    a = np.ascontiguousarray(a, LINE_DTYPE); b = np.ascontiguousarray(b, LINE_DTYPE)
    Rt = np.concatenate([np.asarray(R, np.float64).ravel(), np.asarray(t, np.float64).ravel()])
    lib().orc_optimizeRelmotion(ptr(a), ptr(b), len(a), ptr(Rt))
    return Rt[:9].reshape(3, 3).copy(), Rt[9:].copy()


# ---- third-party restatement probes (compared with cv2 4.13 in tests/test_oracle_cv2.py) ----
def cvnorm(v):
    v = np.ascontiguousarray(v, np.float64)
    return lib().orc_cvnorm(ptr(v), len(v))


def cvnorm_diff72(a, b):
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    return lib().orc_cvnorm_diff72(ptr(a), ptr(b))


def l2sqr_f(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return float(lib().orc_l2sqr_f(ptr(a), ptr(b), len(a)))


def cov_to_DU(cov):
    c = np.ascontiguousarray(cov, np.float64); DU = np.zeros((3, 3)); W = np.zeros(3)
    lib().orc_cov_to_DU(ptr(c), ptr(DU), ptr(W))
    return DU, W


def inv3(a):
    a = np.ascontiguousarray(a, np.float64); r = np.zeros((3, 3))
    det = lib().orc_inv3(ptr(a), ptr(r))
    return r, det


def inv6(a):
    a = np.ascontiguousarray(a, np.float64); r = np.zeros((6, 6))
    ok = lib().orc_inv6(ptr(a), ptr(r))
    return r, ok


def jacobi_sym(a):
    a = np.ascontiguousarray(a, np.float64); n = a.shape[0]
    w = np.zeros(n); V = np.zeros((n, n))
    (lib().orc_jacobi4 if n == 4 else lib().orc_jacobi3)(ptr(a), ptr(w), ptr(V))
    return w, V


def clip_line(W, H, pt1, pt2):
    pts = np.array([pt1[0], pt1[1], pt2[0], pt2[1]], np.int32)
    ok = lib().orc_clip_line(W, H, ptr(pts))
    return bool(ok), (int(pts[0]), int(pts[1])), (int(pts[2]), int(pts[3]))


def get_gradient(gx, gy, p, q):
    """FrameLine::getGradient over f64 gradient planes; returns r (2)."""
    gx = np.ascontiguousarray(gx, np.float64); gy = np.ascontiguousarray(gy, np.float64)
    H, W = gx.shape
    pq = np.array([p[0], p[1], q[0], q[1]], np.float64); r = np.zeros(2)
    lib().orc_get_gradient(ptr(gx), ptr(gy), W, H, ptr(pq), ptr(r))
    return r


def featureMatching_hamming(qdesc, tdesc, nn_ratio=0.5, seed=1):
    """Node::featureMatching for ORB rows (BruteForce-HammingLUT)."""
    q = np.ascontiguousarray(qdesc, np.uint8); t = np.ascontiguousarray(tdesc, np.uint8)
    out = np.zeros(max(len(q), 1), MATCH_DTYPE)
    n = lib().orc_featureMatching_hamming(ptr(q), len(q), ptr(t), len(t), q.shape[1], C.c_double(nn_ratio), C.c_uint32(seed),
                                          ptr(out), len(out))
    return out[:n].copy()
