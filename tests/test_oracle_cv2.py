"""Third-party arithmetic of the oracle against the OpenCV that IS in this image (cv2 4.13, not the reference's 2.4).

The reference calls OpenCV 2.4 for cv::norm, cv::SVD, Mat::inv, minMaxLoc and BFMatcher; its source is not under
/root/reference, so the oracle restates the published 2.4 algorithms. cv2 4.13 is an independent
implementation of the same functions: where it agrees bit for bit the restatement is pinned; where 4.x uses another
summation order (universal intrinsics, FMA) the measured difference is asserted as a bound and written down here
instead of being hidden:

  cv::norm (f64, 36/72-vector)      <= 4 ulp  (4.x accumulates in SIMD lanes with FMA; 2.4 pairs squares two by two)
  BFMatcher L2 knnMatch k = 2       nearest / second nearest INDICES and tie order identical; distances <= 2 ulp (f32)
  cv::invert 3x3 (DECOMP_LU)        bit-exact
  cv::invert 6x6 (DECOMP_LU)        <= 1e-13 relative (same elimination order, 4.x build contracts mul+add)
  cv::SVDecomp 3x3 / 4x4 symmetric  singular values 1e-12 relative, singular vectors up to sign 1e-8 (cyclic Jacobi on
                                    both sides, different sweep order)
  cv::minMaxLoc                     first minimum in row-major order, as the oracle's scan
"""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import pyoracle as po


def _ulp64(a, b):
    return abs(int(np.float64(a).view(np.int64)) - int(np.float64(b).view(np.int64)))


def _ulp32(a, b):
    return abs(int(np.float32(a).view(np.int32)) - int(np.float32(b).view(np.int32)))


@pytest.mark.parametrize("n", [36, 72])
def test_cvnorm_vs_cv2_norm(n):
    rng = np.random.default_rng(100 + n)
    worst = 0
    for _ in range(3000):
        v = rng.random(n) * rng.choice([1e-3, 1.0, 40.0])
        worst = max(worst, _ulp64(po.cvnorm(v), cv2.norm(v.reshape(-1, 1))))
    assert worst <= 4, worst      # measured: 3


def test_cvnorm_diff72_vs_cv2_norm():
    rng = np.random.default_rng(7)
    worst = 0
    for _ in range(3000):
        a, b = rng.random(72), rng.random(72)
        d = po.cvnorm_diff72(a, b)
        assert d == po.cvnorm(a - b)                      # cv::norm(des1 - des2): MatExpr materialises the difference first
        worst = max(worst, _ulp64(d, cv2.norm((a - b).reshape(-1, 1))))
    assert worst <= 4, worst


def _rootsift_rows(rng, n, dim):
    d = rng.random((n, dim)).astype(np.float32) ** 3
    return po.rootsift(d)


@pytest.mark.parametrize("dim", [64, 128])
def test_bfmatcher_knn_indices_ties_distances(dim):
    """Node::featureMatching's BFMatcher(NORM_L2).knnMatch(k = 2) (src/node.cpp:614-616)."""
    rng = np.random.default_rng(dim)
    q = _rootsift_rows(rng, 600, dim)
    t = _rootsift_rows(rng, 600, dim)
    sel = rng.permutation(600)[:300]
    t[:300] = np.abs(q[sel] + rng.normal(0, 0.002, (300, dim)).astype(np.float32))
    t[10] = t[11]                                          # exact ties: the earlier train row must come first
    t[400] = t[20]
    knn = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, k=2)
    worst = 0
    for i in range(0, 600, 3):
        ds = np.array([np.sqrt(np.float32(po.l2sqr_f(q[i], t[j]))) for j in range(600)], np.float32)
        order = np.argsort(ds, kind="stable")
        a, b = knn[i]
        assert (a.trainIdx, b.trainIdx) == (order[0], order[1]), i
        worst = max(worst, _ulp32(a.distance, ds[a.trainIdx]), _ulp32(b.distance, ds[b.trainIdx]))
    assert worst <= 2, worst


def test_feature_matching_vs_cv2_pipeline():
    """The whole BRUTEFORCE branch (src/node.cpp:606-641) re-run with cv2's matcher: same (queryIdx, trainIdx) list in
    the same order; the stored distance (ratio + rand jitter) within the f32 rounding of the two distance ulps."""
    rng = np.random.default_rng(5)
    q = _rootsift_rows(rng, 600, 128)
    t = _rootsift_rows(rng, 600, 128)
    sel = rng.permutation(600)[:350]
    t[:350] = np.abs(q[sel] + rng.normal(0, 0.004, (350, 128)).astype(np.float32))
    got = po.featureMatching(q, t, nn_ratio=0.5, seed=11)
    knn = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, t, k=2)
    draws = po.rand(11, 600)
    exp, seen = [], set()
    for i, (a, b) in enumerate(knn):
        ratio = np.float32(a.distance) / np.float32(b.distance)
        if ratio < 0.5:
            if a.trainIdx in seen:
                continue
            seen.add(a.trainIdx)
            exp.append((i, a.trainIdx, np.float32(ratio + np.float32(draws[len(exp)]) / (1000.0 * 2147483647))))
    assert len(got) == len(exp) > 100
    assert [(int(m["queryIdx"]), int(m["trainIdx"])) for m in got] == [(e[0], e[1]) for e in exp]
    assert np.allclose(got["distance"], [e[2] for e in exp], rtol=0, atol=1e-6)


def test_invert_3x3_bit_exact_and_6x6_close():
    rng = np.random.default_rng(2)
    for _ in range(300):
        a = rng.normal(size=(3, 3)); a = a @ a.T + np.eye(3) * 0.1
        r, det = po.inv3(a)
        ok, c = cv2.invert(a, flags=cv2.DECOMP_LU)
        assert ok and np.array_equal(r, c)                # Mat::inv on 3x3: cofactors, bit for bit
        h = rng.normal(size=(6, 6)); h = h @ h.T + np.eye(6) * 0.1
        r6, ok6 = po.inv6(h)
        _, c6 = cv2.invert(h, flags=cv2.DECOMP_LU)
        assert ok6 and np.abs(r6 - c6).max() <= 1e-13 * np.abs(c6).max()


def test_cov_to_DU_vs_SVDecomp():
    """RandomPoint3d ctor (src/line/lineslam.h:59-81): W_sqrt and DU = diag(W^-1/2) U^T; only the whitened norm
    |DU x| is consumed downstream, which is invariant under the sign of a singular vector."""
    rng = np.random.default_rng(3)
    for _ in range(300):
        a = rng.normal(size=(3, 3)) * [1, 0.1, 0.01]
        cov = a @ a.T + np.eye(3) * 1e-6
        DU, W = po.cov_to_DU(cov)
        w, u, _ = cv2.SVDecomp(cov)
        assert np.abs(W - np.sqrt(w.ravel())).max() <= 1e-12 * W.max()
        DUc = np.diag(1 / np.sqrt(w.ravel())) @ u.T
        assert np.abs(np.abs(DU) - np.abs(DUc)).max() <= 1e-8 * np.abs(DUc).max()
        x = rng.normal(size=3)
        assert abs(np.linalg.norm(DU @ x) - np.linalg.norm(DUc @ x)) <= 1e-9 * np.linalg.norm(DUc @ x)


def test_zhang_quaternion_vs_SVDecomp_4x4():
    """computeRelativeMotion_svd (src/line/motion.cpp:353): q = svd.u.col(3) of the symmetric 4x4 A."""
    rng = np.random.default_rng(4)
    for _ in range(300):
        m = rng.normal(size=(6, 4))
        A = m.T @ m
        w, V = po.jacobi_sym(A)
        wc, u, _ = cv2.SVDecomp(A)
        assert np.abs(w - wc.ravel()).max() <= 1e-12 * w.max()
        q, qc = V[:, 3], u[:, 3]
        assert min(np.abs(q - qc).max(), np.abs(q + qc).max()) <= 1e-8


def test_pca_direction_vs_SVDecomp():
    """computeLine3d_svd (src/line/utils.cpp:471-493): direction = first right singular vector of the centred n x 3 matrix
    = dominant eigenvector of the scatter matrix the oracle diagonalises."""
    rng = np.random.default_rng(6)
    for _ in range(200):
        n = int(rng.integers(10, 100))
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        pts = np.outer(rng.uniform(-1, 1, n), d) + rng.normal(0, 0.01, (n, 3)) + rng.normal(size=3)
        c = pts - pts.mean(0)
        w, V = po.jacobi_sym(c.T @ c)
        _, _, vt = cv2.SVDecomp(c)
        v, vc = V[:, 0], vt[0]
        assert min(np.abs(v - vc).max(), np.abs(v + vc).max()) <= 1e-8


def test_minMaxLoc_returns_first_minimum():
    """Node::lineMatching reads cv::minMaxLoc on a row and on a column of the distance matrix (src/node.cpp:1660-1664);
    the oracle scans with strict <. A tie can never produce a match anyway: the second-best value then equals the best
    and `second * 0.7 > best` fails — the rule is pinned AND immaterial."""
    row = np.array([[3.0, 1.0, 2.0, 1.0, 5.0, 1.0]])
    assert cv2.minMaxLoc(row)[2] == (1, 0)
    col = row.T.copy()
    assert cv2.minMaxLoc(col)[2] == (0, 1)
    rng = np.random.default_rng(8)
    for _ in range(200):
        v = rng.integers(0, 5, size=(1, 40)).astype(np.float64)
        first = int(np.flatnonzero(v[0] == v.min())[0])
        assert cv2.minMaxLoc(v)[2] == (first, 0)
        assert cv2.minMaxLoc(v.T.copy())[2] == (0, first)


def test_clipLine_vs_cv2():
    """cv::clipLine inside the cv::LineIterator constructor (FrameLine::getGradient, src/line/lineslam.cpp:529)."""
    rng = np.random.default_rng(9)
    W, H = 640, 480
    cases = [((-3, 10), (50, 40)), ((600, 470), (660, 500)), ((-5, -5), (-1, 700)), ((639, 0), (640, 479)),
             ((10, -1), (10, 480)), ((-1, 5), (640, 5)), ((700, 600), (800, 700)), ((0, 0), (639, 479))]
    for _ in range(20000):
        cases.append(((int(rng.integers(-60, 700)), int(rng.integers(-60, 540))),
                      (int(rng.integers(-60, 700)), int(rng.integers(-60, 540)))))
    for p1, p2 in cases:
        ok, a, b = po.clip_line(W, H, p1, p2)
        ok2, a2, b2 = cv2.clipLine((0, 0, W, H), p1, p2)
        assert ok == ok2, (p1, p2)
        if ok:
            assert (a, b) == (tuple(a2), tuple(b2)), (p1, p2)


def test_getGradient_border_segments():
    """A segment whose rounded end point leaves the image is clipped and iterated (the reference), not skipped: r is the
    normalised gradient sum over the Bresenham pixels between the cv2.clipLine'd end points; a segment entirely outside
    gives count = 0 and r = NaN (0/0), as `xSum/len` does in the reference."""
    rng = np.random.default_rng(10)
    H, W = 120, 160
    gx = rng.integers(-6570, 6571, (H, W)).astype(np.float64)
    gy = rng.integers(-6570, 6571, (H, W)).astype(np.float64)

    def expect(p, q):
        p1 = (int(np.rint(p[0])), int(np.rint(p[1]))); p2 = (int(np.rint(q[0])), int(np.rint(q[1])))
        inside = all(0 <= a[0] < W and 0 <= a[1] < H for a in (p1, p2))
        if not inside:
            ok, p1, p2 = cv2.clipLine((0, 0, W, H), p1, p2)
            if not ok:
                return None
        (x1, y1), (x2, y2) = p1, p2
        dx, dy = abs(x2 - x1), abs(y2 - y1)
        sx, sy = (1 if x2 >= x1 else -1), (1 if y2 >= y1 else -1)
        steep = dy > dx
        dmaj, dmin = (dy, dx) if steep else (dx, dy)
        err = dmaj - 2 * dmin
        x, y, sxs, sys_ = x1, y1, 0.0, 0.0
        for _ in range(dmaj + 1):
            sxs += gx[y, x]; sys_ += gy[y, x]
            mask = err < 0
            err += -2 * dmin + (2 * dmaj if mask else 0)
            if steep:
                y += sy; x += sx if mask else 0
            else:
                x += sx; y += sy if mask else 0
        n = np.sqrt(sxs * sxs + sys_ * sys_)
        return np.array([sxs / n, sys_ / n])

    segs = [((-2.4, 10.2), (50.3, 40.7)), ((150.2, 100.9), (161.7, 118.2)), ((159.6, 3.0), (100.0, 60.0)),
            ((20.0, 119.6), (90.0, 60.0)), ((-0.6, -0.6), (30.0, 30.0)), ((170.0, 130.0), (200.0, 150.0)),
            ((-0.4, 50.0), (80.0, 50.0))]
    for _ in range(300):
        segs.append(((rng.uniform(-8, W + 8), rng.uniform(-8, H + 8)), (rng.uniform(-8, W + 8), rng.uniform(-8, H + 8))))
    n_clipped = 0
    for p, q in segs:
        r = po.get_gradient(gx, gy, p, q)
        e = expect(p, q)
        if e is None:
            assert np.all(np.isnan(r))
        else:
            assert np.array_equal(r, e), (p, q)
        n_clipped += any(not (0 <= np.rint(a[0]) < W and 0 <= np.rint(a[1]) < H) for a in (p, q))
    assert n_clipped > 20


def test_hamming_feature_matching_vs_cv2():
    """feature_extractor_type ORB: "BruteForce-HammingLUT" knnMatch k = 2 + ratio / unique / jitter pass
    (src/node.cpp:606-641). Integer distances: the whole match list is bit-exact against cv2's matcher."""
    rng = np.random.default_rng(0)
    q = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (480, 32), dtype=np.uint8)
    sel = rng.permutation(500)[:300]
    t[:300] = q[sel]
    flip = rng.random((300, 32)) < 0.05
    t[:300] ^= (flip * rng.integers(0, 256, (300, 32))).astype(np.uint8)
    t[5] = t[6]
    got = po.featureMatching_hamming(q, t, 0.5, seed=3)
    draws = po.rand(3, 600)
    for name in ("BruteForce-HammingLUT", "BruteForce-Hamming"):
        knn = cv2.DescriptorMatcher_create(name).knnMatch(q, t, k=2)
        exp, seen = [], set()
        for i, (a, b) in enumerate(knn):
            r = np.float32(a.distance) / np.float32(b.distance)
            if r < 0.5:
                if a.trainIdx in seen:
                    continue
                seen.add(a.trainIdx)
                exp.append((i, a.trainIdx, np.float32(float(r) + float(np.float32(draws[len(exp)])) / (1000.0 * 2147483647))))
        assert len(got) == len(exp) > 100
        assert [(int(m["queryIdx"]), int(m["trainIdx"])) for m in got] == [(e[0], e[1]) for e in exp]
        assert np.array_equal(got["distance"], np.array([e[2] for e in exp], np.float32))


def test_rootsift_vs_cv2_pipeline():
    """squareroot_descriptor_space (src/node.cpp:1823-1837): cv::abs, cv::reduce(CV_REDUCE_SUM, CV_32FC1), sqrt(v / sum).
    cv2 4.13's reduce accumulates the row sequentially in float, as the oracle does: bit-exact."""
    rng = np.random.default_rng(1)
    for dim in (64, 128):
        d = (rng.normal(size=(200, dim)) * rng.choice([0.01, 1.0, 50.0], size=(200, 1))).astype(np.float32)
        d[7] = 0
        a = cv2.absdiff(d, np.zeros_like(d))
        sums = cv2.reduce(a, 1, cv2.REDUCE_SUM, None, cv2.CV_32F)
        exp = a.copy()
        nz = sums[:, 0] != 0
        exp[nz] = np.sqrt(a[nz] / sums[nz])
        assert np.array_equal(po.rootsift(d), exp)
