"""CPU restatement of the point-feature half of Node::Node (TEST INFRASTRUCTURE — SURVEY.md §8f row 3).

The reference detects and describes point features with OpenCV's SIFT / SURF (src/node.cpp:219-291: detector->detect,
removeDepthless, KeyPointsFilter::retainBest(max_keypoints), extractor->compute, projectTo3D :952-1018, then
squareroot_descriptor_space :304-310). OpenCV is not under /root/reference; this module restates the published SIFT
algorithm the way OpenCV implements it (modules/features2d/src/sift.dispatch.cpp, sift.simd.hpp: float images,
first octave -1, 3 layers per octave, sigma 1.6, contrast 0.04, edge 10, 36-bin orientation histogram, 4 x 4 x 8
descriptor, fastAtan2) and is pinned against cv2.SIFT_create of the OpenCV in this image in tests/test_oracle_sift.py
(keypoint sets, orientations, descriptors). The CUDA path (csrc/k_sift.cu) is compared with this module stage by stage
and with cv2 end to end at a stated tolerance (Tier-T: float filters are summed in another order).
"""
from __future__ import annotations

import numpy as np

N_LAYERS = 3
SIGMA = 1.6
CONTRAST_THR = 0.04
EDGE_THR = 10.0
IMG_BORDER = 5
MAX_INTERP = 5
ORI_BINS = 36
ORI_SIG_FCTR = 1.5
ORI_RADIUS = 3 * ORI_SIG_FCTR
ORI_PEAK_RATIO = 0.8
DESCR_W = 4
DESCR_BINS = 8
DESCR_SCL = 3.0
DESCR_MAG_THR = 0.2
INT_DESCR_FCTR = 512.0


def gaussian_kernel(sigma: float) -> np.ndarray:
    """cv::getGaussianKernel(ksize, sigma, CV_32F) with ksize = cvRound(sigma * 8 + 1) | 1 (GaussianBlur, float input)."""
    ksize = int(np.rint(sigma * 8 + 1)) | 1
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-0.5 * x * x / (sigma * sigma))
    return (k / k.sum()).astype(np.float32)


def _reflect101(i, n):
    """cv::borderInterpolate(BORDER_REFLECT_101), any distance outside the image."""
    if n == 1:
        return np.zeros_like(i)
    p = 2 * n - 2
    i = np.mod(i, p)
    return np.where(i >= n, p - i, i)


def gaussian_blur(img: np.ndarray, sigma: float) -> np.ndarray:
    """Separable float filter, BORDER_REFLECT_101, rows then columns (cv::GaussianBlur)."""
    k = gaussian_kernel(sigma)
    r = len(k) // 2
    H, W = img.shape
    xi = _reflect101(np.arange(-r, W + r), W)
    tmp = np.zeros_like(img, dtype=np.float32)
    src = img[:, xi]
    for t in range(len(k)):
        tmp += k[t] * src[:, t:t + W]
    yi = _reflect101(np.arange(-r, H + r), H)
    out = np.zeros_like(img, dtype=np.float32)
    src = tmp[yi, :]
    for t in range(len(k)):
        out += k[t] * src[t:t + H, :]
    return out


def upsample2(img: np.ndarray) -> np.ndarray:
    """cv::resize(..., INTER_LINEAR) by exactly 2: dst x samples src at x / 2 - 0.25 (clamped at the borders)."""
    def axis(a, ax):
        n = a.shape[ax]
        d = np.arange(2 * n)
        f = d * 0.5 - 0.25
        i0 = np.floor(f).astype(int)
        w1 = (f - i0).astype(np.float32)
        i1 = np.clip(i0 + 1, 0, n - 1); i0 = np.clip(i0, 0, n - 1)
        a0 = np.take(a, i0, axis=ax); a1 = np.take(a, i1, axis=ax)
        shp = [1, 1]; shp[ax] = -1
        w1 = w1.reshape(shp)
        return (a0 * (np.float32(1) - w1) + a1 * w1).astype(np.float32)
    return axis(axis(img.astype(np.float32), 1), 0)


def build_pyramid(gray_u8: np.ndarray):
    """createInitialImage + buildGaussianPyramid + buildDoGPyramid. Returns (gauss[o][i], dog[o][i], n_octaves)."""
    base = upsample2(gray_u8.astype(np.float32))
    sig_diff = np.sqrt(max(SIGMA * SIGMA - 0.5 * 0.5 * 4, 0.01))
    base = gaussian_blur(base, float(np.float32(sig_diff)))
    n_oct = int(np.rint(np.log(min(base.shape)) / np.log(2.0) - 2)) + 1   # cvRound(log2(min) - 2) - firstOctave
    sig = [SIGMA]
    k = 2.0 ** (1.0 / N_LAYERS)
    for i in range(1, N_LAYERS + 3):
        sp = k ** (i - 1) * SIGMA
        sig.append(np.sqrt((sp * k) ** 2 - sp ** 2))
    gauss, dog = [], []
    for o in range(n_oct):
        layers = []
        for i in range(N_LAYERS + 3):
            if o == 0 and i == 0:
                layers.append(base)
            elif i == 0:
                prev = gauss[o - 1][N_LAYERS]                                  # resize(Size(cols / 2, rows / 2), INTER_NEAREST)
                layers.append(np.ascontiguousarray(prev[::2, ::2][:prev.shape[0] // 2, :prev.shape[1] // 2]))
            else:
                layers.append(gaussian_blur(layers[i - 1], float(sig[i])))
        gauss.append(layers)
        dog.append([layers[i + 1] - layers[i] for i in range(N_LAYERS + 2)])
    return gauss, dog, n_oct


def fast_atan2(y, x):
    """cv::fastAtan2 (degrees, the 7th-order polynomial of modules/core/src/mathfuncs_core)."""
    y = np.asarray(y, np.float32); x = np.asarray(x, np.float32)
    p1 = np.float32(0.9997878412794807 * (180 / np.pi)); p3 = np.float32(-0.3258083974640975 * (180 / np.pi))
    p5 = np.float32(0.1555786518463281 * (180 / np.pi)); p7 = np.float32(-0.04432655554792128 * (180 / np.pi))
    ax, ay = np.abs(x), np.abs(y)
    eps = np.float32(2.220446049250313e-16)
    big = ax >= ay
    c = np.where(big, ay / (ax + eps), ax / (ay + eps)).astype(np.float32)
    c2 = c * c
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c
    a = np.where(big, a, np.float32(90.0) - a)
    a = np.where(x < 0, np.float32(180.0) - a, a)
    a = np.where(y < 0, np.float32(360.0) - a, a)
    return a.astype(np.float32)


def _solve3(A, b):
    """Matx33f::solve(b, DECOMP_LU): cv::solve's closed form for 3 x 3 (Cramer's rule, products and determinant in double,
    result cast to float); zeros if the determinant is exactly 0."""
    S = A.astype(np.float64); bf = b.astype(np.float64)
    d = (S[0, 0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0]) +
         S[0, 2] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]))
    if d == 0.0:
        return np.zeros(3, np.float32)
    d = 1.0 / d
    t0 = d * (bf[0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (bf[1] * S[2, 2] - S[1, 2] * bf[2]) + S[0, 2] * (bf[1] * S[2, 1] - S[1, 1] * bf[2]))
    t1 = d * (S[0, 0] * (bf[1] * S[2, 2] - S[1, 2] * bf[2]) - bf[0] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0]) + S[0, 2] * (S[1, 0] * bf[2] - bf[1] * S[2, 0]))
    t2 = d * (S[0, 0] * (S[1, 1] * bf[2] - bf[1] * S[2, 1]) - S[0, 1] * (S[1, 0] * bf[2] - bf[1] * S[2, 0]) + bf[0] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]))
    return np.array([t0, t1, t2]).astype(np.float32)


def _adjust(dog_o, layer, r, c):
    """adjustLocalExtrema; returns None or (layer, r, c, xi, xr, xc, contr)."""
    img_scale = np.float32(1.0 / 255)
    ds, ss, cs = img_scale * np.float32(0.5), img_scale, img_scale * np.float32(0.25)
    H, W = dog_o[0].shape
    xi = xr = xc = 0.0
    for it in range(MAX_INTERP):
        img, prv, nxt = dog_o[layer], dog_o[layer - 1], dog_o[layer + 1]
        dD = np.array([(img[r, c + 1] - img[r, c - 1]) * ds, (img[r + 1, c] - img[r - 1, c]) * ds, (nxt[r, c] - prv[r, c]) * ds], np.float32)
        v2 = img[r, c] * 2
        dxx = (img[r, c + 1] + img[r, c - 1] - v2) * ss
        dyy = (img[r + 1, c] + img[r - 1, c] - v2) * ss
        dss = (nxt[r, c] + prv[r, c] - v2) * ss
        dxy = (img[r + 1, c + 1] - img[r + 1, c - 1] - img[r - 1, c + 1] + img[r - 1, c - 1]) * cs
        dxs = (nxt[r, c + 1] - nxt[r, c - 1] - prv[r, c + 1] + prv[r, c - 1]) * cs
        dys = (nxt[r + 1, c] - nxt[r - 1, c] - prv[r + 1, c] + prv[r - 1, c]) * cs
        Hm = np.array([[dxx, dxy, dxs], [dxy, dyy, dys], [dxs, dys, dss]], np.float32)
        X = _solve3(Hm, dD)
        xi, xr, xc = -X[2], -X[1], -X[0]
        if abs(xi) < 0.5 and abs(xr) < 0.5 and abs(xc) < 0.5:
            break
        if max(abs(xi), abs(xr), abs(xc)) > 2147483647 / 3:
            return None
        c += int(np.rint(xc)); r += int(np.rint(xr)); layer += int(np.rint(xi))
        if layer < 1 or layer > N_LAYERS or c < IMG_BORDER or c >= W - IMG_BORDER or r < IMG_BORDER or r >= H - IMG_BORDER:
            return None
    else:
        return None
    img, prv, nxt = dog_o[layer], dog_o[layer - 1], dog_o[layer + 1]
    dD = np.array([(img[r, c + 1] - img[r, c - 1]) * ds, (img[r + 1, c] - img[r - 1, c]) * ds, (nxt[r, c] - prv[r, c]) * ds], np.float32)
    t = dD[0] * xc + dD[1] * xr + dD[2] * xi
    contr = img[r, c] * img_scale + t * np.float32(0.5)
    if abs(contr) * N_LAYERS < CONTRAST_THR:
        return None
    v2 = img[r, c] * 2
    dxx = (img[r, c + 1] + img[r, c - 1] - v2) * ss
    dyy = (img[r + 1, c] + img[r - 1, c] - v2) * ss
    dxy = (img[r + 1, c + 1] - img[r + 1, c - 1] - img[r - 1, c + 1] + img[r - 1, c - 1]) * cs
    tr, det = dxx + dyy, dxx * dyy - dxy * dxy
    if det <= 0 or tr * tr * EDGE_THR >= (EDGE_THR + 1) ** 2 * det:
        return None
    return layer, r, c, float(xi), float(xr), float(xc), float(contr)


def _ori_hist(img, px, py, radius, sigma):
    n = ORI_BINS
    H, W = img.shape
    ii, jj = np.mgrid[-radius:radius + 1, -radius:radius + 1]
    y, x = py + ii, px + jj
    ok = (y > 0) & (y < H - 1) & (x > 0) & (x < W - 1)
    y, x, ii, jj = y[ok], x[ok], ii[ok], jj[ok]
    dx = img[y, x + 1] - img[y, x - 1]
    dy = img[y - 1, x] - img[y + 1, x]
    w = np.exp((ii * ii + jj * jj).astype(np.float32) * np.float32(-1.0 / (2.0 * sigma * sigma))).astype(np.float32)
    ori = fast_atan2(dy, dx)
    mag = np.sqrt(dx * dx + dy * dy).astype(np.float32)
    b = np.rint(np.float32(n / 360.0) * ori).astype(int)
    b = np.where(b >= n, b - n, b); b = np.where(b < 0, b + n, b)
    tmp = np.zeros(n, np.float32)
    np.add.at(tmp, b, w * mag)
    t = np.concatenate([tmp[-2:], tmp, tmp[:2]])
    hist = (t[:-4] + t[4:]) * np.float32(1 / 16) + (t[1:-3] + t[3:-1]) * np.float32(4 / 16) + t[2:-2] * np.float32(6 / 16)
    return hist.astype(np.float32)


def detect(gray_u8: np.ndarray, pyr=None):
    """findScaleSpaceExtrema. Returns a list of dicts (x, y in INPUT image pixels, size, angle, response, octave, layer)
    before duplicate removal / retainBest, in OpenCV's scan order (octave, layer, row, column, histogram bin)."""
    gauss, dog, n_oct = pyr if pyr is not None else build_pyramid(gray_u8)
    thr = int(np.floor(0.5 * CONTRAST_THR / N_LAYERS * 255))
    kps = []
    for o in range(n_oct):
        H, W = dog[o][0].shape
        if H <= 2 * IMG_BORDER or W <= 2 * IMG_BORDER:
            continue
        stack = np.stack(dog[o])
        for i in range(1, N_LAYERS + 1):
            cur = stack[i]
            nb = np.stack([stack[i + di][IMG_BORDER + dy:H - IMG_BORDER + dy, IMG_BORDER + dx:W - IMG_BORDER + dx]
                           for di in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)])
            v = cur[IMG_BORDER:H - IMG_BORDER, IMG_BORDER:W - IMG_BORDER]
            is_ext = (np.abs(v) > thr) & (((v > 0) & (v >= nb.max(0))) | ((v < 0) & (v <= nb.min(0))))
            for r0, c0 in zip(*np.nonzero(is_ext)):
                res = _adjust(dog[o], i, int(r0) + IMG_BORDER, int(c0) + IMG_BORDER)
                if res is None:
                    continue
                layer, r, c, xi, xr, xc, contr = res
                size = SIGMA * 2.0 ** ((layer + xi) / N_LAYERS) * (1 << o) * 2
                scl_octv = np.float32(size) * np.float32(0.5) / (1 << o)
                hist = _ori_hist(gauss[o][layer], c, r, int(np.rint(ORI_RADIUS * scl_octv)), float(ORI_SIG_FCTR * scl_octv))
                omax = hist.max()
                mag_thr = np.float32(omax * ORI_PEAK_RATIO)
                n = ORI_BINS
                for j in range(n):
                    l, r2 = (j - 1) % n, (j + 1) % n
                    if hist[j] > hist[l] and hist[j] > hist[r2] and hist[j] >= mag_thr:
                        b = j + 0.5 * (hist[l] - hist[r2]) / (hist[l] - 2 * hist[j] + hist[r2])
                        b = b + n if b < 0 else (b - n if b >= n else b)
                        ang = 360.0 - (360.0 / n) * b
                        if abs(ang - 360.0) < 1.1920929e-07:
                            ang = 0.0
                        # firstOctave = -1: coordinates and size are halved back to the input image
                        kps.append(dict(x=float(np.float32((c + xc) * (1 << o))) * 0.5, y=float(np.float32((r + xr) * (1 << o))) * 0.5,
                                        size=float(np.float32(size)) * 0.5, angle=float(np.float32(ang)), response=abs(contr),
                                        octave=o - 1, layer=layer, xi=xi))
    return kps


def describe(gauss, kp) -> np.ndarray:
    """calcSIFTDescriptor for one keypoint dict of detect(); 128 floats (integers 0..255)."""
    d, n = DESCR_W, DESCR_BINS
    o = kp["octave"] + 1
    scale = 1.0 / 2.0 ** kp["octave"] if kp["octave"] >= 0 else float(1 << -kp["octave"])
    size = np.float32(kp["size"] * scale)
    ptx, pty = np.float32(kp["x"] * scale), np.float32(kp["y"] * scale)
    img = gauss[o][kp["layer"]]
    H, W = img.shape
    angle = 360.0 - kp["angle"]
    if abs(angle - 360.0) < 1.1920929e-07:
        angle = 0.0
    scl = np.float32(size * np.float32(0.5))
    px, py = int(np.rint(ptx)), int(np.rint(pty))
    cos_t = np.float32(np.cos(np.float32(angle) * np.float32(np.pi / 180))); sin_t = np.float32(np.sin(np.float32(angle) * np.float32(np.pi / 180)))
    bins_per_rad = np.float32(n / 360.0)
    exp_scale = np.float32(-1.0 / (d * d * 0.5))
    hist_width = np.float32(DESCR_SCL * scl)
    radius = int(np.rint(hist_width * 1.4142135623730951 * (d + 1) * 0.5))
    radius = min(radius, int(np.sqrt(float(H) * H + float(W) * W)))
    cos_t /= hist_width; sin_t /= hist_width
    ii, jj = np.mgrid[-radius:radius + 1, -radius:radius + 1]
    c_rot = (jj * cos_t - ii * sin_t).astype(np.float32)
    r_rot = (jj * sin_t + ii * cos_t).astype(np.float32)
    rbin = r_rot + np.float32(d // 2 - 0.5); cbin = c_rot + np.float32(d // 2 - 0.5)
    r, c = py + ii, px + jj
    ok = (rbin > -1) & (rbin < d) & (cbin > -1) & (cbin < d) & (r > 0) & (r < H - 1) & (c > 0) & (c < W - 1)
    r, c, rbin, cbin, c_rot, r_rot = r[ok], c[ok], rbin[ok], cbin[ok], c_rot[ok], r_rot[ok]
    dx = img[r, c + 1] - img[r, c - 1]
    dy = img[r - 1, c] - img[r + 1, c]
    w = np.exp((c_rot * c_rot + r_rot * r_rot) * exp_scale).astype(np.float32)
    ori = fast_atan2(dy, dx)
    mag = np.sqrt(dx * dx + dy * dy).astype(np.float32) * w
    obin = (ori - np.float32(angle)) * bins_per_rad
    r0, c0, o0 = np.floor(rbin).astype(int), np.floor(cbin).astype(int), np.floor(obin).astype(int)
    rb, cb, ob = rbin - r0, cbin - c0, obin - o0
    o0 = np.where(o0 < 0, o0 + n, o0); o0 = np.where(o0 >= n, o0 - n, o0)
    hist = np.zeros((d + 2, d + 2, n + 2), np.float32)
    v_r1 = mag * rb; v_r0 = mag - v_r1
    v_rc11 = v_r1 * cb; v_rc10 = v_r1 - v_rc11
    v_rc01 = v_r0 * cb; v_rc00 = v_r0 - v_rc01
    for (dr, dc, v) in ((0, 0, v_rc00), (0, 1, v_rc01), (1, 0, v_rc10), (1, 1, v_rc11)):
        v1 = v * ob; v0 = v - v1
        np.add.at(hist, (r0 + 1 + dr, c0 + 1 + dc, o0), v0)
        np.add.at(hist, (r0 + 1 + dr, c0 + 1 + dc, o0 + 1), v1)
    hist[:, :, 0] += hist[:, :, n]; hist[:, :, 1] += hist[:, :, n + 1]
    dst = hist[1:d + 1, 1:d + 1, :n].reshape(-1).astype(np.float32)
    nrm2 = float((dst * dst).sum())
    thr = np.float32(np.sqrt(nrm2) * DESCR_MAG_THR)
    dst = np.minimum(dst, thr)
    nrm2 = float((dst * dst).sum())
    f = np.float32(INT_DESCR_FCTR / max(np.sqrt(nrm2), 1.1920929e-07))
    return np.clip(np.rint(dst * f), 0, 255).astype(np.float32)


def remove_duplicates_and_retain_best(kps, nfeatures: int):
    """KeyPointsFilter::removeDuplicatedSorted (same pt, size, angle) then retainBest by response."""
    seen, out = set(), []
    for k in sorted(kps, key=lambda k: (k["x"], k["y"], k["size"], k["angle"])):
        key = (k["x"], k["y"], k["size"], k["angle"])
        if key not in seen:
            seen.add(key); out.append(k)
    if nfeatures and len(out) > nfeatures:
        out.sort(key=lambda k: -k["response"])
        thr = out[nfeatures - 1]["response"]
        out = [k for k in out if k["response"] >= thr]       # retainBest keeps ties of the n-th response
    return out


def project_to_3d(kps, depth: np.ndarray, K: np.ndarray, max_keypoints: int = 600):
    """removeDepthless + projectTo3D (src/node.cpp:101-130, 952-1018): float arithmetic, depth at the rounded pixel,
    NaN depth drops the feature. Returns (indices kept, xyz1 float32 [n, 4])."""
    H, W = depth.shape
    fx, fy = np.float32(1.0 / K[0, 0]), np.float32(1.0 / K[1, 1])
    cx, cy = np.float32(K[0, 2]), np.float32(K[1, 2])
    keep, xyz = [], []
    for i, k in enumerate(kps):
        x, y = np.float32(k["x"]), np.float32(k["y"])
        if not (0 <= x < W and 0 <= y < H):
            continue
        Z = depth[int(np.rint(y)) if np.rint(y) < H else H - 1, int(np.rint(x)) if np.rint(x) < W else W - 1]
        if np.isnan(Z):
            continue
        keep.append(i)
        xyz.append([(x - cx) * Z * fx, (y - cy) * Z * fy, Z, 1.0])
        if len(keep) >= max_keypoints:
            break
    return keep, np.array(xyz, np.float32).reshape(-1, 4)
