"""Seeded synthetic RGB-D streams.

A 5 x 4 x 3 m box room with piece-wise constant "posters" on the walls and a few
cuboids (shelves, door frames) standing in it, ray-cast through a pin-hole camera
K = [525, 525, 319.5, 239.5] (src/openni_listener.cpp:1256-1259) or the 1280x960
equivalent. Depth follows the reference's own noise model (depthStdDev,
src/line/utils.cpp:671-687), is quantised to 1/5000 m like a TUM PNG
(src/openni_listener.cpp:1233-1244) and has invalid pixels set to NaN.
Nothing here touches the GPU or the oracle: it only makes inputs.
"""
from __future__ import annotations

import numpy as np

ROOM = np.array([5.0, 4.0, 3.0])  # x, y, z extents (z up)


def camera_K(W: int = 640, H: int = 480) -> np.ndarray:
    s = W / 640.0
    return np.array([[525.0 * s, 0, (W - 1) / 2.0], [0, 525.0 * s, (H - 1) / 2.0], [0, 0, 1.0]])


class Scene:
    """Planar rectangles with constant albedo: 6 walls + posters + cuboids."""

    def __init__(self, seed: int):
        rng = np.random.default_rng(seed)
        self.wall_col = rng.integers(60, 200, size=(6, 3)).astype(np.float32)
        # 40-80 posters per wall (the TUM frames the reference targets carry ~400 LSD segments)
        # poster: wall id, (u0, v0, u1, v1) in wall coordinates, colour
        self.posters = []
        for _ in range(int(rng.integers(40, 81)) * 6):
            w = int(rng.integers(0, 6))
            du, dv = self._wall_extent(w)
            su, sv = rng.uniform(0.12, 0.7), rng.uniform(0.12, 0.7)
            u0, v0 = rng.uniform(0, max(du - su, 0.1)), rng.uniform(0, max(dv - sv, 0.1))
            col = rng.integers(30, 226, size=3).astype(np.float32)
            self.posters.append((w, u0, v0, u0 + su, v0 + sv, col))
        n_box = int(rng.integers(5, 10))
        self.boxes = []
        for _ in range(n_box):
            size = rng.uniform([0.1, 0.1, 0.3], [0.45, 0.45, 1.8])
            lo = rng.uniform([0.2, 0.2, 0.0], ROOM - size - [0.2, 0.2, 0.0])
            lo[2] = 0.0 if rng.random() < 0.7 else lo[2] * 0.3
            # keep the middle of the room free for the camera
            c = lo[:2] + size[:2] / 2
            if np.linalg.norm(c - ROOM[:2] / 2) < 1.5:
                continue
            cols = rng.integers(30, 226, size=(6, 3)).astype(np.float32)
            self.boxes.append((lo, lo + size, cols))

    @staticmethod
    def _wall_extent(w: int):
        ax = w // 2  # walls 0,1: x = 0 / Lx ; 2,3: y ; 4,5: z
        o = [a for a in range(3) if a != ax]
        return ROOM[o[0]], ROOM[o[1]]

    def render(self, R_wc: np.ndarray, t_wc: np.ndarray, W: int, H: int, K: np.ndarray):
        """Returns (albedo float32 HxWx3, z-depth float32 HxW) for camera pose x_w = R x_c + t."""
        u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
        d_c = np.stack([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], np.ones_like(u)], -1)
        d_w = d_c @ R_wc.T
        o = t_wc
        tbest = np.full((H, W), np.inf)
        col = np.zeros((H, W, 3), np.float32)
        # room walls (seen from inside)
        for w in range(6):
            ax, side = w // 2, w % 2
            plane = ROOM[ax] * side
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (plane - o[ax]) / d_w[..., ax]
            ok = (t > 1e-6) & (t < tbest)
            if not ok.any():
                continue
            P = o + t[..., None] * d_w
            oa = [a for a in range(3) if a != ax]
            pu, pv = P[..., oa[0]], P[..., oa[1]]
            ok &= (pu >= -1e-9) & (pu <= ROOM[oa[0]] + 1e-9) & (pv >= -1e-9) & (pv <= ROOM[oa[1]] + 1e-9)
            c = np.broadcast_to(self.wall_col[w], (H, W, 3)).copy()
            for (pw, u0, v0, u1, v1, pc) in self.posters:
                if pw != w:
                    continue
                m = (pu >= u0) & (pu <= u1) & (pv >= v0) & (pv <= v1)
                c[m] = pc
            col[ok] = c[ok]
            tbest = np.where(ok, t, tbest)
        # cuboids (slab test, first hit face)
        for lo, hi, cols in self.boxes:
            with np.errstate(divide="ignore", invalid="ignore"):
                t0 = (lo - o) / d_w
                t1 = (hi - o) / d_w
            tn, tf = np.minimum(t0, t1), np.maximum(t0, t1)
            tnear = tn.max(-1)
            tfar = tf.min(-1)
            hit = (tnear < tfar) & (tnear > 1e-6) & (tnear < tbest)
            if not hit.any():
                continue
            ax = tn.argmax(-1)
            side = (np.take_along_axis(d_w, ax[..., None], -1)[..., 0] < 0).astype(np.int64)
            face = ax * 2 + side
            col[hit] = cols[face[hit]]
            tbest = np.where(hit, tnear, tbest)
        z = (tbest * d_c[..., 2]).astype(np.float32)  # d_c z-component is 1 -> z-depth == t
        return col, z


def trajectory_xyz(i: int):
    """fr1/xyz-shape: translation-only sinusoids, +-0.2 m, ~0.3 m/s peak, 30 Hz (SURVEY §8d cfg 2)."""
    t = i / 30.0
    c = np.array([2.5, 2.0, 1.4])
    p = c + 0.2 * np.array([np.sin(1.5 * t), np.sin(1.1 * t + 0.7), 0.5 * np.sin(0.9 * t + 1.3)])
    # camera looks along +x of the world: columns are the camera axes (x right, y down, z forward)
    R = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    return R, p


def trajectory_orbit(i: int):
    """fr2/desk-shape: slow orbit with yaw toward the room centre (cfg 3/4/5)."""
    t = i / 30.0
    ang = 0.2 / 1.2 * t
    c = np.array([2.5, 2.0, 1.3])
    p = c + 1.2 * np.array([np.cos(ang), np.sin(ang), 0.0]) * 0.6
    fwd = c + np.array([0, 0, -0.2]) - p
    fwd[2] = -0.15
    fwd = -fwd / np.linalg.norm(fwd)  # look outward at the walls (more structure than the centre)
    up = np.array([0, 0, 1.0])
    right = np.cross(fwd, up); right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], 1)
    return R, p


def make_frame(scene: Scene, i: int, seed: int, W: int = 640, H: int = 480, traj=trajectory_xyz):
    """Returns (bgr u8 HxWx3, depth f32 HxW metres with NaN holes, R_wc, t_wc)."""
    rng = np.random.default_rng(seed * 1000003 + i)
    K = camera_K(W, H)
    R, p = traj(i)
    alb, z = scene.render(R, p, W, H, K)
    img = alb + rng.normal(0.0, 2.0, alb.shape).astype(np.float32)
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    zz = z.astype(np.float64)
    sig = 0.00273 * zz * zz + 0.00074 * zz - 0.00058
    zn = zz + rng.normal(0.0, 1.0, zz.shape) * np.maximum(sig, 0.0)
    zn = np.rint(zn * 5000.0) / 5000.0
    depth = zn.astype(np.float32)
    # invalid: 5 % random + a 2-px band at depth discontinuities
    bad = rng.random(zz.shape) < 0.05
    gy = np.abs(np.diff(z, axis=0, prepend=z[:1])) > 0.05
    gx = np.abs(np.diff(z, axis=1, prepend=z[:, :1])) > 0.05
    edge = gx | gy
    band = edge.copy()
    for s in (1, 2):
        band[s:, :] |= edge[:-s, :]; band[:-s, :] |= edge[s:, :]
        band[:, s:] |= edge[:, :-s]; band[:, :-s] |= edge[:, s:]
    depth[bad | band | ~np.isfinite(depth)] = np.nan
    return np.ascontiguousarray(img), np.ascontiguousarray(depth), R, p


def relative_pose_q2t(Rq, pq, Rt, pt):
    """4x4 transform mapping query-camera coordinates to train-camera coordinates."""
    T = np.eye(4)
    T[:3, :3] = Rt.T @ Rq
    T[:3, 3] = Rt.T @ (pq - pt)
    return T


def make_stream(n: int, scene_seed: int = 2000, W: int = 640, H: int = 480, traj=trajectory_xyz, start: int = 0,
                stride: int = 1):
    scene = Scene(scene_seed)
    imgs = np.empty((n, H, W, 3), np.uint8)
    deps = np.empty((n, H, W), np.float32)
    poses = []
    for k in range(n):
        img, d, R, p = make_frame(scene, start + k * stride, scene_seed, W, H, traj)
        imgs[k] = img; deps[k] = d; poses.append((R, p))
    return imgs, deps, poses


# ------------------------------------------------------------------ point features (cfg 3) ----
def make_landmarks(scene_seed: int, n: int = 6000, dim: int = 128):
    """World landmarks on the room walls with a RootSIFT-shaped base descriptor each (non-negative, unit L2).
    Stand-in for SIFT keypoints (point features are an INPUT of the hot path)."""
    rng = np.random.default_rng(scene_seed * 7919 + 13)
    wall = rng.integers(0, 6, n)
    P = rng.uniform(0, 1, (n, 3)) * ROOM
    ax, side = wall // 2, wall % 2
    P[np.arange(n), ax] = ROOM[ax] * side
    base = np.abs(rng.normal(0, 1, (n, dim))).astype(np.float32) ** 2
    base = np.sqrt(base / base.sum(1, keepdims=True)).astype(np.float32)
    return P, base


def make_points(scene_seed: int, frame_index: int, depth: np.ndarray, R_wc: np.ndarray, t_wc: np.ndarray, K: np.ndarray,
                max_keypoints: int = 600, dim: int = 128, desc_noise: float = 0.01, outlier_frac: float = 0.1):
    """Point features of one frame: (xyz1 float32 [n,4] camera coordinates with the frame's noisy depth, NaN z where the
    depth map has a hole; desc float32 [n,dim] RootSIFT-like rows; landmark ids). Deterministic in (scene_seed, frame)."""
    P, base = make_landmarks(scene_seed, dim=dim)
    H, W = depth.shape
    Pc = (P - t_wc) @ R_wc          # x_c = R^T (x_w - t)
    z = Pc[:, 2]
    ok = z > 0.3
    u = K[0, 0] * Pc[:, 0] / np.where(ok, z, 1) + K[0, 2]
    v = K[1, 1] * Pc[:, 1] / np.where(ok, z, 1) + K[1, 2]
    ok &= (u >= 2) & (u <= W - 3) & (v >= 2) & (v <= H - 3)
    idx = np.nonzero(ok)[0]
    ui, vi = np.rint(u[idx]).astype(int), np.rint(v[idx]).astype(int)
    d = depth[vi, ui].astype(np.float64)
    vis = ~np.isfinite(d) | (np.abs(d - z[idx]) < 0.05)     # occluded by a cuboid -> not a feature
    idx, ui, vi, d = idx[vis], ui[vis], vi[vis], d[vis]
    idx, ui, vi, d = idx[:max_keypoints], ui[:max_keypoints], vi[:max_keypoints], d[:max_keypoints]
    rng = np.random.default_rng(scene_seed * 1000003 + 7 * frame_index + 1)
    xyz1 = np.ones((len(idx), 4), np.float32)
    xyz1[:, 0] = ((u[idx] - K[0, 2]) / K[0, 0] * d).astype(np.float32)
    xyz1[:, 1] = ((v[idx] - K[1, 2]) / K[1, 1] * d).astype(np.float32)
    xyz1[:, 2] = d.astype(np.float32)
    desc = base[idx] + rng.normal(0, desc_noise, (len(idx), dim)).astype(np.float32)
    out = rng.random(len(idx)) < outlier_frac               # features whose descriptor matches nothing
    desc[out] = np.abs(rng.normal(0, 1, (int(out.sum()), dim))).astype(np.float32)
    desc = np.abs(desc)
    desc = np.sqrt(desc / desc.sum(1, keepdims=True)).astype(np.float32)
    return np.ascontiguousarray(xyz1), np.ascontiguousarray(desc), idx


def border_bands(seed: int, W: int = 320, H: int = 240) -> np.ndarray:
    """Wide colour bands crossing the image border at shallow angles: LSD rectangles whose axis end points round to a
    pixel outside the image (seeds 88 and 307 at 320x240), the case cv::LineIterator clips (tests only)."""
    import cv2
    rng = np.random.default_rng(seed)
    img = np.full((H, W, 3), int(rng.integers(60, 180)), np.uint8)
    for _ in range(10):
        side = rng.integers(0, 4)
        a = rng.uniform(-0.35, 0.35) + (0 if side < 2 else np.pi / 2)
        if side < 2:
            c = np.array([rng.uniform(0, W), rng.choice([rng.uniform(-10, 25), rng.uniform(H - 25, H + 10)])])
        else:
            c = np.array([rng.choice([rng.uniform(-10, 25), rng.uniform(W - 25, W + 10)]), rng.uniform(0, H)])
        L = rng.uniform(100, 300); wd = rng.uniform(25, 70)
        d = np.array([np.cos(a), np.sin(a)]); n = np.array([-d[1], d[0]])
        poly = np.array([c - d * L / 2 - n * wd / 2, c + d * L / 2 - n * wd / 2, c + d * L / 2 + n * wd / 2, c - d * L / 2 + n * wd / 2])
        cv2.fillPoly(img, [np.round(poly).astype(np.int32)], tuple(int(v) for v in rng.integers(20, 235, 3)))
    return np.clip(img + rng.normal(0, 1.5, img.shape), 0, 255).astype(np.uint8)
