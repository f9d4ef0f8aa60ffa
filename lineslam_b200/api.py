"""Host-side mirror of the reference's interface for the line front end, on top of the C ABI
(include/lsl.h -> liblsl_b200.so). Names follow the reference:

  Node(...)                       src/node.h:71-77   (constructor runs detect3DLines, src/node.cpp:214-215)
  Node.lineMatching(other, adj)   src/node.h:288     (src/node.cpp:1619-1694)
  Node.matchNodePair(older)       src/node.h:107     (src/node.cpp:1494-1545)
  getTransform_PtsLines_ransac    src/line/utils.h:147-153 (src/line/motion.cpp:605-849)
  MatchingResult                  src/matching_result.h

There is no CPU path here: if liblsl_b200.so is missing or no CUDA device is present every call
raises. torch is not involved; device buffers may be handed in as raw pointers (see
Context.extract_batch_dev).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass, field

import numpy as np

from .records import LINE_DTYPE, MATCH_DTYPE, POSE_DTYPE, Params, Stats, ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblsl_b200.so")
_LIB = None


class LslError(RuntimeError):
    pass


def lib():
    """Loads liblsl_b200.so (built by __graft_entry__.build() / make -C lineslam_b200/csrc)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise LslError(f"{LIB_PATH} not built: run `make -C lineslam_b200/csrc` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.lsl_strerror.restype = C.c_char_p
        L.lsl_last_error.restype = C.c_char_p
        L.lsl_last_error.argtypes = [C.c_void_p]
        L.lsl_debug_read.restype = C.c_int64
        L.lsl_debug_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        L.lsl_ctx_destroy.argtypes = [C.c_void_p]
        L.lsl_ctx_destroy.restype = None
        L.lsl_frame_free.argtypes = [C.c_void_p]
        L.lsl_frame_free.restype = None
        L.lsl_frame_num_lines.argtypes = [C.c_void_p]
        L.lsl_frame_clear_lines.argtypes = [C.c_void_p]
        L.lsl_kernel_name.restype = C.c_char_p
        L.lsl_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def default_params() -> Params:
    p = Params()
    lib().lsl_params_default(C.byref(p))
    return p


def _check(rc, ctx=None):
    if rc < 0:
        msg = lib().lsl_strerror(rc).decode()
        if ctx is not None:
            extra = lib().lsl_last_error(ctx).decode()
            if extra:
                msg += ": " + extra
        raise LslError(msg)
    return rc


@dataclass
class MatchingResult:
    """src/matching_result.h: the subset the path fills."""
    all_matches: np.ndarray = field(default_factory=lambda: np.zeros(0, MATCH_DTYPE))        # point matches
    inlier_matches: np.ndarray = field(default_factory=lambda: np.zeros(0, MATCH_DTYPE))
    all_line_matches: np.ndarray = field(default_factory=lambda: np.zeros(0, MATCH_DTYPE))
    inlier_line_matches: np.ndarray = field(default_factory=lambda: np.zeros(0, MATCH_DTYPE))
    ransac_line_inliers: np.ndarray = field(default_factory=lambda: np.zeros(0, MATCH_DTYPE))
    ransac_trafo: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    final_trafo: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    rmse: float = 1e9
    found: bool = False
    id1: int = -1   # edge.id1 = older node, edge.id2 = newer node (src/node.cpp:1535-1537)
    id2: int = -1
    informationMatrix: np.ndarray = field(default_factory=lambda: np.eye(6))


class Context:
    """Owns the device workspace (lsl_ctx). max_batch frames can be extracted per call."""

    def __init__(self, params: Params | None = None, device: int = 0, max_batch: int = 1, max_w: int = 640,
                 max_h: int = 480, debug: bool = False):
        self.params = params if params is not None else default_params()
        self._h = C.c_void_p()
        self._frames = weakref.WeakSet()
        _check(lib().lsl_ctx_create(C.byref(self._h), C.byref(self.params), device, max_batch, max_w, max_h))
        self.max_batch = max_batch
        if debug:
            _check(lib().lsl_ctx_set_debug(self._h, 1))

    def close(self):
        if self._h:
            for f in list(self._frames):
                f.free()
            lib().lsl_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def set_stream(self, cuda_stream: int | None):
        """Run on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); None restores ours."""
        _check(lib().lsl_ctx_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None), self._h)

    # ---- pose exchange (one process per GPU; the id travels over the host's plumbing) ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().lsl_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid: bytes, nranks: int, rank: int):
        _check(lib().lsl_comm_init(self._h, C.c_char_p(uid), nranks, rank), self._h)
        self.nranks, self.rank = nranks, rank

    def allgather_poses(self, local_recs=None, nlocal=None):
        """All ranks' pose records. local_recs None: the records of the last match_pair_batch, gathered from the device."""
        if local_recs is None:
            out = np.zeros(nlocal * self.nranks, POSE_DTYPE)
            _check(lib().lsl_allgather_poses(self._h, None, 0, None, nlocal, ptr(out)), self._h)
            return out
        loc = np.ascontiguousarray(local_recs, POSE_DTYPE)
        out = np.zeros(len(loc) * self.nranks, POSE_DTYPE)
        _check(lib().lsl_allgather_poses(self._h, None, 0, ptr(loc), len(loc), ptr(out)), self._h)
        return out

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- extraction -----------------------------------------------------------------------
    def extract_batch(self, imgs, depths, K, seeds=None, dt=0.0, depth_factor=5000.0):
        """imgs: (n,H,W,3) or (n,H,W) u8 host array; depths: (n,H,W) f32 metres, or uint16 raw sensor / TUM PNG values
        (0 = invalid, metres = v / depth_factor: converted on the device, half the PCIe bytes). Returns [Frame]."""
        imgs = np.ascontiguousarray(imgs, np.uint8)
        u16 = np.asarray(depths).dtype == np.uint16
        depths = np.ascontiguousarray(depths, np.uint16 if u16 else np.float32)
        n, H, W = depths.shape
        ch = 3 if imgs.ndim == 4 else 1
        # per-frame pointers without a Python loop (592 numpy slices cost ~10 ms per call)
        ipa = np.uint64(imgs.ctypes.data) + np.arange(n, dtype=np.uint64) * np.uint64(imgs.strides[0])
        dpa = np.uint64(depths.ctypes.data) + np.arange(n, dtype=np.uint64) * np.uint64(depths.strides[0])
        ip, dp = ptr(ipa), ptr(dpa)
        sd = np.ascontiguousarray(seeds if seeds is not None else np.ones(n), np.uint32)
        Kc = np.ascontiguousarray(K, np.float64)
        out = (C.c_void_p * n)()
        if u16:
            _check(lib().lsl_extract_batch_u16(self._h, n, ip, ch, dp, W, H, ptr(Kc), C.c_double(dt), ptr(sd),
                                               C.c_double(depth_factor), out), self._h)
        else:
            _check(lib().lsl_extract_batch(self._h, n, ip, ch, dp, W, H, ptr(Kc), C.c_double(dt), ptr(sd), out), self._h)
        return [Frame(self, out[i]) for i in range(n)]

    def extract_batch_dev(self, d_imgs: int, channels: int, d_depths: int, n: int, W: int, H: int, K, seeds=None,
                          dt=0.0):
        """Same with inputs already in device memory (raw device pointers, e.g. tensor.data_ptr())."""
        sd = np.ascontiguousarray(seeds if seeds is not None else np.ones(n), np.uint32)
        Kc = np.ascontiguousarray(K, np.float64)
        out = (C.c_void_p * n)()
        _check(lib().lsl_extract_batch_dev(self._h, n, C.c_void_p(d_imgs), channels, C.c_void_p(d_depths), W, H,
                                           ptr(Kc), C.c_double(dt), ptr(sd), out), self._h)
        return [Frame(self, out[i]) for i in range(n)]

    def frame_from_lines(self, recs):
        recs = np.ascontiguousarray(recs, LINE_DTYPE)
        h = C.c_void_p()
        _check(lib().lsl_frame_from_lines(self._h, ptr(recs), len(recs), C.byref(h)), self._h)
        return Frame(self, h)

    # ---- pair registration ----------------------------------------------------------------
    def match_lines(self, query: "Frame", train: "Frame", adjacent: bool):
        cap = max(query.num_lines, 1)
        out = np.zeros(cap, MATCH_DTYPE)
        n = C.c_int(0)
        _check(lib().lsl_match_lines(self._h, query._h, train._h, int(adjacent), ptr(out), cap, C.byref(n)), self._h)
        return out[:n.value].copy()

    def match_points(self, query: "Frame", train: "Frame", seed: int = 1):
        """Node::featureMatching (BRUTEFORCE) on the frames' point features."""
        cap = max(query.num_points, 1)
        out = np.zeros(cap, MATCH_DTYPE)
        n = C.c_int(0)
        _check(lib().lsl_match_points(self._h, query._h, train._h, C.c_uint32(seed), ptr(out), cap, C.byref(n)), self._h)
        return out[:n.value].copy()

    def bcast_frame(self, frame, root: int = 0):
        """ncclBroadcast of the query frame's line records from `root` (loop-closure batches sharded over ranks)."""
        out = C.c_void_p()
        _check(lib().lsl_bcast_frame(self._h, root, frame._h if frame is not None else None, C.byref(out)), self._h)
        if frame is not None and out.value == frame._h.value:
            return frame
        return Frame(self, out)

    def compute_inliers_and_error(self, query: "Frame", train: "Frame", matches, tf, squared_max_inlier_dist: float):
        """Node::computeInliersAndError (src/node.cpp:1019-1080): (inlier matches in order, rmse Mahalanobis distance)."""
        m = np.ascontiguousarray(matches, MATCH_DTYPE)
        out = np.zeros(max(len(m), 1), MATCH_DTYPE)
        t = np.ascontiguousarray(tf, np.float32).reshape(16)
        k = C.c_int(0); rmse = C.c_double(0.0)
        _check(lib().lsl_compute_inliers_and_error(self._h, query._h, train._h, ptr(m), len(m), ptr(t), C.c_double(squared_max_inlier_dist),
                                                   ptr(out), len(out), C.byref(k), C.byref(rmse)), self._h)
        return out[:k.value].copy(), rmse.value

    def shift_frame(self, frame):
        """Ring shift of the block tails of one stream split over the ranks: sends `frame` to rank + 1, returns the frame
        received from rank - 1 (ncclSend / ncclRecv, device to device). Collective."""
        out = C.c_void_p()
        _check(lib().lsl_shift_frame(self._h, frame._h, C.byref(out)), self._h)
        return Frame(self, out)

    def set_point_detector(self, kind: str = "SIFT", max_keypoints: int = 600, root_sift: bool = True):
        """Point detector run by every extract call (the other half of Node::Node): 'SIFT' or None."""
        _check(lib().lsl_ctx_set_point_detector(self._h, 1 if kind else 0, max_keypoints, int(root_sift)), self._h)

    def match_tc_stats(self, reset: bool = True):
        """(exact evaluations, rows rescanned in full, rows) of the tensor-core pre-filtered point matcher."""
        out = np.zeros(3, np.int64)
        _check(lib().lsl_match_tc_stats(self._h, ptr(out), int(reset)), self._h)
        return tuple(int(v) for v in out)

    def set_camera(self, fx: float, dt: float = 0.0):
        _check(lib().lsl_ctx_set_camera(self._h, C.c_double(fx), C.c_double(dt)), self._h)

    def pose_ransac(self, train: "Frame", query: "Frame", ln_matches, id_train=0, id_query=1, seed=1, pt_matches=None):
        """getTransform_PtsLines_ransac. Point inlier lists of a hybrid call: pair_matches(0, 4) / (0, 5)."""
        m = np.ascontiguousarray(ln_matches if ln_matches is not None else np.zeros(0, MATCH_DTYPE), MATCH_DTYPE)
        pm = np.ascontiguousarray(pt_matches if pt_matches is not None else np.zeros(0, MATCH_DTYPE), MATCH_DTYPE)
        rec = np.zeros(1, POSE_DTYPE)
        cap = max(len(m), 1)
        inl = np.zeros(cap, MATCH_DTYPE)
        rinl = np.zeros(cap, MATCH_DTYPE)
        n1, n2 = C.c_int(0), C.c_int(0)
        _check(lib().lsl_pose_ransac(self._h, train._h, query._h, id_train, id_query, ptr(pm) if len(pm) else None, len(pm),
                                     ptr(m) if len(m) else None, len(m), C.c_uint32(seed), ptr(rec), ptr(inl), cap,
                                     C.byref(n1), ptr(rinl), cap, C.byref(n2)), self._h)
        return rec[0].copy(), inl[:n1.value].copy(), rinl[:n2.value].copy()

    def relmotion_ransac(self, train: "Frame", query: "Frame", ln_matches, seed=1):
        """computeRelativeMotion_Ransac + optimizeRelmotion (src/line/motion.cpp:367-526) on the matched line pairs."""
        m = np.ascontiguousarray(ln_matches, MATCH_DTYPE)
        R = np.zeros(9); t = np.zeros(3); con = np.zeros(max(len(m), 1), np.int32)
        n, calls, have = C.c_int(0), C.c_int(0), C.c_int(0)
        _check(lib().lsl_relmotion_ransac(self._h, train._h, query._h, ptr(m) if len(m) else None, len(m), C.c_uint32(seed),
                                          ptr(R), ptr(t), ptr(con), len(con), C.byref(n), C.byref(calls), C.byref(have)), self._h)
        return dict(R=R.reshape(3, 3), t=t, conset=con[:n.value].copy(), lm_calls=calls.value, have=bool(have.value))

    def set_points_batch(self, frames, xyz1_list, desc_list, root_sift: bool = False):
        """lsl_frames_set_points_batch: point features of many frames in one device block / two uploads."""
        n = len(frames)
        counts = np.array([len(x) for x in xyz1_list], np.int32)
        u8 = np.asarray(desc_list[0]).dtype == np.uint8
        x = np.ascontiguousarray(np.concatenate([np.asarray(a, np.float32).reshape(-1, 4) for a in xyz1_list]), np.float32)
        d = np.ascontiguousarray(np.concatenate([np.asarray(a).reshape(len(a), -1) for a in desc_list]), np.uint8 if u8 else np.float32)
        fr = (C.c_void_p * n)(*[f._h.value for f in frames])
        _check(lib().lsl_frames_set_points_batch(self._h, n, fr, ptr(x), ptr(d), ptr(counts), d.shape[1], int(u8), int(root_sift)), self._h)
        for f in frames:
            f._pdim, f._pu8 = d.shape[1], bool(u8)

    def relmotion_batch(self, npairs: int):
        """lsl_relmotion_batch on the pairs of the last match_pair_batch: (R [n,3,3], t [n,3], info [n,4])."""
        Rt = np.zeros((npairs, 12)); info = np.zeros((npairs, 4), np.int32)
        _check(lib().lsl_relmotion_batch(self._h, npairs, ptr(Rt), ptr(info)), self._h)
        return Rt[:, :9].reshape(npairs, 3, 3), Rt[:, 9:], info

    def match_pair_batch(self, queries, trains, id_query, id_train, seeds):
        """Node::matchNodePair for a batch of independent pairs (graph_manager.cpp:555). Returns POSE_DTYPE[n]."""
        n = len(queries)
        q = (C.c_void_p * n)(*[f._h.value for f in queries])
        t = (C.c_void_p * n)(*[f._h.value for f in trains])
        iq = np.ascontiguousarray(id_query, np.int32)
        it = np.ascontiguousarray(id_train, np.int32)
        sd = np.ascontiguousarray(seeds, np.uint32)
        out = np.zeros(n, POSE_DTYPE)
        _check(lib().lsl_match_pair_batch(self._h, n, q, t, ptr(iq), ptr(it), ptr(sd), ptr(out)), self._h)
        return out

    def match_pair_batch_begin(self, queries, trains, id_query, id_train, seeds):
        """First half of match_pair_batch: enqueues the batch on the context's pair stream and returns at once, so that
        the extraction of the next batch can run underneath it. The frames must stay alive until match_pair_batch_end."""
        n = len(queries)
        qa = (C.c_void_p * n)(*[q._h.value for q in queries])
        ta = (C.c_void_p * n)(*[t._h.value for t in trains])
        iq = np.ascontiguousarray(id_query, np.int32); it = np.ascontiguousarray(id_train, np.int32)
        sd = np.ascontiguousarray(seeds, np.uint32)
        _check(lib().lsl_match_pair_batch_begin(self._h, n, qa, ta, ptr(iq), ptr(it), ptr(sd)), self._h)
        self._pending_pairs = n

    def match_pair_batch_end(self):
        """Second half: waits for the pair stream and returns POSE_DTYPE[n]."""
        n = getattr(self, "_pending_pairs", 0)
        out = np.zeros(max(n, 1), POSE_DTYPE)
        _check(lib().lsl_match_pair_batch_end(self._h, ptr(out), max(n, 1)), self._h)
        self._pending_pairs = 0
        return out[:n]

    def pair_matches(self, pair: int, what: int):
        """Lists of the last pair call: what 0 = all line matches, 1 = refined line inliers, 2 = line inliers of the
        best RANSAC hypothesis; 3, 4, 5 = the same lists for point matches."""
        k = C.c_int(0)
        lib().lsl_pair_matches(self._h, pair, what, None, 0, C.byref(k))
        out = np.zeros(max(k.value, 1), MATCH_DTYPE)
        _check(lib().lsl_pair_matches(self._h, pair, what, ptr(out), len(out), C.byref(k)), self._h)
        return out[:k.value].copy()

    # ---- introspection --------------------------------------------------------------------
    def stats(self) -> Stats:
        s = Stats()
        _check(lib().lsl_get_stats(self._h, C.byref(s)))
        return s

    def kernel_times(self) -> dict:
        """Device ms per kernel of the last extract call and the last pair call."""
        ms = (C.c_float * 32)()
        n = C.c_int(0)
        _check(lib().lsl_kernel_times(self._h, ms, 32, C.byref(n)))
        return {lib().lsl_kernel_name(i).decode(): float(ms[i]) for i in range(n.value)}

    def last_timing(self):
        a, b = C.c_float(0), C.c_float(0)
        _check(lib().lsl_last_timing(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_read(self, what: int, dtype, count: int):
        buf = np.zeros(count, dtype)
        n = _check(lib().lsl_debug_read(self._h, what, ptr(buf), buf.nbytes), self._h)
        return buf[:n]


class Frame:
    """lsl_frame handle: the `lines` member of a reference Node, resident on the device."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle if isinstance(handle, C.c_void_p) else C.c_void_p(handle)
        ctx._frames.add(self)

    @property
    def num_lines(self) -> int:
        return _check(lib().lsl_frame_num_lines(self._h))

    @property
    def num_points(self) -> int:
        return _check(lib().lsl_frame_num_points(self._h))

    def set_points(self, xyz1, desc, root_sift: bool = False):
        """feature_locations_3d_ (n,4) and feature_descriptors_ (n,dim) of the Node this frame belongs to. uint8 rows are
        ORB descriptors (Hamming matcher); root_sift applies squareroot_descriptor_space on the device (f32 rows)."""
        x = np.ascontiguousarray(xyz1, np.float32).reshape(-1, 4)
        d = np.asarray(desc)
        u8 = d.dtype == np.uint8
        d = np.ascontiguousarray(d, np.uint8 if u8 else np.float32).reshape(len(x), -1) if len(x) else np.zeros((0, 4), np.float32)
        _check(lib().lsl_frame_set_points_ex(self.ctx._h, self._h, ptr(x), ptr(d), len(x), d.shape[1], int(u8), int(root_sift)), self.ctx._h)
        self._pdim, self._pu8 = d.shape[1], bool(u8)
        return self

    def points(self):
        """(xyz1 [n,4], desc [n,dim], kp [n,6] or None) of the frame's point features, read back from the device."""
        n = self.num_points
        xyz = np.zeros((max(n, 1), 4), np.float32); kp = np.zeros((max(n, 1), 6), np.float32)
        k = C.c_int(0)
        rc = lib().lsl_frame_points(self.ctx._h, self._h, ptr(xyz), None, ptr(kp), max(n, 1), C.byref(k))
        have_kp = rc == 0
        if not have_kp:
            _check(lib().lsl_frame_points(self.ctx._h, self._h, ptr(xyz), None, None, max(n, 1), C.byref(k)), self.ctx._h)
        dim = 128 if have_kp else getattr(self, "_pdim", 128)
        desc = np.zeros((max(n, 1), dim), np.float32)
        _check(lib().lsl_frame_points(self.ctx._h, self._h, None, ptr(desc), None, max(n, 1), C.byref(k)), self.ctx._h)
        return xyz[:n], desc[:n], (kp[:n] if have_kp else None)

    def descriptors(self) -> np.ndarray:
        """The frame's descriptor rows as the device holds them (after the optional RootSIFT conditioning)."""
        n, dim, u8 = self.num_points, getattr(self, "_pdim", 0), getattr(self, "_pu8", False)
        out = np.zeros((n, dim), np.uint8 if u8 else np.float32)
        _check(lib().lsl_frame_descriptors(self.ctx._h, self._h, ptr(out), C.c_int64(out.nbytes)), self.ctx._h)
        return out

    def lines(self) -> np.ndarray:
        n = self.num_lines
        out = np.zeros(max(n, 1), LINE_DTYPE)
        k = C.c_int(0)
        _check(lib().lsl_frame_lines(self._h, ptr(out), len(out), C.byref(k)))
        return out[:k.value].copy()

    def segments(self) -> np.ndarray:
        """LSD rows (only for frames of a debug=True context)."""
        k = C.c_int(0)
        lib().lsl_frame_segments(self._h, None, 0, C.byref(k))
        out = np.zeros((max(k.value, 1), 5))
        _check(lib().lsl_frame_segments(self._h, ptr(out), len(out), C.byref(k)))
        return out[:k.value].copy()

    def debug(self):
        n = self.num_lines
        npts = np.zeros(max(n, 1), np.int32)
        idx = np.full((max(n, 1), 101), -1, np.int32)
        seg = np.zeros(max(n, 1), np.int32)
        lm = np.zeros(max(n, 1), np.int32)
        _check(lib().lsl_frame_debug(self._h, ptr(npts), ptr(idx), ptr(seg), ptr(lm)))
        return dict(n_inl=npts[:n], inl_idx=idx[:n], seg_of_line=seg[:n], lm_iters=lm[:n])

    def clear_lines(self):
        """vector<FrameLine>().swap(lines) of the clear_past_point_cloud sweep (src/graph_manager.cpp:845-857)."""
        _check(lib().lsl_frame_clear_lines(self._h))

    def free(self):
        if self._h:
            lib().lsl_frame_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Node:
    """One RGB-D frame (src/node.h). The constructor runs Node::detect3DLines on the device; point features
    (feature_locations_3d_, feature_descriptors_) are handed in by the caller like the detectors' output."""

    def __init__(self, ctx: Context, visual, depth, K, node_id: int = 0, seed: int = 1, frame: Frame | None = None,
                 feature_locations_3d=None, feature_descriptors=None):
        self.ctx = ctx
        self.id_ = node_id
        self.seed = seed
        if frame is None:
            frame = ctx.extract_batch(np.asarray(visual)[None], np.asarray(depth)[None], K, [seed])[0]
        self.frame = frame
        if feature_locations_3d is not None:
            frame.set_points(feature_locations_3d, feature_descriptors)

    @property
    def lines(self) -> np.ndarray:
        return self.frame.lines()

    def lineMatching(self, other: "Node", adjacentFrame: bool) -> np.ndarray:
        return self.ctx.match_lines(self.frame, other.frame, adjacentFrame)

    def featureMatching(self, other: "Node", seed: int | None = None) -> np.ndarray:
        return self.ctx.match_points(self.frame, other.frame, seed if seed is not None else self.seed)

    def matchNodePair(self, older: "Node", seed: int | None = None) -> MatchingResult:
        """src/node.cpp:1494-1545 (USE_LINES): featureMatching + lineMatching -> RANSAC -> edge, one device call."""
        P = self.ctx.params
        mr = MatchingResult()
        sd = seed if seed is not None else self.seed
        rec = self.ctx.match_pair_batch([self.frame], [older.frame], [self.id_], [older.id_], [sd])[0]
        mr.all_line_matches = self.ctx.pair_matches(0, 0)
        mr.all_matches = self.ctx.pair_matches(0, 3)
        if len(mr.all_matches) + len(mr.all_line_matches) * P.line_match_number_weight < P.min_feature_matches:
            return mr
        mr.inlier_line_matches, mr.ransac_line_inliers = self.ctx.pair_matches(0, 1), self.ctx.pair_matches(0, 2)
        mr.inlier_matches = self.ctx.pair_matches(0, 4)
        mr.rmse = float(rec["rmse"])
        mr.found = bool(rec["found"])
        if mr.found:
            mr.final_trafo = rec["tf"].reshape(4, 4).copy()
            mr.ransac_trafo = mr.final_trafo
            mr.id1, mr.id2 = older.id_, self.id_
            w = len(mr.inlier_matches) + len(mr.inlier_line_matches) * P.line_match_number_weight
            mr.informationMatrix = np.eye(6) * (w / (mr.rmse * mr.rmse))
        return mr


def getTransform_PtsLines_ransac(trainNode: Node, queryNode: Node, all_point_matches, all_line_matches, seed: int = 1):
    """src/line/motion.cpp:605-849: (record, output_line_inlier_matches, max_line_inlier_set); the point inlier
    lists of a hybrid call are ctx.pair_matches(0, 4) and (0, 5)."""
    return trainNode.ctx.pose_ransac(trainNode.frame, queryNode.frame, all_line_matches, trainNode.id_,
                                     queryNode.id_, seed, pt_matches=all_point_matches)
