"""End-to-end use of the three ABIs on a TUM raw directory (what rgbdslam's `loadRawData` + `GraphManager::addNode`
+ `write_poses_2file` do): PNG files -> frames (lsl_extract_tum_batch) -> graph insertion with one
lsl_match_pair_batch per phase (lsl_graph_add_frame) -> trajectory file in TUM format.

  python tools/run_tum_directory.py <dir with syncidx.txt> [out.txt] [--launch lineslam|default] [--keep-lines]
  python tools/run_tum_directory.py --synthetic 12 [out.txt] # writes a synthetic TUM-shaped directory first
"""
import argparse
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from lineslam_b200 import api, graph, tum


def write_synthetic(dirname, n):
    from PIL import Image
    from lineslam_b200 import synth
    imgs, deps, poses = synth.make_stream(n, scene_seed=2000)
    os.makedirs(os.path.join(dirname, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(dirname, "depth"), exist_ok=True)
    rows = []
    for i in range(n):
        ts = 1305031102.175304 + i / 30.0
        Image.fromarray(np.ascontiguousarray(imgs[i][:, :, ::-1])).save(os.path.join(dirname, "rgb", f"{ts:.6f}.png"))
        z = np.rint(np.nan_to_num(deps[i].astype(np.float64), nan=0.0) * 5000.0).astype(np.uint16)
        Image.fromarray(z).save(os.path.join(dirname, "depth", f"{ts:.6f}.png"))
        rows.append(f"{ts:.6f} rgb/{ts:.6f}.png {ts:.6f} depth/{ts:.6f}.png")
    open(os.path.join(dirname, "syncidx.txt"), "w").write("\n".join(rows) + "\n")
    return poses


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dirname", nargs="?")
    ap.add_argument("out", nargs="?", default="poses.txt")
    ap.add_argument("--synthetic", type=int, default=0)
    ap.add_argument("--launch", default="lineslam", choices=["lineslam", "default"])
    ap.add_argument("--keep-lines", action="store_true", help="clear_past_point_cloud = false (line loop closure stays possible)")
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    if a.synthetic:
        if a.dirname and a.out == "poses.txt":
            a.out = a.dirname            # with --synthetic the only positional argument is the output file
        a.dirname = tempfile.mkdtemp(prefix="tum_synth_")
        write_synthetic(a.dirname, a.synthetic)
    ctx = api.Context(max_batch=a.batch, max_w=640, max_h=480)
    gp = graph.lineslam_launch_params() if a.launch == "lineslam" else graph.default_graph_params()
    if a.keep_lines:
        gp.clear_past_point_cloud = 0
    gm = graph.GraphManager(gp, seed=1, ctx=ctx)
    t0 = time.time()
    n_frames = n_added = 0
    for stamps, frames in tum.load_raw_data(ctx, a.dirname, batch=a.batch):
        for ts, fr in zip(stamps, frames):
            node = api.Node(ctx, None, None, None, node_id=n_frames, frame=fr)
            # line-only odometry: the feature-count gates of addNode are fed with the number of 3D lines
            res = gm.addNode(node, ts, n_feat2d=fr.num_lines, n_feat3d=fr.num_lines, seed=1000 * (n_frames + 1))
            n_added += int(res.in_graph)
            n_frames += 1
    gm.write_poses_2file(a.out)
    dt = time.time() - t0
    e = gm.edges()
    print(f"{n_frames} frames in {dt:.2f} s ({n_frames / dt:.1f} frames/s incl. file reads), {n_added} nodes, {len(e)} edges, "
          f"{len(gm.keyframe_ids())} keyframes -> {a.out}")
    tum.release(ctx)
    gm.close()
    ctx.close()


if __name__ == "__main__":
    main()
