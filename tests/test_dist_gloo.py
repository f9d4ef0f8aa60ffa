"""world_size-2 gloo test of the host-side sharding: pairs / stream batches are dealt to ranks,
every rank registers its own units, pose records are all-gathered; the union equals the unsharded result."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _fake_register(pair_ids):
    """Deterministic stand-in for the device call: the record depends only on the pair id."""
    from lineslam_b200.records import POSE_DTYPE
    recs = np.zeros(len(pair_ids), POSE_DTYPE)
    for k, i in enumerate(pair_ids):
        recs[k]["id_train"], recs[k]["id_query"], recs[k]["found"] = i, i + 1, i % 3 != 0
        recs[k]["tf"] = np.arange(16, dtype=np.float32) + i
        recs[k]["rmse"] = 0.5 * i
    return recs


def _worker(rank, world, port, n_pairs, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lineslam_b200 import shard
    lo, hi, per = shard.shard_pairs(n_pairs, world, rank)
    local = shard.pad_records(_fake_register(list(range(lo, hi))), per)
    allr = shard.drop_padding(shard.allgather_pose_records(local))
    ranges = shard.shard_stream(100, world, rank, 16)
    q.put((rank, allr.tobytes(), ranges))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_pairs_allgather_equals_unsharded():
    from lineslam_b200.records import POSE_DTYPE
    world, n_pairs = 2, 37
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    ps = [ctxm.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = _fake_register(list(range(n_pairs)))
    cover = []
    for rank, blob, ranges in res:
        got = np.frombuffer(blob, POSE_DTYPE)
        assert got.tobytes() == full.tobytes()      # every rank sees all records, in pair order
        cover += ranges
    cover.sort()
    assert cover[0][0] == 0 and cover[-1][1] == 100 and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))


def test_shard_helpers_edge_cases():
    from lineslam_b200 import shard
    assert shard.shard_pairs(0, 4, 2) == (0, 0, 0)
    assert shard.shard_pairs(5, 8, 7) == (5, 5, 1)
    assert [shard.shard_pairs(256, 8, r)[:2] for r in (0, 7)] == [(0, 32), (224, 256)]
    assert shard.shard_stream(10, 3, 1, 4) == [(4, 8)]
    assert len(shard.pad_records(np.zeros(0, shard.POSE_DTYPE), 3)) == 3


def _graph_worker(rank, world, port, q):
    """Loop-closure insertion with the candidates of every frame sharded over the ranks: each rank
    registers its block of candidates, the 128-byte records are all-gathered, every rank commits the same list."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_graph as TG
    from lineslam_b200 import graph as G, shard
    from lineslam_b200.records import POSE_DTYPE
    po, pp = TG._params(**TG.CASES["octomap_5_5_5"]["kw"])
    rec, stamps, feats = TG._script(321, 60, motion=0.15)
    gm = G.GraphManager(pp, 5)
    found = []
    for i in range(len(stamps)):
        rec.cur[0] = i
        action, nid, cmp_ = gm.node_begin(float(stamps[i]), feats[i], feats[i])
        if action in (G.FIRST, G.SKIPPED):
            found.append(action == G.FIRST); continue
        r0 = rec(nid, cmp_) if action == G.COMPARE_PREDECESSOR else None     # the predecessor pair: every rank (cheap)
        action, ids, res = gm.node_predecessor(r0)
        if action == G.DROPPED:
            found.append(bool(res.found_match)); continue
        lo, hi, per = shard.shard_pairs(len(ids), world, rank)
        mine = np.array([rec(nid, int(c)) for c in ids[lo:hi]], POSE_DTYPE) if hi > lo else np.zeros(0, POSE_DTYPE)
        allr = shard.allgather_pose_records(shard.pad_records(mine, per)) if per > 0 else np.zeros(0, POSE_DTYPE)
        allr = shard.drop_padding(allr)
        found.append(bool(gm.node_commit(allr).found_match))
    q.put((rank, found, gm.edges().tobytes(), gm.nodes().tobytes(), [int(k) for k in gm.keyframe_ids()]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_candidates_give_the_unsharded_graph():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_graph as TG
    world = 2
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    ps = [ctxm.Process(target=_graph_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    po, pp = TG._params(**TG.CASES["octomap_5_5_5"]["kw"])
    rec, stamps, feats = TG._script(321, 60, motion=0.15)
    gm, found, _ = TG._run_product(pp, 5, rec, stamps, feats)          # single process, unsharded
    for rank, f, e, n, k in res:
        assert f == found
        assert e == gm.edges().tobytes() and n == gm.nodes().tobytes()
        assert k == [int(x) for x in gm.keyframe_ids()]
    assert len(gm.edges()) > 60


def _stream_worker(rank, world, port, batch, steps, q):
    """One stream dealt block-wise: every rank 'extracts' its block (a record = a function of the frame id), the block tails
    are ring-shifted, the head pair of every block is registered against the received tail."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lineslam_b200 import shard
    pairs, kept = [], None
    for s in range(steps):
        g, first, own_rank, own_step = shard.stream_block(s, world, rank, batch)
        frames = [np.full(3 + (f % 5), f, np.int64) for f in range(first, first + batch)]       # ragged fake line records
        recv = np.frombuffer(shard.ring_shift_bytes(frames[-1].tobytes()), np.int64)            # lsl_shift_frame
        head = recv if rank > 0 else kept
        if own_rank is None:
            assert head is None
        else:
            assert own_rank == (rank - 1) % world and own_step == (s if rank > 0 else s - 1)
            assert len(head) == 3 + ((first - 1) % 5) and head[0] == first - 1                    # the neighbour's tail, intact
            pairs.append((int(head[0]), first))
        pairs += [(f - 1, f) for f in range(first + 1, first + batch)]
        if rank == 0:
            kept = recv                     # the last rank's tail precedes the head of rank 0's next block
    q.put((rank, pairs))
    dist.barrier()
    dist.destroy_process_group()


def test_one_stream_dealt_blockwise_with_shifted_tails_covers_every_pair_once():
    world, batch, steps = 2, 7, 3
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_stream_worker, args=(r, world, port, batch, steps, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    allp = sorted(sum((g[1] for g in got), []))
    n = world * batch * steps
    assert allp == [(f - 1, f) for f in range(1, n)]          # every consecutive pair of the stream exactly once
