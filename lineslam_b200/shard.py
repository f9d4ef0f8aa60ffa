"""Host-side sharding of the two units the path offers: frames of a stream and (query, train)
pairs of a loop-closure batch. Units are independent, so there is no data-path collective; the only exchange
is the all-gather of fixed-size pose records at graph-insert time (lsl_allgather_poses over NCCL on the GPUs,
`allgather_pose_records` below over whatever torch.distributed backend the host runs, e.g. gloo in tests).
"""
from __future__ import annotations

import numpy as np

from .records import POSE_DTYPE


def shard_stream(n_frames: int, world: int, rank: int, batch: int):
    """Contiguous batches of `batch` frames dealt round-robin to the ranks. Returns a list of (first, last+1)
    frame ranges owned by `rank`; the pair (first-1, first) is registered by the owner of `first`, which fetches
    (or re-extracts) frame first-1."""
    out = []
    nb = (n_frames + batch - 1) // batch
    for b in range(nb):
        if b % world == rank:
            out.append((b * batch, min((b + 1) * batch, n_frames)))
    return out


def shard_pairs(n_pairs: int, world: int, rank: int):
    """Block-wise split of a loop-closure batch (1 query x n keyframes): rank r gets [lo, hi); blocks are padded
    to equal length `per` so that the all-gather moves the same record count from every rank."""
    per = (n_pairs + world - 1) // world
    lo = min(rank * per, n_pairs)
    hi = min(lo + per, n_pairs)
    return lo, hi, per


def pad_records(recs: np.ndarray, per: int) -> np.ndarray:
    """Pads a rank's records to `per` entries; padding carries id_train = id_query = -1, found = 0."""
    out = np.zeros(per, POSE_DTYPE)
    out["id_train"] = -1
    out["id_query"] = -1
    out[:len(recs)] = recs
    return out


def allgather_pose_records(local: np.ndarray, group=None) -> np.ndarray:
    """Host-side all-gather (torch.distributed, any backend) of equal-length POSE_DTYPE arrays, rank-major."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    loc = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).copy())
    outs = [torch.empty_like(loc) for _ in range(world)]
    dist.all_gather(outs, loc, group=group)
    return np.concatenate([o.numpy().view(POSE_DTYPE) for o in outs])


def drop_padding(recs: np.ndarray) -> np.ndarray:
    return recs[recs["id_query"] >= 0]
