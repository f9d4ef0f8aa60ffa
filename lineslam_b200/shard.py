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


def stream_block(step: int, world: int, rank: int, batch: int):
    """ONE stream dealt block-wise: at `step` rank r extracts block g = step * world + r, frames [g * batch, (g + 1) * batch).
    Returns (g, first_frame, prev_owner_rank, prev_owner_step): the head pair (first - 1, first) needs the tail record of
    block g - 1, extracted by rank r - 1 at the same step, or — for rank 0 — by the last rank at the previous step
    (None, None for the very first block). On the GPUs the tails travel with lsl_shift_frame (ncclSend / ncclRecv ring)."""
    g = step * world + rank
    if g == 0:
        return g, 0, None, None
    return g, g * batch, (rank - 1) % world, step if rank > 0 else step - 1


def ring_shift_bytes(payload: bytes, group=None) -> bytes:
    """Host-side stand-in for lsl_shift_frame (any torch.distributed backend): every rank sends `payload` to rank + 1 and
    returns what rank - 1 sent. Collective."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nxt, prv = (rank + 1) % world, (rank - 1) % world
    n_out = torch.tensor([len(payload)], dtype=torch.int64)
    n_in = torch.zeros(1, dtype=torch.int64)
    reqs = [dist.isend(n_out, nxt, group=group), dist.irecv(n_in, prv, group=group)]
    for r in reqs:
        r.wait()
    out = torch.frombuffer(bytearray(payload), dtype=torch.uint8) if payload else torch.zeros(0, dtype=torch.uint8)
    inn = torch.zeros(int(n_in.item()), dtype=torch.uint8)
    reqs = []
    if len(payload):
        reqs.append(dist.isend(out, nxt, group=group))
    if inn.numel():
        reqs.append(dist.irecv(inn, prv, group=group))
    for r in reqs:
        r.wait()
    return inn.numpy().tobytes()


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
