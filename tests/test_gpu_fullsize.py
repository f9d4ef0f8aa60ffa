"""Size-independent properties at the bench's full batch (592 VGA frames, BASELINE config 2): a frame's records do not
depend on its position in the batch or on the batch size, the host-buffer path equals the small-batch path, and a
pair's 128-byte record depends only on the two frames, the ids and the seed — checked bit for bit over the whole batch
(the oracle comparison of the same frames at small sizes is tests/test_gpu_extract.py / test_gpu_pair.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_batch_592_position_and_size_invariance(api):
    from lineslam_b200 import synth
    U, B = 6, 592
    imgs, deps, poses = synth.make_stream(U, scene_seed=2000)
    K = synth.camera_K()
    small = api.Context(max_batch=U, max_w=640, max_h=480)
    ref = small.extract_batch(imgs, deps, K, seeds=list(range(1, U + 1)))
    ref_lines = [f.lines().tobytes() for f in ref]
    assert all(f.num_lines > 50 for f in ref)
    order = np.array([k % U for k in range(B)])
    big = api.Context(max_batch=B, max_w=640, max_h=480)
    frames = big.extract_batch(imgs[order], deps[order], K, seeds=[int(o) + 1 for o in order])
    assert len(frames) == B
    for k in range(B):
        assert frames[k].lines().tobytes() == ref_lines[order[k]], k
    # pairs (k, k-1): the record depends only on (frame k % U, frame (k-1) % U, ids, seed)
    q, t = frames[1:], frames[:-1]
    idq = [int(order[k]) for k in range(1, B)]
    idt = [int(order[k - 1]) for k in range(1, B)]
    seeds = [7 + idq[k] for k in range(B - 1)]
    recs = big.match_pair_batch(q, t, idq, idt, seeds)
    assert int(recs["found"].sum()) >= (B - 1) * (U - 1) // U           # all but the wrap-around pairs register
    first = {}
    for k in range(B - 1):
        key = (idq[k], idt[k])
        if key not in first:
            first[key] = recs[k].tobytes()
        assert recs[k].tobytes() == first[key], k
    assert len(first) == U
    for (a, b), blob in first.items():                                   # and equals a single-pair call on the small context
        one = small.match_pair_batch([ref[a]], [ref[b]], [a], [b], [7 + a])[0]
        assert one.tobytes() == blob, (a, b)
    big.close()
    small.close()
