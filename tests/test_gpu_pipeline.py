"""lsl_match_pair_batch_begin / _end: the pair stage of one batch on the context's pair stream while the next batch is
extracted on the same context. The records must equal the one-call path bit for bit, the pair workspace is fenced
(LSL_ERR_BUSY) while a batch is in flight, and freeing a participating frame waits for the batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_begin_end_equals_the_one_call_path_while_extracting(api, stream4):
    imgs, deps, poses, K = stream4
    ctx = api.Context(max_batch=4, max_w=640, max_h=480)
    frames = ctx.extract_batch(imgs, deps, K, seeds=[1, 2, 3, 4])
    ids = np.arange(4, dtype=np.int32)
    seeds = np.array([5, 6, 7], np.uint32)
    want = ctx.match_pair_batch(frames[1:], frames[:-1], ids[1:], ids[:-1], seeds)
    m_want = ctx.pair_matches(1, 1)
    assert want["found"].all()
    ctx.match_pair_batch_begin(frames[1:], frames[:-1], ids[1:], ids[:-1], seeds)
    # the next batch is extracted on the same context while the pair stage runs
    nxt = ctx.extract_batch(imgs[::-1].copy(), deps[::-1].copy(), K, seeds=[4, 3, 2, 1])
    with pytest.raises(api.LslError, match="in flight"):
        ctx.match_lines(frames[1], frames[0], True)
    with pytest.raises(api.LslError, match="in flight"):
        ctx.match_pair_batch_begin(frames[1:], frames[:-1], ids[1:], ids[:-1], seeds)
    got = ctx.match_pair_batch_end()
    assert got.tobytes() == want.tobytes()
    assert np.array_equal(ctx.pair_matches(1, 1), m_want)
    # the extraction that ran underneath is the one a quiet context produces
    assert nxt[3].lines().tobytes() == frames[0].lines().tobytes() and nxt[0].lines().tobytes() == frames[3].lines().tobytes()
    # freeing a frame of a batch in flight waits for the pair stream: the records are still complete
    ctx.match_pair_batch_begin(frames[1:], frames[:-1], ids[1:], ids[:-1], seeds)
    frames[3].free()
    assert ctx.match_pair_batch_end().tobytes() == want.tobytes()
    with pytest.raises(api.LslError):
        ctx.match_pair_batch_end()                      # nothing in flight any more
    ctx.close()
