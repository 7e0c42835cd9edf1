"""CPU: the shared-memory trajectory ring across real processes."""
import queue

import numpy as np
import pytest
import torch.multiprocessing as mp

from torched_impala_b200 import synth
from torched_impala_b200.ring import RingQueue

T, B, O, A = 10, 8, 5, 3


def _writer(ring, first, count, seed):
    batch = synth.make_batch(seed, T, B * 2, O, A, ragged=True)
    trajs = synth.to_trajectories(batch)
    for i in range(first, first + count):
        tr = trajs[i]
        tr.id = 1000 * seed + i
        ring.put(tr, timeout=30)


def test_two_writer_processes_fill_slabs_in_learner_layout():
    ctx = mp.get_context("fork")
    ring = RingQueue(T, B, O, A, slabs=2)
    try:
        ps = [ctx.Process(target=_writer, args=(ring, 0, B, 3)), ctx.Process(target=_writer, args=(ring, B, B, 3))]
        for p in ps:
            p.start()
        batch = synth.make_batch(3, T, B * 2, O, A, ragged=True)
        seen = set()
        for _ in range(2):  # two full slabs = 2B trajectories
            k, reward = ring.collect_batch(timeout=30)
            v = ring.views(k)
            for b in range(B):
                seed, i = divmod(ring.ids[k][b], 1000)
                assert seed == 3 and i not in seen
                seen.add(i)
                for name in ("obs", "beh_logits", "actions", "rewards", "done"):
                    np.testing.assert_array_equal(v[name][:, b], batch[name][:, i], err_msg=name)
                assert v["lens"][b] == batch["lens"][i]
            ring.release(k)
        assert seen == set(range(2 * B))
        for p in ps:
            p.join(timeout=10)
            assert p.exitcode == 0
    finally:
        ring.close()


def _stream_writer(ring, wid, count):
    trajs = synth.to_trajectories(synth.make_batch(10 + wid, T, count, O, A, ragged=True))
    for i, tr in enumerate(trajs):
        tr.id = 1000 * wid + i
        while True:
            try:
                ring.put(tr, timeout=0.05)  # short timeout: exercises the queue.Full retry of actor.py:116-124
                break
            except queue.Full:
                continue


def test_many_generations_lose_and_duplicate_nothing():
    """Three writers, two slabs of four columns, ten generations per slab: every trajectory
    arrives exactly once, intact, whichever writer got which ticket."""
    ctx = mp.get_context("fork")
    Bs, per = 4, 28  # 84 trajectories = 21 batches
    ring = RingQueue(T, Bs, O, A, slabs=2)
    try:
        ps = [ctx.Process(target=_stream_writer, args=(ring, w, per)) for w in range(3)]
        for p in ps:
            p.start()
        want = {w: synth.make_batch(10 + w, T, per, O, A, ragged=True) for w in range(3)}
        seen = set()
        for _ in range(3 * per // Bs):
            k, reward = ring.collect_batch(timeout=30)
            v = ring.views(k)
            total = 0.0
            for b in range(Bs):
                w, i = divmod(ring.ids[k][b], 1000)
                assert (w, i) not in seen
                seen.add((w, i))
                np.testing.assert_array_equal(v["obs"][:, b], want[w]["obs"][:, i])
                np.testing.assert_array_equal(v["actions"][:, b], want[w]["actions"][:, i])
                assert v["lens"][b] == want[w]["lens"][i]
                total += float(want[w]["rewards"][:, i].astype(np.float64).sum())
            assert abs(reward - total / Bs) < 1e-9
            ring.release(k)
        assert len(seen) == 3 * per
        for p in ps:
            p.join(timeout=10)
            assert p.exitcode == 0
    finally:
        ring.close()


def test_full_ring_raises_queue_full_like_mp_queue():
    ring = RingQueue(T, 2, O, A, slabs=2)
    try:
        trajs = synth.to_trajectories(synth.make_batch(1, T, 5, O, A))
        for tr in trajs[:4]:
            ring.put(tr, timeout=1)
        with pytest.raises(queue.Full):
            ring.put(trajs[4], timeout=0.2)  # actor.py:120 catches queue.Full and retries
        k, _ = ring.collect_batch(timeout=1)
        ring.release(k)
        k, _ = ring.collect_batch(timeout=1)
        ring.release(k)
        with pytest.raises(queue.Empty):  # the learner's timeout path (learner.py:91-100)
            ring.collect_batch(timeout=0.2)
    finally:
        ring.close()


def test_layout_matches_c_abi():
    from torched_impala_b200 import _cabi
    from torched_impala_b200.ring import _layout

    for shape in ((20, 4096, 24, 4), (10, 8, 5, 3), (1000, 32, 4, 2)):
        assert _layout(*shape) == tuple(_cabi.batch_layout(*shape))


def test_malformed_trajectory_never_leaves_a_hole():
    """a bad trajectory must be rejected BEFORE a column is taken, so the batch still
    completes with the good ones (mp.Queue had no such failure mode)."""
    ring = RingQueue(T, 2, O, A, slabs=2)
    try:
        trajs = synth.to_trajectories(synth.make_batch(1, T, 3, O, A))
        bad = synth.to_trajectories(synth.make_batch(2, T + 4, 1, O, A))[0]  # longer than the unroll
        ring.put(trajs[0], timeout=1)
        with pytest.raises(ValueError):
            ring.put(bad, timeout=1)
        trajs[1].obs.pop()  # malformed lists
        with pytest.raises(ValueError):
            ring.put(trajs[1], timeout=1)
        ring.put(trajs[2], timeout=1)
        k, _ = ring.collect_batch(timeout=1)  # two good trajectories fill the two columns
        assert ring.views(k)["lens"].tolist() == [T, T]
    finally:
        ring.close()


def _block_writer(ring, wid, blocks, n):
    for i in range(blocks):
        blk = synth.make_batch(100 * wid + i, T, n, O, A, ragged=True)
        while True:
            try:
                ring.put_block(blk, timeout=0.05)
                break
            except queue.Full:
                continue


def test_put_block_from_two_processes():
    """Pre-stacked payload: blocks of n columns, two writer processes, every block
    lands intact in n consecutive columns and the reward sum matches."""
    ctx = mp.get_context("fork")
    Bs, n, blocks = 8, 4, 6
    ring = RingQueue(T, Bs, O, A, slabs=2)
    try:
        ps = [ctx.Process(target=_block_writer, args=(ring, w, blocks, n)) for w in range(2)]
        for p in ps:
            p.start()
        want = {(w, i): synth.make_batch(100 * w + i, T, n, O, A, ragged=True) for w in range(2) for i in range(blocks)}
        seen = 0
        for _ in range(2 * blocks * n // Bs):
            k, reward = ring.collect_batch(timeout=30)
            v = ring.views(k)
            total = 0.0
            for b0 in range(0, Bs, n):
                hit = [key for key, blk in want.items() if np.array_equal(v["obs"][:, b0:b0 + n], blk["obs"])]
                assert len(hit) == 1
                blk = want.pop(hit[0])
                for name in ("beh_logits", "actions", "rewards", "done"):
                    np.testing.assert_array_equal(v[name][:, b0:b0 + n], blk[name])
                np.testing.assert_array_equal(v["lens"][b0:b0 + n], blk["lens"])
                total += float(blk["rewards"].astype(np.float64).sum())
                seen += 1
            assert abs(reward - total / Bs) < 1e-9
            ring.release(k)
        assert seen == 2 * blocks and not want
        for p in ps:
            p.join(timeout=10)
            assert p.exitcode == 0
        with pytest.raises(ValueError):
            ring.put_block(synth.make_batch(0, T, 3, O, A))  # 3 does not divide 8
    finally:
        ring.close()
