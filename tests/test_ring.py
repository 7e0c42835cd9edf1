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
