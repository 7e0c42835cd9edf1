"""CPU model of the push all-reduce handshake (csrc/optim.cu, tail of csrc/mlp_bwd_tc.cu).

Protocol: the backward of step s on rank `me` STORES its contribution into slot `me` of parity
s & 1 in EVERY rank's gather buffer, then posts flag[me] = s on every rank; the optimizer of step s
waits until its own flag block shows s for all ranks, then reads its own `world` slots.  The
safety argument - parity-double-buffered slots need only the "ready" flags, no acknowledgement
round trip - is a protocol property, independent of CUDA: rank threads with random delays run it
here and check that every read returns exactly the step it expects (never a slot a fast peer has
already overwritten, never a stale one).  The same model with ONE buffer per rank must fail,
which shows the test can see the hazard the second buffer removes."""
import random
import threading
import time

import pytest


def run_ranks(world: int, steps: int, buffers: int, seed: int):
    gather = [[[0] * world for _ in range(buffers)] for _ in range(world)]  # gather[owner][parity][writer] = step
    flags = [[0] * world for _ in range(world)]          # flags[owner][writer] = last step `writer` posted
    errors, stop = [], threading.Event()

    def rank(me: int):
        rng = random.Random(seed * 131 + me)
        for s in range(1, steps + 1):
            if stop.is_set():
                return
            time.sleep(rng.random() * 2e-4)              # forward / V-trace / backward of step s ...
            for p in range(world):                       # ... whose tail pushes the gradient to everyone
                gather[p][s % buffers][me] = s
            for p in range(world):                       # last CTA out: release the flags
                flags[p][me] = s
            t0 = time.time()
            while any(flags[me][r] < s for r in range(world)):   # optimizer: wait on the LOCAL flag block
                if stop.is_set() or time.time() - t0 > 20:
                    return
                time.sleep(0)
            if rng.random() < 0.3:
                time.sleep(rng.random() * 3e-4)          # a slow reader
            got = list(gather[me][s % buffers])          # local slots, rank order
            if got != [s] * world:
                errors.append((me, s, got))
                stop.set()
                return

    ts = [threading.Thread(target=rank, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=60)
    return errors


@pytest.mark.parametrize("world", [2, 4, 8])
def test_parity_buffers_need_no_acknowledgement(world):
    for seed in range(3):
        assert run_ranks(world, steps=300, buffers=2, seed=seed) == []


def test_single_buffer_is_unsafe_without_acknowledgement():
    """Sanity of the model itself: with one buffer a fast rank overwrites what a slow rank has not
    read yet - the hazard must show up within a few attempts."""
    assert any(run_ranks(4, steps=300, buffers=1, seed=seed) for seed in range(8))
