"""CPU model of the barrier protocol of the wide tensor-core backward (csrc/mlp_tcw.cu,
`mlp_bwd_tcw_kernel`): three roles around an in-order asynchronous tensor pipe.

    producer   row-major x tile (xa stage) -> xa_full;  transposed tile + dz (xt stage) -> xt_full
    issuer     UMMA1(t): W1' x xa[t] -> PRE[t & 1], commits d1_full + xa_empty; issued two tiles ahead
               UMMA2(t): DP[t & 1] x xt[t] -> dW1' accumulator, commits xt_empty; `done` once per pass
    epilogue   waits xt_full (dz) and d1_full (PRE), writes DP hi / lo IN PLACE of PRE, arrives dp_full;
               at the end of a pass waits `done`, reads the accumulator out; W1' is restaged per pass

The properties the kernel relies on are protocol properties, independent of CUDA: (1) no deadlock for
any interleaving, (2) every consumer sees the data of exactly the tile it expects - a stage or TMEM
buffer is never overwritten while a reader is still due (the DP_lo buffer has NO barrier of its own:
in-order retirement of the UMMAs protects it), (3) the mbarrier parities keep working when the tile
counter runs on across passes with an odd number of tiles per pass.  Threads with random delays run
the protocol here with tagged resources; a variant that frees the transposed stage one commit too
early (on UMMA1 instead of UMMA2) must be caught, which shows the model sees the hazard.
"""
import queue
import random
import threading
import time

import pytest

STAGES = 3


class MBar:
    """mbarrier with phase parity: wait(p) returns once the phase of parity p has completed."""

    def __init__(self, count):
        self.count, self.arrived, self.phase = count, 0, 0
        self.cv = threading.Condition()

    def arrive(self):
        with self.cv:
            self.arrived += 1
            if self.arrived == self.count:
                self.arrived, self.phase = 0, self.phase + 1
                self.cv.notify_all()

    def wait(self, parity, stop):
        with self.cv:
            while (self.phase & 1) == parity:  # the phase with this parity is the current, incomplete one
                if stop.is_set():
                    raise TimeoutError
                self.cv.wait(0.05)


def run_cta(n_tiles: int, passes: int, seed: int, xt_freed_by: str = "umma2"):
    rng = random.Random(seed)
    stop, errors = threading.Event(), []
    bars = {k: [MBar(1) for _ in range(STAGES)] for k in ("xa_full", "xa_empty", "xt_full", "xt_empty")}
    bars.update(d1_full=[MBar(1), MBar(1)], dp_full=[MBar(1), MBar(1)], done=[MBar(1)])
    res = dict(xa=[None] * STAGES, xt=[None] * STAGES, dz=[None] * STAGES, pre=[None, None], dp_hi=[None, None],
               dp_lo=[None, None], w=None, acc=[])
    pipe: queue.Queue = queue.Queue()  # the tensor pipe: executes in issue order, asynchronously

    def nap(p=0.3):
        if rng.random() < p:
            time.sleep(rng.random() * 2e-4)

    def expect(what, got, want):
        if got != want:
            errors.append((what, got, want))
            stop.set()

    def tensor_pipe():
        while True:
            op = pipe.get()
            if op is None:
                return
            nap(0.5)
            op()

    def producer():
        t = 0
        for hb in range(passes):
            for i in range(n_tiles):
                s, ph = t % STAGES, (t // STAGES) & 1
                bars["xa_empty"][s].wait(ph ^ 1, stop)
                res["xa"][s] = (hb, i)
                nap()
                bars["xa_full"][s].arrive()
                bars["xt_empty"][s].wait(ph ^ 1, stop)
                res["xt"][s] = (hb, i)
                res["dz"][s] = (hb, i)
                nap()
                bars["xt_full"][s].arrive()
                t += 1

    def epilogue():
        t = 0
        for hb in range(passes):
            res["w"] = hb  # begin_pass: W1' of the block into tensor memory (all UMMAs of the last pass retired)
            pass_sync.wait()
            for i in range(n_tiles):
                s, ph, d1, dph = t % STAGES, (t // STAGES) & 1, t & 1, (t >> 1) & 1
                bars["xt_full"][s].wait(ph, stop)
                bars["d1_full"][d1].wait(dph, stop)
                expect("dz", res["dz"][s], (hb, i))
                expect("pre", res["pre"][d1], (hb, i))
                nap()
                res["dp_hi"][d1] = (hb, i)  # in place of PRE
                res["pre"][d1] = None
                res["dp_lo"][d1] = (hb, i)  # no barrier of its own
                bars["dp_full"][d1].arrive()
                t += 1
            bars["done"][0].wait(hb & 1, stop)
            expect("acc", res["acc"], [(hb, i) for i in range(n_tiles)])
            nap()
            pass_sync.wait()  # end_pass

    def issuer():
        t = 0

        def umma1(tt, hb, i):
            s, ph, d1 = tt % STAGES, (tt // STAGES) & 1, tt & 1
            bars["xa_full"][s].wait(ph, stop)

            def op():
                expect("umma1 x", res["xa"][s], (hb, i))
                expect("umma1 w", res["w"], hb)
                res["pre"][d1] = (hb, i)  # overwrites DP_hi of tile tt - 2: UMMA2(tt - 2) was issued before
                res["dp_hi"][d1] = None
                bars["d1_full"][d1].arrive()
                bars["xa_empty"][s].arrive()
                if xt_freed_by == "umma1":  # the broken variant
                    bars["xt_empty"][s].arrive()
            pipe.put(op)

        for hb in range(passes):
            pass_sync.wait()  # begin_pass
            if n_tiles > 0:
                umma1(t, hb, 0)
            if n_tiles > 1:
                umma1(t + 1, hb, 1)
            for i in range(n_tiles):
                s, ph, d1 = t % STAGES, (t // STAGES) & 1, t & 1
                bars["dp_full"][d1].wait((t >> 1) & 1, stop)
                bars["xt_full"][s].wait(ph, stop)

                def op(s=s, d1=d1, hb=hb, i=i):
                    expect("umma2 dp_hi", res["dp_hi"][d1], (hb, i))
                    expect("umma2 dp_lo", res["dp_lo"][d1], (hb, i))
                    expect("umma2 xt", res["xt"][s], (hb, i))
                    if i == 0:
                        res["acc"] = []
                    res["acc"].append((hb, i))
                    if xt_freed_by == "umma2":
                        bars["xt_empty"][s].arrive()
                pipe.put(op)
                nap()
                if i + 2 < n_tiles:
                    umma1(t + 2, hb, i + 2)
                t += 1
            pipe.put(bars["done"][0].arrive)
            pass_sync.wait()  # end_pass

    pass_sync = threading.Barrier(2, timeout=30)  # epilogue + issuer stand for the CTA barrier of a pass boundary

    def guarded(fn):
        def run():
            try:
                fn()
            except (TimeoutError, threading.BrokenBarrierError):
                pass
        return run

    ts = [threading.Thread(target=guarded(f)) for f in (producer, epilogue, issuer)]
    tp = threading.Thread(target=tensor_pipe)
    tp.start()
    for th in ts:
        th.start()
    t_end = time.time() + 30
    for th in ts:
        th.join(timeout=max(0.1, t_end - time.time()))
    hung = any(th.is_alive() for th in ts)
    stop.set()
    pass_sync.abort()
    pipe.put(None)
    tp.join(timeout=5)
    for th in ts:
        th.join(timeout=5)
    if hung and not errors:
        errors.append(("deadlock",))
    return errors


@pytest.mark.parametrize("n_tiles,passes", [(1, 4), (2, 3), (5, 4), (7, 2), (12, 4)])
def test_pipeline_is_deadlock_free_and_hands_over_the_right_tiles(n_tiles, passes):
    for seed in range(4):
        assert run_cta(n_tiles, passes, seed) == []


def test_model_detects_a_stage_freed_too_early():
    """Sanity of the model: freeing the transposed stage when UMMA1 (not UMMA2) retires lets the producer
    overwrite x^T / dz of a tile whose epilogue or UMMA2 is still due."""
    assert any(run_cta(9, 2, seed, xt_freed_by="umma1") for seed in range(12))
