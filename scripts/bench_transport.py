"""Host-only microbenchmark of the actor -> learner transport, no GPU needed.

Two writer processes send N trajectories of T steps in the reference wire format
(utils.Trajectory: ~5T+1 small tensors) to one consumer that assembles dense (T, B) batches:

  mp.Queue : the reference transport (`actor.py:118` -> `learner.py:91`); the consumer unpickles
             every tensor (one shared-memory handle each) and packs it into the batch slab
  RingQueue: the writers pack into shared-memory slabs themselves; the consumer polls B flags

Prints the consumer's wall time per trajectory and per batch for both.

    python scripts/bench_transport.py [--T 20] [--B 64] [--batches 4]
"""
import argparse
import os
import queue
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from torched_impala_b200 import synth  # noqa: E402
from torched_impala_b200.learner import pack_trajectory  # noqa: E402
from torched_impala_b200.ring import RingQueue  # noqa: E402

O, A = 24, 4


def writer(q, T, count, seed, start, finish):
    trajs = synth.to_trajectories(synth.make_batch(seed, T, count, O, A))
    start.wait()
    for i, tr in enumerate(trajs):
        tr.id = i
        while True:
            try:
                q.put(tr, timeout=5)
                break
            except queue.Full:  # retry like actor.py:116-124
                continue
    finish.wait()  # tensors are handed over through this process's fd server: stay alive (as actors do)


def run(kind, T, B, batches):
    ctx = mp.get_context("fork")
    n = B * batches
    q = RingQueue(T, B, O, A, slabs=3) if kind == "ring" else ctx.Queue(maxsize=2 * B)
    start, finish = ctx.Event(), ctx.Event()
    ps = [ctx.Process(target=writer, args=(q, T, n // 2, s, start, finish)) for s in (1, 2)]
    for p in ps:
        p.start()
    time.sleep(1.0)  # writers have built their trajectories
    shapes = {"obs": (T + 1, B, O), "beh_logits": (T, B, A), "actions": (T, B), "rewards": (T, B),
              "done": (T, B), "lens": (B,)}
    dts = {"obs": np.float32, "beh_logits": np.float32, "actions": np.int32, "rewards": np.float32,
           "done": np.uint8, "lens": np.int32}
    views = {k: np.zeros(shapes[k], dts[k]) for k in shapes}
    t0 = time.perf_counter()
    start.set()
    busy = 0.0
    for _ in range(batches):
        if kind == "ring":
            k, _ = q.collect_batch(timeout=60)
            t1 = time.perf_counter()
            q.release(k)
            busy += time.perf_counter() - t1
        else:
            for b in range(B):
                tr = q.get(timeout=60)
                t1 = time.perf_counter()
                pack_trajectory(views, b, tr, T)
                del tr
                busy += time.perf_counter() - t1
    wall = time.perf_counter() - t0
    finish.set()
    for p in ps:
        p.join(timeout=30)
    if kind == "ring":
        q.close()
    return wall, busy


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=20)
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--batches", type=int, default=4)
    a = ap.parse_args()
    for kind in ("queue", "ring"):
        wall, busy = run(kind, a.T, a.B, a.batches)
        n = a.B * a.batches
        print(f"{kind:5s}: {n} trajectories (T={a.T}) in {wall:.3f} s wall = {1e3 * wall / n:.2f} ms/trajectory, "
              f"{1e3 * wall / a.batches:.1f} ms per batch of {a.B}; consumer-side work after arrival "
              f"{1e6 * busy / n:.0f} us/trajectory", flush=True)
