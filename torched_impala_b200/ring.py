"""Shared-memory trajectory ring: SURVEY.md section 8(f)-1, the first "next" row after the hot path.

The reference ships every trajectory through `mp.Queue` as ~5T+1 separately pickled tiny tensors
(one file descriptor each, `utils.py:48-77`, `actor.py:116-124`): about 52 ms per trajectory at
the consumer and an fd-exhaustion failure mode.  `RingQueue` keeps the queue
interface the unmodified `actor.py` uses (`q.put(traj, timeout=...)`, `queue.Full`), but the
payload never travels through the pipe:

  * K batch slabs live in ONE `multiprocessing.shared_memory` segment, each in exactly the
    learner's device layout (time-major, float32, `impala_batch_layout` offsets);
  * `put()` runs in the ACTOR process: it takes a free (slab, column) ticket, packs the trajectory
    straight into that column with the same `pack_trajectory` the learner uses, and posts the
    ticket (three small ints) on a control queue;
  * the learner collects tickets until a slab is complete; with a GPU it registers the segment
    with `cudaHostRegister` once and DMAs each finished slab to the device directly - no second
    host copy, no per-tensor pickling.

`train.py` changes one line (`q = RingQueue(...)` instead of `mp.Queue(...)`); `Learner` detects
the ring by its `collect_batch` method and otherwise speaks the reference wire format.
"""
from __future__ import annotations

import queue
import time
from multiprocessing import shared_memory

import numpy as np
import torch.multiprocessing as mp

_FIELDS = (("obs", np.float32), ("beh_logits", np.float32), ("actions", np.int32),
           ("rewards", np.float32), ("done", np.uint8), ("lens", np.int32))


def _layout(T: int, B: int, O: int, A: int):
    """Same 256-byte aligned layout as include/impala_b200.h::impala_batch_layout (pure python so
    actor processes do not need the CUDA library)."""
    sizes = ((T + 1) * B * O * 4, T * B * A * 4, T * B * 4, T * B * 4, T * B, B * 4)
    offs, off = [], 0
    for s in sizes:
        offs.append(off)
        off = (off + s + 255) // 256 * 256
    return offs, off


class RingQueue:
    """Drop-in for the `mp.Queue` between actors and learner, backed by shared-memory batch slabs."""

    def __init__(self, T: int, B: int, O: int, A: int, slabs: int = 3):
        if slabs < 2:
            raise ValueError("need at least two slabs (one filling while one is consumed)")
        self.T, self.B, self.O, self.A, self.K = T, B, O, A, slabs
        self.offsets, self.slab_bytes = _layout(T, B, O, A)
        self.shm = shared_memory.SharedMemory(create=True, size=self.slab_bytes * slabs)
        self._owner = True
        self.free = mp.Queue()    # (slab, column) tickets an actor may fill
        self.ready = mp.Queue()   # (slab, column, reward_sum) tickets that are filled
        for k in range(slabs):
            for b in range(B):
                self.free.put((k, b))
        self._views = None
        self._counts = [0] * slabs
        self._rewards = [0.0] * slabs
        self.ids = [[None] * B for _ in range(slabs)]  # trajectory id per (slab, column), for logs
        self._next = 0

    # ---- pickling: child processes attach to the same segment by name
    def __getstate__(self):
        d = self.__dict__.copy()
        d["_views"] = None
        d["_owner"] = False
        d["shm_name"] = self.shm.name
        del d["shm"]
        return d

    def __setstate__(self, d):
        name = d.pop("shm_name")
        self.__dict__.update(d)
        self.shm = shared_memory.SharedMemory(name=name)

    def views(self, k: int) -> dict:
        """Numpy views of slab k (the six batch tensors, learner layout)."""
        if self._views is None:
            shapes = {"obs": (self.T + 1, self.B, self.O), "beh_logits": (self.T, self.B, self.A),
                      "actions": (self.T, self.B), "rewards": (self.T, self.B), "done": (self.T, self.B),
                      "lens": (self.B,)}
            self._views = []
            for kk in range(self.K):
                base = kk * self.slab_bytes
                v = {}
                for (name, dt), off in zip(_FIELDS, self.offsets):
                    n = int(np.prod(shapes[name]))
                    v[name] = np.ndarray(shapes[name], dtype=dt, buffer=self.shm.buf, offset=base + off)
                    assert v[name].size == n
                self._views.append(v)
        return self._views[k]

    def slab_address(self, k: int) -> int:
        """Address of slab k in THIS process (for cudaHostRegister / impala_ingest)."""
        return np.ndarray((1,), dtype=np.uint8, buffer=self.shm.buf, offset=k * self.slab_bytes).ctypes.data

    # ---- actor side (same call shape as mp.Queue.put used at actor.py:118)
    def put(self, traj, block: bool = True, timeout: float | None = None):
        from .learner import pack_trajectory

        try:
            k, b = self.free.get(block, timeout)
        except queue.Empty:
            raise queue.Full from None
        rsum = pack_trajectory(self.views(k), b, traj, self.T)
        self.ready.put((k, b, rsum, getattr(traj, "id", None)))

    # ---- learner side
    def collect_batch(self, timeout: float | None = None):
        """Blocks until the next slab (in round-robin order) has all B columns; returns
        (slab index, batch-mean reward).  Raises queue.Empty after `timeout` seconds without a
        ticket, like `mp.Queue.get` does for the reference learner (learner.py:91-100)."""
        k = self._next
        while self._counts[k] < self.B:
            kk, bb, rsum, tid = self.ready.get(True, timeout)
            self._counts[kk] += 1
            self._rewards[kk] += rsum / self.B
            self.ids[kk][bb] = tid
        reward, self._counts[k], self._rewards[k] = self._rewards[k], 0, 0.0
        self._next = (k + 1) % self.K
        return k, reward

    def release(self, k: int) -> None:
        """The learner is done with slab k (its DMA has completed): hand its columns out again."""
        for b in range(self.B):
            self.free.put((k, b))

    def close(self):
        try:
            self._views = None
            self.shm.close()
            if self._owner:
                self.shm.unlink()
        except (FileNotFoundError, BufferError):
            pass


def wait_until(pred, timeout: float, poll: float = 0.01) -> bool:
    end = time.time() + timeout
    while time.time() < end:
        if pred():
            return True
        time.sleep(poll)
    return pred()
