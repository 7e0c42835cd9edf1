"""CPU, world_size 2 over gloo: the data-parallel contract of the learner.

Each rank takes a contiguous B/N slice (synth.shard_batch), scales every local sum by
1/B_GLOBAL, packs [gradient | loss scalars] into one float64 buffer and all-reduces it once;
the result must equal the single-process full-batch update, and both ranks must end with
identical parameters.  The compute stand-in here is the float64 oracle (no GPU in this
container); the GPU twin of this test is test_gpu_parity.py::test_full_size_properties.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.impala_oracle import BatchedLearner
from torched_impala_b200 import synth
from torched_impala_b200.utils import default_hparams

CFG = dict(T=9, B=12, O=5, A=3, H=16)


def _payload(out):
    flat = [g.reshape(-1) for g in out["g_policy"] + out["g_value"]]
    scal = np.array([out["value_fn_loss"], out["policy_loss"], out["policy_entropy"],
                     out["batch_mean_reward"]])
    return np.concatenate(flat + [scal])


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = CFG
    hp = default_hparams(batch_size=c["B"], max_norm=0.5)
    batch = synth.make_batch(4, c["T"], c["B"], c["O"], c["A"], ragged=True)
    lrn = BatchedLearner(synth.init_params(2, c["O"], c["A"], c["H"]), hp)
    out = lrn.forward_backward(synth.shard_batch(batch, rank, world), batch_size=c["B"])
    buf = torch.from_numpy(_payload(out))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)  # the ONE collective of a learner step
    red = buf.numpy()
    sizes = [g.size for g in out["g_policy"] + out["g_value"]]
    parts = np.split(red[:-4], np.cumsum(sizes)[:-1])
    shapes = [g.shape for g in out["g_policy"] + out["g_value"]]
    grads = [p.reshape(s) for p, s in zip(parts, shapes)]
    lrn.apply(grads[:4], grads[4:])  # clip on the REDUCED gradient, then Adam (learner.py:176-183)
    ret[rank] = (red, [p.copy() for p in lrn.pi + lrn.vf])
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_full_batch():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    c = CFG
    hp = default_hparams(batch_size=c["B"], max_norm=0.5)
    batch = synth.make_batch(4, c["T"], c["B"], c["O"], c["A"], ragged=True)
    full = BatchedLearner(synth.init_params(2, c["O"], c["A"], c["H"]), hp)
    out = full.forward_backward(batch)
    want = _payload(out)
    full.apply(out["g_policy"], out["g_value"])
    for r in (0, 1):
        red, params = ret[r]
        np.testing.assert_allclose(red, want, rtol=0, atol=1e-12)
        for got, ref in zip(params, full.pi + full.vf):
            np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)
    for a, b in zip(ret[0][1], ret[1][1]):
        assert np.array_equal(a, b)  # replicas stay bit-identical
