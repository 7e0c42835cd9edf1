"""Diagnose the UMMA2 (dW1) output layout of the tcgen05 backward."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import impala_oracle as orc  # noqa: E402
from torched_impala_b200 import ops, synth  # noqa: E402

PKEYS = ops.PKEYS
np.set_printoptions(linewidth=200, precision=3, suppress=True)
for (M, O, H, N2) in [(32, 24, 128, 1), (64, 24, 256, 4)]:
    rng = np.random.default_rng(1)
    p = synth.init_params(3, O, N2, H)["policy"]
    x = rng.standard_normal((M, O), dtype=np.float32)
    dout = rng.standard_normal((M, N2), dtype=np.float32)
    p64 = [p[k].astype(np.float64) for k in PKEYS]
    _, pre = orc.mlp_forward(x.astype(np.float64), *p64)
    want = orc.mlp_backward(x.astype(np.float64), pre, p64[2], dout.astype(np.float64))
    W = np.concatenate([want[0], want[1][:, None]], 1)  # [H, O+1] incl. db1 column
    os.environ["IMPALA_MLP_TC"] = "1"
    flat = ops.mlp_backward(torch.from_numpy(x).cuda(), ops.pack_params(p), torch.from_numpy(dout).cuda(), O, H, N2)
    g = ops.unpack_grad(flat, O, H, N2)
    G = np.concatenate([g[PKEYS[0]], g[PKEYS[1]][:, None]], 1)
    print(f"== M={M} O={O} H={H} N2={N2}: |G|max={np.abs(G).max():.3e} |W|max={np.abs(W).max():.3e} nonzero frac={np.mean(G != 0):.3f}")
    # column matching: for each got column, which want column correlates best
    def norm(a):
        return a / (np.linalg.norm(a, axis=0, keepdims=True) + 1e-30)
    C = norm(G).T @ norm(W)
    print("best want-col for each got-col:", np.argmax(np.abs(C), 1), "corr:", np.round(np.max(np.abs(C), 1), 2))
    R = norm(G.T).T @ norm(W.T)  # row matching
    print("row match (first 16 got rows -> want row):", np.argmax(np.abs(R), 1)[:16], np.round(np.max(np.abs(R), 1)[:16], 2))
    print("G[:4,:8]\n", G[:4, :8], "\nW[:4,:8]\n", W[:4, :8])
    print("ratio G/W median:", np.median(G[W != 0] / W[W != 0]))
