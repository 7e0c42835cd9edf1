"""Stand-alone V-trace at the long-unroll stress shape (T=100, B=8192, A=4) for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torched_impala_b200 import ops  # noqa: E402

T, B, A = 100, 8192, 4
g = torch.Generator(device="cuda").manual_seed(0)
cur = torch.randn(T, B, A, device="cuda", generator=g)
beh = torch.randn(T, B, A, device="cuda", generator=g)
act = torch.randint(0, A, (T, B), device="cuda", generator=g, dtype=torch.int32)
rew = torch.randn(T, B, device="cuda", generator=g)
don = torch.zeros(T, B, dtype=torch.uint8, device="cuda")
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
v = torch.randn(T + 1, B, device="cuda", generator=g)
for _ in range(3):
    vs, pg = ops.vtrace(cur, beh, act, rew, don, lens, v, 0.99, 1.0, 1.0)
torch.cuda.synchronize()
print("ok", float(vs.abs().mean()))
