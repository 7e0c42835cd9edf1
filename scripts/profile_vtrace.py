"""V-trace kernel under ncu: stand-alone scan at the long-unroll stress shape (default T=100, B=8192,
A=4) or the fused V-trace + loss kernel of the step (`python scripts/profile_vtrace.py 20 4096 loss`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torched_impala_b200 import ops  # noqa: E402
from torched_impala_b200.utils import default_hparams  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
LOSS = len(sys.argv) > 3 and sys.argv[3] == "loss"
A = 4
g = torch.Generator(device="cuda").manual_seed(0)
cur = torch.randn(T, B, A, device="cuda", generator=g)
beh = torch.randn(T, B, A, device="cuda", generator=g)
act = torch.randint(0, A, (T, B), device="cuda", generator=g, dtype=torch.int32)
rew = torch.randn(T, B, device="cuda", generator=g)
don = torch.zeros(T, B, dtype=torch.uint8, device="cuda")
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
v = torch.randn(T + 1, B, device="cuda", generator=g)
hp = default_hparams(batch_size=B)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    if LOSS:
        out = ops.vtrace_loss(cur, beh, act, rew, don, lens, v, hp, 1.0 / B)["vs"]
    else:
        out, pg = ops.vtrace(cur, beh, act, rew, don, lens, v, 0.99, 1.0, 1.0)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
