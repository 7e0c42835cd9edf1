"""Pipeline timeline of CTA 0 of the tcgen05 forward (IMPALA_TC_TRACE=1); cycles since kernel start."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["IMPALA_TC_TRACE"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

from torched_impala_b200 import _cabi, ops, synth  # noqa: E402

PAIR = len(sys.argv) > 1 and sys.argv[1] == "pair"  # c4 pair launch: CTA 0 is a policy CTA
M, O, H, N2 = 86016, 24, 256, 4 if PAIR else (int(sys.argv[1]) if len(sys.argv) > 1 else 1)
rng = np.random.default_rng(0)
p = synth.init_params(0, O, N2, H)["policy"]
x = torch.from_numpy(rng.standard_normal((M, O), dtype=np.float32)).cuda()
pp = ops.pack_params(p)
pv = ops.pack_params(synth.init_params(1, O, 1, H)["policy"])
for _ in range(3):
    if PAIR:
        ops.mlp_forward_pair(x, pp, pv, 81920, M, O, H, H, 4)
    else:
        ops.mlp_forward(x, pp, O, H, N2)
torch.cuda.synchronize()
fn = _cabi.lib().impala_debug_read_trace_fwd
fn.restype = C.c_int
buf = (C.c_longlong * (24 * 16))()
fn(buf, 24 * 16)
t = np.array(buf[:], dtype=np.int64).reshape(24, 16)
t0 = t[0, 15]
names = {0: "E.begin", 1: "E.acc_full", 2: "E.done", 5: "P.begin", 6: "P.loaded", 7: "P.empty_ok", 8: "P.full",
         9: "M.begin", 10: "M.ready", 11: "M.issued"}
print("kernel cycles (CTA 0):", t[1, 15] - t0)
print("tile " + " ".join(f"{n:>11s}" for n in names.values()))
for i in range(24):
    if t[i, :12].any():
        print(f"{i:4d} " + " ".join(f"{(t[i, k] - t0) if t[i, k] else 0:11d}" for k in names))
