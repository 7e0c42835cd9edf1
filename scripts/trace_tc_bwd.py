"""Timeline of CTA 0 of the tcgen05 backward (IMPALA_TC_TRACE=1)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["IMPALA_TC_TRACE"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

from torched_impala_b200 import _cabi, ops, synth  # noqa: E402

M, O, H, N2 = 86016, 24, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 1
rng = np.random.default_rng(0)
p = synth.init_params(0, O, N2, H)["policy"]
x = torch.from_numpy(rng.standard_normal((M, O), dtype=np.float32)).cuda()
d = torch.from_numpy(rng.standard_normal((M, N2), dtype=np.float32)).cuda()
pp = ops.pack_params(p)
for _ in range(3):
    ops.mlp_backward(x, pp, d, O, H, N2)
torch.cuda.synchronize()
lib = _cabi.lib()
lib._FuncPtr  # noqa: B018
fn = lib.impala_debug_read_trace_unused
fn.restype = C.c_int
buf = (C.c_longlong * (3 * 4096))()
n = fn(buf, 4096)
ev = np.array(buf[: 3 * n], dtype=np.int64).reshape(n, 3)
t0 = ev[:, 2].min()
names = {0: "start", 10: "E.begin", 11: "E.d1_full", 12: "E.math_done", 13: "E.dp_free", 14: "E.dp_full",
         15: "E.done", 20: "P.begin", 21: "P.raw_ok", 22: "P.loaded", 23: "P.empty_ok", 24: "P.full",
         30: "M1.begin", 31: "M1.ready", 32: "M1.issued", 33: "M2.wait", 34: "M2.ready", 35: "M2.issued"}
order = np.argsort(ev[:, 2], kind="stable")
print(f"{n} events")
for k in order:
    e, tile, t = ev[k]
    if tile <= 5 or tile >= 17:
        print(f"{t - t0:8d}  tile {tile:3d}  {names.get(int(e), e)}")
