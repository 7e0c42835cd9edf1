"""Hardware check of the 16x256b TMEM fragment layout documented in csrc/tc_common.cuh."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from torched_impala_b200 import _cabi  # noqa: E402

lib = _cabi.lib()
out = torch.zeros(128, 96, device="cuda")
rc = lib.impala_debug_tmem_fragment(C.c_void_p(out.data_ptr()))
assert rc == 0, rc
o = out.cpu().numpy()
bad = 0
tile2 = np.full((128, 16), -1.0)
for tid in range(128):
    w, t = tid // 32, tid % 32
    for half in range(2):          # x4 loads: lanes 32w + 16 half + {t/4, t/4 + 8}
        for k in range(4):
            for lh in range(2):
                for c in range(2):
                    L = 32 * w + 16 * half + t // 4 + 8 * lh
                    col = 8 * k + 2 * (t % 4) + c
                    got = o[tid, 16 * half + 4 * k + 2 * lh + c]
                    bad += got != 100.0 * L + col
        for k in range(2):         # x2 loads at column 16
            for lh in range(2):
                for c in range(2):
                    L = 32 * w + 16 * half + t // 4 + 8 * lh
                    col = 16 + 8 * k + 2 * (t % 4) + c
                    got = o[tid, 32 + 8 * half + 4 * k + 2 * lh + c]
                    bad += got != 100.0 * L + col
                    tile2[L, 8 * k + 2 * (t % 4) + c] = 1000.0 * tid + 8 * half + 4 * k + 2 * lh + c
for L in range(128):
    bad += int((o[L, 48:64] != tile2[L]).sum())
print("mismatches:", bad)
if bad:
    print(o[:8, :48])
print("TMEM_FRAGMENT_OK" if not bad else "TMEM_FRAGMENT_MISMATCH")
