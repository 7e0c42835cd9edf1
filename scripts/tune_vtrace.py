"""Sweep (S, NSEG) of vtrace_lane_kernel: stand-alone scan at T=100,B=8192 and the fused
V-trace + loss kernel at the c4 / c5 shapes.  CUDA events, L2 flushed before every launch."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ctypes as C  # noqa: E402

from torched_impala_b200 import _cabi  # noqa: E402

lib = _cabi.lib()

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def inputs(T, B, A):
    g = torch.Generator(device="cuda").manual_seed(0)
    return dict(cur=torch.randn(T, B, A, device="cuda", generator=g), beh=torch.randn(T, B, A, device="cuda", generator=g),
                act=torch.randint(0, A, (T, B), device="cuda", generator=g, dtype=torch.int32),
                rew=torch.randn(T, B, device="cuda", generator=g), don=torch.zeros(T, B, dtype=torch.uint8, device="cuda"),
                lens=torch.full((B,), T, dtype=torch.int32, device="cuda"), v=torch.randn(T + 1, B, device="cuda", generator=g))


def time_it(fn, iters=15):
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def main():
    for (T, B, A, loss) in ((100, 8192, 4, False), (100, 8192, 4, True), (20, 4096, 4, True), (20, 4096, 4, False)):
        x = inputs(T, B, A)
        o = dict(vs=torch.empty(T + 1, B, device="cuda"), pg=torch.empty(T, B, device="cuda"),
                 dl=torch.empty(T, B, A, device="cuda"), dv=torch.empty(T + 1, B, device="cuda"),
                 sc=torch.empty(4, dtype=torch.float64, device="cuda"),
                 ws=torch.zeros(int(lib.impala_vtrace_loss_workspace(T, B, A)), dtype=torch.uint8, device="cuda"))
        by_in = 4 * T * B * (2 * A + 2) + T * B + 4 * (T + 1) * B
        by = by_in + 4 * (T + 1) * B + 4 * T * B + (4 * T * B * A + 4 * (T + 1) * B if loss else 0)
        combos = [(2, 8, 1), (2, 13, 4), (2, 7, 8), (2, 4, 8), (2, 16, 4), (5, 5, 4), (5, 4, 8)] if T > 32 else \
                 [(2, 10, 1), (1, 20, 1), (2, 5, 2), (2, 3, 4)]
        for S, nseg, cl in combos:
            if True:
                os.environ["IMPALA_VTRACE_S"], os.environ["IMPALA_VTRACE_NSEG"] = str(S), str(nseg)
                os.environ["IMPALA_VTRACE_CLUSTER"] = str(cl)
                P = lambda t: C.c_void_p(t.data_ptr())
                st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                if loss:
                    fn = lambda: _cabi.check(lib.impala_vtrace_loss(
                        P(x["cur"]), P(x["beh"]), P(x["act"]), P(x["rew"]), P(x["don"]), P(x["lens"]), P(x["v"]), P(o["vs"]),
                        P(o["pg"]), P(o["dl"]), P(o["dv"]), P(o["sc"]), P(o["ws"]), o["ws"].numel(), T, B, A, 0.99, 1.0, 1.0,
                        0.5, 1.0, 6e-4, 1.0 / B, 0, st), "vtrace_loss")
                else:
                    fn = lambda: _cabi.check(lib.impala_vtrace(
                        P(x["cur"]), P(x["beh"]), P(x["act"]), P(x["rew"]), P(x["don"]), P(x["lens"]), P(x["v"]), P(o["vs"]),
                        P(o["pg"]), T, B, A, 0.99, 1.0, 1.0, 0, st), "vtrace")
                us = time_it(fn)
                print(f"T={T} B={B} loss={int(loss)} S={S} warps/CTA={nseg} cluster={cl}: {us:7.2f} us  {by / us / 1e3:7.1f} GB/s", flush=True)
    os.environ.pop("IMPALA_VTRACE_S"), os.environ.pop("IMPALA_VTRACE_NSEG"), os.environ.pop("IMPALA_VTRACE_CLUSTER")


if __name__ == "__main__":
    main()
