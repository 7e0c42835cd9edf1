"""Device-resident learner step time for any BASELINE config (c2..c5): CUDA events, L2 flushed."""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torched_impala_b200 import synth  # noqa: E402
from torched_impala_b200.engine import LearnerEngine  # noqa: E402
from torched_impala_b200.utils import default_hparams  # noqa: E402

CFG = {"c4": dict(T=20, B=4096, O=24, A=4, H=256), "c5": dict(T=100, B=8192, O=64, A=4, H=512),
       "c3": dict(T=20, B=1024, O=24, A=4, H=256), "c2": dict(T=20, B=256, O=4, A=2, H=32)}
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c5")
ap.add_argument("--steps", type=int, default=30)
a = ap.parse_args()
w = CFG[a.config]
hp = default_hparams(batch_size=w["B"], max_timesteps=w["T"])
eng = LearnerEngine(w["T"], w["B"], w["O"], w["A"], w["H"], w["H"], hp)
eng.load_state(synth.init_params(0, w["O"], w["A"], w["H"]))
eng.load_device_batch(synth.make_batch(1, w["T"], w["B"], w["O"], w["A"]))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
with torch.cuda.stream(eng.stream):
    for i in range(a.steps + 5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(eng.stream)
        eng.step(0)
        e1.record(eng.stream)
        e1.synchronize()
        if i >= 5:
            ts.append(e0.elapsed_time(e1) * 1e3)
print(f"{a.config} {w}: median {statistics.median(ts):.1f} us/step ({1e6 / statistics.median(ts):.0f} steps/s), "
      f"loss {eng.read_scalars()['total_loss']:.5f}")
