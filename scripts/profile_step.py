"""Run a few eager learner steps so ncu sees every kernel of the update path.

    ncu ... python scripts/profile_step.py [--config c4|c5] [--steps 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from torched_impala_b200 import synth  # noqa: E402
from torched_impala_b200.engine import LearnerEngine  # noqa: E402
from torched_impala_b200.utils import default_hparams  # noqa: E402

CFG = {"c4": dict(T=20, B=4096, O=24, A=4, H=256), "c5": dict(T=100, B=8192, O=64, A=4, H=512),
       "c3": dict(T=20, B=1024, O=24, A=4, H=256), "c2": dict(T=20, B=256, O=4, A=2, H=32)}

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c4")
ap.add_argument("--steps", type=int, default=2)
args = ap.parse_args()
w = CFG[args.config]
hp = default_hparams(batch_size=w["B"], max_timesteps=w["T"])
eng = LearnerEngine(w["T"], w["B"], w["O"], w["A"], w["H"], w["H"], hp, use_graph=False)
eng.load_state(synth.init_params(0, w["O"], w["A"], w["H"]))
eng.load_device_batch(synth.make_batch(1, w["T"], w["B"], w["O"], w["A"]))
for _ in range(args.steps):
    eng.step()
eng.synchronize()
print("ok", eng.read_scalars())
