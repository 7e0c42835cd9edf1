"""TEST INFRASTRUCTURE - import the unmodified reference modules in this container.

`/root/reference/utils.py:6-7` imports `gym` and `pybullet_envs`, neither of which is
installed; two empty stub modules make `learner`, `models`, `utils`, `actor` import
cleanly.  `device` is resolved at import time
(`learner.py:13`), so CUDA is hidden first.  Only usable where `/root/reference`
exists (the authoring container) - never on the GPU box.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_DIR = os.environ.get("IMPALA_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "learner.py"))


def load(env_factory=None):
    """Returns the reference's (learner, models, utils) modules.

    env_factory, if given, becomes `gym.make` (used to run the unmodified actor.py
    against an in-repo old-API environment).
    """
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_DIR}")
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        gym.Env = object
        sys.modules["gym"] = gym
    if env_factory is not None:
        sys.modules["gym"].make = env_factory
    sys.modules.setdefault("pybullet_envs", types.ModuleType("pybullet_envs"))
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import learner as ref_learner  # noqa: E402
    import models as ref_models  # noqa: E402
    import utils as ref_utils  # noqa: E402

    return ref_learner, ref_models, ref_utils


class ListQueue:
    """Duck-typed stand-in for mp.Queue: `get(timeout=)` pops from a python list."""

    def __init__(self, items):
        self.items = list(items)
        self.pos = 0

    def get(self, timeout=None):
        import queue

        if self.pos >= len(self.items):
            raise queue.Empty
        it = self.items[self.pos]
        self.pos += 1
        return it


class ScalarRecorder:
    """Replaces `learner.SummaryWriter`; records every add_scalar call."""

    instances = []
    hook = None  # called as hook(step) right after each update's total_loss is logged

    def __init__(self, *a, **k):
        self.scalars = []
        ScalarRecorder.instances.append(self)

    def add_text(self, *a, **k):
        pass

    def add_scalar(self, tag, value, step):
        try:
            value = float(value)
        except TypeError:
            value = float(value.item())
        self.scalars.append((tag.split("/")[-1], step, value))
        if tag.endswith("total_loss") and ScalarRecorder.hook is not None:
            ScalarRecorder.hook(step)

    def close(self):
        pass
