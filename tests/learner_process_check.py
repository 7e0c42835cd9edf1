"""Run the drop-in Learner as a real forked process fed through a real mp.Queue.

Executed by test_gpu_learner_process.py in a fresh interpreter (the parent of a forked CUDA
process must never have initialised CUDA, exactly as with the reference's train.py:42).
Feeds the golden c1 batches (CartPole-like, ragged) as reference-format trajectories and
checks the shared-memory policy the actors would read against the real reference's weights.
"""
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from conftest import PKEYS, Golden  # noqa: E402
from torched_impala_b200 import synth  # noqa: E402
from torched_impala_b200.learner import Learner  # noqa: E402
from torched_impala_b200.models import MlpPolicy, MlpValueFn  # noqa: E402
from torched_impala_b200.utils import Counter  # noqa: E402


def main():
    mp.set_start_method("fork", force=True)  # what reference train.py:42 does
    g = Golden("c1_cartpole_ragged")
    c = g.case
    hp = g.hp._replace(max_updates=g.updates, verbose=1, eval_every=None, save_every=2)
    policy, value_fn = MlpPolicy(c["O"], c["A"], c["H_pi"]), MlpValueFn(c["O"], c["H_v"])
    init = g.init_params()
    policy.load_state_dict({k: torch.from_numpy(init["policy"][k]).double() for k in PKEYS})
    value_fn.load_state_dict({k: torch.from_numpy(init["value_fn"][k]).double() for k in PKEYS})
    policy.share_memory()  # train.py:67
    use_ring = len(sys.argv) > 2 and sys.argv[2] == "ring"
    if use_ring:  # same actors, shared-memory slabs instead of pickled tensors
        from torched_impala_b200.ring import RingQueue

        q = RingQueue(c["T"], c["B"], c["O"], c["A"], slabs=2)
    else:
        q = mp.Queue(maxsize=hp.queue_lim)
    counter = Counter(0)
    log_dir = sys.argv[1] if len(sys.argv) > 1 else None
    n_dev = int(sys.argv[3]) if len(sys.argv) > 3 else 1   # > 1: data-parallel learner (dp.py), still ONE Learner
    lrn = Learner(1, hp, policy, value_fn, q, counter, log_path=log_dir, timeout=60,
                  devices=[f"cuda:{i}" for i in range(n_dev)])

    def feed():  # stands in for actor processes: same wire format, same bounded queue
        for u in range(g.updates):
            for tr in synth.to_trajectories(g.batch(u)):
                q.put(tr, timeout=60)

    lrn.start()
    t = threading.Thread(target=feed, daemon=True)
    t.start()
    ok = lrn.completion.wait(timeout=180)
    lrn.join()
    t.join(timeout=5)
    assert ok, "learner never signalled completion"
    assert lrn.p.exitcode == 0, f"learner exit code {lrn.p.exitcode}"
    assert counter.value == g.updates, counter.value
    want = g.params_after(g.updates - 1)["policy"]
    worst = 0.0
    for k in PKEYS:  # the parent (like an actor) sees the update through shared memory
        worst = max(worst, float(np.abs(policy.state_dict()[k].numpy() - want[k]).max()))
    assert worst < 5e-5, worst
    assert not np.allclose(policy.state_dict()[PKEYS[0]].numpy(), init["policy"][PKEYS[0]])
    if log_dir is not None:
        ck = os.path.join(log_dir, "l1", f"IMPALA_{hp.env_name}_l1_2.pt")
        assert os.path.exists(ck), ck
        assert set(torch.load(ck)) == {"policy_state_dict", "value_fn_state_dict"}
    if use_ring:
        q.close()
    assert lrn.policy_version >= 2 and lrn.policy_version % 2 == 0, lrn.policy_version  # publications happened, none torn
    print(f"LEARNER_PROCESS_OK updates={counter.value} max|dW|={worst:.2e} ring={use_ring} devices={n_dev} "
          f"version={lrn.policy_version}")


if __name__ == "__main__":
    main()
