"""GPU: first-step parity at the FULL benchmark sizes against the float64 oracle.

BASELINE.json configs[2..4]: c3 (T=20, B=1024, O=24, A=4, H=256), c4 (T=20, B=4096, same nets) and
the c5 shape (T=100, O=64, A=4, H=512; B=1024 here so the numpy oracle stays at a few seconds -
the full B=8192 runs in bench.py --config c5, whose line carries the same `parity` object).
Checked per case: V-trace targets and pg advantages element-wise, the three logged loss scalars
and total loss, the raw (pre-clip) gradient, the clip norms and the parameters after one
clip + Adam step.  Tolerance: ABSOLUTE 1e-5 on vs / pg_adv / scalars (north_star), 5e-5 relative
to the largest entry on gradients, 5e-5 on parameters whose gradient is resolved.
The observed errors are printed (pytest -s / -rP shows them) and asserted.
"""
import json

import pytest
import torch

from oracle.check import first_step_parity
from torched_impala_b200 import synth
from torched_impala_b200.utils import default_hparams

pytestmark = pytest.mark.gpu

CASES = {
    # name: (T, B, O, A, H, ragged)
    "c3": (20, 1024, 24, 4, 256, False),
    "c3_ragged": (20, 1024, 24, 4, 256, True),
    "c4": (20, 4096, 24, 4, 256, False),
    "c5_shape_B1024": (100, 1024, 64, 4, 512, False),
    "c5_shape_ragged_B512": (100, 512, 64, 4, 512, True),
}


@pytest.mark.parametrize("name", list(CASES))
def test_first_step_matches_oracle_at_full_size(name):
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from torched_impala_b200.engine import LearnerEngine

    T, B, O, A, H, ragged = CASES[name]
    hp = default_hparams(batch_size=B, max_timesteps=T, policy_hidden_dims=H, value_fn_hidden_dims=H)
    params = synth.init_params(11, O, A, H)
    batch = synth.make_batch(17, T, B, O, A, ragged=ragged)
    eng = LearnerEngine(T, B, O, A, H, H, hp, use_graph=False)
    par = first_step_parity(eng, params, batch)
    print(name, json.dumps(par))
    assert par["max_abs_vs"] < 1e-5, par
    assert par["max_abs_pg"] < 1e-5, par
    for k, v in par["scalars"].items():
        assert v["abs_err"] < 1e-5, (k, v)      # absolute, as BASELINE.json states it
    assert par["max_rel_grad"] < 5e-5, par
    assert par["max_abs_param_after_1_update"] < 5e-5, par
    assert par["frac_params_off"] < 1e-3, par
    for k in ("norm_policy", "norm_value"):
        assert abs(par[k]["got"] - par[k]["ref"]) < 5e-5 * max(1.0, par[k]["ref"]), par
    assert par["ok"]


def test_graph_replay_equals_eager_at_c4():
    """The captured step (what bench.py times) gives bit-identical parameters to eager launches."""
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from torched_impala_b200.engine import LearnerEngine

    T, B, O, A, H = 20, 4096, 24, 4, 256
    hp = default_hparams(batch_size=B, max_timesteps=T, policy_hidden_dims=H, value_fn_hidden_dims=H)
    params = synth.init_params(3, O, A, H)
    batches = [synth.make_batch(5 + i, T, B, O, A) for i in range(2)]
    out = []
    for graph in (False, True):
        eng = LearnerEngine(T, B, O, A, H, H, hp, use_graph=graph)
        eng.load_state(params)
        for u in range(4):
            eng.fill_host(batches[u % 2], u % 2)
            eng.ingest(u % 2)
            eng.step(u % 2)
        eng.synchronize()
        out.append(eng.params.clone())
    assert torch.equal(out[0], out[1])
