"""TEST INFRASTRUCTURE - first-step parity of a `LearnerEngine` against the float64 oracle.

Used by `tests/` (full-size c3/c4/c5 checks, multi-GPU check), `__graft_entry__.smoke()` and by
`bench.py` as a CHECKER outside every timed region (`"parity"` key of the bench line).  Never on
the product path.

Trajectories are independent units: `vs` / `pg_adv` of a shard only depend on that
shard, the loss scalars and the gradient are sums over trajectories scaled by 1/B_global.  So with
N ranks every rank runs the oracle on ITS shard (with the global batch size) and the per-rank
oracle sums are added with one all-reduce - the result is the oracle of the full batch, and it is
compared with what the engine left in its reduced `[gradient | scalars]` buffer.
"""
from __future__ import annotations

import numpy as np

from .impala_oracle import PKEYS, BatchedLearner, clip_coef

SCALARS = ("value_fn_loss", "policy_loss", "policy_entropy", "batch_mean_reward")
TOL = 1e-5  # BASELINE.json north_star: V-trace targets, advantages, three loss scalars


def _flat_oracle_grad(eng, out):
    """Oracle gradient in the engine's flat `[policy | value_fn]` parameter-block layout."""
    flat = np.zeros(eng.n_total, np.float64)
    per_group = {"policy": out["g_policy"], "value_fn": out["g_value"]}
    for grp, key, off, shp in eng._segments():
        g = per_group[grp][PKEYS.index(key)]
        flat[off:off + g.size] = np.asarray(g, np.float64).reshape(-1)
    return flat


def first_step_parity(eng, params: dict, local_batch: dict, mode: str = "reference", group=None) -> dict:
    """Run ONE eager engine step on `local_batch` from `params` and compare with the oracle.

    Leaves the engine's parameters/optimizer state advanced by that one step (callers that go on
    to time steps do not care; callers that compare later updates reload the state).
    """
    import torch

    hp, world = eng.hp, eng.world
    eng.load_state(params)
    eng.fill_host(local_batch, 0)
    eng.ingest(0)
    eng.step(0)
    sc = eng.read_scalars()
    eng.synchronize()
    vs = eng.vs.detach().cpu().numpy()
    pg = eng.pg_adv.detach().cpu().numpy()
    grad = eng.comm[: eng.n_total].detach().cpu().numpy().copy()   # reduced over ranks, pre-clip
    after = eng.params.detach().cpu().numpy().astype(np.float64)

    orc = BatchedLearner(params, hp)
    out = orc.forward_backward(local_batch, mode=mode, batch_size=eng.global_batch)
    e_vs = float(np.abs(vs - out["vs"]).max())
    e_pg = float(np.abs(pg - out["pg_adv"]).max())
    ref_sc = np.array([out[k] for k in SCALARS], np.float64)
    ref_grad = _flat_oracle_grad(eng, out)
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([e_vs, e_pg], dtype=torch.float64, device=eng.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        e_vs, e_pg = (float(x) for x in t.tolist())
        s = torch.from_numpy(np.concatenate([ref_sc, ref_grad])).to(eng.dev)
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
        s = s.cpu().numpy()
        ref_sc, ref_grad = s[:4], s[4:]
    # clip + Adam of the oracle on the summed gradient (learner.py:176-183)
    g_pi = [ref_grad[off:off + int(np.prod(shp))].reshape(shp) for grp, _, off, shp in eng._segments() if grp == "policy"]
    g_vf = [ref_grad[off:off + int(np.prod(shp))].reshape(shp) for grp, _, off, shp in eng._segments() if grp == "value_fn"]
    orc.apply(g_pi, g_vf)
    st = orc.state()
    want_after = np.zeros(eng.n_total, np.float64)
    for grp, key, off, shp in eng._segments():
        want_after[off:off + int(np.prod(shp))] = st[grp][key].reshape(-1)
    scal = {}
    for i, k in enumerate(SCALARS):
        scal[k] = dict(got=float(sc[k]), ref=float(ref_sc[i]), abs_err=abs(float(sc[k]) - float(ref_sc[i])))
    tot_ref = hp.v_loss_c * ref_sc[0] + hp.policy_loss_c * ref_sc[1] - hp.entropy_c * ref_sc[2]
    scal["total_loss"] = dict(got=float(sc["total_loss"]), ref=float(tot_ref),
                              abs_err=abs(float(sc["total_loss"]) - float(tot_ref)))
    gmax = float(np.abs(ref_grad).max())
    e_grad = float(np.abs(grad - ref_grad).max()) / max(gmax, 1e-30)
    # Adam's first step is lr * g / (|g| + eps'): entries whose gradient is at rounding level may
    # take a different sign - compare where the gradient is resolved
    d_after = np.abs(after - want_after)
    resolved = np.abs(ref_grad) > 1e-3 * gmax
    e_par = float(d_after[resolved].max()) if resolved.any() else 0.0
    n_pi = clip_coef(g_pi, hp.max_norm)[1]
    n_vf = clip_coef(g_vf, hp.max_norm)[1]
    worst_scalar = max(scal[k]["abs_err"] for k in ("value_fn_loss", "policy_loss", "policy_entropy"))
    return dict(max_abs_vs=e_vs, max_abs_pg=e_pg, scalars=scal, max_abs_scalar=worst_scalar,
                max_rel_grad=e_grad, max_abs_param_after_1_update=e_par,
                frac_params_off=float((d_after > 2e-5).mean()),
                norm_policy=dict(got=float(sc["norm_policy"]), ref=float(n_pi)),
                norm_value=dict(got=float(sc["norm_value"]), ref=float(n_vf)),
                tol=TOL, n_ranks=world,
                ok=bool(e_vs < TOL and e_pg < TOL and worst_scalar < TOL and e_grad < 5e-5 and e_par < 5e-5))
