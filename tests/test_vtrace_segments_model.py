"""CPU model of the work split inside vtrace_lane_kernel (csrc/vtrace_loss.cu).

The kernel cuts the T-step backward recurrence  acc_t = fa_t + g_t * acc_(t+1)  of every trajectory
into chunks of S * NSEG steps (walked backwards, carry in a register) and each chunk into NSEG
segments of S steps owned by different warps: a thread scans its segment with carry 0 keeping
acc0_t and the running product P_t = g_t ... g_(segment end), publishes the segment's composed
map (acc0, P), and after one barrier composes the later segments' maps to get the accumulator that
enters its own segment, then fixes up acc_t = acc0_t + P_t * carry.  This test runs exactly that
schedule in numpy (float64, so only the algebra is under test) for many (T, S, NSEG) and ragged
lengths and checks it against the sequential oracle (reference learner.py:126-135)."""
import numpy as np
import pytest

from oracle import impala_oracle as orc
from torched_impala_b200 import synth


def segmented_vtrace(v, cur, beh, actions, rewards, done, lens, gamma, rho_bar, c_bar, S, NSEG, mode="reference"):
    T, B = rewards.shape
    v = np.asarray(v, np.float64)
    t_idx = np.arange(T)[:, None]
    valid = t_idx < lens[None, :]
    ratio = np.exp(orc.taken_log_probs(np.asarray(cur, np.float64), actions)
                   - orc.taken_log_probs(np.asarray(beh, np.float64), actions))
    rho = np.where(valid, np.minimum(ratio, rho_bar), 0.0)
    cc = np.where(valid, np.minimum(ratio, c_bar), 0.0)
    disc = np.where(valid & (done == 0), np.float64(np.float32(gamma)), 0.0)
    g = disc * cc
    r = rewards.astype(np.float64)
    if mode == "reference":
        fa = rho * (r + gamma * v[1:] - v[:1]) - g * v[1:]
    else:
        fa = rho * (r + disc * v[1:] - v[:-1])
    rows = S * NSEG
    nch = (T + rows - 1) // rows
    pad = nch * rows - T
    fa = np.concatenate([fa, np.zeros((pad, B))])   # steps past the unroll: identity-reset maps
    g = np.concatenate([g, np.zeros((pad, B))])
    acc = np.zeros((nch * rows + 1, B))
    chunk_carry = np.zeros(B)
    for c in range(nch - 1, -1, -1):
        acc0 = np.zeros((NSEG, S + 1, B))
        P = np.ones((NSEG, S + 1, B))
        for w in range(NSEG):                            # every (warp, lane) independently
            tb = c * rows + w * S
            for i in range(S - 1, -1, -1):
                acc0[w, i] = fa[tb + i] + g[tb + i] * acc0[w, i + 1]
                P[w, i] = g[tb + i] * P[w, i + 1]
        carry = chunk_carry.copy()                       # after the barrier: compose later segments
        mine = np.zeros((NSEG, B))
        for s in range(NSEG - 1, -1, -1):
            mine[s] = carry
            carry = acc0[s, 0] + P[s, 0] * carry
        chunk_carry = carry
        for w in range(NSEG):                            # fix-up
            tb = c * rows + w * S
            for i in range(S):
                acc[tb + i] = acc0[w, i] + P[w, i] * mine[w]
    acc = acc[: T + 1]
    acc[T] = 0.0
    vs = acc + v
    pg = rho * (r + disc * vs[1:] - v[:-1])
    vs = np.where(np.arange(T + 1)[:, None] <= lens[None, :], vs, 0.0)
    return vs, pg


@pytest.mark.parametrize("T,S,NSEG", [(20, 2, 10), (20, 1, 20), (100, 2, 8), (100, 5, 10), (7, 2, 4), (33, 5, 3),
                                      (129, 2, 16), (1, 1, 1), (16, 2, 8)])
@pytest.mark.parametrize("mode", ["reference", "paper"])
def test_segmented_schedule_equals_sequential_recurrence(T, S, NSEG, mode):
    B, A = 13, 4
    b = synth.make_batch(T * 31 + S, T, B, 3, A, ragged=True)
    rng = np.random.default_rng(T + NSEG)
    cur = rng.standard_normal((T, B, A))
    v = rng.standard_normal((T + 1, B))
    want_vs, want_pg, _ = orc.vtrace(v, cur, b["beh_logits"], b["actions"], b["rewards"], b["done"], b["lens"],
                                     0.97, 0.9, 0.8, mode)
    got_vs, got_pg = segmented_vtrace(v, cur, b["beh_logits"], b["actions"], b["rewards"], b["done"], b["lens"],
                                      0.97, 0.9, 0.8, S, NSEG, mode)
    np.testing.assert_allclose(got_vs, want_vs, rtol=0, atol=1e-10)
    np.testing.assert_allclose(got_pg, want_pg, rtol=0, atol=1e-10)
