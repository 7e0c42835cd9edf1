"""Drop-in `Learner` for the reference's actor/learner split, running the update on a B200.

Same constructor, lifecycle methods, events, checkpoint format and module-level loss
helpers as `/root/reference/learner.py` (constructor `:18-28`, `start/terminate/join`
`:55-65`, `_learn` `:67-275`, `save/load/policy_weights` `:277-295`, helpers `:298-321`),
so `train.py:69` and the unmodified `actor.py` (`:68,70,118,121`) work against it:

  * trajectories arrive as pickled `utils.Trajectory` objects through the same `mp.Queue`;
    `queue.Empty` after `timeout` seconds sets `completion` and re-raises (`:91-100`);
  * after every update the new policy weights are copied IN PLACE into the (shared-memory,
    float64) `policy` module the actors read through `learner.policy_weights`;
  * `update_counter.increment()` once per update; TensorBoard scalars under the same tags;
    checkpoints with keys `policy_state_dict` / `value_fn_state_dict`.

What changes is where the arithmetic happens: `_learn` packs each trajectory into a pinned,
zero-padded time-major host slab (replacing the torch.stack calls at `:104-109,117`) and
hands the batch to `engine.LearnerEngine`, i.e. to the sm_100a kernels behind
include/impala_b200.h.  CUDA is initialised inside the learner process only (`train.py:42`
forces the fork start method, so the parent must never touch the device).

Dropout: the reference MLPs carry `Dropout(p=0.8)` (models.py:15,44).  This learner
implements the deterministic (`.eval()`) forward, the setting every parity number is
quoted in.
"""
from __future__ import annotations

import queue
from pathlib import Path

import numpy as np
import torch
import torch.multiprocessing as mp

PKEYS = ("model.0.weight", "model.0.bias", "model.3.weight", "model.3.bias")


def pack_trajectory(views: dict, b: int, traj, T: int) -> float:
    """Write one reference-format trajectory into column `b` of a host batch slab.

    Replaces learner.py:104-109,117 (five torch.stack calls + `disc`): float64 -> float32,
    int64 -> int32, bool -> u8, zero padding past the trajectory's length.  Returns the
    trajectory's reward sum (learner.py:108)."""
    L = len(traj.r)
    if L < 1 or L > T:
        raise ValueError(f"trajectory {getattr(traj, 'id', '?')} has {L} steps; the learner was "
                         f"built for 1..{T} (hp.max_timesteps)")
    if len(traj.obs) != L + 1 or len(traj.a) != L or len(traj.d) != L or len(traj.logits) != L:
        raise ValueError("malformed trajectory: obs must have one more entry than a/r/d/logits")
    obs = torch.stack(traj.obs).to(torch.float32).numpy()
    views["obs"][:L + 1, b] = obs
    views["obs"][L + 1:, b] = 0
    views["beh_logits"][:L, b] = torch.stack(traj.logits).to(torch.float32).numpy()
    views["beh_logits"][L:, b] = 0
    views["actions"][:L, b] = torch.stack(traj.a).reshape(L).to(torch.int32).numpy()
    views["actions"][L:, b] = 0
    r = torch.stack(traj.r)
    views["rewards"][:L, b] = r.to(torch.float32).numpy()
    views["rewards"][L:, b] = 0
    views["done"][:L, b] = torch.stack(traj.d).to(torch.uint8).numpy()
    views["done"][L:, b] = 0
    views["lens"][b] = L
    return float(r.sum(dtype=torch.float64))  # summed as received (float64 from actor.py), like learner.py:108


def _dims(policy, value_fn):
    sd_p, sd_v = policy.state_dict(), value_fn.state_dict()
    H_pi, O = sd_p[PKEYS[0]].shape
    A = sd_p[PKEYS[2]].shape[0]
    H_v = sd_v[PKEYS[0]].shape[0]
    return int(O), int(A), int(H_pi), int(H_v)


class Learner:
    def __init__(self, id, hparams, policy, value_fn, q, update_counter, log_path=None,
                 timeout=200, device="cuda:0", mode="reference"):
        self.id = id
        self.hp = hparams
        self.policy = policy
        self.value_fn = value_fn
        self.timeout = timeout
        self.q = q
        self.update_counter = update_counter
        self.device = device
        self.mode = mode
        self.log_path = log_path
        if self.log_path is not None:
            self.log_path = Path(log_path) / Path(f"l{self.id}")
            self.log_path.mkdir(parents=True, exist_ok=False)
        self.evaluator = None  # optional callable(policy) -> (mean_reward, std); see _evaluate
        self.completion = mp.Event()
        self.p = mp.Process(target=self._learn, name=f"learner_{self.id}")
        print(f"[main] learner_{self.id} Initialized")

    # -------------------------------------------------------------- lifecycle (learner.py:55-65)
    def start(self):
        self.completion.clear()
        self.p.start()
        print(f"[main] Started learner_{self.id} with pid {self.p.pid}")

    def terminate(self):
        self.p.terminate()
        print(f"[main] Terminated learner_{self.id}")

    def join(self):
        self.p.join()

    # ------------------------------------------------------------------ helpers
    def _make_engine(self):
        from .engine import LearnerEngine

        O, A, H_pi, H_v = _dims(self.policy, self.value_fn)
        eng = LearnerEngine(self.hp.max_timesteps, self.hp.batch_size, O, A, H_pi, H_v, self.hp,
                            device=self.device, mode=self.mode)
        eng.load_state({"policy": self.policy.state_dict(), "value_fn": self.value_fn.state_dict()})
        return eng

    def _publish(self, eng):
        """New weights -> the float64 modules, in place (actors read them lock-free, actor.py:70)."""
        st = eng.state()
        with torch.no_grad():
            for mod, grp in ((self.policy, "policy"), (self.value_fn, "value_fn")):
                for k, t in mod.state_dict().items():
                    t.copy_(st[grp][k].to(t.dtype))

    def _evaluate(self):
        """learner.py:195-214 runs utils.test_policy (a gym rollout) inside the learner.  The
        environment side is outside this package: use `self.evaluator` if set, else the
        reference's own `utils.test_policy` when this class is deployed inside that repo."""
        if self.evaluator is not None:
            return self.evaluator(self.policy)
        try:
            import utils as ref_utils  # the reference's utils.py, if on sys.path

            return ref_utils.test_policy(self.policy, self.hp.env_name, self.hp.eval_eps, True,
                                         self.hp.max_timesteps)
        except Exception as e:  # no gym / not inside the reference tree
            print(f"[learner_{self.id}] evaluation skipped: {e!r}")
            return None

    # ---------------------------------------------------------------- the update loop
    def _collect(self, views, writer):
        """Pull hp.batch_size trajectories off the queue into one host slab (learner.py:89-109)."""
        hp = self.hp
        reward = 0.0
        for b in range(hp.batch_size):
            try:
                traj = self.q.get(timeout=self.timeout)
            except queue.Empty:
                print(f"[learner_{self.id}] queue empty for {self.timeout} s - giving up")
                if writer is not None:
                    writer.close()
                self.completion.set()  # lets the actors leave their put() retry loop (actor.py:121)
                raise
            if hp.verbose >= 2:
                print(f"[learner_{self.id}] packing traj_{traj.id} into column {b}")
            reward += pack_trajectory(views, b, traj, hp.max_timesteps) / hp.batch_size
            del traj  # drop the shared-memory handles of its ~5T tensors right away
        return reward

    def _report(self, writer, n, reward, sc):
        """Console line + the five TensorBoard scalars of learner.py:188-192,217-240."""
        if self.hp.verbose >= 1:
            print(f"[learner_{self.id}] update {n}: batch mean reward {reward:.2f}, "
                  f"loss {sc['total_loss']:.2f}")
        if writer is None:
            return
        tag = f"learner_{self.id}"
        for name, val in (("rewards/batch_mean_reward", reward), ("loss/policy_loss", sc["policy_loss"]),
                          ("loss/value_fn_loss", sc["value_fn_loss"]),
                          ("loss/policy_entropy", sc["policy_entropy"]),
                          ("loss/total_loss", sc["total_loss"])):
            writer.add_scalar(f"{tag}/{name}", val, n)

    def _periodic(self, writer, n):
        hp = self.hp
        if hp.eval_every is not None and n % hp.eval_every == 0:          # learner.py:195-214
            res = self._evaluate()
            if res is not None:
                if hp.verbose >= 1:
                    print(f"[learner_{self.id}] update {n}: evaluation reward {res[0]:.2f} +- {res[1]:.2f}")
                if writer is not None:
                    writer.add_scalar(f"learner_{self.id}/rewards/evaluation_reward", res[0], n)
        if self.log_path is not None and n % hp.save_every == 0:          # learner.py:243-251
            path = self.log_path / f"IMPALA_{hp.env_name}_l{self.id}_{n}.pt"
            self.save(path)
            print(f"[learner_{self.id}] checkpoint -> {path}")

    def _learn(self):
        """Process target (learner.py:67): loop until the shared counter reaches max_updates."""
        writer = None
        try:
            eng = self._make_engine()  # first CUDA call of this process (post-fork)
            if self.log_path is not None:
                from torch.utils.tensorboard import SummaryWriter

                writer = SummaryWriter(self.log_path)
                writer.add_text("hyperparameters", f"{self.hp}")
            done, slot = 0, 0
            ring = self.q if hasattr(self.q, "collect_batch") else None  # ring.RingQueue
            if ring is not None:
                if ring.slab_bytes != eng.slab_bytes:
                    raise ValueError("RingQueue and learner disagree on (T, B, obs, actions)")
                eng.register_host(ring.slab_address(0), ring.slab_bytes * ring.K)
            while done < self.hp.max_updates:
                if ring is None:
                    reward = self._collect(eng.host_batch(slot), writer)
                    eng.ingest(slot)
                else:
                    try:
                        k, reward = ring.collect_batch(self.timeout)
                    except queue.Empty:
                        print(f"[learner_{self.id}] no trajectory for {self.timeout} s - giving up")
                        self.completion.set()
                        raise
                    eng.ingest_from(ring.slab_address(k), slot)  # DMA straight out of shared memory
                    eng.slab_ready[slot].synchronize()
                    ring.release(k)
                eng.step(slot)
                sc = eng.read_scalars()
                self._publish(eng)
                slot ^= 1
                self._report(writer, done + 1, reward, sc)
                self._periodic(writer, done + 1)
                self.update_counter.increment()                            # learner.py:254-255
                done = self.update_counter.value
            print(f"[learner_{self.id}] done after {done} updates")
            self.completion.set()
        except KeyboardInterrupt:
            print(f"[learner_{self.id}] interrupted")
            self.completion.set()
        except Exception:
            # The reference re-raises without setting `completion` (learner.py:271-275), which
            # leaves train.py:84 waiting forever; here the actors are released as well.
            print(f"[learner_{self.id}] failed")
            self.completion.set()
            raise
        finally:
            if writer is not None:
                writer.close()

    # ------------------------------------------------------ checkpoints (learner.py:277-295)
    def save(self, path):
        torch.save({"policy_state_dict": self.policy.state_dict(),
                    "value_fn_state_dict": self.value_fn.state_dict()}, path)

    def load(self, path):
        checkpoint = torch.load(path)
        self.policy.load_state_dict(checkpoint["policy_state_dict"])
        self.value_fn.load_state_dict(checkpoint["value_fn_state_dict"])

    @property
    def policy_weights(self):
        return self.policy.state_dict()


# ----------------------------------------------------------------------------------------------
# Module-level loss helpers with the reference's names, signatures and sign conventions
# (learner.py:298-321), backed by the C ABI.  Inputs are CUDA tensors (any float dtype; the
# kernels compute in float32 and reduce in float64); results come back in the input dtype and
# are differentiable with respect to the logits / advantages exactly where the reference's are.
# ----------------------------------------------------------------------------------------------
def _lib_and_stream():
    import ctypes as C

    from . import _cabi

    return _cabi, _cabi.lib(), C.c_void_p(torch.cuda.current_stream().cuda_stream), C


def _cuda_f32(t):
    from . import _cabi

    if not (torch.is_tensor(t) and t.is_cuda):
        raise _cabi.ImpalaCudaError("loss helpers need CUDA tensors (there is no CPU fallback)")
    return t.detach().to(torch.float32).contiguous()


class _PolicyTerms(torch.autograd.Function):
    """(logits, actions) -> (log pi(a), sum_k p_k log p_k) per row."""

    @staticmethod
    def forward(ctx, logits, actions):
        _cabi, lib, st, C = _lib_and_stream()
        z = _cuda_f32(logits).reshape(-1, logits.shape[-1])
        a = actions.detach().reshape(-1).to(torch.int32).contiguous()
        M, A = z.shape
        lp = torch.empty(M, dtype=torch.float32, device=z.device)
        ne = torch.empty(M, dtype=torch.float32, device=z.device)
        _cabi.check(lib.impala_policy_terms(C.c_void_p(z.data_ptr()), C.c_void_p(a.data_ptr()),
                                            C.c_void_p(lp.data_ptr()), C.c_void_p(ne.data_ptr()), M, A, st),
                    "impala_policy_terms")
        ctx.save_for_backward(z, a)
        ctx.in_shape, ctx.in_dtype = logits.shape, logits.dtype
        return lp.to(logits.dtype), ne.to(logits.dtype)

    @staticmethod
    def backward(ctx, g_lp, g_ne):
        _cabi, lib, st, C = _lib_and_stream()
        z, a = ctx.saved_tensors
        M, A = z.shape
        gl = None if g_lp is None else g_lp.to(torch.float32).contiguous()
        gn = None if g_ne is None else g_ne.to(torch.float32).contiguous()
        dz = torch.empty_like(z)
        _cabi.check(lib.impala_policy_terms_backward(
            C.c_void_p(z.data_ptr()), C.c_void_p(a.data_ptr()),
            C.c_void_p(gl.data_ptr()) if gl is not None else None,
            C.c_void_p(gn.data_ptr()) if gn is not None else None,
            C.c_void_p(dz.data_ptr()), M, A, st), "impala_policy_terms_backward")
        return dz.reshape(ctx.in_shape).to(ctx.in_dtype), None


class _Reduce(torch.autograd.Function):
    """float64-accumulated scalar reductions: mode 0 sum(a), 1 0.5*sum(a^2), 2 sum(a*b) (b constant)."""

    @staticmethod
    def forward(ctx, a, b, mode):
        _cabi, lib, st, C = _lib_and_stream()
        af = _cuda_f32(a).reshape(-1)
        bf = _cuda_f32(b).reshape(-1) if b is not None else None
        out = torch.empty(1, dtype=torch.float64, device=af.device)
        _cabi.check(lib.impala_reduce(C.c_void_p(af.data_ptr()),
                                      C.c_void_p(bf.data_ptr()) if bf is not None else None,
                                      af.numel(), mode, C.c_void_p(out.data_ptr()), st), "impala_reduce")
        ctx.mode, ctx.shape, ctx.dtype = mode, a.shape, a.dtype
        ctx.save_for_backward(af, bf if bf is not None else af)
        return out[0].to(a.dtype)

    @staticmethod
    def backward(ctx, g):
        af, bf = ctx.saved_tensors
        if ctx.mode == 0:
            ga = g.to(torch.float32).expand(af.shape)
        elif ctx.mode == 1:
            ga = g.to(torch.float32) * af
        else:
            ga = g.to(torch.float32) * bf
        return ga.reshape(ctx.shape).to(ctx.dtype), None, None


def action_log_probs(policy_logits, actions):
    """log pi(a|x) of the taken actions, shaped like `actions` (learner.py:298-303)."""
    return _PolicyTerms.apply(policy_logits, actions)[0].view_as(actions)


def compute_baseline_loss(advantages):
    """0.5 * sum(advantages ** 2)  (learner.py:306-307)."""
    return _Reduce.apply(advantages, None, 1)


def compute_entropy_loss(logits):
    """The NEGATIVE entropy sum(p * log p), as in the reference (learner.py:310-314)."""
    return _Reduce.apply(_PolicyTerms.apply(logits, torch.zeros(logits.shape[:-1], dtype=torch.int32,
                                                                 device=logits.device))[1], None, 0)


def compute_policy_gradient_loss(logits, actions, advantages):
    """sum(-log pi(a|x) * advantages.detach())  (learner.py:317-321)."""
    lp = _PolicyTerms.apply(logits, actions)[0]
    return _Reduce.apply(lp, -advantages.detach().reshape(-1), 2)
