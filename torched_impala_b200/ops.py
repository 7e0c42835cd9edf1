"""Thin per-kernel wrappers over the C ABI for torch CUDA tensors.

Each function allocates its outputs with torch (device-memory plumbing), launches one
C-ABI call on the current torch stream and returns the outputs.  They exist for the
parity tests and for users who want a single piece of the update (e.g. V-trace only);
the learner itself goes through `engine.LearnerEngine`, which pre-allocates everything.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi

PKEYS = ("model.0.weight", "model.0.bias", "model.3.weight", "model.3.bias")


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr())


def _need_cuda(*ts):
    for t in ts:
        if not (torch.is_tensor(t) and t.is_cuda and t.is_contiguous()):
            raise _cabi.ImpalaCudaError("expected contiguous CUDA tensors (there is no CPU fallback)")


def pack_params(state_dict: dict, device="cuda") -> torch.Tensor:
    """state_dict of one reference MLP (models.py:13-18) -> flat float32 parameter block."""
    w1 = np.asarray(state_dict[PKEYS[0]], dtype=np.float32)
    w2 = np.asarray(state_dict[PKEYS[2]], dtype=np.float32)
    H, O = w1.shape
    N2 = w2.shape[0]
    offs, total = _cabi.param_layout(O, H, N2)
    flat = np.zeros(total, np.float32)
    for k, off in zip(PKEYS, offs):
        a = np.asarray(state_dict[k], dtype=np.float32).reshape(-1)
        flat[off:off + a.size] = a
    return torch.from_numpy(flat).to(device)


def unpack_grad(flat, O: int, H: int, N2: int) -> dict:
    offs, _ = _cabi.param_layout(O, H, N2)
    shapes = ((H, O), (H,), (N2, H), (N2,))
    a = flat.detach().cpu().numpy()
    return {k: a[off:off + int(np.prod(s))].reshape(s).copy() for k, off, s in zip(PKEYS, offs, shapes)}


def mlp_forward(x, params, O: int, H: int, N2: int):
    _need_cuda(x, params)
    M = x.numel() // O
    out = torch.empty(M, N2, dtype=torch.float32, device=x.device)
    _cabi.check(_cabi.lib().impala_mlp_forward(_p(x), _p(params), _p(out), M, O, H, N2, _st()),
                "impala_mlp_forward")
    return out


def mlp_backward(x, params, dout, O: int, H: int, N2: int):
    _need_cuda(x, params, dout)
    lib = _cabi.lib()
    M = x.numel() // O
    _, total = _cabi.param_layout(O, H, N2)
    nbytes = lib.impala_mlp_backward_workspace(M, O, H, N2)
    if nbytes < 0:
        _cabi.check(int(nbytes), "impala_mlp_backward_workspace")
    ws = torch.zeros(int(nbytes), dtype=torch.uint8, device=x.device)  # control words start at 0
    grad = torch.empty(total, dtype=torch.float64, device=x.device)
    _cabi.check(lib.impala_mlp_backward(_p(x), _p(params), _p(dout), _p(grad), _p(ws), int(nbytes),
                                        M, O, H, N2, _st()), "impala_mlp_backward")
    return grad


def mlp_forward_pair(x, params_pi, params_vf, M_pi: int, M_vf: int, O: int, H_pi: int, H_vf: int, A: int):
    """Policy logits on the first M_pi rows of x and values on the first M_vf rows, one call."""
    _need_cuda(x, params_pi, params_vf)
    logits = torch.empty(M_pi, A, dtype=torch.float32, device=x.device)
    values = torch.empty(M_vf, dtype=torch.float32, device=x.device)
    _cabi.check(_cabi.lib().impala_mlp_forward_pair(_p(x), _p(params_pi), _p(params_vf), _p(logits), _p(values),
                                                    M_pi, M_vf, O, H_pi, H_vf, A, _st()),
                "impala_mlp_forward_pair")
    return logits, values


def mlp_backward_pair(x, params_pi, params_vf, dlogits, dv, O: int, H_pi: int, H_vf: int, A: int):
    _need_cuda(x, params_pi, params_vf, dlogits, dv)
    lib = _cabi.lib()
    M_pi, M_vf = dlogits.numel() // A, dv.numel()
    out, wss = [], []
    for M, H, N2 in ((M_pi, H_pi, A), (M_vf, H_vf, 1)):
        nbytes = lib.impala_mlp_backward_workspace(M, O, H, N2)
        if nbytes < 0:
            _cabi.check(int(nbytes), "impala_mlp_backward_workspace")
        wss.append(torch.zeros(int(nbytes), dtype=torch.uint8, device=x.device))
        out.append(torch.empty(_cabi.param_layout(O, H, N2)[1], dtype=torch.float64, device=x.device))
    _cabi.check(lib.impala_mlp_backward_pair(_p(x), _p(params_pi), _p(params_vf), _p(dlogits), _p(dv),
                                             _p(out[0]), _p(out[1]), _p(wss[0]), wss[0].numel(), _p(wss[1]),
                                             wss[1].numel(), M_pi, M_vf, O, H_pi, H_vf, A, _st()),
                "impala_mlp_backward_pair")
    return out[0], out[1]


def vtrace(cur_logits, beh_logits, actions, rewards, done, lens, v, gamma, rho_bar, c_bar,
           mode="reference"):
    _need_cuda(cur_logits, beh_logits, actions, rewards, done, lens, v)
    T, B, A = cur_logits.shape
    vs = torch.empty(T + 1, B, dtype=torch.float32, device=v.device)
    pg = torch.empty(T, B, dtype=torch.float32, device=v.device)
    _cabi.check(_cabi.lib().impala_vtrace(_p(cur_logits), _p(beh_logits), _p(actions), _p(rewards),
                                          _p(done), _p(lens), _p(v), _p(vs), _p(pg), T, B, A,
                                          float(gamma), float(rho_bar), float(c_bar),
                                          _cabi.MODES[mode], _st()), "impala_vtrace")
    return vs, pg


def vtrace_loss(cur_logits, beh_logits, actions, rewards, done, lens, v, hp, inv_batch,
                mode="reference"):
    _need_cuda(cur_logits, beh_logits, actions, rewards, done, lens, v)
    T, B, A = cur_logits.shape
    dev = v.device
    vs = torch.empty(T + 1, B, dtype=torch.float32, device=dev)
    pg = torch.empty(T, B, dtype=torch.float32, device=dev)
    dlogits = torch.empty(T, B, A, dtype=torch.float32, device=dev)
    dv = torch.empty(T + 1, B, dtype=torch.float32, device=dev)
    scalars = torch.empty(4, dtype=torch.float64, device=dev)
    lib = _cabi.lib()
    ws_bytes = int(lib.impala_vtrace_loss_workspace(T, B, A))
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
    _cabi.check(lib.impala_vtrace_loss(
        _p(cur_logits), _p(beh_logits), _p(actions), _p(rewards), _p(done), _p(lens), _p(v), _p(vs),
        _p(pg), _p(dlogits), _p(dv), _p(scalars), _p(ws), ws_bytes, T, B, A, float(hp.gamma),
        float(hp.rho_bar),
        float(hp.c_bar), float(hp.v_loss_c), float(hp.policy_loss_c), float(hp.entropy_c),
        float(inv_batch), _cabi.MODES[mode], _st()), "impala_vtrace_loss")
    return dict(vs=vs, pg_adv=pg, dlogits=dlogits, dv=dv, scalars=scalars)


def clip_adam(params, grad, m, v, step, n_policy, max_norm, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    _need_cuda(params, grad, m, v, step)
    norms = torch.empty(2, dtype=torch.float64, device=params.device)
    _cabi.check(_cabi.lib().impala_clip_adam(_p(params), _p(grad), _p(m), _p(v), _p(step),
                                             int(n_policy), params.numel(), float(max_norm),
                                             float(lr), float(beta1), float(beta2), float(eps),
                                             _p(norms), _st()), "impala_clip_adam")
    return norms
