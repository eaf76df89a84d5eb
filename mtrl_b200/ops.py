"""Component operators of the update path as stand-alone calls on CUDA tensors (C entries `mtrl_adam_polyak_step`,
`mtrl_sac_losses_fwd_bwd`; `MTSAC.network_forward` is the third, `mtrl_mlp_forward`).  The fused update runs exactly these
kernels; here they can be driven -- and checked against the oracle -- one at a time."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_vp, _i, _f = C.c_void_p, C.c_int, C.c_float


class SacLossesArgsC(C.Structure):
    _fields_ = [("mode", _i), ("rows", _i), ("width", _i), ("num_critics", _i), ("global_batch", _i), ("clip_q", _i), ("gamma", _f),
                ("H_target", _vp * 4), ("H_online", _vp * 4), ("w_target", _vp * 4), ("b_target", _vp * 4), ("w_online", _vp * 4),
                ("b_online", _vp * 4), ("tile_task", _vp), ("row_valid", _vp), ("rewards", _vp), ("dones", _vp), ("logp_next", _vp),
                ("logp", _vp), ("alpha", _vp), ("task_weights", _vp), ("dq", _vp), ("acc", _vp)]


L._EXTRA_DECLS.update({
    "mtrl_adam_polyak_step": ([_vp, _vp, _vp, _vp, _vp, C.c_longlong, _vp, _f, _f, _f, _f, _f, _f, _vp, _vp],),
    "mtrl_sac_losses_fwd_bwd": ([C.POINTER(SacLossesArgsC), _vp],),
})


def adam_polyak_step(params: torch.Tensor, grads: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: torch.Tensor, *,
                     lr: float = 3e-4, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-5, max_grad_norm: float | None = None,
                     target: torch.Tensor | None = None, tau: float = 0.005) -> dict:
    """In place: optax.chain(clip_by_global_norm, adam) + apply_updates (mtrl/config/optim.py:26-43), then
    optax.incremental_update into `target` if given (mtsac.py:607-613).  All tensors flat fp32 CUDA of the same length
    (a multiple of 4); `step` an int32 CUDA scalar (the Adam count, incremented).  Returns device scalars
    {"grad_norm" (pre-clip), "params_norm"}."""
    n = params.numel()
    for t in (params, grads, m, v) + ((target,) if target is not None else ()):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n):
            raise ValueError("adam_polyak_step: flat contiguous fp32 CUDA tensors of one length")
    if not (step.is_cuda and step.dtype == torch.int32 and step.numel() == 1):
        raise ValueError("adam_polyak_step: `step` must be an int32 CUDA scalar")
    scratch = torch.zeros(5, dtype=torch.float64, device=params.device)
    with torch.cuda.device(params.device):
        L.check(L.lib().mtrl_adam_polyak_step(_vp(params.data_ptr()), _vp(grads.data_ptr()), _vp(m.data_ptr()), _vp(v.data_ptr()),
                                              _vp(target.data_ptr() if target is not None else None), n, _vp(step.data_ptr()), lr, b1,
                                              b2, eps, -1.0 if max_grad_norm is None else float(max_grad_norm), tau,
                                              _vp(scratch.data_ptr()), _vp(L.current_stream_ptr())))
    return {"grad_norm": scratch[0].sqrt(), "params_norm": scratch[1].sqrt()}


def sac_losses(mode: str, *, H_online, w_online, b_online, tile_task, row_valid, alpha, task_weights, global_batch: int,
               H_target=None, w_target=None, b_target=None, rewards=None, dones=None, logp_next=None, logp=None, gamma: float = 0.99,
               clip_q: bool = False) -> dict:
    """The fused loss pass on packed rows (128-row tiles of one task each).  mode="critic": returns {"dq" (E, rows),
    "loss_sum" = sum w (Q - y)^2 over members and rows, "q_sum"}; mode="actor": {"dq", "loss_sum" = sum w (alpha logp - min Q)}.
    H_*: lists of (rows, W) tensors per member; w_*: (T, W, 1); b_*: (T, 1)."""
    E = len(H_online)
    rows, W = H_online[0].shape
    a = SacLossesArgsC(mode=0 if mode == "critic" else 1, rows=rows, width=W, num_critics=E, global_batch=global_batch,
                       clip_q=int(clip_q), gamma=gamma)
    keep = []

    def put(field, tensors):
        for e, t in enumerate(tensors or []):
            t = t.contiguous()
            keep.append(t)
            getattr(a, field)[e] = t.data_ptr()
    put("H_online", H_online); put("w_online", w_online); put("b_online", b_online)
    put("H_target", H_target); put("w_target", w_target); put("b_target", b_target)
    for name, t in (("tile_task", tile_task), ("row_valid", row_valid), ("rewards", rewards), ("dones", dones), ("logp_next", logp_next),
                    ("logp", logp), ("alpha", alpha), ("task_weights", task_weights)):
        if t is not None:
            t = t.contiguous()
            keep.append(t)
            setattr(a, name, t.data_ptr())
    dev = H_online[0].device
    dq = torch.zeros(E, rows, dtype=torch.float32, device=dev)
    acc = torch.zeros(4, dtype=torch.float64, device=dev)
    a.dq, a.acc = dq.data_ptr(), acc.data_ptr()
    with torch.cuda.device(dev):
        L.check(L.lib().mtrl_sac_losses_fwd_bwd(C.byref(a), _vp(L.current_stream_ptr())))
    if mode == "critic":
        return {"dq": dq, "loss_sum": acc[0], "q_sum": acc[1]}
    return {"dq": dq, "loss_sum": acc[2]}
