"""Integrated Gradients for the fusion model: the reference's explainability entry points, engine-native.

Mirrors ``compute_attributions`` / ``aggregate_importances`` of the reference (train2.py:776-866; the served copy is
back-end/app/libs/inference.py:268-330), which wrap ``captum.attr.IntegratedGradients`` (captum >= 0.6,
back-end/requirements.txt:16) around ``lambda v, a, mask: model(v, a, mask=mask)[1]``.  Captum's algorithm, as that call
uses it (method "gausslegendre", multiply_by_inputs True, internal_batch_size None):

    x_k   = base + alpha_k (x - base),  alpha_k = (1 + t_k) / 2,  (t_k, w_k) = Gauss-Legendre nodes / weights on [-1, 1]
    g_k   = d logits[:, target] / d x_k        for all k in ONE forward/backward over the n_steps * B expanded batch
    attr  = (x - base) * sum_k (w_k / 2) g_k

Here the expansion and the weighted reduction are one CUDA kernel each (csrc/attribution.cu) and the model evaluation is
the engine's eval-mode forward + a backward that skips every weight-gradient GEMM (``input_grads_only``); the model's
``param.grad`` tensors are not touched.  No Captum, no autograd graph.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._lib import BF16, F32, MmerError
from .engine import Engine, _compute_dtype

__all__ = ["compute_attributions", "aggregate_importances", "gauss_legendre_schedule"]

_MAX_EXPANDED = 32768   # samples per engine pass when the caller sets no internal_batch_size


def gauss_legendre_schedule(n_steps: int) -> Tuple[np.ndarray, np.ndarray]:
    """(alphas, step sizes) of Captum's default "gausslegendre" approximation (captum/attr/_utils/approximation_methods.py
    ``gauss_legendre_builders``): nodes and weights of the n-point rule mapped from [-1, 1] to [0, 1]."""
    if n_steps < 1:
        raise ValueError("n_steps must be positive")
    t, w = np.polynomial.legendre.leggauss(n_steps)
    return 0.5 * (1.0 + t), 0.5 * w


def _dt(t: torch.dtype) -> int:
    if t == torch.float32:
        return F32
    if t == torch.bfloat16:
        return BF16
    raise MmerError(f"unsupported dtype {t}")


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _expand(x: torch.Tensor, base: Optional[torch.Tensor], alphas: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    n = x.numel()
    out = torch.empty((alphas.numel() * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=out_dtype)
    _lib.check(_lib.load().mmer_ig_expand(x.data_ptr(), base.data_ptr() if base is not None else None,
                                          alphas.data_ptr(), out.data_ptr(), n, alphas.numel(),
                                          _dt(x.dtype), _dt(out_dtype), _stream()), "mmer_ig_expand")
    return out


def _reduce(grads: torch.Tensor, x: torch.Tensor, base: Optional[torch.Tensor], weights: torch.Tensor) -> torch.Tensor:
    attr = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().mmer_ig_reduce(grads.data_ptr(), x.data_ptr(), base.data_ptr() if base is not None else None,
                                          weights.data_ptr(), attr.data_ptr(),
                                          x.numel(), weights.numel(), _dt(x.dtype), _dt(grads.dtype), _stream()),
               "mmer_ig_reduce")
    return attr


def _run(model, v: torch.Tensor, a: torch.Tensor, mask: Optional[torch.Tensor], cdt: torch.dtype,
         onehot_of: Optional[torch.Tensor], want_grads: bool):
    """One eval-mode pass of the engine over (v, a); with ``onehot_of`` also the backward for d logits[:, target]."""
    eng: Engine = model._engine
    dev = v.device
    B, T = v.shape[0], v.shape[1]
    m = eng.make(B, T, cdt, False, 0.0, 0.0, 0, 0)
    eng.attach_shadow(m)
    ws = torch.empty(eng.workspace_bytes(m), device=dev, dtype=torch.uint8)
    m.workspace, m.workspace_bytes = ws.data_ptr(), ws.numel()
    m.video, m.audio = v.data_ptr(), a.data_ptr()
    if mask is not None:
        m.mask, m.has_mask = mask.data_ptr(), 1
    logits = torch.empty((B, eng.cfg["classes"]), device=dev, dtype=torch.float32)
    probs = torch.empty_like(logits)
    m.logits, m.probs = logits.data_ptr(), probs.data_ptr()
    Engine.forward(m)
    if not want_grads:
        return logits, None, None
    dl = torch.zeros_like(logits)
    dl.scatter_(1, onehot_of.view(-1, 1), 1.0)
    scratch = torch.zeros(eng.ctx.flat.numel(), device=dev, dtype=torch.float32)   # small reductions land here
    dv = torch.empty(v.shape, device=dev, dtype=cdt)
    da = torch.empty(a.shape, device=dev, dtype=cdt)
    m.grads, m.dlogits, m.dvideo, m.daudio = scratch.data_ptr(), dl.data_ptr(), dv.data_ptr(), da.data_ptr()
    m.input_grads_only = 1
    Engine.backward(m)
    return logits, dv, da


def compute_attributions(model, video_feats: torch.Tensor, audio_feats: torch.Tensor,
                         mask: Optional[torch.Tensor] = None, target: Union[None, int, torch.Tensor] = None,
                         n_steps: int = 50, baseline="zeros", device: str = "cuda",
                         internal_batch_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same contract as the reference's ``compute_attributions`` (train2.py:776-838): returns
    ``(attr_video [B, T, Dv], attr_audio [B, Da])`` (fp32, on ``device``), the model left in eval mode.

    ``target`` None = the predicted class per sample (argmax of the logits, train2.py:819-823); an int is broadcast,
    a tensor is taken per sample.  ``baseline`` is "zeros" -- the only value that works in the reference: its "mean"
    branch reads undefined globals (train2.py:815-819) and the served copy raises ValueError (inference.py:306-310) --
    or, beyond the reference, a ``(video_baseline, audio_baseline)`` pair of tensors shaped like the inputs.
    ``internal_batch_size`` (not in the reference's signature; Captum's name) bounds how many
    of the n_steps * B expanded samples go through the model at once; None = all like the reference, in passes of at
    most 32768 samples.
    """
    if isinstance(baseline, str) and baseline != "zeros":
        raise ValueError("Invalid baseline: only 'zeros' (or a pair of baseline tensors) is implemented")
    if not hasattr(model, "_engine"):
        raise MmerError("compute_attributions needs a mmer_b200 MultimodalEmotionModel")
    model.eval()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise MmerError("compute_attributions runs on a CUDA device only (there is no CPU path)")
    v = video_feats.detach().to(dev)
    a = audio_feats.detach().to(dev)
    if v.dim() != 3 or a.dim() != 2 or v.shape[0] != a.shape[0]:
        raise ValueError("expected video_feats [B, T, Dv] and audio_feats [B, Da]")
    cdt = _compute_dtype(v, getattr(model, "compute_dtype", None))
    if v.dtype not in (torch.float32, torch.bfloat16) or (v.dtype == torch.bfloat16 and cdt == torch.float32):
        v = v.float()
    v = v.contiguous()
    a = a.to(v.dtype).contiguous()
    if v[0].numel() % 8 or a[0].numel() % 8:
        raise MmerError("feature sizes must be multiples of 8")
    B = v.shape[0]
    bv = ba = None
    if not isinstance(baseline, str):
        bv, ba = (t.detach().to(device=dev, dtype=v.dtype).contiguous() for t in baseline)
        if bv.shape != v.shape or ba.shape != a.shape:
            raise ValueError("baseline tensors must have the shapes of the inputs")
    mk = None
    if mask is not None:
        mk = mask.to(device=dev, dtype=torch.bool).contiguous().view(torch.uint8)

    with torch.cuda.device(dev), torch.no_grad():
        vc, ac = (v, a) if v.dtype == cdt else (v.to(cdt), a.to(cdt))
        if target is None:
            logits, _, _ = _run(model, vc, ac, mk, cdt, None, False)
            tgt = logits.argmax(dim=1)
        elif isinstance(target, int):
            tgt = torch.full((B,), target, device=dev, dtype=torch.long)
        else:
            tgt = torch.as_tensor(target, device=dev, dtype=torch.long).view(-1)
            if tgt.numel() != B:
                raise ValueError("target must have one entry per sample")

        alphas_np, steps_np = gauss_legendre_schedule(n_steps)
        # None = everything in one pass like the reference, up to _MAX_EXPANDED samples per pass (the workspace of the
        # engine grows with the expanded batch); the sums over the chunks are the same attributions
        limit = _MAX_EXPANDED if internal_batch_size is None else int(internal_batch_size)
        chunk = min(n_steps, max(1, limit // B))
        v_attr = a_attr = None
        for k0 in range(0, n_steps, chunk):
            k1 = min(n_steps, k0 + chunk)
            n = k1 - k0
            al = torch.tensor(alphas_np[k0:k1], device=dev, dtype=torch.float32)
            wt = torch.tensor(steps_np[k0:k1], device=dev, dtype=torch.float32)
            vs, as_ = _expand(v, bv, al, cdt), _expand(a, ba, al, cdt)   # step-major, like Captum's cat over alphas
            mks = mk.repeat(n, 1) if mk is not None else None
            _, dv, da = _run(model, vs, as_, mks, cdt, tgt.repeat(n), True)
            pv, pa = _reduce(dv, v, bv, wt), _reduce(da, a, ba, wt)
            v_attr = pv if v_attr is None else v_attr + pv
            a_attr = pa if a_attr is None else a_attr + pa
    return v_attr, a_attr


def aggregate_importances(attr_video: torch.Tensor, attr_audio: torch.Tensor, abs_sum: bool = True):
    """train2.py:841-866: per-feature importance -- video summed over time -> [B, Dv]; audio as is -> [B, Da];
    magnitudes when ``abs_sum``."""
    if abs_sum:
        attr_video, attr_audio = attr_video.abs(), attr_audio.abs()
    return attr_video.sum(dim=1), attr_audio
