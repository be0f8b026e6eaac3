"""CPU oracle for the audio-visual fusion step.  TEST INFRASTRUCTURE ONLY.

This file restates, with primitive tensor arithmetic (matmul / exp / sum, no
``nn.TransformerEncoder``, no ``F.cross_entropy``, no ``optim.Adam``), the
algorithm that the reference delegates to PyTorch modules.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; the product path (``mmer_b200``) never does.

Parity pinning: the reference has no tests or golden vectors of its own for this
path, so this oracle is pinned against outputs of the
reference classes themselves, generated in the build container by
``tests/golden/make_golden.py`` (which imports /root/reference/train.py and
train2.py) and committed under ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every function here against them.

Reference citations (relative to the reference repo root):
  FocalLoss                      train.py:20-37  == train2.py:40-70
  CrossModalFusion (v2, LN)      train2.py:77-193 == back-end/app/libs/model.py:7-76
  CrossModalFusion (v1, BN)      train.py:39-106
  EmotionClassifier v2 / v1      train2.py:196-238 / train.py:108-130
  MultimodalEmotionModel v2/v1   train2.py:241-292 / train.py:133-142
  training step v2 / v1          train2.py:570-579 / train.py:293-297
  encoder layer (post-norm)      torch nn.TransformerEncoderLayer as configured at
                                 train.py:54-57, train2.py:111-118
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor
LN_EPS = 1e-5
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------
# optional emulation of reduced-precision ACTIVATION STORAGE (test infrastructure)
# --------------------------------------------------------------------------
# The bf16 product path keeps every activation and activation-gradient tensor in bf16 between
# kernels while all arithmetic inside a kernel is fp32.  Inside ``storage_rounding(torch.bfloat16)``
# the oracle rounds values (forward) and gradients (backward) at those same tensor boundaries, so a
# bf16 run can be compared with an oracle that has the same quantisation points.  Outside the
# context manager ``_q`` is the identity and the oracle is the exact restatement pinned by the
# golden vectors.
_STORAGE = None


class _RoundAtBoundary(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dt):
        ctx.dt = dt
        return x.to(dt).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dt).to(g.dtype), None


def _q(x: Tensor) -> Tensor:
    return x if _STORAGE is None else _RoundAtBoundary.apply(x, _STORAGE)


class storage_rounding:
    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global _STORAGE
        self.prev, _STORAGE = _STORAGE, self.dtype
        return self

    def __exit__(self, *exc):
        global _STORAGE
        _STORAGE = self.prev
        return False


# --------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------
def log_softmax(x: Tensor) -> Tensor:
    m = x.max(dim=-1, keepdim=True).values
    z = x - m
    return z - torch.log(torch.exp(z).sum(dim=-1, keepdim=True))


def focal_loss(logits: Tensor, targets: Tensor, gamma: float = 2.0,
               alpha: Optional[Tensor] = None, reduction: str = "mean") -> Tensor:
    """train.py:27-37.  Plain mean over the batch even when alpha is given."""
    lsm = log_softmax(logits)
    ce = -lsm.gather(1, targets.view(-1, 1)).squeeze(1)
    pt = torch.exp(-ce)
    fl = (1 - pt) ** gamma * ce
    if alpha is not None:
        fl = alpha.to(fl.dtype)[targets] * fl
    if reduction == "mean":
        return fl.mean()
    if reduction == "sum":
        return fl.sum()
    return fl


def focal_loss_grad(logits: Tensor, targets: Tensor, gamma: float = 2.0,
                    alpha: Optional[Tensor] = None, reduction: str = "mean") -> Tensor:
    """Closed-form d loss / d logits."""
    lsm = log_softmax(logits)
    p = torch.exp(lsm)
    ce = -lsm.gather(1, targets.view(-1, 1)).squeeze(1)
    pt = torch.exp(-ce)
    one_m = 1 - pt
    # d fl / d ce = (1-pt)^g + g * pt * (1-pt)^(g-1) * ce
    dfl = one_m ** gamma + gamma * pt * one_m ** (gamma - 1) * ce
    if alpha is not None:
        dfl = dfl * alpha.to(dfl.dtype)[targets]
    if reduction == "mean":
        dfl = dfl / logits.shape[0]
    onehot = torch.zeros_like(p)
    onehot.scatter_(1, targets.view(-1, 1), 1.0)
    return dfl.unsqueeze(1) * (p - onehot)


def weighted_ce(logits: Tensor, targets: Tensor, weight: Optional[Tensor] = None) -> Tensor:
    """nn.CrossEntropyLoss(weight=w) as used at train2.py:523: sum(w_y*ce)/sum(w_y)."""
    lsm = log_softmax(logits)
    ce = -lsm.gather(1, targets.view(-1, 1)).squeeze(1)
    if weight is None:
        return ce.mean()
    w = weight.to(ce.dtype)[targets]
    return (w * ce).sum() / w.sum()


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    y = x @ w.t()
    return y if b is None else y + b


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = LN_EPS) -> Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)  # biased
    return (x - mu) / torch.sqrt(var + eps) * w + b


def batch_norm(x2d: Tensor, w: Tensor, b: Tensor, running_mean: Tensor, running_var: Tensor,
               training: bool, eps: float = BN_EPS, momentum: float = BN_MOMENTUM,
               update: Optional[Dict[str, Tensor]] = None, prefix: str = "") -> Tensor:
    """BatchNorm1d over rows of a (N, C) matrix.  train.py:51-52,66-74,116,125.

    In training mode uses biased batch variance for normalisation; when ``update`` is
    given, writes the new running stats (unbiased variance, momentum 0.1) into it.
    """
    if training:
        n = x2d.shape[0]
        mu = x2d.mean(dim=0)
        var = ((x2d - mu) ** 2).mean(dim=0)
        if update is not None:
            unb = var * (n / max(n - 1, 1))
            update[prefix + "running_mean"] = (1 - momentum) * running_mean + momentum * mu.detach()
            update[prefix + "running_var"] = (1 - momentum) * running_var + momentum * unb.detach()
    else:
        mu, var = running_mean, running_var
    return (x2d - mu) / torch.sqrt(var + eps) * w + b


def mha(x: Tensor, in_w: Tensor, in_b: Tensor, out_w: Tensor, out_b: Tensor,
        key_pad: Optional[Tensor], num_heads: int) -> Tuple[Tensor, Tensor]:
    """Self-attention over batch-major tokens x:(B,S,F); key_pad:(B,S) True = ignore key.

    Returns (output (B,S,F), probabilities (B,H,S,S)).  Equivalent to
    nn.MultiheadAttention(F, H) with src_key_padding_mask, dropout 0.
    """
    B, S, Fd = x.shape
    d = Fd // num_heads
    qkv = _q(linear(x, in_w, in_b))                              # (B,S,3F)
    q, k, v = qkv.split(Fd, dim=-1)
    q = q.view(B, S, num_heads, d).permute(0, 2, 1, 3)           # (B,H,S,d)
    k = k.view(B, S, num_heads, d).permute(0, 2, 1, 3)
    v = v.view(B, S, num_heads, d).permute(0, 2, 1, 3)
    scores = (q @ k.transpose(-1, -2)) / math.sqrt(d)            # (B,H,S,S)
    if key_pad is not None:
        scores = scores.masked_fill(key_pad.view(B, 1, 1, S), float("-inf"))
    m = scores.max(dim=-1, keepdim=True).values
    e = torch.exp(scores - m)
    p = e / e.sum(dim=-1, keepdim=True)
    o = _q((p @ v).permute(0, 2, 1, 3).reshape(B, S, Fd))
    return _q(linear(o, out_w, out_b)), p


def encoder_layer(x: Tensor, P: Dict[str, Tensor], pre: str, key_pad: Optional[Tensor],
                  num_heads: int) -> Tuple[Tensor, Tensor]:
    """Post-norm TransformerEncoderLayer, ReLU, no dropout (p=0 / eval)."""
    a, probs = mha(x, P[pre + "self_attn.in_proj_weight"], P[pre + "self_attn.in_proj_bias"],
                   P[pre + "self_attn.out_proj.weight"], P[pre + "self_attn.out_proj.bias"],
                   key_pad, num_heads)
    x = _q(layer_norm(x + a, P[pre + "norm1.weight"], P[pre + "norm1.bias"]))
    h = _q(torch.relu(linear(x, P[pre + "linear1.weight"], P[pre + "linear1.bias"])))
    f = _q(linear(h, P[pre + "linear2.weight"], P[pre + "linear2.bias"]))
    x = _q(layer_norm(x + f, P[pre + "norm2.weight"], P[pre + "norm2.bias"]))
    return x, probs


def _num_layers(P: Dict[str, Tensor]) -> int:
    n = 0
    while f"fusion.transformer.layers.{n}.norm1.weight" in P:
        n += 1
    return n


def _pool(x: Tensor, full_mask: Optional[Tensor]) -> Tensor:
    """Masked mean pooling, train2.py:184-189 / train.py:100-104."""
    if full_mask is None:
        return x.mean(dim=1)
    keep = (~full_mask).to(x.dtype).unsqueeze(-1)
    return (x * keep).sum(dim=1) / keep.sum(dim=1).clamp(min=1e-6)


# --------------------------------------------------------------------------
# v2 model (train2.py, back-end/app/libs/model.py): LayerNorm variant
# --------------------------------------------------------------------------
def fusion_forward_v2(P: Dict[str, Tensor], video: Tensor, audio: Tensor, mask: Optional[Tensor],
                      num_heads: int = 8) -> Tuple[Tensor, Tensor]:
    """train2.py:128-193 with dropout disabled.  Returns (fused (B,F), attn (L,B,H,S,S))."""
    B, T, _ = video.shape
    v = layer_norm(_q(linear(video, P["fusion.video_proj.weight"], P["fusion.video_proj.bias"])),
                   P["fusion.norm_video.weight"], P["fusion.norm_video.bias"])
    a = layer_norm(_q(linear(audio, P["fusion.audio_proj.weight"], P["fusion.audio_proj.bias"])),
                   P["fusion.norm_audio.weight"], P["fusion.norm_audio.bias"]).unsqueeze(1)
    x = _q(torch.cat([v, a], dim=1) + P["fusion.pos_embed"][:, : T + 1, :])
    full_mask = None
    if mask is not None:
        full_mask = torch.cat([mask, torch.zeros(B, 1, dtype=torch.bool)], dim=1)
    probs = []
    for l in range(_num_layers(P)):
        x, p = encoder_layer(x, P, f"fusion.transformer.layers.{l}.", full_mask, num_heads)
        probs.append(p)
    pooled = _pool(x, full_mask)
    fused = _q(layer_norm(pooled, P["fusion.out_norm.weight"], P["fusion.out_norm.bias"]))
    return fused, torch.stack(probs)


def classifier_forward_v2(P: Dict[str, Tensor], fused: Tensor) -> Tensor:
    """train2.py:217-238, dropout disabled."""
    h = _q(torch.relu(layer_norm(_q(linear(fused, P["classifier.net.0.weight"], P["classifier.net.0.bias"])),
                                 P["classifier.net.1.weight"], P["classifier.net.1.bias"])))
    h = _q(torch.relu(layer_norm(_q(linear(h, P["classifier.net.4.weight"], P["classifier.net.4.bias"])),
                                 P["classifier.net.5.weight"], P["classifier.net.5.bias"])))
    return linear(h, P["classifier.net.8.weight"], P["classifier.net.8.bias"])


def model_forward_v2(P: Dict[str, Tensor], video: Tensor, audio: Tensor, mask: Optional[Tensor],
                     num_heads: int = 8):
    """Returns (probs, logits, fused, attn (L,B,H,S,S)); train2.py:281-292."""
    fused, attn = fusion_forward_v2(P, video, audio, mask, num_heads)
    logits = classifier_forward_v2(P, fused)
    return torch.exp(log_softmax(logits)), logits, fused, attn


# --------------------------------------------------------------------------
# v1 model (train.py): BatchNorm variant
# --------------------------------------------------------------------------
def model_forward_v1(P: Dict[str, Tensor], video: Tensor, audio: Tensor, mask: Optional[Tensor],
                     training: bool, num_heads: int = 8,
                     update: Optional[Dict[str, Tensor]] = None):
    """train.py:64-106,123-130,139-142 with dropout disabled.

    BatchNorm statistics run over ALL (B*T) projected rows, padded rows included
    (train.py:66-69).  Returns (probs, logits, fused, attn).
    """
    B, T, _ = video.shape
    Fd = P["fusion.video_proj.weight"].shape[0]
    pv = linear(video, P["fusion.video_proj.weight"], P["fusion.video_proj.bias"]).reshape(B * T, Fd)
    v = batch_norm(pv, P["fusion.bn_video.weight"], P["fusion.bn_video.bias"],
                   P["fusion.bn_video.running_mean"], P["fusion.bn_video.running_var"],
                   training, update=update, prefix="fusion.bn_video.").reshape(B, T, Fd)
    pa = linear(audio, P["fusion.audio_proj.weight"], P["fusion.audio_proj.bias"])
    a = batch_norm(pa, P["fusion.bn_audio.weight"], P["fusion.bn_audio.bias"],
                   P["fusion.bn_audio.running_mean"], P["fusion.bn_audio.running_var"],
                   training, update=update, prefix="fusion.bn_audio.").unsqueeze(1)
    x = torch.cat([v, a], dim=1) + P["fusion.pos_embed"][:, : T + 1, :]
    full_mask = None
    if mask is not None:
        full_mask = torch.cat([mask, torch.zeros(B, 1, dtype=torch.bool)], dim=1)
    probs = []
    for l in range(_num_layers(P)):
        x, p = encoder_layer(x, P, f"fusion.transformer.layers.{l}.", full_mask, num_heads)
        probs.append(p)
    fused = _pool(x, full_mask)
    h = linear(fused, P["classifier.fc1.weight"], P["classifier.fc1.bias"])
    h = torch.relu(batch_norm(h, P["classifier.bn_fc1.weight"], P["classifier.bn_fc1.bias"],
                              P["classifier.bn_fc1.running_mean"], P["classifier.bn_fc1.running_var"],
                              training, update=update, prefix="classifier.bn_fc1."))
    logits = linear(h, P["classifier.fc2.weight"], P["classifier.fc2.bias"])
    return torch.exp(log_softmax(logits)), logits, fused, torch.stack(probs)


# --------------------------------------------------------------------------
# attention-weight definition for config 4
# --------------------------------------------------------------------------
def cross_modal_attention(attn: Tensor) -> Tuple[Tensor, Tensor]:
    """Last-layer, head-averaged weights (B,S,S) and its audio-query row (B,S)."""
    last = attn[-1].mean(dim=1)
    return last, last[:, -1, :]


# --------------------------------------------------------------------------
# optimiser: optim.Adam(lr, weight_decay) as used at train.py:252 / train2.py:525
# --------------------------------------------------------------------------
def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
              beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
              weight_decay: float = 1e-4) -> Tuple[Tensor, Tensor, Tensor]:
    """One Adam update with coupled L2 (g += wd*p), bias correction, no amsgrad."""
    g = g + weight_decay * p
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = torch.sqrt(v) / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


def clip_coef(grads, max_norm: float = 1.0) -> Tuple[float, float]:
    """torch.nn.utils.clip_grad_norm_ (train2.py:576): returns (total_norm, scale)."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
    coef = max_norm / (total + 1e-6)
    return total, min(coef, 1.0)


# --------------------------------------------------------------------------
# a full training step, used by tests and by bench.py's CPU baseline
# --------------------------------------------------------------------------
PARAM_SKIP = ("running_mean", "running_var", "num_batches_tracked")


def trainable(P: Dict[str, Tensor]) -> Dict[str, Tensor]:
    return {k: v for k, v in P.items() if not k.endswith(PARAM_SKIP)}


def train_step(P: Dict[str, Tensor], state: Dict[str, Dict[str, Tensor]], step: int,
               video: Tensor, audio: Tensor, mask: Optional[Tensor], labels: Tensor, *,
               variant: str = "v2", loss: str = "focal", gamma: float = 2.0,
               alpha: Optional[Tensor] = None, lr: float = 1e-4, weight_decay: float = 1e-4,
               clip: Optional[float] = None, num_heads: int = 8):
    """zero_grad -> forward -> loss -> backward -> [clip] -> Adam.  Returns
    (new_P, new_state, loss_value, logits, grads).  Gradients come from autograd over
    the primitive ops above; Adam and clipping are the explicit formulas.
    """
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in trainable(P).items()}
    full = dict(P)
    full.update(leaf)
    upd: Dict[str, Tensor] = {}
    if variant == "v2":
        _, logits, _, _ = model_forward_v2(full, video, audio, mask, num_heads)
    else:
        _, logits, _, _ = model_forward_v1(full, video, audio, mask, True, num_heads, update=upd)
    if loss == "focal":
        lval = focal_loss(logits, labels, gamma, alpha)
    else:
        lval = weighted_ce(logits, labels, alpha)
    names = list(leaf)
    grads = dict(zip(names, torch.autograd.grad(lval, [leaf[n] for n in names])))
    scale = 1.0
    if clip is not None:
        _, scale = clip_coef(list(grads.values()), clip)
    newP = dict(P)
    newS = {}
    for n in names:
        st = state.get(n) or {"m": torch.zeros_like(P[n]), "v": torch.zeros_like(P[n])}
        p, m, v = adam_step(P[n], grads[n] * scale, st["m"], st["v"], step, lr,
                            weight_decay=weight_decay)
        newP[n] = p
        newS[n] = {"m": m, "v": v}
    newP.update(upd)
    return newP, newS, float(lval.detach()), logits.detach(), grads
