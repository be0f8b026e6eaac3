"""Drop-in module surface of the reference fusion classifier, backed by the sm_100a engine.

Same class names, constructor signatures, forward signatures, attribute names and
``state_dict`` keys/shapes as the reference (train2.py:40-70,77-292 ==
back-end/app/libs/model.py:6-149 for the LayerNorm variant; train.py:20-142 for the
BatchNorm variant in modules_v1.py).  The stock ``nn`` sub-modules are kept ONLY as
parameter containers -- their ``forward`` is never called; all arithmetic runs in
libmmer_sm100.so.  There is no CPU fallback.
"""
from __future__ import annotations

import itertools
from typing import Optional

import torch
from torch import nn

from . import _lib, ops
from .engine import Engine, ModelFn, ParamContext

_seed_counter = itertools.count(1)


class FocalLoss(nn.Module):
    """Fused focal loss forward + gradient (reference: train.py:20-37 == train2.py:40-70).

    ``alpha`` is an optional per-class weight tensor; the reduction is a plain mean over the
    batch even with ``alpha``, exactly as in the reference.
    """

    def __init__(self, gamma: float = 2.0, alpha=None, reduction: str = "mean"):
        super().__init__()
        self.gamma = gamma
        self.alpha = alpha
        self.reduction = reduction

    def forward(self, inputs: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        return _LossFn.apply(inputs, targets, self.alpha, _lib.LOSS_FOCAL, float(self.gamma), self.reduction)


class WeightedCrossEntropyLoss(nn.Module):
    """``nn.CrossEntropyLoss(weight=w)`` as used by train2.py:523,572, fused forward + gradient."""

    def __init__(self, weight=None):
        super().__init__()
        self.weight = weight

    def forward(self, inputs: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        return _LossFn.apply(inputs, targets, self.weight, _lib.LOSS_WCE, 0.0, "mean")


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, alpha, kind, gamma, reduction):
        red = {"mean": _lib.REDUCE_MEAN, "sum": _lib.REDUCE_SUM}.get(reduction, _lib.REDUCE_NONE)
        x = logits.detach().to(torch.float32).contiguous()
        a = None if alpha is None else alpha.detach().to(device=x.device, dtype=torch.float32).contiguous()
        loss, dlogits = ops.loss_fwd_bwd(x, targets.contiguous(), a, kind, gamma, red, want_grad=True)
        ctx.save_for_backward(dlogits)
        ctx.red, ctx.in_dtype = red, logits.dtype
        return loss if red == _lib.REDUCE_NONE else loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        d = dlogits * (g.reshape(-1, 1) if ctx.red == _lib.REDUCE_NONE else g)
        return d.to(ctx.in_dtype), None, None, None, None, None


class _EngineOwner:
    """Mixin: engine plumbing shared by the module classes."""

    compute_dtype: Optional[torch.dtype] = None   # None: follow the input dtype (fp32 in -> fp32 parity mode)

    def _init_owner(self):
        self.__dict__["_engine_obj"] = None
        self.__dict__["_anchor_t"] = None
        self.__dict__["_base_seed"] = torch.initial_seed() & 0xFFFFFFFF

    @property
    def _engine(self) -> Engine:
        if self.__dict__.get("_engine_obj") is None:
            self.__dict__["_engine_obj"] = self._make_engine()
        return self.__dict__["_engine_obj"]

    @property
    def _anchor(self) -> torch.Tensor:
        dev = next(self.parameters()).device
        a = self.__dict__.get("_anchor_t")
        if a is None or a.device != dev:
            a = torch.zeros(1, device=dev, requires_grad=True)
            self.__dict__["_anchor_t"] = a
        return a

    def _next_seed(self) -> int:
        return (self.__dict__["_base_seed"] << 32) | (next(_seed_counter) & 0xFFFFFFFF)


def _layer_slots(layer: nn.TransformerEncoderLayer):
    return {"IN_W": layer.self_attn.in_proj_weight, "IN_B": layer.self_attn.in_proj_bias,
            "OUT_W": layer.self_attn.out_proj.weight, "OUT_B": layer.self_attn.out_proj.bias,
            "FF1_W": layer.linear1.weight, "FF1_B": layer.linear1.bias,
            "FF2_W": layer.linear2.weight, "FF2_B": layer.linear2.bias,
            "N1_W": layer.norm1.weight, "N1_B": layer.norm1.bias,
            "N2_W": layer.norm2.weight, "N2_B": layer.norm2.bias}


class CrossModalFusion(nn.Module, _EngineOwner):
    """Self-attention fusion of T video tokens and one audio token (train2.py:77-193)."""

    def __init__(self, video_dim: int = 768, audio_dim: int = 1024, fused_dim: int = 512, num_layers: int = 4,
                 num_heads: int = 8, dropout: float = 0.1, max_seq_len: int = 101, use_layernorm: bool = True):
        super().__init__()
        self.video_proj = nn.Linear(video_dim, fused_dim)
        self.audio_proj = nn.Linear(audio_dim, fused_dim)
        # use_layernorm=False: the three norms are nn.Identity (train2.py:104-105,121) -- engine flag NORM_FUSION_IDENTITY
        self.norm_video = nn.LayerNorm(fused_dim) if use_layernorm else nn.Identity()
        self.norm_audio = nn.LayerNorm(fused_dim) if use_layernorm else nn.Identity()
        self.pos_embed = nn.Parameter(torch.zeros(1, max_seq_len, fused_dim))
        nn.init.normal_(self.pos_embed, mean=0.0, std=0.02)
        encoder_layer = nn.TransformerEncoderLayer(d_model=fused_dim, nhead=num_heads, dim_feedforward=4 * fused_dim,
                                                   dropout=dropout, batch_first=False)
        self.transformer = nn.TransformerEncoder(encoder_layer, num_layers=num_layers, enable_nested_tensor=False)
        self.dropout_layer = nn.Dropout(dropout)
        self.out_norm = nn.LayerNorm(fused_dim) if use_layernorm else nn.Identity()
        self.use_layernorm = bool(use_layernorm)
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.dropout = dropout  # float, read by the reference's logging (train2.py:541)
        self._init_owner()

    # engine wiring -----------------------------------------------------------
    def _g_slots(self):
        g = {"POS": self.pos_embed, "WV": self.video_proj.weight, "BV": self.video_proj.bias,
             "WA": self.audio_proj.weight, "BA": self.audio_proj.bias}
        if self.use_layernorm:
            g.update({"NV_W": self.norm_video.weight, "NV_B": self.norm_video.bias,
                      "NA_W": self.norm_audio.weight, "NA_B": self.norm_audio.bias,
                      "ON_W": self.out_norm.weight, "ON_B": self.out_norm.bias})
        return g

    def _norms(self) -> int:
        return 0 if self.use_layernorm else _lib.NORM_FUSION_IDENTITY

    def _l_slots(self):
        return [_layer_slots(l) for l in self.transformer.layers]

    def _dims(self):
        return dict(video_dim=self.video_proj.in_features, audio_dim=self.audio_proj.in_features,
                    fused=self.video_proj.out_features, heads=self.num_heads, layers=self.num_layers,
                    ffn=self.transformer.layers[0].linear1.out_features)

    def _make_engine(self) -> Engine:
        ctx = ParamContext(2, self._g_slots(), self._l_slots())
        return Engine(ctx, variant=2, hidden=8, classes=1, norms=self._norms(), **self._dims())

    @property
    def _p_fusion(self):
        return float(self.dropout)

    _p_classifier = 0.0

    def forward(self, video_feats: torch.Tensor, audio_feats: torch.Tensor, mask: Optional[torch.Tensor] = None,
                return_attn: bool = False):
        """Returns (fused_embedding (B,F), attn_weights).  ``attn_weights`` is None unless
        ``return_attn`` (the reference always returns None, train2.py:179); with it, the
        per-layer softmax weights (L,B,H,S,S) fp32."""
        _check_inputs(self, video_feats, audio_feats, mask)
        fused, _, attn = ModelFn.apply(self._anchor, video_feats, audio_feats, None, self, mask, 1, bool(return_attn))
        return fused, (attn if return_attn else None)


class EmotionClassifier(nn.Module, _EngineOwner):
    """Classifier head: Linear-LN-ReLU-Dropout x2 + Linear, returns logits (train2.py:196-238)."""

    def __init__(self, input_dim: int = 512, num_classes: int = 6, hidden_dim: Optional[int] = None,
                 dropout: float = 0.2, use_layernorm: bool = True):
        super().__init__()
        if hidden_dim is None:
            hidden_dim = input_dim // 2
        # use_layernorm=False: nn.BatchNorm1d in both blocks (train2.py:215) -- engine flag NORM_HEAD_BATCHNORM
        Norm = nn.LayerNorm if use_layernorm else nn.BatchNorm1d
        self.net = nn.Sequential(
            nn.Linear(input_dim, hidden_dim), Norm(hidden_dim), nn.ReLU(inplace=True), nn.Dropout(dropout),
            nn.Linear(hidden_dim, hidden_dim), Norm(hidden_dim), nn.ReLU(inplace=True), nn.Dropout(dropout),
            nn.Linear(hidden_dim, num_classes),
        )
        self.use_layernorm = bool(use_layernorm)
        self.dropout = dropout
        self.hidden_dim = hidden_dim
        self._init_owner()

    def _norms(self) -> int:
        return 0 if self.use_layernorm else _lib.NORM_HEAD_BATCHNORM

    def _bn_buffers(self):
        if self.use_layernorm:
            return []
        n = self.net
        return [(n[1], "running_mean"), (n[1], "running_var"), (n[5], "running_mean"), (n[5], "running_var")]

    def _count_batches(self):
        if self.training and not self.use_layernorm:   # nn.BatchNorm1d bookkeeping (momentum is fixed: not used otherwise)
            self.net[1].num_batches_tracked += 1
            self.net[5].num_batches_tracked += 1

    def _g_slots(self):
        n = self.net
        return {"C0_W": n[0].weight, "C0_B": n[0].bias, "C1_W": n[1].weight, "C1_B": n[1].bias,
                "C4_W": n[4].weight, "C4_B": n[4].bias, "C5_W": n[5].weight, "C5_B": n[5].bias,
                "C8_W": n[8].weight, "C8_B": n[8].bias}

    def _make_engine(self) -> Engine:
        ctx = ParamContext(2, self._g_slots(), [], self._bn_buffers())
        fused = self.net[0].in_features
        heads = fused // 64 if fused % 64 == 0 else max(fused // 32, 1)   # unused by the head, must be valid
        return Engine(ctx, variant=2, video_dim=8, audio_dim=8, fused=fused, heads=heads, layers=1,
                      ffn=8, hidden=self.hidden_dim, classes=self.net[8].out_features, norms=self._norms())

    _p_fusion = 0.0

    @property
    def _p_classifier(self):
        return float(self.dropout)

    def forward(self, fused_embedding: torch.Tensor) -> torch.Tensor:
        if fused_embedding.dim() != 2 or fused_embedding.shape[1] != self.net[0].in_features:
            raise RuntimeError(f"expected fused embedding of shape (B, {self.net[0].in_features}), got "
                               f"{tuple(fused_embedding.shape)}")
        logits, _, _ = ModelFn.apply(self._anchor, None, None, fused_embedding, self, None, 2, False)
        self._count_batches()
        return logits


class MultimodalEmotionModel(nn.Module, _EngineOwner):
    """Fusion module + classifier head (train2.py:241-292, back-end/app/libs/model.py:114-149).

    ``forward(video_feats, audio_feats, mask=None, return_attn=False) -> (probs, logits, attn_weights)``.
    ``attn_weights`` is None unless ``return_attn=True``; then it is a dict with the defined
    attention outputs of SURVEY.md section 8a row A9: ``"layers"`` (L,B,H,S,S), ``"last_mean"``
    (B,S,S) head-averaged last layer, ``"audio_row"`` (B,S) = attention of the audio token.
    Set ``model.compute_dtype = torch.bfloat16`` to run bf16 tensor-core GEMMs on fp32 inputs.
    """

    def __init__(self, video_dim: int = 768, audio_dim: int = 1024, fused_dim: int = 512, num_classes: int = 6,
                 max_seq_len: int = 101, fusion_num_layers: int = 2, fusion_num_heads: int = 8,
                 fusion_dropout: float = 0.1, classifier_hidden_dim: Optional[int] = None,
                 classifier_dropout: float = 0.2):
        super().__init__()
        self.fusion = CrossModalFusion(video_dim=video_dim, audio_dim=audio_dim, fused_dim=fused_dim,
                                       num_layers=fusion_num_layers, num_heads=fusion_num_heads,
                                       dropout=fusion_dropout, max_seq_len=max_seq_len)
        self.classifier = EmotionClassifier(input_dim=fused_dim, num_classes=num_classes,
                                            hidden_dim=classifier_hidden_dim, dropout=classifier_dropout)
        self._init_owner()

    def _make_engine(self) -> Engine:
        g = dict(self.fusion._g_slots())
        g.update(self.classifier._g_slots())
        # sub-modules swapped for their use_layernorm=False variants keep working: the flags travel with them
        ctx = ParamContext(2, g, self.fusion._l_slots(), self.classifier._bn_buffers())
        return Engine(ctx, variant=2, hidden=self.classifier.hidden_dim, classes=self.classifier.net[8].out_features,
                      norms=self.fusion._norms() | self.classifier._norms(), **self.fusion._dims())

    @property
    def _p_fusion(self):
        return float(self.fusion.dropout)

    @property
    def _p_classifier(self):
        return float(self.classifier.dropout)

    def forward(self, video_feats: torch.Tensor, audio_feats: torch.Tensor, mask: Optional[torch.Tensor] = None,
                return_attn: bool = False):
        _check_inputs(self.fusion, video_feats, audio_feats, mask)
        logits, probs, attn = ModelFn.apply(self._anchor, video_feats, audio_feats, None, self, mask, 0,
                                            bool(return_attn))
        self.classifier._count_batches()
        return probs, logits, (attention_outputs(attn) if return_attn else None)


def attention_outputs(attn: torch.Tensor):
    last = attn[-1].mean(dim=1)
    return {"layers": attn, "last_mean": last, "audio_row": last[:, -1, :]}


def _check_inputs(fusion, video, audio, mask):
    if video.dim() != 3 or audio.dim() != 2:
        raise RuntimeError(f"expected video (B,T,Dv) and audio (B,Da), got {tuple(video.shape)} and {tuple(audio.shape)}")
    b, t, dv = video.shape
    if dv != fusion.video_proj.in_features or audio.shape[1] != fusion.audio_proj.in_features:
        raise RuntimeError("feature dimension mismatch with video_proj / audio_proj")
    if audio.shape[0] != b:
        raise RuntimeError("video and audio batch sizes differ")
    if t + 1 > fusion.pos_embed.size(1):
        # same failure the reference hits at `combined + self.pos_embed[:, :t + 1, :]` (train2.py:160)
        raise RuntimeError(f"The size of tensor a ({t + 1}) must match the size of tensor b "
                           f"({fusion.pos_embed.size(1)}) at non-singleton dimension 1")
    if mask is not None and tuple(mask.shape) != (b, t):
        raise RuntimeError(f"mask must have shape ({b}, {t})")
