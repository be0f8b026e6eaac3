"""The train.py variant of the model (BatchNorm1d projections, 4 encoder layers, 2-layer head
with softmax inside): reference train.py:39-142.  Same surface and state_dict as the reference."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .engine import Engine, ModelFn, ParamContext
from .modules import FocalLoss, _EngineOwner, _check_inputs, _layer_slots, attention_outputs  # noqa: F401


class CrossModalFusion(nn.Module, _EngineOwner):
    """train.py:39-106.  BatchNorm statistics run over all B*T projected rows, padded ones included."""

    def __init__(self, video_dim=768, audio_dim=1024, fused_dim=512, num_layers=4, num_heads=8, dropout=0.01,
                 max_seq_len=101):
        super().__init__()
        self.video_proj = nn.Linear(video_dim, fused_dim)
        self.audio_proj = nn.Linear(audio_dim, fused_dim)
        self.bn_video = nn.BatchNorm1d(fused_dim)
        self.bn_audio = nn.BatchNorm1d(fused_dim)
        self.pos_embed = nn.Parameter(torch.randn(1, max_seq_len, fused_dim))
        self.transformer = nn.TransformerEncoder(
            nn.TransformerEncoderLayer(d_model=fused_dim, nhead=num_heads, dim_feedforward=2048, dropout=dropout),
            num_layers=num_layers, enable_nested_tensor=False)
        self.pool = nn.AdaptiveAvgPool1d(1)
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.dropout = dropout
        self._init_owner()

    def _g_slots(self):
        return {"POS": self.pos_embed, "WV": self.video_proj.weight, "BV": self.video_proj.bias,
                "WA": self.audio_proj.weight, "BA": self.audio_proj.bias,
                "NV_W": self.bn_video.weight, "NV_B": self.bn_video.bias,
                "NA_W": self.bn_audio.weight, "NA_B": self.bn_audio.bias}

    def _l_slots(self):
        return [_layer_slots(l) for l in self.transformer.layers]

    def _bn_buffers(self):
        return [(self.bn_video, "running_mean"), (self.bn_video, "running_var"),
                (self.bn_audio, "running_mean"), (self.bn_audio, "running_var")]

    def _dims(self):
        return dict(video_dim=self.video_proj.in_features, audio_dim=self.audio_proj.in_features,
                    fused=self.video_proj.out_features, heads=self.num_heads, layers=self.num_layers,
                    ffn=self.transformer.layers[0].linear1.out_features)

    def _make_engine(self) -> Engine:
        # the engine's bn_state layout always has the head's BatchNorm last; give it a dummy slot
        dummy = nn.BatchNorm1d(8)
        self.__dict__["_dummy_bn"] = dummy.to(self.pos_embed.device)
        bufs = self._bn_buffers() + [(self.__dict__["_dummy_bn"], "running_mean"), (self.__dict__["_dummy_bn"], "running_var")]
        ctx = ParamContext(1, self._g_slots(), self._l_slots(), bufs)
        return Engine(ctx, variant=1, hidden=8, classes=1, **self._dims())

    @property
    def _p_fusion(self):
        return float(self.dropout)

    _p_classifier = 0.0

    def _count_batches(self):
        if self.training:
            self.bn_video.num_batches_tracked += 1
            self.bn_audio.num_batches_tracked += 1

    def forward(self, video_feats, audio_feats, mask=None, return_attn=False):
        _check_inputs(self, video_feats, audio_feats, mask)
        fused, _, attn = ModelFn.apply(self._anchor, video_feats, audio_feats, None, self, mask, 1, bool(return_attn))
        self._count_batches()
        return fused, (attn if return_attn else None)


class EmotionClassifier(nn.Module, _EngineOwner):
    """train.py:108-130: fc1 -> BatchNorm -> ReLU -> dropout -> fc2 -> softmax; returns (probs, logits)."""

    def __init__(self, input_dim=512, num_classes=6, dropout=0.01):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, input_dim // 2)
        self.bn_fc1 = nn.BatchNorm1d(input_dim // 2)
        self.dropout1 = nn.Dropout(dropout)
        self.fc2 = nn.Linear(input_dim // 2, num_classes)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout = dropout
        self._init_owner()

    def _g_slots(self):
        return {"C0_W": self.fc1.weight, "C0_B": self.fc1.bias, "C1_W": self.bn_fc1.weight, "C1_B": self.bn_fc1.bias,
                "C8_W": self.fc2.weight, "C8_B": self.fc2.bias}

    def _bn_buffers(self):
        return [(self.bn_fc1, "running_mean"), (self.bn_fc1, "running_var")]

    def forward(self, fused_embedding):
        raise NotImplementedError("call the head through MultimodalEmotionModel (train.py:139-142); the standalone "
                                  "BatchNorm head has no separate CUDA entry")


class MultimodalEmotionModel(nn.Module, _EngineOwner):
    """train.py:133-142.  forward -> (probs, logits, attn_weights)."""

    def __init__(self, video_dim=768, audio_dim=1024, fused_dim=512, num_classes=6, max_seq_len=101, *,
                 sync_batchnorm: bool = False, sync_group=None):
        """``sync_batchnorm`` (keyword-only, not in the reference signature): under ``torch.distributed`` the three
        BatchNorm layers take their batch statistics over ALL replicas (equal shards), so that a data-parallel run
        computes what the reference's single process computes on the global batch (train.py:66-74,125)."""
        super().__init__()
        self.fusion = CrossModalFusion(video_dim, audio_dim, fused_dim, dropout=0.01, max_seq_len=max_seq_len)
        self.classifier = EmotionClassifier(fused_dim, num_classes, dropout=0.01)
        self.sync_batchnorm = bool(sync_batchnorm)
        self.sync_group = sync_group
        self._init_owner()

    def _make_engine(self) -> Engine:
        g = dict(self.fusion._g_slots())
        g.update(self.classifier._g_slots())
        ctx = ParamContext(1, g, self.fusion._l_slots(), self.fusion._bn_buffers() + self.classifier._bn_buffers())
        return Engine(ctx, variant=1, hidden=self.classifier.fc1.out_features,
                      classes=self.classifier.fc2.out_features, **self.fusion._dims())

    @property
    def _p_fusion(self):
        return float(self.fusion.dropout)

    @property
    def _p_classifier(self):
        return float(self.classifier.dropout)

    def forward(self, video_feats, audio_feats, mask=None, return_attn=False):
        _check_inputs(self.fusion, video_feats, audio_feats, mask)
        logits, probs, attn = ModelFn.apply(self._anchor, video_feats, audio_feats, None, self, mask, 0,
                                            bool(return_attn))
        if self.training:
            self.fusion._count_batches()
            self.classifier.bn_fc1.num_batches_tracked += 1
        return probs, logits, (attention_outputs(attn) if return_attn else None)
