"""Latency path for serving: the eval forward of one fixed shape captured once in a CUDA graph.

The served call (back-end/app/libs/inference.py:494-495: ``probs, logits, _ = fusion_model(video, audio, mask=mask)``
on one clip of a few 32-frame chunks) is ~30 small launches; from Python they cost 180-200 us per call, replayed from a
graph 135-145 us (profiles/r01_summary.md section 5).  The graph re-casts the bf16 shadow from the current fp32 weights
on every replay, so ``load_state_dict`` / optimizer steps between calls are picked up.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._lib import MmerError

__all__ = ["GraphedInference"]


class GraphedInference:
    """``run = GraphedInference(model, batch=1, frames=5); probs, logits = run(video, audio, mask)``.

    Inputs of any float dtype are copied into static buffers (``input_dtype``); ``mask`` (True = padded) is optional
    when the graph was built with ``use_mask=True`` (an absent mask means no padding).  The returned tensors are
    copies, valid across later calls."""

    def __init__(self, model, batch: int, frames: int, use_mask: bool = True, input_dtype: torch.dtype = torch.float32,
                 device="cuda"):
        if not hasattr(model, "_engine"):
            raise MmerError("GraphedInference needs a mmer_b200 model")
        dev = torch.device(device)
        model.eval()
        self.model, self.use_mask = model, use_mask
        dv, da = model.fusion.video_proj.in_features, model.fusion.audio_proj.in_features
        self.video = torch.zeros((batch, frames, dv), device=dev, dtype=input_dtype)
        self.audio = torch.zeros((batch, da), device=dev, dtype=input_dtype)
        self.mask = torch.zeros((batch, frames), device=dev, dtype=torch.bool) if use_mask else None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):                       # first-call work (attribute setting, tensor-map cache) stays outside
                model(self.video, self.audio, mask=self.mask)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.probs, self.logits, _ = model(self.video, self.audio, mask=self.mask)

    @torch.no_grad()
    def __call__(self, video: torch.Tensor, audio: torch.Tensor, mask: Optional[torch.Tensor] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
        if video.shape != self.video.shape or audio.shape != self.audio.shape:
            raise MmerError(f"this graph was captured for video {tuple(self.video.shape)} / audio {tuple(self.audio.shape)}")
        if mask is not None and not self.use_mask:
            raise MmerError("this graph was captured without a padding mask")
        self.video.copy_(video, non_blocking=True)
        self.audio.copy_(audio, non_blocking=True)
        if self.use_mask:
            if mask is None:
                self.mask.zero_()
            else:
                self.mask.copy_(mask, non_blocking=True)
        self.graph.replay()
        return self.probs.clone(), self.logits.clone()
