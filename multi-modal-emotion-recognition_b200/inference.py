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

__all__ = ["GraphedInference", "ServingForward"]


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


class ServingForward:
    """The served call -- one clip window, ``probs, logits, _ = fusion_model(video[1, T, 768], audio[1, 1024], mask)``
    (back-end/app/libs/inference.py:494-495, T <= window_size = 5) -- as ONE kernel launch: a thread-block cluster walks
    the whole model (``mmer_serve_forward``, csrc/serve.cu) instead of ~30 launch-bound kernels.

    ``run = ServingForward(model, frames=5); probs, logits = run(video, audio, mask)``.  bf16 weights (the shadow is
    re-cast from the live fp32 parameters inside the captured graph, so ``load_state_dict`` between calls is picked up),
    fp32 residual stream, eval mode.  Supports the LayerNorm (train2.py) model, batch 1, ``frames + 1 <= 16`` tokens;
    other shapes use ``GraphedInference`` / the module call.  Returned tensors are copies."""

    def __init__(self, model, frames: int, device="cuda", use_graph: bool = True):
        import ctypes as C
        from . import _lib
        if not hasattr(model, "_engine") or model._engine.cfg["variant"] != 2:
            raise MmerError("ServingForward needs the mmer_b200 LayerNorm (train2.py) model")
        if not 1 <= frames <= 15:
            raise MmerError("ServingForward serves one clip of 1..15 chunks")
        dev = torch.device(device)
        model.eval()
        self.model, self.frames = model, frames
        eng = model._engine
        dv, da = eng.cfg["video_dim"], eng.cfg["audio_dim"]
        self.video = torch.zeros((1, frames, dv), device=dev, dtype=torch.bfloat16)
        self.audio = torch.zeros((1, da), device=dev, dtype=torch.bfloat16)
        self.mask = torch.zeros((1, frames), device=dev, dtype=torch.bool)
        self.logits = torch.zeros((1, eng.cfg["classes"]), device=dev, dtype=torch.float32)
        self.probs = torch.zeros_like(self.logits)
        lib = _lib.load()
        self.scratch = torch.zeros(int(lib.mmer_serve_scratch_bytes()), device=dev, dtype=torch.uint8)
        self._lib, self._C = lib, C

        def launch():
            m = eng.make(1, frames, torch.bfloat16, False, 0.0, 0.0, 0, 0)
            eng.attach_shadow(m)                      # casts the live fp32 weights (captured in the graph)
            m.video, m.audio = self.video.data_ptr(), self.audio.data_ptr()
            m.mask, m.has_mask = self.mask.view(torch.uint8).data_ptr(), 1
            m.logits, m.probs = self.logits.data_ptr(), self.probs.data_ptr()
            _lib.check(lib.mmer_serve_forward(C.byref(m), C.c_void_p(self.scratch.data_ptr()),
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mmer_serve_forward")

        self._launch = launch
        self.graph = None
        with torch.no_grad():
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):                    # first-call work (function attributes, cluster size probe) outside the graph
                    launch()
            torch.cuda.current_stream(dev).wait_stream(side)
            if use_graph:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    launch()

    @torch.no_grad()
    def __call__(self, video: torch.Tensor, audio: torch.Tensor, mask: Optional[torch.Tensor] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
        if tuple(video.shape) != tuple(self.video.shape) or tuple(audio.shape) != tuple(self.audio.shape):
            raise MmerError(f"this server was built for video {tuple(self.video.shape)} / audio {tuple(self.audio.shape)}")
        self.video.copy_(video, non_blocking=True)
        self.audio.copy_(audio, non_blocking=True)
        if mask is None:
            self.mask.zero_()
        else:
            self.mask.copy_(mask, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._launch()
        return self.probs.clone(), self.logits.clone()
