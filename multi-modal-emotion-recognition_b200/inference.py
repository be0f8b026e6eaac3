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
    the whole model (``mmer_serve_forward``: csrc/serve_small.cu for up to 8 tokens, csrc/serve.cu beyond) instead of ~30
    launch-bound kernels.

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
        self.packed = None                            # the bf16 weights in MMA-fragment order (mmer_serve_pack)
        self._lib, self._C = lib, C

        def launch(cast: bool = True):
            m = eng.make(1, frames, torch.bfloat16, False, 0.0, 0.0, 0, 0)
            if cast:
                eng.attach_shadow(m)                  # casts the live fp32 weights into the bf16 shadow
            m.shadow = eng.ctx.shadow.data_ptr()
            m.video, m.audio = self.video.data_ptr(), self.audio.data_ptr()
            m.mask, m.has_mask = self.mask.view(torch.uint8).data_ptr(), 1
            m.logits, m.probs = self.logits.data_ptr(), self.probs.data_ptr()
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            if self.packed is None or self.packed.numel() != eng.ctx.shadow.numel():
                self.packed = torch.zeros_like(eng.ctx.shadow)
                cast = True
            if cast:                                  # ... and from the shadow into the fragment-order copy
                _lib.check(lib.mmer_serve_pack(C.byref(m), C.c_void_p(self.packed.data_ptr()), stream), "mmer_serve_pack")
            _lib.check(lib.mmer_serve_forward(C.byref(m), C.c_void_p(self.scratch.data_ptr()),
                                              C.c_void_p(self.packed.data_ptr()), stream), "mmer_serve_forward")

        self._launch = launch
        self.graph = self.graph_cast = None
        self._cast_state = None       # (flat pointer, parameter versions) the shadow was last cast for
        with torch.no_grad():
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):                    # first-call work (function attributes, cluster size probe) outside the graphs
                    launch()
            torch.cuda.current_stream(dev).wait_stream(side)
            if use_graph:
                # two graphs: shadow cast + kernel (after any parameter change), and the kernel alone (steady state)
                self.graph_cast = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_cast):
                    launch(True)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    launch(False)

    def _weights_state(self):
        ctx = self.model._engine.ctx
        return (ctx.flat.data_ptr() if ctx.flat is not None else 0, ctx.shadow.data_ptr() if ctx.shadow is not None else 0,
                ctx.param_versions())

    def refresh(self) -> None:
        """Force a re-cast of the bf16 weights on the next call (needed only after writes through ``param.data``, which
        carry no version counter; ``load_state_dict``, optimizer steps and in-place ops on the parameters are seen)."""
        self._cast_state = None

    def phase_times(self):
        """Nanosecond offsets of the kernel's phase boundaries during the LAST call (stamped by CTA 0 from
        ``%globaltimer``): projections, token assembly, then per layer in_proj / attention / out_proj / norm1+linear1 /
        linear2, then norm2+pooling, head."""
        st = self.scratch.view(torch.int64)[-16 * 4:]           # SC_STAMPS: the last 128 floats = 64 int64
        st = st.cpu().tolist()
        n = int(st[0])
        return [t - st[1] for t in st[1:1 + n]]

    @torch.no_grad()
    def replay(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """The zero-copy form for a server that owns the buffers: write the request into ``self.video`` [1, T, 768],
        ``self.audio`` [1, 1024] (bf16) and ``self.mask`` [1, T] (bool, True = padded chunk) -- e.g. as the output buffers
        of the feature extractors -- then ``probs, logits = run.replay()``.  ONE launch on the stream, no staging copies;
        the returned tensors are the static output buffers (valid until the next call)."""
        fresh = self._weights_state() == self._cast_state
        if self.graph is not None:
            (self.graph if fresh else self.graph_cast).replay()
        else:
            self._launch(not fresh)
        self._cast_state = self._weights_state()
        return self.probs, self.logits

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
        probs, logits = self.replay()
        return probs.clone(), logits.clone()
