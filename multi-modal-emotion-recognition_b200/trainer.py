"""Fused optimizer and the one-call training step.

``FusedAdam`` mirrors ``torch.optim.Adam(params, lr, weight_decay)`` as used by the reference
(train.py:252, train2.py:525): coupled L2, bias correction, lr read from ``param_groups`` each
step so ``ReduceLROnPlateau`` works unchanged.  ``FusedTrainStep`` is the reference's hot loop
body (train2.py:570-579 / train.py:293-297): zero_grad -> forward -> loss -> backward ->
[clip_grad_norm_] -> Adam, plus the data-parallel gradient all-reduce, with no host sync.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import BF16, MmerError
from .engine import Engine, ParamContext


def allreduce_flat_gradients(flat: torch.Tensor, group=None) -> float:
    """Sum the flat gradient buffer over the data-parallel ranks (NCCL on GPUs, gloo in the CPU
    tests) and return the factor that turns the sum of per-rank mean-loss gradients into the
    global-batch gradient (1 / world size; equal shards)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def allreduce_buckets(flat: torch.Tensor, ranges, group=None) -> float:
    """The same sum as ``allreduce_flat_gradients``, issued bucket by bucket (``ranges`` = [lo, hi) element ranges
    in the order backward completes them).  On GPUs ``OverlappedAllReduce`` launches each bucket as soon as its
    gradients are final; this plain version is the host logic the gloo CPU tests exercise."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    for lo, hi in ranges:
        if hi > lo:
            dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def shard_range(n: int, world: int, rank: int):
    """[lo, hi) of the flat buffer that ``rank`` reduces, updates and broadcasts in the multicast step: equal chunks
    rounded up to 4 elements (one 16-byte multimem access), the last ranks possibly short or empty."""
    chunk = ((n + world - 1) // world + 3) // 4 * 4
    return min(n, rank * chunk), min(n, (rank + 1) * chunk)


class OverlappedAllReduce:
    """NCCL all-reduce of the flat gradient buffer overlapped with backward.

    ``mmer_model_backward`` records one CUDA event per gradient bucket as soon as that bucket is final
    (classifier + out_norm, then the encoder layers from last to first, then the input projections).  A side
    stream waits for event k and all-reduces bucket k over NVLink while the remaining backward kernels run on
    the main stream; Adam waits for the side stream."""

    def __init__(self, ctx: ParamContext, n_layers: int, device, group=None):
        self.ctx, self.group = ctx, group
        self.n = n_layers + 2
        lib = _lib.load()
        self.events = (C.c_void_p * self.n)()
        for k in range(self.n):
            ev = C.c_void_p()
            _lib.check(lib.mmer_event_create(C.byref(ev)), "mmer_event_create")
            self.events[k] = ev
        self.stream = torch.cuda.Stream(device=device)
        self.world = dist.get_world_size(group)

    def attach(self, m) -> None:
        m.grad_events = C.cast(self.events, C.POINTER(C.c_void_p))
        m.n_grad_events = self.n

    def reduce(self) -> float:
        """Call right after mmer_model_backward has been enqueued on the current stream."""
        lib = _lib.load()
        ranges = self.ctx.bucket_ranges()
        assert len(ranges) == self.n
        grads = self.ctx.grads
        side = C.c_void_p(self.stream.cuda_stream)
        for k, (lo, hi) in enumerate(ranges):
            _lib.check(lib.mmer_stream_wait_event(side, self.events[k]), "mmer_stream_wait_event")
            if hi > lo:
                with torch.cuda.stream(self.stream):
                    dist.all_reduce(grads[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        torch.cuda.current_stream().wait_stream(self.stream)
        return 1.0 / self.world

    def __del__(self):
        try:
            lib = _lib.load()
            for ev in self.events:
                lib.mmer_event_destroy(ev)
        except Exception:
            pass


class FusedAdam(torch.optim.Optimizer):
    """Adam over the flat parameter buffer of an mmer_b200 model: one kernel per step."""

    def __init__(self, model_or_params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: Optional[float] = None):
        if isinstance(model_or_params, torch.nn.Module):
            model = model_or_params
            params = list(model.parameters())
            self._owner = model
        else:
            params = list(model_or_params)
            self._owner = None
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self._step = 0
        self._m = self._v = None
        self._sumsq = None
        self._ctx: Optional[ParamContext] = None

    def _context(self) -> ParamContext:
        if self._owner is not None and hasattr(self._owner, "_engine"):
            ctx = self._owner._engine.ctx
        else:
            raise MmerError("FusedAdam needs the mmer_b200 module (pass the model, not model.parameters())")
        ctx.ensure()
        mine = {id(p) for g in self.param_groups for p in g["params"]}
        if mine != {id(p) for p in ctx.params}:
            raise MmerError("FusedAdam must own exactly the parameters of the model")
        return ctx

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        ctx = self._context()
        if self._m is None or getattr(self, "_ctx_flat_ptr", 0) != ctx.flat.data_ptr():
            self._m = torch.zeros_like(ctx.flat)
            self._v = torch.zeros_like(ctx.flat)
            self._sumsq = torch.zeros(1, device=ctx.flat.device, dtype=torch.float32)
            self._ctx_flat_ptr = ctx.flat.data_ptr()
        g = self.param_groups[0]
        self._step += 1
        sumsq = None
        if self.max_grad_norm is not None:
            sumsq = ops.grad_sumsq(ctx.grads, self._sumsq)
        ops.adam_step(ctx.flat, ctx.grads, self._m, self._v, ctx.shadow, self._step, g["lr"], g["betas"][0],
                      g["betas"][1], g["eps"], g["weight_decay"], grad_scale, sumsq, self.max_grad_norm or 0.0)
        ctx.shadow_fresh = ctx.shadow is not None
        return loss

    def zero_grad(self, set_to_none: bool = True):
        # gradients live in one flat buffer: a single memset instead of one per tensor
        ctx = self._context()
        ctx.grads.zero_()
        ctx.attach_grads()


class FusedTrainStep:
    """zero_grad + forward + loss + backward + [gradient exchange] + [clip] + Adam as one host call.

    Under ``torch.distributed`` (one process per GPU) the ranks' gradients are averaged.  ``dp_mode="nvls"`` (what
    "auto" picks on an NVSwitch node) does that inside the optimizer kernel: every rank reduces, updates and broadcasts
    its own 1/world shard of the flat parameter buffer through multicast addresses, so after a step ``param.grad`` still
    holds this rank's LOCAL gradient, Adam's moments exist only for the rank's shard, and all ranks hold bit-identical
    weights.  ``dp_mode="nccl"`` all-reduces the flat gradient buffer (``param.grad`` = sum over ranks; the 1/world factor
    is applied inside Adam) and keeps full optimizer state on every rank."""

    def __init__(self, model, *, lr: float = 1e-4, weight_decay: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 loss: str = "focal", gamma: float = 2.0, alpha: Optional[torch.Tensor] = None,
                 clip_grad_norm: Optional[float] = None, compute_dtype: torch.dtype = torch.bfloat16,
                 process_group=None, overlap_allreduce: bool = True, dp_mode: str = "auto"):
        self.model = model
        self.engine: Engine = model._engine
        self.ctx: ParamContext = self.engine.ctx
        self.opt = FusedAdam(model, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                             max_grad_norm=clip_grad_norm)
        self.loss_kind = {"focal": _lib.LOSS_FOCAL, "wce": _lib.LOSS_WCE}[loss]
        self.gamma = gamma
        self.alpha = alpha
        self.compute_dtype = compute_dtype
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.overlap = overlap_allreduce
        self._overlapped: Optional[OverlappedAllReduce] = None
        # data-parallel gradient exchange: "nvls" = one kernel doing reduce-scatter + Adam on this rank's shard +
        # all-gather over NVSwitch multicast (mmer_adam_step_multicast; with clipping, the norm of the reduced gradient
        # is summed shard-wise and exchanged through a symmetric slot array first); "nccl" = NCCL all-reduce (overlapped
        # with backward or not) + the ordinary Adam kernel; "auto" = nvls when the fabric offers multicast, else nccl
        if dp_mode not in ("auto", "nvls", "nccl"):
            raise ValueError("dp_mode must be auto, nvls or nccl")
        self.dp_mode = "nccl"
        if self.world > 1 and dp_mode != "nccl":
            if self.world > 64:
                if dp_mode == "nvls":
                    raise MmerError("dp_mode='nvls' supports at most 64 ranks")
            elif self.ctx.make_symmetric(process_group):
                self.dp_mode = "nvls"
            elif dp_mode == "nvls":
                raise MmerError("dp_mode='nvls': symmetric memory / NVSwitch multicast is not available here")
        self._key = None
        self._ws = None
        self.launch_count = 0

    @property
    def lr(self):
        return self.opt.param_groups[0]["lr"]

    @lr.setter
    def lr(self, v):
        self.opt.param_groups[0]["lr"] = v

    def _prepare(self, B: int, T: int, dev):
        key = (B, T, self.compute_dtype, self.ctx.flat.data_ptr() if self.ctx.flat is not None else 0)
        if key == self._key:
            return
        m = self.engine.make(B, T, self.compute_dtype, True, self.model._p_fusion, self.model._p_classifier, 0, 0)
        self._ws = torch.empty(Engine.workspace_bytes(m), device=dev, dtype=torch.uint8)
        Cn = self.engine.cfg["classes"]
        self._logits = torch.empty((B, Cn), device=dev, dtype=torch.float32)
        self._probs = torch.empty((B, Cn), device=dev, dtype=torch.float32)
        self._dlogits = torch.empty((B, Cn), device=dev, dtype=torch.float32)
        self._loss = torch.zeros(1, device=dev, dtype=torch.float32)
        self._scratch = torch.zeros(2, device=dev, dtype=torch.float32)
        if self.alpha is not None:
            self.alpha = self.alpha.to(device=dev, dtype=torch.float32).contiguous()
        self._key = (B, T, self.compute_dtype, self.ctx.flat.data_ptr())

    def _nvls_adam(self, stream) -> None:
        """barrier | reduce-scatter + Adam(shard) + all-gather in one kernel over the multicast addresses | barrier."""
        ctx, opt = self.ctx, self.opt
        n = ctx.flat.numel()
        rank = dist.get_rank(self.group)
        lo, hi = shard_range(n, self.world, rank)
        if opt._m is None or getattr(opt, "_ctx_flat_ptr", 0) != ctx.flat.data_ptr():
            opt._m = torch.zeros(n, device=ctx.flat.device, dtype=torch.float32)
            opt._v = torch.zeros(n, device=ctx.flat.device, dtype=torch.float32)
            opt._ctx_flat_ptr = ctx.flat.data_ptr()
        g = opt.param_groups[0]
        opt._step += 1
        p_mc, g_mc, s_mc, slots_mc = ctx.multicast_ptrs()
        use_shadow = ctx.shadow is not None and self.compute_dtype == torch.bfloat16
        lib = _lib.load()
        ctx.sym_hdl.barrier(channel=0)      # every rank's backward has finished: all gradients are complete
        slots, max_norm = None, 0.0
        if opt.max_grad_norm is not None:   # clip_grad_norm_ on the reduced gradient (train2.py:576)
            if opt._sumsq is None:
                opt._sumsq = torch.zeros(1, device=ctx.flat.device, dtype=torch.float32)
            _lib.check(lib.mmer_grad_sumsq_multicast(C.c_void_p(g_mc), lo, hi, opt._sumsq.data_ptr(), C.c_void_p(slots_mc),
                                                     rank, stream), "mmer_grad_sumsq_multicast")
            ctx.sym_hdl.barrier(channel=0)  # every rank's partial sum has landed in everybody's slot array
            slots, max_norm = ctx.sym_slots.data_ptr(), float(opt.max_grad_norm)
        _lib.check(lib.mmer_adam_step_multicast(
            ctx.flat.data_ptr(), C.c_void_p(p_mc), C.c_void_p(g_mc), opt._m.data_ptr(), opt._v.data_ptr(),
            C.c_void_p(s_mc) if use_shadow else None, lo, hi, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
            float(g["eps"]), float(g["weight_decay"]), opt._step, 1.0 / self.world, slots, self.world, max_norm, stream),
            "mmer_adam_step_multicast")
        ctx.sym_hdl.barrier(channel=1)      # every rank's shard has landed everywhere: weights are complete
        ctx.shadow_fresh = use_shadow

    @torch.no_grad()
    def step(self, video: torch.Tensor, audio: torch.Tensor, mask: Optional[torch.Tensor], labels: torch.Tensor):
        """Runs one optimisation step; returns (loss [1] fp32 device tensor, probs (B,C))."""
        if not video.is_cuda:
            raise MmerError("FusedTrainStep needs CUDA tensors (no CPU fallback)")
        B, T = video.shape[0], video.shape[1]
        self.ctx.ensure()
        self._prepare(B, T, video.device)
        eng, ctx = self.engine, self.ctx
        m = eng.make(B, T, self.compute_dtype, True, self.model._p_fusion, self.model._p_classifier,
                     self.model._next_seed(), 0)
        eng.attach_shadow(m)
        v = video if video.dtype == self.compute_dtype else video.to(self.compute_dtype)
        a = audio if audio.dtype == self.compute_dtype else audio.to(self.compute_dtype)
        v, a = v.contiguous(), a.contiguous()
        m.video, m.audio = v.data_ptr(), a.data_ptr()
        if mask is not None:
            mk = mask.contiguous().view(torch.uint8)
            m.mask, m.has_mask = mk.data_ptr(), 1
        m.workspace, m.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        m.logits, m.probs, m.dlogits = self._logits.data_ptr(), self._probs.data_ptr(), self._dlogits.data_ptr()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        lib = _lib.load()
        ctx.grads.zero_()
        _lib.check(lib.mmer_model_forward(C.byref(m), stream), "mmer_model_forward")
        _lib.check(lib.mmer_loss_fwd_bwd(self._logits.data_ptr(), labels.data_ptr(),
                                         self.alpha.data_ptr() if self.alpha is not None else None, self.loss_kind,
                                         float(self.gamma), _lib.REDUCE_MEAN, self._loss.data_ptr(), None,
                                         self._dlogits.data_ptr(), self._scratch.data_ptr(), B,
                                         self.engine.cfg["classes"], 1.0, stream), "mmer_loss_fwd_bwd")
        nvls = self.dp_mode == "nvls" and ctx.multicast_ptrs() is not None
        if self.dp_mode == "nvls" and not nvls:
            raise MmerError("the model's parameters were re-allocated after FusedTrainStep made them symmetric")
        if self.world > 1 and self.overlap and not nvls:
            if self._overlapped is None or self._overlapped.ctx.grads is not ctx.grads:
                self._overlapped = OverlappedAllReduce(ctx, self.engine.cfg["layers"], video.device, self.group)
            self._overlapped.attach(m)
        _lib.check(lib.mmer_model_backward(C.byref(m), stream), "mmer_model_backward")
        if nvls:
            self._nvls_adam(stream)
            return self._loss, self._probs
        if self.world > 1 and self.overlap:
            scale = self._overlapped.reduce()
        else:
            scale = allreduce_flat_gradients(ctx.grads, self.group) if self.world > 1 else 1.0
        self.opt.step(grad_scale=scale)
        return self._loss, self._probs
