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
    """Adam over the flat parameter buffer of an mmer_b200 model: one kernel per step.

    ``state_dict()`` / ``load_state_dict()`` use ``torch.optim.Adam``'s own format (per-parameter ``step``, ``exp_avg``,
    ``exp_avg_sq`` keyed by the parameter's index in ``model.parameters()`` order), so optimizer checkpoints move freely
    between this class and the stock optimizer the reference constructs (train2.py:525).  In the multicast data-parallel
    mode the moments exist for the rank's shard only; ``state_dict()`` is then a COLLECTIVE that gathers them."""

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
        self._mkey = None            # (flat length, lo, hi) the moments were allocated for
        self._shard = None           # (lo, hi, group) in the multicast data-parallel mode, else None
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

    def _moments(self, ctx: ParamContext):
        """Adam's moments for [lo, hi) of the flat buffer (the whole buffer outside the multicast mode).  The flat layout
        depends only on the parameter shapes, so the moments survive a re-allocation of the flat buffer (``model.to()``,
        a sub-module call that re-points ``param.data``); they are reset (with the step count, and a warning) only when
        the layout itself changed."""
        n = ctx.flat.numel()
        lo, hi = (self._shard[0], self._shard[1]) if self._shard is not None else (0, n)
        key, dev = (n, lo, hi), ctx.flat.device
        if self._m is not None and self._mkey != key:
            if self._mkey[0] == n:                 # same layout, different shard (mode switch): re-slice
                full_m, full_v = self._full_moments(ctx)
                self._m, self._v = full_m[lo:hi].clone(), full_v[lo:hi].clone()
            else:
                import warnings
                warnings.warn("FusedAdam: the parameter layout changed; optimizer state and step count are reset")
                self._m = self._v = None
                self._step = 0
        if self._m is None:
            self._m = torch.zeros(hi - lo, device=dev, dtype=torch.float32)
            self._v = torch.zeros(hi - lo, device=dev, dtype=torch.float32)
        elif self._m.device != dev:
            self._m, self._v = self._m.to(dev), self._v.to(dev)
        self._mkey = key
        if self._sumsq is None or self._sumsq.device != dev:
            self._sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
        return self._m, self._v

    def _full_moments(self, ctx: ParamContext):
        """(m, v) over the whole flat buffer; gathers the shards in the multicast mode (collective)."""
        n = self._mkey[0]
        if self._mkey[1:] == (0, n):
            return self._m, self._v
        lo, hi = self._mkey[1:]
        out = []
        for t in (self._m, self._v):
            full = torch.zeros(n, device=t.device, dtype=torch.float32)
            full[lo:hi] = t
            dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self._shard[2] if self._shard is not None else None)
            out.append(full)
        return out[0], out[1]

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        ctx = self._context()
        if self._shard is not None:
            raise MmerError("this optimizer belongs to a multicast data-parallel FusedTrainStep: call its step()")
        m, v = self._moments(ctx)
        g = self.param_groups[0]
        self._step += 1
        sumsq = None
        if self.max_grad_norm is not None:
            sumsq = ops.grad_sumsq(ctx.grads, self._sumsq)
        ops.adam_step(ctx.flat, ctx.grads, m, v, ctx.shadow, self._step, g["lr"], g["betas"][0],
                      g["betas"][1], g["eps"], g["weight_decay"], grad_scale, sumsq, self.max_grad_norm or 0.0)
        ctx.mark_shadow_written()
        return loss

    def zero_grad(self, set_to_none: bool = True):
        # gradients live in one flat buffer: a single memset instead of one per tensor
        ctx = self._context()
        ctx.grads.zero_()
        ctx.attach_grads()

    # ------------------------------------------------------------------ checkpointing (torch.optim.Adam's format)
    def state_dict(self):
        plist = [p for g in self.param_groups for p in g["params"]]
        groups, start = [], 0
        for g in self.param_groups:
            d = {k: v for k, v in g.items() if k != "params"}
            d["params"] = list(range(start, start + len(g["params"])))
            start += len(g["params"])
            groups.append(d)
        state = {}
        if self._m is not None:
            ctx = self._context()
            m, v = self._full_moments(ctx)
            for i, p in enumerate(plist):
                o = ctx.offsets[id(p)]
                state[i] = {"step": torch.tensor(float(self._step)),
                            "exp_avg": m[o:o + p.numel()].view(p.shape).clone(),
                            "exp_avg_sq": v[o:o + p.numel()].view(p.shape).clone()}
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        plist = [p for g in self.param_groups for p in g["params"]]
        for g, sg in zip(self.param_groups, state_dict["param_groups"]):
            if len(sg["params"]) != len(g["params"]):
                raise ValueError("loaded state dict has a parameter group of a different size")
            for k, val in sg.items():
                if k != "params":
                    g[k] = val
        st = state_dict.get("state", {})
        if not st:
            self._m = self._v = None
            self._step = 0
            return
        ctx = self._context()
        n = ctx.flat.numel()
        m = torch.zeros(n, device=ctx.flat.device, dtype=torch.float32)
        v = torch.zeros(n, device=ctx.flat.device, dtype=torch.float32)
        steps = set()
        for i, p in enumerate(plist):
            s = st.get(i, st.get(str(i)))
            if s is None:
                raise ValueError(f"optimizer state for parameter {i} is missing")
            if tuple(s["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"optimizer state {i}: shape {tuple(s['exp_avg'].shape)} != parameter {tuple(p.shape)}")
            o = ctx.offsets[id(p)]
            m[o:o + p.numel()].copy_(s["exp_avg"].reshape(-1))
            v[o:o + p.numel()].copy_(s["exp_avg_sq"].reshape(-1))
            steps.add(int(float(s["step"])))
        if len(steps) != 1:
            raise ValueError("FusedAdam keeps ONE step count for all parameters; the loaded state has several")
        self._step = steps.pop()
        lo, hi = (self._shard[0], self._shard[1]) if self._shard is not None else (0, n)
        self._m, self._v, self._mkey = m[lo:hi].clone(), v[lo:hi].clone(), (n, lo, hi)


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
                 process_group=None, overlap_allreduce: bool = True, dp_mode: str = "auto", check_labels: bool = False):
        self.model = model
        self.check_labels = check_labels
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
        if self.world > 1:
            has_bn = self.engine.cfg["variant"] == 1 or (self.engine.cfg.get("norms", 0) & _lib.NORM_HEAD_BATCHNORM)
            if has_bn and not (self.engine.cfg["variant"] == 1 and getattr(model, "sync_batchnorm", False)):
                raise MmerError("a model with BatchNorm layers (train.py) is not invariant under data parallelism: per-rank batch "
                                "statistics differ from the global-batch reference (SURVEY 8e).  Construct the model with "
                                "sync_batchnorm=True or train it on one GPU")
            # DistributedDataParallel semantics: every replica starts from rank 0's weights (and BatchNorm buffers)
            self.ctx.ensure()
            src = dist.get_global_rank(process_group, 0) if process_group is not None else 0
            dist.broadcast(self.ctx.flat, src=src, group=process_group)
            if self.ctx.bn_state is not None:
                dist.broadcast(self.ctx.bn_state, src=src, group=process_group)
            self.ctx.invalidate_shadow()
            if self.dp_mode == "nvls":
                n = self.ctx.flat.numel()
                lo, hi = shard_range(n, self.world, dist.get_rank(process_group))
                self.opt._shard = (lo, hi, process_group)
        self._key = None
        self._ws = None
        self._trace = None      # measure_barrier_wait: list of CUDA events around the synchronisation points
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
        m_sh, v_sh = opt._moments(ctx)
        g = opt.param_groups[0]
        opt._step += 1
        p_mc, g_mc, s_mc, slots_mc = ctx.multicast_ptrs()
        use_shadow = ctx.shadow is not None and self.compute_dtype == torch.bfloat16
        lib = _lib.load()
        tr = self._trace
        mark = (lambda: tr.append(self._mark())) if tr is not None else (lambda: None)
        mark()                              # [0] backward finished on this rank
        ctx.sym_hdl.barrier(channel=0)      # every rank's backward has finished: all gradients are complete
        mark()                              # [1]
        slots, max_norm = None, 0.0
        if opt.max_grad_norm is not None:   # clip_grad_norm_ on the reduced gradient (train2.py:576)
            _lib.check(lib.mmer_grad_sumsq_multicast(C.c_void_p(g_mc), lo, hi, opt._sumsq.data_ptr(), C.c_void_p(slots_mc),
                                                     rank, stream), "mmer_grad_sumsq_multicast")
            ctx.sym_hdl.barrier(channel=0)  # every rank's partial sum has landed in everybody's slot array
            slots, max_norm = ctx.sym_slots.data_ptr(), float(opt.max_grad_norm)
        mark()                              # [2]
        _lib.check(lib.mmer_adam_step_multicast(
            ctx.flat.data_ptr(), C.c_void_p(p_mc), C.c_void_p(g_mc), m_sh.data_ptr(), v_sh.data_ptr(),
            C.c_void_p(s_mc) if use_shadow else None, lo, hi, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
            float(g["eps"]), float(g["weight_decay"]), opt._step, 1.0 / self.world, slots, self.world, max_norm, stream),
            "mmer_adam_step_multicast")
        mark()                              # [3]
        ctx.sym_hdl.barrier(channel=1)      # every rank's shard has landed everywhere: weights are complete
        mark()                              # [4]
        if use_shadow:
            ctx.mark_shadow_written()
        else:
            ctx.invalidate_shadow()

    @staticmethod
    def _mark():
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    @torch.no_grad()
    def measure_barrier_wait(self, video, audio, mask, labels, steps: int = 10) -> dict:
        """Where a data-parallel step waits (collective; every rank calls it): ``steps`` training steps with CUDA events
        around the cross-rank synchronisation points, averaged per rank and gathered.  In the multicast mode:
        ``pre_barrier_ms`` = wait until EVERY rank has finished backward (rank skew + barrier latency),
        ``clip_ms`` = norm exchange (when clipping), ``adam_ms`` = the reduce-scatter + Adam + all-gather kernel,
        ``post_barrier_ms`` = wait until every rank's shard has landed.  In the NCCL mode: ``exchange_ms`` = from the end
        of backward to the start of Adam (the part of the bucketed all-reduce that backward did not hide)."""
        names = ["pre_barrier_ms", "clip_ms", "adam_ms", "post_barrier_ms"]
        acc = [0.0] * 4
        step_ms = 0.0
        # No host synchronisation inside the loop: the host must stay AHEAD of the device as it does in a real run,
        # otherwise the gaps between the events measure Python launch latency instead of device-side waiting.
        for _ in range(3):
            self.step(video, audio, mask, labels)
        traces = []
        for _ in range(steps):
            self._trace = []
            e0 = self._mark()
            self.step(video, audio, mask, labels)
            e1 = self._mark()
            traces.append((e0, e1, self._trace))
            self._trace = None
        torch.cuda.synchronize()
        for e0, e1, tr in traces:
            step_ms += e0.elapsed_time(e1)
            if len(tr) == 5:
                for i in range(4):
                    acc[i] += tr[i].elapsed_time(tr[i + 1])
            elif len(tr) == 2:
                acc[0] += tr[0].elapsed_time(tr[1])
        mine = [a / steps for a in acc] + [step_ms / steps]
        out = {"mode": self.dp_mode, "steps": steps}
        if self.world > 1:
            t = torch.tensor(mine, device=video.device, dtype=torch.float32)
            allv = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(allv, t, group=self.group)
            rows = [[round(float(x), 4) for x in r.tolist()] for r in allv]
        else:
            rows = [[round(x, 4) for x in mine]]
        keys = (names if self.dp_mode == "nvls" else ["exchange_ms", "_", "_", "_"]) + ["step_ms"]
        for i, k in enumerate(keys):
            if k != "_":
                out[k + "_per_rank"] = [r[i] for r in rows]
        waits = [r[0] + (r[3] if self.dp_mode == "nvls" else 0.0) for r in rows]
        out["wait_ms_mean"], out["wait_ms_max"] = sum(waits) / len(waits), max(waits)
        return out

    @torch.no_grad()
    def step(self, video: torch.Tensor, audio: torch.Tensor, mask: Optional[torch.Tensor], labels: torch.Tensor):
        """Runs one optimisation step; returns (loss [1] fp32 device tensor, probs (B,C)).  Both are PERSISTENT buffers
        that the next ``step`` overwrites: ``.clone()`` (or ``.item()``) them before the next call if they are kept.
        Inputs are validated for shape, dtype and device here, as ``ModelFn`` does.  An out-of-range label never reads
        out of bounds: the loss kernel poisons the loss and the gradients with NaN (torch device-asserts there);
        ``check_labels=True`` raises on the host instead (one host sync per step)."""
        if not video.is_cuda:
            raise MmerError("FusedTrainStep needs CUDA tensors (no CPU fallback)")
        from .modules import _check_inputs
        _check_inputs(self.model.fusion, video, audio, mask)
        dev = video.device
        if audio.device != dev or labels.device != dev:
            raise MmerError("video, audio and labels must live on the same CUDA device")
        B, T = video.shape[0], video.shape[1]
        if labels.dtype != torch.int64 or tuple(labels.shape) != (B,):
            raise MmerError(f"labels must be an int64 tensor of shape ({B},) (train2.py:568), got {labels.dtype} "
                            f"{tuple(labels.shape)}")
        labels = labels.contiguous()
        if self.check_labels:
            lo, hi = int(labels.min()), int(labels.max())      # host sync: debugging aid, off by default
            if lo < 0 or hi >= self.engine.cfg["classes"]:
                raise MmerError(f"label out of range [0, {self.engine.cfg['classes']}): min {lo}, max {hi}")
        self.ctx.ensure()
        self._prepare(B, T, video.device)
        eng, ctx = self.engine, self.ctx
        m = eng.make(B, T, self.compute_dtype, True, self.model._p_fusion, self.model._p_classifier,
                     self.model._next_seed(), 0)
        eng.attach_shadow(m, trust_optimizer=True)
        v = video if video.dtype == self.compute_dtype else video.to(self.compute_dtype)
        a = audio if audio.dtype == self.compute_dtype else audio.to(self.compute_dtype)
        v, a = v.contiguous(), a.contiguous()
        m.video, m.audio = v.data_ptr(), a.data_ptr()
        if mask is not None:
            mk = mask.to(device=dev, dtype=torch.bool).contiguous().view(torch.uint8)
            m.mask, m.has_mask = mk.data_ptr(), 1
        m.workspace, m.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        keep = []
        eng.attach_bn_sync(m, self._ws, self.model, keep)
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
        if self.engine.cfg["variant"] == 1:
            # BatchNorm bookkeeping of a training-mode forward (train.py:66-74,125): the module path does this in forward()
            self.model.fusion._count_batches()
            self.model.classifier.bn_fc1.num_batches_tracked += 1
        elif self.engine.cfg.get("norms", 0) & _lib.NORM_HEAD_BATCHNORM:   # train2's head with use_layernorm=False
            self.model.classifier._count_batches()
        if nvls:
            self._nvls_adam(stream)
            return self._loss, self._probs
        if self._trace is not None:
            self._trace.append(self._mark())
        if self.world > 1 and self.overlap:
            scale = self._overlapped.reduce()
        else:
            scale = allreduce_flat_gradients(ctx.grads, self.group) if self.world > 1 else 1.0
        if self._trace is not None:
            self._trace.append(self._mark())
        self.opt.step(grad_scale=scale)
        return self._loss, self._probs
