"""Host-side engine: flat parameter storage + one-call forward/backward through the C ABI.

The module classes (modules.py) keep stock ``nn.Linear`` / ``nn.LayerNorm`` /
``nn.TransformerEncoder`` objects purely as parameter containers, so ``state_dict()`` has the
reference's key names and shapes.  At run time the parameters are
views into ONE flat fp32 buffer (plus a flat fp32 gradient buffer and, in bf16 mode, a flat
bf16 shadow); the C engine addresses them by element offset.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _lib, ops
from ._lib import BF16, F32, G, L, MmerError, Model

PAD = 64  # every parameter starts on a 256-byte boundary of the fp32 buffer

# parameters whose gradients become final LAST in backward (token assembly and the two input projections)
EMBED_SLOTS = ("POS", "WV", "BV", "WA", "BA", "NV_W", "NV_B", "NA_W", "NA_B")


def _pad(n: int) -> int:
    return (n + PAD - 1) // PAD * PAD


class ParamContext:
    """Flat master / gradient / shadow storage for the parameters of one module tree."""

    def __init__(self, variant: int, g_slots: Dict[str, nn.Parameter], l_slots: List[Dict[str, nn.Parameter]],
                 bn_buffers: Optional[List[Tuple[nn.Module, str]]] = None):
        self.variant = variant
        self.g_slots = g_slots
        self.l_slots = l_slots
        self.bn_buffers = bn_buffers or []   # [(module, 'running_mean'), (module, 'running_var'), ...] in engine order
        # flat layout = reverse order of gradient completion in backward:  embed | layer 0 | ... | layer L-1 | head.
        # Each group is one contiguous range, so the data-parallel all-reduce can go out bucket by bucket while
        # backward is still running (see bucket_ranges / mmer_model.grad_events).
        self.embed_params = [p for n, p in g_slots.items() if n in EMBED_SLOTS]
        self.layer_params = [list(d.values()) for d in l_slots]
        self.head_params = [p for n, p in g_slots.items() if n not in EMBED_SLOTS]
        self.params: List[nn.Parameter] = self.embed_params + [p for d in self.layer_params for p in d] + self.head_params
        self.flat: Optional[torch.Tensor] = None
        self.grads: Optional[torch.Tensor] = None
        self.shadow: Optional[torch.Tensor] = None
        self.bn_state: Optional[torch.Tensor] = None
        self.offsets: Dict[int, int] = {}
        self._ptrs: List[int] = []
        self.shadow_fresh = False

    # ------------------------------------------------------------------ layout
    def layout(self) -> Tuple[Dict[int, int], int]:
        """Element offset of every parameter in the flat buffers, and the buffer length (device independent)."""
        off = 0
        offsets: Dict[int, int] = {}
        for p in self.params:
            if id(p) in offsets:
                continue
            if p.dtype != torch.float32:
                raise MmerError("parameters must be float32 masters (bf16 compute uses an internal shadow copy)")
            offsets[id(p)] = off
            off += _pad(p.numel())
        return offsets, off

    def bucket_ranges(self) -> List[Tuple[int, int]]:
        """[lo, hi) element ranges of the flat gradient buffer in the order backward completes them:
        head (classifier + out_norm), layer L-1, ..., layer 0, embed.  They tile the buffer exactly."""
        offsets, total = self.layout()
        starts = []
        for group in [self.embed_params] + self.layer_params + [self.head_params]:
            starts.append(min((offsets[id(p)] for p in group), default=None))
        # an empty group (e.g. a classifier-only module tree) collapses onto its successor
        bounds = []
        nxt = total
        for st in reversed(starts):
            lo = nxt if st is None else st
            bounds.append((lo, nxt))
            nxt = lo
        return bounds   # already in completion order: head first, embed last

    def _build(self, device: torch.device) -> None:
        offsets, off = self.layout()
        flat = torch.zeros(off, device=device, dtype=torch.float32)
        grads = torch.zeros(off, device=device, dtype=torch.float32)
        with torch.no_grad():
            for p in self.params:
                o = offsets[id(p)]
                view = flat[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                gview = grads[o:o + p.numel()].view(p.shape)
                if p.grad is not None:
                    gview.copy_(p.grad)
                    p.grad = gview
            if self.bn_buffers:
                n = sum(getattr(m, name).numel() for m, name in self.bn_buffers)
                self.bn_state = torch.zeros(n, device=device, dtype=torch.float32)
                o = 0
                for m, name in self.bn_buffers:
                    buf = getattr(m, name)
                    v = self.bn_state[o:o + buf.numel()]
                    v.copy_(buf)
                    m._buffers[name] = v
                    o += buf.numel()
        self.flat, self.grads, self.offsets, self.shadow = flat, grads, offsets, None
        self.shadow_fresh = False
        self._ptrs = [p.data_ptr() for p in self.params] + [getattr(m, n).data_ptr() for m, n in self.bn_buffers]

    def make_symmetric(self, group) -> bool:
        """Data-parallel jobs: move master weights, gradients and the bf16 shadow into ONE symmetric-memory allocation
        (same layout on every rank: flat fp32 | grads fp32 | shadow bf16) and rendezvous it over ``group``, so that the
        fused reduce-scatter + Adam + all-gather kernel (``mmer_adam_step_multicast``) can address all ranks through the
        NVSwitch multicast pointer.  Collective: every rank of the group must call it.  Returns False (and changes
        nothing) when symmetric memory or multicast is not available on this system."""
        import torch.distributed as dist
        self.ensure()
        n = self.flat.numel()
        dev = self.flat.device
        ok = 1
        sym = hdl = None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            sym = symm_mem.empty(2 * n + n // 2 + 64, dtype=torch.float32, device=dev)   # + 64 slots for the clip norm
            hdl = symm_mem.rendezvous(sym, group if group is not None else dist.group.WORLD)
            if int(hdl.multicast_ptr) == 0:
                ok = 0
        except Exception:   # no symmetric-memory support in this build / on this fabric
            ok = 0
        flag = torch.tensor([ok], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            return False
        flat, grads = sym[:n], sym[n:2 * n]
        shadow = sym[2 * n:2 * n + n // 2].view(torch.bfloat16)
        self.sym_slots = sym[2 * n + n // 2:]
        self.sym_slots.zero_()
        with torch.no_grad():
            flat.copy_(self.flat)
            grads.copy_(self.grads)
            for p in self.params:
                o = self.offsets[id(p)]
                p.data = flat[o:o + p.numel()].view(p.shape)
                if p.grad is not None:
                    p.grad = grads[o:o + p.numel()].view(p.shape)
        self.flat, self.grads, self.shadow = flat, grads, shadow
        self.shadow_fresh = False
        self._ptrs = [p.data_ptr() for p in self.params] + [getattr(m, nm).data_ptr() for m, nm in self.bn_buffers]
        self.sym, self.sym_hdl = sym, hdl
        self.sym_flat_ptr = flat.data_ptr()
        return True

    def multicast_ptrs(self):
        """(weights, gradients, shadow, clip-norm slots) multicast addresses, or None when the storage is not (or no
        longer) symmetric."""
        hdl = getattr(self, "sym_hdl", None)
        if hdl is None or self.flat is None or self.flat.data_ptr() != self.sym_flat_ptr:
            return None
        n = self.flat.numel()
        # the tensor may sit at an offset inside the rendezvoused block: same offset in the multicast mapping
        base = int(hdl.multicast_ptr) + (self.sym.data_ptr() - int(hdl.buffer_ptrs[hdl.rank]))
        return base, base + 4 * n, base + 8 * n, base + 10 * n

    def ensure(self) -> None:
        """(Re)build the flat storage if parameters were moved or replaced (``.to()``, ``.cuda()``...)."""
        dev = self.params[0].device
        if dev.type != "cuda":
            raise MmerError("mmer_b200 modules run on CUDA only (no CPU fallback): call model.cuda() first")
        _lib.bind_device(dev.index if dev.index is not None else torch.cuda.current_device())
        ptrs = [p.data_ptr() for p in self.params] + [getattr(m, n).data_ptr() for m, n in self.bn_buffers]
        if self.flat is None or ptrs != self._ptrs or self.flat.device != dev:
            self._build(dev)

    def grad_view(self, p: nn.Parameter) -> torch.Tensor:
        o = self.offsets[id(p)]
        return self.grads[o:o + p.numel()].view(p.shape)

    def bind_grads(self) -> bool:
        """Make every ``param.grad`` a view of the flat gradient buffer.  Returns True when the
        buffer must be zeroed first (all grads were None: the usual ``zero_grad()`` state)."""
        none = [p.grad is None for p in self.params]
        if all(none):
            return True
        for p in self.params:
            gv = self.grad_view(p)
            if p.grad is None:
                gv.zero_()
            elif p.grad.data_ptr() != gv.data_ptr():
                gv.copy_(p.grad)
        return False

    def attach_grads(self) -> None:
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.grads.data_ptr() + 4 * self.offsets[id(p)]:
                p.grad = self.grad_view(p)

    def param_versions(self) -> int:
        """Sum of the autograd version counters of the parameters: changes on ``load_state_dict``, a stock
        ``torch.optim`` step or any other in-place write through the Parameter objects.  (Writes through ``p.data``
        carry their own counter and are NOT seen: call ``invalidate_shadow()`` after those.)"""
        return sum(p._version for p in self.params)

    def mark_shadow_written(self) -> None:
        """The optimizer kernel has just written bf16(weights) into the shadow."""
        self.shadow_fresh = self.shadow is not None
        self._shadow_versions = self.param_versions()

    def invalidate_shadow(self) -> None:
        self.shadow_fresh = False

    def refresh_shadow(self, trust_optimizer: bool = False) -> None:
        """Make the bf16 shadow equal bf16(fp32 masters).  The cast kernel (one 47 MB pass, ~8 us) runs on EVERY call
        unless ``trust_optimizer`` is set (only FusedTrainStep's own loop does that) AND the last writer of the shadow
        was the fused Adam kernel AND no parameter was modified in place since (``load_state_dict``, ``torch.optim``).
        Evaluation, attribution and graph-captured forwards therefore always read the live weights."""
        if self.shadow is None:
            self.shadow = torch.empty(self.flat.numel(), device=self.flat.device, dtype=torch.bfloat16)
            self.shadow_fresh = False
        if not (trust_optimizer and self.shadow_fresh and getattr(self, "_shadow_versions", -1) == self.param_versions()):
            ops.cast_bf16(self.flat, self.shadow)
            self.shadow_fresh = False

    # ------------------------------------------------------------------ C struct
    def fill_offsets(self, m: Model) -> None:
        for i in range(_lib.G_COUNT):
            m.off_g[i] = -1
        for name, p in self.g_slots.items():
            m.off_g[G[name]] = self.offsets[id(p)]
        for l, d in enumerate(self.l_slots):
            for name, p in d.items():
                m.off_l[l][L[name]] = self.offsets[id(p)]
        m.n_params = self.flat.numel()


class Engine:
    """Builds the ``mmer_model`` struct for one call and runs the C forward / backward."""

    def __init__(self, ctx: ParamContext, *, variant: int, video_dim: int, audio_dim: int, fused: int, heads: int,
                 layers: int, ffn: int, hidden: int, classes: int, norms: int = 0):
        self.ctx = ctx
        self.cfg = dict(variant=variant, video_dim=video_dim, audio_dim=audio_dim, fused=fused, heads=heads,
                        layers=layers, ffn=ffn, hidden=hidden, classes=classes, norms=norms)
        if layers > _lib.MAX_LAYERS:
            raise MmerError(f"at most {_lib.MAX_LAYERS} encoder layers are supported")

    def make(self, B: int, T: int, dtype: torch.dtype, training: bool, p_fusion: float, p_classifier: float, seed: int,
             stage: int = 0) -> Model:
        self.ctx.ensure()
        m = Model()
        for k, v in self.cfg.items():
            setattr(m, k, v)
        m.dtype = BF16 if dtype == torch.bfloat16 else F32
        m.B, m.T = B, T
        m.training = int(training)
        m.p_fusion, m.p_classifier = float(p_fusion), float(p_classifier)
        m.seed = seed & 0xFFFFFFFFFFFFFFFF
        m.stage = stage
        self.ctx.fill_offsets(m)
        m.params = self.ctx.flat.data_ptr()
        m.grads = self.ctx.grads.data_ptr()
        if self.ctx.bn_state is not None:
            m.bn_state = self.ctx.bn_state.data_ptr()
        return m

    def attach_bn_sync(self, m: Model, ws: torch.Tensor, owner, keep: list) -> None:
        """SyncBatchNorm (train.py variant under data parallelism): when the owning module was built with
        ``sync_batchnorm=True`` and a process group of more than one rank is up, every BatchNorm layer of the C engine
        calls back here (on the host, while it enqueues its kernels) with a buffer of partial column sums inside the
        workspace; the callback enqueues an in-place SUM all-reduce of that buffer on the current stream.  Nine tiny
        all-reduces per training step (two per layer in forward, one in backward); none in eval mode."""
        import torch.distributed as dist
        if self.cfg["variant"] != 1 or not getattr(owner, "sync_batchnorm", False):
            return
        if not (dist.is_available() and dist.is_initialized()):
            return
        group = getattr(owner, "sync_group", None)
        world = dist.get_world_size(group)
        if world <= 1:
            return
        base, nbytes = ws.data_ptr(), ws.numel()

        def cb(_user, buf, n, _stream):
            try:
                off = int(buf) - base
                if off < 0 or off + 4 * n > nbytes:
                    return -2
                dist.all_reduce(ws[off:off + 4 * n].view(torch.float32), op=dist.ReduceOp.SUM, group=group)
                return 0
            except Exception:      # never unwind through the C frames
                return -1

        fn = _lib.BN_SYNC_FN(cb)
        keep.append(fn)            # the ctypes thunk must outlive every call that may invoke it
        m.bn_sync = C.cast(fn, C.c_void_p)
        m.bn_world = world

    @staticmethod
    def workspace_bytes(m: Model) -> int:
        n = _lib.load().mmer_workspace_bytes(C.byref(m))
        if n < 0:
            raise MmerError("mmer_workspace_bytes: " + _lib.last_error())
        return int(n)

    def attach_shadow(self, m: Model, trust_optimizer: bool = False) -> None:
        if m.dtype == BF16:
            self.ctx.refresh_shadow(trust_optimizer)
            m.shadow = self.ctx.shadow.data_ptr()

    @staticmethod
    def forward(m: Model) -> None:
        _lib.check(_lib.load().mmer_model_forward(C.byref(m), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "mmer_model_forward")

    @staticmethod
    def backward(m: Model) -> None:
        _lib.check(_lib.load().mmer_model_backward(C.byref(m), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "mmer_model_backward")


def _compute_dtype(video: torch.Tensor, requested: Optional[torch.dtype]) -> torch.dtype:
    if requested is not None:
        return requested
    return torch.bfloat16 if video.dtype == torch.bfloat16 else torch.float32


class ModelFn(torch.autograd.Function):
    """Autograd node for a whole forward pass of the engine (stage 0, 1 or 2).

    Parameter gradients are accumulated straight into the flat gradient buffer that
    ``param.grad`` views alias (the Megatron ``main_grad`` convention); the node returns
    gradients only for the data inputs (needed by Captum-style attribution, train2.py:808-836).
    """

    @staticmethod
    def forward(ctx, anchor, video, audio, fused_in, owner, mask, stage, return_attn):
        eng: Engine = owner._engine
        dev = anchor.device
        training = owner.training
        if stage == 2:
            B, T = fused_in.shape[0], 1
            cdt = _compute_dtype(fused_in, owner.compute_dtype)
        else:
            B, T = video.shape[0], video.shape[1]
            cdt = _compute_dtype(video, owner.compute_dtype)
        seed = owner._next_seed() if training else 0
        m = eng.make(B, T, cdt, training, owner._p_fusion, owner._p_classifier, seed, stage)
        eng.attach_shadow(m)
        ws = torch.empty(eng.workspace_bytes(m), device=dev, dtype=torch.uint8)
        m.workspace, m.workspace_bytes = ws.data_ptr(), ws.numel()
        keep = [ws]
        if training:
            eng.attach_bn_sync(m, ws, owner, keep)
        if stage != 2:
            v = video.detach().to(cdt).contiguous()
            a = audio.detach().to(cdt).contiguous()
            m.video, m.audio = v.data_ptr(), a.data_ptr()
            keep += [v, a]
            if mask is not None:
                mk = mask.to(device=dev, dtype=torch.bool).contiguous().view(torch.uint8)
                m.mask, m.has_mask = mk.data_ptr(), 1
                keep.append(mk)
        else:
            f = fused_in.detach().to(cdt).contiguous()
            m.fused_in = f.data_ptr()
            keep.append(f)
        F_, Cn = eng.cfg["fused"], eng.cfg["classes"]
        outs = []
        logits = probs = fused = attn = None
        if stage != 1:
            logits = torch.empty((B, Cn), device=dev, dtype=torch.float32)
            probs = torch.empty((B, Cn), device=dev, dtype=torch.float32)
            m.logits, m.probs = logits.data_ptr(), probs.data_ptr()
        if stage == 1:
            fused = torch.empty((B, F_), device=dev, dtype=cdt)
            m.fused_out = fused.data_ptr()
        if return_attn and stage != 2:
            S = T + 1
            attn = torch.empty((eng.cfg["layers"], B, eng.cfg["heads"], S, S), device=dev, dtype=torch.float32)
            m.attn_probs = attn.data_ptr()
        Engine.forward(m)
        m.attn_probs = None
        m.fused_out = None
        ctx.m, ctx.keep, ctx.owner, ctx.stage, ctx.cdt = m, keep, owner, stage, cdt
        ctx.in_dtypes = (video.dtype if video is not None else None, audio.dtype if audio is not None else None,
                         fused_in.dtype if fused_in is not None else None)
        ctx.mark_non_differentiable(*[t for t in (probs, attn) if t is not None])
        if stage == 1:
            out_main = fused if fused.dtype == video.dtype else fused.to(video.dtype)
        else:
            out_main = logits
        empty = torch.empty(0, device=dev)
        return out_main, (probs if probs is not None else empty), (attn if attn is not None else empty)

    @staticmethod
    def backward(ctx, d_main, _dp, _da):
        m, owner, stage, cdt = ctx.m, ctx.owner, ctx.stage, ctx.cdt
        eng: Engine = owner._engine
        pc = eng.ctx
        if pc.flat is None or m.params != pc.flat.data_ptr():
            raise MmerError("parameters were re-allocated between forward and backward")
        need_v, need_a, need_f = ctx.needs_input_grad[1], ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        if pc.bind_grads():
            pc.grads.zero_()
        keep = list(ctx.keep)
        B, T = m.B, m.T
        dvideo = daudio = dfused_out = None
        if stage == 1:
            df = d_main.detach().to(cdt).contiguous()
            m.dfused_in = df.data_ptr()
            keep.append(df)
        else:
            dl = d_main.detach().to(torch.float32).contiguous()
            m.dlogits = dl.data_ptr()
            keep.append(dl)
        dev = d_main.device
        if stage != 2:
            if need_v:
                dvideo = torch.empty((B, T, eng.cfg["video_dim"]), device=dev, dtype=cdt)
                m.dvideo = dvideo.data_ptr()
            if need_a:
                daudio = torch.empty((B, eng.cfg["audio_dim"]), device=dev, dtype=cdt)
                m.daudio = daudio.data_ptr()
        elif need_f:
            dfused_out = torch.empty((B, eng.cfg["fused"]), device=dev, dtype=cdt)
            m.dfused_out = dfused_out.data_ptr()
        Engine.backward(m)
        pc.attach_grads()
        ctx.keep = None
        vd, ad, fd = ctx.in_dtypes
        return (None,
                dvideo.to(vd) if dvideo is not None else None,
                daudio.to(ad) if daudio is not None else None,
                dfused_out.to(fd) if dfused_out is not None else None,
                None, None, None, None)
