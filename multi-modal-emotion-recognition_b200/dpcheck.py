"""Self-check of the data-parallel training step, for multi-GPU jobs (bench.py at N > 1, tests, user sanity checks).

The reference has no distributed code; what a data-parallel run of its loop body (train2.py:570-579) must
compute is fixed by the single-process semantics: the gradient of the mean loss over the GLOBAL batch, identical weights
on every replica.  ``dp_selfcheck`` runs ONE optimisation step through the default exchange (``dp_mode="auto"``: the
fused multicast reduce-scatter + Adam + all-gather kernel on an NVSwitch node) and ONE through the plain NCCL all-reduce
path, both from the same initial weights and the same per-rank batches, and reports

* ``mode``                        which exchange "auto" resolved to on this fabric,
* ``ranks_bit_identical``         every rank's fp32 weights after the step have the same bits (checked by all-gathering
                                  a 64-bit checksum AND by broadcasting rank 0's buffer),
* ``max_abs_vs_nccl``             largest weight difference between the two exchanges, next to ``max_abs_moved`` (how far
                                  the step moved the weights): only the summation order of the gradients differs,
* ``shadow_equals_bf16_weights``  (bf16 compute) the bf16 shadow the next forward reads equals bf16(fp32 weights),
* ``checksums``                   the per-rank checksums themselves.

Adam runs with eps = 1 here so that the update is a smooth function of the gradient: with 1e-8 the first step is
lr * sign(g), which turns fp32 summation-order noise of near-zero gradients into +-lr flips that say nothing about the
exchange.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

__all__ = ["dp_selfcheck", "weights_checksum"]


def weights_checksum(flat: torch.Tensor) -> int:
    """Order-sensitive 62-bit checksum of the BITS of a float32 buffer (computed on the device, one host read)."""
    bits = flat.detach().contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    idx = torch.arange(bits.numel(), device=bits.device, dtype=torch.int64)
    mixed = (bits * 2654435761 + idx * 40503) & 0x3FFFFFFFFFFFFFFF
    return int(mixed.sum().item() & 0x3FFFFFFFFFFFFFFF)


def _one_step(mode: str, dev, group, samples: int, frames: int, dtype: torch.dtype, clip: Optional[float]):
    from . import FusedTrainStep, MultimodalEmotionModel
    rank = dist.get_rank(group)
    torch.manual_seed(0)
    model = MultimodalEmotionModel(max_seq_len=frames + 1, fusion_num_layers=2, classifier_hidden_dim=512,
                                   fusion_dropout=0.0, classifier_dropout=0.0).to(dev).train()
    step = FusedTrainStep(model, lr=1e-2, weight_decay=1e-4, eps=1.0, loss="focal",
                          alpha=torch.tensor([1, 1, 1, 1, 1.2, 1.2]), compute_dtype=dtype, overlap_allreduce=mode != "nccl_seq",
                          dp_mode="nccl" if mode.startswith("nccl") else mode, clip_grad_norm=clip, process_group=group)
    ctx = model._engine.ctx
    before = ctx.flat.clone()
    g = torch.Generator().manual_seed(7 + rank)
    v = torch.randn(samples, frames, 768, generator=g).to(dev).to(dtype)
    a = torch.randn(samples, 1024, generator=g).to(dev).to(dtype)
    y = torch.randint(0, 6, (samples,), generator=g).to(dev)
    loss, _ = step.step(v, a, None, y)
    torch.cuda.synchronize(dev)
    shadow = ctx.shadow.clone() if ctx.shadow is not None else None
    return before, ctx.flat.clone(), shadow, float(loss), step.dp_mode


def dp_selfcheck(dev, group=None, samples: int = 512, frames: int = 16, dtype: torch.dtype = torch.bfloat16,
                 clip: Optional[float] = None) -> dict:
    """Collective: every rank of ``group`` must call it.  Returns the report described in the module docstring (the
    same dict on every rank except ``checksum_this_rank``)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
        return {"mode": "single", "ranks_bit_identical": True, "max_abs_vs_nccl": 0.0, "world": 1}
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    p0, pa, sa, loss_a, mode = _one_step("auto", dev, group, samples, frames, dtype, clip)
    _, pb, _, loss_b, _ = _one_step("nccl", dev, group, samples, frames, dtype, clip)
    mine = weights_checksum(pa)
    sums = torch.zeros(world, dtype=torch.int64, device=dev)
    sums[rank] = mine
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    ref = pa.clone()
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(ref, src=src, group=group)
    same = torch.tensor([int(torch.equal(ref, pa))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN, group=group)
    stats = torch.tensor([float((pa - pb).abs().max()), float((pa - p0).abs().max()),
                          0.0 if (sa is None or dtype != torch.bfloat16 or torch.equal(sa, pa.bfloat16())) else 1.0, abs(loss_a - loss_b)],
                         device=dev, dtype=torch.float64)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX, group=group)
    checks = [int(x) for x in sums.tolist()]
    return {"mode": mode, "world": world,
            "ranks_bit_identical": bool(int(same.item()) == 1 and len(set(checks)) == 1),
            "max_abs_vs_nccl": float(stats[0]), "max_abs_moved": float(stats[1]),
            "shadow_equals_bf16_weights": bool(float(stats[2]) == 0.0),
            "max_loss_diff_vs_nccl": float(stats[3]), "checksums": checks, "checksum_this_rank": mine,
            "samples_per_rank": samples, "frames": frames, "dtype": str(dtype).replace("torch.", ""), "clip": clip}
