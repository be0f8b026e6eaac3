"""N-rank check of the fused NVLS data-parallel step (run under torchrun on >= 2 GPUs of one NVSwitch node):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/nvls_check.py

(a) dp_mode="nvls" (barrier | multimem reduce-scatter + Adam shard + multimem all-gather in one kernel | barrier) and
(b) dp_mode="nccl" (NCCL all-reduce after backward + the ordinary Adam kernel) start from the same weights, see the same
per-rank batches and must end with the same weights after three steps; in mode (a) every rank must hold bit-identical
weights and a bf16 shadow equal to the rounded fp32 weights.  Adam runs with eps = 1 here so that the update is a smooth
function of the gradient (with the usual 1e-8 the first steps are lr * sign(g), which turns the fp32-atomic summation
noise of near-zero gradients into +-lr differences that say nothing about the exchange).  Then both modes are timed at
the bench shape (4096 samples per rank).
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmer_b200 as mm  # noqa: E402


def make(mode, dev, dropout):
    torch.manual_seed(0)
    model = mm.MultimodalEmotionModel(max_seq_len=17, fusion_num_layers=2, classifier_hidden_dim=512, fusion_dropout=dropout,
                                      classifier_dropout=dropout).to(dev).train()
    return model


def run(mode, rank, dev, nsteps, clip=None):
    model = make(mode, dev, 0.0)
    step = mm.FusedTrainStep(model, lr=1e-2, weight_decay=1e-4, eps=1.0, loss="focal", alpha=torch.tensor([1, 1, 1, 1, 1.2, 1.2]),
                             compute_dtype=torch.bfloat16, overlap_allreduce=False, dp_mode=mode, clip_grad_norm=clip)
    g = torch.Generator().manual_seed(7 + rank)
    for _ in range(nsteps):
        v = torch.randn(512, 16, 768, generator=g).to(dev).bfloat16()
        a = torch.randn(512, 1024, generator=g).to(dev).bfloat16()
        y = torch.randint(0, 6, (512,), generator=g).to(dev)
        loss, _ = step.step(v, a, None, y)
    torch.cuda.synchronize()
    ctx = model._engine.ctx
    return ctx.flat.clone(), ctx.shadow.clone(), float(loss), step.dp_mode


def timed(mode, overlap, rank, dev, steps=100, warmup=10):
    model = make(mode, dev, 0.1)
    step = mm.FusedTrainStep(model, lr=1e-4, weight_decay=1e-4, loss="focal", alpha=torch.tensor([1, 1, 1, 1, 1.2, 1.2]),
                             compute_dtype=torch.bfloat16, overlap_allreduce=overlap, dp_mode=mode)
    g = torch.Generator().manual_seed(11 + rank)
    v = torch.randn(4096, 16, 768, generator=g).to(dev).bfloat16()
    a = torch.randn(4096, 1024, generator=g).to(dev).bfloat16()
    y = torch.randint(0, 6, (4096,), generator=g).to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(warmup + steps):
        if i == warmup:
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
        step.step(v, a, None, y)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t), step.dp_mode


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pa, sa, la, mode_a = run("auto", rank, dev, 1)
    if rank == 0:
        print("auto mode resolved to:", mode_a, flush=True)
    ok = True
    if mode_a == "nvls":
        init = make("x", dev, 0.0)._engine
        init.ctx.ensure()
        p0 = init.ctx.flat.clone()
        # one step: both modes see the same weights, so only the summation order of the gradients differs
        pb, sb, lb, _ = run("nccl", rank, dev, 1)
        upd1, moved1 = float((pb - pa).abs().max()), float((pb - p0).abs().max())
        # three steps: bf16 rounding of slightly different weights feeds back, a looser bound applies
        pa3, sa3, la3, _ = run("nvls", rank, dev, 3)
        pb3, sb3, lb3, _ = run("nccl", rank, dev, 3)
        upd3, moved3 = float((pb3 - pa3).abs().max()), float((pb3 - p0).abs().max())
        ref = pa3.clone()
        dist.broadcast(ref, src=0)
        agree = bool(torch.equal(ref, pa3))
        shadow_ok = bool(torch.equal(sa3, pa3.bfloat16()))
        # gradient clipping (train2.py:576): the norm of the reduced gradient, exchanged through the symmetric slot array
        pc, _, _, mode_c = run("nvls", rank, dev, 1, clip=0.05)
        pd, _, _, _ = run("nccl", rank, dev, 1, clip=0.05)
        updc, movedc = float((pd - pc).abs().max()), float((pd - p0).abs().max())
        refc = pc.clone()
        dist.broadcast(refc, src=0)
        # (the clipped step moves the weights by ~3e-5: one fp32 ulp of a weight, 3.7e-9, is already 1e-4 of that)
        clip_ok = mode_c == "nvls" and updc < max(1e-4 * movedc, 1e-8) and movedc < 0.5 * moved1 and bool(torch.equal(refc, pc))
        print(f"rank {rank}: clip 0.05, 1 step: max |nvls - nccl| weights = {updc:.3e} of {movedc:.3e} moved "
              f"(unclipped step moved {moved1:.3e}); ok: {clip_ok}", flush=True)
        ok = upd1 < 1e-4 * moved1 and upd3 < 1e-2 * moved3 and agree and shadow_ok and moved1 > 1e-5 and clip_ok
        print(f"rank {rank}: 1 step: max |nvls - nccl| weights = {upd1:.3e} of {moved1:.3e} moved; 3 steps: {upd3:.3e} of "
              f"{moved3:.3e}; ranks bit-identical: {agree}; shadow == bf16(weights): {shadow_ok}; "
              f"loss {la3:.6f} / {lb3:.6f}; ok: {ok}", flush=True)
    if os.environ.get("NVLS_CHECK_NO_TIMING"):
        dist.barrier()
        dist.destroy_process_group()
        if not ok:
            raise SystemExit(1)
        return
    for mode, overlap in (("nccl", True), ("nvls", False), ("nccl", False), ("nvls", False), ("nccl", True), ("nvls", False)):
        if mode == "nvls" and mode_a != "nvls":
            continue
        ms, used = timed(mode, overlap, rank, dev)
        if rank == 0:
            print(f"timing: dp_mode={used} overlap={overlap}: {ms:.3f} ms/step, {4096 * dist.get_world_size() / ms * 1e3:.0f} samples/s",
                  flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
