"""2-rank check of the overlapped gradient all-reduce (run under torchrun on >= 2 GPUs):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py

Three optimisation steps from the same initial weights with (a) bucketed all-reduce overlapped with backward and
(b) one all-reduce after backward must give the same summed gradient buffer (up to the fp32-atomic summation order of the
split-K weight gradients), and the ranks must agree bit for bit.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmer_b200 as mm  # noqa: E402


def run(overlap, rank, dev):
    torch.manual_seed(0)
    model = mm.MultimodalEmotionModel(max_seq_len=17, fusion_num_layers=2, classifier_hidden_dim=512, fusion_dropout=0.0,
                                      classifier_dropout=0.0).to(dev).train()
    # lr = 0: the step leaves the weights alone, so the all-reduced gradient buffer of the LAST step can be compared
    step = mm.FusedTrainStep(model, lr=0.0, weight_decay=0.0, loss="focal", alpha=torch.tensor([1, 1, 1, 1, 1.2, 1.2]),
                             compute_dtype=torch.bfloat16, overlap_allreduce=overlap, dp_mode="nccl")
    g = torch.Generator().manual_seed(7 + rank)
    for _ in range(3):
        v = torch.randn(512, 16, 768, generator=g).to(dev).bfloat16()
        a = torch.randn(512, 1024, generator=g).to(dev).bfloat16()
        y = torch.randint(0, 6, (512,), generator=g).to(dev)
        loss, _ = step.step(v, a, None, y)
    torch.cuda.synchronize()
    return model._engine.ctx.grads.clone(), float(loss)


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ga, la = run(True, rank, dev)
    gb, lb = run(False, rank, dev)
    # split-K weight gradients use fp32 atomics, so two runs differ in the last bits
    diff = float((ga - gb).abs().max() / gb.abs().max())
    same = diff < 1e-5
    other = ga.clone()
    dist.broadcast(other, src=0)
    agree = bool(torch.equal(other, ga))          # the all-reduced gradients are bit-identical on every rank
    print(f"rank {rank}: max |overlapped - sequential| / max|g| = {diff:.3e} (ok: {same}); ranks agree bitwise: {agree}; "
          f"loss {la:.6f} / {lb:.6f}; |g|max {float(gb.abs().max()):.3e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not (same and agree):
        raise SystemExit(1)


if __name__ == "__main__":
    main()
