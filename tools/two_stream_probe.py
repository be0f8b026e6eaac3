"""Probe: is one GPU faster running TWO half-batch training steps side by side on disjoint SM sets than one full-batch step
on all SMs?  (GEMMs are power-capped and row kernels HBM-bound, so the two kinds of phases could overlap.)

  python tools/two_stream_probe.py            (uses MMER_DEBUG=9=<reserve> semantics through mmer_debug_set)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmer_b200 as mm  # noqa: E402
from mmer_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
ALPHA = torch.tensor([1, 1, 1, 1, 1.2, 1.2])


def make(B):
    torch.manual_seed(0)
    model = mm.MultimodalEmotionModel(max_seq_len=17, fusion_num_layers=2, classifier_hidden_dim=512).to(dev).train()
    step = mm.FusedTrainStep(model, lr=1e-4, weight_decay=1e-4, loss="focal", alpha=ALPHA, compute_dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(1)
    v = torch.randn(B, 16, 768, generator=g).to(dev).bfloat16()
    a = torch.randn(B, 1024, generator=g).to(dev).bfloat16()
    y = torch.randint(0, 6, (B,), generator=g).to(dev)
    return step, v, a, y


def run(nstreams, B, reserve, steps=60, warmup=8, stagger=False):
    _lib.load().mmer_debug_set(_lib.DEBUG_RESERVE_SMS, reserve)
    jobs = [make(B) for _ in range(nstreams)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(warmup + steps):
        if i == warmup:
            torch.cuda.synchronize()
            e0.record()
            for s in streams:
                s.wait_event(e0)
        for (step, v, a, y), s in zip(jobs, streams):
            with torch.cuda.stream(s):
                step.step(v, a, None, y)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"streams={nstreams} batch/stream={B} SMs/kernel={148 - reserve}: {ms:.3f} ms per round, "
          f"{nstreams * B / ms * 1e3:.0f} samples/s", flush=True)
    _lib.load().mmer_debug_set(_lib.DEBUG_RESERVE_SMS, 0)


run(1, 4096, 0)
run(2, 2048, 74)
run(2, 2048, 48)
run(2, 2048, 0)
run(1, 2048, 0)
run(1, 2048, 74)
run(2, 4096, 74)
run(1, 4096, 0)
