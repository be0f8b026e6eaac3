"""Timings of the other BASELINE.json configs (parity-test cases, not the bench line): cfg4 (T=256, attention weights
returned) and cfg5 (inference only: batch 1 latency, batch 8192 throughput), bf16, CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmer_b200 as mm  # noqa: E402

dev = torch.device("cuda")


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def model_for(T, p=0.0):
    torch.manual_seed(0)
    m = mm.MultimodalEmotionModel(max_seq_len=T + 1, fusion_num_layers=2, classifier_hidden_dim=512, fusion_dropout=p,
                                  classifier_dropout=p).to(dev)
    m.compute_dtype = torch.bfloat16
    return m


def main():
    print(torch.cuda.get_device_name(0))
    # cfg5: inference
    for B, T in ((1, 5), (1, 16), (8192, 16)):
        m = model_for(T).eval()
        v = torch.randn(B, T, 768, device=dev).bfloat16()
        a = torch.randn(B, 1024, device=dev).bfloat16()
        with torch.no_grad():
            ms = timed(lambda: m(v, a), 50 if B == 1 else 20)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                m(v, a)
            msg = timed(g.replay, 50 if B == 1 else 20)
        print(f"cfg5 inference forward B={B:5d} T={T:3d}: {ms * 1e3:9.1f} us per call ({B / ms * 1e3:12.0f} samples/s); "
              f"CUDA-graph replay {msg * 1e3:9.1f} us ({B / msg * 1e3:12.0f} samples/s)")
    # cfg4: long sequence, attention weights
    B, T = 512, 256
    m = model_for(T, 0.1).train()
    step = mm.FusedTrainStep(m, lr=1e-4, weight_decay=1e-4, loss="wce", alpha=torch.tensor([1, 1, 1, 1, 1.2, 1.2]),
                             clip_grad_norm=1.0)
    v = torch.randn(B, T, 768, device=dev).bfloat16()
    a = torch.randn(B, 1024, device=dev).bfloat16()
    y = torch.randint(0, 6, (B,), device=dev)
    ms = timed(lambda: step.step(v, a, None, y), 5, warm=2)
    print(f"cfg4 training step B={B} T={T} (weighted CE, clip 1.0): {ms:8.2f} ms ({B / ms * 1e3:10.0f} samples/s)")
    m.eval()
    with torch.no_grad():
        ms = timed(lambda: m(v, a, None, return_attn=True), 5, warm=2)
    print(f"cfg4 eval forward with attention weights B={B} T={T}: {ms:8.2f} ms")


if __name__ == "__main__":
    main()
