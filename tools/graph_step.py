"""How much of the step is launch gaps?  Times the cfg2 training step launched kernel by kernel from Python and
replayed from a CUDA graph of the same launches (measurement only: a captured graph replays one dropout seed)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmer_b200 as mm  # noqa: E402

dev = torch.device("cuda")
B, T = 4096, 16
torch.manual_seed(0)
model = mm.MultimodalEmotionModel(max_seq_len=T + 1, fusion_num_layers=2, classifier_hidden_dim=512, fusion_dropout=0.1,
                                  classifier_dropout=0.1).to(dev).train()
step = mm.FusedTrainStep(model, lr=1e-4, weight_decay=1e-4, loss="focal", alpha=torch.tensor([1, 1, 1, 1, 1.2, 1.2]))
v = torch.randn(B, T, 768, device=dev).bfloat16()
a = torch.randn(B, 1024, device=dev).bfloat16()
y = torch.randint(0, 6, (B,), device=dev)
for _ in range(5):
    step.step(v, a, None, y)
torch.cuda.synchronize()


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


eager = timed(lambda: step.step(v, a, None, y), 30)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step.step(v, a, None, y)
g.replay()
torch.cuda.synchronize()
graph = timed(g.replay, 30)
print(f"step launched from Python: {eager:.3f} ms; same launches replayed from a CUDA graph: {graph:.3f} ms")
