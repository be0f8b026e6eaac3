"""ncu target: a few cfg4 training steps (train2 model, B=512, T=256, bf16, weighted CE + clip + Adam) and one eval
forward with attention weights.  `python tools/cfg4_step.py [steps]`"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmer_b200 as mm  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda")
B, T = 512, 256
torch.manual_seed(0)
m = mm.MultimodalEmotionModel(max_seq_len=T + 1, fusion_num_layers=2, classifier_hidden_dim=512, fusion_dropout=0.1,
                              classifier_dropout=0.1).to(dev).train()
m.compute_dtype = torch.bfloat16
step = mm.FusedTrainStep(m, lr=1e-4, weight_decay=1e-4, loss="wce", alpha=torch.tensor([1, 1, 1, 1, 1.2, 1.2]), clip_grad_norm=1.0)
v = torch.randn(B, T, 768, device=dev).bfloat16()
a = torch.randn(B, 1024, device=dev).bfloat16()
y = torch.randint(0, 6, (B,), device=dev)
for _ in range(steps):
    loss, _ = step.step(v, a, None, y)
torch.cuda.synchronize()
m.eval()
with torch.no_grad():
    m(v, a, None, return_attn=True)
torch.cuda.synchronize()
print("ok", float(loss))
