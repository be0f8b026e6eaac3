"""ncu target: a few launches of the fused Linear + dropout + residual + LayerNorm kernel at the step's shapes.
  python tools/prof_gemm_ln.py [K]      (K = 512: out_proj -> norm1; 2048: linear2 -> norm2)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmer_b200 import ops  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
M, F = 4096 * 17, 512
dev, bf = torch.device("cuda"), torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
sets = [(torch.randn(M, K, device=dev, generator=g).to(bf), torch.randn(M, F, device=dev, generator=g).to(bf)) for _ in range(3)]
w = (torch.randn(F, K, device=dev, generator=g) * K ** -0.5).to(bf)
gam, bet, bias = torch.ones(F, device=dev), torch.zeros(F, device=dev), torch.zeros(F, device=dev)
for i in range(6):
    a, r = sets[i % 3]
    z, y, st = ops.gemm_ln_fwd(a, w, bias, r, gam, bet, drop_p=0.1, site=1, seed=1)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
