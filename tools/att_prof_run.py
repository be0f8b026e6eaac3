import sys, torch
sys.path.insert(0, "/root/repo")
from mmer_b200 import ops
B, T, H, D, F = 4096, 16, 8, 64, 512
M = B * (T + 1)
qkv = torch.randn(M, 3 * F, device="cuda").bfloat16(); do = torch.randn(M, F, device="cuda").bfloat16()
db = torch.zeros(3 * F, device="cuda")
for p in (0.0, 0.1):
    print("== p", p, flush=True)
    ops.mha_bwd(qkv, None, do, B, T, H, D, drop_p=p, seed=1, site=1, dbias=db)
    torch.cuda.synchronize()
