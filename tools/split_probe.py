import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmer_b200 import _lib, ops
dev = torch.device("cuda"); bf = torch.bfloat16
M = 4096 * 17
lib = _lib.load()
def bench(n_lin, k_lin, splits, reps=30):
    lib.mmer_debug_set(10, splits)
    NS = 3
    dys = [torch.randn(M, n_lin, device=dev, dtype=bf) for _ in range(NS)]
    xs = [torch.randn(M, k_lin, device=dev, dtype=bf) for _ in range(NS)]
    gw = torch.zeros(n_lin, k_lin, device=dev)
    fn = lambda i: ops.linear_wgrad(dys[i % NS], xs[i % NS], gw)
    for i in range(3): fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for name, n_lin, k_lin in (("ffn1", 2048, 512), ("ffn2", 512, 2048), ("qkv", 1536, 512), ("out", 512, 512), ("video", 512, 768)):
    for rep in range(2):
        row = []
        for s in (0, 4, 6, 9, 13, 18, 23):
            row.append(f"s={s}:{bench(n_lin, k_lin, s):6.1f}")
        print(name, " ".join(row), flush=True)
