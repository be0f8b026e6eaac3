import sys, torch
sys.path.insert(0, "/root/repo")
from mmer_b200 import _lib, ops
dev = "cuda"; bf = torch.bfloat16
M = 69632
def run(name, K, N, **kw):
    x = torch.randn(M, K, device=dev).to(bf); w = torch.randn(N, K, device=dev).to(bf)
    bias = torch.zeros(N, device=dev); out = torch.empty(M, N, device=dev, dtype=bf)
    torch.cuda.synchronize()
    print("==", name, flush=True)
    ops.gemm(x, w, M=M, N=N, K=K, bias=bias, out=out, **kw)
    torch.cuda.synchronize()
run("ffn1 fwd+bias K=512 N=2048", 512, 2048)
run("ffn1 fwd+bias+relu+drop", 512, 2048, relu=True, drop_p=0.1, seed=1, site=1)
run("ffn2 fwd+bias K=2048 N=512", 2048, 512)
run("out_proj K=512 N=512", 512, 512)
