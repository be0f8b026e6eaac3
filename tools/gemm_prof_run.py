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


def run_wgrad(name, Ntok, Nout, Kin, with_bias=False):
    dy = torch.randn(Ntok, Nout, device=dev).to(bf); x = torch.randn(Ntok, Kin, device=dev).to(bf)
    gw = torch.zeros(Nout, Kin, device=dev); db = torch.zeros(Nout, device=dev)
    torch.cuda.synchronize()
    print("==", name, flush=True)
    ops.linear_wgrad(dy, x, gw, dbias=db if with_bias else None)
    torch.cuda.synchronize()


# round 2: no weight-gradient GEMM of the step carries the row-sum MMA any more (linear1's bias gradient comes out of
# linear2's dgrad epilogue), so the captures below are the kernels exactly as the step launches them
run_wgrad("ffn1 wgrad (dW[2048,512], K=69632)", M, 2048, 512)
run_wgrad("ffn2 wgrad (dW[512,2048])", M, 512, 2048)
