"""Launch one family of kernels at the cfg2 shapes a few times (target for `ncu --set full -k regex:...`).

  python tools/run_kernel.py mha|add_ln|gemm|colsum|embed|adam [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmer_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda")
B, T, H, D, F = 4096, 16, 8, 64, 512
S = T + 1
M = B * S
bf = torch.bfloat16


def main():
    what = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    g = torch.Generator(device="cuda").manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=g).to(bf)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    if what == "mha":
        qkv, do = rnd(M, 3 * F), rnd(M, F)
        fn = lambda: (ops.mha_fwd(qkv, None, B, T, H, D, drop_p=0.1, seed=1, site=1),
                      ops.mha_bwd(qkv, None, do, B, T, H, D, drop_p=0.1, seed=1, site=1))
    elif what == "add_ln":
        x, a, dy = rnd(M, F), rnd(M, F), rnd(M, F)
        gam, bet = torch.ones(F, device=dev), torch.zeros(F, device=dev)
        dg, db, dbias = torch.zeros(F, device=dev), torch.zeros(F, device=dev), torch.zeros(F, device=dev)

        def fn():
            y, st = ops.add_ln_fwd(x, a, gam, bet, drop_a_p=0.1, site_a=1, seed=1)
            ops.add_ln_bwd(dy, x, a, st, gam, bet, dg, db, dbias, drop_a_p=0.1, site_a=1, seed=1)
    elif what == "colsum":
        x = rnd(M, 2048)
        out = torch.zeros(2048, device=dev)
        fn = lambda: ops.colsum(x, out)
    elif what == "gemm":
        x, w, bias = rnd(M, 512), rnd(2048, 512), torch.zeros(2048, device=dev)
        out = torch.empty(M, 2048, device=dev, dtype=bf)
        fn = lambda: ops.gemm(x, w, M=M, N=2048, K=512, bias=bias, relu=True, drop_p=0.1, seed=1, site=1, out=out)
    elif what == "gemm_relu":   # the kernel bench.py reports as `roofline` (FFN1 forward as the step launches it: bias + ReLU + dropout + 1-bit mask)
        x, w, bias = rnd(M, 512), rnd(2048, 512), torch.zeros(2048, device=dev)
        out = torch.empty(M, 2048, device=dev, dtype=bf)
        mask = torch.empty(M * 2048 // 8, device=dev, dtype=torch.uint8)
        fn = lambda: ops.gemm(x, w, M=M, N=2048, K=512, bias=bias, relu=True, drop_p=0.1, seed=1, site=1, out=out,
                              relu_mask_out=mask)
    elif what == "gemm_wgrad":   # linear1's weight gradient as the step launches it: dW[2048,512] += dY^T X, + row sums
        dy, x = rnd(M, 2048), rnd(M, 512)
        gw, gb = torch.zeros(2048, 512, device=dev), torch.zeros(2048, device=dev)
        fn = lambda: ops.linear_wgrad(dy, x, gw, dbias=gb)
    elif what == "gemm_gate":
        dy, w, h = rnd(M, 512), rnd(512, 2048), rnd(M, 2048)
        out = torch.empty(M, 2048, device=dev, dtype=bf)
        fn = lambda: ops.gemm(dy, w, M=M, N=2048, K=512, b_major=_lib.MAJOR_MN, gate=h, gate_scale=1.1, out=out)
    elif what == "gemm_res":
        dy, w, r = rnd(M, 2048), rnd(2048, 512), rnd(M, 512)
        out = torch.empty(M, 512, device=dev, dtype=bf)
        fn = lambda: ops.gemm(dy, w, M=M, N=512, K=2048, b_major=_lib.MAJOR_MN, residual=r, out=out)
    elif what == "gemm_plain":
        x, w, bias = rnd(M, 512), rnd(1536, 512), torch.zeros(1536, device=dev)
        out = torch.empty(M, 1536, device=dev, dtype=bf)
        fn = lambda: ops.gemm(x, w, M=M, N=1536, K=512, bias=bias, out=out)
    else:
        raise SystemExit("unknown kernel family")
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    print(f"{what}: {ev[0].elapsed_time(ev[1]) / reps * 1e3:.1f} us per iteration")


if __name__ == "__main__":
    main()
