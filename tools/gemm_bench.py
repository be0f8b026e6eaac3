"""Times the tcgen05 GEMM at every shape of the cfg2 training step (run under gpurun).

  python tools/gemm_bench.py [reps]

Each case rotates over 3 operand sets so that a launch never finds its inputs in L2 from the
previous one (every set is > 126 MB for the big shapes); times are CUDA-event averages.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmer_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda")
B, T = 4096, 16
M, Mv = B * (T + 1), B * T
bf = torch.bfloat16


def timeit(fn, reps):
    for i in range(3):
        fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20
    if "--generic-epilogue" in sys.argv:
        _lib.load().mmer_debug_set(_lib.DEBUG_GENERIC_EPI, 1)
    print(torch.cuda.get_device_name(0))
    NS = 3
    total_ms = 0.0
    total_flop = 0.0

    def rnd(*s):
        return torch.randn(*s, device=dev, dtype=bf)

    # (name, rows, K, N, count per step)
    lin = [("video_proj", Mv, 768, 512, 1), ("audio_proj", B, 1024, 512, 1), ("qkv", M, 512, 1536, 2),
           ("out_proj", M, 512, 512, 2), ("ffn1", M, 512, 2048, 2), ("ffn2", M, 2048, 512, 2), ("head", B, 512, 512, 2)]
    for name, m, k, n, cnt in lin:
        xs = [rnd(m, k) for _ in range(NS)]
        w = rnd(n, k)
        bias = torch.zeros(n, device=dev)
        ys = [torch.empty(m, n, device=dev, dtype=bf) for _ in range(NS)]
        dys = [rnd(m, n) for _ in range(NS)]
        dxs = [torch.empty(m, k, device=dev, dtype=bf) for _ in range(NS)]
        aux = [rnd(m, k) for _ in range(NS)]
        gw = torch.zeros(n, k, device=dev)
        gbits = torch.randint(0, 256, (m * k // 8,), device=dev, dtype=torch.uint8)
        mbits = torch.zeros(m * n // 8, device=dev, dtype=torch.uint8)
        flop = 2.0 * m * n * k
        cases = [
            ("fwd +bias", lambda i: ops.gemm(xs[i % NS], w, M=m, N=n, K=k, bias=bias, out=ys[i % NS])),
            ("fwd +bias+relu+drop", lambda i: ops.gemm(xs[i % NS], w, M=m, N=n, K=k, bias=bias, relu=True, drop_p=0.1,
                                                       seed=1, site=1, out=ys[i % NS])),
            ("dgrad", lambda i: ops.gemm(dys[i % NS], w, M=m, N=k, K=n, b_major=_lib.MAJOR_MN, out=dxs[i % NS])),
            ("dgrad +residual", lambda i: ops.gemm(dys[i % NS], w, M=m, N=k, K=n, b_major=_lib.MAJOR_MN,
                                                   residual=aux[i % NS], out=dxs[i % NS])),
            ("dgrad +gate", lambda i: ops.gemm(dys[i % NS], w, M=m, N=k, K=n, b_major=_lib.MAJOR_MN,
                                               gate=aux[i % NS], gate_scale=1.1, out=dxs[i % NS])),
            ("dgrad +gate bits", (lambda i: ops.gemm(dys[i % NS], w, M=m, N=k, K=n, b_major=_lib.MAJOR_MN,
                                                     gate_bits=gbits, gate_scale=1.1, out=dxs[i % NS])) if k % 64 == 0 else None),
            ("fwd +bias+relu+drop+mask", (lambda i: ops.gemm(xs[i % NS], w, M=m, N=n, K=k, bias=bias, relu=True, drop_p=0.1, seed=1,
                                                             site=1, out=ys[i % NS], relu_mask_out=mbits)) if n % 64 == 0 else None),
            ("wgrad", lambda i: ops.gemm(dys[i % NS], xs[i % NS], M=n, N=k, K=m, a_major=_lib.MAJOR_MN,
                                         b_major=_lib.MAJOR_MN, out=gw, accumulate=True)),
        ]
        if "--cublas" in sys.argv:
            # library reference on the same operands (torch.matmul -> cuBLASLt): not part of the product path
            wt = w.t().contiguous()
            cases += [
                ("cuBLAS fwd (x@W^T)", lambda i: torch.matmul(xs[i % NS], w.t(), out=ys[i % NS])),
                ("cuBLAS dgrad (dy@W)", lambda i: torch.matmul(dys[i % NS], w, out=dxs[i % NS])),
                ("cuBLAS wgrad (dy^T@x)", lambda i: torch.matmul(dys[i % NS].t(), xs[i % NS])),
            ]
        for cname, fn in cases:
            if fn is None:
                continue
            ms = timeit(fn, reps)
            print(f"{name:11s} {cname:22s} M={m:6d} K={k:5d} N={n:5d}  {ms * 1e3:8.1f} us  {flop / ms / 1e9:8.1f} TFLOP/s",
                  flush=True)
            if cname in ("fwd +bias", "dgrad", "wgrad"):
                total_ms += ms * cnt
                total_flop += flop * cnt
        del xs, ys, dys, dxs, aux
    print(f"sum over one step's GEMMs (plain fwd + dgrad + wgrad): {total_ms:.3f} ms, "
          f"{total_flop / total_ms / 1e9:.1f} TFLOP/s average")


if __name__ == "__main__":
    main()
