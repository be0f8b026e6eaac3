"""Per-kernel timing at the cfg2 shapes (B=4096, T=16, F=512): CUDA events on the launching stream, rotating
buffers larger than L2, algorithmic bytes / time against the measured HBM peak.

  python tools/kernel_bench.py [mha] [add_ln] [colsum] [embed] [pool] [adam] [head]     (default: all)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmer_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda")
B, T, H, D, F, FF = 4096, 16, 8, 64, 512, 2048
S = T + 1
M = B * S
bf = torch.bfloat16
NBUF = 3   # rotating sets: every kernel below touches >= 140 MB per launch, 3 sets > 126 MB L2 by a wide margin

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0

g = torch.Generator(device="cuda").manual_seed(0)


def rnd(*s, dt=bf):
    return torch.randn(*s, device=dev, generator=g).to(dt)


def timeit(name, fns, bytes_per_launch, reps=20):
    """fns: list of NBUF closures, each launching the kernel once on its own buffer set"""
    for f in fns:
        f()
    torch.cuda.synchronize()
    # replay from a CUDA graph so that short kernels are not timed by the Python launch overhead
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(reps):
            fns[i % len(fns)]()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    gbs = bytes_per_launch / us * 1e-3
    print(f"{name:34s} {us:8.1f} us   {bytes_per_launch / 1e6:8.1f} MB   {gbs:7.0f} GB/s   {gbs / PEAK:5.2f} of measured HBM peak")
    return us


def bench_mha():
    lib = _lib.load()
    sets = [(rnd(M, 3 * F), rnd(M, F)) for _ in range(NBUF)]
    fb = M * 3 * F * 2 + M * F * 2
    bb = 2 * M * 3 * F * 2 + M * F * 2
    for variant in ("tma-tile", "bulk-row"):
        lib.mmer_debug_set(_lib.DEBUG_ATT_ROWS, 0 if variant == "tma-tile" else 1)
        tag = variant
        for p in (0.0, 0.1):
            timeit(f"mha_fwd[{tag}] p={p}", [lambda q=q: ops.mha_fwd(q, None, B, T, H, D, drop_p=p, seed=1, site=1)
                                             for q, _ in sets], fb)
            timeit(f"mha_bwd[{tag}] p={p}", [lambda q=q, d=d: ops.mha_bwd(q, None, d, B, T, H, D, drop_p=p, seed=1, site=1)
                                             for q, d in sets], bb)
            if p > 0:   # as the training step launches it: with the in_proj bias gradient (column sums of what is stored)
                dbias = torch.zeros(3 * F, device=dev)
                timeit(f"mha_bwd[{tag}] p={p} +dbias", [lambda q=q, d=d: ops.mha_bwd(q, None, d, B, T, H, D, drop_p=p, seed=1,
                                                                                    site=1, dbias=dbias) for q, d in sets], bb)
    lib.mmer_debug_set(_lib.DEBUG_ATT_ROWS, 0)


def bench_add_ln():
    gam, bet = torch.ones(F, device=dev), torch.zeros(F, device=dev)
    dg, db, dbias = (torch.zeros(F, device=dev) for _ in range(3))
    sets = [(rnd(M, F), rnd(M, F), rnd(M, F)) for _ in range(NBUF)]
    stats = ops.add_ln_fwd(sets[0][0], sets[0][1], gam, bet)[1]
    e = 2
    for p in (0.0, 0.1):
        timeit(f"add_ln_fwd p={p}", [lambda x=x, a=a: ops.add_ln_fwd(x, a, gam, bet, drop_a_p=p, site_a=1, seed=1)
                                     for x, a, _ in sets], 3 * M * F * e)
        nb = (5 if p > 0 else 4) * M * F * e
        timeit(f"add_ln_bwd p={p}",
               [lambda x=x, a=a, dy=dy: ops.add_ln_bwd(dy, x, a, stats, gam, bet, dg, db, dbias, drop_a_p=p, site_a=1, seed=1)
                for x, a, dy in sets], nb)


def bench_colsum():
    for N in (512, 1536, 2048):
        sets = [rnd(M, N) for _ in range(NBUF)]
        out = torch.zeros(N, device=dev)
        timeit(f"colsum N={N}", [lambda x=x: ops.colsum(x, out) for x in sets], M * N * 2)


def bench_adam():
    n = 7_765_510
    p, gr, m, v = (torch.randn(n, device=dev) for _ in range(4))
    v.abs_()
    sh = torch.empty(n, device=dev, dtype=bf)
    timeit("adam (7.77M params, 30 B each)", [lambda: ops.adam_step(p, gr, m, v, sh, 3, 1e-4, weight_decay=1e-4)], n * 30, reps=50)


def bench_embed():
    gv, bv, ga, ba = (torch.ones(F, device=dev) for _ in range(4))
    pos = torch.zeros(S, F, device=dev)
    sets = [(rnd(B * T, F), rnd(B, F), rnd(M, F)) for _ in range(NBUF)]
    dg = [torch.zeros(F, device=dev) for _ in range(4)]
    dpos = torch.zeros(S, F, device=dev)
    stats = ops.embed_fwd(sets[0][0], sets[0][1], gv, bv, ga, ba, pos, B, T)[1]
    timeit("embed_fwd p=0.1", [lambda pv=pv, pa=pa: ops.embed_fwd(pv, pa, gv, bv, ga, ba, pos, B, T, drop_p=0.1, seed=1, site=1)
                               for pv, pa, _ in sets], (B * T * F + B * F + M * F) * 2)
    timeit("embed_bwd p=0.1", [lambda pv=pv, pa=pa, dx=dx: ops.embed_bwd(dx, pv, pa, stats, gv, ga, B, T, dg[0], dg[1], dg[2],
                                                                         dg[3], dpos, drop_p=0.1, seed=1, site=1)
                               for pv, pa, dx in sets], (2 * (B * T * F + B * F) + M * F) * 2)


def bench_head():
    Hd, Cn = 512, 6
    h = rnd(B, Hd)
    W, dW, db = torch.randn(Cn, Hd, device=dev), torch.zeros(Cn, Hd, device=dev), torch.zeros(Cn, device=dev)
    dl = torch.randn(B, Cn, device=dev)
    timeit("head_out_bwd (4096x512 -> 6)", [lambda: ops.head_out_bwd(dl, h, W, dW, db)], 2 * B * Hd * 2, reps=50)


def bench_collate():
    """Batch assembly (mmer_collate): 4096 samples of up to 16 frames gathered from a resident ragged feature set,
    z-scored, padded and cast to bf16, against the reference's host collate_fn + .to(device) for the same batch."""
    import time
    import mmer_b200 as mm
    from torch.nn.utils.rnn import pad_sequence
    gen = torch.Generator().manual_seed(0)
    n = 3 * B
    lens = torch.randint(8, T + 1, (n,), generator=gen).tolist()
    videos = [torch.randn(t, 768, generator=gen) for t in lens]
    audios = [torch.randn(1024, generator=gen) for _ in range(n)]
    labels = torch.randint(0, 6, (n,), generator=gen).tolist()
    ds = mm.DeviceFeatureSet(videos, audios, labels)
    batches = [list(range(k * B, (k + 1) * B)) for k in range(3)]
    frames = [sum(lens[i] for i in b) for b in batches]
    tmax = [max(lens[i] for i in b) for b in batches]
    byts = sum(f * 768 * 4 + B * 1024 * 4 + B * tm * 768 * 2 + B * 1024 * 2 + B * tm for f, tm in zip(frames, tmax)) / 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for b in batches:
        ds.collate(b, torch.bfloat16)
    torch.cuda.synchronize()
    reps = 12
    e0.record()
    for i in range(reps):
        ds.collate(batches[i % 3], torch.bfloat16)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{'collate (device, incl. host index upload)':34s} {us:8.1f} us   {byts / 1e6:8.1f} MB   {byts / us * 1e-3:7.0f} GB/s   "
          f"{byts / us * 1e-3 / PEAK:5.2f} of measured HBM peak   {B / us:6.1f} M samples/s")
    vm, vs, am, as_ = (t.cpu() for t in (ds.video_mean, ds.video_std, ds.audio_mean, ds.audio_std))
    nv = [(v - vm) / vs for v in videos[:B]]
    na = [(a - am) / as_ for a in audios[:B]]
    t0 = time.perf_counter()
    vp = pad_sequence(nv, batch_first=True, padding_value=0.0)
    ap = torch.stack(na)
    mk = pad_sequence([torch.zeros(len(v), dtype=torch.bool) for v in nv], batch_first=True, padding_value=True)
    vp, ap, mk = vp.to(dev), ap.to(dev), mk.to(dev)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    print(f"{'reference collate_fn + .to(device)':34s} {ms * 1e3:8.1f} us   (host pad_sequence/stack of pre-normalised tensors, "
          f"pageable H2D)   {B / ms / 1e3:6.3f} M samples/s")


def bench_gemm_ln():
    """Fused Linear + dropout + residual + LayerNorm (gemm_ln.cu) against the two kernels it replaces, at the two shapes
    of the step (out_proj: K = 512, linear2: K = 2048), dropout 0.1 as in training."""
    gam, bet, bias = torch.ones(F, device=dev), torch.zeros(F, device=dev), torch.zeros(F, device=dev)
    for K in (F, FF):
        sets = [(rnd(M, K), rnd(M, F)) for _ in range(NBUF)]
        w = rnd(F, K) * K ** -0.5
        flop = 2.0 * M * F * K
        byts = M * K * 2 + 3 * M * F * 2
        us = timeit(f"gemm_ln_fwd K={K} (p=0.1)", [lambda a=a, r=r: ops.gemm_ln_fwd(a, w, bias, r, gam, bet, drop_p=0.1, site=1, seed=1)
                                                   for a, r in sets], byts)
        print(f"{'':34s} {flop / us * 1e-6:8.0f} TFLOP/s")
        outs = [torch.empty(M, F, device=dev, dtype=bf) for _ in range(NBUF)]
        us_g = timeit(f"  unfused: gemm K={K} bias", [lambda a=a, o=o: ops.gemm(a, w, M=M, N=F, K=K, bias=bias, out=o)
                                                      for (a, _), o in zip(sets, outs)], M * K * 2 + M * F * 2)
        us_l = timeit("  unfused: add_ln_fwd (p=0.1)", [lambda r=r, o=o: ops.add_ln_fwd(r, o, gam, bet, drop_a_p=0.1, site_a=1, seed=1)
                                                        for (_, r), o in zip(sets, outs)], 3 * M * F * 2)
        print(f"{'':34s} fused {us:.1f} us vs unfused {us_g + us_l:.1f} us")
        del sets, outs
    # backward that reads the stored z instead of x and the sub-layer output
    dg, db, dbias = (torch.zeros(F, device=dev) for _ in range(3))
    sets = [(rnd(M, F), rnd(M, F), rnd(M, F)) for _ in range(NBUF)]
    stats = ops.add_ln_fwd(sets[0][0], sets[0][1], gam, bet)[1]
    timeit("add_ln_bwd (x, a) (p=0.1)", [lambda x=x, a=a, dy=dy: ops.add_ln_bwd(dy, x, a, stats, gam, bet, dg, db, dbias, drop_a_p=0.1,
                                                                               site_a=1, seed=1) for x, a, dy in sets], 5 * M * F * 2)
    timeit("add_ln_bwd_z (p=0.1)", [lambda a=a, dy=dy: ops.add_ln_bwd_z(dy, a, stats, gam, dg, db, dbias, drop_a_p=0.1, site_a=1, seed=1)
                                    for _, a, dy in sets], 4 * M * F * 2)


def main():
    which = sys.argv[1:] or ["mha", "add_ln", "embed", "head", "adam", "collate", "gemm_ln"]
    print(torch.cuda.get_device_name(0), f"HBM peak {PEAK:.0f} GB/s (measured)")
    for w in which:
        {"mha": bench_mha, "add_ln": bench_add_ln, "colsum": bench_colsum, "adam": bench_adam, "embed": bench_embed,
         "head": bench_head, "collate": bench_collate, "gemm_ln": bench_gemm_ln}[w]()


if __name__ == "__main__":
    main()
