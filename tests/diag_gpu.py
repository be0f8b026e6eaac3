"""One-shot GPU diagnostics (run under gpurun): GEMM descriptor variants, op parity, model parity, timing."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mmer_b200 as mm  # noqa: E402
from mmer_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


def gemm_case(M, N, K, amaj, bmaj, dtype, acc=False):
    A = torch.randn(M, K, device=dev).to(dtype)
    B = torch.randn(N, K, device=dev).to(dtype)
    ref = A.float() @ B.float().t()
    As = A if amaj == 0 else A.t().contiguous()
    Bs = B if bmaj == 0 else B.t().contiguous()
    if acc:
        out = torch.ones(M, N, device=dev, dtype=torch.float32)
        ops.gemm(As, Bs, M=M, N=N, K=K, a_major=amaj, b_major=bmaj, out=out, accumulate=True)
        ref = ref + 1
    else:
        out = ops.gemm(As, Bs, M=M, N=N, K=K, a_major=amaj, b_major=bmaj)
    torch.cuda.synchronize()
    return rel(out, ref)


def section(name):
    print("\n==== " + name, flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    if which in ("all", "gemm"):
        diag_gemm()
    if which in ("all", "model"):
        diag_model()
    if which in ("all", "timing"):
        diag_timing()


def diag_gemm():
    section("tcgen05 GEMM variants")
    for bn in (256, 128):
        _lib.load().mmer_debug_set(_lib.DEBUG_FORCE_BN, bn)
        for (amaj, bmaj) in ((0, 0), (0, 1), (1, 1)):
            for swap in ((0,) if (amaj, bmaj) == (0, 0) else (0, 1)):
                _lib.load().mmer_debug_set(_lib.DEBUG_MN_SWAP, swap)
                for (M, N, K) in ((256, 256, 64), (384, 512, 512), (300, 200, 136), (4096, 1536, 512)):
                    try:
                        e = gemm_case(M, N, K, amaj, bmaj, torch.bfloat16)
                        print(f"bn={bn} majors=({amaj},{bmaj}) swap={swap} M{M} N{N} K{K}: rel={e:.3e}", flush=True)
                    except Exception as ex:
                        print(f"bn={bn} majors=({amaj},{bmaj}) swap={swap} M{M} N{N} K{K}: EXC {ex}", flush=True)
                        raise
    _lib.load().mmer_debug_set(_lib.DEBUG_MN_SWAP, 0)
    _lib.load().mmer_debug_set(_lib.DEBUG_FORCE_BN, 0)
    print("wgrad accumulate (MN,MN) M512 N768 K8192:", gemm_case(512, 768, 8192, 1, 1, torch.bfloat16, acc=True))
    section("SIMT GEMM")
    for (amaj, bmaj) in ((0, 0), (0, 1), (1, 1)):
        print((amaj, bmaj), gemm_case(300, 200, 136, amaj, bmaj, torch.float32))



def diag_model():
    section("model parity fp32 vs oracle (v2)")
    from oracle import fusion_oracle as O
    import detgen
    for variant in ("v2", "v1"):
        try:
            B, T = 8, 5
            dims = dict(max_seq_len=T + 1)
            if variant == "v2":
                dims["hidden"] = 512
                model = mm.MultimodalEmotionModel(max_seq_len=T + 1, classifier_hidden_dim=512, fusion_dropout=0.0,
                                                  classifier_dropout=0.0)
            else:
                model = mm.v1.MultimodalEmotionModel(max_seq_len=T + 1)
                model.fusion.dropout = 0.0
                model.classifier.dropout = 0.0
            P = {k: torch.from_numpy(np.asarray(v)) for k, v in detgen.make_params(variant, **dims).items()}
            model.load_state_dict(P, strict=True)
            model.cuda().train()
            v, a, m, y = detgen.make_batch(B, T, tag="diag")
            video, audio, mask, labels = (torch.from_numpy(x) for x in (v, a, m, y))
            P64 = {k: (t.double() if t.is_floating_point() else t) for k, t in P.items()}
            if variant == "v2":
                _, lo, fo, at = O.model_forward_v2(P64, video.double(), audio.double(), mask)
            else:
                _, lo, fo, at = O.model_forward_v1(P64, video.double(), audio.double(), mask, training=True)
            for cdt in (torch.float32, torch.bfloat16):
                model.compute_dtype = cdt
                vg = video.cuda().requires_grad_(True)
                ag = audio.cuda().requires_grad_(True)
                probs, logits, attn = model(vg, ag, mask=mask.cuda(), return_attn=True)
                print(variant, cdt, "logits rel", rel(logits.cpu(), lo), "attn rel", rel(attn["layers"].cpu(), at))
                crit = mm.FocalLoss(gamma=2.0)
                loss = crit(logits, labels.cuda())
                model.zero_grad()
                loss.backward()
                # oracle grads
                leaf = {k: t.clone().requires_grad_(True) for k, t in O.trainable(P64).items()}
                full = dict(P64); full.update(leaf)
                vd = video.double().requires_grad_(True); ad = audio.double().requires_grad_(True)
                if variant == "v2":
                    _, l2, _, _ = O.model_forward_v2(full, vd, ad, mask)
                else:
                    _, l2, _, _ = O.model_forward_v1(full, vd, ad, mask, training=True)
                lref = O.focal_loss(l2, labels)
                lref.backward()
                print("   loss", float(loss), float(lref))
                worst = 0
                for k, p in model.named_parameters():
                    g = leaf[k].grad
                    e = float((p.grad.cpu().double() - g).norm() / (g.norm() + 1e-9 * 1))
                    if e > (1e-3 if cdt == torch.float32 else 5e-2) and float(g.norm()) > 1e-7:
                        print("   GRAD MISMATCH", k, e, float(g.norm()))
                    worst = max(worst, e if float(g.norm()) > 1e-7 else 0)
                print("   worst param grad rel-norm err", worst)
                print("   dvideo rel", rel(vg.grad.cpu(), vd.grad), "daudio rel", rel(ag.grad.cpu(), ad.grad))
        except Exception:
            traceback.print_exc()



def diag_timing():
    section("train step timing cfg2 (B=4096,T=16,bf16)")
    try:
        B, T = 4096, 16
        model = mm.MultimodalEmotionModel(max_seq_len=T + 1, classifier_hidden_dim=512).cuda().train()
        step = mm.FusedTrainStep(model, loss="focal", alpha=torch.tensor([1, 1, 1, 1, 1.2, 1.2]))
        video = torch.randn(B, T, 768, device=dev, dtype=torch.bfloat16)
        audio = torch.randn(B, 1024, device=dev, dtype=torch.bfloat16)
        labels = torch.randint(0, 6, (B,), device=dev)
        for _ in range(3):
            loss, _ = step.step(video, audio, None, labels)
        torch.cuda.synchronize()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            loss, _ = step.step(video, audio, None, labels)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"ms/step {ms:.3f}  samples/s {B / ms * 1e3:.0f}  loss {float(loss):.4f}  host {1e3 * (time.time() - t0) / n:.3f} ms")
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
