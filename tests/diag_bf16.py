"""bf16 gradient-error diagnosis (run under gpurun): per-parameter relative gradient error of the bf16 path
against the fp64 oracle on bf16-rounded weights, for two GEMM epilogue routes (tcgen05 staged epilogue, tcgen05 direct stores)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import detgen  # noqa: E402
import mmer_b200 as mm  # noqa: E402
from mmer_b200 import _lib  # noqa: E402
from oracle import fusion_oracle as O  # noqa: E402

ALPHA = torch.tensor([1, 1, 1, 1, 1.2, 1.2])


def main():
    B, T = 16, 16
    P = {k: torch.from_numpy(np.asarray(v)) for k, v in detgen.make_params("v2", max_seq_len=T + 1, hidden=512).items()}
    v, a, m, y = detgen.make_batch(B, T, tag="bf16case")
    video, audio, mask, labels = (torch.from_numpy(x) for x in (v, a, m, y))
    last = "classifier.net.8.weight"
    rounded = {k: (t.bfloat16().float() if (t.dim() == 2 and k != last) else t) for k, t in P.items()}
    leaf = {k: t.double().clone().requires_grad_(True) for k, t in O.trainable(rounded).items()}
    full = {k: (t.double() if t.is_floating_point() else t) for k, t in rounded.items()}
    full.update(leaf)
    vr = video.bfloat16().double().requires_grad_(True)
    ar = audio.bfloat16().double().requires_grad_(True)
    _, lref, _, _ = O.model_forward_v2(full, vr, ar, mask)
    O.focal_loss(lref, labels, 2.0, ALPHA.double()).backward()
    leaf_q = {k: t.double().clone().requires_grad_(True) for k, t in O.trainable(rounded).items()}
    full_q = dict(full)
    full_q.update(leaf_q)
    vq = video.bfloat16().double().requires_grad_(True)
    aq = audio.bfloat16().double().requires_grad_(True)
    with O.storage_rounding(torch.bfloat16):
        _, lq, _, _ = O.model_forward_v2(full_q, vq, aq, mask)
        O.focal_loss(lq, labels, 2.0, ALPHA.double()).backward()
    lib = _lib.load()
    for route, key in (("tcgen05 staged", None), ("tcgen05 direct", _lib.DEBUG_DIRECT_STORE)):
        for k in (_lib.DEBUG_DIRECT_STORE,):
            lib.mmer_debug_set(k, 0)
        if key is not None:
            lib.mmer_debug_set(key, 1)
        model = mm.MultimodalEmotionModel(max_seq_len=T + 1, classifier_hidden_dim=512, fusion_dropout=0.0,
                                          classifier_dropout=0.0)
        model.load_state_dict(P)
        model.cuda().train()
        model.compute_dtype = torch.bfloat16
        vg, ag = video.cuda().requires_grad_(True), audio.cuda().requires_grad_(True)
        _, logits, _ = model(vg, ag, mask=mask.cuda())
        mm.FocalLoss(2.0, ALPHA.cuda())(logits, labels.cuda()).backward()
        print(f"== {route}: logits err {float((logits.detach().cpu().double() - lref.detach()).abs().max()):.4e}")
        for k, p in model.named_parameters():
            gr = leaf[k].grad
            if float(gr.norm()) < 1e-6:
                continue
            err = float((p.grad.cpu().double() - gr).norm() / gr.norm())
            gq = leaf_q[k].grad
            errq = float((p.grad.cpu().double() - gq).norm() / gq.norm())
            print(f"   {k:55s} vs exact {err:.4f}   vs bf16-storage oracle {errq:.4f}")
        print(f"   dvideo {float((vg.grad.cpu().double() - vr.grad).norm() / vr.grad.norm()):.4f} "
              f"{float((vg.grad.cpu().double() - vq.grad).norm() / vq.grad.norm()):.4f}")
        print(f"   logits vs bf16-storage oracle {float((logits.detach().cpu().double() - lq.detach()).abs().max()):.4e}")


if __name__ == "__main__":
    main()
