"""Golden vectors for the `use_layernorm=False` variants of the train2.py classes, from the UNMODIFIED reference:
CrossModalFusion with nn.Identity norms (train2.py:96,104-105,121) and EmotionClassifier with nn.BatchNorm1d
(train2.py:208,215).  Run once in the build container (needs /root/reference):

    python tests/golden/make_golden_nolayernorm.py

Writes tests/golden/v2_nolayernorm_b8_t5_mask.npz: eval / train outputs of each module on its own and of the two chained
(fusion -> classifier -> weighted cross-entropy), every parameter gradient (whole tensor up to detgen.FULL_GRAD_MAX
elements, 256 seeded projections above), input gradients and the BatchNorm running statistics after one training
forward."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import detgen  # noqa: E402
from make_golden import import_reference, zero_dropout  # noqa: E402

B, T, HID = 8, 5, 512


def put_grads(out, prefix, named):
    for k, p in named:
        g = p.grad.detach().numpy()
        if g.size <= detgen.FULL_GRAD_MAX:
            out[f"{prefix}/gradfull/{k}"] = g.copy()
        else:
            out[f"{prefix}/gradproj/{k}"] = detgen.project(g, k)


def main():
    _, ref = import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(4)
    params = detgen.make_params("v2", max_seq_len=T + 1, hidden=HID)
    fusion = ref.CrossModalFusion(num_layers=2, dropout=0.0, max_seq_len=T + 1, use_layernorm=False)
    head = ref.EmotionClassifier(input_dim=512, hidden_dim=HID, dropout=0.0, use_layernorm=False)
    zero_dropout(fusion)
    zero_dropout(head)
    fsd = {k[len("fusion."):]: torch.from_numpy(v) for k, v in params.items()
           if k.startswith("fusion.") and "norm_video" not in k and "norm_audio" not in k and "out_norm" not in k}
    fusion.load_state_dict(fsd, strict=True)
    hsd = {k[len("classifier."):]: torch.from_numpy(v) for k, v in params.items() if k.startswith("classifier.")}
    for i in (1, 5):   # BatchNorm buffers (the LayerNorm weights / biases of the detgen set become the affine parameters)
        hsd[f"net.{i}.running_mean"] = torch.from_numpy(detgen.det_array((HID,), f"nolayernorm/rm{i}", 0.1))
        hsd[f"net.{i}.running_var"] = torch.from_numpy(detgen.det_array((HID,), f"nolayernorm/rv{i}", 0.25, 1.0))
        hsd[f"net.{i}.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    head.load_state_dict(hsd, strict=True)
    v, a, m, y = detgen.make_batch(B, T, tag="nolayernorm")
    video, audio, mask, labels = torch.from_numpy(v), torch.from_numpy(a), torch.from_numpy(m), torch.from_numpy(y)
    fused_in = torch.from_numpy(detgen.det_array((B, 512), "nolayernorm/fused_in", 1.0))
    wr = torch.from_numpy(detgen.det_array((B, 512), "nolayernorm/wr", 1.0))
    alpha = torch.tensor([1, 1, 1, 1, 1.2, 1.2], dtype=torch.float32)
    wce = torch.nn.CrossEntropyLoss(weight=alpha)
    out = {"B": B, "T": T, "hidden": HID}
    for k, t in hsd.items():
        if "running" in k:
            out["bn_init/" + k] = t.numpy().copy()

    # ---- fusion alone
    fusion.eval()
    with torch.no_grad():
        out["fusion/eval_fused"] = fusion(video, audio, mask=mask)[0].numpy()
        out["fusion/eval_fused_nomask"] = fusion(video, audio)[0].numpy()
    fusion.train()
    vg, ag = video.clone().requires_grad_(True), audio.clone().requires_grad_(True)
    fused = fusion(vg, ag, mask=mask)[0]
    out["fusion/train_fused"] = fused.detach().numpy()
    fusion.zero_grad()
    (fused * wr).sum().backward()
    put_grads(out, "fusion", fusion.named_parameters())
    out["fusion/grad_video"], out["fusion/grad_audio"] = vg.grad.numpy().copy(), ag.grad.numpy().copy()

    # ---- classifier alone: eval (running statistics), then one training forward / backward
    head.eval()
    with torch.no_grad():
        out["head/eval_logits"] = head(fused_in).numpy()
    head.train()
    fg = fused_in.clone().requires_grad_(True)
    logits = head(fg)
    out["head/train_logits"] = logits.detach().numpy()
    head.zero_grad()
    loss = wce(logits, labels)
    out["head/loss"] = loss.item()
    loss.backward()
    put_grads(out, "head", head.named_parameters())
    out["head/grad_fused"] = fg.grad.numpy().copy()
    for k, t in head.state_dict().items():
        if "running" in k or "tracked" in k:
            out["head/bn_after_fwd/" + k] = t.numpy().copy()

    # ---- chained (what MultimodalEmotionModel.forward does with these two sub-modules, train2.py:281-292)
    head.load_state_dict(hsd, strict=True)
    head.train()
    vg, ag = video.clone().requires_grad_(True), audio.clone().requires_grad_(True)
    logits = head(fusion(vg, ag, mask=mask)[0])
    out["chain/train_logits"] = logits.detach().numpy()
    out["chain/train_probs"] = torch.softmax(logits, -1).detach().numpy()
    fusion.zero_grad()
    head.zero_grad()
    loss = wce(logits, labels)
    out["chain/loss"] = loss.item()
    loss.backward()
    put_grads(out, "chain", [("fusion." + k, p) for k, p in fusion.named_parameters()] +
              [("classifier." + k, p) for k, p in head.named_parameters()])
    out["chain/grad_video"], out["chain/grad_audio"] = vg.grad.numpy().copy(), ag.grad.numpy().copy()
    head.eval()
    fusion.eval()
    with torch.no_grad():
        out["chain/eval_logits_after"] = head(fusion(video, audio, mask=mask)[0]).numpy()   # running stats moved once
    np.savez_compressed(os.path.join(HERE, "v2_nolayernorm_b8_t5_mask.npz"), **out)
    print("v2_nolayernorm_b8_t5_mask", "head loss", out["head/loss"], "chain loss", out["chain/loss"], len(out), "arrays")


if __name__ == "__main__":
    main()
