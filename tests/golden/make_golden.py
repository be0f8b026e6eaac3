"""Generate golden vectors from the UNMODIFIED reference classes.

Run once in the build container (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

It imports /root/reference/train.py and train2.py (with a stub for the missing
``captum`` package, which the model classes do not use), loads the deterministic
weights of tests/detgen.py into the reference modules with ``load_state_dict(strict=
True)``, runs forward / loss / backward / optimizer step with the reference's own code
(nn.TransformerEncoder, F.cross_entropy, nn.CrossEntropyLoss, optim.Adam,
clip_grad_norm_) and stores the results as small .npz fixtures next to this script.
Full gradients are 30 MB, so each parameter's gradient is stored as
[l2 norm, sum, first 16 values] (``grad/<name>``), plus, element-wise: the whole tensor when it has at most
detgen.FULL_GRAD_MAX elements (``gradfull/<name>``) and 256 seeded sparse +-1 projections otherwise
(``gradproj/<name>``, plan in detgen.projection_plan).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import detgen  # noqa: E402

REF = "/root/reference"


def import_reference():
    cap = types.ModuleType("captum")
    attr = types.ModuleType("captum.attr")
    attr.IntegratedGradients = object
    cap.attr = attr
    sys.modules.setdefault("captum", cap)
    sys.modules.setdefault("captum.attr", attr)
    sys.path.insert(0, REF)
    import train as ref_v1  # noqa
    import train2 as ref_v2  # noqa
    return ref_v1, ref_v2


def summarize(t: torch.Tensor) -> np.ndarray:
    f = t.detach().double().flatten()
    head = f[:16].numpy()
    head = np.pad(head, (0, 16 - head.size))
    return np.concatenate([[float(f.norm()), float(f.sum())], head])


def zero_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0


def load(model, params):
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}
    model.load_state_dict(sd, strict=True)


def attn_oracle(model, video, audio, mask):
    """re-invoke each layer's self_attn with need_weights=True."""
    store = []

    def hook(mod, args, kwargs, out):
        kw = dict(kwargs)
        kw["need_weights"] = True
        kw["average_attn_weights"] = False
        _, w = torch.nn.MultiheadAttention.forward(mod, *args, **kw)
        store.append(w.detach())

    hs = [l.self_attn.register_forward_hook(hook, with_kwargs=True) for l in model.fusion.transformer.layers]
    with torch.no_grad():
        model(video, audio, mask=mask)
    for h in hs:
        h.remove()
    return torch.stack(store)  # (L,B,H,S,S)


def run_case(name, variant, ref_mod, B, T, use_mask, steps_cfg):
    torch.manual_seed(0)
    torch.set_num_threads(4)
    max_seq_len = T + 1
    if variant == "v2":
        model = ref_mod.MultimodalEmotionModel(max_seq_len=max_seq_len, fusion_num_layers=2,
                                               fusion_dropout=0.0, classifier_hidden_dim=512,
                                               classifier_dropout=0.0)
        params = detgen.make_params("v2", max_seq_len=max_seq_len, hidden=512)
    else:
        model = ref_mod.MultimodalEmotionModel(max_seq_len=max_seq_len)
        params = detgen.make_params("v1", max_seq_len=max_seq_len)
    zero_dropout(model)
    load(model, params)
    v, a, m, y = detgen.make_batch(B, T, tag=name)
    video, audio = torch.from_numpy(v), torch.from_numpy(a)
    mask = torch.from_numpy(m) if use_mask else None
    labels = torch.from_numpy(y)
    alpha = torch.tensor([1, 1, 1, 1, 1.2, 1.2], dtype=torch.float32)
    out = {"B": B, "T": T, "use_mask": int(use_mask)}

    # ---- eval forward + attention oracle
    model.eval()
    with torch.no_grad():
        probs, logits, _ = model(video, audio, mask=mask)
        fused, _ = model.fusion(video, audio, mask=mask)
    out["eval/probs"], out["eval/logits"], out["eval/fused"] = probs.numpy(), logits.numpy(), fused.numpy()
    attn = attn_oracle(model, video, audio, mask)
    out["eval/attn_last_mean"] = attn[-1].mean(dim=1).numpy()          # (B,S,S)
    out["eval/attn_layer0_head0"] = attn[0][:, 0].numpy()

    # ---- train-mode forward, losses, backward (input grads too)
    model.train()
    load(model, params)  # restore BN running stats
    video.requires_grad_(True)
    audio.requires_grad_(True)
    probs, logits, _ = model(video, audio, mask=mask)
    out["train/logits"], out["train/probs"] = logits.detach().numpy(), probs.detach().numpy()
    focal = ref_mod.FocalLoss(gamma=2.0)
    focal_a = ref_mod.FocalLoss(gamma=2.0, alpha=alpha)
    wce = torch.nn.CrossEntropyLoss(weight=alpha)
    out["loss/focal"] = focal(logits, labels).item()
    out["loss/focal_alpha"] = focal_a(logits, labels).item()
    out["loss/focal_sum"] = ref_mod.FocalLoss(gamma=2.0, reduction="sum")(logits, labels).item()
    out["loss/focal_none"] = ref_mod.FocalLoss(gamma=2.0, alpha=alpha, reduction="none")(logits, labels).detach().numpy()
    out["loss/wce"] = wce(logits, labels).item()
    dl = torch.autograd.grad(focal_a(logits, labels), logits, retain_graph=True)[0]
    out["dlogits/focal_alpha"] = dl.numpy()
    dl = torch.autograd.grad(wce(logits, labels), logits, retain_graph=True)[0]
    out["dlogits/wce"] = dl.numpy()
    if variant == "v1":
        for k, val in model.state_dict().items():
            if "running" in k or "tracked" in k:
                out["bn_after_fwd/" + k] = val.numpy().copy()

    loss_kind = steps_cfg["loss"]
    loss = {"focal": focal, "focal_alpha": focal_a, "wce": wce}[loss_kind](logits, labels)
    model.zero_grad()
    loss.backward()
    out["grad_in/video"], out["grad_in/audio"] = video.grad[:4].numpy(), audio.grad[:4].numpy()  # first 4 samples
    out["grad_in/video_sum"], out["grad_in/audio_sum"] = summarize(video.grad), summarize(audio.grad)
    for k, p in model.named_parameters():
        out["grad/" + k] = summarize(p.grad)
        if p.numel() <= detgen.FULL_GRAD_MAX:
            out["gradfull/" + k] = p.grad.detach().numpy().copy()
        else:
            out["gradproj/" + k] = detgen.project(p.grad.detach().numpy(), k)

    # ---- optimizer step exactly as the reference training loop does it
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    if steps_cfg.get("clip"):
        tn = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        out["clip/total_norm"] = float(tn)
    opt.step()
    for k, p in model.named_parameters():
        out["step1/" + k] = summarize(p)
        out["delta1/" + k] = summarize(p.detach() - torch.from_numpy(params[k]))
    # second step on the same batch (exercises Adam state / bias correction)
    opt.zero_grad()
    video.grad = None
    audio.grad = None
    _, logits2, _ = model(video, audio, mask=mask)
    loss2 = {"focal": focal, "focal_alpha": focal_a, "wce": wce}[loss_kind](logits2, labels)
    out["step2/loss"] = loss2.item()
    out["step2/logits"] = logits2.detach().numpy()
    loss2.backward()
    if steps_cfg.get("clip"):
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
    opt.step()
    for k, p in model.named_parameters():
        out["step2p/" + k] = summarize(p)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", float(loss), "->", float(loss2), "files ok")


def main():
    ref_v1, ref_v2 = import_reference()
    run_case("v2_b8_t5_mask", "v2", ref_v2, 8, 5, True, {"loss": "wce", "clip": True})
    run_case("v2_b4_t16_nomask", "v2", ref_v2, 4, 16, False, {"loss": "focal_alpha"})
    run_case("v1_b8_t5_mask", "v1", ref_v1, 8, 5, True, {"loss": "focal"})
    run_case("v1_b32_t16_cfg1", "v1", ref_v1, 32, 16, True, {"loss": "focal"})


if __name__ == "__main__":
    main()
