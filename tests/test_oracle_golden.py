"""Pin the CPU oracle (oracle/fusion_oracle.py) against golden vectors produced by the
unmodified reference classes (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

import detgen
from oracle import fusion_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = [("v2_b8_t5_mask", "v2", "wce", True), ("v2_b4_t16_nomask", "v2", "focal_alpha", False),
         ("v1_b8_t5_mask", "v1", "focal", False), ("v1_b32_t16_cfg1", "v1", "focal", False)]
ALPHA = torch.tensor([1, 1, 1, 1, 1.2, 1.2], dtype=torch.float64)


def summarize(t):
    f = t.detach().double().flatten()
    head = f[:16].numpy()
    head = np.pad(head, (0, 16 - head.size))
    return np.concatenate([[float(f.norm()), float(f.sum())], head])


def load_case(name, variant):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    B, T = int(g["B"]), int(g["T"])
    dims = dict(max_seq_len=T + 1)
    if variant == "v2":
        dims["hidden"] = 512
    P = {k: torch.from_numpy(np.asarray(v)) for k, v in detgen.make_params(variant, **dims).items()}
    v, a, m, y = detgen.make_batch(B, T, tag=name)
    mask = torch.from_numpy(m) if int(g["use_mask"]) else None
    return g, P, torch.from_numpy(v), torch.from_numpy(a), mask, torch.from_numpy(y)


def to64(P):
    return {k: (v.double() if v.is_floating_point() else v) for k, v in P.items()}


@pytest.mark.parametrize("name,variant,loss,clip", CASES)
def test_eval_forward_and_attention(name, variant, loss, clip):
    g, P, video, audio, mask, labels = load_case(name, variant)
    P = to64(P)
    if variant == "v2":
        probs, logits, fused, attn = O.model_forward_v2(P, video.double(), audio.double(), mask)
    else:
        probs, logits, fused, attn = O.model_forward_v1(P, video.double(), audio.double(), mask, training=False)
    np.testing.assert_allclose(logits.numpy(), g["eval/logits"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(probs.numpy(), g["eval/probs"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(fused.numpy(), g["eval/fused"], rtol=2e-5, atol=2e-5)
    assert (probs.argmax(1).numpy() == g["eval/probs"].argmax(1)).all()
    last, audio_row = O.cross_modal_attention(attn)
    np.testing.assert_allclose(last.numpy(), g["eval/attn_last_mean"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(attn[0][:, 0].numpy(), g["eval/attn_layer0_head0"], rtol=1e-4, atol=1e-6)
    assert (audio_row.argmax(1).numpy() == g["eval/attn_last_mean"][:, -1, :].argmax(1)).all()
    if mask is not None:  # masked keys get exactly zero weight
        S = mask.shape[1] + 1
        full = torch.cat([mask, torch.zeros(mask.shape[0], 1, dtype=torch.bool)], 1)
        assert float(last.masked_select(full.view(-1, 1, S).expand_as(last)).abs().max()) == 0.0


@pytest.mark.parametrize("name,variant,loss,clip", CASES)
def test_losses_and_dlogits(name, variant, loss, clip):
    g, *_ = load_case(name, variant)
    logits = torch.from_numpy(g["train/logits"]).double()
    labels = torch.from_numpy(detgen.make_batch(int(g["B"]), int(g["T"]), tag=name)[3])
    assert abs(float(O.focal_loss(logits, labels)) - float(g["loss/focal"])) < 2e-6
    assert abs(float(O.focal_loss(logits, labels, alpha=ALPHA)) - float(g["loss/focal_alpha"])) < 2e-6
    assert abs(float(O.focal_loss(logits, labels, reduction="sum")) - float(g["loss/focal_sum"])) < 2e-5
    np.testing.assert_allclose(O.focal_loss(logits, labels, alpha=ALPHA, reduction="none").numpy(),
                               g["loss/focal_none"], rtol=1e-5, atol=1e-6)
    assert abs(float(O.weighted_ce(logits, labels, ALPHA)) - float(g["loss/wce"])) < 2e-6
    np.testing.assert_allclose(O.focal_loss_grad(logits, labels, alpha=ALPHA).numpy(),
                               g["dlogits/focal_alpha"], rtol=1e-4, atol=1e-7)
    lg = logits.clone().requires_grad_(True)
    O.weighted_ce(lg, labels, ALPHA).backward()
    np.testing.assert_allclose(lg.grad.numpy(), g["dlogits/wce"], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("name,variant,loss,clip", CASES)
def test_train_step_grads_and_adam(name, variant, loss, clip):
    g, P, video, audio, mask, labels = load_case(name, variant)
    P64 = to64(P)
    kw = dict(variant=variant, loss="wce" if loss == "wce" else "focal",
              alpha=ALPHA if loss in ("wce", "focal_alpha") else None, clip=1.0 if clip else None)
    vid = video.double().requires_grad_(True)
    aud = audio.double().requires_grad_(True)
    P1, S1, l1, logits, grads = O.train_step(P64, {}, 1, vid, aud, mask, labels, **kw)
    np.testing.assert_allclose(logits.numpy(), g["train/logits"], rtol=2e-5, atol=2e-5)
    for k, gr in grads.items():
        ref = g["grad/" + k]
        tol = 2e-4 * ref[0] + 1e-6  # pre-BatchNorm biases have an exactly-zero true gradient
        assert abs(summarize(gr)[0] - ref[0]) <= tol, k
        np.testing.assert_allclose(summarize(gr)[2:], ref[2:], rtol=0, atol=5e-4 * ref[0] + 1e-6, err_msg=k)
        # every element (tensors <= 64 K elements) / 256 seeded projections (larger ones); fp64 oracle vs fp32 reference
        detgen.check_gradient_elementwise(g, k, gr.detach().numpy(), rel=2e-3)
    if clip:
        total, _ = O.clip_coef(list(grads.values()), 1.0)
        assert abs(total - float(g["clip/total_norm"])) < 1e-4 * total
    for k in grads:
        ref = g["delta1/" + k]
        # Adam's first step moves every coordinate by ~lr; compare the update itself
        np.testing.assert_allclose(summarize(P1[k] - P64[k])[2:], ref[2:], rtol=0, atol=2e-6, err_msg=k)
        assert abs(summarize(P1[k] - P64[k])[0] - ref[0]) <= 2e-3 * ref[0] + 1e-9, k
    # second step on the same batch: Adam state and bias correction
    P2, S2, l2, logits2, _ = O.train_step(P1, S1, 2, vid, aud, mask, labels, **kw)
    assert abs(l2 - float(g["step2/loss"])) < 5e-4
    np.testing.assert_allclose(logits2.numpy(), g["step2/logits"], rtol=0, atol=3e-3)
    if variant == "v1":
        for k in P1:
            if "running" in k:
                np.testing.assert_allclose(P1[k].numpy(), g["bn_after_fwd/" + k], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,variant", [("v2_b8_t5_mask", "v2"), ("v1_b8_t5_mask", "v1")])
def test_input_gradients(name, variant):
    g, P, video, audio, mask, labels = load_case(name, variant)
    P64 = to64(P)
    vid = video.double().requires_grad_(True)
    aud = audio.double().requires_grad_(True)
    if variant == "v2":
        _, logits, _, _ = O.model_forward_v2(P64, vid, aud, mask)
        lval = O.weighted_ce(logits, labels, ALPHA)
    else:
        _, logits, _, _ = O.model_forward_v1(P64, vid, aud, mask, training=True)
        lval = O.focal_loss(logits, labels)
    lval.backward()
    np.testing.assert_allclose(vid.grad[:4].numpy(), g["grad_in/video"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(aud.grad[:4].numpy(), g["grad_in/audio"], rtol=0, atol=1e-6)
    if variant == "v2":  # padded video rows receive exactly zero gradient
        assert float(vid.grad[mask].abs().max()) == 0.0


def test_bf16_storage_noise_floor():
    """The exact oracle vs the same oracle with bf16 rounding at every activation / activation-gradient tensor
    boundary (what the bf16 product path stores between kernels): logits move by < 2e-2, gradients by 4-12 % in
    norm because ReLU masks of near-zero pre-activations flip.  This pins the noise floor that the GPU bf16
    gradient test (tests/test_gpu_model.py::test_bf16_against_oracle_on_rounded_weights) is judged against."""
    B, T = 16, 16
    P = {k: torch.from_numpy(np.asarray(v)) for k, v in detgen.make_params("v2", max_seq_len=T + 1, hidden=512).items()}
    v, a, m, y = detgen.make_batch(B, T, tag="bf16case")
    video, audio, mask, labels = (torch.from_numpy(x) for x in (v, a, m, y))
    alpha = torch.tensor([1, 1, 1, 1, 1.2, 1.2]).double()
    last = "classifier.net.8.weight"
    rounded = {k: (t.bfloat16().float() if (t.dim() == 2 and k != last) else t) for k, t in P.items()}

    def run(emulate):
        leaf = {k: t.double().clone().requires_grad_(True) for k, t in O.trainable(rounded).items()}
        full = {k: (t.double() if t.is_floating_point() else t) for k, t in rounded.items()}
        full.update(leaf)
        vr, ar = video.bfloat16().double(), audio.bfloat16().double()
        if emulate:
            with O.storage_rounding(torch.bfloat16):
                _, logits, _, _ = O.model_forward_v2(full, vr, ar, mask)
                O.focal_loss(logits, labels, 2.0, alpha).backward()
        else:
            _, logits, _, _ = O.model_forward_v2(full, vr, ar, mask)
            O.focal_loss(logits, labels, 2.0, alpha).backward()
        return logits.detach(), {k: t.grad for k, t in leaf.items()}

    l0, g0 = run(False)
    l1, g1 = run(True)
    assert float((l0 - l1).abs().max()) < 2e-2 * float(l0.abs().max())
    errs = {k: float((g0[k] - g1[k]).norm() / g0[k].norm()) for k in g0 if k.startswith("fusion.transformer")}
    assert 0.04 < min(errs.values()) and max(errs.values()) < 0.12, errs


def test_stock_torch_restatement_matches_reference_golden():
    """oracle/eager_torch.py (the reference architecture on stock torch.nn modules, used by bench.py as the
    'what the reference launches on a GPU' comparison) reproduces the reference's golden logits."""
    from oracle import eager_torch as E
    for name in ("v2_b8_t5_mask", "v2_b4_t16_nomask"):
        g, P, video, audio, mask, labels = load_case(name, "v2")
        T = video.shape[1]
        model = E.EagerModel(max_seq_len=T + 1, classifier_hidden_dim=512).eval()
        model.load_state_dict(P, strict=True)
        with torch.no_grad():
            probs, logits = model(video, audio, mask)
        np.testing.assert_allclose(logits.numpy(), g["eval/logits"], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(probs.numpy(), g["eval/probs"], rtol=2e-4, atol=1e-6)
        lf = E.focal_loss(logits, labels, 2.0, ALPHA.float())
        np.testing.assert_allclose(float(lf), float(g["loss/focal_alpha"]), rtol=1e-4)


def test_stock_torch_restatement_of_train_py_matches_reference_golden():
    """oracle/eager_torch.py::EagerModelV1 (train.py's BatchNorm model on stock torch.nn; bench.py's cfg1 CPU leg)
    reproduces the reference's golden logits in eval and train mode, its loss and its BatchNorm running statistics."""
    from oracle import eager_torch as E
    for name in ("v1_b8_t5_mask", "v1_b32_t16_cfg1"):
        g, P, video, audio, mask, labels = load_case(name, "v1")
        T = video.shape[1]
        model = E.EagerModelV1(max_seq_len=T + 1, dropout=0.0)
        model.load_state_dict(P, strict=True)
        model.eval()
        with torch.no_grad():
            probs, logits = model(video, audio, mask)
        np.testing.assert_allclose(logits.numpy(), g["eval/logits"], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(probs.numpy(), g["eval/probs"], rtol=2e-4, atol=1e-6)
        model.train()
        probs, logits = model(video, audio, mask)
        np.testing.assert_allclose(logits.detach().numpy(), g["train/logits"], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(float(E.focal_loss(logits, labels, 2.0, None)), float(g["loss/focal"]), rtol=1e-4)
        for k, val in model.state_dict().items():
            if "running" in k:
                np.testing.assert_allclose(val.numpy(), g["bn_after_fwd/" + k], rtol=1e-4, atol=1e-6, err_msg=k)


# ---------------------------------------------------------------- Integrated Gradients (train2.py:776-866)
def _ig_case():
    from oracle import ig_oracle
    g = np.load(os.path.join(GOLD, "ig_v2_b8_t5_mask.npz"))
    _, P, video, audio, mask, _ = load_case("v2_b8_t5_mask", "v2")
    P64 = to64(P)
    fn = lambda v, a, mk: O.model_forward_v2(P64, v, a, mk)[1]  # noqa: E731
    return ig_oracle, g, fn, video.double(), audio.double(), mask


def test_integrated_gradients_oracle_matches_reference_model_golden():
    """The model restatement under the IG restatement == the unmodified reference model class under it
    (tests/golden/make_golden_ig.py), including the predicted-class target rule."""
    ig, g, fn, video, audio, mask = _ig_case()
    logits = fn(video, audio, mask)
    target = logits.argmax(dim=-1)
    assert np.array_equal(target.numpy(), g["target"])
    av, aa = ig.integrated_gradients(fn, (video, audio), (torch.zeros_like(video), torch.zeros_like(audio)), mask, target,
                                     int(g["n_steps"]))
    np.testing.assert_allclose(av.numpy(), g["attr_video"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(aa.numpy(), g["attr_audio"], rtol=0, atol=1e-9)
    assert float(av[mask].abs().max()) == 0.0           # padded frames get exactly zero attribution


def test_integrated_gradients_completeness_converges():
    """sum(attr) -> f(x) - f(baseline) as n_steps grows (slowly: LayerNorm makes the zero baseline a sharp corner)."""
    ig, g, fn, video, audio, mask = _ig_case()
    target = torch.from_numpy(g["target"])
    zero = (torch.zeros_like(video), torch.zeros_like(audio))
    gaps = []
    for n in (int(g["n_steps"]), 96):
        av, aa = ig.integrated_gradients(fn, (video, audio), zero, mask, target, n)
        gaps.append(float((av.flatten(1).sum(1) + aa.sum(1) - torch.from_numpy(g["delta"])).abs().max()))
    assert gaps[1] < 0.4 * gaps[0] and gaps[1] < 0.05, gaps


def test_gauss_legendre_schedule_and_aggregation():
    from oracle import ig_oracle
    from mmer_b200 import attribution as A
    for n in (1, 2, 7, 50):
        al, st = ig_oracle.gauss_legendre(n)
        al2, st2 = A.gauss_legendre_schedule(n)
        assert np.array_equal(np.asarray(al), al2) and np.array_equal(np.asarray(st), st2)
        assert abs(sum(st) - 1.0) < 1e-12 and all(0 < x < 1 for x in al)
        assert np.allclose(np.asarray(al) + np.asarray(al)[::-1], 1.0)
    # degree-(2n-1) exactness of the rule: integral of alpha^3 over [0,1] with two nodes
    al, st = ig_oracle.gauss_legendre(2)
    assert abs(sum(s * a ** 3 for a, s in zip(al, st)) - 0.25) < 1e-12
    av, aa = torch.randn(3, 4, 8), torch.randn(3, 16)
    for fn in (ig_oracle.aggregate_importances, A.aggregate_importances):
        vi, ai = fn(av, aa)
        assert torch.equal(vi, av.abs().sum(1)) and torch.equal(ai, aa.abs())
        vi, ai = fn(av, aa, abs_sum=False)
        assert torch.equal(vi, av.sum(1)) and torch.equal(ai, aa)
    with pytest.raises(ValueError):
        A.compute_attributions(object(), av, aa, baseline="mean")
