"""GPU parity of the drop-in modules against (1) golden vectors produced by the unmodified reference
classes (tests/golden/*.npz) and (2) the CPU oracle, on identical inputs and weights.

Tolerances (BASELINE.json north_star): logits / loss within 1e-4 relative in fp32 and 2e-2 in bf16;
argmax of predictions and of the attention weights identical; gradients within the same tolerances.
bf16 runs are compared with the fp32 oracle evaluated on the SAME bf16-rounded GEMM weights and inputs
(SURVEY.md 8c "oracle hygiene"): rounding the weights alone moves this model's gradients by 5-8 %,
which is quantisation of the problem, not kernel error.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import detgen  # noqa: E402
import mmer_b200 as mm  # noqa: E402
from oracle import fusion_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ALPHA = torch.tensor([1, 1, 1, 1, 1.2, 1.2])
CASES = [("v2_b8_t5_mask", "v2", "wce", True), ("v2_b4_t16_nomask", "v2", "focal_alpha", False),
         ("v1_b8_t5_mask", "v1", "focal", False), ("v1_b32_t16_cfg1", "v1", "focal", False)]


def summarize(t):
    f = t.detach().double().flatten().cpu()
    head = f[:16].numpy()
    head = np.pad(head, (0, 16 - head.size))
    return np.concatenate([[float(f.norm()), float(f.sum())], head])


def build(name, variant, dropout0=True):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    B, T = int(g["B"]), int(g["T"])
    if variant == "v2":
        model = mm.MultimodalEmotionModel(max_seq_len=T + 1, fusion_num_layers=2, classifier_hidden_dim=512,
                                          fusion_dropout=0.0, classifier_dropout=0.0)
        P = detgen.make_params("v2", max_seq_len=T + 1, hidden=512)
    else:
        model = mm.v1.MultimodalEmotionModel(max_seq_len=T + 1)
        model.fusion.dropout = 0.0
        model.classifier.dropout = 0.0
        P = detgen.make_params("v1", max_seq_len=T + 1)
    P = {k: torch.from_numpy(np.asarray(v)) for k, v in P.items()}
    model.load_state_dict(P, strict=True)
    model.cuda()
    v, a, m, y = detgen.make_batch(B, T, tag=name)
    mask = torch.from_numpy(m).cuda() if int(g["use_mask"]) else None
    return g, model, P, torch.from_numpy(v).cuda(), torch.from_numpy(a).cuda(), mask, torch.from_numpy(y).cuda()


def criterion(loss):
    if loss == "wce":
        return mm.WeightedCrossEntropyLoss(ALPHA.cuda())
    return mm.FocalLoss(gamma=2.0, alpha=ALPHA.cuda() if loss == "focal_alpha" else None)


@pytest.mark.parametrize("name,variant,loss,clip", CASES)
def test_fp32_eval_forward_matches_reference_golden(name, variant, loss, clip):
    g, model, P, video, audio, mask, labels = build(name, variant)
    model.eval()
    with torch.no_grad():
        probs, logits, attn = model(video, audio, mask=mask, return_attn=True)
        none_attn = model(video, audio, mask=mask)[2]
    assert none_attn is None                                   # reference returns None (train2.py:179)
    np.testing.assert_allclose(logits.cpu().numpy(), g["eval/logits"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(probs.cpu().numpy(), g["eval/probs"], rtol=1e-4, atol=1e-5)
    assert (probs.argmax(1).cpu().numpy() == g["eval/probs"].argmax(1)).all()
    np.testing.assert_allclose(attn["last_mean"].cpu().numpy(), g["eval/attn_last_mean"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(attn["layers"][0][:, 0].cpu().numpy(), g["eval/attn_layer0_head0"], rtol=1e-4, atol=1e-6)
    assert (attn["audio_row"].argmax(1).cpu().numpy() == g["eval/attn_last_mean"][:, -1, :].argmax(1)).all()
    if variant == "v2":
        with torch.no_grad():
            fused, _ = model.fusion(video, audio, mask=mask)
            logits2 = model.classifier(fused)
        np.testing.assert_allclose(fused.cpu().numpy(), g["eval/fused"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(logits2.cpu().numpy(), g["eval/logits"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name,variant,loss,clip", CASES)
def test_fp32_training_step_matches_reference_golden(name, variant, loss, clip):
    """forward, loss, backward (parameter + input grads), clip, Adam x2 -- the reference's hot loop."""
    g, model, P, video, audio, mask, labels = build(name, variant)
    model.train()
    video.requires_grad_(True)
    audio.requires_grad_(True)
    crit = criterion(loss)
    opt = mm.FusedAdam(model, lr=1e-4, weight_decay=1e-4, max_grad_norm=1.0 if clip else None)
    opt.zero_grad()
    probs, logits, _ = model(video, audio, mask=mask)
    np.testing.assert_allclose(logits.detach().cpu().numpy(), g["train/logits"], rtol=1e-4, atol=1e-4)
    lval = crit(logits, labels)
    key = {"wce": "loss/wce", "focal": "loss/focal", "focal_alpha": "loss/focal_alpha"}[loss]
    assert abs(float(lval) - float(g[key])) < 1e-4 * max(1.0, abs(float(g[key])))
    lval.backward()
    np.testing.assert_allclose(video.grad[:4].cpu().numpy(), g["grad_in/video"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(audio.grad[:4].cpu().numpy(), g["grad_in/audio"], rtol=0, atol=2e-6)
    if mask is not None and variant == "v2":
        assert float(video.grad[mask].abs().max()) == 0.0       # padded rows: exactly zero (LayerNorm model)
    # (train.py's BatchNorm statistics include the padded rows, so the reference itself gives them a small
    #  non-zero gradient; the golden comparison above covers those rows for v1.)
    for k, p in model.named_parameters():
        ref = g["grad/" + k]
        got = summarize(p.grad)
        assert abs(got[0] - ref[0]) <= 2e-4 * ref[0] + 2e-6, k
        np.testing.assert_allclose(got[2:], ref[2:], rtol=0, atol=5e-4 * ref[0] + 2e-6, err_msg=k)
        # element-wise: the whole tensor (<= 64 K elements) or 256 seeded sparse projections of it
        detgen.check_gradient_elementwise(g, k, p.grad.detach().cpu().numpy(), rel=5e-3)
    if variant == "v1":
        for k, val in model.state_dict().items():
            if "running" in k:
                np.testing.assert_allclose(val.cpu().numpy(), g["bn_after_fwd/" + k], rtol=1e-4, atol=1e-5, err_msg=k)
            if "tracked" in k:
                assert int(val) == int(g["bn_after_fwd/" + k])
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    # A bias that feeds straight into a BatchNorm has an analytically ZERO gradient (train.py:66-74,125: the batch
    # mean removes it).  What either implementation holds there is summation-order noise (~1e-9), which Adam's first
    # step g / (|g| + 1e-8) turns into an O(lr) update of arbitrary sign: not comparable, so check smallness instead.
    zero_grad_keys = {"fusion.video_proj.bias", "fusion.audio_proj.bias", "classifier.fc1.bias"} if variant == "v1" else set()
    for k in zero_grad_keys:
        assert float(dict(model.named_parameters())[k].grad.abs().max()) < 1e-6, k
    opt.step()
    for k, p in model.named_parameters():
        if k in zero_grad_keys:
            continue
        ref = g["delta1/" + k]
        got = summarize(p.detach() - before[k])
        np.testing.assert_allclose(got[2:], ref[2:], rtol=0, atol=3e-6, err_msg=k)
    # second step on the same batch
    opt.zero_grad()
    _, logits2, _ = model(video, audio, mask=mask)
    l2 = crit(logits2, labels)
    assert abs(float(l2) - float(g["step2/loss"])) < 2e-3
    np.testing.assert_allclose(logits2.detach().cpu().numpy(), g["step2/logits"], rtol=0, atol=1e-2)


@pytest.mark.parametrize("variant", ["v2", "v1"])
def test_bf16_against_oracle_on_rounded_weights(variant):
    B, T = 16, 16
    name = "bf16case"
    if variant == "v2":
        model = mm.MultimodalEmotionModel(max_seq_len=T + 1, classifier_hidden_dim=512, fusion_dropout=0.0,
                                          classifier_dropout=0.0)
        P = detgen.make_params("v2", max_seq_len=T + 1, hidden=512)
    else:
        model = mm.v1.MultimodalEmotionModel(max_seq_len=T + 1)
        model.fusion.dropout = model.classifier.dropout = 0.0
        P = detgen.make_params("v1", max_seq_len=T + 1)
    P = {k: torch.from_numpy(np.asarray(v)) for k, v in P.items()}
    model.load_state_dict(P)
    model.cuda().train()
    model.compute_dtype = torch.bfloat16
    v, a, m, y = detgen.make_batch(B, T, tag=name)
    video, audio, mask, labels = (torch.from_numpy(x) for x in (v, a, m, y))
    last = "classifier.net.8.weight" if variant == "v2" else "classifier.fc2.weight"
    rounded = {k: (t.bfloat16().float() if (t.dim() == 2 and k != last) else t) for k, t in P.items()}
    leaf = {k: t.double().clone().requires_grad_(True) for k, t in O.trainable(rounded).items()}
    full = {k: (t.double() if t.is_floating_point() else t) for k, t in rounded.items()}
    full.update(leaf)
    vr = video.bfloat16().double().requires_grad_(True)
    ar = audio.bfloat16().double().requires_grad_(True)
    if variant == "v2":
        _, lref, _, attn_ref = O.model_forward_v2(full, vr, ar, mask)
    else:
        _, lref, _, attn_ref = O.model_forward_v1(full, vr, ar, mask, training=True)
    loss_ref = O.focal_loss(lref, labels, 2.0, ALPHA.double())
    loss_ref.backward()

    vg, ag = video.cuda().requires_grad_(True), audio.cuda().requires_grad_(True)
    probs, logits, attn = model(vg, ag, mask=mask.cuda(), return_attn=True)
    loss = mm.FocalLoss(2.0, ALPHA.cuda())(logits, labels.cuda())
    loss.backward()
    scale = float(lref.abs().max())
    assert float((logits.detach().cpu().double() - lref.detach()).abs().max()) < 2e-2 * scale
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * float(loss_ref)
    assert (logits.argmax(1).cpu() == lref.argmax(1)).all()
    ref_row = attn_ref[-1].mean(1)[:, -1, :]
    top2 = ref_row.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 2e-2                      # skip numerically tied rows
    assert (attn["audio_row"].argmax(1).cpu()[clear] == ref_row.argmax(1)[clear]).all()
    # Gradients.  With every activation stored in bf16, a fraction ~3e-3 of the ReLU pre-activations sits within
    # rounding distance of zero and flips its mask relative to the exact evaluation; each flip moves a gradient
    # element by its full magnitude, so ANY bf16-storage evaluation of this model differs from the exact one by
    # sqrt(3e-3) ~ 6-8 % in gradient norm.  tests/test_oracle_golden.py::test_bf16_storage_noise_floor measures the
    # same 6-9 % on the fp64 oracle with bf16 rounding at the tensor boundaries (no GPU involved), so the bound
    # below is the noise floor of the problem, not slack for the kernels: the kernels themselves are held to 2e-2
    # op by op in test_gpu_ops.py and to 1e-4 end to end in fp32 mode above.
    worst, worst_cos = 0.0, 1.0
    all_got, all_ref = [], []
    for k, p in model.named_parameters():
        gr = leaf[k].grad
        if float(gr.norm()) < 1e-6:
            continue
        got = p.grad.cpu().double()
        all_got.append(got.flatten())
        all_ref.append(gr.flatten())
        err = float((got - gr).norm() / gr.norm())
        cos = float((got * gr).sum() / (got.norm() * gr.norm()))
        worst, worst_cos = max(worst, err), min(worst_cos, cos)
        # per tensor: bias gradients that are sums of LayerNorm-backward rows cancel strongly (their norm is a few
        # per cent of the terms'), which amplifies the same flip noise; hence the wider per-tensor band
        assert err < 0.2 and cos > 0.98, (k, err, cos)
        assert abs(float(got.norm() / gr.norm()) - 1.0) < 3e-2, k           # gradient magnitude
    got, gr = torch.cat(all_got), torch.cat(all_ref)
    total_err = float((got - gr).norm() / gr.norm())
    total_cos = float((got * gr).sum() / (got.norm() * gr.norm()))
    print(f"bf16 {variant}: whole-gradient error {total_err:.4f} cosine {total_cos:.5f}; worst tensor {worst:.4f} / {worst_cos:.5f}")
    assert total_err < 0.12 and total_cos > 0.992
    got = vg.grad.cpu().double()
    assert float((got - vr.grad).norm() / vr.grad.norm()) < 0.12
    assert float((got * vr.grad).sum() / (got.norm() * vr.grad.norm())) > 0.992


def test_fused_train_step_matches_reference_golden_fp32():
    name, variant = "v2_b8_t5_mask", "v2"
    g, model, P, video, audio, mask, labels = build(name, variant)
    model.train()
    step = mm.FusedTrainStep(model, lr=1e-4, weight_decay=1e-4, loss="wce", alpha=ALPHA, clip_grad_norm=1.0,
                             compute_dtype=torch.float32)
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    l1, _ = step.step(video, audio, mask, labels)
    assert abs(float(l1) - float(g["loss/wce"])) < 1e-4
    for k, p in model.named_parameters():
        got = summarize(p.detach() - before[k])
        np.testing.assert_allclose(got[2:], g["delta1/" + k][2:], rtol=0, atol=3e-6, err_msg=k)
    l2, _ = step.step(video, audio, mask, labels)
    assert abs(float(l2) - float(g["step2/loss"])) < 2e-3


def test_fused_train_step_bf16_learns_and_dropout_runs():
    torch.manual_seed(0)
    B, T = 256, 16
    model = mm.MultimodalEmotionModel(max_seq_len=T + 1, classifier_hidden_dim=512).cuda().train()
    step = mm.FusedTrainStep(model, lr=3e-4, loss="focal", alpha=ALPHA)
    gen = torch.Generator().manual_seed(1)
    video = torch.randn(B, T, 768, generator=gen).cuda().bfloat16()
    audio = torch.randn(B, 1024, generator=gen).cuda().bfloat16()
    labels = torch.randint(0, 6, (B,), generator=gen).cuda()
    losses = [float(step.step(video, audio, None, labels)[0]) for _ in range(30)]
    assert all(np.isfinite(losses))
    assert np.mean(losses[-5:]) < 0.6 * np.mean(losses[:3])       # memorises a fixed batch


def test_eval_input_gradients_for_integrated_gradients():
    """Captum's IG path (train2.py:808-836): eval mode, gradient of one logit w.r.t. the inputs."""
    g, model, P, video, audio, mask, labels = build("v2_b8_t5_mask", "v2")
    model.eval()
    video.requires_grad_(True)
    audio.requires_grad_(True)
    _, logits, _ = model(video, audio, mask=mask)
    target = logits.gather(1, labels.view(-1, 1)).sum()
    gv, ga = torch.autograd.grad(target, (video, audio))
    P64 = {k: (t.double() if t.is_floating_point() else t) for k, t in P.items()}
    vr, ar = video.detach().cpu().double().requires_grad_(True), audio.detach().cpu().double().requires_grad_(True)
    _, lref, _, _ = O.model_forward_v2(P64, vr, ar, mask.cpu())
    lref.gather(1, labels.cpu().view(-1, 1)).sum().backward()
    assert float((gv.cpu().double() - vr.grad).abs().max()) < 1e-5 * float(vr.grad.abs().max()) + 1e-7
    assert float((ga.cpu().double() - ar.grad).abs().max()) < 1e-5 * float(ar.grad.abs().max()) + 1e-7


def test_batch_padding_changes_audio_position_like_the_reference():
    """the audio token sits at pos_embed[T_padded]; trimming changes the logits."""
    g, model, P, video, audio, mask, labels = build("v2_b4_t16_nomask", "v2")
    model.eval()
    T = video.shape[1]
    m = torch.zeros(video.shape[0], T, dtype=torch.bool, device="cuda")
    m[:, T // 2:] = True
    with torch.no_grad():
        _, padded, _ = model(video, audio, mask=m)
        _, trimmed, _ = model(video[:, : T // 2].contiguous(), audio, mask=None)
    P64 = {k: (t.double() if t.is_floating_point() else t) for k, t in P.items()}
    _, rp, _, _ = O.model_forward_v2(P64, video.cpu().double(), audio.cpu().double(), m.cpu())
    _, rt, _, _ = O.model_forward_v2(P64, video[:, : T // 2].cpu().double(), audio.cpu().double(), None)
    assert float((padded.cpu().double() - rp).abs().max()) < 1e-4
    assert float((trimmed.cpu().double() - rt).abs().max()) < 1e-4
    assert float((rp - rt).abs().max()) > 1e-4                   # the two really differ


def test_sequence_longer_than_pos_embed_raises_like_reference():
    model = mm.MultimodalEmotionModel(max_seq_len=6).cuda()
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, 6, 768, device="cuda"), torch.zeros(2, 1024, device="cuda"))


def test_torch_optimizer_and_clip_work_on_flat_views():
    """The reference's own loop body (train2.py:570-578) runs unchanged on the drop-in model."""
    g, model, P, video, audio, mask, labels = build("v2_b8_t5_mask", "v2")
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    crit = mm.WeightedCrossEntropyLoss(ALPHA.cuda())
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    opt.zero_grad()
    _, logits, _ = model(video, audio, mask=mask)
    crit(logits, labels).backward()
    tn = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
    assert abs(float(tn) - float(g["clip/total_norm"])) < 1e-3 * float(g["clip/total_norm"])
    opt.step()
    for k, p in model.named_parameters():
        got = summarize(p.detach() - before[k])
        np.testing.assert_allclose(got[2:], g["delta1/" + k][2:], rtol=0, atol=3e-6, err_msg=k)


# ----------------------------------------------------------------------------- BASELINE.json full sizes
def _full_size_model(T, layers=2):
    torch.manual_seed(0)
    return mm.MultimodalEmotionModel(max_seq_len=T + 1, fusion_num_layers=layers, classifier_hidden_dim=512,
                                     fusion_dropout=0.0, classifier_dropout=0.0).cuda()


def test_cfg2_full_batch_is_sample_independent_and_matches_small_batches_bf16():
    """cfg2 size (B=4096, T=16, bf16): the oracle cannot run this in seconds, so check the size-independent
    properties the model has by construction (LayerNorm variant: no cross-sample coupling):
    a permutation of the batch permutes logits bit for bit, and slices evaluated on their own give the same logits
    (different GEMM tilings / CTA-pair vs single-CTA kernels) to bf16 tolerance."""
    B, T = 4096, 16
    model = _full_size_model(T).eval()
    model.compute_dtype = torch.bfloat16
    g = torch.Generator().manual_seed(5)
    video = torch.randn(B, T, 768, generator=g).cuda().bfloat16()
    audio = torch.randn(B, 1024, generator=g).cuda().bfloat16()
    lens = torch.randint(1, T + 1, (B,), generator=g)
    mask = (torch.arange(T)[None] >= lens[:, None]).cuda()
    with torch.no_grad():
        probs, logits, _ = model(video, audio, mask)
        perm = torch.randperm(B, generator=g).cuda()
        _, logits_p, _ = model(video[perm], audio[perm], mask[perm])
        assert torch.equal(logits_p, logits[perm])
        _, logits_s, _ = model(video[:96], audio[:96], mask[:96])
    err = float((logits_s - logits[:96]).abs().max() / logits.abs().max())
    assert err < 2e-2
    assert float((logits_s.argmax(1) == logits[:96].argmax(1)).float().mean()) > 0.97
    assert torch.allclose(probs.sum(1), torch.ones(B, device="cuda"), atol=1e-5)


def test_cfg2_full_size_training_step_gradient_is_mean_of_shard_gradients_bf16():
    """cfg2 size: the gradient of the mean loss over 4096 samples equals the average of the gradients over its four
    1024-sample shards (the identity the data-parallel all-reduce relies on), to bf16 accumulation noise."""
    B, T = 4096, 16
    model = _full_size_model(T).train()
    model.compute_dtype = torch.bfloat16
    g = torch.Generator().manual_seed(6)
    video = torch.randn(B, T, 768, generator=g).cuda().bfloat16()
    audio = torch.randn(B, 1024, generator=g).cuda().bfloat16()
    labels = torch.randint(0, 6, (B,), generator=g).cuda()
    crit = mm.FocalLoss(gamma=2.0, alpha=ALPHA.cuda())

    def grad_of(sl):
        model.zero_grad(set_to_none=True)
        _, logits, _ = model(video[sl], audio[sl])
        crit(logits, labels[sl]).backward()
        return model._engine.ctx.grads.clone()

    full = grad_of(slice(0, B))
    parts = sum(grad_of(slice(i * 1024, (i + 1) * 1024)) for i in range(4)) / 4
    rel_err = float((full - parts).norm() / full.norm())
    assert rel_err < 2e-2, rel_err


def test_cfg4_long_sequence_attention_weights_bf16():
    """cfg4 (train2 variant, T=256, return_attn=True): rows of every attention map sum to 1, masked keys get exactly
    zero weight, the audio row is the last row of the head-mean of the last layer, and bf16 agrees with the fp32
    path on the argmax of the audio row for almost every sample."""
    B, T = 16, 256
    model = _full_size_model(T).eval()
    g = torch.Generator().manual_seed(7)
    video = torch.randn(B, T, 768, generator=g).cuda()
    audio = torch.randn(B, 1024, generator=g).cuda()
    lens = torch.randint(32, T + 1, (B,), generator=g)
    mask = (torch.arange(T)[None] >= lens[:, None]).cuda()
    outs = {}
    for dt in (torch.float32, torch.bfloat16):
        model.compute_dtype = dt
        with torch.no_grad():
            probs, logits, attn = model(video.to(dt), audio.to(dt), mask, return_attn=True)
        outs[dt] = (logits.float(), attn)
        last = attn["layers"][-1]                       # (B, H, S, S)
        assert torch.allclose(last.sum(-1), torch.ones_like(last.sum(-1)), atol=1e-4)
        full = torch.cat([mask, torch.zeros(B, 1, dtype=torch.bool, device="cuda")], 1)
        assert float(last.masked_select(full.view(B, 1, 1, T + 1).expand_as(last)).abs().max()) == 0.0
        assert torch.allclose(attn["last_mean"], last.mean(1), atol=1e-6)
        assert torch.equal(attn["audio_row"], attn["last_mean"][:, -1, :])
    err = float((outs[torch.bfloat16][0] - outs[torch.float32][0]).abs().max() / outs[torch.float32][0].abs().max())
    assert err < 5e-2
    same = (outs[torch.bfloat16][1]["audio_row"].argmax(1) == outs[torch.float32][1]["audio_row"].argmax(1)).float().mean()
    assert float(same) >= 0.8


@pytest.mark.parametrize("B", [1, 8192])
def test_cfg5_inference_batch_1_and_8192_bf16(B):
    """cfg5: inference-only forward at batch 1 (serving latency shape, T=5 as in routers/infer.py:9) and batch 8192."""
    T = 5 if B == 1 else 16
    model = _full_size_model(T).eval()
    model.compute_dtype = torch.bfloat16
    g = torch.Generator().manual_seed(8)
    video = torch.randn(B, T, 768, generator=g).cuda().bfloat16()
    audio = torch.randn(B, 1024, generator=g).cuda().bfloat16()
    with torch.no_grad():
        probs, logits, attn = model(video, audio)
        assert attn is None and probs.shape == (B, 6) and bool(torch.isfinite(logits).all())
        # the same sample inside a batch of copies gives the same logits
        probs2, logits2, _ = model(video[:1].expand(3, -1, -1).contiguous(), audio[:1].expand(3, -1).contiguous())
    assert float((logits2 - logits[:1]).abs().max()) < 2e-2 * float(logits.abs().max())


# ---------------------------------------------------------------- Integrated Gradients (train2.py:776-866)
def _ig_golden():
    return np.load(os.path.join(GOLD, "ig_v2_b8_t5_mask.npz"))


def test_compute_attributions_matches_reference_golden_fp32():
    """compute_attributions (engine forward/backward + the expand / reduce kernels) against attributions of the
    unmodified reference model class (tests/golden/make_golden_ig.py); fp32 tolerance 1e-4 of the largest value."""
    ig = _ig_golden()
    _, model, P, video, audio, mask, _ = build("v2_b8_t5_mask", "v2")
    model.train()
    before = {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    av, aa = mm.compute_attributions(model, video, audio, mask=mask, n_steps=int(ig["n_steps"]))
    assert not model.training                                   # left in eval mode like the reference
    assert av.shape == video.shape and aa.shape == audio.shape and av.is_cuda and av.dtype == torch.float32
    scale = float(np.abs(ig["attr_video"]).max())
    assert float(np.abs(av.cpu().numpy() - ig["attr_video"]).max()) < 1e-4 * scale
    assert float(np.abs(aa.cpu().numpy() - ig["attr_audio"]).max()) < 1e-4 * scale
    assert float(av[mask].abs().max()) == 0.0                   # padded frames: exactly zero
    for k, p in model.named_parameters():                       # parameter gradients are not touched
        assert (p.grad is None) == (before[k] is None) and (p.grad is None or torch.equal(p.grad, before[k]))
    vi, ai = mm.aggregate_importances(av, aa)
    assert vi.shape == (video.shape[0], video.shape[2]) and ai.shape == audio.shape


def test_compute_attributions_targets_baselines_and_chunking():
    from oracle import ig_oracle
    _, model, P, video, audio, mask, labels = build("v2_b8_t5_mask", "v2")
    P64 = {k: (t.double() if t.is_floating_point() else t) for k, t in P.items()}
    fn = lambda v, a, mk: O.model_forward_v2(P64, v, a, mk)[1]  # noqa: E731
    g = torch.Generator().manual_seed(5)
    bv = (0.1 * torch.randn(video.shape, generator=g)).cuda()
    ba = (0.1 * torch.randn(audio.shape, generator=g)).cuda()
    n = 7
    av, aa = mm.compute_attributions(model, video, audio, mask=mask, target=labels, n_steps=n, baseline=(bv, ba))
    rv, ra = ig_oracle.integrated_gradients(fn, (video.cpu().double(), audio.cpu().double()),
                                            (bv.cpu().double(), ba.cpu().double()), mask.cpu(), labels.cpu(), n)
    # fp32 GEMMs (K up to 2048) under 7 summed gradient evaluations: measured 1.0e-4 of the largest attribution
    scale = float(max(rv.abs().max(), ra.abs().max()))
    assert float((av.cpu().double() - rv).abs().max()) < 2e-4 * scale
    assert float((aa.cpu().double() - ra).abs().max()) < 2e-4 * scale
    # Captum's internal_batch_size: the steps in chunks give the same sums
    cv, ca = mm.compute_attributions(model, video, audio, mask=mask, target=labels, n_steps=n, baseline=(bv, ba),
                                     internal_batch_size=3 * video.shape[0])
    # (not bitwise: another batch size changes the GEMM tiling / split-K summation order; both are within 1e-4 of exact)
    assert float((cv - av).abs().max()) < 2e-4 * scale and float((ca - aa).abs().max()) < 2e-4 * scale
    # an int target is broadcast
    iv, _ = mm.compute_attributions(model, video, audio, mask=mask, target=2, n_steps=3)
    tv, _ = mm.compute_attributions(model, video, audio, mask=mask, target=torch.full((video.shape[0],), 2), n_steps=3)
    assert torch.equal(iv, tv)
    with pytest.raises(ValueError):
        mm.compute_attributions(model, video, audio, mask=mask, baseline="mean")


def test_compute_attributions_bf16_and_completeness():
    """Completeness against the engine's own logits, and bf16 compute within 5e-2 of the fp32 attributions in l2 norm
    (measured 1.4e-2: the bf16 gradient noise floor of tests/test_oracle_golden.py::test_bf16_storage_noise_floor,
    averaged over the integration steps)."""
    _, model, P, video, audio, mask, _ = build("v2_b8_t5_mask", "v2")
    n = 64
    av, aa = mm.compute_attributions(model, video, audio, mask=mask, n_steps=n)
    model.eval()
    with torch.no_grad():
        _, lx, _ = model(video, audio, mask=mask)
        _, l0, _ = model(torch.zeros_like(video), torch.zeros_like(audio), mask=mask)
    tgt = lx.argmax(1, keepdim=True)
    delta = (lx - l0).gather(1, tgt).squeeze(1)
    gap = float((av.flatten(1).sum(1) + aa.sum(1) - delta).abs().max())
    assert gap < 0.06, gap
    model.compute_dtype = torch.bfloat16
    bv_, ba_ = mm.compute_attributions(model, video, audio, mask=mask, target=tgt.squeeze(1), n_steps=n)
    rel = float((bv_ - av).norm() / av.norm())
    assert rel < 5e-2, rel


def test_graphed_inference_equals_eager_and_tracks_weight_updates():
    """cfg5 serving shape (1 clip, 5 chunks) and a small batch: graph replay == eager forward, for new inputs, with and
    without a mask, in fp32 and bf16 compute, and after the weights change."""
    for cdt in (None, torch.bfloat16):
        g, model, P, video, audio, mask, labels = build("v2_b8_t5_mask", "v2")
        model.eval()
        model.compute_dtype = cdt
        run = mm.GraphedInference(model, batch=video.shape[0], frames=video.shape[1])
        one = mm.GraphedInference(model, batch=1, frames=video.shape[1])
        with torch.no_grad():
            for v, a, m in ((video, audio, mask), (video.flip(0), audio * 0.5, None)):
                pe, le, _ = model(v, a, mask=m)
                pg, lg = run(v, a, m)
                assert torch.equal(pe, pg) and torch.equal(le, lg)
            pe, le, _ = model(video[:1], audio[:1], mask=mask[:1])
            pg, lg = one(video[:1], audio[:1], mask[:1])
            assert torch.equal(le, lg)
            for p in model.parameters():
                p.mul_(0.9)
            pe, le, _ = model(video, audio, mask=mask)
            pg, lg = run(video, audio, mask)
            assert torch.equal(le, lg)
        with pytest.raises(mm.MmerError):
            run(video[:2], audio[:2])


def test_reference_epoch_loop_runs_on_the_device_pieces():
    """The body of the reference's train_model (train2.py:523-647) with the device-side pieces swapped in: DeviceLoader
    for DataLoader + collate_fn + .to(device), FusedTrainStep for zero_grad/forward/criterion/backward/clip/step,
    EvalAccumulator for the .item()/.cpu() bookkeeping; ReduceLROnPlateau, best-state copy and the checkpoint stay as
    they are in the reference.  The trained weights then load into the reference architecture on stock torch.nn
    (oracle/eager_torch.py) and give the same logits."""
    from oracle import eager_torch as E
    gen = torch.Generator().manual_seed(11)
    n, T = 192, 6
    labels = torch.randint(0, 6, (n,), generator=gen).tolist()
    lens = torch.randint(2, T + 1, (n,), generator=gen).tolist()
    proto_v, proto_a = torch.randn(6, 768, generator=gen), torch.randn(6, 1024, generator=gen)   # class-dependent means
    videos = [proto_v[y] * 0.5 + torch.randn(t, 768, generator=gen) for y, t in zip(labels, lens)]
    audios = [proto_a[y] * 0.5 + torch.randn(1024, generator=gen) for y in labels]
    ds = mm.DeviceFeatureSet(videos, audios, labels)
    train_idx, val_idx, _ = mm.data.stratified_split(labels)
    class_weights = mm.data.balanced_class_weights([labels[i] for i in train_idx])
    torch.manual_seed(0)
    model = mm.MultimodalEmotionModel(max_seq_len=T + 1, fusion_num_layers=2, classifier_hidden_dim=512,
                                      fusion_dropout=0.1, classifier_dropout=0.1).cuda()
    step = mm.FusedTrainStep(model, lr=3e-4, weight_decay=1e-4, loss="wce", alpha=class_weights, clip_grad_norm=1.0,
                             compute_dtype=torch.float32)                               # train2.py:523-526,576
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(step.opt, mode="min", factor=0.3, patience=0)
    criterion = mm.WeightedCrossEntropyLoss(class_weights.cuda())
    history, best, best_state = [], float("inf"), None
    for epoch in range(4):
        model.train()
        train_loss = torch.zeros(1, device="cuda")
        loader = ds.loader(train_idx, batch_size=32, shuffle=True)
        for v, a, y, m in loader:
            loss, _ = step.step(v, a, m, y)
            train_loss += loss
        model.eval()
        acc = mm.EvalAccumulator(6)
        with torch.no_grad():
            for v, a, y, m in ds.loader(val_idx, batch_size=32):
                probs, logits, _ = model(v, a, mask=m)
                acc.update(probs, y, criterion(logits, y))
        out = acc.result()
        history.append((float(train_loss) / len(loader), out["avg_loss"], out["accuracy"], step.lr))
        scheduler.step(out["avg_loss"] if epoch != 2 else 1e9)        # force one plateau: lr must drop by 0.3
        if out["avg_loss"] < best:
            best, best_state = out["avg_loss"], {k: t.clone() for k, t in model.state_dict().items()}
    assert history[-1][0] < 0.7 * history[0][0], history                 # it learns the class prototypes
    assert abs(step.lr - 3e-4 * 0.3) < 1e-12 and step.opt.param_groups[0]["lr"] == step.lr
    assert out["total"] == len(val_idx) and 0.0 <= out["macro_f1"] <= 1.0 and history[-1][2] > 100.0 / 6
    # checkpoint interchange with the reference architecture (same state_dict keys and shapes)
    ref = E.EagerModel(max_seq_len=T + 1, fusion_num_layers=2, classifier_hidden_dim=512).cuda().eval()
    ref.load_state_dict(best_state, strict=True)
    model.load_state_dict(best_state, strict=True)
    model.eval()
    v, a, y, m = ds.collate(val_idx[:16])
    with torch.no_grad():
        _, lo, _ = model(v, a, mask=m)
        _, lr_ = ref(v, a, m)
    assert float((lo - lr_).abs().max()) < 1e-4 * max(1.0, float(lr_.abs().max()))


# ---------------------------------------------------------------- post-norm sub-layer tails: three kernel arrangements
def test_layernorm_tail_modes_agree_bf16():
    """engine.cu ln_mode: (0) GEMM -> a, add_ln(x, a); (1) the fully fused tcgen05 GEMM + LayerNorm kernel (gemm_ln.cu);
    (2, default) GEMM with the residual epilogue -> z, LayerNorm(z).  Same model, same batch, same dropout seed: logits,
    loss and every parameter gradient must agree to bf16 rounding, with and without dropout (the three arrangements draw
    the SAME dropout mask: one counter hash of (seed, site, row * 512 + col))."""
    from mmer_b200 import _lib
    lib = _lib.load()
    B, T = 64, 16                                   # M = 1088 >= 256: the fused arrangements are eligible
    g = torch.Generator().manual_seed(3)
    video = torch.randn(B, T, 768, generator=g).cuda().bfloat16()
    audio = torch.randn(B, 1024, generator=g).cuda().bfloat16()
    labels = torch.randint(0, 6, (B,), generator=g).cuda()
    lens = torch.randint(1, T + 1, (B,), generator=g)
    mask = (torch.arange(T)[None] >= lens[:, None]).cuda()
    try:
        for p_drop in (0.0, 0.2):
            out = {}
            for knob, name in ((1, "unfused"), (2, "fused kernel"), (0, "residual epilogue")):
                lib.mmer_debug_set(_lib.DEBUG_NO_LN_FUSE, knob)
                torch.manual_seed(0)
                model = mm.MultimodalEmotionModel(max_seq_len=T + 1, classifier_hidden_dim=512, fusion_dropout=p_drop,
                                                  classifier_dropout=0.0).cuda().train()
                model.compute_dtype = torch.bfloat16
                model.__dict__["_base_seed"] = 1234          # same dropout stream for the three runs
                import mmer_b200.modules as M_
                M_._seed_counter = __import__("itertools").count(77)
                _, logits, _ = model(video, audio, mask=mask)
                loss = mm.FocalLoss(2.0, ALPHA.cuda())(logits, labels)
                loss.backward()
                out[name] = (logits.detach().float().clone(), float(loss), model._engine.ctx.grads.clone())
            ref = out["unfused"]
            for name in ("fused kernel", "residual epilogue"):
                lg, ls, gr = out[name]
                assert float((lg - ref[0]).abs().max()) < 3e-2 * float(ref[0].abs().max()), (name, p_drop)
                assert abs(ls - ref[1]) < 2e-2 * abs(ref[1]), (name, p_drop)
                cos = float((gr * ref[2]).sum() / (gr.norm() * ref[2].norm()))
                assert cos > 0.99 and abs(float(gr.norm() / ref[2].norm()) - 1) < 3e-2, (name, p_drop, cos)
    finally:
        lib.mmer_debug_set(_lib.DEBUG_NO_LN_FUSE, 0)


# ---------------------------------------------------------------- batch-1 serving forward (one cluster kernel)
@pytest.mark.parametrize("T,masked", [(5, False), (5, True), (1, False), (7, True), (8, True), (15, True)])
def test_serving_forward_single_launch_matches_module_and_stock_reference(T, masked):
    """mmer_serve_forward (csrc/serve.cu) on the served shape (back-end/app/libs/inference.py:494-495): same logits as
    the layer-by-layer bf16 engine and as the stock fp32 modules on bf16-rounded weights, to bf16 tolerance; argmax equal;
    parameter changes between calls are picked up (the graph re-casts the shadow)."""
    from oracle import eager_torch as E
    torch.manual_seed(T)
    model = mm.MultimodalEmotionModel(max_seq_len=T + 1, fusion_num_layers=2, classifier_hidden_dim=512).cuda().eval()
    ref = E.EagerModel(max_seq_len=T + 1, fusion_num_layers=2, classifier_hidden_dim=512).cuda().eval()
    last = "classifier.net.8.weight"
    ref.load_state_dict({k: (v.bfloat16().float() if (v.dim() == 2 and k != last) else v) for k, v in model.state_dict().items()})
    g = torch.Generator().manual_seed(100 + T)
    video = torch.randn(1, T, 768, generator=g).bfloat16().cuda()
    audio = torch.randn(1, 1024, generator=g).bfloat16().cuda()
    mask = None
    if masked:
        mask = torch.zeros(1, T, dtype=torch.bool, device="cuda")
        mask[0, T - max(1, T // 3):] = True
    run = mm.ServingForward(model, frames=T)
    probs, logits = run(video, audio, mask)
    model.compute_dtype = torch.bfloat16
    with torch.no_grad():
        p_eng, l_eng, _ = model(video, audio, mask=mask)
        p_ref, l_ref = ref(video.float(), audio.float(), mask)
    scale = float(l_ref.abs().max())
    assert float((logits - l_ref).abs().max()) < 2e-2 * scale, (logits, l_ref)
    assert float((logits - l_eng).abs().max()) < 3e-2 * scale
    assert float((probs - p_ref).abs().max()) < 2e-2
    assert abs(float(probs.sum()) - 1.0) < 1e-5
    top2 = l_ref.topk(2).values[0]
    if float(top2[0] - top2[1]) > 3e-2 * scale:
        assert int(logits.argmax()) == int(l_ref.argmax())
    # same call again: identical bits; after an in-place parameter change: follows the module
    # (S <= 8 sums split-K partial products with fp32 reductions in L2: the order varies, the last bits may)
    probs2, logits2 = run(video, audio, mask)
    if T + 1 > 8:
        assert torch.equal(logits2, logits)
    else:
        assert float((logits2 - logits).abs().max()) < 5e-3 * scale
    with torch.no_grad():
        for p_ in model.parameters():
            p_.mul_(1.05)
        _, l_new, _ = model(video, audio, mask=mask)
    _, logits3 = run(video, audio, mask)
    assert float((logits3 - l_new).abs().max()) < 3e-2 * float(l_new.abs().max())
    assert float((logits3 - logits).abs().max()) > 1e-4
    with pytest.raises(mm.MmerError):
        run(video[:, :-1] if T > 1 else torch.zeros(1, 2, 768, device="cuda"), audio, None)
    # zero-copy form: the request written into the server's own buffers, outputs read in place
    run.video.copy_(video)
    run.audio.copy_(audio)
    run.mask.copy_(mask) if mask is not None else run.mask.zero_()
    p4, l4 = run.replay()
    assert l4.data_ptr() == run.logits.data_ptr()
    assert float((l4 - logits3).abs().max()) < 5e-3 * float(logits3.abs().max())


@pytest.mark.parametrize("T", [1, 5, 7])
def test_serving_forward_variants_agree(T):
    """S <= 8 runs csrc/serve_small.cu (head-local attention, split-K sums in L2); MMER_DEBUG_SERVE_GLOBAL = 1 forces
    csrc/serve.cu (output-feature split, L2 exchange; bit-reproducible run to run), 2 csrc/serve_dsmem.cu (shared-memory
    broadcast).  Same roundings at the same places, different summation orders: the logits agree to a few bf16 flips."""
    from mmer_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(40 + T)
    model = mm.MultimodalEmotionModel(max_seq_len=T + 1, fusion_num_layers=2, classifier_hidden_dim=512).cuda().eval()
    g = torch.Generator().manual_seed(7 + T)
    video = torch.randn(1, T, 768, generator=g).bfloat16().cuda()
    audio = torch.randn(1, 1024, generator=g).bfloat16().cuda()
    mask = torch.zeros(1, T, dtype=torch.bool, device="cuda")
    if T > 2:
        mask[0, T - 2:] = True
    run = mm.ServingForward(model, frames=T, use_graph=False)
    out = {}
    try:
        for knob in (0, 1, 2):
            lib.mmer_debug_set(_lib.DEBUG_SERVE_GLOBAL, knob)
            _, l = run(video, audio, mask)
            out[knob] = l.clone()
            if knob == 1:
                _, l2 = run(video, audio, mask)
                assert torch.equal(l2, out[1])
    finally:
        lib.mmer_debug_set(_lib.DEBUG_SERVE_GLOBAL, 0)
    scale = float(out[1].abs().max())
    assert float((out[0] - out[1]).abs().max()) < 5e-3 * scale, (out[0], out[1])
    assert float((out[2] - out[1]).abs().max()) < 5e-3 * scale


# ----------------------------------------------------------------------------- use_layernorm=False (train2.py:96,208)
def _nolayernorm_modules():
    """The two train2.py sub-modules built with use_layernorm=False, loaded like tests/golden/make_golden_nolayernorm.py."""
    g = np.load(os.path.join(GOLD, "v2_nolayernorm_b8_t5_mask.npz"))
    B, T, HID = int(g["B"]), int(g["T"]), int(g["hidden"])
    params = detgen.make_params("v2", max_seq_len=T + 1, hidden=HID)
    fusion = mm.CrossModalFusion(num_layers=2, dropout=0.0, max_seq_len=T + 1, use_layernorm=False)
    head = mm.EmotionClassifier(input_dim=512, hidden_dim=HID, dropout=0.0, use_layernorm=False)
    assert isinstance(fusion.norm_video, torch.nn.Identity) and isinstance(fusion.out_norm, torch.nn.Identity)
    assert isinstance(head.net[1], torch.nn.BatchNorm1d) and isinstance(head.net[5], torch.nn.BatchNorm1d)
    fsd = {k[len("fusion."):]: torch.from_numpy(v) for k, v in params.items()
           if k.startswith("fusion.") and "norm_video" not in k and "norm_audio" not in k and "out_norm" not in k}
    fusion.load_state_dict(fsd, strict=True)
    hsd = {k[len("classifier."):]: torch.from_numpy(v) for k, v in params.items() if k.startswith("classifier.")}
    for i in (1, 5):
        hsd[f"net.{i}.running_mean"] = torch.from_numpy(g[f"bn_init/net.{i}.running_mean"])
        hsd[f"net.{i}.running_var"] = torch.from_numpy(g[f"bn_init/net.{i}.running_var"])
        hsd[f"net.{i}.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    head.load_state_dict(hsd, strict=True)
    v, a, m, y = detgen.make_batch(B, T, tag="nolayernorm")
    fused_in = torch.from_numpy(detgen.det_array((B, 512), "nolayernorm/fused_in", 1.0)).cuda()
    wr = torch.from_numpy(detgen.det_array((B, 512), "nolayernorm/wr", 1.0)).cuda()
    return (g, fusion.cuda(), head.cuda(), hsd, torch.from_numpy(v).cuda(), torch.from_numpy(a).cuda(),
            torch.from_numpy(m).cuda(), torch.from_numpy(y).cuda(), fused_in, wr)


def _check_golden_grads(g, prefix, named, rtol=2e-4):
    n = 0
    for k, p in named:
        got = p.grad.detach().cpu().numpy()
        if f"{prefix}/gradfull/{k}" in g:
            want = g[f"{prefix}/gradfull/{k}"]
        else:
            want, got = g[f"{prefix}/gradproj/{k}"], detgen.project(got, k)
        # absolute floor: the bias of a Linear that feeds BatchNorm has a mathematically zero gradient (1e-9 of noise)
        assert float(np.abs(got - want).max()) < rtol * float(np.abs(want).max()) + 2e-7, (prefix, k)
        n += 1
    return n


def test_use_layernorm_false_variants_match_reference_golden_fp32():
    """CrossModalFusion with nn.Identity norms and EmotionClassifier with nn.BatchNorm1d (train2.py:104-105,121,215),
    each on its own and chained, against outputs, gradients and BatchNorm statistics of the unmodified reference."""
    g, fusion, head, hsd, video, audio, mask, labels, fused_in, wr = _nolayernorm_modules()
    # ---- fusion alone
    fusion.eval()
    with torch.no_grad():
        np.testing.assert_allclose(fusion(video, audio, mask=mask)[0].cpu().numpy(), g["fusion/eval_fused"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(fusion(video, audio)[0].cpu().numpy(), g["fusion/eval_fused_nomask"], rtol=1e-4, atol=1e-5)
    fusion.train()
    vg, ag = video.clone().requires_grad_(True), audio.clone().requires_grad_(True)
    fused = fusion(vg, ag, mask=mask)[0]
    np.testing.assert_allclose(fused.detach().cpu().numpy(), g["fusion/train_fused"], rtol=1e-4, atol=1e-5)
    fusion.zero_grad()
    (fused * wr).sum().backward()
    assert _check_golden_grads(g, "fusion", fusion.named_parameters()) == len(list(fusion.parameters()))
    np.testing.assert_allclose(vg.grad.cpu().numpy(), g["fusion/grad_video"], rtol=1e-3, atol=2e-6)
    np.testing.assert_allclose(ag.grad.cpu().numpy(), g["fusion/grad_audio"], rtol=1e-3, atol=2e-6)
    # ---- classifier alone
    head.eval()
    with torch.no_grad():
        np.testing.assert_allclose(head(fused_in).cpu().numpy(), g["head/eval_logits"], rtol=1e-4, atol=1e-5)
    head.train()
    fg = fused_in.clone().requires_grad_(True)
    logits = head(fg)
    np.testing.assert_allclose(logits.detach().cpu().numpy(), g["head/train_logits"], rtol=1e-4, atol=1e-5)
    head.zero_grad()
    loss = mm.WeightedCrossEntropyLoss(ALPHA.cuda())(logits, labels)
    assert abs(float(loss) - float(g["head/loss"])) < 1e-5
    loss.backward()
    assert _check_golden_grads(g, "head", head.named_parameters()) == len(list(head.parameters()))
    np.testing.assert_allclose(fg.grad.cpu().numpy(), g["head/grad_fused"], rtol=1e-3, atol=1e-7)
    sd = head.state_dict()
    for i in (1, 5):
        np.testing.assert_allclose(sd[f"net.{i}.running_mean"].cpu().numpy(), g[f"head/bn_after_fwd/net.{i}.running_mean"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(sd[f"net.{i}.running_var"].cpu().numpy(), g[f"head/bn_after_fwd/net.{i}.running_var"], rtol=1e-5, atol=1e-6)
        assert int(sd[f"net.{i}.num_batches_tracked"]) == int(g[f"head/bn_after_fwd/net.{i}.num_batches_tracked"]) == 1
    # ---- chained inside MultimodalEmotionModel (one engine, both flags): sub-modules swapped in like a user would
    model = mm.MultimodalEmotionModel(max_seq_len=int(g["T"]) + 1, fusion_num_layers=2, classifier_hidden_dim=int(g["hidden"]),
                                      fusion_dropout=0.0, classifier_dropout=0.0)
    head.load_state_dict(hsd, strict=True)
    model.fusion, model.classifier = fusion, head
    model.cuda().train()
    vg, ag = video.clone().requires_grad_(True), audio.clone().requires_grad_(True)
    probs, logits, _ = model(vg, ag, mask=mask)
    np.testing.assert_allclose(logits.detach().cpu().numpy(), g["chain/train_logits"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(probs.detach().cpu().numpy(), g["chain/train_probs"], rtol=1e-4, atol=1e-6)
    model.zero_grad()
    loss = mm.WeightedCrossEntropyLoss(ALPHA.cuda())(logits, labels)
    assert abs(float(loss) - float(g["chain/loss"])) < 1e-5
    loss.backward()
    assert _check_golden_grads(g, "chain", model.named_parameters()) == len(list(model.parameters()))
    np.testing.assert_allclose(vg.grad.cpu().numpy(), g["chain/grad_video"], rtol=1e-3, atol=2e-7)
    model.eval()
    with torch.no_grad():
        np.testing.assert_allclose(model(video, audio, mask=mask)[1].cpu().numpy(), g["chain/eval_logits_after"], rtol=1e-4, atol=1e-5)


def test_use_layernorm_false_variants_bf16_close_to_fp32():
    g, fusion, head, hsd, video, audio, mask, labels, fused_in, wr = _nolayernorm_modules()
    fusion.eval()
    head.eval()
    with torch.no_grad():
        f32 = fusion(video, audio, mask=mask)[0]
        l32 = head(fused_in)
        fusion.compute_dtype = torch.bfloat16
        head.compute_dtype = torch.bfloat16
        f16 = fusion(video, audio, mask=mask)[0]
        l16 = head(fused_in)
    assert float((f16.float() - f32).abs().max() / f32.abs().max()) < 2e-2
    assert float((l16.float() - l32).abs().max() / l32.abs().max()) < 2e-2
