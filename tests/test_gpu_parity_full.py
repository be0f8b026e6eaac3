"""Model-level GPU parity AT THE BENCHMARKED SIZES (BASELINE.json configs 2, 4, 5) against the reference architecture
on stock torch.nn modules in fp32 on the same GPU (oracle/eager_torch.py -- pinned to the reference's golden vectors by
tests/test_oracle_golden.py::test_stock_torch_restatement_matches_reference_golden), plus the bf16 gradient against the
storage-rounded CPU oracle and regression tests for the round-1 advisor findings.

Tolerances (north_star): logits / loss within 1e-4 relative in fp32 and 2e-2 in bf16, argmax identical outside numerical
ties, gradient norms within 3e-2.  bf16 runs are compared with the fp32 reference evaluated on the SAME bf16-rounded GEMM
weights and inputs (SURVEY 8c oracle hygiene).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import detgen  # noqa: E402
import mmer_b200 as mm  # noqa: E402
from oracle import eager_torch as E  # noqa: E402
from oracle import fusion_oracle as O  # noqa: E402

ALPHA = torch.tensor([1, 1, 1, 1, 1.2, 1.2])
LAST = "classifier.net.8.weight"          # the N=6 layer runs in fp32 on CUDA cores: its weight is not rounded


def _pair(T, *, bf16, seed=0):
    """(product model, stock-torch fp32 reference holding the weights the product computes with)."""
    torch.manual_seed(seed)
    model = mm.MultimodalEmotionModel(max_seq_len=T + 1, fusion_num_layers=2, classifier_hidden_dim=512,
                                      fusion_dropout=0.0, classifier_dropout=0.0).cuda()
    ref = E.EagerModel(max_seq_len=T + 1, fusion_num_layers=2, classifier_hidden_dim=512, fusion_dropout=0.0,
                       classifier_dropout=0.0).cuda()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    if bf16:
        sd = {k: (v.bfloat16().float() if (v.dim() == 2 and k != LAST) else v) for k, v in sd.items()}
        model.compute_dtype = torch.bfloat16
    ref.load_state_dict(sd, strict=True)
    return model, ref


def _batch(B, T, seed, masked):
    g = torch.Generator().manual_seed(seed)
    video = torch.randn(B, T, 768, generator=g).bfloat16().float().cuda()     # bf16-representable: same inputs both sides
    audio = torch.randn(B, 1024, generator=g).bfloat16().float().cuda()
    labels = torch.randint(0, 6, (B,), generator=g).cuda()
    mask = None
    if masked:
        lens = torch.randint(max(1, T // 8), T + 1, (B,), generator=g)
        mask = (torch.arange(T)[None] >= lens[:, None]).cuda()
    return video, audio, mask, labels


def _argmax_equal_outside_ties(got, ref, margin):
    top2 = ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > margin
    assert float(clear.float().mean()) > 0.5
    assert bool((got.argmax(1)[clear] == ref.argmax(1)[clear]).all())


@pytest.mark.parametrize("masked", [False, True])
def test_cfg2_full_size_training_step_matches_stock_fp32_reference_bf16(masked):
    """cfg2 exactly as benchmarked: B = 4096, T = 16, bf16, FocalLoss(gamma 2, alpha), train mode (dropout 0 for
    parity).  Logits, loss, argmax, every parameter gradient and the input gradients against stock torch fp32."""
    B, T = 4096, 16
    model, ref = _pair(T, bf16=True)
    model.train()
    ref.train()
    video, audio, mask, labels = _batch(B, T, 11, masked)
    alpha = ALPHA.cuda()
    vr, ar = video.clone().requires_grad_(True), audio.clone().requires_grad_(True)
    _, lref = ref(vr, ar, mask)
    loss_ref = E.focal_loss(lref, labels, 2.0, alpha)
    loss_ref.backward()
    vg, ag = video.bfloat16().requires_grad_(True), audio.bfloat16().requires_grad_(True)
    probs, logits, _ = model(vg, ag, mask=mask)
    loss = mm.FocalLoss(2.0, alpha)(logits, labels)
    loss.backward()
    scale = float(lref.abs().max())
    assert float((logits.detach() - lref.detach()).abs().max()) < 2e-2 * scale
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * float(loss_ref)
    _argmax_equal_outside_ties(logits.detach(), lref.detach(), 2e-2 * scale)
    ref_grads = dict(ref.named_parameters())
    got_all, ref_all = [], []
    for k, p in model.named_parameters():
        gr = ref_grads[k].grad
        if float(gr.norm()) < 1e-7:
            continue
        got = p.grad
        assert abs(float(got.norm() / gr.norm()) - 1.0) < 3e-2, (k, float(got.norm()), float(gr.norm()))
        cos = float((got * gr).sum() / (got.norm() * gr.norm()))
        assert cos > 0.98, (k, cos)
        got_all.append(got.flatten())
        ref_all.append(gr.flatten())
    got, gr = torch.cat(got_all), torch.cat(ref_all)
    err = float((got - gr).norm() / gr.norm())
    cos = float((got * gr).sum() / (got.norm() * gr.norm()))
    print(f"cfg2 masked={masked}: whole-gradient rel l2 {err:.4f} cosine {cos:.5f}")
    assert err < 0.12 and cos > 0.992          # bf16-storage noise floor, see test_bf16_gradient_vs_storage_rounded_oracle
    gv = vg.grad.float()
    assert abs(float(gv.norm() / vr.grad.norm()) - 1.0) < 3e-2
    assert float((gv * vr.grad).sum() / (gv.norm() * vr.grad.norm())) > 0.99
    if mask is not None:
        assert float(gv[mask].abs().max()) == 0.0 and float(vr.grad[mask].abs().max()) == 0.0


def test_cfg2_full_size_fp32_mode_matches_stock_fp32_reference():
    """The fp32 parity mode at the benchmarked size: 1e-4 on logits and loss against the stock fp32 modules.  Gradients
    are sums over 69,632 token rows, where two fp32 evaluations differ by their summation order (and a handful of ReLU
    pre-activations that round to the other side of zero): the truth is the stock model in FLOAT64 on the same GPU,
    and the product must be as close to it as the stock fp32 evaluation is (x3), and within 3e-3 per tensor."""
    B, T = 4096, 16
    model, ref = _pair(T, bf16=False)
    model.train()
    ref.train()
    video, audio, mask, labels = _batch(B, T, 12, True)
    alpha = ALPHA.cuda()
    _, lref = ref(video, audio, mask)
    loss_ref = E.focal_loss(lref, labels, 2.0, alpha)
    loss_ref.backward()
    ref64 = E.EagerModel(max_seq_len=T + 1, fusion_num_layers=2, classifier_hidden_dim=512, fusion_dropout=0.0,
                         classifier_dropout=0.0).cuda().double().train()
    ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    _, l64 = ref64(video.double(), audio.double(), mask)
    E.focal_loss(l64, labels, 2.0, alpha.double()).backward()
    probs, logits, _ = model(video, audio, mask=mask)
    loss = mm.FocalLoss(2.0, alpha)(logits, labels)
    loss.backward()
    scale = max(float(lref.detach().abs().max()), 1.0)
    assert float((logits.detach() - lref.detach()).abs().max()) < 1e-4 * scale
    assert float((logits.detach().double() - l64.detach()).abs().max()) < 1e-4 * scale
    assert abs(float(loss) - float(loss_ref)) < 1e-4 * float(loss_ref)
    _argmax_equal_outside_ties(logits.detach(), lref.detach(), 1e-4 * scale)
    g32, g64 = dict(ref.named_parameters()), dict(ref64.named_parameters())
    worst = (0.0, 0.0, "")
    for k, p in model.named_parameters():
        truth = g64[k].grad
        if float(truth.norm()) < 1e-7:
            continue
        ours = float((p.grad.double() - truth).norm() / truth.norm())
        stock = float((g32[k].grad.double() - truth).norm() / truth.norm())
        worst = max(worst, (ours, stock, k))
        assert ours < 3e-3 and ours < 3 * stock + 1e-4, (k, ours, stock)
    print(f"cfg2 fp32: worst gradient error vs float64 truth {worst[0]:.2e} (stock fp32: {worst[1]:.2e}) at {worst[2]}")


def _attention_of_stock_model(ref, video, audio, mask):
    """re-invoke every layer's self_attn with need_weights=True (the A9 oracle)."""
    store = []

    def hook(mod, args, kwargs, out):
        kw = dict(kwargs)
        kw["need_weights"] = True
        kw["average_attn_weights"] = False
        _, w = torch.nn.MultiheadAttention.forward(mod, *args, **kw)
        store.append(w.detach())

    hs = [l.self_attn.register_forward_hook(hook, with_kwargs=True) for l in ref.fusion.transformer.layers]
    with torch.no_grad():
        _, logits = ref(video, audio, mask)
    for h in hs:
        h.remove()
    return logits, torch.stack(store)          # (L, B, H, S, S)


@pytest.mark.parametrize("bf16", [False, True])
def test_cfg4_long_sequence_with_attention_weights_matches_stock_reference(bf16):
    """cfg4: T = 256 (S = 257), padding mask, return_attn=True, B = 64: logits and EVERY attention map against the
    need_weights=True hook on the stock modules; the audio-query row's argmax identical outside ties."""
    B, T = 64, 256
    model, ref = _pair(T, bf16=bf16)
    model.eval()
    ref.eval()
    video, audio, mask, _ = _batch(B, T, 13, True)
    lref, aref = _attention_of_stock_model(ref, video, audio, mask)
    dt = torch.bfloat16 if bf16 else torch.float32
    with torch.no_grad():
        probs, logits, attn = model(video.to(dt), audio.to(dt), mask, return_attn=True)
    tol = 2e-2 if bf16 else 1e-4
    scale = max(float(lref.abs().max()), 1.0 if not bf16 else 0.0)
    assert float((logits - lref).abs().max()) < tol * scale
    _argmax_equal_outside_ties(logits, lref, tol * scale)
    layers = attn["layers"]
    assert layers.shape == aref.shape
    # probabilities live in [0, 1]: absolute tolerance, relative to the largest weight of the row for bf16
    assert float((layers - aref).abs().max()) < (2e-2 if bf16 else 1e-5) * max(float(aref.max()), 1e-3) + (0 if bf16 else 1e-6)
    full = torch.cat([mask, torch.zeros(B, 1, dtype=torch.bool, device="cuda")], 1)
    assert float(layers[-1].masked_select(full.view(B, 1, 1, T + 1).expand_as(layers[-1])).abs().max()) == 0.0
    ref_row = aref[-1].mean(1)[:, -1, :]
    top2 = ref_row.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > (2e-2 if bf16 else 1e-5) * top2[:, 0]
    assert bool((attn["audio_row"].argmax(1)[clear] == ref_row.argmax(1)[clear]).all())
    assert torch.allclose(attn["last_mean"], aref[-1].mean(1), atol=(2e-2 if bf16 else 1e-5) * float(aref.max()) + 1e-6)


def test_cfg4_training_step_gradients_match_stock_reference_bf16():
    """cfg4 train step (T = 256, masked): loss and gradient norms against stock fp32."""
    B, T = 32, 256
    model, ref = _pair(T, bf16=True)
    model.train()
    ref.train()
    video, audio, mask, labels = _batch(B, T, 14, True)
    alpha = ALPHA.cuda()
    _, lref = ref(video, audio, mask)
    loss_ref = E.focal_loss(lref, labels, 2.0, alpha)
    loss_ref.backward()
    probs, logits, _ = model(video.bfloat16(), audio.bfloat16(), mask=mask)
    loss = mm.FocalLoss(2.0, alpha)(logits, labels)
    loss.backward()
    assert float((logits.detach() - lref.detach()).abs().max()) < 2e-2 * float(lref.abs().max())
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * float(loss_ref)
    ref_grads = dict(ref.named_parameters())
    for k, p in model.named_parameters():
        gr = ref_grads[k].grad
        if float(gr.norm()) < 1e-7:
            continue
        assert abs(float(p.grad.norm() / gr.norm()) - 1.0) < 4e-2, k
        assert float((p.grad * gr).sum() / (p.grad.norm() * gr.norm())) > 0.97, k


@pytest.mark.parametrize("B,T", [(1, 5), (8192, 16)])
def test_cfg5_inference_matches_stock_reference_bf16(B, T):
    """cfg5: the served shape (1 clip, 5 chunks, routers/infer.py:9) and the throughput shape (8192 x 16), eval, bf16."""
    model, ref = _pair(T, bf16=True)
    model.eval()
    ref.eval()
    video, audio, mask, _ = _batch(B, T, 15, B > 1)
    with torch.no_grad():
        pref, lref = ref(video, audio, mask)
        probs, logits, attn = model(video.bfloat16(), audio.bfloat16(), mask)
    assert attn is None
    scale = float(lref.abs().max())
    assert float((logits - lref).abs().max()) < 2e-2 * scale
    assert float((probs - pref).abs().max()) < 2e-2
    if B > 1:
        _argmax_equal_outside_ties(logits, lref, 2e-2 * scale)
    else:
        assert int(logits.argmax()) == int(lref.argmax()) or float(lref.topk(2).values.diff().abs()) < 2e-2 * scale


def test_bf16_gradient_vs_storage_rounded_oracle():
    """The bf16 GPU gradient against BOTH CPU oracles: the exact one (fp64 on bf16-rounded weights) and the same oracle
    with bf16 rounding at every tensor boundary (O.storage_rounding).  The distance between the two oracles IS the
    noise floor of bf16 storage (ReLU masks of near-zero pre-activations flip); the GPU path -- which rounds at the
    same boundaries, possibly to the other side for individual elements -- must sit within that floor of BOTH, and be
    no further from the exact answer than the emulation is (x1.25)."""
    B, T = 16, 16
    P = {k: torch.from_numpy(np.asarray(v)) for k, v in detgen.make_params("v2", max_seq_len=T + 1, hidden=512).items()}
    v, a, m, y = detgen.make_batch(B, T, tag="bf16case")
    video, audio, mask, labels = (torch.from_numpy(x) for x in (v, a, m, y))
    rounded = {k: (t.bfloat16().float() if (t.dim() == 2 and k != LAST) else t) for k, t in P.items()}

    def oracle(emulate):
        leaf = {k: t.double().clone().requires_grad_(True) for k, t in O.trainable(rounded).items()}
        full = {k: (t.double() if t.is_floating_point() else t) for k, t in rounded.items()}
        full.update(leaf)
        vr, ar = video.bfloat16().double(), audio.bfloat16().double()
        if emulate:
            with O.storage_rounding(torch.bfloat16):
                _, logits, _, _ = O.model_forward_v2(full, vr, ar, mask)
                O.focal_loss(logits, labels, 2.0, ALPHA.double()).backward()
        else:
            _, logits, _, _ = O.model_forward_v2(full, vr, ar, mask)
            O.focal_loss(logits, labels, 2.0, ALPHA.double()).backward()
        return logits.detach(), {k: t.grad for k, t in leaf.items()}

    l_exact, g_exact = oracle(False)
    l_round, g_round = oracle(True)
    model = mm.MultimodalEmotionModel(max_seq_len=T + 1, classifier_hidden_dim=512, fusion_dropout=0.0,
                                      classifier_dropout=0.0)
    model.load_state_dict(P)
    model.cuda().train()
    model.compute_dtype = torch.bfloat16
    probs, logits, _ = model(video.cuda(), audio.cuda(), mask=mask.cuda())
    mm.FocalLoss(2.0, ALPHA.cuda())(logits, labels.cuda()).backward()
    keys = [k for k in g_exact if float(g_exact[k].norm()) > 1e-6]
    cat = lambda d: torch.cat([d[k].flatten() for k in keys])  # noqa: E731
    got = torch.cat([dict(model.named_parameters())[k].grad.cpu().double().flatten() for k in keys])
    ge, gr = cat(g_exact), cat(g_round)
    rel = lambda x, y_: float((x - y_).norm() / y_.norm())  # noqa: E731
    floor = rel(gr, ge)
    d_exact, d_round = rel(got, ge), rel(got, gr)
    print(f"bf16 gradient: oracle-vs-oracle floor {floor:.4f}; GPU vs exact {d_exact:.4f}; GPU vs storage-rounded {d_round:.4f}")
    assert 0.02 < floor < 0.12
    assert d_exact < 1.25 * floor + 0.01
    assert d_round < 1.25 * floor + 0.01      # measured: 0.065 vs a floor of 0.069 (closer to the emulation than the exact oracle is)
    assert d_round < 1.6 * floor + 0.01       # two independent realisations of the same flip noise: ~sqrt(2) x floor
    assert float((logits.detach().cpu().double() - l_round).abs().max()) < 2e-2 * float(l_exact.abs().max())


# ------------------------------------------------------------------------- round-1 advisor findings (regressions)
def test_bf16_shadow_follows_load_state_dict_after_fused_training():
    """after FusedTrainStep steps in bf16, load_state_dict / in-place parameter changes must be seen
    by every later bf16 forward (evaluation, attribution, graphs) and by the next training step."""
    B, T = 64, 8
    torch.manual_seed(3)
    model = mm.MultimodalEmotionModel(max_seq_len=T + 1, classifier_hidden_dim=512, fusion_dropout=0.0,
                                      classifier_dropout=0.0).cuda().train()
    best = {k: v.detach().clone() for k, v in model.state_dict().items()}
    step = mm.FusedTrainStep(model, lr=1e-2, loss="focal", alpha=ALPHA)
    video, audio, mask, labels = _batch(B, T, 21, True)
    vb, ab = video.bfloat16(), audio.bfloat16()
    for _ in range(3):
        step.step(vb, ab, mask, labels)
    model.load_state_dict(best)                     # train2.py:721: best weights restored before the test evaluation
    model.eval()
    model.compute_dtype = torch.bfloat16
    with torch.no_grad():
        _, logits, _ = model(vb, ab, mask)
    torch.manual_seed(3)
    fresh = mm.MultimodalEmotionModel(max_seq_len=T + 1, classifier_hidden_dim=512, fusion_dropout=0.0,
                                      classifier_dropout=0.0).cuda().eval()
    fresh.load_state_dict(best)
    fresh.compute_dtype = torch.bfloat16
    with torch.no_grad():
        _, want, _ = fresh(vb, ab, mask)
    assert torch.equal(logits, want)
    # graph captured right after an optimizer step must still re-cast the shadow on every replay
    model.train()
    step.step(vb, ab, mask, labels)
    run = mm.GraphedInference(model, batch=B, frames=T, input_dtype=torch.bfloat16)
    model.load_state_dict(best)
    _, got = run(vb, ab, mask)
    assert torch.equal(got, want)
    # and the next fused step computes its gradient at the loaded weights
    model.train()
    l_after, _ = step.step(vb, ab, mask, labels)
    fresh.train()
    _, lg, _ = fresh(vb, ab, mask)
    want_loss = mm.FocalLoss(2.0, ALPHA.cuda())(lg, labels)
    assert abs(float(l_after) - float(want_loss)) < 1e-6 * max(1.0, abs(float(want_loss)))


def test_fused_adam_state_dict_interchanges_with_torch_adam():
    """optimizer state must survive state_dict()/load_state_dict(), in torch.optim.Adam's format."""
    g, T = torch.Generator().manual_seed(4), 5

    def make():
        torch.manual_seed(9)
        return mm.MultimodalEmotionModel(max_seq_len=T + 1, classifier_hidden_dim=512, fusion_dropout=0.0,
                                         classifier_dropout=0.0).cuda().train()

    video = torch.randn(8, T, 768, generator=g).cuda()
    audio = torch.randn(8, 1024, generator=g).cuda()
    labels = torch.randint(0, 6, (8,), generator=g).cuda()
    crit = mm.FocalLoss(2.0)

    def one(model, opt):
        opt.zero_grad()
        crit(model(video, audio)[1], labels).backward()
        opt.step()

    a, b = make(), make()
    oa = mm.FusedAdam(a, lr=1e-3, weight_decay=1e-4)
    for _ in range(2):
        one(a, oa)
    sd = oa.state_dict()
    assert len(sd["state"]) == len(list(a.parameters())) and int(sd["state"][0]["step"]) == 2
    # (1) into the stock optimizer of a model holding the same weights
    b.load_state_dict(a.state_dict())
    ob = torch.optim.Adam(b.parameters(), lr=1e-3, weight_decay=1e-4)
    ob.load_state_dict(sd)
    one(a, oa)
    one(b, ob)
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert float((pa - pb).abs().max()) < 2e-6, k
    # (2) back from the stock optimizer into a fresh FusedAdam (resume)
    c = make()
    c.load_state_dict(b.state_dict())
    oc = mm.FusedAdam(c, lr=1e-3, weight_decay=1e-4)
    oc.load_state_dict(ob.state_dict())
    assert oc._step == 3
    one(b, ob)
    one(c, oc)
    for (k, pb), (_, pc) in zip(b.named_parameters(), c.named_parameters()):
        assert float((pb - pc).abs().max()) < 2e-6, k
    # (3) the moments survive a re-allocation of the flat buffer (model.to / a sub-module call)
    m_before = oc._m.clone()
    c._engine.ctx._ptrs = []          # force a rebuild of the flat storage
    one(c, oc)
    assert oc._step == 5 and float(oc._m.abs().sum()) > 0 and oc._m.shape == m_before.shape


def test_fused_train_step_validates_its_inputs():
    """FusedTrainStep.step must reject what ModelFn rejects instead of reading garbage."""
    T = 4
    model = mm.MultimodalEmotionModel(max_seq_len=T + 1, classifier_hidden_dim=512).cuda().train()
    step = mm.FusedTrainStep(model, compute_dtype=torch.float32)
    v, a = torch.zeros(4, T, 768, device="cuda"), torch.zeros(4, 1024, device="cuda")
    y = torch.zeros(4, dtype=torch.int64, device="cuda")
    with pytest.raises(mm.MmerError):
        step.step(v, a, None, y.int())                                  # int32 labels
    with pytest.raises(mm.MmerError):
        step.step(v, a, None, y.cpu())                                  # CPU labels
    with pytest.raises(mm.MmerError):
        step.step(v, a, None, y[:3])                                    # wrong batch
    with pytest.raises(RuntimeError):
        step.step(torch.zeros(4, T + 1, 768, device="cuda"), a, None, y)    # T + 1 > max_seq_len (train2.py:160)
    with pytest.raises(RuntimeError):
        step.step(v, a[:, :512].contiguous(), None, y)                  # feature dimension
    with pytest.raises(RuntimeError):
        step.step(v, a, torch.zeros(4, T + 2, dtype=torch.bool, device="cuda"), y)
    loss, _ = step.step(v, a, torch.zeros(4, T, dtype=torch.uint8, device="cuda"), y)   # non-bool mask is converted
    assert np.isfinite(float(loss))
    bad = y.clone()
    bad[1] = 17
    loss, _ = step.step(v, a, None, bad)                                # out-of-range label: NaN, no illegal address
    assert np.isnan(float(loss))
    torch.cuda.synchronize()
    strict = mm.FusedTrainStep(model, compute_dtype=torch.float32, check_labels=True)
    with pytest.raises(mm.MmerError):
        strict.step(v, a, None, bad)


@pytest.mark.parametrize("B,T", [(3, 1), (5, 2), (17, 7), (33, 13), (9, 31), (130, 9), (7, 40), (300, 16), (2, 100),
                                 (600, 5)])
@pytest.mark.parametrize("bf16", [False, True])
def test_odd_shapes_training_step_matches_stock_reference(B, T, bf16):
    """Batch / sequence shapes that hit every kernel-selection branch around the model -- position-stable token assembly
    with every warp count and grid the plan can choose (S = 2 ... 41), its fallback for long sequences (S = 101), ragged
    last tiles, short / long attention, one-CTA-per-sample pooling -- against stock torch fp32 with a ragged padding
    mask: logits, loss and every parameter gradient."""
    model, ref = _pair(T, bf16=bf16)
    model.train()
    ref.train()
    # fixed seeds: the forward pass is deterministic, so whether some ReLU pre-activation lands within an fp32 ulp of zero
    # (and flips between this implementation and cuBLAS, moving one sample's gradient by a few percent -- seen once while
    # writing this test, confined to a single sample of the batch) is decided by the seed, not by the run
    video, audio, mask, labels = _batch(B, T, 1000 * B + T, True)
    alpha = ALPHA.cuda()
    _, lref = ref(video, audio, mask)
    loss_ref = torch.nn.functional.cross_entropy(lref, labels, weight=alpha)
    loss_ref.backward()
    vin, ain = (video.bfloat16(), audio.bfloat16()) if bf16 else (video, audio)
    _, logits, _ = model(vin, ain, mask=mask)
    loss = mm.WeightedCrossEntropyLoss(alpha)(logits, labels)
    loss.backward()
    tol = 2e-2 if bf16 else 1e-4
    assert float((logits.detach().float() - lref.detach()).abs().max()) < tol * float(lref.abs().max())
    assert abs(float(loss) - float(loss_ref)) < tol * abs(float(loss_ref))
    got = dict(model.named_parameters())
    for k, p in ref.named_parameters():
        g, r = got[k].grad.double(), p.grad.double()
        if float(r.norm()) < 1e-12:
            continue
        if bf16:   # direction and size (element-wise agreement is bounded by the bf16 storage noise floor)
            assert abs(float(g.norm()) / float(r.norm()) - 1) < 6e-2, k
            assert float((g * r).sum() / (g.norm() * r.norm())) > 0.97, k
        else:      # l2: one ReLU pre-activation within an ulp of zero may flip between two fp32 implementations and move a
            #           few elements by their full size (the full-size test compares both with float64 for that reason)
            assert float((g - r).norm()) < 2e-3 * float(r.norm()) + 1e-9, k
