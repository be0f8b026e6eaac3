"""The callers on either side of the model: batch assembly (train2.py:296-492) and evaluation
bookkeeping (row N3; train2.py:593-667).

CPU: the oracle restatement of load_data / collate_fn against what the unmodified reference returned for the same
synthetic feature files (tests/golden/data_v2_small.npz), plus the host-side helpers of mmer_b200.data.
GPU: DeviceFeatureSet / DeviceLoader (feature statistics + gather/pad/normalise kernel) against the same golden.
"""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_data import BATCH, synthetic_dataset  # noqa: E402
from oracle import data_oracle as D  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "data_v2_small.npz"))


def kept():
    names, videos, audios = synthetic_dataset()
    keep = [i for i, n in enumerate(names) if D.label_of(n) is not None]
    return [names[i] for i in keep], [videos[i] for i in keep], [audios[i] for i in keep]


def check_batches(tag, got, atol):
    assert len(got) == int(GOLD[f"{tag}/n"])
    for i, (v, a, y, m) in enumerate(got):
        v, a = v.float().cpu().numpy(), a.float().cpu().numpy()
        assert v.shape == GOLD[f"{tag}/{i}/video"].shape, (tag, i)
        np.testing.assert_array_equal(y.cpu().numpy(), GOLD[f"{tag}/{i}/labels"])
        np.testing.assert_array_equal(m.cpu().numpy(), GOLD[f"{tag}/{i}/mask"])
        np.testing.assert_allclose(v, GOLD[f"{tag}/{i}/video"], rtol=0, atol=atol)
        np.testing.assert_allclose(a, GOLD[f"{tag}/{i}/audio"], rtol=0, atol=atol)
        assert float(np.abs(v[m.cpu().numpy()]).max(initial=0.0)) == 0.0       # padding is exactly zero


def test_oracle_matches_reference_load_data():
    names, videos, audios = synthetic_dataset()
    dataset, (train, val, test), max_chunks, cw, _ = D.load_data(names, videos, audios)
    assert max_chunks == int(GOLD["max_chunks"])
    np.testing.assert_array_equal(cw.numpy(), GOLD["class_weights"])
    check_batches("val", D.batches(dataset, val, BATCH), 0.0)                  # same torch calls: bit-exact
    check_batches("test", D.batches(dataset, test, BATCH), 0.0)


def test_host_helpers_match_oracle():
    from mmer_b200 import data as P
    names, videos, audios = synthetic_dataset()
    for n in names + ["1001_DFA_NEU_XX.npy", "03-02-08-01-01-01-01.npy"]:
        assert P.label_from_filename(n) == D.label_of(n)
    labels = [D.label_of(n) for n in kept()[0]]
    _, (train, val, test), _, cw, _ = D.load_data(names, videos, audios)
    assert P.stratified_split(labels) == (train, val, test)
    np.testing.assert_allclose(P.balanced_class_weights([labels[i] for i in train]).numpy(), GOLD["class_weights"], rtol=1e-7)


def test_loader_shuffle_order_is_the_random_samplers():
    """DeviceLoader draws its permutation exactly like torch's RandomSampler (checked against a real DataLoader)."""
    from mmer_b200.data import DeviceLoader

    class Recorder:   # stands in for the feature set: records the index lists it is asked to collate
        def collate(self, idx, dtype):
            return list(idx)

    indices = [7, 3, 9, 11, 2, 5, 8, 1, 0, 4, 6]
    torch.manual_seed(99)
    ref = [b.tolist() for b in torch.utils.data.DataLoader(indices, batch_size=4, shuffle=True)]
    torch.manual_seed(99)
    got = list(DeviceLoader(Recorder(), indices, 4, True, torch.float32))
    assert got == ref and len(got) == 3 and len(got[-1]) == 3
    assert list(DeviceLoader(Recorder(), indices, 4, False, torch.float32)) == [indices[0:4], indices[4:8], indices[8:]]
    # an UNSHUFFLED pass also advances the global RNG exactly like a DataLoader pass (its iterator draws a base seed),
    # so that the next epoch's shuffle matches the reference's (train2.py:564-667: train, val, test passes per epoch)
    torch.manual_seed(7)
    list(torch.utils.data.DataLoader(indices, batch_size=4, shuffle=False))
    ref2 = [b.tolist() for b in torch.utils.data.DataLoader(indices, batch_size=4, shuffle=True)]
    torch.manual_seed(7)
    list(DeviceLoader(Recorder(), indices, 4, False, torch.float32))
    assert list(DeviceLoader(Recorder(), indices, 4, True, torch.float32)) == ref2


# ------------------------------------------------------------------------------------------------ GPU
def device_set(normalize=True):
    import mmer_b200 as mm
    names, videos, audios = kept()
    labels = [D.label_of(n) for n in names]
    return mm.DeviceFeatureSet(videos, audios, labels, device="cuda", normalize=normalize), labels


@pytest.mark.gpu
def test_device_feature_set_matches_reference_golden():
    import mmer_b200 as mm
    ds, labels = device_set()
    names, videos, audios = synthetic_dataset()
    _, (train, val, test), max_chunks, cw, stats = D.load_data(names, videos, audios)
    assert ds.max_chunks == max_chunks and len(ds) == len(labels)
    for got, ref in zip((ds.video_mean, ds.video_std, ds.audio_mean, ds.audio_std), stats):
        np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=2e-6, atol=1e-6)
    # statistics come from a different (double, two-pass) summation than torch's: normalised values within 1e-5
    check_batches("val", list(ds.loader(val, BATCH)), 1e-5)
    check_batches("test", list(ds.loader(test, BATCH)), 1e-5)
    torch.manual_seed(1234)
    check_batches("train", list(ds.loader(train, BATCH, shuffle=True)), 1e-5)   # the reference's shuffled order
    assert mm.data.stratified_split(labels) == (train, val, test)


@pytest.mark.gpu
def test_collate_is_bit_exact_given_the_reference_statistics():
    """With the reference's own mean / std the kernel's (x - mean) / std equals torch's to the bit."""
    ds, _ = device_set()
    names, videos, audios = synthetic_dataset()
    dataset, (train, val, test), _, _, stats = D.load_data(names, videos, audios)
    ds.video_mean, ds.video_std, ds.audio_mean, ds.audio_std = (s.cuda() for s in stats)
    check_batches("val", list(ds.loader(val, BATCH)), 0.0)
    check_batches("test", list(ds.loader(test, BATCH)), 0.0)


@pytest.mark.gpu
def test_collate_bf16_unnormalised_and_errors():
    ds, labels = device_set(normalize=False)          # train.py behaviour: features as they are
    names, videos, audios = kept()
    idx = [5, 0, 17, 3]
    v, a, y, m = ds.collate(idx, dtype=torch.bfloat16)
    rv, ra, ry, rm = D.collate([(torch.from_numpy(videos[i]), torch.from_numpy(audios[i]), labels[i]) for i in idx])
    assert v.dtype == torch.bfloat16 and torch.equal(v.cpu(), rv.bfloat16()) and torch.equal(a.cpu(), ra.bfloat16())
    assert torch.equal(y.cpu(), ry) and torch.equal(m.cpu(), rm)
    with pytest.raises(IndexError):
        ds.collate([len(ds)])
    with pytest.raises(ValueError):
        ds.collate([])


@pytest.mark.gpu
def test_collated_batch_feeds_the_model_at_full_width():
    """768 / 1024-wide features, ragged lengths, straight into the fused step (bf16)."""
    import mmer_b200 as mm
    g = torch.Generator().manual_seed(3)
    n = 96
    lens = torch.randint(1, 17, (n,), generator=g).tolist()
    videos = [torch.randn(t, 768, generator=g) * 2 + 1 for t in lens]
    audios = [torch.randn(1024, generator=g) for _ in range(n)]
    labels = torch.randint(0, 6, (n,), generator=g).tolist()
    ds = mm.DeviceFeatureSet(videos, audios, labels)
    ref_v = (torch.cat(videos).double().mean(0), torch.cat(videos).double().std(0) + 1e-6)
    np.testing.assert_allclose(ds.video_mean.cpu().numpy(), ref_v[0].numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ds.video_std.cpu().numpy(), ref_v[1].numpy(), rtol=1e-5, atol=1e-6)
    model = mm.MultimodalEmotionModel(max_seq_len=17, fusion_num_layers=2, classifier_hidden_dim=512).cuda().train()
    step = mm.FusedTrainStep(model, lr=1e-4, loss="focal", compute_dtype=torch.bfloat16)
    losses = []
    for v, a, y, m in ds.loader(range(n), batch_size=32, dtype=torch.bfloat16):
        assert v.shape[1] == max(lens[i] for i in range(len(losses) * 32, len(losses) * 32 + v.shape[0]))
        loss, _ = step.step(v, a, m, y)
        losses.append(float(loss))
    assert len(losses) == 3 and all(np.isfinite(losses))


# ------------------------------------------------------------------------------------------------ evaluation (N3)
def _sk_metrics(y, p):
    from sklearn.metrics import precision_recall_fscore_support
    out = {}
    for avg in ("macro", "micro"):
        pr, rc, f1, _ = precision_recall_fscore_support(y, p, average=avg, zero_division=0)   # train2.py:633-644
        out[avg + "_precision"], out[avg + "_recall"], out[avg + "_f1"] = pr, rc, f1
    return out


def test_metrics_from_confusion_equal_sklearn():
    """Every number the reference logs per epoch, including absent classes and 0/0 cases."""
    from sklearn.metrics import confusion_matrix
    from mmer_b200.evaluation import metrics_from_confusion
    rng = np.random.default_rng(0)
    cases = [(rng.integers(0, 6, 500), rng.integers(0, 6, 500)),
             (rng.integers(0, 3, 40), rng.integers(2, 6, 40)),            # classes never predicted / never true
             (np.array([0, 0, 1, 1]), np.array([2, 2, 2, 2])),            # nothing right: 0/0 in F1
             (np.array([4]), np.array([4]))]
    for y, p in cases:
        got = metrics_from_confusion(confusion_matrix(y, p, labels=list(range(6))))
        for k, v in _sk_metrics(y, p).items():
            assert abs(got[k] - v) < 1e-12, (k, got[k], v)
        assert got["accuracy"] == 100.0 * float((y == p).sum()) / len(y) and got["total"] == len(y)


@pytest.mark.gpu
def test_eval_accumulator_matches_the_reference_loop():
    """Three batches through EvalAccumulator == the reference's per-batch bookkeeping (train2.py:593-609) + sklearn."""
    import mmer_b200 as mm
    from sklearn.metrics import confusion_matrix
    g = torch.Generator().manual_seed(0)
    acc = mm.EvalAccumulator(6, keep_predictions=True)
    all_p, all_y, losses = [], [], []
    for B in (4096, 33, 1):
        probs = torch.softmax(3 * torch.randn(B, 6, generator=g), dim=1)
        probs[0, :] = 1.0 / 6                                             # an exact tie: first index, like torch.max
        labels = torch.randint(0, 6, (B,), generator=g)
        loss = torch.rand((), generator=g)
        acc.update(probs.cuda(), labels.cuda(), loss.cuda())
        all_p.extend(torch.max(probs, dim=1)[1].tolist())
        all_y.extend(labels.tolist())
        losses.append(float(loss))
    out = acc.result()
    np.testing.assert_array_equal(acc.confusion_matrix(), confusion_matrix(all_y, all_p, labels=list(range(6))))
    assert acc.predictions().tolist() == all_p
    assert out["total"] == len(all_y) and out["correct"] == sum(int(a == b) for a, b in zip(all_p, all_y))
    assert abs(out["accuracy"] - 100.0 * out["correct"] / out["total"]) < 1e-9
    assert abs(out["avg_loss"] - sum(losses) / 3) < 1e-6
    for k, v in _sk_metrics(all_y, all_p).items():
        assert abs(out[k] - v) < 1e-12, k


@pytest.mark.gpu
@pytest.mark.parametrize("dv,da,lens", [(10, 6, [3, 1, 4]), (16, 8, [5]), (12, 4, [2, 2, 2, 2]), (768, 1024, [1, 16, 7])])
def test_collate_edge_shapes(dv, da, lens):
    """Widths that are not multiples of 4 (scalar path), a single sample, no padding at all, one-frame samples; fp32 output
    is bit-exact against pad_sequence / stack given the same statistics."""
    import mmer_b200 as mm
    g = torch.Generator().manual_seed(len(lens) * 100 + dv)
    videos = [torch.randn(t, dv, generator=g) * 3 - 1 for t in lens]
    audios = [torch.randn(da, generator=g) for _ in lens]
    labels = list(range(len(lens)))
    ds = mm.DeviceFeatureSet(videos, audios, labels, normalize=len(lens) > 1)
    if len(lens) > 1:
        vm, vs, am, as_ = D.global_stats(videos, audios)
        ds.video_mean, ds.video_std, ds.audio_mean, ds.audio_std = vm.cuda(), vs.cuda(), am.cuda(), as_.cuda()
        nv, na = [(v - vm) / vs for v in videos], [(a - am) / as_ for a in audios]
    else:
        nv, na = videos, audios
    order = list(reversed(range(len(lens))))
    v, a, y, m = ds.collate(order)
    rv, ra, ry, rm = D.collate([(nv[i], na[i], labels[i]) for i in order])
    assert torch.equal(v.cpu(), rv) and torch.equal(a.cpu(), ra) and torch.equal(y.cpu(), ry) and torch.equal(m.cpu(), rm)
    assert ds.max_chunks == max(lens)


@pytest.mark.gpu
@pytest.mark.parametrize("dv,da", [(768, 1024), (20, 6)])
def test_bf16_resident_set_gives_the_same_batches(dv, da):
    """store_dtype=bfloat16 (normalise + round once, then 2-byte gathers) == bf16 batches of the fp32-resident set, bit
    for bit, including padding, mask and labels; with and without normalisation."""
    import mmer_b200 as mm
    g = torch.Generator().manual_seed(dv)
    n = 40
    lens = torch.randint(1, 9, (n,), generator=g).tolist()
    videos = [torch.randn(t, dv, generator=g) * 2 + 0.5 for t in lens]
    audios = [torch.randn(da, generator=g) for _ in range(n)]
    labels = torch.randint(0, 6, (n,), generator=g).tolist()
    for normalize in (True, False):
        a32 = mm.DeviceFeatureSet(videos, audios, labels, normalize=normalize)
        a16 = mm.DeviceFeatureSet(videos, audios, labels, normalize=normalize, store_dtype=torch.bfloat16)
        assert a16.frames.dtype == torch.bfloat16 and a16.frames.shape == a32.frames.shape
        for idx in ([3, 17, 0, 39, 8], list(range(n))):
            x, y = a32.collate(idx, torch.bfloat16), a16.collate(idx, torch.bfloat16)
            assert all(torch.equal(p, q) for p, q in zip(x, y))
        with pytest.raises(mm.MmerError):
            a16.collate([0], torch.float32)
