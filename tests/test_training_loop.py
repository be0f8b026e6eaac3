"""SURVEY 8f rows N2 / N3 as product code: ``mmer_b200.load_data`` (``.npy`` directories), ``HostBatchStager`` and
``mmer_b200.train_model`` against what the UNMODIFIED reference produced on the same synthetic files:
tests/golden/data_v2_small.npz (train2.load_data) and tests/golden/train_v2_small.npz (train2.train_model, 3 epochs;
generator tests/golden/make_golden_train.py)."""
import glob
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import detgen  # noqa: E402
from make_golden_data import BATCH, DA, DV, synthetic_dataset  # noqa: E402

DATA = np.load(os.path.join(HERE, "golden", "data_v2_small.npz"))
TRAIN = np.load(os.path.join(HERE, "golden", "train_v2_small.npz"))


def write_files(tmp, skip_video=None):
    names, videos, audios = synthetic_dataset()
    vdir, adir = os.path.join(tmp, "v"), os.path.join(tmp, "a")
    os.makedirs(vdir)
    os.makedirs(adir)
    for n, v, a in zip(names, videos, audios):
        if n != skip_video:
            np.save(os.path.join(vdir, n), v)
        np.save(os.path.join(adir, n), a)
    return vdir, adir, names


def test_pairing_by_zip_reproduces_the_reference_and_by_stem_repairs_it(tmp_path):
    """train2.py:318-325 zips the two sorted globs; one missing video file shifts every later pair."""
    from mmer_b200.training import list_feature_pairs
    names = synthetic_dataset()[0]
    vdir, adir, _ = write_files(str(tmp_path), skip_video=names[10])
    zipped = list_feature_pairs(vdir, adir, "zip")
    stem = list_feature_pairs(vdir, adir, "stem")
    assert len(zipped) == len(names) - 1 and len(stem) == len(names) - 1
    assert all(os.path.basename(v) == os.path.basename(a) for v, a in stem)
    wrong = [i for i, (v, a) in enumerate(zipped) if os.path.basename(v) != os.path.basename(a)]
    assert wrong == list(range(10, len(names) - 1))               # everything after the gap is misaligned, like the reference
    with pytest.raises(ValueError):
        list_feature_pairs(vdir, adir, "name")


@pytest.mark.gpu
def test_load_data_from_npy_directories_matches_reference_loaders(tmp_path):
    import mmer_b200 as mm
    vdir, adir, _ = write_files(str(tmp_path))
    train_loader, val_loader, test_loader, max_chunks, class_weights = mm.load_data(vdir, adir, batch_size=BATCH,
                                                                                    verbose=False)
    assert max_chunks == int(DATA["max_chunks"])
    np.testing.assert_allclose(class_weights.numpy(), DATA["class_weights"], rtol=1e-7)
    for tag, loader in (("val", val_loader), ("test", test_loader)):
        got = list(loader)
        assert len(got) == int(DATA[f"{tag}/n"]) == len(loader)
        for i, (v, a, y, m) in enumerate(got):
            assert v.is_cuda and m.dtype == torch.bool
            np.testing.assert_array_equal(y.cpu().numpy(), DATA[f"{tag}/{i}/labels"])
            np.testing.assert_array_equal(m.cpu().numpy(), DATA[f"{tag}/{i}/mask"])
            np.testing.assert_allclose(v.cpu().numpy(), DATA[f"{tag}/{i}/video"], rtol=0, atol=1e-5)
            np.testing.assert_allclose(a.cpu().numpy(), DATA[f"{tag}/{i}/audio"], rtol=0, atol=1e-5)
    torch.manual_seed(1234)
    got = list(train_loader)                                       # the reference's shuffled order after the same seed
    assert len(got) == int(DATA["train/n"])
    for i, (v, a, y, m) in enumerate(got):
        np.testing.assert_array_equal(y.cpu().numpy(), DATA[f"train/{i}/labels"])
        np.testing.assert_allclose(v.cpu().numpy(), DATA[f"train/{i}/video"], rtol=0, atol=1e-5)
    assert len(train_loader.dataset) + len(val_loader.dataset) + len(test_loader.dataset) == \
        sum(int(DATA[f"{t}/{i}/labels"].size) for t in ("train", "val", "test") for i in range(int(DATA[f"{t}/n"])))


@pytest.mark.gpu
def test_host_batch_stager_delivers_every_batch_in_order_and_feeds_the_step():
    import mmer_b200 as mm
    g = torch.Generator().manual_seed(5)
    batches = []
    for i in range(7):
        b, t = 4 + (i % 3), 3 + (i % 4)                           # shapes change from batch to batch (padded length varies)
        v, a = torch.randn(b, t, 768, generator=g), torch.randn(b, 1024, generator=g)
        y = torch.randint(0, 6, (b,), generator=g)
        m = torch.zeros(b, t, dtype=torch.bool)
        m[:, t - 1] = i % 2 == 0
        if i % 2:
            v, a = v.pin_memory(), a.pin_memory()                  # pinned and pageable sources both work
        batches.append((v, a, m if i != 3 else None, y))
    stager = mm.HostBatchStager("cuda", depth=3)
    seen = 0
    for i, (v, a, m, y) in enumerate(stager.pipeline(batches)):
        hv, ha, hm, hy = batches[i]
        assert v.is_cuda and torch.equal(v.cpu(), hv) and torch.equal(a.cpu(), ha) and torch.equal(y.cpu(), hy)
        assert (m is None) == (hm is None) and (m is None or torch.equal(m.cpu(), hm))
        seen += 1
    assert seen == len(batches) and stager.bytes_staged > 0
    assert list(stager.pipeline([])) == []

    def losses(feed):
        torch.manual_seed(0)
        model = mm.MultimodalEmotionModel(max_seq_len=8, classifier_hidden_dim=512, fusion_dropout=0.0,
                                          classifier_dropout=0.0).cuda().train()
        step = mm.FusedTrainStep(model, lr=1e-3, compute_dtype=torch.float32)
        return [float(step.step(v, a, m, y)[0]) for v, a, m, y in feed]

    direct = losses([(v.cuda(), a.cuda(), None if m is None else m.cuda(), y.cuda()) for v, a, m, y in batches])
    staged = losses(mm.HostBatchStager("cuda", depth=2).pipeline(batches))
    # staging changes nothing but where the data waits (split-K fp32 atomics make two runs differ in the last bit)
    np.testing.assert_allclose(staged, direct, rtol=1e-5)


@pytest.mark.gpu
def test_train_model_matches_the_references_own_train_model(tmp_path):
    """3 epochs of train2.train_model (weighted CE, clip 1.0, Adam, ReduceLROnPlateau, early-stop bookkeeping, test pass,
    confusion matrix, results_*.json / best_model_*.pth / final_model_*.pth) on the fused device path, fp32 mode,
    against the log the unmodified reference wrote for the same files, weights and seed."""
    import mmer_b200 as mm
    vdir, adir, _ = write_files(str(tmp_path))
    train_loader, val_loader, test_loader, max_chunks, class_weights = mm.load_data(vdir, adir, batch_size=BATCH,
                                                                                    verbose=False)
    model = mm.MultimodalEmotionModel(video_dim=DV, audio_dim=DA, fused_dim=64, num_classes=6, max_seq_len=max_chunks + 1,
                                      fusion_num_layers=2, fusion_num_heads=2, fusion_dropout=0.0,
                                      classifier_hidden_dim=32, classifier_dropout=0.0)
    P = detgen.make_params("v2", max_seq_len=max_chunks + 1, video_dim=DV, audio_dim=DA, fused=64, hidden=32)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in P.items()}, strict=True)
    ref_log = json.loads(str(TRAIN["log_json"]))
    out_dir = str(tmp_path / "training_runs_2")
    torch.manual_seed(1234)
    out = mm.train_model(model, train_loader, val_loader, test_loader, class_weights, num_epochs=3, lr=1e-3,
                         batch_size=BATCH, device="cuda", out_dir=out_dir, verbose=False)
    # ---- the log: same keys, same epochs, losses to fp32 training noise, classification metrics equal
    assert len(out["training_progress"]) == len(ref_log["training_progress"]) == 3
    for got, ref in zip(out["training_progress"], ref_log["training_progress"]):
        assert list(got.keys()) == list(ref.keys())
        assert got["epoch"] == ref["epoch"]
        assert abs(got["train_loss"] - ref["train_loss"]) < 3e-3, (got, ref)
        assert abs(got["val_loss"] - ref["val_loss"]) < 3e-3, (got, ref)
        for k in ref:
            if k.startswith(("val_", "test_")) and k != "val_loss":
                assert abs(got[k] - ref[k]) < 1e-9, (k, got[k], ref[k])
    assert out["best_model"] == ref_log["best_model"]
    hp, hp_ref = out["hyperparameters"], dict(ref_log["hyperparameters"])
    hp_ref["device"] = "cuda"                                       # the golden ran on the container's CPU
    assert hp == hp_ref
    np.testing.assert_array_equal(out["confusion_matrix"], TRAIN["confusion_matrix"])
    # ---- the files: reference names and layout
    names = sorted(os.path.basename(p).rsplit("_", 2)[0] for p in glob.glob(os.path.join(out_dir, "*")))
    assert names == [str(x) for x in TRAIN["files"]]
    on_disk = json.load(open(out["paths"]["results"]))
    assert list(on_disk.keys()) == ["training_progress", "best_model", "hyperparameters"]
    final = torch.load(out["paths"]["final_model"])
    best = torch.load(out["paths"]["best_model"])
    assert list(final.keys()) == list(model.state_dict().keys()) == list(best.keys())
    for k, v in final.items():
        ref = TRAIN["final/" + k]
        f = v.detach().double().flatten().cpu()
        assert abs(float(f.norm()) - ref[0]) < 2e-3 * ref[0] + 1e-5, k
        head = np.pad(f[:16].numpy(), (0, max(0, 16 - f.numel())))
        np.testing.assert_allclose(head, ref[2:], rtol=0, atol=3e-3, err_msg=k)
        # the reference's "best" state aliases the live weights (train2.py:619: a shallow .copy()): best == final
        np.testing.assert_array_equal(TRAIN["best/" + k], TRAIN["final/" + k])
        assert torch.equal(best[k].cpu(), v.cpu())
    # ---- a checkpoint written here loads into the stock-module restatement of the reference and vice versa
    from oracle import eager_torch as E
    stock = E.EagerModel(video_dim=DV, audio_dim=DA, fused_dim=64, max_seq_len=max_chunks + 1, fusion_num_layers=2,
                         fusion_num_heads=2, classifier_hidden_dim=32)
    stock.load_state_dict(final, strict=True)


@pytest.mark.gpu
def test_train_model_early_stopping_and_true_best_state(tmp_path):
    """The reference's rule (train2.py:622-631): stop once `patience` epochs failed to improve the validation loss by
    1e-4 over the previous epoch; the stopping epoch is not logged.  copy_best_state=True keeps a real snapshot."""
    import mmer_b200 as mm
    vdir, adir, _ = write_files(str(tmp_path))
    loaders = mm.load_data(vdir, adir, batch_size=BATCH, verbose=False)
    model = mm.MultimodalEmotionModel(video_dim=DV, audio_dim=DA, fused_dim=64, num_classes=6, max_seq_len=loaders[3] + 1,
                                      fusion_num_layers=1, fusion_num_heads=2, fusion_dropout=0.0,
                                      classifier_hidden_dim=32, classifier_dropout=0.0)
    torch.manual_seed(7)
    out = mm.train_model(model, *loaders[:3], loaders[4], num_epochs=40, lr=3e-2, patience=2, batch_size=BATCH,
                         out_dir=None, copy_best_state=True, verbose=False)
    logged = out["training_progress"]
    assert 1 <= len(logged) < 40                                    # an lr this large overfits 60 samples: val loss turns
    val = [e["val_loss"] for e in logged]
    assert out["best_model"]["epoch"] >= 1 and out["paths"] == {}
    # replay the rule on the logged losses: the run must have stopped exactly where the reference's counter says
    prev, bad = float("inf"), 0
    for v in val:
        bad = bad + 1 if prev - v < 1e-4 else 0
        assert bad < 2
        prev = v
    # the model now holds the weights of the best epoch (snapshot restored before the final test pass)
    assert abs(min(val) - out["best_val_loss"]) < 1e-12 or out["best_val_loss"] <= min(val)
