"""Golden epoch-loop vectors from the UNMODIFIED reference ``train2.load_data`` + ``train2.train_model`` (build container
only; needs /root/reference):

    python tests/golden/make_golden_train.py

The synthetic feature files of make_golden_data.synthetic_dataset() are written to a temporary directory, the reference's
own load_data builds the loaders, a small reference MultimodalEmotionModel (dropout 0, deterministic detgen weights)
is trained by the reference's own train_model for 3 epochs on the CPU after ``torch.manual_seed(1234)``, and what it
writes is stored: the per-epoch log of results_*.json (train / validation loss, accuracies, macro / micro P / R / F1),
best epoch, hyperparameters, a summary of every parameter of the final and of the "best" checkpoint, and the confusion
matrix it prints (captured by wrapping sklearn's confusion_matrix in the reference module's namespace).
"""
import glob
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import detgen  # noqa: E402
from make_golden import import_reference, summarize  # noqa: E402
from make_golden_data import BATCH, DA, DV, synthetic_dataset  # noqa: E402

DIMS = dict(video_dim=DV, audio_dim=DA, fused=64, hidden=32)
EPOCHS, LR = 3, 1e-3


def main():
    _, ref = import_reference()
    names, videos, audios = synthetic_dataset()
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        vdir, adir = os.path.join(tmp, "v"), os.path.join(tmp, "a")
        os.makedirs(vdir)
        os.makedirs(adir)
        for n, v, a in zip(names, videos, audios):
            np.save(os.path.join(vdir, n), v)
            np.save(os.path.join(adir, n), a)
        os.chdir(tmp)
        try:
            train_loader, val_loader, test_loader, max_chunks, class_weights = ref.load_data(vdir, adir, batch_size=BATCH)
            model = ref.MultimodalEmotionModel(video_dim=DV, audio_dim=DA, fused_dim=64, num_classes=6,
                                               max_seq_len=max_chunks + 1, fusion_num_layers=2, fusion_num_heads=2,
                                               fusion_dropout=0.0, classifier_hidden_dim=32, classifier_dropout=0.0)
            for m in model.modules():
                if isinstance(m, torch.nn.MultiheadAttention):
                    m.dropout = 0.0
            params = detgen.make_params("v2", max_seq_len=max_chunks + 1, **DIMS)
            model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}, strict=True)
            captured = {}
            real_cm = ref.confusion_matrix

            def spy(*a, **k):
                captured["cm"] = real_cm(*a, **k)
                return captured["cm"]

            ref.confusion_matrix = spy
            torch.manual_seed(1234)
            torch.set_num_threads(4)
            ref.train_model(model, train_loader, val_loader, test_loader, class_weights, num_epochs=EPOCHS, lr=LR,
                            batch_size=BATCH, device="cpu")
            ref.confusion_matrix = real_cm
            log = json.load(open(glob.glob(os.path.join(tmp, "training_runs_2", "results_*.json"))[0]))
            best = torch.load(glob.glob(os.path.join(tmp, "training_runs_2", "best_model_*.pth"))[0])
            final = torch.load(glob.glob(os.path.join(tmp, "training_runs_2", "final_model_*.pth"))[0])
            out["files"] = np.array(sorted(os.path.basename(p).rsplit("_", 2)[0] for p in
                                           glob.glob(os.path.join(tmp, "training_runs_2", "*"))))
        finally:
            os.chdir(cwd)
    out["max_chunks"] = max_chunks
    out["class_weights"] = class_weights.numpy()
    out["log_json"] = np.array(json.dumps(log))
    out["confusion_matrix"] = captured["cm"]
    for k, v in final.items():
        out["final/" + k] = summarize(v)
        out["best/" + k] = summarize(best[k])
    path = os.path.join(HERE, "train_v2_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for e in log["training_progress"]:
        print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in e.items() if k in
               ("epoch", "train_loss", "val_loss", "val_acc", "test_acc", "val_macro_f1")})
    print("best epoch", log["best_model"], "cm\n", captured["cm"])


if __name__ == "__main__":
    main()
