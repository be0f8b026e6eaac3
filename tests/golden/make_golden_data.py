"""Golden data-path vectors from the UNMODIFIED reference ``train2.load_data`` (build container only).

    python tests/golden/make_golden_data.py

Writes the synthetic feature files of ``synthetic_dataset()`` (RAVDESS- and CREMA-D-style names, including the two
RAVDESS classes the reference skips) into a temporary directory, calls the reference's load_data on it and stores what
its loaders yield: every validation / test batch (unshuffled), the first pass over the shuffled training loader after
``torch.manual_seed(1234)``, max_chunks and the class weights.  Feature widths are small (16 / 8) to keep the fixture
small; nothing in the reference's data path depends on them.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
DV, DA, BATCH = 16, 8, 5


def synthetic_dataset():
    """(sorted names, videos [T_i, DV], audios [DA]) -- deterministic; T_i in 1..6; non-trivial means and scales."""
    rng = np.random.default_rng(20240607)
    names = []
    for actor in range(1, 8):
        for emo in (1, 2, 3, 4, 5, 6, 7, 8):
            names.append(f"03-01-{emo:02d}-01-02-01-{actor:02d}.npy")
    for spk in range(1001, 1008):
        for emo in ("ANG", "DIS", "FEA", "HAP", "NEU", "SAD"):
            names.append(f"{spk}_DFA_{emo}_XX.npy")
    names = sorted(names)
    scale_v = rng.uniform(0.2, 3.0, DV).astype(np.float32)
    shift_v = rng.uniform(-5.0, 5.0, DV).astype(np.float32)
    scale_a = rng.uniform(0.5, 2.0, DA).astype(np.float32)
    shift_a = rng.uniform(-1.0, 1.0, DA).astype(np.float32)
    videos, audios = [], []
    for _ in names:
        t = int(rng.integers(1, 7))
        videos.append((rng.standard_normal((t, DV)).astype(np.float32) * scale_v + shift_v).astype(np.float32))
        audios.append((rng.standard_normal(DA).astype(np.float32) * scale_a + shift_a).astype(np.float32))
    return names, videos, audios


def main():
    sys.path.insert(0, HERE)
    from make_golden import import_reference
    _, ref_v2 = import_reference()
    names, videos, audios = synthetic_dataset()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        vdir, adir = os.path.join(tmp, "v"), os.path.join(tmp, "a")
        os.makedirs(vdir)
        os.makedirs(adir)
        for n, v, a in zip(names, videos, audios):
            np.save(os.path.join(vdir, n), v)
            np.save(os.path.join(adir, n), a)
        train_loader, val_loader, test_loader, max_chunks, class_weights = ref_v2.load_data(vdir, adir, batch_size=BATCH)
        out["max_chunks"] = max_chunks
        out["class_weights"] = class_weights.numpy()
        for tag, loader in (("val", val_loader), ("test", test_loader)):
            for i, (v, a, y, m) in enumerate(loader):
                out[f"{tag}/{i}/video"], out[f"{tag}/{i}/audio"] = v.numpy(), a.numpy()
                out[f"{tag}/{i}/labels"], out[f"{tag}/{i}/mask"] = y.numpy(), m.numpy()
            out[f"{tag}/n"] = len(loader)
        torch.manual_seed(1234)
        for i, (v, a, y, m) in enumerate(train_loader):
            out[f"train/{i}/video"], out[f"train/{i}/audio"] = v.numpy(), a.numpy()
            out[f"train/{i}/labels"], out[f"train/{i}/mask"] = y.numpy(), m.numpy()
        out["train/n"] = len(train_loader)
    path = os.path.join(HERE, "data_v2_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", out["train/n"], "train batches,", out["val/n"], "val,", out["test/n"], "test")


if __name__ == "__main__":
    main()
