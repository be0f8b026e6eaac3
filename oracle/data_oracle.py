"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's data path between the feature
files and the model -- ``load_data`` and its closure ``collate_fn`` (train2.py:296-492) -- on in-memory arrays.

Pinned by tests/golden/data_v2_small.npz, which holds what the UNMODIFIED ``train2.load_data`` returned for the same
synthetic feature files (tests/golden/make_golden_data.py): statistics-normalised batches of the validation and test
loaders, a seeded pass over the shuffled training loader, ``max_chunks`` and the class weights.
"""
import numpy as np
import torch
from sklearn.model_selection import train_test_split
from sklearn.utils.class_weight import compute_class_weight
from torch.nn.utils.rnn import pad_sequence


def label_of(basename):
    """train2.py:327-352.  None = the reference skips the file."""
    if "-" in basename:
        label_num = int(basename.split("-")[2])
        if label_num in [2, 8]:
            return None
        return {1: 0, 3: 1, 4: 2, 5: 3, 6: 4, 7: 5}[label_num]
    label_num = {"ANG": 5, "DIS": 7, "FEA": 6, "HAP": 3, "NEU": 1, "SAD": 4}[basename.split("_")[2]]
    return {1: 0, 3: 1, 4: 2, 5: 3, 6: 4, 7: 5}[label_num]


def global_stats(video_features, audio_features):
    """train2.py:430-441."""
    all_video = torch.cat(video_features, dim=0)
    all_audio = torch.stack(audio_features, dim=0)
    return (all_video.mean(dim=0), all_video.std(dim=0) + 1e-6, all_audio.mean(dim=0), all_audio.std(dim=0) + 1e-6)


def collate(batch):
    """train2.py:418-440 (the closure collate_fn)."""
    videos, audios, labels = zip(*batch)
    videos_padded = pad_sequence(videos, batch_first=True, padding_value=0.0)
    audios_stacked = torch.stack(audios)
    labels_tensor = torch.tensor(labels, dtype=torch.long)
    masks = [torch.zeros(len(v), dtype=torch.bool) for v in videos]
    masks_padded = pad_sequence(masks, batch_first=True, padding_value=True)
    return videos_padded, audios_stacked, labels_tensor, masks_padded


def load_data(names, videos, audios):
    """names sorted like ``sorted(glob(...))``; returns (dataset, (train, val, test) indices, max_chunks, class_weights,
    stats) following train2.py:312-488."""
    vf, af, labels = [], [], []
    for name, v, a in zip(names, videos, audios):
        lab = label_of(name)
        if lab is None:
            continue
        vf.append(torch.from_numpy(np.asarray(v, dtype=np.float32)))
        af.append(torch.from_numpy(np.asarray(a, dtype=np.float32)))
        labels.append(lab)
    stats = global_stats(vf, af)
    vm, vs, am, as_ = stats
    vf = [(v - vm) / vs for v in vf]
    af = [(a - am) / as_ for a in af]
    max_chunks = max(v.shape[0] for v in vf)
    dataset = list(zip(vf, af, labels))
    indices = list(range(len(dataset)))
    train, temp = train_test_split(indices, test_size=0.2, random_state=42, stratify=labels)
    val, test = train_test_split(temp, test_size=0.5, random_state=42, stratify=[labels[i] for i in temp])
    train_labels = [labels[i] for i in train]
    cw = torch.tensor(compute_class_weight(class_weight="balanced", classes=np.unique(train_labels), y=train_labels),
                      dtype=torch.float32)
    cw[4] = cw[4] * 1.2
    cw[5] = cw[5] * 1.2
    return dataset, (train, val, test), max_chunks, cw, stats


def batches(dataset, indices, batch_size):
    """DataLoader(..., shuffle=False, collate_fn=collate_fn): consecutive slices, last one short."""
    sub = [dataset[i] for i in indices]
    return [collate(sub[i:i + batch_size]) for i in range(0, len(sub), batch_size)]
