"""Batch assembly on the device: the reference's ``load_data`` post-processing and ``collate_fn``, HBM-resident.

The reference (train2.py:296-492; train.py:176-230 is the same without the z-score) keeps the pre-extracted features
as a Python list of CPU tensors, normalises them one by one with global statistics and lets ``DataLoader`` call a
Python ``collate_fn`` (``pad_sequence`` + ``stack``) per batch, followed by ``.to(device)`` of the padded batch.  At
the rate the fused step consumes samples that host loop is the bottleneck, so here the whole feature set lives on the
GPU as one ragged ``[total_frames, Dv]`` array and a batch is ONE gather kernel (``mmer_collate``) that normalises,
pads, casts and builds the mask in a single pass.  Results equal the reference's
``(videos_padded, audios_stacked, labels_tensor, masks_padded)``.

Host-side helpers restate the reference's small pure-Python pieces: the filename -> label maps (train2.py:327-352),
the stratified 80/10/10 split (train2.py:399-413) and the boosted balanced class weights (train2.py:474-488).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import BF16, F32, MmerError

__all__ = ["DeviceFeatureSet", "DeviceLoader", "label_from_filename", "stratified_split", "balanced_class_weights"]

_RAVDESS = {1: 0, 3: 1, 4: 2, 5: 3, 6: 4, 7: 5}                       # train2.py:338
_CREMAD = {"ANG": 5, "DIS": 7, "FEA": 6, "HAP": 3, "NEU": 1, "SAD": 4}    # train2.py:343


def label_from_filename(basename: str) -> Optional[int]:
    """train2.py:327-352: RAVDESS ``03-01-05-...`` (third field; classes 02 and 08 are skipped -> None) or CREMA-D
    ``1001_DFA_ANG_XX`` (third field), both mapped to NEU 0, HAP 1, SAD 2, ANG 3, FEA 4, DIS 5."""
    if "-" in basename:
        num = int(basename.split("-")[2])
        if num in (2, 8):
            return None
        return _RAVDESS[num]
    return _RAVDESS[_CREMAD[basename.split("_")[2]]]


def stratified_split(labels: Sequence[int]) -> Tuple[List[int], List[int], List[int]]:
    """train2.py:399-413: 80 / 10 / 10 split, stratified, random_state 42 (scikit-learn, like the reference)."""
    from sklearn.model_selection import train_test_split
    indices = list(range(len(labels)))
    train, temp = train_test_split(indices, test_size=0.2, random_state=42, stratify=list(labels))
    val, test = train_test_split(temp, test_size=0.5, random_state=42, stratify=[labels[i] for i in temp])
    return train, val, test


def balanced_class_weights(train_labels: Sequence[int], boost_factor: float = 1.2) -> torch.Tensor:
    """train2.py:474-488: sklearn 'balanced' weights n / (k * count_c), Fear (4) and Disgust (5) boosted by 1.2."""
    y = np.asarray(train_labels)
    classes, counts = np.unique(y, return_counts=True)
    w = torch.tensor(len(y) / (len(classes) * counts.astype(np.float64)), dtype=torch.float32)
    w[4] = w[4] * boost_factor
    w[5] = w[5] * boost_factor
    return w


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceFeatureSet:
    """All samples of ``load_data`` (train2.py:312-360) resident in HBM.

    ``video_features``: list of ``[T_i, Dv]`` arrays/tensors; ``audio_features``: list of ``[Da]``; ``labels``: ints.
    ``normalize=True`` computes the global statistics of train2.py:430-441 on the device (mean, unbiased std + 1e-6
    over all frames / all samples); the z-score of train2.py:443-447 is applied inside the collate kernel, so the raw
    features are stored once.  ``normalize=False`` is the train.py behaviour (features used as they are).
    ``store_dtype=torch.bfloat16`` keeps the z-scored features in bf16 instead (normalised and rounded once): half the
    memory and half the bytes per batch, bit-identical to bf16 batches of the fp32-resident set.
    """

    def __init__(self, video_features: Sequence, audio_features: Sequence, labels: Sequence[int], device="cuda",
                 normalize: bool = True, store_dtype: torch.dtype = torch.float32):
        if not (len(video_features) == len(audio_features) == len(labels)) or len(labels) == 0:
            raise ValueError("need the same, non-zero number of video, audio and label entries")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise MmerError("DeviceFeatureSet lives on a CUDA device (there is no CPU path)")
        vids = [torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v, dtype=torch.float32) for v in video_features]
        auds = [torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a, dtype=torch.float32) for a in audio_features]
        self.Dv = int(vids[0].shape[1])
        self.Da = int(auds[0].shape[0])
        if any(v.dim() != 2 or v.shape[1] != self.Dv for v in vids) or any(a.shape != (self.Da,) for a in auds):
            raise ValueError("video features must be [T_i, Dv] and audio features [Da] with constant Dv, Da")
        self.lengths = [int(v.shape[0]) for v in vids]
        self.lengths_np = np.asarray(self.lengths, dtype=np.int64)
        self.max_chunks = max(self.lengths)                                        # train2.py:455
        off = np.zeros(len(vids) + 1, dtype=np.int64)
        np.cumsum(self.lengths, out=off[1:])
        self.n = len(vids)
        self.device = dev
        self.labels_host = [int(x) for x in labels]
        self.frames = torch.cat(vids, dim=0).to(dev).contiguous()                  # [total_frames, Dv], raw
        self.audio = torch.stack(auds, dim=0).to(dev).contiguous()                 # [N, Da], raw
        self.offsets = torch.from_numpy(off).to(dev)
        self.labels = torch.tensor(self.labels_host, dtype=torch.long, device=dev)
        self.video_mean = self.video_std = self.audio_mean = self.audio_std = None
        if normalize:
            self.video_mean, self.video_std = self._stats(self.frames)
            self.audio_mean, self.audio_std = self._stats(self.audio)
        if store_dtype not in (torch.float32, torch.bfloat16):
            raise MmerError(f"unsupported store dtype {store_dtype}")
        self.store_dtype = store_dtype
        if store_dtype == torch.bfloat16:
            # z-score and round ONCE; batches are then pure 2-byte gathers (and only ever bf16)
            self.frames = self._normalized_bf16(self.frames, self.video_mean, self.video_std)
            self.audio = self._normalized_bf16(self.audio, self.audio_mean, self.audio_std)

    @staticmethod
    def _stats(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        R, D = x.shape
        mean = torch.empty(D, device=x.device, dtype=torch.float32)
        std = torch.empty(D, device=x.device, dtype=torch.float32)
        scratch = torch.empty(2 * D, device=x.device, dtype=torch.float64)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().mmer_feature_stats(x.data_ptr(), R, D, 1e-6, mean.data_ptr(), std.data_ptr(),
                                                      scratch.data_ptr(), _stream()), "mmer_feature_stats")
        return mean, std

    @staticmethod
    def _normalized_bf16(x: torch.Tensor, mean: Optional[torch.Tensor], std: Optional[torch.Tensor]) -> torch.Tensor:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().mmer_normalize_rows(x.data_ptr(), mean.data_ptr() if mean is not None else None,
                                                       std.data_ptr() if std is not None else None, out.data_ptr(),
                                                       x.shape[0], x.shape[1], _stream()), "mmer_normalize_rows")
        return out

    def __len__(self) -> int:
        return self.n

    def collate(self, indices: Sequence[int], dtype: torch.dtype = torch.float32):
        """The reference's ``collate_fn`` (train2.py:418-440) for the samples ``indices``:
        ``(videos_padded [B, T_max, Dv], audios_stacked [B, Da], labels_tensor [B], masks_padded [B, T_max] bool)``,
        already on the device and in ``dtype`` (float32 like the reference, or bfloat16 for the tensor-core step)."""
        idx_host = np.ascontiguousarray(np.asarray(indices, dtype=np.int64).reshape(-1))
        B = int(idx_host.size)
        if B == 0:
            raise ValueError("empty batch")
        if int(idx_host.min()) < 0 or int(idx_host.max()) >= self.n:
            raise IndexError("sample index out of range")
        if dtype not in (torch.float32, torch.bfloat16):
            raise MmerError(f"unsupported batch dtype {dtype}")
        tmax = int(self.lengths_np[idx_host].max())     # the only thing the host has to know about the batch
        dev = self.device
        idx = torch.from_numpy(idx_host).to(dev)
        video = torch.empty((B, tmax, self.Dv), device=dev, dtype=dtype)
        audio = torch.empty((B, self.Da), device=dev, dtype=dtype)
        labels = torch.empty(B, device=dev, dtype=torch.long)
        mask = torch.empty((B, tmax), device=dev, dtype=torch.bool)
        p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        if self.store_dtype == torch.bfloat16:
            if dtype != torch.bfloat16:
                raise MmerError("a bf16-resident feature set produces bf16 batches only")
            with torch.cuda.device(dev):
                _lib.check(_lib.load().mmer_collate_bf16(
                    self.frames.data_ptr(), self.offsets.data_ptr(), self.audio.data_ptr(), self.labels.data_ptr(),
                    idx.data_ptr(), video.data_ptr(), audio.data_ptr(), labels.data_ptr(), mask.data_ptr(), B, tmax,
                    self.Dv, self.Da, _stream()), "mmer_collate_bf16")
            return video, audio, labels, mask
        with torch.cuda.device(dev):
            _lib.check(_lib.load().mmer_collate(
                self.frames.data_ptr(), self.offsets.data_ptr(), self.audio.data_ptr(), self.labels.data_ptr(), idx.data_ptr(),
                p(self.video_mean), p(self.video_std), p(self.audio_mean), p(self.audio_std), video.data_ptr(),
                audio.data_ptr(), labels.data_ptr(), mask.data_ptr(), B, tmax, self.Dv, self.Da,
                BF16 if dtype == torch.bfloat16 else F32, _stream()), "mmer_collate")
        return video, audio, labels, mask

    def loader(self, indices: Sequence[int], batch_size: int = 32, shuffle: bool = False,
               dtype: torch.dtype = torch.float32) -> "DeviceLoader":
        return DeviceLoader(self, indices, batch_size, shuffle, dtype)


class DeviceLoader:
    """Iterates like ``DataLoader([dataset[i] for i in indices], batch_size, shuffle, collate_fn)`` (train2.py:443-462):
    same batch boundaries (the last batch may be short), same tuple order, and with ``shuffle=True`` the same permutation
    torch's ``DataLoader`` + ``RandomSampler`` would draw from the global RNG (the loader's base seed, then the sampler's
    seed from ``torch.empty((), int64).random_()``, then ``torch.randperm(n, generator)``), so that a seeded run visits the samples in the reference's order."""

    def __init__(self, data: DeviceFeatureSet, indices: Sequence[int], batch_size: int, shuffle: bool, dtype: torch.dtype):
        if batch_size < 1:
            raise ValueError("batch_size must be positive")
        self.data, self.indices = data, np.asarray([int(i) for i in indices], dtype=np.int64)
        self.batch_size, self.shuffle, self.dtype = batch_size, shuffle, dtype

    def __len__(self) -> int:
        return (len(self.indices) + self.batch_size - 1) // self.batch_size

    @property
    def dataset(self):
        """``len(loader.dataset)`` as the reference's main() prints it (train2.py:956-960): the loader's sample indices."""
        return self.indices

    def __iter__(self) -> Iterator:
        order = np.arange(len(self.indices))
        # every DataLoader iterator -- shuffled or not -- first draws its workers' base seed from the global RNG
        # (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__); the validation and test passes of an epoch
        # therefore advance the RNG that the NEXT epoch's training shuffle draws from (train2.py:564, 596, 658)
        torch.empty((), dtype=torch.int64).random_()
        if self.shuffle:
            # ... then RandomSampler its own seed
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
            order = torch.randperm(len(self.indices), generator=g).numpy()
        for i in range(0, len(order), self.batch_size):
            yield self.data.collate(self.indices[order[i:i + self.batch_size]], self.dtype)
