"""The callers on either side of the hot path, as product code:

* ``load_data(video_feat_dir, audio_feat_dir, batch_size)``   -- train2.py:299-488 (and train.py:144-246 with
  ``normalize=False``): the ``.npy`` feature directories -> HBM-resident feature set + three device loaders.
* ``HostBatchStager``                                          -- pinned, ring-buffered host -> device staging for callers
  whose batches live in host memory (the reference's ``.to(device)`` x4 per batch, train2.py:565-568, made asynchronous).
* ``train_model(model, train_loader, val_loader, test_loader, class_weights, ...)`` -- train2.py:495-774: the epoch loop
  (fused training step, device-side evaluation bookkeeping, ReduceLROnPlateau, early stopping, best-state bookkeeping,
  ``results_*.json`` and ``.pth`` writers) with the reference's names, argument meaning, file names and JSON layout.

Same return values as the reference: ``load_data`` returns ``(train_loader, val_loader, test_loader, max_chunks,
class_weights)``; ``train_model`` returns ``None`` in the reference -- here it returns the dict it also writes to disk
(a superset that callers of the reference simply ignore).
"""
from __future__ import annotations

import glob
import json
import os
from concurrent.futures import ThreadPoolExecutor
from datetime import datetime
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import MmerError
from .data import DeviceFeatureSet, DeviceLoader, balanced_class_weights, label_from_filename, stratified_split
from .evaluation import EvalAccumulator
from .trainer import FusedTrainStep

__all__ = ["load_data", "list_feature_pairs", "HostBatchStager", "train_model", "bind_host_to_gpu"]


# --------------------------------------------------------------------------------------------- N2: load_data
def list_feature_pairs(video_feat_dir: str, audio_feat_dir: str, pair_by: str = "zip") -> List[Tuple[str, str]]:
    """(video file, audio file) pairs of the two feature directories.

    ``pair_by="zip"`` is what the reference does (train2.py:318-325, train.py:149-152): ``zip`` of the two SORTED globs,
    i.e. the i-th video file goes with the i-th audio file whatever their names.  On the reference's own data one audio
    file (``1076_MTI_SAD_XX``) has no video twin, which silently misaligns every later pair (2694 of them).
    ``pair_by="stem"`` pairs by file name and drops files without a twin -- what the author meant.  The default keeps
    the reference's behaviour so that a drop-in run reproduces its numbers; use ``"stem"`` for correct training."""
    vids = sorted(glob.glob(os.path.join(video_feat_dir, "*.npy")))
    auds = sorted(glob.glob(os.path.join(audio_feat_dir, "*.npy")))
    if pair_by == "zip":
        return list(zip(vids, auds))
    if pair_by == "stem":
        by_stem = {os.path.basename(a): a for a in auds}
        return [(v, by_stem[os.path.basename(v)]) for v in vids if os.path.basename(v) in by_stem]
    raise ValueError("pair_by must be 'zip' (reference behaviour) or 'stem'")


def load_data(video_feat_dir: str, audio_feat_dir: str, batch_size: int = 32, *, device="cuda",
              dtype: torch.dtype = torch.float32, store_dtype: torch.dtype = torch.float32, normalize: bool = True,
              pair_by: str = "zip", io_threads: int = 16, verbose: bool = True):
    """train2.py:299-488.  Reads every ``.npy`` pair (label from the VIDEO file name; RAVDESS classes 02 / 08 skipped),
    uploads the whole set once, computes the global z-score statistics on the device (``normalize=False``: train.py's
    variant, features as they are), splits 80 / 10 / 10 stratified with ``random_state=42`` and returns loaders that
    yield ``(videos_padded, audios_stacked, labels_tensor, masks_padded)`` already on the device.

    Returns ``(train_loader, val_loader, test_loader, max_chunks, class_weights)`` like the reference."""
    pairs = list_feature_pairs(video_feat_dir, audio_feat_dir, pair_by)
    if verbose:
        print("Example video/audio file pairs:")
        for v, a in pairs[:10]:
            print(os.path.basename(v), "<--->", os.path.basename(a))
    keep, labels = [], []
    for v, a in pairs:
        lab = label_from_filename(os.path.basename(v))
        if lab is None:
            continue
        keep.append((v, a))
        labels.append(lab)
    if not keep:
        raise MmerError(f"no usable feature files under {video_feat_dir!r} / {audio_feat_dir!r}")

    def read(pair):
        return np.load(pair[0]).astype(np.float32), np.load(pair[1]).astype(np.float32)      # train2.py:355-356

    with ThreadPoolExecutor(max_workers=max(1, io_threads)) as ex:
        loaded = list(ex.map(read, keep))
    data = DeviceFeatureSet([x[0] for x in loaded], [x[1] for x in loaded], labels, device=device, normalize=normalize,
                            store_dtype=store_dtype)
    if verbose:
        print(f"Maximum number of video chunks: {data.max_chunks}")
    train_idx, val_idx, test_idx = stratified_split(labels)
    train_loader = data.loader(train_idx, batch_size, shuffle=True, dtype=dtype)
    val_loader = data.loader(val_idx, batch_size, shuffle=False, dtype=dtype)
    test_loader = data.loader(test_idx, batch_size, shuffle=False, dtype=dtype)
    train_labels = [labels[i] for i in train_idx]
    if verbose:
        from collections import Counter
        print("Train label distribution:", Counter(train_labels))
        print("Val label distribution:", Counter(labels[i] for i in val_idx))
        print("Test label distribution:", Counter(labels[i] for i in test_idx))
    return train_loader, val_loader, test_loader, data.max_chunks, balanced_class_weights(train_labels)


# --------------------------------------------------------------------------------------------- N2: host staging
def bind_host_to_gpu(device_index: int) -> Optional[list]:
    """Pin the calling process to the CPU cores (and thereby, through first-touch, its future pinned host buffers to the
    memory) of the NUMA node the GPU hangs off.  With eight ranks streaming 27 GB/s each from host memory, buffers that
    all sit on one socket make the inter-socket link the bottleneck of the host -> device stream.  Call it first thing
    in each rank, BEFORE allocating pinned memory.  Returns the CPU list, or None when NVML / the affinity call is not
    available (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                idx = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c < n_cpu]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class HostBatchStager:
    """Asynchronous host -> device staging of batches that live in HOST memory.

    The reference copies every batch synchronously from pageable memory (``videos.to(device)`` ..., train2.py:565-568).
    Here a ring of ``depth`` device slots is filled on a side stream from pinned host memory while earlier batches are
    being consumed: ``for batch in stager.pipeline(host_batches): step.step(*batch)`` sees device tensors whose copies
    were issued ``depth - 1`` batches ahead.  Tensors that are not pinned are first copied into the slot's own pinned
    staging buffer (a host memcpy; pin the source once to avoid it).  A yielded batch stays valid until the consumer
    asks for the next one; ``None`` entries (e.g. an absent mask) pass through."""

    def __init__(self, device="cuda", depth: int = 3):
        if depth < 2:
            raise ValueError("depth must be at least 2 (one slot in use, one being filled)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise MmerError("HostBatchStager stages onto a CUDA device")
        self.depth = depth
        self.stream = torch.cuda.Stream(device=self.device)
        self._dev: List[Optional[list]] = [None] * depth
        self._pin: List[Optional[list]] = [None] * depth
        self._ready = [torch.cuda.Event() for _ in range(depth)]
        self._freed = [torch.cuda.Event() for _ in range(depth)]
        self.bytes_staged = 0

    def _buffers(self, slot: int, batch: Sequence[Optional[torch.Tensor]]):
        dev, pin = self._dev[slot], self._pin[slot]
        ok = dev is not None and len(dev) == len(batch) and all(
            (t is None and d is None) or (t is not None and d is not None and d.shape == t.shape and d.dtype == t.dtype)
            for t, d in zip(batch, dev))
        if not ok:
            dev = [None if t is None else torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch]
            pin = [None] * len(batch)
            self._dev[slot], self._pin[slot] = dev, pin
            # fresh blocks come from the consumer stream's pool: whatever that stream did with them must finish before
            # the copy stream writes them
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.stream.wait_event(ev)
        return dev, pin

    def _stage(self, slot: int, batch: Sequence[Optional[torch.Tensor]]) -> None:
        dev, pin = self._buffers(slot, batch)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self._freed[slot])          # the consumer is done with what this slot held
            for i, t in enumerate(batch):
                if t is None:
                    continue
                if t.is_cuda:
                    dev[i].copy_(t, non_blocking=True)
                    continue
                src = t.contiguous()
                if not src.is_pinned():
                    if pin[i] is None or pin[i].shape != src.shape or pin[i].dtype != src.dtype:
                        pin[i] = torch.empty(src.shape, dtype=src.dtype).pin_memory()
                    # the previous copy out of this pinned buffer was ordered before _freed[slot]; make the HOST wait too
                    self._freed[slot].synchronize()
                    pin[i].copy_(src)
                    src = pin[i]
                dev[i].copy_(src, non_blocking=True)
                self.bytes_staged += src.numel() * src.element_size()
            self._ready[slot].record(self.stream)

    def pipeline(self, batches: Iterable[Sequence[Optional[torch.Tensor]]]) -> Iterator[tuple]:
        main = torch.cuda.current_stream(self.device)
        for s in range(self.depth):
            self._freed[s].record(main)
        it = iter(batches)
        staged = 0          # batches handed to the copy stream
        done = 0            # batches yielded
        pending = []
        for _ in range(self.depth - 1):
            nxt = next(it, None)
            if nxt is None:
                break
            self._stage(staged % self.depth, nxt)
            pending.append(staged)
            staged += 1
        while pending:
            k = pending.pop(0)
            slot = k % self.depth
            nxt = next(it, None)
            if nxt is not None:                                 # keep depth - 1 copies in flight
                self._stage(staged % self.depth, nxt)
                pending.append(staged)
                staged += 1
            main = torch.cuda.current_stream(self.device)
            main.wait_event(self._ready[slot])
            yield tuple(self._dev[slot])
            self._freed[slot].record(torch.cuda.current_stream(self.device))
            done += 1


# --------------------------------------------------------------------------------------------- N3: train_model
def _evaluate(model, loader, criterion_kind: str, class_weights, acc: EvalAccumulator, with_loss: bool):
    from .modules import FocalLoss, WeightedCrossEntropyLoss
    crit = None
    if with_loss:
        crit = WeightedCrossEntropyLoss(class_weights) if criterion_kind == "wce" else FocalLoss(2.0, class_weights)
    acc.reset()
    with torch.no_grad():
        for videos, audios, labels, masks in loader:
            probs, logits, _ = model(videos, audios, mask=masks)
            acc.update(probs, labels, crit(logits, labels) if crit is not None else None)
    return acc.result()


def train_model(model, train_loader, val_loader, test_loader, class_weights: torch.Tensor, num_epochs: int = 100,
                lr: float = 1e-4, weight_decay: float = 1e-4, patience: int = 8, batch_size: int = 128,
                device: str = "cuda", *, loss: str = "wce", compute_dtype: torch.dtype = torch.float32,
                out_dir: Optional[str] = "training_runs_2", copy_best_state: bool = False, verbose: bool = True,
                process_group=None):
    """train2.py:495-774 on the fused device path.

    Per epoch: one ``FusedTrainStep.step`` per training batch (class-weighted cross-entropy, ``clip_grad_norm_`` 1.0, Adam
    with coupled weight decay; train2.py:523-525, 570-578) with the running loss kept on the device; validation and test
    passes through ``EvalAccumulator`` (no per-batch ``.item()`` / ``.cpu()``); ``ReduceLROnPlateau(mode="min", factor=0.3,
    patience=20)`` on the validation loss (train2.py:526, 614); the reference's early-stopping rule -- stop after
    ``patience`` epochs in which the validation loss did not improve by 1e-4 over the PREVIOUS epoch (train2.py:622-631;
    the epoch that triggers the stop is not logged, like the reference's ``break`` before its metrics block).

    ``copy_best_state=False`` reproduces the reference's bookkeeping exactly: ``model.state_dict().copy()`` is a shallow
    copy whose tensors alias the live parameters (train2.py:619), so the "best" checkpoint and the final confusion matrix
    are those of the LAST weights.  ``copy_best_state=True`` snapshots the weights of the best epoch instead.

    Batches may come from ``load_data`` above (device tensors) or from any iterable of host tensors (e.g. the
    reference's own ``DataLoader``): host batches are staged through ``HostBatchStager``.
    Writes ``results_bs{B}_ep{E}_lr{lr}_{timestamp}.json``, ``best_model_*.pth`` and ``final_model_*.pth`` under
    ``out_dir`` (None: nothing is written) with the reference's layout, and returns the same dictionary plus
    ``confusion_matrix`` and the file paths."""
    if not torch.cuda.is_available():
        raise MmerError("train_model runs on a CUDA device (no CPU fallback)")
    from torch.optim.lr_scheduler import ReduceLROnPlateau
    dev = torch.device(device)
    model.to(dev)
    class_weights = class_weights.to(dev)
    step = FusedTrainStep(model, lr=lr, weight_decay=weight_decay, loss=loss, gamma=2.0, alpha=class_weights,
                          clip_grad_norm=1.0, compute_dtype=compute_dtype, process_group=process_group)
    if compute_dtype == torch.bfloat16:
        model.compute_dtype = torch.bfloat16
    scheduler = ReduceLROnPlateau(step.opt, mode="min", factor=0.3, patience=20)
    hyperparameters = {                                            # train2.py:529-548, same keys and values
        "num_epochs": num_epochs, "lr": lr, "weight_decay": weight_decay, "patience": patience, "batch_size": batch_size,
        "device": str(device),
        "video_dim": model.fusion.video_proj.in_features, "audio_dim": model.fusion.audio_proj.in_features,
        "fused_dim": model.fusion.video_proj.out_features, "num_classes": model.classifier.net[-1].out_features,
        "max_seq_len": model.fusion.pos_embed.size(1), "fusion_dropout": model.fusion.dropout,
        "classifier_dropout": model.classifier.dropout, "num_layers": model.fusion.num_layers,
        "num_heads": model.fusion.num_heads, "scheduler_factor": 0.3, "scheduler_patience": 5, "focal_gamma": 2.0,
    }
    n_cls = hyperparameters["num_classes"]
    acc = EvalAccumulator(n_cls, device=dev)
    stager = HostBatchStager(dev)

    def on_device(loader):
        # (no peeking: iterating a shuffled loader draws from the global RNG exactly once per epoch, like the reference)
        return iter(loader) if isinstance(loader, DeviceLoader) else stager.pipeline(loader)

    results = []
    best_val_loss, best_model_state, best_epoch = float("inf"), None, 0
    epochs_without_improvement, previous_val_loss = 0, float("inf")
    train_loss_sum = torch.zeros(1, device=dev, dtype=torch.float32)
    for epoch in range(num_epochs):
        model.train()
        train_loss_sum.zero_()
        n_batches = 0
        for videos, audios, labels, masks in on_device(train_loader):
            l, _ = step.step(videos, audios, masks, labels)
            train_loss_sum += l                                    # stays on the device (train2.py:579 syncs per step)
            n_batches += 1
        avg_train_loss = float(train_loss_sum.item()) / max(n_batches, 1)

        model.eval()
        val = _evaluate(model, on_device(val_loader), loss, class_weights, acc, True)
        avg_val_loss, val_acc = val["avg_loss"], val["accuracy"]
        scheduler.step(avg_val_loss)
        if avg_val_loss < best_val_loss:
            best_val_loss, best_epoch = avg_val_loss, epoch + 1
            sd = model.state_dict()
            best_model_state = {k: v.detach().clone() for k, v in sd.items()} if copy_best_state else sd.copy()
        if previous_val_loss - avg_val_loss < 1e-4:
            epochs_without_improvement += 1
            if epochs_without_improvement >= patience:
                if verbose:
                    print(f"Early stopping at epoch {epoch + 1}")
                break
        else:
            epochs_without_improvement = 0
        previous_val_loss = avg_val_loss
        test = _evaluate(model, on_device(test_loader), loss, class_weights, acc, False)
        if verbose:
            print(f"Epoch {epoch + 1}/{num_epochs}, Train Loss: {avg_train_loss:.4f}, Val Loss: {avg_val_loss:.4f}, "
                  f"Val Acc: {val_acc:.2f}%")
            print(f"Val Macro P/R/F1: {val['macro_precision']:.4f}/{val['macro_recall']:.4f}/{val['macro_f1']:.4f}, "
                  f"Micro P/R/F1: {val['micro_precision']:.4f}/{val['micro_recall']:.4f}/{val['micro_f1']:.4f}")
            print(f"Test Acc: {test['accuracy']:.2f}%, Test Macro P/R/F1: {test['macro_precision']:.4f}/"
                  f"{test['macro_recall']:.4f}/{test['macro_f1']:.4f}, Micro P/R/F1: {test['micro_precision']:.4f}/"
                  f"{test['micro_recall']:.4f}/{test['micro_f1']:.4f}")
        results.append({
            "epoch": epoch + 1, "train_loss": avg_train_loss, "val_loss": avg_val_loss, "val_acc": val_acc,
            "val_macro_precision": val["macro_precision"], "val_macro_recall": val["macro_recall"],
            "val_macro_f1": val["macro_f1"], "val_micro_precision": val["micro_precision"],
            "val_micro_recall": val["micro_recall"], "val_micro_f1": val["micro_f1"],
            "test_acc": test["accuracy"], "test_macro_precision": test["macro_precision"],
            "test_macro_recall": test["macro_recall"], "test_macro_f1": test["macro_f1"],
            "test_micro_precision": test["micro_precision"], "test_micro_recall": test["micro_recall"],
            "test_micro_f1": test["micro_f1"],
        })

    cm = None
    if best_model_state is not None:                               # train2.py:716-745
        if verbose:
            print("\nEvaluating BEST model on test set for confusion matrix ...")
        model.load_state_dict(best_model_state)
        model.to(dev)
        model.eval()
        _evaluate(model, on_device(test_loader), loss, class_weights, acc, False)
        cm = acc.confusion_matrix()
        if verbose:
            print("Confusion matrix (rows = true, cols = pred):")
            print(cm)

    out = {"training_progress": results, "best_model": {"epoch": best_epoch}, "hyperparameters": hyperparameters}
    paths = {}
    if out_dir is not None:                                        # train2.py:750-773
        os.makedirs(out_dir, exist_ok=True)
        timestamp = datetime.now().strftime("%Y%m%d_%H%M%S")
        tag = f"bs{batch_size}_ep{num_epochs}_lr{lr}_{timestamp}"
        paths["results"] = os.path.join(out_dir, f"results_{tag}.json")
        with open(paths["results"], "w") as f:
            json.dump(out, f, indent=4)
        paths["best_model"] = os.path.join(out_dir, f"best_model_{tag}.pth")
        torch.save(best_model_state, paths["best_model"])
        paths["final_model"] = os.path.join(out_dir, f"final_model_{tag}.pth")
        torch.save(model.state_dict(), paths["final_model"])
        if verbose:
            print(f"Training results saved to {paths['results']}")
            print(f"Best model (epoch {best_epoch}, val_loss {best_val_loss:.4f}) saved to {paths['best_model']}")
            print(f"Final model saved to {paths['final_model']}")
    ret = dict(out)
    ret["confusion_matrix"] = cm
    ret["paths"] = paths
    ret["best_val_loss"] = best_val_loss
    return ret
