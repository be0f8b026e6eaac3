"""Evaluation bookkeeping on the device: the reference's validation / test loop bodies without per-batch host syncs.

The reference (train2.py:586-667, 724-745) calls ``loss.item()``, ``(predicted == labels).sum().item()`` and two
``.cpu().numpy()`` copies for every batch, then hands the collected label lists to scikit-learn.  ``EvalAccumulator``
keeps a confusion matrix, the loss sum and the batch count in device memory (one small kernel per batch, no sync) and
derives everything the reference logs from the matrix at the end: ``avg_val_loss = total_val_loss / len(loader)``,
``val_acc = 100 * correct / total``, macro / micro precision, recall, F1 as
``precision_recall_fscore_support(..., average=..., zero_division=0)`` computes them, and
``confusion_matrix(labels, preds, labels=[0..C-1])``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import MmerError

__all__ = ["EvalAccumulator", "metrics_from_confusion"]


def metrics_from_confusion(conf: np.ndarray) -> Dict[str, float]:
    """What the reference logs per epoch, from a [C, C] confusion matrix (rows = true class).

    scikit-learn's ``precision_recall_fscore_support`` (train2.py:633-644) averages over the labels that occur in
    ``y_true`` or ``y_pred``; a class with neither a true nor a predicted sample does not enter the macro mean, and
    ``zero_division=0`` turns 0/0 into 0."""
    conf = np.asarray(conf, dtype=np.float64)
    tp = np.diag(conf)
    pred, true = conf.sum(axis=0), conf.sum(axis=1)
    total = conf.sum()
    present = (pred + true) > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        prec = np.where(pred > 0, tp / pred, 0.0)
        rec = np.where(true > 0, tp / true, 0.0)
        f1 = np.where(prec + rec > 0, 2 * prec * rec / (prec + rec), 0.0)
    micro = float(tp.sum() / total) if total > 0 else 0.0     # single-label: micro P = R = F1 = accuracy
    n = max(int(present.sum()), 1)
    return {"accuracy": 100.0 * micro, "total": int(total), "correct": int(tp.sum()),
            "macro_precision": float(prec[present].sum() / n), "macro_recall": float(rec[present].sum() / n),
            "macro_f1": float(f1[present].sum() / n),
            "micro_precision": micro, "micro_recall": micro, "micro_f1": micro}


class EvalAccumulator:
    """``acc = EvalAccumulator(6); for batch: acc.update(probs, labels, loss); out = acc.result()``."""

    def __init__(self, num_classes: int = 6, device="cuda", keep_predictions: bool = False):
        if not 1 <= num_classes <= 16:
            raise MmerError("EvalAccumulator supports 1..16 classes")
        self.C = num_classes
        self.device = torch.device(device)
        self.keep = keep_predictions
        self.reset()

    def reset(self) -> None:
        self.conf = torch.zeros(self.C * self.C, device=self.device, dtype=torch.int64)
        self.loss_sum = torch.zeros(1, device=self.device, dtype=torch.float32)
        self.batches = 0
        self.preds = []

    @torch.no_grad()
    def update(self, probs: torch.Tensor, labels: torch.Tensor, loss: Optional[torch.Tensor] = None) -> None:
        if not probs.is_cuda:
            raise MmerError("EvalAccumulator needs CUDA tensors (no CPU fallback)")
        if probs.dim() != 2 or probs.shape[1] != self.C or labels.shape != (probs.shape[0],):
            raise ValueError("expected probs [B, C] and labels [B]")
        p = probs.detach().float().contiguous()
        y = labels.detach().to(device=p.device, dtype=torch.long).contiguous()
        pred = torch.empty_like(y) if self.keep else None
        with torch.cuda.device(p.device):
            _lib.check(_lib.load().mmer_eval_accumulate(p.data_ptr(), y.data_ptr(), pred.data_ptr() if self.keep else None,
                                                        self.conf.data_ptr(), p.shape[0], self.C,
                                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       "mmer_eval_accumulate")
        if self.keep:
            self.preds.append(pred)
        if loss is not None:
            self.loss_sum += loss.detach().reshape(-1)[:1].float()      # stays on the device: no .item() per batch
        self.batches += 1

    def confusion_matrix(self) -> np.ndarray:
        return self.conf.view(self.C, self.C).cpu().numpy()

    def predictions(self) -> torch.Tensor:
        return torch.cat(self.preds) if self.preds else torch.empty(0, dtype=torch.long, device=self.device)

    def result(self) -> Dict[str, float]:
        """One device -> host copy for the whole epoch."""
        out = metrics_from_confusion(self.confusion_matrix())
        out["avg_loss"] = float(self.loss_sum.item()) / max(self.batches, 1)       # train2.py:609
        return out
