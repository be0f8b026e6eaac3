"""mmer_b200: B200-native (sm_100a) drop-in for the fusion classifier step of
EvanZJ/multi-modal-emotion-recognition.  Import as ``mmer_b200`` (alias package at the repo
root; this directory's name is not a valid Python identifier).

Public surface (reference names):
    FocalLoss, CrossModalFusion, EmotionClassifier, MultimodalEmotionModel   (train2.py variant)
    v1.CrossModalFusion, v1.EmotionClassifier, v1.MultimodalEmotionModel     (train.py variant)
plus FusedAdam, FusedTrainStep, WeightedCrossEntropyLoss and the raw ``ops``.
"""
from . import _lib, ops  # noqa: F401
from ._lib import MmerError  # noqa: F401
from .modules import (CrossModalFusion, EmotionClassifier, FocalLoss, MultimodalEmotionModel,  # noqa: F401
                      WeightedCrossEntropyLoss)
from . import modules_v1 as v1  # noqa: F401
from .trainer import FusedAdam, FusedTrainStep  # noqa: F401
from .attribution import aggregate_importances, compute_attributions  # noqa: F401
from .data import DeviceFeatureSet, DeviceLoader  # noqa: F401
from . import data  # noqa: F401
from .evaluation import EvalAccumulator, metrics_from_confusion  # noqa: F401
from .inference import GraphedInference, ServingForward  # noqa: F401
from .dpcheck import dp_selfcheck, weights_checksum  # noqa: F401
from .training import HostBatchStager, bind_host_to_gpu, list_feature_pairs, load_data, train_model  # noqa: F401

__all__ = ["FocalLoss", "WeightedCrossEntropyLoss", "CrossModalFusion", "EmotionClassifier",
           "MultimodalEmotionModel", "v1", "FusedAdam", "FusedTrainStep", "ops", "MmerError",
           "compute_attributions", "aggregate_importances",
           "DeviceFeatureSet", "DeviceLoader", "data", "EvalAccumulator", "metrics_from_confusion", "GraphedInference", "ServingForward",
           "dp_selfcheck", "weights_checksum", "load_data", "train_model", "HostBatchStager", "list_feature_pairs", "bind_host_to_gpu"]
