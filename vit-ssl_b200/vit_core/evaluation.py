"""Evaluation-path helpers on the GPU.

The reference's evaluators copy `model.inference_forward` features to the CPU and hand them to
scikit-learn (evaluators/unsupervised_evaluators/evaluator_utils.py:8-22,
evaluators/unsupervised_evaluator.py:38-66). `knn_predict` is the same classifier —
`KNeighborsClassifier(n_neighbors=k, metric="cosine")`, uniform vote — run where the features
already are; `run_knn_evaluation` mirrors the reference function's signature and result dict so an
evaluator can swap it in with one import.
"""
from __future__ import annotations

import torch

from ._backend import ops


def knn_predict(train_features: torch.Tensor, train_labels: torch.Tensor, val_features: torch.Tensor,
                n_neighbors: int, num_classes: int | None = None) -> torch.Tensor:
    """Predicted labels (int64, on the features' device) of `val_features` by cosine k-NN."""
    dev = val_features.device if val_features.is_cuda else torch.device("cuda")
    labels = torch.as_tensor(train_labels).to(dev)
    if num_classes is None:
        num_classes = int(labels.max().item()) + 1
    pred, _ = ops.knn_cosine(val_features.to(dev), train_features.to(dev), labels, int(n_neighbors), int(num_classes))
    return pred.long()


def run_knn_evaluation(train_features, train_labels, val_features, val_labels, num_classes):
    """Same contract as evaluators/unsupervised_evaluator.py:38-66 (`accuracy`, `predictions`, ...)."""
    preds = knn_predict(torch.as_tensor(train_features), torch.as_tensor(train_labels), torch.as_tensor(val_features),
                        num_classes, None)
    vl = torch.as_tensor(val_labels).to(preds.device)
    accuracy = (preds == vl).float().mean().item()
    return {"method": "knn", "accuracy": accuracy, "predictions": preds.cpu().numpy(), "num_neighbors": num_classes}
