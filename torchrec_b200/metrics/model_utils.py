"""Turn model outputs into per-task metric inputs (reference metrics/model_utils.py:20-195)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from .rec_metric import RecTaskInfo


def session_ids_to_tensor(session_ids: List[str], device: Optional[torch.device] = None) -> torch.Tensor:
    """Consecutive-run encoding of session id strings: equal neighbours get the same integer (what session-level metrics group by)."""
    out, cur = [], -1
    prev: Optional[str] = None
    for s in session_ids:
        if s != prev:
            cur += 1
            prev = s
        out.append(cur)
    return torch.tensor(out, dtype=torch.int64, device=device)


def is_empty_signals(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> bool:
    return torch.numel(labels) <= 0 and torch.numel(predictions) <= 0 and torch.numel(weights) <= 0


def parse_model_outputs(label_name: str, prediction_name: str, weight_name: str, model_out: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """(labels, predictions, weights) of one task; binary tasks are squeezed to 1-D, multiclass keeps ``[batch, classes]`` predictions."""
    labels = model_out[label_name]
    predictions = model_out[prediction_name]
    weights = model_out.get(weight_name) if weight_name else None
    if labels.dim() == predictions.dim():
        labels, predictions = labels.squeeze(), predictions.squeeze()
        assert labels.size() == predictions.size(), f"labels {tuple(labels.shape)} and predictions {tuple(predictions.shape)} disagree"
    else:
        assert predictions.dim() == labels.dim() + 1 and predictions.shape[0] == labels.shape[0], "multiclass: predictions must be [batch, classes]"
    if weights is not None:
        weights = weights.squeeze()
        assert weights.shape[0] == labels.shape[0] if weights.dim() else True
    return labels, predictions, weights


def parse_required_inputs(model_out: Dict[str, torch.Tensor], required_inputs_list: List[str], ndcg_transform_input: bool = False, device: Optional[torch.device] = None) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    for name in required_inputs_list:
        v = model_out[name]
        if ndcg_transform_input and isinstance(v, list):
            v = session_ids_to_tensor(v, device=device)
        out[name] = v.squeeze() if isinstance(v, torch.Tensor) else v
    return out


def parse_task_model_outputs(tasks: List[RecTaskInfo], model_out: Dict[str, torch.Tensor], required_inputs_list: Optional[List[str]] = None):
    from .metric_module import parse_task_model_outputs as _impl

    return _impl(tasks, model_out, required_inputs_list)
