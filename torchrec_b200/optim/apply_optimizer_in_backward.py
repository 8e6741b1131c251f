"""``apply_optimizer_in_backward``: tag parameters with the optimizer that the embedding backward
kernel must fuse (reference torchrec/optim/apply_optimizer_in_backward.py:16)."""
from typing import Any, Dict, Iterable, Type

import torch


def apply_optimizer_in_backward(optimizer_class: Type[torch.optim.Optimizer], params: Iterable[torch.nn.Parameter], optimizer_kwargs: Dict[str, Any]) -> None:
    """Mark ``params`` so that sharders build tables whose backward applies ``optimizer_class`` with
    ``optimizer_kwargs`` in-kernel. Sets both the legacy ``_optimizer_class(es)`` /
    ``_optimizer_kwargs`` attributes and ``_in_backward_optimizers`` (used by
    ``in_backward_optimizer_filter``)."""
    for param in params:
        param._optimizer_class = optimizer_class  # type: ignore[attr-defined]
        param._optimizer_kwargs = dict(optimizer_kwargs)  # type: ignore[attr-defined]
        param._optimizer_classes = [optimizer_class]  # type: ignore[attr-defined]
        param._optimizer_kwargs_list = [dict(optimizer_kwargs)]  # type: ignore[attr-defined]
        if not param.is_meta and param.device.type != "meta":
            try:
                param._in_backward_optimizers = [optimizer_class([param], **optimizer_kwargs)]  # type: ignore[attr-defined]
            except Exception:
                param._in_backward_optimizers = [None]  # type: ignore[attr-defined]
        else:
            param._in_backward_optimizers = [None]  # type: ignore[attr-defined]
