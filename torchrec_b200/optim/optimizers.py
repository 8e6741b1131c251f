"""Shell optimizer classes used only as *tags* for ``apply_optimizer_in_backward`` — the real
update runs inside the fused embedding backward kernel (reference torchrec/optim/optimizers.py)."""
from typing import Any, Dict, Iterable, Iterator, Tuple

import torch
from torch import nn
from torch.optim.optimizer import Optimizer


def in_backward_optimizer_filter(named_parameters: Iterator[Tuple[str, nn.Parameter]], include: bool = False) -> Iterator[Tuple[str, nn.Parameter]]:
    """Yield parameters that are NOT (or, with ``include=True``, that ARE) optimized in backward."""
    for fqn, param in named_parameters:
        if hasattr(param, "_in_backward_optimizers") == include:
            yield fqn, param


class _Shell(Optimizer):
    _defaults: Dict[str, Any] = {}

    def __init__(self, params: Iterable[torch.nn.Parameter], **kwargs: Any) -> None:
        d = dict(self._defaults)
        d.update(kwargs)
        self._params = params
        super().__init__(params, d)

    @torch.no_grad()
    def step(self, closure: Any = None) -> torch.Tensor:
        raise NotImplementedError(f"{type(self).__name__} is a tag for apply_optimizer_in_backward; the update is fused into the embedding kernel")


class SGD(_Shell):
    _defaults = dict(lr=1e-2, momentum=0.0, weight_decay=0.0)


class LarsSGD(_Shell):
    _defaults = dict(lr=1e-2, momentum=0.9, eps=1e-8, eta=0.001, weight_decay=0.0)


class LAMB(_Shell):
    _defaults = dict(lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)


class PartialRowWiseLAMB(_Shell):
    _defaults = dict(lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)


class Adam(_Shell):
    _defaults = dict(lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)


class PartialRowWiseAdam(_Shell):
    _defaults = dict(lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)


class Adagrad(_Shell):
    _defaults = dict(lr=1e-2, eps=1e-8, weight_decay=0.0)
