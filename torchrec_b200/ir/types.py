"""Interfaces of the IR package (reference ``torchrec/ir/types.py``): what a per-module serializer implements."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Type

import torch
from torch import nn


class SerializerInterface:
    """``serialize(module) -> (json dict, children names)`` / ``deserialize(dict, device, unflatten children) -> module``."""

    module_cls: Type[nn.Module]

    @classmethod
    def serialize_to_dict(cls, module: nn.Module) -> Dict[str, Any]:
        raise NotImplementedError

    @classmethod
    def deserialize_from_dict(cls, d: Dict[str, Any], device: Optional[torch.device] = None, children: Optional[Dict[str, nn.Module]] = None) -> nn.Module:
        raise NotImplementedError

    @classmethod
    def children(cls, module: nn.Module) -> List[str]:
        return []
