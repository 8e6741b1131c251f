"""KeyedTensor regrouping modules (reference torchrec/modules/regroup.py:37-301)."""
from typing import Dict, List, Optional, Tuple

import torch

from ..sparse.jagged_tensor import KeyedTensor, regroup_kts
from ..types import CacheMixin


def _build_plan(keyed_tensors: List[KeyedTensor], groups: List[List[str]]) -> List[torch.Tensor]:
    """Per group: the column indices into cat(values of all KTs, dim=1)."""
    base = 0
    where: Dict[str, Tuple[int, int]] = {}
    for kt in keyed_tensors:
        off = kt.offset_per_key()
        for i, k in enumerate(kt.keys()):
            if k not in where:
                where[k] = (base + off[i], kt.length_per_key()[i])
        base += off[-1]
    plans = []
    for g in groups:
        idx = [torch.arange(where[k][0], where[k][0] + where[k][1]) for k in g]
        plans.append(torch.cat(idx) if idx else torch.zeros(0, dtype=torch.long))
    return plans


class PermuteMultiEmbedding(torch.nn.Module):
    """Regroup columns of several pooled-embedding tensors with a plan computed once."""

    def __init__(self, groups: List[List[str]], multi_device: bool = False) -> None:
        super().__init__()
        self._groups = groups
        self._plans: Optional[List[torch.Tensor]] = None
        self._multi_device = multi_device  # inputs may arrive on another device than the one the plan was built on (inference over several devices)

    def init_tensors(self, keyed_tensors: List[KeyedTensor]) -> None:
        self._plans = [p.to(keyed_tensors[0].device()) for p in _build_plan(keyed_tensors, self._groups)]

    def forward(self, values: List[torch.Tensor]) -> List[torch.Tensor]:
        assert self._plans is not None, "call init_tensors first"
        cat = values[0] if len(values) == 1 else torch.cat(values, dim=1)
        if self._multi_device and self._plans and self._plans[0].device != cat.device:
            self._plans = [p.to(cat.device) for p in self._plans]
        return [cat.index_select(1, p) for p in self._plans]


class KTRegroupAsDict(torch.nn.Module, CacheMixin):
    """``KeyedTensor.regroup_as_dict`` with the permutation cached after the first batch."""

    def __init__(self, groups: List[List[str]], keys: List[str], emb_dtype: Optional[torch.dtype] = None, multi_device: bool = False) -> None:
        super().__init__()
        torch._C._log_api_usage_once(f"torchrec_b200.modules.{self.__class__.__name__}")
        assert len(groups) == len(keys), "Groups and keys should have same length"
        self._groups = groups
        self._keys = keys
        self._emb_dtype = emb_dtype
        self._is_inited = False
        self._dim: int = 1
        self._permute = PermuteMultiEmbedding(groups, multi_device)

    def forward(self, keyed_tensors: List[KeyedTensor]) -> Dict[str, torch.Tensor]:
        if not self._is_inited:
            assert len(keyed_tensors) > 0, "Empty list provided"
            self._dim = keyed_tensors[0].key_dim()
            self._permute.init_tensors(keyed_tensors)
            self._is_inited = True
        if self._dim == 1:
            vals = [kt.values() if self._emb_dtype is None else kt.values().to(self._emb_dtype) for kt in keyed_tensors]
            out = self._permute(vals)
        else:
            out = regroup_kts(keyed_tensors, self._groups)
        return dict(zip(self._keys, out))

    def clear_cache(self) -> None:
        self._is_inited = False
