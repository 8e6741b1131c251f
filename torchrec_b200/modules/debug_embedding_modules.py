"""Debug wrappers around the embedding collections (reference torchrec/modules/debug_embedding_modules.py:46,133): with ``debug_mode`` the
forward checks that every id is inside its table and every output is finite, and an identity autograd function on the outputs raises
when a NaN / Inf gradient ENTERS the embedding backward (i.e. names the module the bad gradient came through). The wrapped collection
is the attribute ``ec`` / ``ebc``, so sharders find and shard it like any other collection; outputs of a sharded collection (awaitables)
are waited for in debug mode."""
import logging
from typing import Dict, List, Optional, Tuple, Union

import torch
from torch import nn

from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor, KeyedTensor
from .embedding_configs import EmbeddingBagConfig, EmbeddingConfig
from .embedding_modules import EmbeddingBagCollection, EmbeddingCollection

logger = logging.getLogger(__name__)


class _GradCheck(torch.autograd.Function):
    """Identity in forward; in backward the incoming gradient must be finite."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, tag: str):
        ctx.tag = tag
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g: torch.Tensor) -> Tuple[torch.Tensor, None]:
        check = g.values() if getattr(g, "is_sparse", False) else g
        if not bool(torch.isfinite(check).all()):
            raise RuntimeError(f"NaN/Inf detected in gradient entering {ctx.tag}")
        return g, None


def _check_ids(features: KeyedJaggedTensor, hash_sizes: Dict[str, int]) -> None:
    for k, f in features.to_dict().items():
        if k in hash_sizes and f.values().numel():
            mx, mn = int(f.values().max()), int(f.values().min())
            if mn < 0 or mx >= hash_sizes[k]:
                raise ValueError(f"feature {k}: ids in [{mn}, {mx}] outside of the table range [0, {hash_sizes[k]})")


class _DebugBase(nn.Module):
    def __init__(self, debug_mode: bool) -> None:
        super().__init__()
        self.debug_mode = debug_mode

    def _wrap_tensor(self, t: torch.Tensor, tag: str) -> torch.Tensor:
        return _GradCheck.apply(t, tag) if self.debug_mode and t.requires_grad else t


class DebugEmbeddingCollection(_DebugBase):
    """``DebugEmbeddingCollection(tables, device, debug_mode)`` builds the collection; ``DebugEmbeddingCollection(ec=existing)`` wraps one."""

    def __init__(self, tables: Optional[Union[List[EmbeddingConfig], EmbeddingCollection]] = None, device: Optional[torch.device] = None, debug_mode: bool = False,
                 ec: Optional[EmbeddingCollection] = None) -> None:
        super().__init__(debug_mode)
        if isinstance(tables, nn.Module):  # positional wrap of an existing collection
            ec, tables, self.debug_mode = tables, None, True
        self.ec = ec if ec is not None else EmbeddingCollection(tables=tables, device=device)
        self._hash_sizes = {f: c.num_embeddings for c in self.ec.embedding_configs() for f in c.feature_names}

    def forward(self, features: KeyedJaggedTensor) -> Dict[str, JaggedTensor]:
        if self.debug_mode:
            _check_ids(features, self._hash_sizes)
        out = self.ec(features)
        if not self.debug_mode:
            return out
        out = out.wait() if hasattr(out, "wait") else out
        wrapped: Dict[str, JaggedTensor] = {}
        for k, jt in out.items():
            if not bool(torch.isfinite(jt.values()).all()):
                raise ValueError(f"non-finite embeddings for {k}")
            wrapped[k] = JaggedTensor(values=self._wrap_tensor(jt.values(), f"ec[{k}].values"), lengths=jt.lengths(), weights=jt.weights_or_none())
        return wrapped


class DebugEmbeddingBagCollection(_DebugBase):
    """``DebugEmbeddingBagCollection(tables, device, debug_mode, is_weighted)`` builds the collection; ``(ebc=existing)`` wraps one."""

    def __init__(self, tables: Optional[Union[List[EmbeddingBagConfig], EmbeddingBagCollection]] = None, device: Optional[torch.device] = None, debug_mode: bool = False,
                 is_weighted: bool = False, ebc: Optional[EmbeddingBagCollection] = None) -> None:
        super().__init__(debug_mode)
        if isinstance(tables, nn.Module):
            ebc, tables, self.debug_mode = tables, None, True
        self.ebc = ebc if ebc is not None else EmbeddingBagCollection(tables=tables, is_weighted=is_weighted, device=device)
        self._hash_sizes = {f: c.num_embeddings for c in self.ebc.embedding_bag_configs() for f in c.feature_names}

    def forward(self, features: KeyedJaggedTensor) -> KeyedTensor:
        if self.debug_mode:
            _check_ids(features, self._hash_sizes)
        out = self.ebc(features)
        if not self.debug_mode:
            return out
        out = out.wait() if hasattr(out, "wait") else out
        if not bool(torch.isfinite(out.values()).all()):
            raise ValueError("non-finite pooled embeddings")
        logger.debug("EBC out: keys=%s shape=%s", out.keys(), tuple(out.values().shape))
        return KeyedTensor(keys=out.keys(), length_per_key=out.length_per_key(), values=self._wrap_tensor(out.values(), "ebc.values"), key_dim=out.key_dim())
