"""Single-process fused embedding collections: one table-batched kernel + optimizer fused into the backward
(reference torchrec/modules/fused_embedding_modules.py:279,529,796)."""
from __future__ import annotations

import copy
from collections import OrderedDict
from typing import Any, Dict, Iterator, List, Optional, Tuple, Type

import torch
from torch import nn

from ..ops.tbe import OptimType, PoolingMode, TableBatchedEmbeddingBags
from ..optim.fused import FusedOptimizer, FusedOptimizerModule
from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor, KeyedTensor
from .embedding_configs import EmbeddingBagConfig, EmbeddingConfig, PoolingType, data_type_to_dtype
from .embedding_modules import EmbeddingBagCollectionInterface, EmbeddingCollectionInterface, get_embedding_names_by_table


def _optim_type(optimizer_type: Type[torch.optim.Optimizer]) -> OptimType:
    from ..optim import optimizers as shells
    from ..optim.rowwise_adagrad import RowWiseAdagrad

    m = {torch.optim.SGD: OptimType.EXACT_SGD, torch.optim.Adagrad: OptimType.EXACT_ADAGRAD, torch.optim.Adam: OptimType.ADAM, torch.optim.AdamW: OptimType.ADAMW,
         RowWiseAdagrad: OptimType.EXACT_ROWWISE_ADAGRAD, shells.SGD: OptimType.EXACT_SGD, shells.Adagrad: OptimType.EXACT_ADAGRAD, shells.Adam: OptimType.ADAM,
         shells.LAMB: OptimType.LAMB, shells.LarsSGD: OptimType.LARS_SGD, shells.PartialRowWiseAdam: OptimType.PARTIAL_ROWWISE_ADAM,
         shells.PartialRowWiseLAMB: OptimType.PARTIAL_ROWWISE_LAMB}
    if optimizer_type not in m:
        raise ValueError(f"Cannot fuse optimizer_type={optimizer_type} with the embedding backward")
    return m[optimizer_type]


def convert_optimizer_type_and_kwargs(optimizer_type: Type[torch.optim.Optimizer], optimizer_kwargs: Dict[str, Any]) -> Optional[Tuple[OptimType, Dict[str, Any]]]:
    """A torch (or shell) optimizer class + its kwargs -> (fused kernel optimizer, kernel kwargs: ``lr`` becomes ``learning_rate``); None
    when the class has no fused counterpart."""
    import copy

    kwargs = copy.deepcopy(optimizer_kwargs)
    if "lr" in kwargs:
        kwargs["learning_rate"] = kwargs.pop("lr")
    try:
        return _optim_type(optimizer_type), kwargs
    except ValueError:
        return None


class _TbeFusedOptimizer(FusedOptimizer):
    def __init__(self, tbes: List[TableBatchedEmbeddingBags], params: Dict[str, torch.Tensor], state: Dict[Any, Any]) -> None:
        self._tbes = tbes
        super().__init__(params, state, [{"params": list(params.values()), "lr": tbes[0].get_learning_rate() if tbes else 0.0}])

    def zero_grad(self, set_to_none: bool = False) -> None:
        pass

    def step(self, closure: Any = None) -> None:
        for t in self._tbes:
            t.set_learning_rate(self.param_groups[0]["lr"])

    def set_optimizer_step(self, step: int) -> None:
        for t in self._tbes:
            t.set_optimizer_step(step)


class FusedEmbeddingBagCollection(EmbeddingBagCollectionInterface, FusedOptimizerModule):
    """EBC backed by the table-batched kernel with ``optimizer_type`` applied inside the backward.
    ``forward(KJT) -> KeyedTensor``; parameters are exposed as ``embedding_bags.<table>.weight`` views."""

    def __init__(self, tables: List[EmbeddingBagConfig], optimizer_type: Type[torch.optim.Optimizer], optimizer_kwargs: Dict[str, Any], is_weighted: bool = False,
                 device: Optional[torch.device] = None, location: Optional[Any] = None) -> None:
        super().__init__()
        self._is_weighted = is_weighted
        self._embedding_bag_configs = tables
        self._device = torch.device(device) if device is not None else torch.device("cpu")
        self._optimizer_type, self._optimizer_kwargs = optimizer_type, optimizer_kwargs
        kw = dict(optimizer_kwargs)
        betas = kw.get("betas", (0.9, 0.999))
        common = dict(optimizer=_optim_type(optimizer_type), learning_rate=kw.get("lr", 0.01), eps=kw.get("eps", 1e-8), beta1=betas[0], beta2=betas[1],
                      weight_decay=kw.get("weight_decay", 0.0), device=self._device)
        self._groups: List[Tuple[List[int], TableBatchedEmbeddingBags, List[str]]] = []
        self._tbes = nn.ModuleList()
        for pt in (PoolingType.SUM, PoolingType.MEAN):
            idxs = [i for i, t in enumerate(tables) if t.pooling == pt]
            if not idxs:
                continue
            tbe = TableBatchedEmbeddingBags([(tables[i].num_embeddings, tables[i].embedding_dim) for i in idxs],
                                            [k for k, i in enumerate(idxs) for _ in tables[i].feature_names],
                                            pooling_mode=PoolingMode.MEAN if pt == PoolingType.MEAN else PoolingMode.SUM,
                                            table_names=[tables[i].name for i in idxs], **common)
            tbe.init_parameters([(tables[i].get_weight_init_min(), tables[i].get_weight_init_max()) for i in idxs])
            self._tbes.append(tbe)
            self._groups.append((idxs, tbe, [f for i in idxs for f in tables[i].feature_names]))
        self._embedding_names = [n for ns in get_embedding_names_by_table(tables) for n in ns]
        self._length_per_key = [t.embedding_dim for t in tables for _ in t.feature_names]
        self.embedding_bags = nn.ModuleDict()
        params: Dict[str, torch.Tensor] = {}
        state: Dict[Any, Any] = {}
        for idxs, tbe, _ in self._groups:
            ws, sts = tbe.split_embedding_weights(), tbe.split_optimizer_states()
            for k, i in enumerate(idxs):
                holder = nn.Module()
                p = nn.Parameter(ws[k], requires_grad=False)
                p._in_backward_optimizers = [None]  # type: ignore[attr-defined]
                holder.weight = p
                self.embedding_bags[tables[i].name] = holder
                params[f"embedding_bags.{tables[i].name}.weight"] = p
                state[p] = {f"{tables[i].name}.{n}": v for n, v in sts[k].items()}
        self._optim = _TbeFusedOptimizer(list(self._tbes), params, state)
        flat = [(i, f) for i, t in enumerate(tables) for f in t.feature_names]
        order = [(i, f) for idxs, _, _ in self._groups for i in idxs for f in tables[i].feature_names]
        self._regroup = None if order == flat else (order, flat)

    def forward(self, features: KeyedJaggedTensor) -> KeyedTensor:
        outs = []
        for idxs, tbe, feats in self._groups:
            sub = features if feats == features.keys() else features.permute([features.keys().index(f) for f in feats])
            outs.append(tbe(sub.values(), sub.offsets(), sub.weights_or_none() if self._is_weighted else None, batch_size=sub.stride()))
        vals = outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)
        if self._regroup is not None:
            order, flat = self._regroup
            dims = {k: self._embedding_bag_configs[k[0]].embedding_dim for k in order}
            start, c = {}, 0
            for k in order:
                start[k] = c
                c += dims[k]
            vals = vals.index_select(1, torch.cat([torch.arange(start[k], start[k] + dims[k], device=vals.device) for k in flat]))
        return KeyedTensor(keys=self._embedding_names, length_per_key=self._length_per_key, values=vals)

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, nn.Parameter]]:
        for name, holder in self.embedding_bags.items():
            yield (prefix + "." if prefix else "") + f"embedding_bags.{name}.weight", holder.weight

    def state_dict(self, destination=None, prefix: str = "", keep_vars: bool = False):
        destination = OrderedDict() if destination is None else destination
        for name, holder in self.embedding_bags.items():
            destination[f"{prefix}embedding_bags.{name}.weight"] = holder.weight if keep_vars else holder.weight.detach()
        return destination

    @torch.no_grad()
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        for name, holder in self.embedding_bags.items():
            holder.weight.copy_(state_dict[f"embedding_bags.{name}.weight"])
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def fused_optimizer(self):  # type: ignore[override]
        return self._optim

    fused_optimizer = property(lambda self: self._optim)  # type: ignore[assignment]

    def embedding_bag_configs(self) -> List[EmbeddingBagConfig]:
        return self._embedding_bag_configs

    def is_weighted(self) -> bool:
        return self._is_weighted

    @property
    def device(self) -> torch.device:
        return self._device


class FusedEmbeddingCollection(EmbeddingCollectionInterface, FusedOptimizerModule):
    """Sequence-embedding twin of FusedEmbeddingBagCollection."""

    def __init__(self, tables: List[EmbeddingConfig], optimizer_type: Type[torch.optim.Optimizer], optimizer_kwargs: Dict[str, Any],
                 device: Optional[torch.device] = None, need_indices: bool = False, location: Optional[Any] = None) -> None:
        super().__init__()
        self._embedding_configs = tables
        self._need_indices = need_indices
        self._device = torch.device(device) if device is not None else torch.device("cpu")
        self._embedding_dim = tables[0].embedding_dim
        kw = dict(optimizer_kwargs)
        betas = kw.get("betas", (0.9, 0.999))
        self._tbe = TableBatchedEmbeddingBags([(t.num_embeddings, t.embedding_dim) for t in tables], [k for k, t in enumerate(tables) for _ in t.feature_names],
                                              pooling_mode=PoolingMode.NONE, optimizer=_optim_type(optimizer_type), learning_rate=kw.get("lr", 0.01),
                                              eps=kw.get("eps", 1e-8), beta1=betas[0], beta2=betas[1], weight_decay=kw.get("weight_decay", 0.0), device=self._device,
                                              table_names=[t.name for t in tables])
        self._tbe.init_parameters([(t.get_weight_init_min(), t.get_weight_init_max()) for t in tables])
        self._feature_names = [f for t in tables for f in t.feature_names]
        self._embedding_names_by_table = get_embedding_names_by_table(tables)
        self._embedding_names = [n for ns in self._embedding_names_by_table for n in ns]
        self.embeddings = nn.ModuleDict()
        params, state = {}, {}
        ws, sts = self._tbe.split_embedding_weights(), self._tbe.split_optimizer_states()
        for k, t in enumerate(tables):
            holder = nn.Module()
            p = nn.Parameter(ws[k], requires_grad=False)
            p._in_backward_optimizers = [None]  # type: ignore[attr-defined]
            holder.weight = p
            self.embeddings[t.name] = holder
            params[f"embeddings.{t.name}.weight"] = p
            state[p] = {f"{t.name}.{n}": v for n, v in sts[k].items()}
        self._optim = _TbeFusedOptimizer([self._tbe], params, state)

    def forward(self, features: KeyedJaggedTensor) -> Dict[str, JaggedTensor]:
        sub = features if self._feature_names == features.keys() else features.permute([features.keys().index(f) for f in self._feature_names])
        emb = self._tbe(sub.values(), sub.offsets(), None, batch_size=sub.stride())
        lpk = sub.length_per_key()
        parts = torch.split(emb, lpk, dim=0)
        vals = torch.split(sub.values(), lpk) if self._need_indices else None
        lengths = sub.lengths().view(len(self._feature_names), sub.stride())
        return {n: JaggedTensor(values=parts[i], lengths=lengths[i], weights=vals[i] if vals is not None else None) for i, n in enumerate(self._embedding_names)}

    fused_optimizer = property(lambda self: self._optim)  # type: ignore[assignment]

    def embedding_configs(self) -> List[EmbeddingConfig]:
        return self._embedding_configs

    def need_indices(self) -> bool:
        return self._need_indices

    def embedding_dim(self) -> int:
        return self._embedding_dim

    def embedding_names_by_table(self) -> List[List[str]]:
        return self._embedding_names_by_table

    @property
    def device(self) -> torch.device:
        return self._device


def fuse_embedding_optimizer(model: nn.Module, optimizer_type: Type[torch.optim.Optimizer], optimizer_kwargs: Dict[str, Any], device: torch.device,
                             location: Optional[Any] = None) -> nn.Module:
    """Swap every EmbeddingBagCollection / EmbeddingCollection of ``model`` for its fused version (weights copied). ``location``: where the
    fused bag tables live (``EmbeddingLocation``: HBM / host-mapped / cached)."""
    from .embedding_modules import EmbeddingBagCollection, EmbeddingCollection

    def swap(m: nn.Module) -> nn.Module:
        if isinstance(m, EmbeddingBagCollection):
            f = FusedEmbeddingBagCollection(m.embedding_bag_configs(), optimizer_type, optimizer_kwargs, m.is_weighted(), device, location)
            with torch.no_grad():
                for n, bag in m.embedding_bags.items():
                    if bag.weight.device.type != "meta":
                        f.embedding_bags[n].weight.copy_(bag.weight)
            return f
        if isinstance(m, EmbeddingCollection):
            f = FusedEmbeddingCollection(m.embedding_configs(), optimizer_type, optimizer_kwargs, device, m.need_indices())
            with torch.no_grad():
                for n, e in m.embeddings.items():
                    if e.weight.device.type != "meta":
                        f.embeddings[n].weight.copy_(e.weight)
            return f
        for name, child in list(m.named_children()):
            setattr(m, name, swap(child))
        return m

    return swap(model)


EmbeddingFusedOptimizer = _TbeFusedOptimizer  # the fused optimizer handle of the fused collections under its reference name
