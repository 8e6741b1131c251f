"""JSON (de)serialisation of sparse-module metadata, so a model can be rebuilt (or exported) without pickling code
(reference torchrec/ir/serializer.py:60-500, ir/schema.py, ir/utils.py:encapsulate_ir_modules)."""
from __future__ import annotations

import json
from typing import Any, Callable, Dict, List, Optional, Tuple, Type

import torch
from torch import nn

from ..modules.embedding_configs import DataType, EmbeddingBagConfig, EmbeddingConfig, PoolingType
from ..modules.embedding_modules import EmbeddingBagCollection, EmbeddingCollection
from ..modules.feature_processor_ import PositionWeightedModule, PositionWeightedModuleCollection
from ..modules.fp_embedding_modules import FeatureProcessedEmbeddingBagCollection
from ..modules.regroup import KTRegroupAsDict


def _cfg_to_dict(c) -> Dict[str, Any]:
    d = {"name": c.name, "embedding_dim": c.embedding_dim, "num_embeddings": c.num_embeddings, "feature_names": list(c.feature_names),
         "data_type": c.data_type.value if hasattr(c.data_type, "value") else str(c.data_type)}
    if hasattr(c, "pooling"):
        d["pooling"] = c.pooling.value if hasattr(c.pooling, "value") else str(c.pooling)
    for k in ("weight_init_max", "weight_init_min", "need_pos"):
        if getattr(c, k, None) is not None:
            d[k] = getattr(c, k)
    return d


def _dict_to_bag_cfg(d: Dict[str, Any]) -> EmbeddingBagConfig:
    return EmbeddingBagConfig(name=d["name"], embedding_dim=d["embedding_dim"], num_embeddings=d["num_embeddings"], feature_names=d["feature_names"],
                              data_type=DataType(d["data_type"]), pooling=PoolingType(d.get("pooling", PoolingType.SUM.value)),
                              weight_init_max=d.get("weight_init_max"), weight_init_min=d.get("weight_init_min"))


def _dict_to_cfg(d: Dict[str, Any]) -> EmbeddingConfig:
    return EmbeddingConfig(name=d["name"], embedding_dim=d["embedding_dim"], num_embeddings=d["num_embeddings"], feature_names=d["feature_names"],
                           data_type=DataType(d["data_type"]), weight_init_max=d.get("weight_init_max"), weight_init_min=d.get("weight_init_min"))




from .types import SerializerInterface  # noqa: E402,F401


class EBCJsonSerializer(SerializerInterface):
    module_cls = EmbeddingBagCollection

    @classmethod
    def serialize_to_dict(cls, module: EmbeddingBagCollection) -> Dict[str, Any]:
        return {"tables": [_cfg_to_dict(c) for c in module.embedding_bag_configs()], "is_weighted": module.is_weighted(), "device": str(module.device)}

    @classmethod
    def deserialize_from_dict(cls, d, device=None, children=None) -> EmbeddingBagCollection:
        return EmbeddingBagCollection(tables=[_dict_to_bag_cfg(t) for t in d["tables"]], is_weighted=d["is_weighted"], device=device or torch.device(d.get("device", "cpu")))


class ECJsonSerializer(SerializerInterface):
    module_cls = EmbeddingCollection

    @classmethod
    def serialize_to_dict(cls, module: EmbeddingCollection) -> Dict[str, Any]:
        return {"tables": [_cfg_to_dict(c) for c in module.embedding_configs()], "need_indices": module.need_indices(), "device": str(module.device)}

    @classmethod
    def deserialize_from_dict(cls, d, device=None, children=None) -> EmbeddingCollection:
        return EmbeddingCollection(tables=[_dict_to_cfg(t) for t in d["tables"]], need_indices=d["need_indices"], device=device or torch.device(d.get("device", "cpu")))


class PWMJsonSerializer(SerializerInterface):
    module_cls = PositionWeightedModule

    @classmethod
    def serialize_to_dict(cls, module: PositionWeightedModule) -> Dict[str, Any]:
        return {"max_feature_length": int(module.position_weight.numel())}

    @classmethod
    def deserialize_from_dict(cls, d, device=None, children=None) -> PositionWeightedModule:
        return PositionWeightedModule(d["max_feature_length"], device)


class PWMCJsonSerializer(SerializerInterface):
    module_cls = PositionWeightedModuleCollection

    @classmethod
    def serialize_to_dict(cls, module: PositionWeightedModuleCollection) -> Dict[str, Any]:
        return {"max_feature_lengths": dict(module.max_feature_lengths)}

    @classmethod
    def deserialize_from_dict(cls, d, device=None, children=None) -> PositionWeightedModuleCollection:
        return PositionWeightedModuleCollection(d["max_feature_lengths"], device)


class FPEBCJsonSerializer(SerializerInterface):
    module_cls = FeatureProcessedEmbeddingBagCollection

    @classmethod
    def children(cls, module) -> List[str]:
        return ["_feature_processors", "_embedding_bag_collection"]

    @classmethod
    def serialize_to_dict(cls, module) -> Dict[str, Any]:
        fp = module._feature_processors
        is_collection = isinstance(fp, PositionWeightedModuleCollection)
        return {"is_fp_collection": is_collection, "feature_list": list(module._embedding_bag_collection._feature_names) if hasattr(module._embedding_bag_collection, "_feature_names") else []}

    @classmethod
    def deserialize_from_dict(cls, d, device=None, children=None):
        assert children is not None
        fp = children["_feature_processors"]
        if not d["is_fp_collection"] and hasattr(fp, "_feature_processors"):
            fp = dict(fp._feature_processors.items())
        return FeatureProcessedEmbeddingBagCollection(children["_embedding_bag_collection"], fp)


class KTRegroupAsDictJsonSerializer(SerializerInterface):
    module_cls = KTRegroupAsDict

    @classmethod
    def serialize_to_dict(cls, module: KTRegroupAsDict) -> Dict[str, Any]:
        return {"groups": [list(g) for g in module._groups], "keys": list(module._keys)}

    @classmethod
    def deserialize_from_dict(cls, d, device=None, children=None) -> KTRegroupAsDict:
        return KTRegroupAsDict(d["groups"], d["keys"])


class JsonSerializer:
    """Registry front end: serialises every known sparse module of a model into ``{fqn: {"type", "meta", "children"}}``."""

    module_to_serializer_cls: Dict[str, Type[SerializerInterface]] = {
        "EmbeddingBagCollection": EBCJsonSerializer, "EmbeddingCollection": ECJsonSerializer, "PositionWeightedModule": PWMJsonSerializer,
        "PositionWeightedModuleCollection": PWMCJsonSerializer, "FeatureProcessedEmbeddingBagCollection": FPEBCJsonSerializer, "KTRegroupAsDict": KTRegroupAsDictJsonSerializer}

    @classmethod
    def serialize(cls, module: nn.Module) -> bytes:
        name = type(module).__name__
        if name not in cls.module_to_serializer_cls:
            raise ValueError(f"no IR serializer registered for {name}")
        ser = cls.module_to_serializer_cls[name]
        kids = {c: json.loads(cls.serialize(getattr(module, c)).decode()) for c in ser.children(module) if type(getattr(module, c)).__name__ in cls.module_to_serializer_cls or c == "_feature_processors"} \
            if ser.children(module) else {}
        return json.dumps({"type": name, "meta": ser.serialize_to_dict(module), "children": kids}).encode()

    @classmethod
    def deserialize(cls, blob: bytes, device: Optional[torch.device] = None) -> nn.Module:
        d = json.loads(blob.decode()) if isinstance(blob, (bytes, bytearray)) else blob
        ser = cls.module_to_serializer_cls[d["type"]]
        kids = {k: cls.deserialize(json.dumps(v).encode(), device) for k, v in d.get("children", {}).items()}
        return ser.deserialize_from_dict(d["meta"], device, kids or None)


def serialize_sparse_modules(model: nn.Module) -> Dict[str, bytes]:
    """fqn -> blob for every top-most serialisable sparse module."""
    out: Dict[str, bytes] = {}

    def walk(m: nn.Module, prefix: str) -> None:
        for name, child in m.named_children():
            fqn = f"{prefix}.{name}" if prefix else name
            if type(child).__name__ in JsonSerializer.module_to_serializer_cls:
                out[fqn] = JsonSerializer.serialize(child)
            else:
                walk(child, fqn)

    walk(model, "")
    return out


def encapsulate_ir_modules(model: nn.Module, serializer=JsonSerializer) -> Tuple[nn.Module, List[str]]:
    """Attach ``ir_metadata`` to every serialisable module (export time). Returns (model, fqns)."""
    blobs = serialize_sparse_modules(model)
    for fqn, blob in blobs.items():
        model.get_submodule(fqn).ir_metadata = blob  # type: ignore[assignment]
    return model, list(blobs)


def decapsulate_ir_modules(model: nn.Module, serializer=JsonSerializer, device: Optional[torch.device] = None) -> nn.Module:
    """Rebuild every module carrying ``ir_metadata`` from its blob (load time), keeping parameters when shapes match."""
    for fqn, m in list(model.named_modules()):
        blob = getattr(m, "ir_metadata", None)
        if blob is None or not fqn:
            continue
        new = serializer.deserialize(blob, device)
        try:
            new.load_state_dict(m.state_dict(), strict=False)
        except Exception:
            pass
        parent = model.get_submodule(fqn.rsplit(".", 1)[0]) if "." in fqn else model
        setattr(parent, fqn.rsplit(".", 1)[-1], new)
    return model
