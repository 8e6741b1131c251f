"""Sharders pinned to ONE sharding type and compute kernel (reference ``torchrec/distributed/test_utils/emb_sharder.py``): tests ask for a specific
placement instead of letting the planner choose."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

from ..embedding import EmbeddingCollectionSharder
from ..embeddingbag import EmbeddingBagCollectionSharder
from ..fused_embeddingbag import FusedEmbeddingBagCollectionSharder, FusedEmbeddingCollectionSharder
from ..types import QuantizedCommCodecs


def _pinned(base: type) -> type:
    class Pinned(base):  # type: ignore[misc, valid-type]
        __test__ = False

        def __init__(self, sharding_type: str, kernel_type: str, fused_params: Optional[Dict[str, Any]] = None,
                     qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None, **kw: Any) -> None:
            if fused_params is None:
                fused_params = {}
            try:
                super().__init__(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry, **kw)
            except TypeError:
                super().__init__(fused_params=fused_params)
            self._sharding_type = sharding_type
            self._kernel_type = kernel_type

        def sharding_types(self, compute_device_type: str) -> List[str]:
            return [self._sharding_type]

        def compute_kernels(self, sharding_type: str, compute_device_type: str) -> List[str]:
            return [self._kernel_type]

    Pinned.__name__ = Pinned.__qualname__ = f"Test{base.__name__}"
    return Pinned


TestEBCSharder = _pinned(EmbeddingBagCollectionSharder)
TestECSharder = _pinned(EmbeddingCollectionSharder)
TestFusedEBCSharder = _pinned(FusedEmbeddingBagCollectionSharder)
TestFusedECSharder = _pinned(FusedEmbeddingCollectionSharder)
