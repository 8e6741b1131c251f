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


# ---- the other sharders of the reference's test kit ----------------------------------------------------------------------------------------------------------
from ..embedding_tower_sharding import EmbeddingTowerCollectionSharder, EmbeddingTowerSharder  # noqa: E402
from ..embeddingbag import EmbeddingBagSharder  # noqa: E402
from ..mc_embedding import ManagedCollisionEmbeddingCollectionSharder  # noqa: E402
from ..mc_embeddingbag import ManagedCollisionEmbeddingBagCollectionSharder  # noqa: E402
from ..mc_modules import ManagedCollisionCollectionSharder  # noqa: E402

TestEBSharder = _pinned(EmbeddingBagSharder)  # nn.EmbeddingBag
TestETSharder = _pinned(EmbeddingTowerSharder)
TestETCSharder = _pinned(EmbeddingTowerCollectionSharder)


class TestMCSharder(ManagedCollisionCollectionSharder):
    """Managed-collision collection pinned to one sharding type."""

    __test__ = False

    def __init__(self, sharding_type: str, qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None) -> None:
        super().__init__()
        self._sharding_type = sharding_type

    def sharding_types(self, compute_device_type: str) -> List[str]:
        return [self._sharding_type]


class TestEBCSharderMCH(ManagedCollisionEmbeddingBagCollectionSharder):
    """MC-EBC whose bags and collision modules are pinned to one sharding type / kernel."""

    __test__ = False

    def __init__(self, sharding_type: str, kernel_type: str, fused_params: Optional[Dict[str, Any]] = None,
                 qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None) -> None:
        super().__init__(TestEBCSharder(sharding_type, kernel_type, fused_params, qcomm_codecs_registry), TestMCSharder(sharding_type, qcomm_codecs_registry),
                         fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)
        self._sharding_type = sharding_type

    def sharding_types(self, compute_device_type: str) -> List[str]:
        return [self._sharding_type]


class TestECSharderMCH(ManagedCollisionEmbeddingCollectionSharder):
    __test__ = False

    def __init__(self, sharding_type: str, kernel_type: str, fused_params: Optional[Dict[str, Any]] = None,
                 qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None) -> None:
        super().__init__(TestECSharder(sharding_type, kernel_type, fused_params, qcomm_codecs_registry), TestMCSharder(sharding_type, qcomm_codecs_registry),
                         fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)
        self._sharding_type = sharding_type

    def sharding_types(self, compute_device_type: str) -> List[str]:
        return [self._sharding_type]
