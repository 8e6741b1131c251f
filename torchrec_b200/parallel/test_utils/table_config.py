"""Table generation config (reference ``torchrec/distributed/test_utils/table_config.py``: ``ManagedCollisionConfig`` :82, ``TableExtendedConfigs`` :108,
``EmbeddingTablesConfig`` :120)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

from ...modules.embedding_configs import DataType, EmbeddingBagConfig, EmbeddingConfig


@dataclass
class ManagedCollisionConfig:
    zch_size: Optional[int] = None
    eviction_interval: int = 1
    eviction_policy: str = "lfu"


@dataclass
class TableExtendedConfigs:
    """Per-table overrides: ``{"table_3": {"num_embeddings": 1_000_000, "embedding_dim": 256, "feature_names": [...]}}``."""

    overrides: Dict[str, Dict[str, Any]] = field(default_factory=dict)


@dataclass
class EmbeddingTablesConfig:
    num_unweighted_features: int = 100
    num_weighted_features: int = 100
    embedding_feature_dim: int = 128
    num_embeddings: int = 100_000
    data_type: str = "FP32"
    sequence: bool = False
    additional_tables: List[List[Dict[str, Any]]] = field(default_factory=list)
    table_extended_configs: Optional[TableExtendedConfigs] = None

    def generate_tables(self) -> Tuple[List[Any], List[Any]]:
        """(unweighted tables, weighted tables): table ``i`` has ``(i + 1) * num_embeddings / n`` ... no: every table ``max(i + 1, num_embeddings)`` rows like
        the reference, one feature each, so hash sizes are distinct and mis-routed ids show up in tests."""
        cls = EmbeddingConfig if self.sequence else EmbeddingBagConfig
        dt = DataType[self.data_type] if isinstance(self.data_type, str) else self.data_type
        ov = self.table_extended_configs.overrides if self.table_extended_configs else {}

        def mk(i: int, prefix: str, feat_prefix: str) -> Any:
            kw = dict(num_embeddings=max(i + 1, self.num_embeddings), embedding_dim=self.embedding_feature_dim, name=f"{prefix}{i}", feature_names=[f"{feat_prefix}{i}"], data_type=dt)
            kw.update(ov.get(kw["name"], {}))
            return cls(**kw)

        tables = [mk(i, "table_", "feature_") for i in range(self.num_unweighted_features)]
        weighted = [mk(i, "weighted_table_", "weighted_feature_") for i in range(self.num_weighted_features)]
        for group in self.additional_tables:
            for spec in group:
                tables.append(cls(**{**dict(data_type=dt), **spec}))
        return tables, weighted
