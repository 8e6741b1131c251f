"""Group tables into TBE buckets by row width (reference torchrec/distributed/embedding_dim_bucketer.py:18-154).

Why: a table-batched kernel sizes its per-warp work (registers, vector count ``MAXV``) and a software-managed cache sizes its slots
for the WIDEST table it contains; batching a 16-wide table with a 1024-wide one wastes occupancy / cache capacity. Three policies:
one bucket, one bucket per distinct width, one bucket per 128-byte line count (the default of the UVM-caching path)."""
from __future__ import annotations

from enum import Enum, unique
from typing import Any, Dict, List

from ..modules.embedding_configs import DATA_TYPE_NUM_BITS, DataType


@unique
class EmbDimBucketerPolicy(Enum):
    SINGLE_BUCKET = "single_bucket"
    ALL_BUCKETS = "all_buckets"
    CACHELINE_BUCKETS = "cacheline_buckets"


class EmbDimBucketer:
    def __init__(self, embedding_tables: List[Any], cfg: EmbDimBucketerPolicy, cacheline: int = 128) -> None:
        """``embedding_tables``: objects with ``local_cols`` (or ``embedding_dim``) and ``data_type`` (ShardedEmbeddingTable, configs)."""
        self.cacheline = cacheline
        self.num_buckets = 1
        cfg = EmbDimBucketerPolicy(cfg)
        widths = [self.dim_in_bytes(self._cols(t), t.data_type) for t in embedding_tables]
        self.emb_dim_buckets: Dict[int, int] = {}
        if cfg == EmbDimBucketerPolicy.SINGLE_BUCKET:
            self.emb_dim_buckets = {w: 0 for w in widths}
        elif cfg == EmbDimBucketerPolicy.ALL_BUCKETS:
            for w in widths:
                self.emb_dim_buckets.setdefault(w, len(self.emb_dim_buckets))
            self.num_buckets = max(1, len(self.emb_dim_buckets))
        else:
            by_lines: Dict[int, int] = {}
            for w in widths:
                lines = (w + cacheline - 1) // cacheline
                by_lines.setdefault(lines, len(by_lines))
                self.emb_dim_buckets[w] = by_lines[lines]
            self.num_buckets = max(1, len(by_lines))

    @staticmethod
    def _cols(t: Any) -> int:
        return int(getattr(t, "local_cols", None) or t.embedding_dim)

    def bucket_count(self) -> int:
        return self.num_buckets

    def get_bucket(self, embedding_dim: int, dtype: DataType) -> int:
        return 0 if self.num_buckets == 1 else self.bucket(embedding_dim, dtype)

    def bucket(self, dim: int, dtype: DataType) -> int:
        return self.emb_dim_buckets[self.dim_in_bytes(dim, dtype)]

    def dim_in_bytes(self, dim: int, dtype: DataType) -> int:
        return (dim * DATA_TYPE_NUM_BITS[DataType(dtype)] + 7) // 8


def should_do_dim_bucketing(embedding_tables: List[Any]) -> bool:
    """Bucket only when it can matter: cached tables taking part in the prefetch pipeline, with more than one row width
    (reference embedding_dim_bucketer.py:157-195)."""
    widths = set()
    for t in embedding_tables:
        kernel = getattr(getattr(t, "compute_kernel", None), "value", str(getattr(t, "compute_kernel", "")))
        if kernel in ("fused_uvm_caching", "quant_uvm_caching", "key_value") and bool((getattr(t, "fused_params", None) or {}).get("prefetch_pipeline", False)):
            widths.add(int(getattr(t, "local_cols", None) or t.embedding_dim))
    return len(widths) > 1
