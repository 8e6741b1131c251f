"""Table-wise-column-wise sharding: column blocks of a table restricted to the ranks of ONE host (blocks talk over NVLink only).

Reference: ``torchrec/distributed/sharding/twcw_sharding.py:18`` - identical mechanics to column-wise; only the planner's placement differs.
"""
from __future__ import annotations

from .cw_sharding import CwPooledEmbeddingSharding


class TwCwPooledEmbeddingSharding(CwPooledEmbeddingSharding):
    pass
