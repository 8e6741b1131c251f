"""Reference import path ``torchrec/distributed/fused_embedding.py`` (``ShardedFusedEmbeddingCollection`` :31, ``FusedEmbeddingCollectionSharder`` :93);
both live next to their bag twins in ``fused_embeddingbag.py``."""
from .fused_embeddingbag import FusedEmbeddingCollectionSharder, ShardedFusedEmbeddingCollection  # noqa: F401
