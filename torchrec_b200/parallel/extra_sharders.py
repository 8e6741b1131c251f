"""Registry of the additional default sharders (feature-processed EBC, managed-collision, ITEP,
fused collections, quantized inference modules). Imported lazily by ``get_default_sharders``."""
from typing import List

from .types import ModuleSharder


def default_extra_sharders() -> List[ModuleSharder]:
    sharders: List[ModuleSharder] = []
    for mod, name in (
        ("fp_embeddingbag", "FeatureProcessedEmbeddingBagCollectionSharder"),
        ("mc_embeddingbag", "ManagedCollisionEmbeddingBagCollectionSharder"),
        ("mc_embedding", "ManagedCollisionEmbeddingCollectionSharder"),
        ("itep_embeddingbag", "ITEPEmbeddingBagCollectionSharder"),
        ("itep_embeddingbag", "ITEPEmbeddingCollectionSharder"),
        ("fused_embeddingbag", "FusedEmbeddingBagCollectionSharder"),
        ("fused_embeddingbag", "FusedEmbeddingCollectionSharder"),
        ("embeddingbag", "EmbeddingBagSharder"),
        ("pec_embedding", "PECEmbeddingCollectionSharder"),
        ("embedding_tower_sharding", "EmbeddingTowerSharder"),
        ("embedding_tower_sharding", "EmbeddingTowerCollectionSharder"),
        ("quant_embeddingbag", "QuantEmbeddingBagCollectionSharder"),
        ("quant_embedding", "QuantEmbeddingCollectionSharder"),
    ):
        try:
            m = __import__(f"torchrec_b200.parallel.{mod}", fromlist=[name])
            sharders.append(getattr(m, name)())
        except (ImportError, AttributeError):
            continue
    return sharders
