from .itep_modules import ITEPEmbeddingBagCollection, ITEPEmbeddingCollection  # noqa: F401
