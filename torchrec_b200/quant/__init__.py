from .embedding_modules import EmbeddingBagCollection, EmbeddingCollection, quantize_state_dict  # noqa: F401
