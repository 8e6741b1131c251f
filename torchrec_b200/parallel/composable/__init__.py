from .table_batched_embedding_slice import TableBatchedEmbeddingSlice  # noqa: F401
