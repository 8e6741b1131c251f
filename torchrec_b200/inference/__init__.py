from .modules import (  # noqa: F401
    BatchingMetadata,
    PredictFactory,
    PredictModule,
    quantize_dense,
    quantize_embeddings,
    quantize_feature,
    quantize_inference_model,
    shard_quant_model,
)
