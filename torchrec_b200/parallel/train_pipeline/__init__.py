from .pipeline_context import EmbeddingTrainPipelineContext, PrefetchTrainPipelineContext, TrainPipelineContext  # noqa: F401
from .train_pipelines import (  # noqa: F401
    EvalPipelineFusedSparseDist,
    EvalPipelineSparseDist,
    ModelDetachedException,
    TorchCompileConfig,
    PipelinedForward,
    PipelineStage,
    PrefetchTrainPipelineSparseDist,
    StagedTrainPipeline,
    TrainPipeline,
    TrainPipelineBase,
    TrainPipelineFusedSparseDist,
    TrainPipelineSemiSync,
    TrainPipelineSparseDist,
    TrainPipelineSparseDistLite,
    TrainPipelinePT2,
    TrainPipelineSparseDistCompAutograd,
)
from .utils import DataLoadingThread, PipelinedPostproc, SparseDataDistUtil  # noqa: F401
