from .apply_optimizer_in_backward import apply_optimizer_in_backward  # noqa: F401
from .clipping import GradientClipping, GradientClippingOptimizer  # noqa: F401
from .fused import EmptyFusedOptimizer, FusedOptimizer, FusedOptimizerModule  # noqa: F401
from .keyed import CombinedOptimizer, KeyedOptimizer, KeyedOptimizerWrapper, OptimizerWrapper  # noqa: F401
from .optimizers import LAMB, SGD, Adagrad, Adam, LarsSGD, PartialRowWiseAdam, PartialRowWiseLAMB, in_backward_optimizer_filter  # noqa: F401
from .rowwise_adagrad import RowWiseAdagrad  # noqa: F401
from .semi_sync import SemisyncOptimizer  # noqa: F401
from .warmup import WarmupOptimizer, WarmupPolicy, WarmupStage  # noqa: F401
