"""Testing toolkit shipped with the framework (reference ``torchrec/distributed/test_utils/``): synthetic model inputs, small reference models, sharders with
pinned sharding type / kernel, multi-process harness, and the config dataclasses the benchmarks are driven by."""
from .model_input import ModelInput, TdModelInput, VariableBatchModelInput  # noqa: F401
from .multi_process import MultiProcessContext, MultiProcessTestBase, run_multi_process_func  # noqa: F401
