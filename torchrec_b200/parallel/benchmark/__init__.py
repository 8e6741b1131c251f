"""``torchrec.distributed.benchmark`` import paths. The benchmark harness and the larger benchmarks live in ``torchrec_b200.benchmarks``
(``base``, ``benchmark_train_pipeline``, ``benchmark_comms``, ``benchmark_inference``, ``benchmark_model_lifecycle``, ``benchmark_zch``):
they are registered here under the reference's module names (same module objects); the micro-benchmarks of the reference's package
(``benchmark_train``, ``benchmark_split_table_batched_embeddings``, ``benchmark_set_sharding_context_post_a2a``,
``embedding_collection_wrappers``, ``utils``) are modules of this package."""
import importlib
import sys

for _name in ("base", "benchmark_train_pipeline", "benchmark_comms", "benchmark_inference", "benchmark_model_lifecycle", "benchmark_zch"):
    _mod = importlib.import_module(f"torchrec_b200.benchmarks.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    globals()[_name] = _mod
