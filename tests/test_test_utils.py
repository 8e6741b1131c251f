"""The shipped testing toolkit (``torchrec_b200.distributed.test_utils``): synthetic inputs, reference models, pinned sharders, configs, harness."""
import pytest
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def test_model_input_generation_and_reference_model():
    from torchrec_b200.distributed.test_utils import ModelInput, TdModelInput, VariableBatchModelInput
    from torchrec_b200.distributed.test_utils.emb_sharder import TestEBCSharder, TestECSharder
    from torchrec_b200.distributed.test_utils.input_config import ModelInputConfig
    from torchrec_b200.distributed.test_utils.model_config import ModelSelectionConfig, create_model_config
    from torchrec_b200.distributed.test_utils.table_config import EmbeddingTablesConfig, TableExtendedConfigs
    from torchrec_b200.distributed.test_utils.test_model import TestSparseNN

    tables, weighted = EmbeddingTablesConfig(num_unweighted_features=3, num_weighted_features=2, embedding_feature_dim=8, num_embeddings=50,
                                             table_extended_configs=TableExtendedConfigs({"table_1": {"embedding_dim": 16}})).generate_tables()
    assert [t.embedding_dim for t in tables] == [8, 16, 8] and [t.name for t in weighted] == ["weighted_table_0", "weighted_table_1"]
    g, locals_ = ModelInput.generate_global_and_local_batches(2, batch_size=4, tables=tables, weighted_tables=weighted, num_float_features=10, pooling_avg=3)
    assert g.float_features.shape == (8, 10) and g.idlist_features.stride() == 8 and g.idscore_features.weights().numel() == g.idscore_features.values().numel()
    m = TestSparseNN(tables, weighted, num_float_features=10)
    loss, pred = m(g)
    loss.backward()
    m.eval()
    torch.testing.assert_close(torch.cat([m(b) for b in locals_]), m(g))          # the global batch is the rank-ordered concatenation of the local ones
    hot = ModelInput.generate(batch_size=256, tables=tables, power_law_alpha=1.2, pooling_avg=5).idlist_features.values()
    uni = ModelInput.generate(batch_size=256, tables=tables, pooling_avg=5).idlist_features.values()
    assert int(torch.bincount(hot).max()) > 3 * int(torch.bincount(uni).max())     # power-law ids concentrate on few rows
    off = ModelInput.generate(batch_size=4, tables=tables, use_offsets=True, indices_dtype=torch.int32)
    assert off.idlist_features.values().dtype == torch.int32 and off.idlist_features.offsets().numel() == 3 * 4 + 1
    vb = VariableBatchModelInput.generate(batch_size=6, tables=tables, pooling_avg=2)
    assert vb.idlist_features.variable_stride_per_key() and len(vb.idlist_features.stride_per_key_per_rank()) == 3
    assert sorted(TdModelInput.generate(batch_size=2, tables=tables).idlist_features.keys()) == ["feature_0", "feature_1", "feature_2"]
    s = TestEBCSharder("row_wise", "fused", fused_params={"learning_rate": 0.1})
    assert s.sharding_types("cuda") == ["row_wise"] and s.compute_kernels("row_wise", "cuda") == ["fused"] and s.fused_params["learning_rate"] == 0.1
    assert TestECSharder("table_wise", "fused").sharding_types("cpu") == ["table_wise"]
    assert len(ModelInputConfig(num_batches=2, batch_size=4, pin_memory=False).generate_batches(tables, weighted)) == 2
    assert type(create_model_config("dlrm", num_float_features=13, embedding_dim=8)).__name__ == "DLRMConfig"
    model = ModelSelectionConfig(model_name="test_model_with_preproc").create_test_model(tables, weighted, torch.device("cpu"))
    loss2, _ = model(g)
    assert torch.isfinite(loss2)


def _train_with_configs(rank: int, world_size: int, steps: int = 3) -> None:
    import torch.distributed as dist

    from torchrec_b200.distributed.test_utils import ModelInput
    from torchrec_b200.distributed.test_utils.pipeline_config import PipelineConfig
    from torchrec_b200.distributed.test_utils.sharding_config import PlannerConfig, ShardingConfig
    from torchrec_b200.distributed.test_utils.table_config import EmbeddingTablesConfig
    from torchrec_b200.distributed.test_utils.test_model import TestSparseNN

    dev = torch.device("cpu")
    tables, weighted = EmbeddingTablesConfig(num_unweighted_features=4, num_weighted_features=0, embedding_feature_dim=8, num_embeddings=200).generate_tables()
    model = TestSparseNN(tables, weighted, num_float_features=10, sparse_device=torch.device("meta"))
    cfg = ShardingConfig(planner=PlannerConfig(world_size=world_size, compute_device="cpu", batch_size=8), sparse_lr=0.05, dense_lr=0.05)
    dmp, opt = cfg.generate_sharded_model_and_optimizer(model, dist.group.WORLD, dev)
    pipe = PipelineConfig(pipeline="sparse").generate_pipeline(dmp, opt, dev)
    g = torch.Generator().manual_seed(rank)
    batches = [ModelInput.generate(batch_size=8, tables=tables, num_float_features=10, pooling_avg=3, generator=g) for _ in range(steps + 2)]
    it = iter(batches)
    losses = []
    for _ in range(steps):
        out = pipe.progress(it)
        pred = out[1] if isinstance(out, tuple) else out          # the pipeline hands back the model output (loss is consumed by backward)
        assert pred.shape == (8,) and bool(torch.isfinite(pred).all()) and 0.0 <= float(pred.min()) and float(pred.max()) <= 1.0


def test_configs_drive_a_sharded_training_run():
    from torchrec_b200.distributed.test_utils import run_multi_process_func

    run_multi_process_func(_train_with_configs, world_size=2, backend="gloo", steps=3)


def test_infer_utils_quantize_shard_compare():
    """The inference test kit: TestSparseNN behind a plain-tensor signature, its quantized copy, the quantized copy sharded over two local
    devices with a forced sharding type - all three agree; expected shards are checked against the plan; mock kernels keep shapes."""
    import torch

    from torchrec_b200.parallel.test_utils.infer_utils import (KJTInputExportWrapper, MockTBE, assert_close, create_cw_min_partition_constraints, create_test_model,
                                                               create_test_model_ebc_only, model_input_to_forward_args, model_input_to_forward_args_kjt, prep_inputs,
                                                               prep_inputs_multiprocess, quantize, replace_registered_tbes_with_mock_tbes, shard_qebc)
    from torchrec_b200.parallel.types import ShardingType

    dev = torch.device("cpu")
    mi = create_test_model(num_embeddings=64, emb_dim=16, world_size=2, batch_size=4, dense_device=dev, sparse_device=dev, num_features=2, num_weighted_features=1)
    inputs = prep_inputs(mi, 2, batch_size=4, count=2)
    assert len(prep_inputs_multiprocess(mi, 2, 4, count=3)) == 3
    for st in (ShardingType.TABLE_WISE, ShardingType.ROW_WISE, ShardingType.COLUMN_WISE):
        sharded = shard_qebc(mi, st, dev, shard_score_ebc=True)
        assert type(sharded._module.sparse.ebc).__name__ == "ShardedQuantEmbeddingBagCollection"
        for b in inputs:
            args = model_input_to_forward_args(b)
            assert_close(mi.quant_model(*args), sharded(*args))
            # int8 rows: close to the float model, not equal
            assert float((mi.model(*args) - mi.quant_model(*args)).abs().max()) < 0.05
    rw = [[((0, 0, 32, 16), "rank:0/cpu"), ((32, 0, 32, 16), "rank:1/cpu")]] * 2
    shard_qebc(mi, ShardingType.ROW_WISE, dev, expected_shards=rw)
    with pytest.raises(AssertionError):
        shard_qebc(mi, ShardingType.ROW_WISE, dev, expected_shards=[[((0, 0, 64, 16), "rank:0/cpu")]] * 2)
    cons = create_cw_min_partition_constraints([("table_0", 8)])
    assert cons["table_0"].min_partition == 8 and cons["table_0"].sharding_types == ["column_wise"]
    # bags only, KJT arguments, kernels registered as sub-modules
    mi2 = create_test_model_ebc_only(64, 16, 2, 4, dev, dev, num_features=2, compute_device="cpu")
    b = prep_inputs(mi2, 2, 4, 1)[0]
    a = model_input_to_forward_args_kjt(b)
    out_q = mi2.quant_model(*a)
    sh = shard_qebc(mi2, ShardingType.TABLE_WISE, dev, ebc_fqn="_module_kjt_input.sparse.ebc")
    assert_close(out_q, sh(*a))
    export = KJTInputExportWrapper(mi2.quant_model._module_kjt_input, a[0])
    assert_close([out_q.values()], export(a[1], a[3]))
    with pytest.raises(AssertionError):
        assert_close(out_q, mi2.model(*a))  # float vs int8 rows differ beyond the default tolerance
    int4 = quantize(mi2.model, inplace=False, weight_dtype=torch.quint4x2)
    assert float((int4(*a).values() - mi2.model(*a).values()).abs().max()) < 0.5
    replace_registered_tbes_with_mock_tbes(mi2.quant_model)
    mocks = [m for m in mi2.quant_model.modules() if isinstance(m, MockTBE)]
    assert mocks and sum(len(m.embedding_specs) for m in mocks) == 2
    assert mi2.quant_model(*a).values().shape == out_q.values().shape and float(mi2.quant_model(*a).values().abs().sum()) == 0.0
