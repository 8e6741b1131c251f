"""Quantized comm codecs, composable shard() API and collective_utils on 2 CPU ranks."""
import pytest
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def test_qcomm_codecs_roundtrip_and_registry():
    from torchrec_b200.parallel.qcomm_codec import CommType, QCommsConfig, get_qcomm_codec, get_qcomm_codecs_registry
    from torchrec_b200.parallel.types import CommOp

    x = torch.randn(64 * 32)
    for ct, tol in ((CommType.FP32, 0.0), (CommType.FP16, 2e-3), (CommType.BF16, 2e-2), (CommType.INT8, 5e-2), (CommType.FP8, 1.5e-1)):
        codec = get_qcomm_codec(ct, None, 32)
        enc = codec.encode(x)
        assert enc.numel() == codec.calc_quantized_size(x.numel()) and enc.dtype == codec.quantized_dtype
        dec = codec.decode(enc).float().reshape(-1)[: x.numel()]
        assert float((dec - x).abs().max()) <= tol * float(x.abs().max()) + 1e-6, ct
    scaled = get_qcomm_codec(CommType.FP16, 128.0, None)
    torch.testing.assert_close(scaled.decode(scaled.encode(x)).float(), x, rtol=2e-3, atol=2e-3)
    assert get_qcomm_codecs_registry(QCommsConfig()) is None
    reg = get_qcomm_codecs_registry(QCommsConfig(forward_precision=CommType.INT8, backward_precision=CommType.BF16), device=torch.device("cpu"))
    assert set(reg) == {CommOp.POOLED_EMBEDDINGS_ALL_TO_ALL.name, CommOp.POOLED_EMBEDDINGS_REDUCE_SCATTER.name, CommOp.SEQUENCE_EMBEDDINGS_ALL_TO_ALL.name}
    rs = reg[CommOp.POOLED_EMBEDDINGS_REDUCE_SCATTER.name]
    assert rs.forward.quantized_dtype == torch.float16 and rs.backward.quantized_dtype == torch.float16   # no int8 sums, no bf16 on gloo
    with pytest.raises(ValueError):
        QCommsConfig(forward_precision=CommType.FP16, fp8_quantize_dim=32)


def _qcomm_sharded(ctx):
    import torch.distributed as dist

    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import collective_utils as CU
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder, ShardedEmbeddingBagCollection
    from torchrec_b200.parallel.qcomm_codec import CommType, QCommsConfig, get_qcomm_codecs_registry
    from torchrec_b200.parallel.shard import shard, shard_modules
    from torchrec_b200.parallel.types import ShardingEnv, ShardingPlan
    from torchrec_b200.sparse import KeyedJaggedTensor

    W, dev = ctx.world_size, ctx.device

    # ---- collective_utils ------------------------------------------------------------------------------------------------------------------
    calls = []

    def plan_like():
        calls.append(dist.get_rank())
        return {"made_on": dist.get_rank(), "payload": [1, 2, 3]}

    res = CU.invoke_on_rank_and_broadcast_result(dist.group.WORLD, 1, plan_like)
    assert res == {"made_on": 1, "payload": [1, 2, 3]} and calls == ([1] if ctx.rank == 1 else [])
    assert CU.is_leader(dist.group.WORLD, 0) == (ctx.rank == 0)
    shared = CU.create_on_rank_and_share_result(dist.group.WORLD, 0, lambda: torch.arange(6.0))
    pair = CU.create_on_rank_and_share_result(dist.group.WORLD, 0, lambda n: {"a": torch.ones(n), "b": torch.zeros(2)}, lambda d: [d["a"], d["b"]],
                                              lambda ts: {"a": ts[0], "b": ts[1]}, 3)
    assert pair["a"].tolist() == [1.0, 1.0, 1.0] and pair["b"].tolist() == [0.0, 0.0]
    assert torch.equal(shared, torch.arange(6.0))

    # ---- quantized comms: a sharded EBC with fp16 forward / int8... gradients stays close to the fp32 one ---------------------------------
    def build(qcomms):
        torch.manual_seed(0)
        ebc = EmbeddingBagCollection([EmbeddingBagConfig(name="t0", embedding_dim=32, num_embeddings=40, feature_names=["f0"]),
                                      EmbeddingBagConfig(name="t1", embedding_dim=32, num_embeddings=50, feature_names=["f1"])])
        apply_optimizer_in_backward(torch.optim.SGD, ebc.parameters(), {"lr": 0.1})
        reg = get_qcomm_codecs_registry(qcomms, device=dev) if qcomms is not None else None
        sharder = EmbeddingBagCollectionSharder(qcomm_codecs_registry=reg)
        plan = sp.construct_module_sharding_plan(ebc, {"t0": sp.table_wise(rank=0), "t1": sp.row_wise()}, sharder=sharder, world_size=W, local_size=W, device_type=dev.type)
        return shard(ebc, plan, env=ShardingEnv.from_process_group(dist.group.WORLD), device=dev, sharder=sharder)

    exact = build(None)
    quant = build(QCommsConfig(forward_precision=CommType.FP16, backward_precision=CommType.BF16))
    assert isinstance(exact, ShardedEmbeddingBagCollection)
    g = torch.Generator().manual_seed(3 + ctx.rank)
    lengths = torch.randint(1, 4, (2 * 6,), generator=g)
    kjt = KeyedJaggedTensor(keys=["f0", "f1"], values=torch.randint(0, 40, (int(lengths.sum()),), generator=g), lengths=lengths)
    a, b = exact(kjt).values(), quant(kjt).values()
    torch.testing.assert_close(b, a, rtol=2e-3, atol=2e-3)        # fp16 on the wire
    a.sum().backward()
    b.sum().backward()
    for ta, tb in zip(exact._engine._tbes, quant._engine._tbes):
        if hasattr(ta, "weights"):
            torch.testing.assert_close(tb.weights.detach(), ta.weights.detach(), rtol=2e-2, atol=2e-2)

    # ---- shard_modules: the DMP-less composable entry point walks the module tree with a ShardingPlan -----------------------------------------
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(1)
            self.ebc = EmbeddingBagCollection([EmbeddingBagConfig(name="u", embedding_dim=8, num_embeddings=30, feature_names=["f0"])])
            self.lin = torch.nn.Linear(8, 1)

        def forward(self, kjt):
            return self.lin(self.ebc(kjt).values())

    m = M()
    plan = ShardingPlan({"ebc": sp.construct_module_sharding_plan(m.ebc, {"u": sp.row_wise()}, sharder=EmbeddingBagCollectionSharder(), world_size=W, local_size=W,
                                                                  device_type=dev.type)})
    sm = shard_modules(m, env=ShardingEnv.from_process_group(dist.group.WORLD), device=dev, plan=plan, sharders=[EmbeddingBagCollectionSharder()])
    assert isinstance(sm.ebc, ShardedEmbeddingBagCollection) and isinstance(sm.lin, torch.nn.Linear)
    one = KeyedJaggedTensor(keys=["f0"], values=torch.randint(0, 30, (6,), generator=g), lengths=torch.ones(6, dtype=torch.int64))
    assert sm(one).shape == (6, 1)


def test_qcomms_shard_api_and_collective_utils():
    run_multi_process(_qcomm_sharded, world_size=2, backend="gloo")


def test_topology_group_sizes_and_storage_accounting(monkeypatch):
    import torch

    from torchrec_b200.parallel import comm
    from torchrec_b200.parallel.types import (ComputeDevice, ComputeKernel, DeviceToHostTensorAwaitable, LazyAwaitable, LazyGetItemMixin, ParameterStorage, StorageUsageType,
                                              compute_storage_usage)

    monkeypatch.setenv("LOCAL_WORLD_SIZE", "4")
    monkeypatch.delenv("TOPOLOGY_DOMAIN_MULTIPLE", raising=False)
    assert comm.get_topology_domain_multiple() is None and comm.get_topology_group_world_size(16) == 4 and comm.get_node_group_size(16) == 4
    monkeypatch.setenv("TOPOLOGY_DOMAIN_MULTIPLE", "2")
    assert comm.get_topology_domain_multiple() == 2 and comm.get_topology_group_world_size(16) == 8
    with pytest.raises(ValueError):
        comm.get_topology_group_world_size(12)
    t = torch.empty(10, 4)
    assert compute_storage_usage(t, "cuda", "fused", StorageUsageType.BASE) == {"hbm": 160}
    assert compute_storage_usage(t, "cuda", "fused_uvm_caching", StorageUsageType.BASE) == {"ddr": 160}
    assert compute_storage_usage(t, "cuda", "quant", StorageUsageType.BASE_QUANT) == {"hbm": 200} and compute_storage_usage(t, "cpu", "dense", StorageUsageType.DEFAULT) == {"ddr": 160}
    assert ParameterStorage.HBM.value == "hbm" and ComputeDevice.CUDA.value == "cuda" and ComputeKernel.DEFAULT.value == "default"

    class Out(LazyGetItemMixin, LazyAwaitable):
        def __init__(self):
            super().__init__()
            self.waits = 0

        def _wait_impl(self):
            self.waits += 1
            return {"a": torch.ones(2)}

    out = Out()
    item = out["a"]
    assert out.waits == 0 and (item + 1).tolist() == [2.0, 2.0] and out.waits == 1
    assert DeviceToHostTensorAwaitable(torch.arange(3)).wait().tolist() == [0, 1, 2]


def test_plan_shape_bucket_and_precision_helpers():
    import torch

    from torchrec_b200.modules.embedding_configs import DataType
    from torchrec_b200.ops.tbe import TableBatchedEmbeddingBags
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.types import EmbeddingEvent, ParameterSharding, ShardingType, ShardMetadata
    from torchrec_b200.parallel.types import EnumerableShardingSpec
    from torchrec_b200.parallel.utils import (EmbeddingQuantizationUtils, convert_to_fbgemm_types, create_global_tensor_shape_stride_from_metadata,
                                              get_bucket_metadata_from_shard_metadata, maybe_annotate_embedding_event)

    rw = [ShardMetadata([0, 0], [40, 8], "rank:0/cpu"), ShardMetadata([40, 0], [60, 8], "rank:1/cpu")]
    m = get_bucket_metadata_from_shard_metadata(rw, 10)
    assert (m.num_buckets_per_shard, m.bucket_offsets_per_shard, m.bucket_size) == ([4, 6], [0, 4], 10)
    with pytest.raises(AssertionError):
        get_bucket_metadata_from_shard_metadata(rw, 8)  # buckets would straddle the shard boundary

    def ps(st, shards):
        return ParameterSharding(sharding_type=st, compute_kernel="fused", ranks=list(range(len(shards))), sharding_spec=EnumerableShardingSpec(shards))

    assert create_global_tensor_shape_stride_from_metadata(ps("row_wise", rw)) == (torch.Size([100, 8]), (8, 1))
    cw = [ShardMetadata([0, 0], [100, 16], "rank:0/cpu"), ShardMetadata([0, 16], [100, 16], "rank:1/cpu")]
    assert create_global_tensor_shape_stride_from_metadata(ps("column_wise", cw)) == (torch.Size([100, 32]), (32, 1))
    grid = [ShardMetadata([r * 50, c * 16], [50, 16], f"rank:{c * 2 + r}/cpu") for c in range(2) for r in range(2)]
    assert create_global_tensor_shape_stride_from_metadata(ps("grid_shard", grid), devices_per_node=2)[0] == torch.Size([100, 32])
    assert create_global_tensor_shape_stride_from_metadata(ps("table_wise", [ShardMetadata([0, 0], [7, 4], "rank:0/cpu")]))[0] == torch.Size([7, 4])
    assert convert_to_fbgemm_types({"cache_precision": DataType.FP16, "lr": 0.1}) == {"cache_precision": "fp16", "lr": 0.1}
    with maybe_annotate_embedding_event(EmbeddingEvent.LOOKUP, "sparse.ebc", "row_wise"):
        pass
    with maybe_annotate_embedding_event(EmbeddingEvent.LOOKUP, None, None):
        pass
    assert sp.get_sharding_constructor_from_type(ShardingType.TABLE_ROW_WISE) is sp.table_row_wise and sp.placement_helper("cuda", 1, 3) == "rank:3/cuda:1"
    # narrow the kernels of a model and restore them
    t = TableBatchedEmbeddingBags([(10, 4), (5, 4)])
    w0 = t.weights.detach().clone()
    u = EmbeddingQuantizationUtils()
    u.quantize_embedding_modules(t, DataType.FP16)
    assert t.weights.dtype == torch.float16
    u.recreate_embedding_modules(t)
    assert t.weights.dtype == torch.float32 and float((t.weights - w0).abs().max()) < 1e-3
