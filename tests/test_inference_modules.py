import pytest
import torch

from torchrec_b200.inference import quantize_inference_model, shard_quant_model
from torchrec_b200.models.dlrm import DLRM
from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig, PoolingType
from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
from torchrec_b200.sparse import KeyedJaggedTensor
from torchrec_b200.types import DataType


def test_quantize_and_shard_dlrm_for_inference_cpu():
    torch.manual_seed(0)
    keys = ["a", "b", "c"]
    ebc = EmbeddingBagCollection([EmbeddingBagConfig(name=f"t{i}", embedding_dim=32, num_embeddings=50 + 10 * i, feature_names=[k],
                                                     pooling=PoolingType.MEAN if i == 1 else PoolingType.SUM) for i, k in enumerate(keys)])
    model = DLRM(ebc, 4, [16, 32], [16, 1]).eval()
    kjt = KeyedJaggedTensor(keys=keys, values=torch.randint(0, 50, (14,)), lengths=torch.tensor([2, 3, 1, 4, 2, 2]))
    x = torch.randn(2, 4)
    ref = model(x, kjt)
    q = quantize_inference_model(model, per_table_weight_dtype={"t2": DataType.FP16})
    out = q(x, kjt)
    assert float((out - ref).detach().abs().max()) < 0.05
    for ws in (1, 2):
        sharded, plan = shard_quant_model(quantize_inference_model(DLRM(EmbeddingBagCollection(ebc.embedding_bag_configs()), 4, [16, 32], [16, 1]).eval()),
                                          world_size=ws, compute_device="cpu")
        assert "sparse_arch.embedding_bag_collection" in plan.plan
    # sharded quantized module reproduces the unsharded quantized one
    q2 = quantize_inference_model(DLRM(EmbeddingBagCollection(ebc.embedding_bag_configs()), 4, [16, 32], [16, 1]).eval(), per_table_weight_dtype={"t2": DataType.FP16})
    q2.load_state_dict(q.state_dict(), strict=False)
    unsharded = q2.sparse_arch.embedding_bag_collection(kjt).values().clone()
    sharded, _ = shard_quant_model(q2, world_size=2, compute_device="cpu")
    got = sharded.sparse_arch.embedding_bag_collection(kjt).values()
    assert float((got - unsharded).abs().max()) < 0.05


def test_dlrm_predict_factory_packager_and_request(tmp_path):
    import numpy as np
    import torch

    from torchrec_b200.inference.client import create_request
    from torchrec_b200.inference.dlrm_predict import DLRMModelConfig, DLRMPredictFactory, create_training_batch
    from torchrec_b200.inference.model_packager import PredictFactoryPackager, load_config_text, load_pickle_config, load_predict_factory

    keys = ["cat_0", "cat_1", "cat_2"]
    cfg = DLRMModelConfig(dense_arch_layer_sizes=[16, 8], dense_in_features=4, embedding_dim=8, id_list_features_keys=keys, num_embeddings_per_feature=[50, 60, 70],
                          num_embeddings=100, over_arch_layer_sizes=[16, 1])
    factory = DLRMPredictFactory(cfg)
    assert set(factory.batching_metadata()) == {"float_features", "id_list_features"} and "sparse" in factory.batching_metadata_json()
    module = factory.create_predict_module(world_size=1, device="cpu")
    batch = create_training_batch(4, keys, 50, batch_size=6, ids_per_feature=2)
    out = module({"float_features": batch.dense_features, "id_list_features.lengths": batch.sparse_features.lengths(),
                  "id_list_features.values": batch.sparse_features.values()})
    assert out["default"].shape == (6,) and bool(((out["default"] >= 0) & (out["default"] <= 1)).all())

    archive = tmp_path / "dlrm.zip"
    PredictFactoryPackager.save_predict_factory(DLRMPredictFactory, {"model_config": cfg}, archive, extra_files={"readme.txt": "dlrm int8"},
                                                state_dict={"w": torch.ones(2)})
    again = load_predict_factory(archive)
    assert isinstance(again, DLRMPredictFactory) and again.model_config.id_list_features_keys == keys
    assert load_config_text(archive, "readme.txt") == "dlrm int8" and load_pickle_config(archive, "model_config").embedding_dim == 8

    req = create_request(batch, num_dense=4, num_id_list_features=3)
    assert req.batch_size == 6 and req.float_features.num_features == 4
    assert np.frombuffer(req.id_list_features.lengths, dtype=np.int32).tolist() == batch.sparse_features.lengths().tolist()
    assert np.frombuffer(req.float_features.values, dtype=np.float32).size == 24


def _sd_transform(ctx):
    import torch
    import torch.distributed as dist
    from torch.distributed._shard.sharded_tensor import Shard, ShardedTensor

    from torchrec_b200.inference.state_dict_transform import state_dict_all_gather_keys, state_dict_gather, state_dict_to_device

    r, W = ctx.rank, ctx.world_size
    local = torch.full((2, 4), float(r + 1))
    st = ShardedTensor._init_from_local_shards([Shard.from_tensor_and_offsets(local, [2 * r, 0], r)], (2 * W, 4), process_group=dist.group.WORLD)
    sd = {"emb.weight": st, "dense.bias": torch.arange(3.0)}
    if r == 0:
        sd["only_rank0"] = torch.zeros(1)
    assert state_dict_all_gather_keys(sd, dist.group.WORLD) == ["dense.bias", "emb.weight", "only_rank0"]
    moved = state_dict_to_device(sd, dist.group.WORLD, torch.device("cpu"))
    assert isinstance(moved["emb.weight"], ShardedTensor) and torch.equal(moved["emb.weight"].local_shards()[0].tensor, local)
    dst = {"emb.weight": torch.zeros(2 * W, 4), "dense.bias": torch.zeros(3)}
    state_dict_gather(sd, dst)
    assert torch.equal(dst["dense.bias"], torch.arange(3.0))
    if r == 0:
        assert torch.equal(dst["emb.weight"][:, 0], torch.tensor([1.0, 1.0, 2.0, 2.0]))


def test_state_dict_transform_gloo():
    from torchrec_b200.utils.multiprocess import run_multi_process

    run_multi_process(_sd_transform, world_size=2, backend="gloo")


def test_quant_state_specs():
    import torch

    from torchrec_b200.inference.modules import quantize_inference_model, shard_quant_model
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.parallel.planner.types import ParameterConstraints
    from torchrec_b200.parallel.quant_state import ShardedQuantEmbeddingModuleState, sharded_tbes_weights_spec
    from torchrec_b200.parallel.shards_wrapper import LocalShardsWrapper

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ebc = EmbeddingBagCollection([EmbeddingBagConfig(name="a", embedding_dim=16, num_embeddings=40, feature_names=["fa"]),
                                               EmbeddingBagConfig(name="b", embedding_dim=16, num_embeddings=64, feature_names=["fb"])])

        def forward(self, kjt):
            return self.ebc(kjt).values()

    q = quantize_inference_model(M())
    sharded, _ = shard_quant_model(q, world_size=2, compute_device="cpu", sharding_device="cpu",
                                   constraints={"a": ParameterConstraints(sharding_types=["table_wise"]), "b": ParameterConstraints(sharding_types=["row_wise"])})
    specs = sharded_tbes_weights_spec(sharded)
    by_table = {}
    for k, s in specs.items():
        assert k.startswith("ebc.tbes.") and s.fqn.startswith("ebc.embedding_bags.")
        by_table.setdefault(s.fqn.split(".")[2], []).append(s)
    assert len(by_table["a"]) == 1 and by_table["a"][0].sharding_type == "table_wise" and by_table["a"][0].shard_sizes == [40, 16]
    assert len(by_table["b"]) == 2 and {s.sharding_type for s in by_table["b"]} == {"row_wise"} and sorted(s.shard_offsets[0] for s in by_table["b"]) == [0, 32]
    sd = ShardedQuantEmbeddingModuleState.sharded_state_dict(sharded.ebc, prefix="ebc.")
    assert sd["ebc.embedding_bags.a.weight"].dtype == torch.uint8 and sd["ebc.embedding_bags.a.weight"].shape[0] == 40
    assert isinstance(sd["ebc.embedding_bags.b.weight"], LocalShardsWrapper) and sd["ebc.embedding_bags.b.weight"].local_offsets() == [(0, 0), (32, 0)]


@pytest.mark.parametrize("placement", ["table_wise", "row_wise", "column_wise", "table_row_wise", "mixed"])
def test_sharded_quant_embedding_collection_every_placement(placement):
    """Sequence quantized lookups sharded TW / RW / CW / TWRW over 4 (virtual) devices == the unsharded quantized module
    (reference distributed/quant_embedding.py:597, tests test_quant_sequence_model_parallel.py)."""
    import torch

    from torchrec_b200.modules.embedding_configs import DataType, EmbeddingConfig
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.quant_embedding import QuantEmbeddingCollectionSharder
    from torchrec_b200.parallel.types import ShardingEnv
    from torchrec_b200.quant.embedding_modules import EmbeddingCollection as QuantEC
    from torchrec_b200.sparse.jagged_tensor import KeyedJaggedTensor

    torch.manual_seed(3)
    W = 4
    tables = [EmbeddingConfig(name="a", embedding_dim=64, num_embeddings=103, feature_names=["fa", "fa2"], data_type=DataType.FP16),
              EmbeddingConfig(name="b", embedding_dim=64, num_embeddings=57, feature_names=["fb"], data_type=DataType.FP16)]
    from torchrec_b200.ops.quant_tbe import quantize_rows

    qw = {t.name: (quantize_rows(torch.randn(t.num_embeddings, t.embedding_dim), t.data_type), None) for t in tables}
    qec = QuantEC(tables, device=torch.device("cpu"), need_indices=True, table_name_to_quantized_weights=qw)
    gens = {"table_wise": {"a": sp.table_wise(rank=1), "b": sp.table_wise(rank=3)}, "row_wise": {"a": sp.row_wise(), "b": sp.row_wise()},
            "column_wise": {"a": sp.column_wise(ranks=[0, 2]), "b": sp.column_wise(ranks=[3, 1])},
            "table_row_wise": {"a": sp.table_row_wise(host_index=0), "b": sp.table_row_wise(host_index=1)},
            "mixed": {"a": sp.row_wise(), "b": sp.column_wise(ranks=[2, 0])}}[placement]
    sharder = QuantEmbeddingCollectionSharder()
    plan = sp.construct_module_sharding_plan(qec, gens, sharder=sharder, world_size=W, local_size=2, device_type="cpu")
    sharded = sharder.shard(qec, plan, ShardingEnv.from_local(W, 0), device=torch.device("cpu"))
    lengths = torch.tensor([2, 0, 1, 3, 1, 1, 0, 2, 2, 1, 1, 1])  # 3 features x batch 4
    n = int(lengths.sum())
    values = torch.cat([torch.randint(0, 103, (int(lengths[:8].sum()),)), torch.randint(0, 57, (int(lengths[8:].sum()),))])
    assert values.numel() == n
    kjt = KeyedJaggedTensor(keys=["fa", "fa2", "fb"], values=values, lengths=lengths)
    want, got = qec(kjt), sharded(kjt)
    assert set(got.keys()) == {"fa", "fa2", "fb"}
    for k in want:
        torch.testing.assert_close(got[k].values().float(), want[k].values().float(), rtol=2e-3, atol=2e-3)
        assert torch.equal(got[k].lengths(), want[k].lengths())
        assert torch.equal(got[k].weights(), want[k].weights())  # need_indices: the ids ride along as weights


def test_dlrm_packager_cli_round_trip(tmp_path):
    """``dlrm_packager`` writes an archive that ``load_predict_factory`` turns back into a working predict module
    (reference inference/dlrm_packager.py + inference_legacy import paths)."""
    import torch

    from torchrec_b200.inference import dlrm_packager
    from torchrec_b200.inference.inference_legacy import model_packager as legacy_packager
    from torchrec_b200.inference.model_packager import load_predict_factory

    assert legacy_packager.load_predict_factory is load_predict_factory
    out = tmp_path / "dlrm.zip"
    dlrm_packager.main(["--output_path", str(out), "--num_embeddings_per_feature", "50,60,70", "--sparse_feature_names", "a,b,c", "--embedding_dim", "8",
                        "--dense_arch_layer_sizes", "16,8", "--over_arch_layer_sizes", "16,1", "--num_dense_features", "4", "--weight_dtype", "INT8"])
    factory = load_predict_factory(out)
    cfg = factory.model_config
    assert cfg.id_list_features_keys == ["a", "b", "c"] and cfg.num_embeddings_per_feature == [50, 60, 70] and cfg.sample_input.dense_features.shape == (2, 4)
    module = factory.create_predict_module(world_size=1, device="cpu")
    b = cfg.sample_input
    res = module.predict_forward({"float_features": b.dense_features, "id_list_features.lengths": b.sparse_features.lengths(), "id_list_features.values": b.sparse_features.values()})
    assert res["default"].shape == (2,) and bool(((res["default"] >= 0) & (res["default"] <= 1)).all())
    with pytest.raises(ValueError):
        dlrm_packager.main(["--output_path", str(out), "--num_embeddings_per_feature", "50,60", "--sparse_feature_names", "a,b,c"])


def test_quantized_weight_publish_and_pruning_metadata():
    import torch

    from torchrec_b200.inference.modules import (MODULE_ATTR_EMB_CONFIG_NAME_TO_NUM_ROWS_POST_PRUNING_DICT, PredictFactory, QualNameMetadata, assign_weights_to_tbe,
                                                 get_table_to_weights_from_tbe, set_pruning_data)
    from torchrec_b200.modules.embedding_configs import DataType, EmbeddingConfig
    from torchrec_b200.ops.quant_tbe import quantize_rows
    from torchrec_b200.quant.embedding_modules import EmbeddingCollection as QuantEC

    tables = [EmbeddingConfig(name="a", embedding_dim=16, num_embeddings=30, feature_names=["fa"], data_type=DataType.INT8)]
    qw = {"a": (quantize_rows(torch.randn(30, 16), DataType.INT8), None)}
    qec = QuantEC(tables, device=torch.device("cpu"), table_name_to_quantized_weights=qw)
    got = get_table_to_weights_from_tbe(qec)
    assert set(got) == {"a"} and torch.equal(got["a"], qw["a"][0])
    fresh = quantize_rows(torch.randn(30, 16), DataType.INT8)
    assign_weights_to_tbe(qec, {"a": fresh})
    assert torch.equal(get_table_to_weights_from_tbe(qec)["a"], fresh)
    with pytest.raises(AssertionError):
        assign_weights_to_tbe(qec, {"a": fresh[:10]})

    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection

    ebc = EmbeddingBagCollection([EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=100, feature_names=["f0"])])
    set_pruning_data(torch.nn.Sequential(ebc), {"t0": 40})
    assert getattr(ebc, MODULE_ATTR_EMB_CONFIG_NAME_TO_NUM_ROWS_POST_PRUNING_DICT) == {"t0": 40} and ebc.embedding_bag_configs()[0].num_embeddings_post_pruning == 40

    class F(PredictFactory):
        def create_predict_module(self):
            return torch.nn.Identity()

        def batching_metadata(self):
            return {}

        def result_metadata(self):
            return "dict_of_tensor"

        def run_weights_independent_tranformations(self, m):
            return m

        def run_weights_dependent_transformations(self, m):
            return m

        def qualname_metadata(self):
            return {"model.preproc": QualNameMetadata(need_preproc=True)}

    assert F().qualname_metadata_json() == '{"model.preproc": {"need_preproc": true}}'


@pytest.mark.parametrize("seed", list(range(8)))
def test_sharded_quant_ebc_random_placements(seed):
    """Quantized pooled lookups: random tables (row formats INT8 / INT4 / FP16 / FP8, SUM / MEAN, shared tables), random TW / RW / CW
    placements over 4 virtual devices - sharded output == unsharded quantized module (re-quantisation of column slices bounds the error)."""
    import random

    import torch

    from torchrec_b200.modules.embedding_configs import DataType, EmbeddingBagConfig, PoolingType
    from torchrec_b200.ops.quant_tbe import quantize_rows
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.quant_embeddingbag import QuantEmbeddingBagCollectionSharder
    from torchrec_b200.parallel.types import ShardingEnv
    from torchrec_b200.quant.embedding_modules import EmbeddingBagCollection as QuantEBC
    from torchrec_b200.sparse.jagged_tensor import KeyedJaggedTensor

    rng = random.Random(seed)
    torch.manual_seed(seed)
    W, B = 4, 5
    tables, gens, keys, hashes, qw, fid = [], {}, [], [], {}, 0
    for t in range(rng.randint(2, 5)):
        dt = rng.choice([DataType.INT8, DataType.FP16, DataType.INT4, DataType.FP8])
        dim = rng.choice([32, 64])
        rows = rng.randint(9, 60)
        feats = [f"f{fid + i}" for i in range(rng.choice([1, 1, 2]))]
        fid += len(feats)
        cfg = EmbeddingBagConfig(name=f"t{t}", embedding_dim=dim, num_embeddings=rows, feature_names=feats, data_type=dt, pooling=rng.choice([PoolingType.SUM, PoolingType.MEAN]))
        tables.append(cfg)
        keys += feats
        hashes += [rows] * len(feats)
        qw[cfg.name] = (quantize_rows(torch.randn(rows, dim) * 0.5, dt), None)
        kind = rng.choice(["tw", "rw", "cw"])
        if kind == "cw" and dt == DataType.FP8 and dim < 64:
            kind = "tw"  # fp8 column shards must be multiples of the 32-element scale block (the sharder says so: checked below)
        gens[cfg.name] = sp.table_wise(rank=rng.randrange(W)) if kind == "tw" else (sp.row_wise() if kind == "rw" else sp.column_wise(ranks=rng.sample(range(W), 2)))
    weighted = rng.random() < 0.3
    q = QuantEBC(tables, is_weighted=weighted, device=torch.device("cpu"), table_name_to_quantized_weights=qw)
    sharder = QuantEmbeddingBagCollectionSharder()
    plan = sp.construct_module_sharding_plan(q, gens, sharder=sharder, world_size=W, local_size=W, device_type="cpu")
    sharded = sharder.shard(q, plan, ShardingEnv.from_local(W, 0), device=torch.device("cpu"))
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(0, 4, (len(keys) * B,), generator=g)
    values = torch.cat([torch.randint(0, hashes[i], (int(lengths[i * B : (i + 1) * B].sum()),), generator=g) for i in range(len(keys))])
    kjt = KeyedJaggedTensor(keys=keys, values=values, lengths=lengths, weights=torch.rand(values.numel(), generator=g) if weighted else None)
    want, got = q(kjt), sharded(kjt)
    assert got.keys() == want.keys()
    desc = {n: (plan[n].sharding_type, str(t.data_type), str(t.pooling)) for n, t in zip(plan, tables)}
    # column slices are re-quantised with their own row scale: int4 steps are coarse, so the tolerance follows the format
    tol = 0.2 if any(t.data_type == DataType.INT4 and plan[t.name].sharding_type == "column_wise" for t in tables) else 3e-2
    torch.testing.assert_close(got.values().float(), want.values().float(), rtol=tol, atol=tol, msg=lambda m: f"seed {seed} {desc}: {m}")
    if seed == 0:  # a 32-wide fp8 table cannot be split into 16-wide column shards
        bad = [EmbeddingBagConfig(name="b", embedding_dim=32, num_embeddings=10, feature_names=["fb"], data_type=DataType.FP8)]
        qb = QuantEBC(bad, is_weighted=False, device=torch.device("cpu"), table_name_to_quantized_weights={"b": (quantize_rows(torch.randn(10, 32), DataType.FP8), None)})
        pb = sp.construct_module_sharding_plan(qb, {"b": sp.column_wise(ranks=[0, 1])}, sharder=sharder, world_size=W, local_size=W, device_type="cpu")
        with pytest.raises(ValueError, match="multiples of 32"):
            sharder.shard(qb, pb, ShardingEnv.from_local(W, 0), device=torch.device("cpu"))


def test_shard_quantized_fp_and_mc_collections():
    """The quantized FP-EBC / MC-EC / MC-EBC shard over two local devices with the default inference sharders and give the unsharded
    quantized module's results."""
    import copy

    import torch

    from torchrec_b200.inference.modules import quantize_embeddings, shard_quant_model
    from torchrec_b200.modules.embedding_configs import DataType, EmbeddingBagConfig, EmbeddingConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection, EmbeddingCollection
    from torchrec_b200.modules.feature_processor_ import PositionWeightedModuleCollection
    from torchrec_b200.modules.fp_embedding_modules import FeatureProcessedEmbeddingBagCollection
    from torchrec_b200.modules.mc_embedding_modules import ManagedCollisionEmbeddingBagCollection, ManagedCollisionEmbeddingCollection
    from torchrec_b200.modules.mc_modules import DistanceLFU_EvictionPolicy, ManagedCollisionCollection, MCHManagedCollisionModule
    from torchrec_b200.sparse import KeyedJaggedTensor

    torch.manual_seed(0)
    bag_cfgs = [EmbeddingBagConfig(name=f"t{i}", embedding_dim=16, num_embeddings=64, feature_names=[f"f{i}"]) for i in range(2)]
    seq_cfgs = [EmbeddingConfig(name=f"t{i}", embedding_dim=16, num_embeddings=64, feature_names=[f"f{i}"]) for i in range(2)]
    fp = PositionWeightedModuleCollection({"f0": 4, "f1": 4})
    with torch.no_grad():
        for p in fp.parameters():
            p.copy_(torch.linspace(0.5, 1.5, p.numel()))

    def mcc(cfgs):
        return ManagedCollisionCollection({c.name: MCHManagedCollisionModule(zch_size=64, device=torch.device("cpu"), eviction_policy=DistanceLFU_EvictionPolicy(),
                                                                            eviction_interval=2, input_hash_size=10**9) for c in cfgs}, cfgs)

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fp = FeatureProcessedEmbeddingBagCollection(EmbeddingBagCollection(bag_cfgs, is_weighted=True), fp)
            self.mc_ec = ManagedCollisionEmbeddingCollection(EmbeddingCollection(seq_cfgs), mcc(seq_cfgs), return_remapped_features=True)
            self.mc_ebc = ManagedCollisionEmbeddingBagCollection(EmbeddingBagCollection(bag_cfgs), mcc(bag_cfgs), return_remapped_features=True)

        def forward(self, small, raw):
            a = self.fp(small).values()
            b, rb = self.mc_ec(raw)
            c, rc = self.mc_ebc(raw)
            return a, {k: v.values() for k, v in b.items()}, c.values(), rb.values(), rc.values()

    small = KeyedJaggedTensor(keys=["f0", "f1"], values=torch.randint(0, 64, (12,)), lengths=torch.tensor([3, 1, 2, 2, 0, 4]))
    raw = KeyedJaggedTensor(keys=["f0", "f1"], values=torch.randint(0, 10**6, (12,)), lengths=torch.tensor([3, 1, 2, 2, 0, 4]))
    m = M()
    m.train()
    for _ in range(6):
        m(small, raw)
    m.eval()
    q = quantize_embeddings(m, DataType.INT8, inplace=False)
    want = q(small, raw)
    sharded, plan = shard_quant_model(copy.deepcopy(q), world_size=2, compute_device="cpu", sharding_device="cpu")
    assert set(plan.plan.keys()) == {"fp", "mc_ec", "mc_ebc"}
    assert type(sharded.fp).__name__ == "ShardedQuantFeatureProcessedEmbeddingBagCollection" and type(sharded.mc_ec).__name__ == "ShardedQuantManagedCollisionEmbeddingCollection"
    got = sharded(small, raw)
    torch.testing.assert_close(got[0], want[0])
    for k in want[1]:
        torch.testing.assert_close(got[1][k], want[1][k])
    torch.testing.assert_close(got[2], want[2])
    assert torch.equal(got[3], want[3]) and torch.equal(got[4], want[4])
