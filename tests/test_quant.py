"""Quantized embedding modules vs float modules (CPU reference path; GPU kernel covered in test_quant_gpu)."""
import pytest
import torch

from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig, EmbeddingConfig, PoolingType, QuantConfig
from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection, EmbeddingCollection
from torchrec_b200.ops import quant_tbe as Q
from torchrec_b200.quant import EmbeddingBagCollection as QEBC
from torchrec_b200.quant import EmbeddingCollection as QEC
from torchrec_b200.sparse import KeyedJaggedTensor
from torchrec_b200.types import DataType


@pytest.mark.parametrize("dt,tol", [(DataType.FP16, 2e-3), (DataType.INT8, 2e-2), (DataType.INT4, 0.2), (DataType.INT2, 0.6), (DataType.FP8, 0.08), (DataType.BF16, 1e-2)])
def test_quantize_dequantize_rows(dt, tol):
    w = torch.randn(37, 64)
    q = Q.quantize_rows(w, dt)
    assert q.shape == (37, Q.row_bytes(64, dt)) and q.shape[1] % 16 == 0
    d = Q.dequantize_rows(q, 64, dt)
    assert float((d - w).abs().max()) < tol * float(w.abs().max()) + 1e-3


def _kjt():
    return KeyedJaggedTensor(keys=["f0", "f1", "f2"], values=torch.tensor([1, 2, 3, 0, 5, 7, 7, 2]), lengths=torch.tensor([2, 0, 1, 1, 3, 1]))


def test_quant_ebc_from_float_matches():
    tables = [EmbeddingBagConfig(name="t0", embedding_dim=32, num_embeddings=10, feature_names=["f0", "f1"]),
              EmbeddingBagConfig(name="t1", embedding_dim=64, num_embeddings=8, feature_names=["f2"], pooling=PoolingType.MEAN)]
    ebc = EmbeddingBagCollection(tables)
    ref = ebc(_kjt())
    ebc.qconfig = QuantConfig(activation=torch.float32, weight=torch.qint8, per_table_weight_dtype={"t1": torch.float16})
    q = QEBC.from_float(ebc)
    out = q(_kjt())
    assert out.keys() == ref.keys() and out.length_per_key() == ref.length_per_key()
    torch.testing.assert_close(out.values(), ref.values(), rtol=0.05, atol=0.02)
    sd = q.state_dict()
    assert sd["embedding_bags.t0.weight"].dtype == torch.uint8 and sd["embedding_bags.t0.weight"].shape == (10, Q.row_bytes(32, DataType.INT8))


def test_quant_ec_from_float_matches():
    tables = [EmbeddingConfig(name="t0", embedding_dim=32, num_embeddings=10, feature_names=["f0", "f1"]),
              EmbeddingConfig(name="t1", embedding_dim=32, num_embeddings=8, feature_names=["f2"])]
    ec = EmbeddingCollection(tables)
    ref = ec(_kjt())
    ec.qconfig = QuantConfig(activation=torch.float32, weight=DataType.FP8)
    q = QEC.from_float(ec)
    out = q(_kjt())
    for k in ref:
        torch.testing.assert_close(out[k].values(), ref[k].values(), rtol=0.1, atol=0.03)
        assert torch.equal(out[k].lengths(), ref[k].lengths())


def test_quant_utils_fx_names_and_meta_to_cpu():
    import torch

    from torchrec_b200.modules.embedding_configs import DataType, EmbeddingBagConfig
    from torchrec_b200.quant.embedding_modules import EmbeddingBagCollection as QEBC
    from torchrec_b200.quant.utils import meta_to_cpu_placement, recursive_populate_fx_names

    tables = [EmbeddingBagConfig(name="a", embedding_dim=8, num_embeddings=10, feature_names=["fa"], data_type=DataType.INT8),
              EmbeddingBagConfig(name="b", embedding_dim=8, num_embeddings=12, feature_names=["fb"], data_type=DataType.INT8)]

    class M(torch.nn.Module):
        def __init__(self, dev):
            super().__init__()
            self.q = QEBC(tables, is_weighted=False, device=torch.device(dev))

    m = M("cpu")
    recursive_populate_fx_names(m)
    assert m.q._tbes[0]._fx_path == "emb_module.a,b"
    meta = M("meta")
    assert meta.q.device.type == "meta"
    meta_to_cpu_placement(meta)
    assert meta.q.device.type == "cpu" and [c.name for c in meta.q.embedding_bag_configs()] == ["a", "b"]


def test_quantized_feature_processed_and_managed_collision_collections():
    """``quantize_embeddings`` swaps FP-EBC / MC-EC / MC-EBC for their quantized versions: position weights still apply, raw ids are
    remapped exactly like in the trained float module, serving unseen ids never changes the collision state."""
    import copy

    from torchrec_b200.inference.modules import quantize_embeddings
    from torchrec_b200.modules.embedding_configs import DataType, EmbeddingBagConfig, EmbeddingConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection, EmbeddingCollection
    from torchrec_b200.modules.feature_processor_ import PositionWeightedModuleCollection
    from torchrec_b200.modules.fp_embedding_modules import FeatureProcessedEmbeddingBagCollection
    from torchrec_b200.modules.mc_embedding_modules import ManagedCollisionEmbeddingBagCollection, ManagedCollisionEmbeddingCollection
    from torchrec_b200.modules.mc_modules import DistanceLFU_EvictionPolicy, ManagedCollisionCollection, MCHManagedCollisionModule
    from torchrec_b200.quant.embedding_modules import (FeatureProcessedEmbeddingBagCollection as QFP, QuantManagedCollisionEmbeddingBagCollection,
                                                       QuantManagedCollisionEmbeddingCollection, for_each_module_of_type_do, quant_prep_enable_cache_features_order,
                                                       quant_prep_enable_quant_state_dict_split_scale_bias_for_types)
    from torchrec_b200.sparse import KeyedJaggedTensor

    torch.manual_seed(0)
    # feature processed bags
    tables = [EmbeddingBagConfig(name="t0", embedding_dim=16, num_embeddings=50, feature_names=["f0"]), EmbeddingBagConfig(name="t1", embedding_dim=16, num_embeddings=40, feature_names=["f1"])]
    fp = PositionWeightedModuleCollection({"f0": 5, "f1": 5})
    with torch.no_grad():
        for p in fp.parameters():
            p.copy_(torch.linspace(0.5, 2.0, p.numel()))
    model = torch.nn.Sequential(FeatureProcessedEmbeddingBagCollection(EmbeddingBagCollection(tables, is_weighted=True), fp))
    kjt = KeyedJaggedTensor(keys=["f0", "f1"], values=torch.tensor([1, 2, 3, 4, 5, 6, 7]), lengths=torch.tensor([3, 1, 0, 3]))
    want = model(kjt).values()
    seen = []
    for_each_module_of_type_do(model, [FeatureProcessedEmbeddingBagCollection], lambda m: seen.append(type(m).__name__))
    assert seen == ["FeatureProcessedEmbeddingBagCollection"]
    quant_prep_enable_cache_features_order(model, [FeatureProcessedEmbeddingBagCollection])
    quant_prep_enable_quant_state_dict_split_scale_bias_for_types(model, [EmbeddingBagCollection])
    q = quantize_embeddings(model, DataType.INT8, inplace=False)
    assert isinstance(q[0], QFP) and q[0]._get_name() == "QuantFeatureProcessedEmbeddingBagCollection"
    got = q(kjt).values()
    assert got.shape == want.shape and float((got - want).abs().max()) < 0.05
    unweighted = FeatureProcessedEmbeddingBagCollection(EmbeddingBagCollection(tables, is_weighted=True), PositionWeightedModuleCollection({"f0": 5, "f1": 5}))
    unweighted._embedding_bag_collection.load_state_dict(model[0]._embedding_bag_collection.state_dict())
    assert float((unweighted(kjt).values() - want).abs().max()) > 0.01, "the position weights matter in this test"

    # managed collision: sequence and bags
    def mc_model(kind):
        if kind == "ec":
            cfgs = [EmbeddingConfig(name="t0", embedding_dim=8, num_embeddings=64, feature_names=["f0"]), EmbeddingConfig(name="t1", embedding_dim=8, num_embeddings=64, feature_names=["f1"])]
            emb, wrap = EmbeddingCollection(cfgs), ManagedCollisionEmbeddingCollection
        else:
            cfgs = [EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=64, feature_names=["f0"]), EmbeddingBagConfig(name="t1", embedding_dim=8, num_embeddings=64, feature_names=["f1"])]
            emb, wrap = EmbeddingBagCollection(cfgs), ManagedCollisionEmbeddingBagCollection
        mcc = ManagedCollisionCollection({c.name: MCHManagedCollisionModule(zch_size=64, device=torch.device("cpu"), eviction_policy=DistanceLFU_EvictionPolicy(), eviction_interval=2,
                                                                            input_hash_size=10**9) for c in cfgs}, cfgs)
        return wrap(emb, mcc, return_remapped_features=True)

    g = torch.Generator().manual_seed(1)
    raw = torch.randint(0, 10**6, (20,), generator=g)
    batch = KeyedJaggedTensor(keys=["f0", "f1"], values=torch.cat([raw[:10], raw[10:]]), lengths=torch.full((10,), 2))
    for kind, qcls in (("ec", QuantManagedCollisionEmbeddingCollection), ("ebc", QuantManagedCollisionEmbeddingBagCollection)):
        m = mc_model(kind)
        m.train()
        for _ in range(6):
            m(batch)  # the ids get admitted
        m.eval()
        out_f, remap_f = m(batch)
        qm = quantize_embeddings(torch.nn.Sequential(copy.deepcopy(m)), DataType.INT8, inplace=True)[0]
        assert isinstance(qm, qcls)
        out_q, remap_q = qm(batch)
        assert torch.equal(remap_q.values(), remap_f.values()), kind
        if kind == "ec":
            for k in out_f:
                assert float((out_q[k].values() - out_f[k].values()).abs().max()) < 0.05
        else:
            assert float((out_q.values() - out_f.values()).abs().max()) < 0.05
        state = {k: v.clone() for k, v in qm._managed_collision_collection.state_dict().items()}
        qm.train()  # a parent switched to train mode must not unfreeze the collision modules
        unseen = KeyedJaggedTensor(keys=["f0", "f1"], values=torch.randint(10**7, 10**8, (20,), generator=g), lengths=torch.full((10,), 2))
        for _ in range(4):
            qm(unseen)
        for k, v in qm._managed_collision_collection.state_dict().items():
            assert torch.equal(v, state[k]), (kind, k)
