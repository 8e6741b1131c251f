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
