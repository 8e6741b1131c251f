"""Sharded quantized inference on GPUs of ONE process: the fused path (kjt_route peer stores + quantized lookup kernels storing into the
result on device 0 + staging reduce) against the unsharded quantized module, for table- / column- / row-wise placements, FP8 block-scaled
and INT8 / INT4 rows, weighted and mean pooling. Runs with however many GPUs the box has (1 GPU: every "device" is cuda:0's only rank)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt_name", ["FP8", "INT8", "INT4"])
@pytest.mark.parametrize("kind", ["tw", "cw", "rw", "mixed"])
@pytest.mark.parametrize("weighted,pooling", [(False, "sum"), (True, "sum"), (False, "mean")])
def test_fused_sharded_quant_matches_unsharded(fmt_name, kind, weighted, pooling):
    from torchrec_b200.inference.modules import quantize_inference_model
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig, PoolingType
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.quant_embeddingbag import QuantEmbeddingBagCollectionSharder, ShardedQuantEmbeddingBagCollection
    from torchrec_b200.parallel.types import ShardingEnv
    from torchrec_b200.sparse import KeyedJaggedTensor
    from torchrec_b200.types import DataType

    n_dev = torch.cuda.device_count()
    W = 2 if n_dev >= 2 else 1
    fmt = getattr(DataType, fmt_name)
    torch.manual_seed(0)
    pt = PoolingType.MEAN if pooling == "mean" else PoolingType.SUM
    tables = [EmbeddingBagConfig(name="a", embedding_dim=128, num_embeddings=300, feature_names=["fa"], pooling=pt),
              EmbeddingBagConfig(name="b", embedding_dim=64, num_embeddings=1000, feature_names=["fb", "fb2"], pooling=pt),
              EmbeddingBagConfig(name="c", embedding_dim=128, num_embeddings=77, feature_names=["fc"], pooling=pt)]

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ebc = EmbeddingBagCollection(tables=tables, is_weighted=weighted)

        def forward(self, k):
            return self.ebc(k)

    qm = quantize_inference_model(M(), quantization_dtype=fmt, output_dtype=torch.float32)
    qebc = qm.ebc
    gens = {"tw": {"a": sp.table_wise(rank=0), "b": sp.table_wise(rank=W - 1), "c": sp.table_wise(rank=0)},
            "cw": {"a": sp.column_wise(ranks=[0, W - 1]), "b": sp.column_wise(ranks=[W - 1, 0]), "c": sp.table_wise(rank=W - 1)},
            "rw": {"a": sp.row_wise(), "b": sp.row_wise(), "c": sp.row_wise()},
            "mixed": {"a": sp.row_wise(), "b": sp.column_wise(ranks=[0, W - 1]), "c": sp.table_wise(rank=W - 1)}}[kind]
    if W == 1 and kind in ("rw", "mixed"):
        gens = {k: (sp.table_wise(rank=0) if k != "b" or kind == "rw" else v) for k, v in gens.items()}
    plan = sp.construct_module_sharding_plan(qebc, gens, sharder=QuantEmbeddingBagCollectionSharder(), world_size=W, local_size=W, device_type="cuda")
    sharded = ShardedQuantEmbeddingBagCollection(qebc, plan, ShardingEnv.from_local(W, 0), device=torch.device("cuda:0"))
    g = torch.Generator().manual_seed(4)
    for B in (33, 7, 64):  # growing and shrinking batches reuse / regrow the buffers
        keys = ["fb2", "fa", "fc", "fb"]
        hs = {"fa": 300, "fb": 1000, "fb2": 1000, "fc": 77}
        lens = torch.randint(0, 5, (4 * B,), generator=g)
        vals = torch.cat([torch.randint(0, hs[k], (int(lens[i * B : (i + 1) * B].sum()),), generator=g) for i, k in enumerate(keys)])
        wts = torch.rand(vals.numel(), generator=g) if weighted else None
        kjt = KeyedJaggedTensor(keys=keys, values=vals, lengths=lens, weights=wts, stride=B)
        ref = qebc(kjt)  # unsharded quantized module on the CPU reference path
        got = sharded(kjt.to(torch.device("cuda:0")))
        assert got.keys() == ref.keys()
        tol = 2e-2 if fmt_name == "INT4" and kind != "tw" else 2e-3  # column / row shards are re-quantised with their own row scales
        torch.testing.assert_close(got.values().cpu().float(), ref.values().float(), rtol=tol, atol=tol * 5)
