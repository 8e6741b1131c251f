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
