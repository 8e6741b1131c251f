"""Unsharded model families and cross networks vs their defining formulas (reference models/tests/test_dlrm.py, test_deepfm.py,
modules/tests/test_crossnet.py)."""
import pytest
import torch

from torchrec_b200.datasets.utils import Batch
from torchrec_b200.models.deepfm import SimpleDeepFMNN
from torchrec_b200.models.dlrm import DLRM, DLRM_DCN, DLRM_Projection, DLRMTrain, InteractionArch
from torchrec_b200.modules.crossnet import CrossNet, LowRankCrossNet, LowRankMixtureCrossNet, VectorCrossNet
from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
from torchrec_b200.sparse import KeyedJaggedTensor


def _ebc(F=3, D=8, rows=50):
    return EmbeddingBagCollection([EmbeddingBagConfig(name=f"t{i}", embedding_dim=D, num_embeddings=rows, feature_names=[f"f{i}"]) for i in range(F)])


def _batch(B=6, F=3, rows=50, seed=0):
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(0, 4, (F * B,), generator=g)
    kjt = KeyedJaggedTensor(keys=[f"f{i}" for i in range(F)], values=torch.randint(0, rows, (int(lengths.sum()),), generator=g), lengths=lengths)
    return Batch(dense_features=torch.randn(B, 5, generator=g), sparse_features=kjt, labels=torch.randint(0, 2, (B,), generator=g).float())


def test_interaction_arch_is_pairwise_dots():
    B, F, D = 4, 3, 8
    dense, sparse = torch.randn(B, D), torch.randn(B, F, D)
    out = InteractionArch(F)(dense, sparse)
    x = torch.cat([dense.unsqueeze(1), sparse], 1)
    gram = torch.bmm(x, x.transpose(1, 2))
    pairs = torch.stack([gram[:, i, j] for i in range(F + 1) for j in range(i)], 1)
    assert out.shape == (B, D + (F + 1) * F // 2)
    torch.testing.assert_close(out[:, :D], dense)
    torch.testing.assert_close(out[:, D:].sort(1).values, pairs.sort(1).values)


def test_dlrm_family_forward_backward():
    b = _batch()
    models = [
        DLRM(_ebc(), dense_in_features=5, dense_arch_layer_sizes=[16, 8], over_arch_layer_sizes=[16, 1]),
        DLRM_Projection(_ebc(), dense_in_features=5, dense_arch_layer_sizes=[16, 8], over_arch_layer_sizes=[16, 1], interaction_branch1_layer_sizes=[16, 16],
                        interaction_branch2_layer_sizes=[16, 8]),
        DLRM_DCN(_ebc(), dense_in_features=5, dense_arch_layer_sizes=[16, 8], over_arch_layer_sizes=[16, 1], dcn_num_layers=2, dcn_low_rank_dim=4),
    ]
    for m in models:
        logits = m(b.dense_features, b.sparse_features)
        assert logits.shape == (6, 1)
        loss, (dl, dlogits, dlabels) = DLRMTrain(m)(b)
        assert loss.dim() == 0 and dlogits.shape == (6,) and torch.equal(dlabels, b.labels)
        loss.backward()
        grads = [p.grad for p in m.parameters()]
        assert all(g is not None for g in grads) and any(float(g.abs().sum()) > 0 for g in grads)
    ref = torch.nn.functional.binary_cross_entropy_with_logits(models[0](b.dense_features, b.sparse_features).squeeze(-1), b.labels)
    torch.testing.assert_close(DLRMTrain(models[0])(b)[0], ref)


def test_deepfm_forward_backward():
    b = _batch()
    m = SimpleDeepFMNN(num_dense_features=5, embedding_bag_collection=_ebc(), hidden_layer_size=16, deep_fm_dimension=4)
    out = m(b.dense_features, b.sparse_features)
    assert out.shape == (6, 1) and bool(((out >= 0) & (out <= 1)).all())
    out.sum().backward()
    assert all(p.grad is not None for p in m.parameters())


def test_crossnets_match_their_formulas():
    torch.manual_seed(0)
    B, N = 5, 12
    x0 = torch.randn(B, N)
    full = CrossNet(N, 2)
    x = x0
    for W, b in zip(full.kernels, full.bias):
        x = x0 * (x @ W.t() + b.squeeze(1)) + x
    torch.testing.assert_close(full(x0), x)

    low = LowRankCrossNet(N, 3, low_rank=4)
    with torch.no_grad():
        for b in low.bias:
            b.normal_()
    x = x0
    for W, V, b in zip(low.W_kernels, low.V_kernels, low.bias):
        x = x0 * ((x @ V.t()) @ W.t() + b) + x
    torch.testing.assert_close(low(x0), x)

    vec = VectorCrossNet(N, 2)
    x = x0
    for w, b in zip(vec.kernels, vec.bias):
        x = x0 * (x @ w) + b.squeeze(1) + x
    torch.testing.assert_close(vec(x0), x)

    mix = LowRankMixtureCrossNet(N, 2, num_experts=3, low_rank=4)
    out = mix(x0)
    assert out.shape == (B, N)
    out.sum().backward()
    assert all(p.grad is not None for p in mix.parameters())
    one = LowRankMixtureCrossNet(N, 1, num_experts=1, low_rank=4, activation=torch.nn.Identity())
    U, V, C, b = one.U_kernels[0][0], one.V_kernels[0][0], one.C_kernels[0][0], one.bias[0].squeeze(1)
    torch.testing.assert_close(one(x0), x0 * (((x0 @ V.t()) @ C.t()) @ U.t() + b) + x0)


def test_transformer_dlrm_and_experimental_marker():
    import warnings

    from torchrec_b200.models.experimental.transformerdlrm import DLRM_Transformer, InteractionTransformerArch
    from torchrec_b200.utils.experimental import experimental

    b = _batch()
    m = DLRM_Transformer(_ebc(D=8), dense_in_features=5, dense_arch_layer_sizes=[16, 8], over_arch_layer_sizes=[16, 1], nhead=2, ntransformer_layers=1)
    m.eval()
    out = m(b.dense_features, b.sparse_features)
    assert out.shape == (6, 1)
    m.train()
    m(b.dense_features, b.sparse_features).sum().backward()
    assert all(p.grad is not None for p in m.parameters())
    assert InteractionTransformerArch(0, 8, nhead=2, ntransformer_layers=1)(torch.ones(2, 8), torch.ones(2, 0, 8)).shape == (2, 8)

    @experimental
    def f(x):
        return x + 1

    @experimental(feature="Thing", since="0.1")
    class Thing:
        def __init__(self, v):
            self.v = v

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert f(1) == 2 and f(2) == 3 and Thing(3).v == 3 and Thing(4).v == 4
    assert len([x for x in w if "experimental" in str(x.message)]) == 2      # once per object


def test_inference_mode_does_not_poison_cached_helpers_for_training():
    """Index / workspace tensors cached by helpers are created outside inference mode: a serving-style forward under
    ``torch.inference_mode()`` followed by a training step in the same process must work (caches used to hold inference tensors)."""
    import torch

    from torchrec_b200.models.dlrm import DLRM
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.ops import dense as D
    from torchrec_b200.sparse.jagged_tensor import KeyedJaggedTensor

    D._TRIU_CACHE.clear()
    ebc = EmbeddingBagCollection([EmbeddingBagConfig(name=f"t{i}", embedding_dim=8, num_embeddings=20, feature_names=[f"f{i}"]) for i in range(3)])
    model = DLRM(ebc, 4, [16, 8], [16, 1])
    kjt = KeyedJaggedTensor(keys=["f2", "f0", "f1"], values=torch.randint(0, 20, (9,)), lengths=torch.tensor([1, 2, 1, 1, 1, 1, 0, 1, 1]))
    dense = torch.randn(3, 4)
    with torch.inference_mode():
        model(dense, kjt.permute([1, 2, 0]))
    out = model(dense, kjt.permute([1, 2, 0]))  # same permutation, same interaction size: the cached helpers are reused
    out.sum().backward()
    assert all(p.grad is not None for p in model.over_arch.parameters())


def test_itep_prunes_resets_rows_and_reports_stats():
    """Unsharded ITEP: hot ids take over physical rows at the pruning interval, the re-assigned rows are re-initialised and the
    eviction statistics add up (reference modules/itep_modules.py:170-452)."""
    import torch

    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.modules.itep_embedding_modules import ITEPEmbeddingBagCollection
    from torchrec_b200.modules.itep_modules import GenericITEPModule
    from torchrec_b200.sparse.jagged_tensor import KeyedJaggedTensor

    torch.manual_seed(0)
    ebc = EmbeddingBagCollection([EmbeddingBagConfig(name="t0", embedding_dim=4, num_embeddings=8, feature_names=["f0"])])  # 8 physical rows
    itep = GenericITEPModule({"t0": 100}, table_name_to_pruned_hash_sizes={"t0": 8}, pruning_interval=3)
    m = ITEPEmbeddingBagCollection(ebc, itep)
    m.train()
    with torch.no_grad():
        ebc.embedding_bags["t0"].weight.fill_(7.0)  # sentinel: a reset row no longer holds it
    hot = torch.tensor([50, 60, 70, 50, 60, 70])     # logical rows far outside the initially mapped 0..6
    for _ in range(4):
        m(KeyedJaggedTensor(keys=["f0"], values=hot, lengths=torch.tensor([3, 3])))
    st = itep.eviction_stats()["t0"]
    assert st["prunes"] == 1 and st["last_evicted"] == 3 and st["evicted_total"] == 3 and st["physical_rows"] == 8 and st["logical_rows"] == 100
    addr = itep._addr("t0")
    phys = addr[torch.tensor([50, 60, 70])]
    assert phys.unique().numel() == 3 and int(phys.max()) < 7, "every hot logical row got its own physical row (the last slot stays the shared one)"
    w = ebc.embedding_bags["t0"].weight
    assert bool((w[phys].abs() < 1.0).all()) and bool((w[7] == 7.0).all()), "re-assigned rows were re-initialised, untouched rows kept their values"
    assert st["access_share_resident"] > 0.99  # all (decayed) accesses now hit rows that own a physical row
    # cold ids share the last physical row
    assert int(addr[99]) == 7


def test_legacy_position_weighted_module_over_jt_dict():
    from torchrec_b200.modules.feature_processor import PositionWeightedModule, offsets_to_range_traceble
    from torchrec_b200.sparse import JaggedTensor

    m = PositionWeightedModule({"a": 3})
    with torch.no_grad():
        m.position_weights["a"].copy_(torch.tensor([1.0, 2.0, 3.0]))
    feats = {"a": JaggedTensor(values=torch.arange(6), lengths=torch.tensor([2, 0, 4])), "b": JaggedTensor(values=torch.arange(2), lengths=torch.tensor([1, 1, 0]))}
    out = m(feats)
    assert out["a"].weights().tolist() == [1.0, 2.0, 1.0, 2.0, 3.0, 3.0] and out["b"].weights_or_none() is None and list(out) == ["a", "b"]
    out["a"].weights().sum().backward()
    assert m.position_weights["a"].grad.tolist() == [2.0, 2.0, 2.0]
    assert offsets_to_range_traceble(torch.tensor([0, 2, 2, 5]), torch.arange(5)).tolist() == [0, 1, 0, 1, 2]


def test_debug_embedding_collections_catch_bad_ids_and_gradients():
    from torchrec_b200.modules.debug_embedding_modules import DebugEmbeddingBagCollection, DebugEmbeddingCollection
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig, EmbeddingConfig
    from torchrec_b200.sparse import KeyedJaggedTensor

    good = KeyedJaggedTensor(keys=["f"], values=torch.tensor([1, 2, 3]), lengths=torch.tensor([2, 1]))
    bad = KeyedJaggedTensor(keys=["f"], values=torch.tensor([1, 99]), lengths=torch.tensor([1, 1]))
    ebc = DebugEmbeddingBagCollection([EmbeddingBagConfig(name="t", embedding_dim=4, num_embeddings=10, feature_names=["f"])], torch.device("cpu"), debug_mode=True)
    out = ebc(good)
    out.values().sum().backward()  # finite gradient: fine
    with pytest.raises(ValueError):
        ebc(bad)
    with pytest.raises(RuntimeError, match="NaN/Inf detected in gradient entering ebc.values"):
        (ebc(good).values() * float("nan")).sum().backward()
    quiet = DebugEmbeddingBagCollection([EmbeddingBagConfig(name="t", embedding_dim=4, num_embeddings=10, feature_names=["f"])], torch.device("cpu"))
    (quiet(good).values() * float("nan")).sum().backward()  # debug mode off: a plain collection
    ec = DebugEmbeddingCollection([EmbeddingConfig(name="t", embedding_dim=4, num_embeddings=10, feature_names=["f"])], torch.device("cpu"), debug_mode=True)
    with pytest.raises(RuntimeError, match=r"ec\[f\].values"):
        (ec(good)["f"].values() * float("inf")).sum().backward()
    assert DebugEmbeddingCollection(ec=ec.ec).ec is ec.ec and type(ebc.ebc).__name__ == "EmbeddingBagCollection"
