"""Dynamic re-sharding and the model delta tracker on 2 CPU ranks (gloo)
(reference tests: distributed/tests/test_dynamic_sharding.py, model_tracker/tests/test_model_delta_tracker.py)."""
import pytest
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def _tables():
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig

    return [EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=40, feature_names=["f0"]),
            EmbeddingBagConfig(name="t1", embedding_dim=16, num_embeddings=30, feature_names=["f1"])]


def _kjt(rank, B, seed=0):
    from torchrec_b200.sparse import KeyedJaggedTensor

    g = torch.Generator().manual_seed(7 + rank + 13 * seed)
    lengths = torch.randint(0, 4, (2 * B,), generator=g)
    v0 = torch.randint(0, 40, (int(lengths[:B].sum()),), generator=g)
    v1 = torch.randint(0, 30, (int(lengths[B:].sum()),), generator=g)
    return KeyedJaggedTensor(keys=["f0", "f1"], values=torch.cat([v0, v1]), lengths=lengths)


class _Wrap(torch.nn.Module):
    def __init__(self, ebc):
        super().__init__()
        self.ebc = ebc

    def forward(self, kjt):
        return self.ebc(kjt).values()


def _build(ctx, gens, tracker=None):
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.comm_ops import set_gradient_division
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan

    set_gradient_division(False)
    torch.manual_seed(0)
    dev, W = ctx.device, ctx.world_size
    gold = EmbeddingBagCollection(_tables(), device=dev)
    local = EmbeddingBagCollection(_tables(), device=dev)
    local.load_state_dict(gold.state_dict())
    apply_optimizer_in_backward(RowWiseAdagrad, local.parameters(), {"lr": 0.1, "eps": 1e-8})
    plan = sp.construct_module_sharding_plan(local, gens, sharder=EmbeddingBagCollectionSharder(), world_size=W, local_size=W, device_type=dev.type)
    model = DistributedModelParallel(_Wrap(local), device=dev, plan=ShardingPlan({"ebc": plan}), sharders=[EmbeddingBagCollectionSharder()], model_tracker_config=tracker)
    return gold, local, model, RowWiseAdagrad(gold.parameters(), lr=0.1, eps=1e-8)


def _step(ctx, model, gold, gold_opt, seed):
    W, dev = ctx.world_size, ctx.device
    batches = [_kjt(r, 6, seed).to(dev) for r in range(W)]
    out = model(batches[ctx.rank])
    gouts = [gold(b).values() for b in batches]
    torch.testing.assert_close(out.float(), gouts[ctx.rank], rtol=1e-5, atol=1e-5)
    out.sum().backward()
    gold_opt.zero_grad()
    sum(o.sum() for o in gouts).backward()
    gold_opt.step()
    return batches


def _run_reshard(ctx):
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder

    W, dev = ctx.world_size, ctx.device
    gold, local, model, gold_opt = _build(ctx, {"t0": sp.table_wise(rank=0), "t1": sp.table_wise(rank=1)})
    _step(ctx, model, gold, gold_opt, 0)
    # move t0 to rank 1 (table-wise) and make t1 row-wise, keeping weights AND adagrad state
    meta = EmbeddingBagCollection(_tables(), device=torch.device("meta"))
    new = sp.construct_module_sharding_plan(meta, {"t0": sp.table_wise(rank=1), "t1": sp.row_wise()}, sharder=EmbeddingBagCollectionSharder(), world_size=W,
                                            local_size=W, device_type=dev.type)
    model.reshard("ebc", dict(new))
    for s in (1, 2):
        _step(ctx, model, gold, gold_opt, s)
    # and back again
    back = sp.construct_module_sharding_plan(meta, {"t0": sp.row_wise(), "t1": sp.table_wise(rank=0)}, sharder=EmbeddingBagCollectionSharder(), world_size=W,
                                             local_size=W, device_type=dev.type)
    model.reshard("ebc", dict(back))
    _step(ctx, model, gold, gold_opt, 3)
    sd = model.state_dict()
    for name in ("t0", "t1"):
        st = sd[f"ebc.embedding_bags.{name}.weight"]
        ref = gold.embedding_bags[name].weight.detach()
        for sh in st.local_shards():
            o, s_ = sh.metadata.shard_offsets, sh.metadata.shard_sizes
            torch.testing.assert_close(sh.tensor, ref[o[0] : o[0] + s_[0], o[1] : o[1] + s_[1]], rtol=1e-4, atol=1e-5)
    assert "ebc.embedding_bags.t0.weight" in model.fused_optimizer.state_dict()["state"]


def test_reshard_preserves_training_state():
    run_multi_process(_run_reshard, world_size=2, backend="gloo")


def _run_tracker(ctx, mode: str):
    import torch.distributed as dist

    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.model_tracker import DeltaTrackerConfig, TrackingMode

    W = ctx.world_size
    cfg = DeltaTrackerConfig(tracking_mode=TrackingMode(mode), consumers=["a", "b"], delete_on_read=True)
    gold, local, model, gold_opt = _build(ctx, {"t0": sp.row_wise(), "t1": sp.table_wise(rank=1)}, tracker=cfg)
    tr = model.get_model_tracker()
    assert set(tr.fqn_to_feature_names()) == {"ebc.embedding_bags.t0", "ebc.embedding_bags.t1"}
    w_before = {n: gold.embedding_bags[n].weight.detach().clone() for n in ("t0", "t1")}
    seen = {"t0": set(), "t1": set()}
    for s in range(3):
        batches = _step(ctx, model, gold, gold_opt, s)
        tr.step()
        for b in batches:
            jt = b.to_dict()
            seen["t0"].update(jt["f0"].values().tolist())
            seen["t1"].update(jt["f1"].values().tolist())
        if s == 0:
            first = tr.get_unique(consumer="a")  # consumer a reads after step 0 ...
    delta_a = tr.get_unique(consumer="a")  # ... and gets only steps 1-2 now
    delta_b = tr.get_unique(consumer="b")  # consumer b gets everything
    # union over ranks of b's ids == all ids seen
    for table in ("t0", "t1"):
        fqn = f"ebc.embedding_bags.{table}"
        mine = delta_b[fqn].ids.tolist() if fqn in delta_b else []
        gathered = [None] * W
        dist.all_gather_object(gathered, mine)
        assert set(x for g in gathered for x in g) == seen[table], (table, gathered, seen[table])
        if fqn in delta_b and fqn in delta_a:
            assert set(delta_a[fqn].ids.tolist()) <= set(delta_b[fqn].ids.tolist())
    if mode == "embedding":
        # FIRST update mode: the stored state is the row BEFORE the first update of the window
        for fqn, rows in delta_b.items():
            table = fqn.rsplit(".", 1)[-1]
            torch.testing.assert_close(rows.states, w_before[table][rows.ids][:, : rows.states.shape[1]], rtol=1e-5, atol=1e-6)
    if mode == "rowwise_adagrad":
        for fqn, rows in delta_b.items():
            assert rows.states is not None and rows.states.shape[0] == rows.ids.numel() and bool((rows.states >= 0).all())
    assert tr.get_unique(consumer="b") == {} or all(v.ids.numel() == 0 for v in tr.get_unique(consumer="b").values())
    # manual recording (modules without an engine hook): ids + states supplied by the caller land in the same store
    from torchrec_b200.sparse.jagged_tensor import KeyedJaggedTensor

    kjt = KeyedJaggedTensor(keys=["f0"], values=torch.tensor([3, 3, 5]), lengths=torch.tensor([2, 1]))
    if mode == "embedding":
        tr.record_embeddings(None, kjt, torch.arange(12, dtype=torch.float32).view(3, 4))
        got = tr.get_unique(consumer="b")["ebc.embedding_bags.t0"]
        assert got.ids.tolist() == [3, 5] and torch.equal(got.states[:, :4], torch.tensor([[0.0, 1, 2, 3], [8, 9, 10, 11]]))  # FIRST occurrence wins
    elif mode == "id_only":
        tr.record_ids(kjt)
        assert tr.get_unique_ids(consumer="b")["ebc.embedding_bags.t0"].tolist() == [3, 5]


@pytest.mark.parametrize("mode", ["id_only", "embedding", "rowwise_adagrad"])
def test_model_delta_tracker(mode):
    run_multi_process(_run_tracker, world_size=2, backend="gloo", mode=mode)
