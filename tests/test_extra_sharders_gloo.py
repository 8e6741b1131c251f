"""Sharded feature-processed / managed-collision / ITEP embedding collections on 2 CPU ranks (gloo).
Methodology of the reference's test_fp_embeddingbag.py / test_mc_embeddingbag.py: sharded vs unsharded golden."""
import pytest
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def _kjt(rank, B, keys, hashes, seed=0, max_len=4):
    from torchrec_b200.sparse import KeyedJaggedTensor

    g = torch.Generator().manual_seed(100 + rank + seed)
    lengths = torch.randint(0, max_len, (len(keys) * B,), generator=g)
    vals = [torch.randint(0, h, (int(lengths[i * B : (i + 1) * B].sum()),), generator=g) for i, h in enumerate(hashes)]
    return KeyedJaggedTensor(keys=keys, values=torch.cat(vals), lengths=lengths)


class _Wrap(torch.nn.Module):
    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, kjt):
        out = self.m(kjt)
        # DDP (static graph) must see tensors in the module output to arm its first-iteration reduction
        return out.wait() if hasattr(out, "wait") else out


def _run_fp(ctx, sharding: str):
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.modules.feature_processor_ import PositionWeightedModuleCollection
    from torchrec_b200.modules.fp_embedding_modules import FeatureProcessedEmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.comm_ops import set_gradient_division
    from torchrec_b200.parallel.fp_embeddingbag import FeatureProcessedEmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan

    set_gradient_division(False)
    torch.manual_seed(0)
    W, B, dev = ctx.world_size, 5, ctx.device
    tables = [EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=40, feature_names=["f0"]),
              EmbeddingBagConfig(name="t1", embedding_dim=8, num_embeddings=30, feature_names=["f1", "f2"])]

    def make():
        torch.manual_seed(1)
        fp = PositionWeightedModuleCollection({"f0": 6, "f1": 6, "f2": 6}, device=dev)
        with torch.no_grad():
            for p in fp.parameters():
                p.copy_(torch.linspace(0.5, 1.5, p.numel()))
        return FeatureProcessedEmbeddingBagCollection(EmbeddingBagCollection(tables, is_weighted=True, device=dev), fp)

    gold, local = make(), make()
    local.load_state_dict(gold.state_dict())
    apply_optimizer_in_backward(torch.optim.SGD, local._embedding_bag_collection.parameters(), {"lr": 0.1})
    gens = {"tw": {"t0": sp.table_wise(rank=0), "t1": sp.table_wise(rank=1)}, "rw": {"t0": sp.row_wise(), "t1": sp.row_wise()},
            "cw": {"t0": sp.column_wise(ranks=[0, 1]), "t1": sp.table_wise(rank=0)}}[sharding]
    sharder = FeatureProcessedEmbeddingBagCollectionSharder()
    plan = sp.construct_module_sharding_plan(local, gens, sharder=sharder, world_size=W, local_size=W, device_type=dev.type)
    model = DistributedModelParallel(_Wrap(local), device=dev, plan=ShardingPlan({"m": plan}), sharders=[sharder])
    dense = [p for _, p in model.named_parameters() if p.requires_grad]
    assert len(dense) == 3, [n for n, _ in model.named_parameters()]
    batches = [_kjt(r, B, ["f0", "f1", "f2"], [40, 30, 30]).to(dev) for r in range(W)]
    out = model(batches[ctx.rank])
    out = out.values()
    gouts = [gold(b).values() for b in batches]
    torch.testing.assert_close(out.float(), gouts[ctx.rank], rtol=1e-5, atol=1e-5)
    proj = torch.linspace(0.5, 1.5, out.shape[1], device=dev)
    (out * proj).sum().backward()
    sum((o * proj).sum() for o in gouts).backward()
    # DDP averages the position-weight gradient over ranks; golden sums
    gp = dict(gold._feature_processors.named_parameters())
    for n, p in model.named_parameters():
        if p.requires_grad:
            key = n.split("_feature_processors.")[-1]
            torch.testing.assert_close(p.grad * W, gp[key].grad, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("sharding", ["tw", "rw", "cw"])
def test_sharded_fp_ebc(sharding):
    run_multi_process(_run_fp, world_size=2, backend="gloo", sharding=sharding)


def _run_mc(ctx, sharding: str, kind: str):
    import torch.distributed as dist

    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig, EmbeddingConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection, EmbeddingCollection
    from torchrec_b200.modules.mc_embedding_modules import ManagedCollisionEmbeddingBagCollection, ManagedCollisionEmbeddingCollection
    from torchrec_b200.modules.mc_modules import DistanceLFU_EvictionPolicy, ManagedCollisionCollection, MCHManagedCollisionModule
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.mc_embedding import ManagedCollisionEmbeddingCollectionSharder
    from torchrec_b200.parallel.mc_embeddingbag import ManagedCollisionEmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan

    torch.manual_seed(0)
    W, B, dev = ctx.world_size, 4, ctx.device
    zch = 256
    if kind == "ebc":
        tables = [EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=zch, feature_names=["f0"]),
                  EmbeddingBagConfig(name="t1", embedding_dim=8, num_embeddings=zch, feature_names=["f1"])]
        emb = EmbeddingBagCollection(tables, device=dev)
    else:
        tables = [EmbeddingConfig(name="t0", embedding_dim=8, num_embeddings=zch, feature_names=["f0"]),
                  EmbeddingConfig(name="t1", embedding_dim=8, num_embeddings=zch, feature_names=["f1"])]
        emb = EmbeddingCollection(tables, device=dev)
    mcs = {t.name: MCHManagedCollisionModule(zch_size=zch, device=dev, eviction_policy=DistanceLFU_EvictionPolicy(), eviction_interval=2, input_hash_size=10**9)
           for t in tables}
    mcc = ManagedCollisionCollection(mcs, tables)
    cls = ManagedCollisionEmbeddingBagCollection if kind == "ebc" else ManagedCollisionEmbeddingCollection
    local = cls(emb, mcc, return_remapped_features=True)
    apply_optimizer_in_backward(torch.optim.SGD, emb.parameters(), {"lr": 0.1})
    gens = {"tw": {"t0": sp.table_wise(rank=0), "t1": sp.table_wise(rank=1)}, "rw": {"t0": sp.row_wise(), "t1": sp.row_wise()}}[sharding]
    sharder = ManagedCollisionEmbeddingBagCollectionSharder() if kind == "ebc" else ManagedCollisionEmbeddingCollectionSharder()
    plan = sp.construct_module_sharding_plan(local, gens, sharder=sharder, world_size=W, local_size=W, device_type=dev.type)
    model = DistributedModelParallel(_Wrap(local), device=dev, plan=ShardingPlan({"m": plan}), sharders=[sharder])
    seen = {}
    for step in range(5):
        kjt = _kjt(ctx.rank, B, ["f0", "f1"], [10**6, 10**6], seed=step % 2).to(dev)  # two alternating id sets -> stable mapping once learnt
        out, remapped = model(kjt)
        assert remapped.keys() == ["f0", "f1"] and remapped.values().numel() == kjt.values().numel()
        assert int(remapped.values().max()) < zch and int(remapped.values().min()) >= 0
        if kind == "ebc":
            out.values().sum().backward()
        else:
            sum(v.values().sum() for v in out.values()).backward()
        # the same raw id must map to the same slot on every rank (the owner decides)
        gathered = [None] * W
        dist.all_gather_object(gathered, (kjt.values().tolist(), remapped.values().tolist(), kjt.length_per_key()))
        if step >= 3:
            for f in range(2):
                m = {}
                for raw, slot, lpk in gathered:
                    lo = sum(lpk[:f])
                    for r_, s_ in zip(raw[lo : lo + lpk[f]], slot[lo : lo + lpk[f]]):
                        assert m.setdefault(r_, s_) == s_, (f, r_, s_, m[r_])
                seen[(step, f)] = m
    # once learnt, the mapping of the recurring id set is stable
    for f in range(2):
        common = set(seen[(3, f)]) & set(seen.get((1, f), {})) if (1, f) in seen else set()
    a, b_ = seen[(4, 0)], seen[(4, 1)]
    assert len(set(a.values())) == len(a) and len(set(b_.values())) == len(b_), "distinct ids share a slot although the table has room"
    assert model.module.m._managed_collision_collection.open_slots()["t0"].numel() == 1


@pytest.mark.parametrize("kind,sharding", [("ebc", "tw"), ("ebc", "rw"), ("ec", "rw")])
def test_sharded_mc(kind, sharding):
    run_multi_process(_run_mc, world_size=2, backend="gloo", sharding=sharding, kind=kind)


def _run_itep(ctx):
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.modules.itep_modules import GenericITEPModule, ITEPEmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.itep_embeddingbag import ITEPEmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan

    torch.manual_seed(0)
    W, B, dev = ctx.world_size, 8, ctx.device
    tables = [EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=16, feature_names=["f0"])]
    ebc = EmbeddingBagCollection(tables, device=dev)
    apply_optimizer_in_backward(torch.optim.SGD, ebc.parameters(), {"lr": 0.1})
    local = ITEPEmbeddingBagCollection(ebc, GenericITEPModule({"t0": 1000}, table_name_to_pruned_hash_sizes={"t0": 16}, pruning_interval=2, device=dev))
    sharder = ITEPEmbeddingBagCollectionSharder()
    plan = sp.construct_module_sharding_plan(local, {"t0": sp.row_wise()}, sharder=sharder, world_size=W, local_size=W, device_type=dev.type)
    model = DistributedModelParallel(_Wrap(local), device=dev, plan=ShardingPlan({"m": plan}), sharders=[sharder])
    for step in range(5):
        kjt = _kjt(0, 4, ["f0"], [1000], seed=0).to(dev)  # same hot set (<= 12 ids) on both ranks, 15 private rows available
        out = model(kjt)
        out.values().sum().backward()
    it = model.module.m._itep_module
    addr = it._addr("t0")
    # hot logical rows got private physical rows, and all ranks agree on the address table
    import torch.distributed as dist

    gathered = [torch.empty_like(addr) for _ in range(W)]
    dist.all_gather(gathered, addr)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    hot = kjt.values().unique()
    assert int((addr[hot] != 15).sum()) == hot.numel(), (addr[hot], hot)


def test_sharded_itep():
    run_multi_process(_run_itep, world_size=2, backend="gloo")


def _towers(ctx):
    import torch.distributed as dist

    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.modules.embedding_tower import EmbeddingTower, EmbeddingTowerCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embedding_tower_sharding import EmbeddingTowerCollectionSharder, EmbeddingTowerSharder, ShardedEmbeddingTower
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.sharding_plan import get_default_sharders
    from torchrec_b200.parallel.types import ShardingEnv, ShardingPlan
    from torchrec_b200.sparse import KeyedJaggedTensor

    W, dev = ctx.world_size, ctx.device

    class Inter(torch.nn.Module):
        def __init__(self, in_dim, out_dim):
            super().__init__()
            self.lin = torch.nn.Linear(in_dim, out_dim)

        def forward(self, kt):
            return torch.relu(self.lin(kt.values()))

    def build():
        torch.manual_seed(0)
        ebc_a = EmbeddingBagCollection([EmbeddingBagConfig(name="a0", embedding_dim=8, num_embeddings=30, feature_names=["fa0"]),
                                        EmbeddingBagConfig(name="a1", embedding_dim=8, num_embeddings=40, feature_names=["fa1"])])
        ebc_b = EmbeddingBagCollection([EmbeddingBagConfig(name="b0", embedding_dim=16, num_embeddings=50, feature_names=["fb0"])])
        return EmbeddingTowerCollection([EmbeddingTower(ebc_a, Inter(16, 6)), EmbeddingTower(ebc_b, Inter(16, 4))])

    class Model(torch.nn.Module):
        def __init__(self, towers, single):
            super().__init__()
            self.towers, self.single = towers, single

        def forward(self, kjt):
            return torch.cat([self.towers(kjt), self.single(kjt)], dim=1)

    def single_tower():
        torch.manual_seed(1)
        return EmbeddingTower(EmbeddingBagCollection([EmbeddingBagConfig(name="s0", embedding_dim=8, num_embeddings=20, feature_names=["fs0"])]), Inter(8, 3))

    golden = Model(build(), single_tower())
    model = Model(build(), single_tower())
    for m in (model.towers.towers[0].embedding, model.towers.towers[1].embedding, model.single.embedding):
        apply_optimizer_in_backward(torch.optim.SGD, m.parameters(), {"lr": 0.1})
    tower_plan = sp.construct_module_sharding_plan(model.towers, {"a0": sp.table_wise(rank=0), "a1": sp.row_wise(), "b0": sp.column_wise(ranks=list(range(W)))},
                                                   sharder=EmbeddingTowerCollectionSharder(), world_size=W, local_size=W, device_type=dev.type)
    single_plan = sp.construct_module_sharding_plan(model.single, {"s0": sp.table_wise(rank=W - 1)}, sharder=EmbeddingTowerSharder(), world_size=W, local_size=W,
                                                    device_type=dev.type)
    dmp = DistributedModelParallel(model, env=ShardingEnv.from_process_group(dist.group.WORLD), device=dev,
                                   plan=ShardingPlan({"towers": tower_plan, "single": single_plan}), sharders=get_default_sharders())
    assert isinstance(dmp.module.single, ShardedEmbeddingTower)
    dmp.load_state_dict(golden.state_dict())                       # unsharded keys: towers.towers.0.embedding.embedding_bags.a0.weight, ...interaction.lin.weight
    assert set(dmp.state_dict().keys()) == set(golden.state_dict().keys())
    g = torch.Generator().manual_seed(10 + ctx.rank)
    keys = ["fa0", "fa1", "fb0", "fs0"]
    lengths = torch.randint(0, 3, (len(keys) * 5,), generator=g)
    kjt = KeyedJaggedTensor(keys=keys, values=torch.randint(0, 20, (int(lengths.sum()),), generator=g), lengths=lengths)
    out, ref = dmp(kjt), golden(kjt)
    assert out.shape == (5, 6 + 4 + 3)
    torch.testing.assert_close(out, ref)
    out.sum().backward()
    gw = dmp.module.module.towers.towers[0].interaction.lin.weight.grad if hasattr(dmp.module, "module") else None
    lin = [p for n, p in dmp.named_parameters() if n.endswith("towers.0.interaction.lin.weight")][0]
    mine = lin.grad.clone()
    allg = [torch.empty_like(mine) for _ in range(W)]
    dist.all_gather(allg, mine)
    for a in allg[1:]:
        torch.testing.assert_close(a, allg[0])                      # interaction gradients are DDP-averaged


def test_embedding_tower_sharders():
    run_multi_process(_towers, world_size=2, backend="gloo")


def _pec(ctx):
    """Sharded PEC == plain sharded EC (weights after training), overlap masks are right, early non-overlapped lookup is exact."""
    import torch.distributed as dist

    from torchrec_b200.modules.embedding_configs import EmbeddingConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingCollection
    from torchrec_b200.modules.pec_embedding_modules import PECEmbeddingCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embedding import EmbeddingCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.pec_embedding import PECEmbeddingCollectionSharder, ShardedPECEmbeddingCollection
    from torchrec_b200.parallel.types import ShardingEnv, ShardingPlan
    from torchrec_b200.sparse import KeyedJaggedTensor

    W, dev = ctx.world_size, ctx.device

    def tables():
        return [EmbeddingConfig(name="t0", embedding_dim=8, num_embeddings=40, feature_names=["f0", "f0b"]),
                EmbeddingConfig(name="t1", embedding_dim=8, num_embeddings=30, feature_names=["f1"])]

    def build(pec: bool):
        torch.manual_seed(0)
        ec = EmbeddingCollection(tables())
        apply_optimizer_in_backward(torch.optim.SGD, ec.parameters(), {"lr": 0.5})
        mod = PECEmbeddingCollection(ec) if pec else ec
        sharder = PECEmbeddingCollectionSharder() if pec else EmbeddingCollectionSharder()
        plan = sp.construct_module_sharding_plan(mod, {"t0": sp.row_wise(), "t1": sp.table_wise(rank=W - 1)}, sharder=sharder, world_size=W, local_size=W, device_type=dev.type)
        return DistributedModelParallel(mod, env=ShardingEnv.from_process_group(dist.group.WORLD), device=dev, plan=ShardingPlan({"": plan}), sharders=[sharder])

    a, b = build(True), build(False)
    assert isinstance(a.module, ShardedPECEmbeddingCollection)
    g = torch.Generator().manual_seed(5 + ctx.rank)
    keys = ["f0", "f0b", "f1"]
    batches = []
    for _ in range(4):
        lengths = torch.randint(0, 4, (len(keys) * 6,), generator=g)
        batches.append(KeyedJaggedTensor(keys=keys, values=torch.randint(0, 12, (int(lengths.sum()),), generator=g), lengths=lengths))   # small id range: plenty of overlap
    for kjt in batches:
        oa, ob = a(kjt), b(kjt)
        for k in keys:
            torch.testing.assert_close(oa[k].values(), ob[k].values())
        sum(v.values().sum() for v in oa.values()).backward()
        sum(v.values().sum() for v in ob.values()).backward()
    for ta, tb in zip(a.module._embedding_collection._engine._tbes, b.module._engine._tbes):
        if hasattr(ta, "weights"):
            torch.testing.assert_close(ta.weights.detach(), tb.weights.detach())
    st = a.module.stats
    tot = torch.tensor([st["values"], st["overlapped"]])
    dist.all_reduce(tot)
    assert 0 < int(tot[1]) < int(tot[0])                                      # some ids repeated between consecutive batches, not all

    # explicit early stage: partition + gather the non-overlapped rows of the next batch BEFORE this batch's backward
    m = a.module
    out_cur = a(batches[0])
    ctx_next = m.create_context()
    dist_next = m.input_dist(ctx_next, batches[1]).wait().wait()
    m.prefetch_nonoverlapped(ctx_next, dist_next)
    sum(v.values().sum() for v in out_cur.values()).backward()               # updates rows of batches[0]
    early = m.compute_and_output_dist(ctx_next, dist_next)
    early = {k: v.values().detach().clone() for k, v in early.items()}
    b(batches[0])["f0"].values().sum().backward() if False else None
    ob = b(batches[0])
    sum(v.values().sum() for v in ob.values()).backward()
    ref = b(batches[1])
    for k in keys:
        torch.testing.assert_close(early[k], ref[k].values())


def test_pec_embedding_collection():
    run_multi_process(_pec, world_size=2, backend="gloo")


def _run_tower_models(ctx):
    """The tower test models sharded by the planner (towers are placed as units) give the unsharded model's predictions."""
    import copy

    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.test_utils.model_config import create_model_config
    from torchrec_b200.parallel.test_utils.model_input import ModelInput

    tables = [EmbeddingBagConfig(name=f"table_{i}", embedding_dim=8, num_embeddings=20 + i, feature_names=[f"feature_{i}"]) for i in range(4)]
    weighted = [EmbeddingBagConfig(name="weighted_table_0", embedding_dim=8, num_embeddings=15, feature_names=["weighted_feature_0"])]
    for name in ("test_tower_sparse_nn", "test_tower_collection_sparse_nn"):
        torch.manual_seed(0)
        model = create_model_config(name, num_float_features=6).generate_model(tables, weighted, torch.device("cpu"))
        gold = copy.deepcopy(model).to_empty(device="cpu") if any(p.is_meta for p in model.parameters()) else copy.deepcopy(model)
        torch.manual_seed(1)
        for p in gold.parameters():
            torch.nn.init.uniform_(p, -0.1, 0.1)
        dmp = DistributedModelParallel(model, device=torch.device("cpu"))
        # same weights on both sides: unsharded state dict -> sharded module
        sd = dmp.state_dict()
        gsd = gold.state_dict()
        for k, v in sd.items():
            if hasattr(v, "local_shards"):
                for sh in v.local_shards():
                    o, s = sh.metadata.shard_offsets, sh.metadata.shard_sizes
                    sh.tensor.copy_(gsd[k][o[0] : o[0] + s[0], o[1] : o[1] + s[1]])
            else:
                v.copy_(gsd[k])
        dmp.eval()
        gold.eval()
        g = torch.Generator().manual_seed(5)
        _, locals_ = ModelInput.generate_global_and_local_batches(ctx.world_size, batch_size=4, tables=tables, weighted_tables=weighted, num_float_features=6, generator=g)
        with torch.no_grad():
            torch.testing.assert_close(dmp(locals_[ctx.rank]), gold(locals_[ctx.rank]), rtol=1e-5, atol=1e-6, msg=lambda m: f"{name}: {m}")
        kinds = {type(m).__name__ for m in dmp.modules()}
        assert ("ShardedEmbeddingTower" in kinds) or ("ShardedEmbeddingTowerCollection" in kinds), (name, kinds)


def test_tower_test_models_sharded():
    run_multi_process(_run_tower_models, world_size=2, backend="gloo")
