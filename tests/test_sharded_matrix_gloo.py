"""Randomised sharding matrix on 4 CPU ranks (gloo): random tables (dims, rows, SUM / MEAN pooling, shared tables with two features,
weighted or not), a random placement per table out of table-wise / row-wise / column-wise / table-row-wise / table-column-wise / grid /
data-parallel, random jagged batches with empty bags - the sharded collection must match the unsharded one in the forward output and
in the weights after two fused-SGD steps. One process group, several seeds per launch (reference: distributed/test_utils/
test_model_parallel*.py matrices)."""
import random

import pytest
import torch

from torchrec_b200.utils.multiprocess import run_multi_process

SEEDS_PER_LAUNCH = 5


def _case(seed: int, weighted: bool, W: int, local: int, no_col_split: bool = False, allow_dp: bool = True):
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig, PoolingType
    from torchrec_b200.parallel import sharding_plan as sp

    rng = random.Random(1000 * seed + (7 if weighted else 0))
    n_tables = rng.randint(3, 5)
    tables, gens, fid = [], {}, 0
    for t in range(n_tables):
        dim = rng.choice([4, 8, 12, 16])
        rows = rng.randint(9, 40) if rng.random() < 0.8 else rng.randint(1, 6)  # tiny tables: row shards of some ranks are empty
        feats = [f"f{fid + i}" for i in range(rng.choice([1, 1, 2]))]
        fid += len(feats)
        pooling = PoolingType.SUM if (weighted or rng.random() < 0.6) else PoolingType.MEAN
        name = f"t{t}"
        tables.append(EmbeddingBagConfig(name=name, embedding_dim=dim, num_embeddings=rows, feature_names=feats, pooling=pooling))
        hosts = W // local
        kind = rng.choice(["tw", "rw", "twrw", "dp"] if no_col_split else ["tw", "rw", "cw", "twrw", "twcw", "grid", "dp"])
        if kind == "dp" and not allow_dp:
            kind = "rw"
        if kind == "cw" and dim % 8 != 0:
            kind = "tw"
        if kind in ("twcw", "grid") and dim % 8 != 0:
            kind = "twrw"
        if kind == "tw":
            gens[name] = sp.table_wise(rank=rng.randrange(W))
        elif kind == "rw":
            gens[name] = sp.row_wise()
        elif kind == "cw":
            gens[name] = sp.column_wise(ranks=rng.sample(range(W), 2))
        elif kind == "twrw":
            gens[name] = sp.table_row_wise(host_index=rng.randrange(hosts))
        elif kind == "twcw":
            gens[name] = sp.table_column_wise(ranks=rng.sample(range(W), 2)) if hasattr(sp, "table_column_wise") else sp.column_wise(ranks=rng.sample(range(W), 2))
        elif kind == "grid":
            gens[name] = sp.grid_shard(host_indexes=list(range(hosts)))
        else:
            gens[name] = sp.data_parallel()
    return tables, gens


def _batch(tables, seed: int, rank: int, B: int, weighted: bool):
    from torchrec_b200.sparse import KeyedJaggedTensor

    g = torch.Generator().manual_seed(7919 * seed + 31 * rank + 1)
    keys, hashes = [], []
    for t in tables:
        for f in t.feature_names:
            keys.append(f)
            hashes.append(t.num_embeddings)
    lengths = torch.randint(0, 4, (len(keys) * B,), generator=g)
    lengths[torch.rand(lengths.shape, generator=g) < 0.2] = 0  # plenty of empty bags
    vals = [torch.randint(0, h, (int(lengths[i * B : (i + 1) * B].sum()),), generator=g) for i, h in enumerate(hashes)]
    values = torch.cat(vals) if vals else torch.zeros(0, dtype=torch.long)
    w = torch.rand(values.numel(), generator=g) + 0.5 if weighted else None
    return KeyedJaggedTensor(keys=keys, values=values, lengths=lengths, weights=w)


def _run(ctx, weighted: bool, first_seed: int, n_seeds: int = SEEDS_PER_LAUNCH, opt: str = "sgd", planner: bool = False, uneven: bool = False):
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.comm_ops import set_gradient_division
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan

    set_gradient_division(False)
    W, local, B, dev = ctx.world_size, 2, 5, ctx.device
    for seed in range(first_seed, first_seed + n_seeds):
        # row-wise Adagrad normalises per (row, column shard): compared on plans without column splits, as in the reference's matrix
        tables, gens = _case(seed, weighted, W, local, no_col_split=(opt == "rowwise_adagrad"))
        torch.manual_seed(seed)
        gold = EmbeddingBagCollection(tables, is_weighted=weighted, device=dev)
        sharded_src = EmbeddingBagCollection(tables, is_weighted=weighted, device=dev)
        sharded_src.load_state_dict(gold.state_dict())
        if opt == "rowwise_adagrad":
            from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad

            apply_optimizer_in_backward(RowWiseAdagrad, sharded_src.parameters(), {"lr": 0.1, "eps": 1e-8})
        else:
            apply_optimizer_in_backward(torch.optim.SGD, sharded_src.parameters(), {"lr": 0.1})
        sharder = EmbeddingBagCollectionSharder()
        if planner:
            # the plan comes from EmbeddingShardingPlanner (collective: rank 0 plans, everybody receives it) under random per-table constraints
            import torch.distributed as dist

            from torchrec_b200.parallel.planner import EmbeddingShardingPlanner, Topology
            from torchrec_b200.parallel.planner.types import ParameterConstraints

            rng = random.Random(seed)
            allowed = ["table_wise", "row_wise", "column_wise", "table_row_wise", "data_parallel"] if opt == "sgd" else ["table_wise", "row_wise", "table_row_wise"]
            cons = {t.name: ParameterConstraints(sharding_types=rng.sample(allowed, rng.randint(1, len(allowed))), min_partition=4) for t in tables}

            class Holder(torch.nn.Module):
                def __init__(self, ebc):
                    super().__init__()
                    self.ebc = ebc

            pl = EmbeddingShardingPlanner(topology=Topology(world_size=W, local_world_size=local, compute_device=dev.type), batch_size=B, constraints=cons)
            plan = pl.collective_plan(Holder(sharded_src), [sharder], dist.group.WORLD).plan["ebc"]
        else:
            plan = sp.construct_module_sharding_plan(sharded_src, gens, sharder=sharder, world_size=W, local_size=local, device_type=dev.type)
        desc = {n: plan[n].sharding_type for n in plan}

        class Wrap(torch.nn.Module):
            def __init__(self, ebc):
                super().__init__()
                self.ebc = ebc

            def forward(self, kjt):
                return self.ebc(kjt).values()

        model = DistributedModelParallel(Wrap(sharded_src), device=dev, plan=ShardingPlan({"ebc": plan}), sharders=[sharder])
        dense_params = [p for _, p in model.named_parameters() if p.requires_grad]  # data-parallel tables
        dense_opt = torch.optim.SGD(dense_params, lr=0.1) if dense_params else None
        if opt == "rowwise_adagrad":
            from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad

            dp_names = {n for n in plan if plan[n].sharding_type == "data_parallel"}  # replicated tables train with the dense (SGD) optimizer
            mp_params = [p for n, p in gold.named_parameters() if n.split(".")[1] not in dp_names]
            gold_opt = RowWiseAdagrad(mp_params, lr=0.1, eps=1e-8) if mp_params else torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.1)
            gold_dp_opt = torch.optim.SGD([p for n, p in gold.named_parameters() if n.split(".")[1] in dp_names], lr=0.1) if dp_names else None
        else:
            gold_opt, gold_dp_opt = torch.optim.SGD(gold.parameters(), lr=0.1), None
        for step in range(2):
            # uneven: every rank feeds its own batch size (the last partial batch of an epoch)
            batches = [_batch(tables, 10 * seed + step, r, B + (r + step) % 3 if uneven else B, weighted).to(dev) for r in range(W)]
            out = model(batches[ctx.rank])
            gouts = [gold(b).values() for b in batches]
            torch.testing.assert_close(out.float(), gouts[ctx.rank], rtol=1e-5, atol=1e-5, msg=lambda m: f"seed {seed} step {step} plan {desc}: {m}")
            proj = torch.linspace(0.5, 1.5, out.shape[1], device=dev)
            (out * proj).sum().backward()
            if dense_opt is not None:
                for p in dense_params:  # DDP averages over ranks, the golden sums the per-rank losses
                    p.grad.mul_(W)
                dense_opt.step()
                dense_opt.zero_grad()
            gold_opt.zero_grad()
            if gold_dp_opt is not None:
                gold_dp_opt.zero_grad()
            sum((o * proj).sum() for o in gouts).backward()
            gold_opt.step()
            if gold_dp_opt is not None:
                gold_dp_opt.step()
        sd = model.state_dict()
        for t in tables:
            st = sd[f"ebc.embedding_bags.{t.name}.weight"]
            ref = gold.embedding_bags[t.name].weight.detach()
            if hasattr(st, "local_shards"):
                for sh in st.local_shards():
                    o, s = sh.metadata.shard_offsets, sh.metadata.shard_sizes
                    torch.testing.assert_close(sh.tensor, ref[o[0] : o[0] + s[0], o[1] : o[1] + s[1]], rtol=1e-4, atol=1e-5,
                                               msg=lambda m: f"seed {seed} table {t.name} ({desc[t.name]}): {m}")
            else:
                torch.testing.assert_close(st, ref, rtol=1e-4, atol=1e-5, msg=lambda m: f"seed {seed} table {t.name} ({desc[t.name]}): {m}")


@pytest.mark.parametrize("weighted", [False, True])
def test_random_sharding_matrix_4_ranks(weighted):
    run_multi_process(_run, world_size=4, backend="gloo", weighted=weighted, first_seed=0)


def test_random_sharding_matrix_uneven_batch_per_rank():
    run_multi_process(_run, world_size=4, backend="gloo", weighted=False, first_seed=500, n_seeds=4, uneven=True)


def test_random_sharding_matrix_rowwise_adagrad():
    run_multi_process(_run, world_size=4, backend="gloo", weighted=False, first_seed=100, n_seeds=4, opt="rowwise_adagrad")


def test_random_sharding_matrix_planner_plans():
    run_multi_process(_run, world_size=4, backend="gloo", weighted=False, first_seed=200, n_seeds=4, planner=True)


def _run_ec(ctx, dedup: bool, first_seed: int, n_seeds: int = SEEDS_PER_LAUNCH, uneven: bool = False):
    """Sequence embeddings: random tables / TW / RW / CW placements / jagged batches, index de-duplication on or off."""
    from torchrec_b200.modules.embedding_configs import EmbeddingConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.comm_ops import set_gradient_division
    from torchrec_b200.parallel.embedding import EmbeddingCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan
    from torchrec_b200.sparse import KeyedJaggedTensor

    set_gradient_division(False)
    W, B = ctx.world_size, 4
    for seed in range(first_seed, first_seed + n_seeds):
        rng = random.Random(31 * seed + (3 if dedup else 0))
        dim = rng.choice([8, 16])  # one dim per collection (EmbeddingCollection contract)
        tables, gens, keys, hashes, fid = [], {}, [], [], 0
        for t in range(rng.randint(2, 4)):
            rows = rng.randint(9, 40)
            feats = [f"f{fid + i}" for i in range(rng.choice([1, 1, 2]))]
            fid += len(feats)
            tables.append(EmbeddingConfig(name=f"t{t}", embedding_dim=dim, num_embeddings=rows, feature_names=feats))
            keys += feats
            hashes += [rows] * len(feats)
            kind = rng.choice(["tw", "rw", "cw"])
            gens[f"t{t}"] = sp.table_wise(rank=rng.randrange(W)) if kind == "tw" else (sp.row_wise() if kind == "rw" else sp.column_wise(ranks=rng.sample(range(W), 2)))
        torch.manual_seed(seed)
        gold = EmbeddingCollection(tables)
        src = EmbeddingCollection(tables)
        src.load_state_dict(gold.state_dict())
        apply_optimizer_in_backward(torch.optim.SGD, src.parameters(), {"lr": 0.1})
        sharder = EmbeddingCollectionSharder(use_index_dedup=dedup)
        plan = sp.construct_module_sharding_plan(src, gens, sharder=sharder, world_size=W, local_size=W, device_type="cpu")
        desc = {n: plan[n].sharding_type for n in plan}

        class Wrap(torch.nn.Module):
            def __init__(self, ec):
                super().__init__()
                self.ec = ec

            def forward(self, kjt):
                return self.ec(kjt)

        model = DistributedModelParallel(Wrap(src), device=torch.device("cpu"), plan=ShardingPlan({"ec": plan}), sharders=[sharder])
        gold_opt = torch.optim.SGD(gold.parameters(), lr=0.1)

        def batch(s, Bq=B):
            g = torch.Generator().manual_seed(s)
            lengths = torch.randint(0, 4, (len(keys) * Bq,), generator=g)
            vals = torch.cat([torch.randint(0, hashes[i], (int(lengths[i * Bq : (i + 1) * Bq].sum()),), generator=g) for i in range(len(keys))])
            return KeyedJaggedTensor(keys=keys, values=vals, lengths=lengths)

        for step in range(2):
            batches = [batch(1000 * seed + 10 * step + r, B + (r + step) % 3 if uneven else B) for r in range(W)]
            out = model(batches[ctx.rank])
            gouts = [gold(b) for b in batches]
            loss = 0
            for k in keys:
                v, gv = out[k].values(), gouts[ctx.rank][k].values()
                torch.testing.assert_close(v.float(), gv, rtol=1e-5, atol=1e-5, msg=lambda m: f"EC seed {seed} step {step} key {k} plan {desc}: {m}")
                assert torch.equal(out[k].lengths(), gouts[ctx.rank][k].lengths())
                loss = loss + (v * torch.linspace(0.5, 1.5, v.shape[1])).sum()
            loss.backward()
            gold_opt.zero_grad()
            sum((go[k].values() * torch.linspace(0.5, 1.5, dim)).sum() for go in gouts for k in keys).backward()
            gold_opt.step()
        sd = model.state_dict()
        for t in tables:
            st = sd[f"ec.embeddings.{t.name}.weight"]
            ref = gold.embeddings[t.name].weight.detach()
            shards = st.local_shards() if hasattr(st, "local_shards") else []
            for sh in shards:
                o, s = sh.metadata.shard_offsets, sh.metadata.shard_sizes
                torch.testing.assert_close(sh.tensor, ref[o[0] : o[0] + s[0], o[1] : o[1] + s[1]], rtol=1e-4, atol=1e-5,
                                           msg=lambda m: f"EC seed {seed} table {t.name} ({desc[t.name]}): {m}")


@pytest.mark.parametrize("dedup", [False, True])
def test_random_sequence_sharding_matrix_4_ranks(dedup):
    run_multi_process(_run_ec, world_size=4, backend="gloo", dedup=dedup, first_seed=0)


def test_random_sequence_sharding_matrix_uneven_batch_per_rank():
    run_multi_process(_run_ec, world_size=4, backend="gloo", dedup=False, first_seed=700, n_seeds=4, uneven=True)


def _run_ckpt(ctx, tmp: str, first_seed: int, n_seeds: int = 3):
    """Checkpoint written under a random plan, loaded under another random plan (weights + fused row-wise Adagrad state), then both
    models take one more identical step and must stay identical (reference: test_model_parallel checkpoint / resharding tests)."""
    import os

    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad
    from torchrec_b200.parallel import checkpoint as ckpt
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.comm_ops import set_gradient_division
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan

    set_gradient_division(False)
    W, local, B, dev = ctx.world_size, 2, 5, ctx.device
    for seed in range(first_seed, first_seed + n_seeds):
        # (no data-parallel tables: they train with the dense optimizer, so A and B would not hold the same optimizer-state keys)
        tables, gens_a = _case(seed, False, W, local, no_col_split=True, allow_dp=False)
        _, gens_b = _case(seed + 5000, False, W, local, no_col_split=True, allow_dp=False)
        gens_b = {t.name: gens_b.get(t.name, sp.row_wise()) for t in tables}  # B's placements on A's table list

        def build(gens, init_seed):
            torch.manual_seed(init_seed)

            class M(torch.nn.Module):
                def __init__(self):
                    super().__init__()
                    self.ebc = EmbeddingBagCollection(tables, device=dev)

                def forward(self, k):
                    return self.ebc(k).values()

            m = M()
            apply_optimizer_in_backward(RowWiseAdagrad, m.ebc.parameters(), {"lr": 0.1, "eps": 1e-8})
            sharder = EmbeddingBagCollectionSharder()
            plan = sp.construct_module_sharding_plan(m.ebc, gens, sharder=sharder, world_size=W, local_size=local, device_type=dev.type)
            return DistributedModelParallel(m, device=dev, plan=ShardingPlan({"ebc": plan}), sharders=[sharder]), {n: plan[n].sharding_type for n in plan}

        a, da = build(gens_a, seed)
        kjt = _batch(tables, seed, ctx.rank, B, False).to(dev)
        for _ in range(2):
            a(kjt).sum().backward()
        path = os.path.join(tmp, f"ck{seed}")
        ckpt.save(a, a.fused_optimizer, path)
        b, db = build(gens_b, seed + 17)
        ckpt.load(b, b.fused_optimizer, path)
        msg = lambda m: f"ckpt seed {seed}: {da} -> {db}: {m}"  # noqa: E731
        torch.testing.assert_close(b(kjt), a(kjt), msg=msg)
        if not any(v == "data_parallel" for v in list(da.values()) + list(db.values())):  # (replicated tables train with the dense optimizer)
            a(kjt).sum().backward()
            b(kjt).sum().backward()
            torch.testing.assert_close(b(kjt), a(kjt), rtol=1e-5, atol=1e-6, msg=msg)


def test_random_checkpoint_resharding_4_ranks(tmp_path):
    run_multi_process(_run_ckpt, world_size=4, backend="gloo", tmp=str(tmp_path), first_seed=0)


def _run_reshard(ctx, first_seed: int, n_seeds: int = 3):
    """Live re-sharding: a model training under a random plan is moved to another random plan (weights + fused Adagrad state travel over
    point-to-point messages), trains on, is moved again - and tracks an unsharded golden model all the way."""
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.comm_ops import set_gradient_division
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan

    set_gradient_division(False)
    W, local, B, dev = ctx.world_size, 2, 5, ctx.device
    for seed in range(first_seed, first_seed + n_seeds):
        tables, gens = _case(seed, False, W, local, no_col_split=True, allow_dp=False)
        torch.manual_seed(seed)
        gold = EmbeddingBagCollection(tables, device=dev)
        src = EmbeddingBagCollection(tables, device=dev)
        src.load_state_dict(gold.state_dict())
        apply_optimizer_in_backward(RowWiseAdagrad, src.parameters(), {"lr": 0.1, "eps": 1e-8})
        sharder = EmbeddingBagCollectionSharder()

        class Wrap(torch.nn.Module):
            def __init__(self, ebc):
                super().__init__()
                self.ebc = ebc

            def forward(self, kjt):
                return self.ebc(kjt).values()

        plan = sp.construct_module_sharding_plan(src, gens, sharder=sharder, world_size=W, local_size=local, device_type=dev.type)
        model = DistributedModelParallel(Wrap(src), device=dev, plan=ShardingPlan({"ebc": plan}), sharders=[sharder])
        gold_opt = RowWiseAdagrad(gold.parameters(), lr=0.1, eps=1e-8)
        history = [{n: plan[n].sharding_type for n in plan}]

        def step(k):
            batches = [_batch(tables, 100 * seed + k, r, B, False).to(dev) for r in range(W)]
            out = model(batches[ctx.rank])
            gouts = [gold(b).values() for b in batches]
            torch.testing.assert_close(out.float(), gouts[ctx.rank], rtol=1e-5, atol=1e-5, msg=lambda m: f"reshard seed {seed} step {k} plans {history}: {m}")
            out.sum().backward()
            gold_opt.zero_grad()
            sum(o.sum() for o in gouts).backward()
            gold_opt.step()

        step(0)
        meta = EmbeddingBagCollection(tables, device=torch.device("meta"))
        for hop in (1, 2):
            _, new_gens = _case(seed + 7000 * hop, False, W, local, no_col_split=True, allow_dp=False)
            new_gens = {t.name: new_gens.get(t.name, sp.row_wise()) for t in tables}
            new = sp.construct_module_sharding_plan(meta, new_gens, sharder=sharder, world_size=W, local_size=local, device_type=dev.type)
            history.append({n: new[n].sharding_type for n in new})
            model.reshard("ebc", dict(new))
            step(2 * hop - 1)
            step(2 * hop)


def test_random_live_resharding_4_ranks():
    run_multi_process(_run_reshard, world_size=4, backend="gloo", first_seed=0)


def _vbe_batch(tables, seed: int, rank: int, full_B: int):
    """Variable batch per feature: feature k carries b_k <= full_B distinct bags and inverse indices that expand them to the full batch."""
    from torchrec_b200.sparse import KeyedJaggedTensor

    g = torch.Generator().manual_seed(104729 * seed + 13 * rank + 5)
    keys, hashes = [], []
    for t in tables:
        for f in t.feature_names:
            keys.append(f)
            hashes.append(t.num_embeddings)
    bks = [int(torch.randint(1, full_B + 1, (1,), generator=g)) for _ in keys]
    lens = [torch.randint(0, 4, (b,), generator=g) for b in bks]
    vals = [torch.randint(0, h, (int(l.sum()),), generator=g) for l, h in zip(lens, hashes)]
    inv = torch.stack([torch.randint(0, b, (full_B,), generator=g) for b in bks])
    return KeyedJaggedTensor(keys=keys, values=torch.cat(vals), lengths=torch.cat(lens), stride_per_key_per_rank=[[b] for b in bks], inverse_indices=(keys, inv))


def _run_vbe(ctx, first_seed: int, n_seeds: int = 4):
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.comm_ops import set_gradient_division
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan

    set_gradient_division(False)
    W, local, dev = ctx.world_size, 2, ctx.device
    for seed in range(first_seed, first_seed + n_seeds):
        tables, gens = _case(seed, False, W, local, allow_dp=False)
        torch.manual_seed(seed)
        gold = EmbeddingBagCollection(tables, device=dev)
        src = EmbeddingBagCollection(tables, device=dev)
        src.load_state_dict(gold.state_dict())
        apply_optimizer_in_backward(torch.optim.SGD, src.parameters(), {"lr": 0.1})
        sharder = EmbeddingBagCollectionSharder()
        plan = sp.construct_module_sharding_plan(src, gens, sharder=sharder, world_size=W, local_size=local, device_type=dev.type)
        desc = {n: plan[n].sharding_type for n in plan}

        class Wrap(torch.nn.Module):
            def __init__(self, ebc):
                super().__init__()
                self.ebc = ebc

            def forward(self, kjt):
                return self.ebc(kjt).values()

        model = DistributedModelParallel(Wrap(src), device=dev, plan=ShardingPlan({"ebc": plan}), sharders=[sharder])
        gold_opt = torch.optim.SGD(gold.parameters(), lr=0.1)
        for step in range(2):
            batches = [_vbe_batch(tables, 10 * seed + step, r, 4 + r) for r in range(W)]  # every rank has its own full batch size
            out = model(batches[ctx.rank])
            gouts = [gold(b).values() for b in batches]
            torch.testing.assert_close(out.float(), gouts[ctx.rank], rtol=1e-5, atol=1e-5, msg=lambda m: f"VBE seed {seed} step {step} plan {desc}: {m}")
            out.sum().backward()
            gold_opt.zero_grad()
            sum(o.sum() for o in gouts).backward()
            gold_opt.step()
        sd = model.state_dict()
        for t in tables:
            st = sd[f"ebc.embedding_bags.{t.name}.weight"]
            ref = gold.embedding_bags[t.name].weight.detach()
            for sh in (st.local_shards() if hasattr(st, "local_shards") else []):
                o, s = sh.metadata.shard_offsets, sh.metadata.shard_sizes
                torch.testing.assert_close(sh.tensor, ref[o[0] : o[0] + s[0], o[1] : o[1] + s[1]], rtol=1e-4, atol=1e-5,
                                           msg=lambda m: f"VBE seed {seed} table {t.name} ({desc[t.name]}): {m}")


def test_random_variable_batch_matrix_4_ranks():
    run_multi_process(_run_vbe, world_size=4, backend="gloo", first_seed=0)


def _run_pipelines(ctx, first_seed: int, n_seeds: int = 2):
    """Every pipelined training loop == the plain loop (losses step by step, weights at the end) on random DLRMs / plans at 4 ranks; the
    last batch of the stream is smaller than the others (end of an epoch)."""
    from torchrec_b200.datasets.utils import Batch
    from torchrec_b200.models.dlrm import DLRM, DLRMTrain
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.optim.keyed import CombinedOptimizer, KeyedOptimizerWrapper
    from torchrec_b200.optim.optimizers import in_backward_optimizer_filter
    from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel import train_pipeline as tp
    from torchrec_b200.parallel.comm_ops import set_gradient_division
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan

    set_gradient_division(True)
    W, local = ctx.world_size, 2
    try:
        for seed in range(first_seed, first_seed + n_seeds):
            tables, gens = _case(seed, False, W, local, no_col_split=True, allow_dp=False)
            tables = [type(t)(name=t.name, embedding_dim=8, num_embeddings=t.num_embeddings, feature_names=t.feature_names, pooling=t.pooling) for t in tables]  # DLRM: one dim

            def build():
                torch.manual_seed(seed)
                ebc = EmbeddingBagCollection(tables)
                apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": 0.05})
                model = DLRMTrain(DLRM(ebc, 5, [16, 8], [16, 1]))
                sharder = EmbeddingBagCollectionSharder()
                plan = sp.construct_module_sharding_plan(ebc, gens, sharder=sharder, world_size=W, local_size=local, device_type="cpu")
                dmp = DistributedModelParallel(model, device=torch.device("cpu"), plan=ShardingPlan({"model.sparse_arch.embedding_bag_collection": plan}), sharders=[sharder])
                dense_opt = KeyedOptimizerWrapper(dict(in_backward_optimizer_filter(dmp.named_parameters())), lambda p: torch.optim.SGD(p, lr=0.1))
                return dmp, CombinedOptimizer([dmp.fused_optimizer, dense_opt]), {n: plan[n].sharding_type for n in plan}

            def batches():
                out = []
                for i, Bq in enumerate([6, 6, 6, 6, 3]):  # the last batch is a partial one (same size on every rank)
                    g = torch.Generator().manual_seed(97 * seed + 11 * i + ctx.rank)
                    kjt = _batch(tables, 50 * seed + i, ctx.rank, Bq, False)
                    out.append(Batch(dense_features=torch.randn(Bq, 5, generator=g), sparse_features=kjt, labels=torch.randint(0, 2, (Bq,), generator=g).float()))
                return out

            data = batches()
            ref, opt_ref, desc = build()
            ref_losses = []
            for b in data:
                opt_ref.zero_grad()
                loss, _ = ref(b)
                loss.backward()
                opt_ref.step()
                ref_losses.append(loss.detach().clone())
            sa = ref.state_dict()
            for name in ("TrainPipelineBase", "TrainPipelineSparseDist", "TrainPipelineSparseDistLite", "TrainPipelineFusedSparseDist", "PrefetchTrainPipelineSparseDist"):
                dmp, opt, _ = build()
                pipe = getattr(tp, name)(dmp, opt, torch.device("cpu"))
                it = iter(data)
                losses = []
                while True:
                    try:
                        losses.append(pipe.progress(it)[0].clone())
                    except StopIteration:
                        break
                msg = lambda m: f"{name} seed {seed} plan {desc}: {m}"  # noqa: E731
                assert len(losses) == len(ref_losses), msg(f"{len(losses)} steps instead of {len(ref_losses)}")
                for a, b in zip(losses, ref_losses):
                    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6, msg=msg)
                sb = dmp.state_dict()
                for k in sa:
                    ta, tb = sa[k], sb[k]
                    if hasattr(ta, "local_shards"):
                        for x, y in zip(ta.local_shards(), tb.local_shards()):
                            torch.testing.assert_close(x.tensor, y.tensor, rtol=1e-5, atol=1e-6, msg=msg)
                    else:
                        torch.testing.assert_close(ta, tb, rtol=1e-5, atol=1e-6, msg=msg)
    finally:
        set_gradient_division(False)


def test_random_train_pipelines_match_plain_loop_4_ranks():
    run_multi_process(_run_pipelines, world_size=4, backend="gloo", first_seed=0)


def _run_fp(ctx, first_seed: int, n_seeds: int = 4):
    """Feature-processed (position-weighted) bags: random tables / placements; the forward output, the embedding rows after two fused
    SGD steps and the position-weight gradients (data-parallel) match the unsharded module."""
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
    W, local, B, dev = ctx.world_size, 2, 5, ctx.device
    for seed in range(first_seed, first_seed + n_seeds):
        tables, gens = _case(seed, True, W, local, allow_dp=False)
        feats = [f for t in tables for f in t.feature_names]

        def make():
            torch.manual_seed(seed)
            fp = PositionWeightedModuleCollection({f: 4 for f in feats}, device=dev)
            with torch.no_grad():
                for i, p in enumerate(fp.parameters()):
                    p.copy_(torch.linspace(0.5, 1.5, p.numel()) + 0.1 * i)
            return FeatureProcessedEmbeddingBagCollection(EmbeddingBagCollection(tables, is_weighted=True, device=dev), fp)

        gold, src = make(), make()
        src.load_state_dict(gold.state_dict())
        apply_optimizer_in_backward(torch.optim.SGD, src._embedding_bag_collection.parameters(), {"lr": 0.1})
        sharder = FeatureProcessedEmbeddingBagCollectionSharder()
        plan = sp.construct_module_sharding_plan(src, gens, sharder=sharder, world_size=W, local_size=local, device_type=dev.type)
        desc = {n: plan[n].sharding_type for n in plan}
        msg = lambda m: f"FP seed {seed} plan {desc}: {m}"  # noqa: E731

        class Wrap(torch.nn.Module):
            def __init__(self, m):
                super().__init__()
                self.m = m

            def forward(self, kjt):
                out = self.m(kjt)
                return (out.wait() if hasattr(out, "wait") else out).values()

        model = DistributedModelParallel(Wrap(src), device=dev, plan=ShardingPlan({"m": plan}), sharders=[sharder])
        gold_emb_opt = torch.optim.SGD(gold._embedding_bag_collection.parameters(), lr=0.1)
        gp = dict(gold._feature_processors.named_parameters())
        for step in range(2):
            batches = [_batch(tables, 10 * seed + step, r, B, False).to(dev) for r in range(W)]
            out = model(batches[ctx.rank])
            gouts = [gold(b).values() for b in batches]
            torch.testing.assert_close(out.float(), gouts[ctx.rank], rtol=1e-5, atol=1e-5, msg=msg)
            proj = torch.linspace(0.5, 1.5, out.shape[1], device=dev)
            model.zero_grad()
            (out * proj).sum().backward()
            gold.zero_grad()
            sum((o * proj).sum() for o in gouts).backward()
            gold_emb_opt.step()
            n_checked = 0
            for n, p in model.named_parameters():
                if p.requires_grad:
                    key = n.split("_feature_processors.")[-1]
                    torch.testing.assert_close(p.grad * W, gp[key].grad, rtol=1e-4, atol=1e-5, msg=msg)  # DDP averages, the golden sums
                    n_checked += 1
            assert n_checked == len(feats), msg(f"{n_checked} position-weight parameters for {len(feats)} features")
        sd = model.state_dict()
        for t in tables:
            st = sd[f"m._embedding_bag_collection.embedding_bags.{t.name}.weight"]
            ref = gold._embedding_bag_collection.embedding_bags[t.name].weight.detach()
            for sh in st.local_shards():
                o, s = sh.metadata.shard_offsets, sh.metadata.shard_sizes
                torch.testing.assert_close(sh.tensor, ref[o[0] : o[0] + s[0], o[1] : o[1] + s[1]], rtol=1e-4, atol=1e-5, msg=msg)


def test_random_feature_processed_matrix_4_ranks():
    run_multi_process(_run_fp, world_size=4, backend="gloo", first_seed=0)
