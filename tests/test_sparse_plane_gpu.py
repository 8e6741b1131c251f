"""The NVLink sparse plane on ONE GPU: W virtual ranks (``LoopbackGroup``) run the real kernels - kjt_route (bucketize + permute + peer
write), table-batched lookup with multi-destination stores, staging reduce, gradient push, fused backward over per-source id
regions, the device barrier - with every "peer" buffer on the local device, driven in lock step, and are checked against an fp32
PyTorch reference of the whole sharded module (forward values and weights / optimizer state after the update).

What the single-GPU box of the driver can verify of the multi-GPU data plane: pointer tables, layouts, routing, reductions, optimizer.
"""
import ctypes

import pytest
import torch

from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig, PoolingType
from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
from torchrec_b200.ops import _lib
from torchrec_b200.ops.tbe import OptimType
from torchrec_b200.parallel import sharding_plan as sp
from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
from torchrec_b200.parallel.engine import OptimizerSpec, ShardedLookupEngine
from torchrec_b200.parallel.sparse_plane import LoopbackGroup
from torchrec_b200.parallel.types import ShardingEnv
from torchrec_b200.sparse.jagged_tensor import KeyedJaggedTensor

pytestmark = pytest.mark.gpu


def _tables(dims, rows, pooling):
    return [EmbeddingBagConfig(name=f"t{i}", embedding_dim=d, num_embeddings=r, feature_names=[f"f{i}"] if i != 1 else ["f1", "f1b"], pooling=pooling)
            for i, (d, r) in enumerate(zip(dims, rows))]


def _plan(kind, tables, W):
    gens = {}
    for i, t in enumerate(tables):
        k = kind if kind != "mixed" else ["tw", "rw", "cw", "tw", "rw"][i % 5]
        if k == "tw":
            gens[t.name] = sp.table_wise(rank=i % W)
        elif k == "rw":
            if i % 2 == 0:
                gens[t.name] = sp.row_wise()
            else:  # uneven row shards, one of them empty
                sizes = [0] * W
                left = t.num_embeddings
                for r in range(W - 1):
                    sizes[r] = 0 if r == 1 else min(left, t.num_embeddings // W + 3 * r)
                    left -= sizes[r]
                sizes[W - 1] = left
                gens[t.name] = sp.row_wise((sizes, "cuda"))
        elif k == "cw":
            n = 2 if t.embedding_dim % 8 == 0 else 1
            gens[t.name] = sp.column_wise(ranks=[(i + j) % W for j in range(n)])
    return gens


def _kjt(tables, B, maxL, weighted, seed, device, one_hot_first=True):
    g = torch.Generator().manual_seed(seed)
    keys, lens, vals = [], [], []
    for ti, t in enumerate(tables):
        for fn in t.feature_names:
            keys.append(fn)
            if one_hot_first and ti == 0:
                l = torch.ones(B, dtype=torch.int64)
            else:
                l = torch.randint(0, maxL + 1, (B,), generator=g)
            lens.append(l)
            v = torch.randint(0, t.num_embeddings, (int(l.sum()),), generator=g)
            if ti == 2 and v.numel() > 3:
                v[:3] = torch.tensor([-1, t.num_embeddings, t.num_embeddings + 7])[: min(3, v.numel())]  # invalid ids contribute zero
            vals.append(v)
    values = torch.cat(vals)
    w = torch.rand(values.numel(), generator=g) if weighted else None
    return KeyedJaggedTensor(keys=keys, values=values.to(device), lengths=torch.cat(lens).to(device), weights=w.to(device) if w is not None else None, stride=B)


def _reference_forward(tables, weights, kjt, mean):
    """fp32 pooled embeddings [B, sum D] of one rank's batch from the FULL tables."""
    B = kjt.stride()
    outs = []
    jd = kjt.to_dict()
    for t in tables:
        W_ = weights[t.name]
        for fn in t.feature_names:
            jt = jd[fn]
            v, l = jt.values().long(), jt.lengths().long()
            ok = (v >= 0) & (v < t.num_embeddings)
            psw = jt.weights_or_none()
            w = ok.float() * (psw.float() if psw is not None else 1.0)
            rows = W_[v.clamp(0, t.num_embeddings - 1)] * w.unsqueeze(1)
            seg = torch.repeat_interleave(torch.arange(B, device=v.device), l)
            out = torch.zeros(B, t.embedding_dim, device=v.device).index_add_(0, seg, rows)
            if mean:
                out = out / l.clamp(min=1).unsqueeze(1)
            outs.append(out)
    return torch.cat(outs, 1)


@pytest.mark.parametrize("W", [2, 4, 8])
@pytest.mark.parametrize("kind", ["tw", "rw", "cw", "mixed"])
@pytest.mark.parametrize("weighted,mean", [(False, False), (True, False), (False, True)])
def test_plane_forward_backward_vs_reference(W, kind, weighted, mean):
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    dims = [128, 64, 128, 32, 256]
    rows = [500, 37, 1201, 64, 90]
    B = 64 if W != 4 else 48  # 48: chunks of 32 with a tail inside every source region
    pooling = PoolingType.MEAN if mean else PoolingType.SUM
    tables = _tables(dims, rows, pooling)
    ebc = EmbeddingBagCollection(tables=tables, is_weighted=weighted, device=torch.device("cpu"))
    plan = sp.construct_module_sharding_plan(ebc, _plan(kind, tables, W), sharder=EmbeddingBagCollectionSharder(), world_size=W, local_size=W, device_type="cuda")
    full = {t.name: (torch.randn(t.num_embeddings, t.embedding_dim, device=dev) * 0.1) for t in tables}
    feature_names = [f for t in tables for f in t.feature_names]
    feature_table = [ti for ti, t in enumerate(tables) for _ in t.feature_names]
    total_cols = sum(tables[ti].embedding_dim for ti in feature_table)
    lr = 0.1
    group = LoopbackGroup(W, dev)
    engines, kjts = [], []
    for r in range(W):
        env = ShardingEnv.from_loopback(W, r, group)
        eng = ShardedLookupEngine(tables, feature_names, feature_table, plan, env, dev, pooled=True, is_weighted=weighted,
                                  opt_specs={t.name: OptimizerSpec(optim=OptimType.EXACT_ROWWISE_ADAGRAD, lr=lr, eps=1e-3) for t in tables}, output_dtype=torch.float32)
        for shard, wview, _st, _tbe in eng.local_shard_views():
            wview.copy_(full[shard.name][shard.row_off : shard.row_off + shard.rows, shard.col_off : shard.col_off + shard.cols])
        engines.append(eng)
        kjts.append(_kjt(tables, B, 5, weighted, 100 + r, dev))
    # ---- phase 1: input dist (every virtual rank pushes into every owner's receive regions) ----
    ids = [engines[r].plane_input_dist(kjts[r], None, total_cols, capacity=6 * B * 5) for r in range(W)]
    planes = [i.plane for i in ids]
    torch.cuda.synchronize()
    assert all(int(p.overflow.item()) == 0 for p in planes)
    # ---- phase 2: lookup + output dist, then the staging reduce ----
    for r in range(W):
        planes[r].forward_kernels(ids[r], 0)
    outs = [planes[r].forward_finish(0).clone() for r in range(W)]
    # the engine reduces MEAN pooling of row-sharded tables after the sum (divisor applied by the sharded module): mirror it
    for r in range(W):
        ref = _reference_forward(tables, full, kjts[r], mean)
        got = outs[r]
        if mean and any(engines[r]._post_mean_feature):
            lens = kjts[r].lengths().view(len(feature_names), B).t().float().clamp(min=1)
            c = 0
            for fi, ti in enumerate(feature_table):
                d = tables[ti].embedding_dim
                if engines[r]._post_mean_feature[fi]:
                    got[:, c : c + d] = got[:, c : c + d] / lens[:, fi : fi + 1]
                c += d
        torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5)
    # ---- phase 3: gradient dist + fused row-wise Adagrad ----
    grads = [torch.randn(B, total_cols, device=dev) for _ in range(W)]
    scale = 1.0 / W
    for r in range(W):
        g_in = grads[r].clone()
        if mean and any(engines[r]._post_mean_feature):  # the sharded module applies 1 / L outside the engine for these features
            lens = kjts[r].lengths().view(len(feature_names), B).t().float().clamp(min=1)
            c = 0
            for fi, ti in enumerate(feature_table):
                d = tables[ti].embedding_dim
                if engines[r]._post_mean_feature[fi]:
                    g_in[:, c : c + d] /= lens[:, fi : fi + 1]
                c += d
        planes[r].backward_push(g_in)
    if W == 4:  # split backward: id-dependent half (keys + sort) first, e.g. while the forward is still running, then the gradient half
        for r in range(W):
            planes[r].prepare_backward(ids[r].slot, fork=(r % 2 == 0))
    for r in range(W):
        planes[r].backward_apply(ids[r], scale)
    torch.cuda.synchronize()
    # reference: dense gradient of every table from all ranks' batches -> row-wise Adagrad (per column shard state)
    col_shards = {t.name: sorted({(s.col_off, s.cols) for s in engines[0]._table_shards[ti]}) for ti, t in enumerate(tables)}
    for ti, t in enumerate(tables):
        gw = torch.zeros_like(full[t.name])
        for r in range(W):
            jd = kjts[r].to_dict()
            c = 0
            for fi, tj in enumerate(feature_table):
                d = tables[tj].embedding_dim
                if tj == ti:
                    jt = jd[feature_names[fi]]
                    v, l = jt.values().long(), jt.lengths().long()
                    ok = (v >= 0) & (v < t.num_embeddings)
                    seg = torch.repeat_interleave(torch.arange(B, device=dev), l)
                    g = grads[r][:, c : c + d][seg] * scale
                    psw = jt.weights_or_none()
                    if psw is not None:
                        g = g * psw.float().unsqueeze(1)
                    if mean:
                        g = g / l.clamp(min=1)[seg].unsqueeze(1)
                    gw.index_add_(0, v[ok], g[ok])
                c += d
        expect = full[t.name].clone()
        for co, cn in col_shards[t.name]:  # row-wise Adagrad state is per (row, column shard)
            gs = gw[:, co : co + cn]
            touched = gs.abs().sum(1) > 0
            st = (gs * gs).sum(1) / cn
            expect[:, co : co + cn] -= (lr / (st.sqrt() + 1e-3)).unsqueeze(1) * gs * touched.unsqueeze(1)
        got = torch.zeros_like(expect)
        for r in range(W):
            for shard, wview, _st, _tbe in engines[r].local_shard_views():
                if shard.name == t.name:
                    got[shard.row_off : shard.row_off + shard.rows, shard.col_off : shard.col_off + shard.cols] = wview
        torch.testing.assert_close(got, expect, rtol=2e-4, atol=2e-5)


def test_routed_ids_jagged_view_matches_eager_route():
    """kjt_route vs the PyTorch route + all-to-all emulation: same ids per (unit, source rank, sample)."""
    dev = torch.device("cuda:0")
    W, B = 4, 32
    tables = _tables([64, 32, 128, 32, 64], [300, 41, 999, 64, 77], PoolingType.SUM)
    ebc = EmbeddingBagCollection(tables=tables, device=torch.device("cpu"))
    plan = sp.construct_module_sharding_plan(ebc, _plan("mixed", tables, W), sharder=EmbeddingBagCollectionSharder(), world_size=W, local_size=W, device_type="cuda")
    feature_names = [f for t in tables for f in t.feature_names]
    feature_table = [ti for ti, t in enumerate(tables) for _ in t.feature_names]
    total_cols = sum(tables[ti].embedding_dim for ti in feature_table)
    group = LoopbackGroup(W, dev)
    engines = [ShardedLookupEngine(tables, feature_names, feature_table, plan, ShardingEnv.from_loopback(W, r, group), dev, pooled=True, is_weighted=False, opt_specs={})
               for r in range(W)]
    kjts = [_kjt(tables, B, 9, False, 5 + r, dev, one_hot_first=False) for r in range(W)]
    ids = [engines[r].plane_input_dist(kjts[r], None, total_cols, capacity=6 * B * 9) for r in range(W)]
    import os

    os.environ["TRB_ROUTE_EAGER"] = "1"  # PyTorch reference of the routing (index arithmetic + stable sort)
    try:
        ref = [engines[r].route(kjts[r]) for r in range(W)]
    finally:
        os.environ.pop("TRB_ROUTE_EAGER", None)
    routed = [x[0] for x in ref]
    # the portable transport's device-side routing (one local destination) produces the very same routed KJT + unbucketize permutation
    for r in range(W):
        got, unb = engines[r].route(kjts[r])
        assert torch.equal(got.lengths().cpu(), routed[r].lengths().cpu()) and torch.equal(got.values().cpu(), routed[r].values().cpu())
        assert torch.equal(unb.cpu(), ref[r][1].cpu())
    for d in range(W):
        got = ids[d].to_kjt()
        eng = engines[d]
        U0, U1 = eng._unit_start[d], eng._unit_start[d + 1]
        exp_vals, exp_lens = [], []
        for u in range(U0, U1):
            for s in range(W):
                off = routed[s].offsets()
                a, b = int(off[u * B]), int(off[(u + 1) * B])
                exp_vals.append(routed[s].values()[a:b])
                exp_lens.append(routed[s].lengths()[u * B : (u + 1) * B])
        assert torch.equal(got.lengths().cpu(), torch.cat(exp_lens).cpu().to(got.lengths().dtype))
        assert torch.equal(got.values().cpu(), torch.cat(exp_vals).cpu())


def test_input_dist_overflow_is_reported():
    dev = torch.device("cuda:0")
    W, B = 2, 32
    tables = _tables([32, 32], [50, 60], PoolingType.SUM)
    ebc = EmbeddingBagCollection(tables=tables, device=torch.device("cpu"))
    plan = sp.construct_module_sharding_plan(ebc, _plan("tw", tables, W), sharder=EmbeddingBagCollectionSharder(), world_size=W, local_size=W, device_type="cuda")
    feature_names = [f for t in tables for f in t.feature_names]
    feature_table = [ti for ti, t in enumerate(tables) for _ in t.feature_names]
    group = LoopbackGroup(W, dev)
    eng = ShardedLookupEngine(tables, feature_names, feature_table, plan, ShardingEnv.from_loopback(W, 0, group), dev, pooled=True, is_weighted=False, opt_specs={})
    small = _kjt(tables, B, 1, False, 1, dev)
    ids = eng.plane_input_dist(small, None, 96)
    big = _kjt(tables, B, 40, False, 2, dev, one_hot_first=False)
    eng.plane_input_dist(big, None, 96)
    torch.cuda.synchronize()
    assert int(ids.plane.overflow.item()) == 1


@pytest.mark.parametrize("W", [2, 8])
def test_device_barrier_virtual_ranks(W):
    """W single-CTA barrier kernels on W streams of one device: each publishes its epoch to every pad and spins for the others."""
    dev = torch.device("cuda:0")
    L = _lib.lib()
    pads = [torch.zeros(64, dtype=torch.int32, device=dev) for _ in range(W)]
    epochs = torch.zeros(W, dtype=torch.int32, device=dev)
    arr = _lib.ptr_array([p.data_ptr() for p in pads])
    streams = [torch.cuda.Stream(dev) for _ in range(W)]
    torch.cuda.synchronize()
    for it in range(3):
        for r in range(W):
            with torch.cuda.stream(streams[r]):
                code = L.trb_barrier(arr, W, r, ctypes.c_void_p(epochs.data_ptr() + 4 * r), _lib.stream_ptr(dev))
                assert code == 0
        torch.cuda.synchronize()
        assert epochs.tolist() == [it + 1] * W
        for p in pads:
            assert p[:W].tolist() == [it + 1] * W


def _train_dlrm(plane: bool, weighted_fp: bool):
    import os

    from torchrec_b200.datasets.random import RandomRecDataset
    from torchrec_b200.models.dlrm import DLRM, DLRMTrain
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan

    os.environ["TRB_PLANE_SINGLE"] = "1" if plane else "0"
    try:
        dev = torch.device("cuda:0")
        torch.manual_seed(5)
        keys = [f"f{i}" for i in range(6)]
        hashes = [300, 17, 4000, 64, 9, 1000]
        ebc = EmbeddingBagCollection([EmbeddingBagConfig(name=f"t{i}", embedding_dim=64, num_embeddings=h, feature_names=[keys[i]]) for i, h in enumerate(hashes)],
                                     device=torch.device("meta"))
        apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": 0.05})
        model = DLRMTrain(DLRM(ebc, 13, [32, 64], [64, 1], dense_device=dev))
        plan = sp.construct_module_sharding_plan(ebc, {f"t{i}": sp.table_wise(rank=0) for i in range(6)}, sharder=EmbeddingBagCollectionSharder(), world_size=1,
                                                 local_size=1, device_type="cuda")
        torch.manual_seed(9)
        dmp = DistributedModelParallel(model, device=dev, plan=ShardingPlan({"model.sparse_arch.embedding_bag_collection": plan}), sharders=[EmbeddingBagCollectionSharder()])
        with torch.no_grad():
            g = torch.Generator(device="cpu").manual_seed(1)
            for _, w, _st, _tbe in dmp.module.model.sparse_arch.embedding_bag_collection._engine.local_shard_views():
                w.copy_(torch.randn(w.shape, generator=g) * 0.1)
            for p in dmp.parameters():
                if p.requires_grad:
                    p.copy_(torch.randn(p.shape, generator=g).to(p.device) * 0.05)
        opt = torch.optim.SGD([p for p in dmp.parameters() if p.requires_grad], lr=0.02)
        ds = RandomRecDataset(keys, 96, hash_sizes=hashes, ids_per_features=[1, 3, 2, 1, 4, 2], min_ids_per_features=[1, 0, 0, 1, 0, 0], num_dense=13, manual_seed=3,
                              num_generated_batches=6)
        it = iter(ds)
        losses = []
        for _ in range(6):  # > 3 steps: the id slots wrap around and the backward replays as a CUDA graph
            b = next(it).to(dev)
            opt.zero_grad()
            loss, _ = dmp(b)
            loss.backward()
            opt.step()
            losses.append(float(loss))
        eng = dmp.module.model.sparse_arch.embedding_bag_collection._engine
        used_plane = any(p.capacity > 0 for p in eng.__dict__.get("_planes", {}).values())
        weights = torch.cat([w.flatten() for _, w, _s, _t in eng.local_shard_views()]).clone()
        return losses, weights, used_plane
    finally:
        os.environ.pop("TRB_PLANE_SINGLE", None)


def test_single_gpu_training_through_the_plane_matches_the_plain_kernels():
    """One GPU: static-shape plane (route kernel, per-slot presort, graph replay, gradient read in place) vs the eager table-batched path."""
    l_plane, w_plane, used = _train_dlrm(True, False)
    l_plain, w_plain, used_plain = _train_dlrm(False, False)
    assert used and not used_plain
    torch.testing.assert_close(torch.tensor(l_plane), torch.tensor(l_plain), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(w_plane, w_plain, rtol=1e-4, atol=1e-6)
