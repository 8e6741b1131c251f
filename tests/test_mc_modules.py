import torch

from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
from torchrec_b200.modules.mc_embedding_modules import ManagedCollisionEmbeddingBagCollection
from torchrec_b200.modules.mc_modules import DistanceLFU_EvictionPolicy, LFU_EvictionPolicy, LRU_EvictionPolicy, ManagedCollisionCollection, MCHManagedCollisionModule
from torchrec_b200.sparse import JaggedTensor, KeyedJaggedTensor


def _jt(v):
    return JaggedTensor(values=torch.tensor(v, dtype=torch.int64), lengths=torch.tensor([len(v)]))


def test_sorted_zch_admits_frequent_ids_and_is_stable():
    mc = MCHManagedCollisionModule(zch_size=4, device=torch.device("cpu"), eviction_policy=LFU_EvictionPolicy(), eviction_interval=1)
    mc.train()
    ids = [1000, 1000, 2000, 3000, 3000, 3000, 4000, 5000]
    mc({"f": _jt(ids)})  # profile -> after the interval the 4 most frequent ids own the slots
    out = mc({"f": _jt([1000, 3000, 2000])})["f"].values()
    assert len(set(out.tolist())) == 3 and all(0 <= x < 4 for x in out.tolist())
    again = mc({"f": _jt([1000, 3000, 2000])})["f"].values()
    assert torch.equal(out, again)  # admitted ids keep their slot
    assert int(mc.open_slots()) == 0
    # a new very frequent id evicts the weakest resident
    mc({"f": _jt([9000] * 20)})
    ev = mc.evict()
    assert ev is not None and ev.numel() == 1
    nine = mc({"f": _jt([9000])})["f"].values()
    assert int(nine) == int(ev[0])


def test_lru_and_distance_lfu_policies_run():
    for pol in (LRU_EvictionPolicy(), DistanceLFU_EvictionPolicy()):
        mc = MCHManagedCollisionModule(zch_size=8, device=torch.device("cpu"), eviction_policy=pol, eviction_interval=2, mch_size=2)
        mc.train()
        for step in range(6):
            out = mc({"f": _jt(list(range(step * 3, step * 3 + 5)))})["f"].values()
            assert out.min() >= 0 and out.max() < 8


def test_mc_ebc_end_to_end():
    tables = [EmbeddingBagConfig(name="t", embedding_dim=8, num_embeddings=16, feature_names=["f"])]
    mcc = ManagedCollisionCollection({"t": MCHManagedCollisionModule(zch_size=16, device=torch.device("cpu"), eviction_policy=LFU_EvictionPolicy(), eviction_interval=1)}, tables)
    m = ManagedCollisionEmbeddingBagCollection(EmbeddingBagCollection(tables), mcc, return_remapped_features=True)
    kjt = KeyedJaggedTensor(keys=["f"], values=torch.tensor([10**12, 5, 10**12 + 7, 5]), lengths=torch.tensor([2, 2]))
    out, remapped = m(kjt)
    assert out.values().shape == (2, 8) and remapped.values().max() < 16


def test_mc_adapters_and_pruning_logger():
    import torch

    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig, EmbeddingConfig
    from torchrec_b200.modules.mc_adapter import McEmbeddingBagCollectionAdapter, McEmbeddingCollectionAdapter
    from torchrec_b200.modules.pruning_logger import PruningLoggerDefault
    from torchrec_b200.sparse import KeyedJaggedTensor

    dev = torch.device("cpu")
    kjt = KeyedJaggedTensor(keys=["f0", "f1"], values=torch.tensor([10**12 + 5, 7, 10**12 + 5, 99, 123456789]), lengths=torch.tensor([2, 1, 1, 1]))
    for method in ("mpzch", "sort_zch"):
        ec = McEmbeddingCollectionAdapter([EmbeddingConfig(name="t0", embedding_dim=8, num_embeddings=64, feature_names=["f0"]),
                                           EmbeddingConfig(name="t1", embedding_dim=8, num_embeddings=64, feature_names=["f1"])], input_hash_size=2**50, device=dev, world_size=1,
                                          zch_method=method, mpzch_num_buckets=4, embedding_device=dev)
        ec.train()
        out = ec(kjt)
        assert out["f0"].values().shape == (3, 8) and ec.remapped_ids is not None
        rem = ec.remapped_ids["f0"].values()
        assert int(rem.max()) < 64 and int(rem[0]) == int(rem[2])                 # huge raw ids land inside the table, same id -> same slot
        assert len(list(ec.parameters())) == 2 and [c.name for c in ec.embedding_bag_configs()] == ["t0", "t1"]
    ebc = McEmbeddingBagCollectionAdapter([EmbeddingBagConfig(name="b0", embedding_dim=8, num_embeddings=32, feature_names=["f0", "f1"])], input_hash_size=2**50, device=dev,
                                          world_size=1, zch_method="mpzch", mpzch_num_buckets=2, embedding_device=dev)
    ebc.train()
    pooled = ebc(kjt)
    assert pooled.values().shape == (2, 16)
    import pytest

    with pytest.raises(NotImplementedError):
        McEmbeddingCollectionAdapter([EmbeddingConfig(name="t", embedding_dim=8, num_embeddings=8, feature_names=["f"])], 100, dev, 1, zch_method="")
    with PruningLoggerDefault.pruning_logger(event="prune", trainer="t0") as log:
        log.rows_pruned = 5
    assert log.rows_pruned == 5


def test_raw_id_tracker_streams_latest_raw_id_per_row():
    import torch

    from torchrec_b200.modules.embedding_configs import EmbeddingConfig
    from torchrec_b200.modules.mc_modules import LFU_EvictionPolicy, ManagedCollisionCollection, MCHManagedCollisionModule
    from torchrec_b200.parallel.model_tracker.trackers import RawIdTracker
    from torchrec_b200.sparse import KeyedJaggedTensor

    dev = torch.device("cpu")
    mcc = ManagedCollisionCollection({"t": MCHManagedCollisionModule(zch_size=16, device=dev, eviction_policy=LFU_EvictionPolicy(), eviction_interval=1, input_hash_size=1 << 40)},
                                     [EmbeddingConfig(name="t", num_embeddings=16, embedding_dim=4, feature_names=["f"])])
    tracker = RawIdTracker(mcc, consumers=["a", "b"])
    assert list(tracker.get_tracked_modules()) == ["_managed_collision_modules.t"]
    raw1 = torch.tensor([1001, 2002, 1001, 3003])
    out1 = mcc(KeyedJaggedTensor.from_lengths_sync(["f"], raw1, torch.tensor([2, 2])))
    tracker.step()
    raw2 = torch.tensor([2002, 4004])
    out2 = mcc(KeyedJaggedTensor.from_lengths_sync(["f"], raw2, torch.tensor([1, 1])))
    m = tracker.get_raw_id_map("a")["_managed_collision_modules.t"]
    rows = torch.cat([out1.values(), out2.values()])
    raws = torch.cat([raw1, raw2])
    assert set(m["ids"].tolist()) == set(rows.tolist())
    for row, raw in zip(m["ids"].tolist(), m["raw_ids"].tolist()):
        last = [int(r) for r, x in zip(raws.tolist(), rows.tolist()) if x == row][-1]
        assert raw == last
    assert tracker.get_raw_id_map("a") == {}                       # consumer a is up to date
    assert "_managed_collision_modules.t" in tracker.get_unique("b")  # consumer b still sees everything
    mcc.eval()
    mcc(KeyedJaggedTensor.from_lengths_sync(["f"], raw2, torch.tensor([1, 1])))
    assert tracker.get_raw_id_map("a") == {}                       # eval lookups are not tracked


def test_train_input_mapper_and_merged_serving_module():
    """MPZCH trained row-wise sharded, served from one merged identity table: every id is probed inside the range of the rank that
    owned it in training and gets the slot training gave it; ids never seen fall back inside that range."""
    from torchrec_b200.modules.hash_mc_modules import HashZchManagedCollisionModule, TrainInputMapper
    from torchrec_b200.modules.mc_modules import apply_mc_method_to_jt_dict
    from torchrec_b200.parallel.mc_modules import _owner_of
    from torchrec_b200.sparse.jagged_tensor import JaggedTensor

    W = 4
    full = HashZchManagedCollisionModule(zch_size=256, device=torch.device("cpu"), total_num_buckets=W)
    segs = [64 * i for i in range(W + 1)]
    shards = [full.rebuild_with_output_id_range((segs[r], segs[r + 1]), segs) for r in range(W)]
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 10**9, (150,), generator=g)
    owner = _owner_of(ids, W)
    train_slots = torch.empty_like(ids)
    for r in range(W):
        m = owner == r
        shards[r].train()
        train_slots[m] = shards[r]({"f": JaggedTensor(values=ids[m], lengths=torch.ones(int(m.sum()), dtype=torch.long))})["f"].values()
    merged = HashZchManagedCollisionModule.merge_trained_shards(shards)
    feats = {"f": JaggedTensor(values=ids, lengths=torch.ones(150, dtype=torch.long))}
    got = apply_mc_method_to_jt_dict(merged, "remap", feats)["f"].values()
    assert torch.equal(got, train_slots)
    before = merged._hash_zch_identities.clone()
    unseen = torch.randint(10**9, 2 * 10**9, (40,), generator=g)
    out = merged({"f": JaggedTensor(values=unseen, lengths=torch.ones(40, dtype=torch.long))})["f"].values()
    assert torch.equal(merged._hash_zch_identities, before), "serving never inserts"
    lo = torch.tensor(segs)[_owner_of(unseen, W)]
    assert bool(((out >= lo) & (out < lo + 64)).all())
    # the reference's dispatch rules
    sizes, offs = torch.tensor([10, 20, 30]), torch.tensor([0, 10, 30])
    mod = TrainInputMapper(input_hash_size=0, total_num_buckets=3, size_per_rank=sizes, train_rank_offsets=offs, inference_dispatch_div_train_world_size=True)
    v, s, o = mod(torch.tensor([0, 1, 2, 7]))
    assert v.tolist() == [0, 0, 0, 2] and s.tolist() == [10, 20, 30, 20] and o.tolist() == [0, 10, 30, 10]
    blk = TrainInputMapper(input_hash_size=100, total_num_buckets=3, size_per_rank=sizes, train_rank_offsets=offs)
    v, s, o = blk(torch.tensor([0, 33, 34, 99]), output_offset=torch.tensor(10))
    assert v.tolist() == [0, 33, 34, 99] and s.tolist() == [10, 10, 20, 30] and o.tolist() == [-10, -10, 0, 20]
