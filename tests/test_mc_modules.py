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
