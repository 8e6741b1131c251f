"""Native dynamic-embedding runtime: id transformer, PS round trip, end-to-end cache training
(reference csrc/dynamic_embedding/details/*_test.cpp, contrib/dynamic_embedding/tests)."""
import os

import pytest
import torch


def test_id_transformer_basic_and_eviction():
    from torchrec_b200.dynamic_embedding import IDTransformer

    tr = IDTransformer(8, {"type": "lru"}, {"partitions": 1})
    ids = torch.tensor([100, 200, 100, 300])
    c, ok, fetch = tr.transform(ids, time=1)
    assert ok and c[0] == c[2] and len(set(c.tolist())) == 3 and fetch.shape == (3, 2)
    assert set(fetch[:, 0].tolist()) == {100, 200, 300}
    c2, ok, fetch2 = tr.transform(torch.tensor([200, 400, 500, 600, 700, 800]), time=2)
    assert ok and fetch2.shape[0] == 5 and len(tr) == 8
    c3, ok, _ = tr.transform(torch.tensor([900]), time=3)
    assert not ok and c3[0] == -1
    ev = tr.evict(2)  # LRU: ids last used at time 1 and not refreshed -> 100, 300
    assert set(ev[:, 0].tolist()) == {100, 300}
    c4, ok, f4 = tr.transform(torch.tensor([900]), time=4)
    assert ok and c4[0] in ev[:, 1].tolist() and f4[0, 0] == 900
    assert tr.save().shape == (7, 3)


@pytest.mark.parametrize("strategy", ["mixed_lru_lfu", "lfu", "distance_lfu"])
def test_id_transformer_partitions_consistent(strategy):
    from torchrec_b200.dynamic_embedding import IDTransformer

    tr = IDTransformer(1 << 14, {"type": strategy}, {"partitions": 8, "threads": 4})
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 1 << 40, (20000,), generator=g)
    ids = torch.cat([ids[:5000], ids[:5000], ids[5000:12000]])
    c, ok, fetch = tr.transform(ids, time=1)
    assert ok
    m = {}
    for gid, slot in zip(ids.tolist(), c.tolist()):
        assert m.setdefault(gid, slot) == slot
    assert len(set(m.values())) == len(m) and max(m.values()) < (1 << 14) and fetch.shape[0] == len(m)
    # frequently used ids survive eviction
    hot = ids[:100]
    for t in range(2, 40):
        tr.transform(hot, time=t)
    ev = tr.evict(len(m) - 1000)  # frequency outranks recency: keep enough room for the randomly promoted cold ids
    assert not (set(ev[:, 0].tolist()) & set(hot.tolist()))


@pytest.mark.parametrize("scheme", ["memory", "file"])
def test_ps_round_trip(tmp_path, scheme):
    from torchrec_b200.dynamic_embedding import PS

    url = f"memory://t_{os.getpid()}_{scheme}" if scheme == "memory" else f"file://{tmp_path}/ps"
    w, m = torch.randn(6, 4), torch.rand(6)
    ps = PS("tbl", [w, m], url, init_fn=lambda n: [torch.full((n, 4), 7.0), torch.zeros(n)])
    ps.evict(torch.tensor([[1001, 2], [1002, 5]]))
    ps.wait()
    assert len(ps) == 2
    w2, m2 = torch.zeros(6, 4), torch.zeros(6)
    ps2 = PS("tbl", [w2, m2], url, init_fn=lambda n: [torch.full((n, 4), 7.0), torch.zeros(n)])
    ps2.fetch(torch.tensor([[1002, 0], [1001, 3], [555, 1]]))
    torch.testing.assert_close(w2[0], w[5])
    torch.testing.assert_close(w2[3], w[2])
    torch.testing.assert_close(m2[3], m[2])
    assert bool((w2[1] == 7.0).all())  # unknown id -> init_fn


def test_dynamic_embedding_training_matches_full_table():
    """Cache of 16 rows over an id space of 200: training through the PS-backed cache must equal training a full table."""
    from torchrec_b200.dynamic_embedding import wrap
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.sparse import KeyedJaggedTensor

    torch.manual_seed(0)
    N, C, D = 200, 16, 4
    full = EmbeddingBagCollection([EmbeddingBagConfig(name="t", embedding_dim=D, num_embeddings=N, feature_names=["f"])])
    cache = EmbeddingBagCollection([EmbeddingBagConfig(name="t", embedding_dim=D, num_embeddings=C, feature_names=["f"])])

    class Batch:
        def __init__(self, kjt):
            self.sparse_features = kjt

    g = torch.Generator().manual_seed(1)
    batches = []
    for _ in range(30):
        lengths = torch.randint(1, 3, (4,), generator=g)
        batches.append(Batch(KeyedJaggedTensor(keys=["f"], values=torch.randint(0, N, (int(lengths.sum()),), generator=g), lengths=lengths)))
    # seed the PS with the full table so both models start from the same rows
    from torchrec_b200.dynamic_embedding import PS

    url = f"memory://dyn_{os.getpid()}"
    seed = PS("ebc.t", [full.embedding_bags["t"].weight.data.clone()], url)
    seed.evict(torch.stack([torch.arange(N), torch.arange(N)], 1))
    seed.wait()

    class M(torch.nn.Module):
        def __init__(self, ebc):
            super().__init__()
            self.ebc = ebc

    loader, colls = wrap(url, batches, M(cache))
    opt_f = torch.optim.SGD(full.parameters(), lr=0.5)
    opt_c = torch.optim.SGD(cache.parameters(), lr=0.5)
    for raw, b in zip([Batch(x.sparse_features) for x in batches], loader):
        of = full(raw.sparse_features).values()
        oc = cache(b.sparse_features).values()
        torch.testing.assert_close(oc, of)
        for o, opt in ((of, opt_f), (oc, opt_c)):
            opt.zero_grad()
            (o * o).sum().backward()
            opt.step()
    colls[0].save()
    out = torch.zeros(N, D)
    chk = PS("ebc.t", [out], url)
    chk.fetch(torch.stack([torch.arange(N), torch.arange(N)], 1))
    torch.testing.assert_close(out, full.embedding_bags["t"].weight.data)


def test_native_cpp_tests_and_benchmark():
    """The C++ test executable of the id transformer (reference test/cpp + benchmarks/cpp dynamic_embedding)."""
    import subprocess

    from torchrec_b200.csrc.build import build_native_test

    exe = build_native_test("dynemb")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "all passed" in r.stdout, r.stderr
    b = subprocess.run([exe, "--bench"], capture_output=True, text=True, timeout=300)
    assert b.returncode == 0 and "M ids/s" in b.stdout


def test_id_transformer_group_and_tensor_list(tmp_path):
    """Two collections of one model translated concurrently; rows survive eviction through the PS; TensorList exposes raw pointers."""
    import torch

    from torchrec_b200.dynamic_embedding import IDTransformerGroup, TensorList
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.sparse import KeyedJaggedTensor

    class Two(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.user = EmbeddingBagCollection([EmbeddingBagConfig(name="u", embedding_dim=4, num_embeddings=8, feature_names=["uf"])])
            self.item = EmbeddingBagCollection([EmbeddingBagConfig(name="i", embedding_dim=4, num_embeddings=8, feature_names=["if"])])

    m = Two()
    cfgs = {"user": m.user.embedding_bag_configs(), "item": m.item.embedding_bag_configs()}
    group = IDTransformerGroup(f"file://{tmp_path}", m, cfgs)
    assert "user" in group and "nope" not in group

    def kjt(key, ids):
        return KeyedJaggedTensor(keys=[key], values=torch.tensor(ids), lengths=torch.ones(len(ids), dtype=torch.int64))

    out, handles = group.transform({"user": kjt("uf", [10**12 + 1, 10**12 + 2, 10**12 + 1]), "item": kjt("if", [7_000_000_007])})
    u = out["user"].values().tolist()
    assert u[0] == u[2] != u[1] and all(0 <= x < 8 for x in u) and 0 <= int(out["item"].values()[0]) < 8 and set(handles) == {"user", "item"}
    # mark the row of id 10^12+1, then push it out of the 8-row cache with 20 other ids, then bring it back: the marked values return
    with torch.no_grad():
        m.user.embedding_bags["u"].weight[u[0]].fill_(42.0)
    for k in range(0, 20, 4):
        group.transform({"user": kjt("uf", [5_000 + k, 5_001 + k, 5_002 + k, 5_003 + k])})
    back, _ = group.transform({"user": kjt("uf", [10**12 + 1])})
    assert m.user.embedding_bags["u"].weight[int(back["user"].values()[0])].tolist() == [42.0] * 4
    group.save()
    group.close()

    tl = TensorList([torch.zeros(4, 3), torch.zeros(4, dtype=torch.int64)])
    assert len(tl) == 2 and tl.row_bytes() == [12, 8] and list(tl.nbytes()) == [48, 32] and list(tl.dtype_codes()) == [0, 3] and tl.pointers()[0] == tl[0].data_ptr()
    with pytest.raises(ValueError):
        TensorList([torch.zeros(4, 3).t()])
