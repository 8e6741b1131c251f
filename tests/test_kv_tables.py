"""Key-value (virtual) embedding tables: arbitrary 64-bit keys, first-touch rows, DRAM and SSD (memory-mapped file) stores behind the row
cache, store eviction, sparse snapshot, and training equivalence with a dense table that holds the same rows. CPU path (TRB_UVM_ON_CPU-style:
the cache module runs its PyTorch kernels); the GPU path shares everything but the row mover."""
import os

import pytest
import torch

from torchrec_b200.ops.kv_tbe import KeyValueEmbeddingBags
from torchrec_b200.ops.tbe import OptimType, TableBatchedEmbeddingBags


def _batch(keys_per_feature, B, seed):
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(0, 4, (len(keys_per_feature) * B,), generator=g)
    vals = torch.cat([pool[torch.randint(0, pool.numel(), (int(lengths[f * B : (f + 1) * B].sum()),), generator=g)] for f, pool in enumerate(keys_per_feature)])
    return vals, torch.cat([torch.zeros(1, dtype=torch.int64), lengths.cumsum(0)])


@pytest.mark.parametrize("backend", ["dram", "ssd"])
def test_kv_training_matches_dense_rows(backend, tmp_path):
    """Sparse, huge keys train exactly like a dense table holding the same rows (declared key-space width 2^12 only sets the init range)."""
    torch.manual_seed(0)
    D, B = 16, 12
    kw = dict(optimizer=OptimType.EXACT_ROWWISE_ADAGRAD, learning_rate=0.1, eps=1e-3)
    kv = KeyValueEmbeddingBags([(1 << 12, D), (1 << 12, D)], [0, 1], store_rows=[64, 64], backend=backend, ssd_storage_directory=str(tmp_path / "kv"),
                               cache_load_factor=0.5, min_cache_rows=40, **kw)
    pools = [torch.randint(1 << 33, 1 << 39, (30,)), torch.randint(1 << 20, 1 << 38, (25,))]
    dense = TableBatchedEmbeddingBags([(30, D), (25, D)], [0, 1], **kw)
    # touch every key once so that all first-touch rows exist, then mirror them into the dense twin by pool position
    for f in range(2):
        kv._keys_to_slots(f, pools[f])
        keys_f, rows_f = kv.key_value_snapshot(f)
        order = (keys_f.unsqueeze(1) == pools[f].unsqueeze(0)).long().argmax(1)
        dense.split_embedding_weights()[f][order] = rows_f
    for step in range(4):
        keys, off = _batch(pools, B, step)
        pos = keys.clone()
        bounds = [int(off[0]), int(off[B]), int(off[2 * B])]
        for f in range(2):
            seg = keys[bounds[f] : bounds[f + 1]]
            pos[bounds[f] : bounds[f + 1]] = (seg.unsqueeze(1) == pools[f].unsqueeze(0)).long().argmax(1)
        out_kv, out_d = kv(keys, off, batch_size=B), dense(pos, off, batch_size=B)
        torch.testing.assert_close(out_kv, out_d, rtol=1e-5, atol=1e-6)  # step > 0: also proves the previous fused updates agree
        g = torch.randn_like(out_kv)
        out_kv.backward(g)
        out_d.backward(g)
    for f in range(2):
        keys_f, rows_f = kv.key_value_snapshot(f)
        order = (keys_f.unsqueeze(1) == pools[f].unsqueeze(0)).long().argmax(1)
        torch.testing.assert_close(rows_f, dense.split_embedding_weights()[f][order], rtol=1e-5, atol=1e-6)
    assert kv.kv_stats["inserted"] == 55
    if backend == "ssd":
        kv.close()
        assert os.path.getsize(os.path.join(str(tmp_path / "kv"), "weights.bin")) == (64 + 64) * D * 4


def test_kv_store_eviction_and_fresh_rows_are_deterministic():
    D = 8
    kv = KeyValueEmbeddingBags([(1 << 50, D)], [0], store_rows=[16], cache_load_factor=1.0, min_cache_rows=16, optimizer=OptimType.EXACT_SGD, learning_rate=0.0)
    off = torch.arange(0, 9)
    a = kv(torch.arange(100, 108), off, batch_size=8).detach().clone()
    kv(torch.arange(200, 208), off, batch_size=8)
    kv(torch.arange(300, 308), off, batch_size=8)  # 24 distinct keys through a 16-row store: the coldest were forgotten
    assert kv.kv_stats["store_evictions"] >= 8
    again = kv(torch.arange(100, 108), off, batch_size=8).detach()
    torch.testing.assert_close(again, a)  # lr = 0: a forgotten key comes back with the same deterministic first-touch row
    other = KeyValueEmbeddingBags([(1 << 50, D)], [0], store_rows=[16], cache_load_factor=1.0, min_cache_rows=16, optimizer=OptimType.EXACT_SGD, learning_rate=0.0)
    torch.testing.assert_close(other(torch.arange(100, 108), off, batch_size=8).detach(), a)  # and on any other rank / after a restart
    with pytest.raises(RuntimeError, match="more distinct keys"):
        kv(torch.arange(1000, 1032), torch.arange(0, 33), batch_size=32)


def test_ssd_store_survives_reopen(tmp_path):
    D = 8
    d = str(tmp_path / "store")
    kw = dict(store_rows=[32], backend="ssd", ssd_storage_directory=d, cache_load_factor=0.5, min_cache_rows=8, optimizer=OptimType.EXACT_SGD, learning_rate=0.5)
    kv = KeyValueEmbeddingBags([(1 << 30, D)], [0], **kw)
    keys, off = torch.tensor([7, 9, 7, 11]), torch.arange(0, 5)
    kv(keys, off, batch_size=4).sum().backward()
    keys_a, rows_a = kv.key_value_snapshot(0)
    kv.close()
    mapped = torch.from_numpy(__import__("numpy").memmap(os.path.join(d, "weights.bin"), dtype="float32", mode="r", shape=(32 * D,)).copy()).view(32, D)
    triples = kv.id_maps[0].save()
    torch.testing.assert_close(mapped[triples[:, 1]], rows_a)
