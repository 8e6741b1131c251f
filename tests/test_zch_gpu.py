"""MPZCH probe kernel (ops/csrc/zch.cu) against the PyTorch mirror and against the properties a zero-collision hash must have.
Reference behaviour: torchrec/modules/hash_mc_modules.py:196-560 (fbgemm zero_collision_hash)."""
import pytest
import torch

from torchrec_b200.modules.hash_mc_modules import HashZchEvictionConfig, HashZchEvictionPolicyName, HashZchManagedCollisionModule
from torchrec_b200.sparse.jagged_tensor import JaggedTensor

pytestmark = pytest.mark.gpu


def _mod(device, zch_size=4096, buckets=4, max_probe=64, **kw):
    return HashZchManagedCollisionModule(zch_size=zch_size, device=torch.device(device), total_num_buckets=buckets, max_probe=max_probe, **kw)


def _remap(m, ids):
    jt = JaggedTensor(values=ids, lengths=torch.ones(ids.numel(), dtype=torch.int64, device=ids.device))
    return m.remap({"f": jt})["f"].values()


def test_insert_then_hit_unique_slots_and_cpu_parity():
    dev = "cuda"
    m = _mod(dev, tb_logging_frequency=1000)
    m.train()
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 1 << 50, (1500,), generator=g).unique().to(dev)
    dup = torch.cat([ids, ids[:200]])  # duplicates inside one batch share the slot of their first copy
    slots = _remap(m, dup)
    n = ids.numel()
    assert slots.min() >= 0 and slots.max() < 4096
    assert slots[:n].unique().numel() == n, "two different ids share a slot although the table is 1/3 full"
    assert torch.equal(slots[n:], slots[:200])
    ident = m._hash_zch_identities.view(-1)
    assert torch.equal(ident[slots[:n]], ids)
    assert torch.equal(ident[slots], dup)
    assert int((ident != -1).sum()) == n
    # second pass (training) and eval pass: pure hits, same slots, nothing inserted
    again = _remap(m, ids)
    m.eval()
    ro = _remap(m, ids)
    assert torch.equal(again, slots[:n]) and torch.equal(ro, slots[:n])
    stats = m.flush_statistics()
    assert stats["insert"] == n and stats["hit"] == 200 + 2 * n and stats["collision"] == 0
    # unseen ids in eval mode: no insertion, they fall back to their start slot (a collision)
    unseen = torch.randint(1 << 51, 1 << 52, (64,), generator=g).to(dev)
    fb = _remap(m, unseen)
    assert int((m._hash_zch_identities.view(-1) != -1).sum()) == n and fb.min() >= 0
    # the PyTorch mirror on CPU finds every id of the GPU-built table at the same slot (same hash, same layout)
    c = _mod("cpu")
    c.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()}, strict=False)
    c.eval()
    assert torch.equal(_remap(c, ids.cpu()), slots[:n].cpu())


def test_full_window_collides_and_fallback_switch():
    dev = "cuda"
    m = _mod(dev, zch_size=64, buckets=1, max_probe=64)
    m.train()
    first = torch.arange(1000, 1064, device=dev)
    s1 = _remap(m, first)
    assert s1.unique().numel() == 64  # the table is exactly full, still collision free
    more = torch.arange(5000, 5016, device=dev)
    s2 = _remap(m, more)
    assert s2.min() >= 0 and s2.max() < 64  # no room, no eviction policy: fall back to the start slot
    assert torch.equal(m._hash_zch_identities.view(-1)[s1], first), "a colliding id must not overwrite an owner"
    m2 = _mod(dev, zch_size=64, buckets=1, max_probe=64, disable_fallback=True)
    m2.train()
    _remap(m2, first)
    assert bool((_remap(m2, more) - m2._offset == -1).all())


def test_ttl_eviction_reuses_expired_slots_and_reports_them():
    dev = "cuda"
    m = _mod(dev, zch_size=128, buckets=1, max_probe=128, eviction_policy_name=HashZchEvictionPolicyName.SINGLE_TTL_EVICTION,
             eviction_config=HashZchEvictionConfig(features=["f"], single_ttl=24))
    m.train()
    old = torch.arange(0, 128, device=dev) + 10_000
    s_old = _remap(m, old)
    assert s_old.unique().numel() == 128 and m.evict() is None
    m._hash_zch_metadata -= 48  # two days ago: everything is expired
    new = torch.arange(0, 40, device=dev) + 90_000
    s_new = _remap(m, new)
    assert s_new.unique().numel() == 40
    ev = m.evict()
    assert ev is not None and torch.equal(ev.sort().values, s_new.sort().values), "evicted slots are exactly the slots the new ids took"
    ident = m._hash_zch_identities.view(-1)
    assert torch.equal(ident[s_new], new)
    # the survivors are still found where they were
    keep = torch.ones(128, dtype=torch.bool, device=dev)
    keep[(s_old.unsqueeze(1) == s_new.unsqueeze(0)).any(1)] = False
    assert torch.equal(_remap(m, old[keep]), s_old[keep])


def test_lru_eviction_takes_the_least_recently_seen_slot():
    dev = "cuda"
    m = _mod(dev, zch_size=32, buckets=1, max_probe=32, eviction_policy_name=HashZchEvictionPolicyName.LRU_EVICTION)
    m.train()
    ids = torch.arange(0, 32, device=dev) + 777
    s = _remap(m, ids)
    meta = m._hash_zch_metadata.view(-1)
    meta -= 5
    meta[s[7]] -= 100  # id 7 is by far the oldest
    newcomer = torch.tensor([123456789], device=dev)
    sn = _remap(m, newcomer)
    assert int(sn) == int(s[7]) and int(m.evict()) == int(s[7])
