"""LocalShardsWrapper: container semantics, the tensor-protocol ops state-dict plumbing applies, checkpoint (DCP) hooks."""
import pickle

import pytest
import torch

from torchrec_b200.parallel.shards_wrapper import LocalShardsWrapper


def _cw():
    a, b = torch.arange(12.0).view(3, 4), torch.arange(100.0, 112.0).view(3, 4)
    return LocalShardsWrapper([a, b], [(0, 0), (0, 8)]), a, b


def test_container_and_bounding_shape():
    w, a, b = _cw()
    assert tuple(w.shape) == (3, 8) and w.local_offsets() == [(0, 0), (0, 8)] and w.local_sizes() == [(3, 4), (3, 4)]
    rows = LocalShardsWrapper([torch.zeros(2, 4), torch.zeros(5, 4)], [(0, 0), (7, 0)])
    assert tuple(rows.shape) == (7, 4)
    assert tuple(LocalShardsWrapper([], []).shape) == (0,)
    full = w.full_tensor((3, 12))
    assert torch.equal(full[:, :4], a) and torch.equal(full[:, 8:], b) and float(full[:, 4:8].abs().sum()) == 0.0


def test_protocol_ops():
    w, a, b = _cw()
    d = w.detach().clone()
    assert isinstance(d, LocalShardsWrapper) and d.local_shards()[0] is not a and torch.equal(d, w)
    assert isinstance(w.view_as(w), LocalShardsWrapper)
    with pytest.raises(NotImplementedError):
        w.view(24)
    with pytest.raises(NotImplementedError):
        w + 1
    d.zero_()
    assert float(d.local_shards()[1].sum()) == 0.0 and not torch.equal(d, w)
    d.copy_(w)
    assert torch.equal(d, w)
    full = torch.arange(36.0).view(3, 12)
    d.copy_(full)  # a full tensor: every shard takes its window
    assert torch.equal(d.local_shards()[1], full[:, 8:12])
    h = w.to(torch.float16)
    assert h.dtype == torch.float16 and h.local_shards()[0].dtype == torch.float16
    z = torch.zeros_like(w)
    assert isinstance(z, LocalShardsWrapper) and float(z.local_shards()[0].sum()) == 0.0
    w2 = pickle.loads(pickle.dumps(w))
    assert torch.equal(w2, w)


def test_checkpoint_hooks():
    from torch.distributed.checkpoint.metadata import MetadataIndex

    w, a, b = _cw()
    chunks = w.__create_chunk_list__()
    assert [tuple(c.offsets) for c in chunks] == [(0, 0), (0, 8)] and [tuple(c.sizes) for c in chunks] == [(3, 4), (3, 4)]
    items = w.__create_write_items__("tables.t.weight", torch.empty(3, 16, device="meta"))
    assert [tuple(i.index.offset) for i in items] == [(0, 0), (0, 8)] and all(tuple(i.tensor_data.size) == (3, 16) for i in items)
    assert w.__get_tensor_shard__(MetadataIndex("tables.t.weight", torch.Size((0, 8)))) is b
    assert w.__get_tensor_shard__(MetadataIndex("tables.t.weight", torch.Size((0, 0)), index=0)) is a
    with pytest.raises(ValueError):
        w.__get_tensor_shard__(MetadataIndex("tables.t.weight", torch.Size((1, 1))))
    names, meta = w.__tensor_flatten__()
    again = LocalShardsWrapper.__tensor_unflatten__({n: getattr(w, n) for n in names}, meta)
    assert torch.equal(again, w)
