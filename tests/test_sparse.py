"""JaggedTensor / KeyedJaggedTensor / KeyedTensor against a plain-python model (a dict key -> list of bags, a bag = list of
(value, weight)): every structural operation is checked on random inputs - fixed and variable (per key) batch sizes, weighted or
not, empty bags / empty keys. Methodology of the reference's sparse/tests/test_jagged_tensor.py, test_keyed_jagged_tensor.py and
test_keyed_tensor.py, model-based instead of hand-written cases."""
import random
from typing import Dict, List, Optional, Tuple

import pytest
import torch

from torchrec_b200.sparse.jagged_tensor import (ComputeJTDictToKJT, ComputeKJTToJTDict, JaggedTensor, KeyedJaggedTensor, KeyedTensor, flatten_kjt_list,
                                                kjt_is_equal, permute_multi_embedding, regroup_kts, unflatten_kjt_list)

Bag = List[Tuple[int, float]]
Model = Dict[str, List[Bag]]


def _random_model(rng: random.Random, variable: bool = False, n_keys: Optional[int] = None) -> Tuple[List[str], Model]:
    n_keys = rng.randint(1, 5) if n_keys is None else n_keys
    keys = [f"k{i}" for i in range(n_keys)]
    B = rng.randint(1, 5)
    model: Model = {}
    for k in keys:
        b = rng.randint(0, 5) if variable else B
        empty_key = rng.random() < 0.15
        model[k] = [[(rng.randrange(1000), round(rng.random() + 0.25, 3)) for _ in range(0 if empty_key or rng.random() < 0.25 else rng.randint(1, 4))] for _ in range(b)]
    return keys, model


def _build(keys: List[str], model: Model, weighted: bool, use_offsets: bool, variable: bool) -> KeyedJaggedTensor:
    vals = [v for k in keys for bag in model[k] for v, _ in bag]
    ws = [w for k in keys for bag in model[k] for _, w in bag]
    lens = [len(bag) for k in keys for bag in model[k]]
    lengths = torch.tensor(lens, dtype=torch.int64)
    kw = {}
    if use_offsets:
        kw["offsets"] = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(lengths, 0)])
    else:
        kw["lengths"] = lengths
    if variable:
        kw["stride_per_key_per_rank"] = [[len(model[k])] for k in keys]
    return KeyedJaggedTensor(keys=keys, values=torch.tensor(vals, dtype=torch.int64), weights=torch.tensor(ws, dtype=torch.float32) if weighted else None, **kw)


def _check(kjt: KeyedJaggedTensor, keys: List[str], model: Model, weighted: bool, what: str = "") -> None:
    """``kjt`` holds exactly ``model`` in key order ``keys`` (keys may repeat)."""
    assert kjt.keys() == keys, (what, kjt.keys(), keys)
    lens = [len(bag) for k in keys for bag in model[k]]
    assert kjt.lengths().tolist() == lens, (what, kjt.lengths().tolist(), lens)
    offs = [0]
    for x in lens:
        offs.append(offs[-1] + x)
    assert kjt.offsets().tolist() == offs, what
    assert kjt.values().tolist() == [v for k in keys for bag in model[k] for v, _ in bag], what
    if weighted:
        torch.testing.assert_close(kjt.weights(), torch.tensor([w for k in keys for bag in model[k] for _, w in bag], dtype=torch.float32), msg=what)
    else:
        assert kjt.weights_or_none() is None, what
    lpk = [sum(len(bag) for bag in model[k]) for k in keys]
    assert kjt.length_per_key() == lpk, (what, kjt.length_per_key(), lpk)
    opk = [0]
    for x in lpk:
        opk.append(opk[-1] + x)
    assert kjt.offset_per_key() == opk, what
    assert kjt.stride_per_key() == [len(model[k]) for k in keys], (what, kjt.stride_per_key())
    if keys:
        assert kjt.stride() == max(len(model[k]) for k in keys), what


CASES = [(w, o, v) for w in (False, True) for o in (False, True) for v in (False, True)]


@pytest.mark.parametrize("weighted,use_offsets,variable", CASES)
def test_kjt_accessors_to_dict_getitem(weighted, use_offsets, variable):
    for seed in range(25):
        rng = random.Random(seed)
        keys, model = _random_model(rng, variable)
        kjt = _build(keys, model, weighted, use_offsets, variable)
        _check(kjt, keys, model, weighted, f"seed {seed}")
        assert kjt.variable_stride_per_key() == variable
        d = kjt.to_dict()
        assert list(d.keys()) == keys
        for k in keys:
            for jt in (d[k], kjt[k]):
                assert jt.lengths().tolist() == [len(b) for b in model[k]]
                assert jt.values().tolist() == [v for b in model[k] for v, _ in b]
                assert jt.offsets().tolist()[0] == 0 and jt.offsets().tolist()[-1] == sum(len(b) for b in model[k])
                if weighted:
                    torch.testing.assert_close(jt.weights(), torch.tensor([w for b in model[k] for _, w in b], dtype=torch.float32))
                else:
                    assert jt.weights_or_none() is None
                dense = jt.to_dense()
                assert [t.tolist() for t in dense] == [[v for v, _ in b] for b in model[k]]
        # a dict of jagged tensors makes the same KJT again
        if not variable:
            back = KeyedJaggedTensor.from_jt_dict(d)
            _check(back, keys, model, weighted, f"from_jt_dict seed {seed}")
            back2 = ComputeJTDictToKJT()(ComputeKJTToJTDict()(kjt))
            _check(back2, keys, model, weighted, f"modules seed {seed}")


@pytest.mark.parametrize("weighted,use_offsets,variable", CASES)
def test_kjt_split_and_permute(weighted, use_offsets, variable):
    for seed in range(25):
        rng = random.Random(100 + seed)
        keys, model = _random_model(rng, variable)
        kjt = _build(keys, model, weighted, use_offsets, variable)
        # split: random segments (zero-length ones included) that sum to the number of keys
        segs, left = [], len(keys)
        while left:
            s = rng.randint(0, left)
            segs.append(s)
            left -= s
        if rng.random() < 0.3:
            segs.append(0)
        parts = kjt.split(segs)
        assert len(parts) == len(segs)
        at = 0
        for s, p in zip(segs, parts):
            _check(p, keys[at : at + s], model, weighted, f"split seed {seed} segs {segs}")
            at += s
        # permute: subsets, repeats, any order
        idx = [rng.randrange(len(keys)) for _ in range(rng.randint(0, len(keys) + 2))]
        perm = kjt.permute(idx)
        _check(perm, [keys[i] for i in idx], model, weighted, f"permute seed {seed} idx {idx}")
        perm2 = kjt.permute(idx, torch.tensor(idx, dtype=torch.int64))
        _check(perm2, [keys[i] for i in idx], model, weighted, f"permute(tensor) seed {seed} idx {idx}")
        # a permutation of a permutation
        if idx:
            idx2 = [rng.randrange(len(idx)) for _ in range(rng.randint(1, len(idx)))]
            _check(perm.permute(idx2), [keys[idx[i]] for i in idx2], model, weighted, f"permute^2 seed {seed}")
            # splits of a permuted KJT
            parts = perm.split([1, len(idx) - 1])
            _check(parts[0], [keys[idx[0]]], model, weighted, "split of permuted")
            _check(parts[1], [keys[i] for i in idx[1:]], model, weighted, "split of permuted")


@pytest.mark.parametrize("weighted,variable", [(False, False), (True, False), (False, True), (True, True)])
def test_kjt_concat(weighted, variable):
    for seed in range(20):
        rng = random.Random(200 + seed)
        keys, model = _random_model(rng, variable, n_keys=rng.randint(2, 6))
        cuts = sorted(rng.sample(range(1, len(keys)), k=min(len(keys) - 1, rng.randint(1, 2))))
        bounds = [0] + cuts + [len(keys)]
        pieces = [_build(keys[a:b], model, weighted, rng.random() < 0.5, variable) for a, b in zip(bounds[:-1], bounds[1:])]
        cat = KeyedJaggedTensor.concat(pieces)
        _check(cat, keys, model, weighted, f"concat seed {seed} cuts {cuts}")
        # split undoes concat
        for p, q in zip(cat.split([b - a for a, b in zip(bounds[:-1], bounds[1:])]), pieces):
            assert kjt_is_equal(p, q)


def test_kjt_sync_unsync_flatten_lengths_and_equality():
    rng = random.Random(7)
    keys, model = _random_model(rng, False, n_keys=3)
    kjt = _build(keys, model, True, False, False)
    assert kjt.length_per_key_or_none() is None and kjt.offset_per_key_or_none() is None
    assert kjt.sync() is kjt
    assert kjt.length_per_key_or_none() is not None and kjt.offset_per_key_or_none() is not None
    kjt.unsync()
    assert kjt.length_per_key_or_none() is None
    _check(kjt, keys, model, True)
    flat = kjt.flatten_lengths()
    assert flat.lengths().dim() == 1
    same = _build(keys, model, True, True, False)
    assert kjt_is_equal(kjt, same)
    other = KeyedJaggedTensor(keys=keys, values=kjt.values() + 1, lengths=kjt.lengths(), weights=kjt.weights())
    assert not kjt_is_equal(kjt, other)
    assert not kjt_is_equal(kjt, _build(keys, model, False, True, False))  # weights only on one side
    renamed = KeyedJaggedTensor(keys=[k + "x" for k in keys], values=kjt.values(), lengths=kjt.lengths(), weights=kjt.weights())
    assert not kjt_is_equal(kjt, renamed)
    with pytest.raises(KeyError):
        kjt["nope"]
    s = str(kjt)
    assert "KeyedJaggedTensor" in s and keys[0] in s


def test_kjt_from_sync_constructors_and_empty():
    keys = ["a", "b"]
    values = torch.arange(6)
    lengths = torch.tensor([1, 2, 0, 3])
    a = KeyedJaggedTensor.from_lengths_sync(keys, values, lengths)
    b = KeyedJaggedTensor.from_offsets_sync(keys, values, torch.tensor([0, 1, 3, 3, 6]))
    assert a.length_per_key_or_none() == [3, 3] and b.length_per_key_or_none() == [3, 3]
    assert kjt_is_equal(a, b)
    assert a.stride() == 2 and a.stride_per_key() == [2, 2]
    e = KeyedJaggedTensor.empty(is_weighted=True)
    assert e.keys() == [] and e.values().numel() == 0 and e.weights().numel() == 0 and e.lengths().numel() == 0
    el = KeyedJaggedTensor.empty_like(a)
    assert el.keys() == [] and el.values().dtype == a.values().dtype and el.weights_or_none() is None
    # stride given explicitly for a KJT without keys
    z = KeyedJaggedTensor(keys=[], values=torch.zeros(0, dtype=torch.long), lengths=torch.zeros(0, dtype=torch.long), stride=4)
    assert z.stride() == 4 and z.split([0, 0])[1].keys() == []


def test_kjt_variable_batch_inverse_indices_and_to():
    # two keys with batch sizes 3 and 1 (de-duplicated), inverse indices expand them to the full batch of 4
    keys = ["a", "b"]
    kjt = KeyedJaggedTensor(keys=keys, values=torch.tensor([1, 2, 3, 4, 5]), lengths=torch.tensor([2, 0, 1, 2]), stride_per_key_per_rank=[[3], [1]],
                            inverse_indices=(keys, torch.tensor([[0, 1, 2, 0], [0, 0, 0, 0]])))
    assert kjt.variable_stride_per_key() and kjt.stride_per_key() == [3, 1] and kjt.stride_per_key_per_rank() == [[3], [1]]
    assert kjt.length_per_key() == [3, 2] and kjt.lengths_offset_per_key() == [0, 3, 4]
    names, inv = kjt.inverse_indices()
    assert names == keys and inv.shape == (2, 4)
    assert kjt["b"].values().tolist() == [4, 5] and kjt["a"].lengths().tolist() == [2, 0, 1]
    p = kjt.permute([1, 0])
    assert p.stride_per_key() == [1, 3] and p.values().tolist() == [4, 5, 1, 2, 3] and p.lengths().tolist() == [2, 2, 0, 1]
    assert p.inverse_indices_or_none() is not None
    sp = kjt.split([1, 1])
    assert sp[0].stride_per_key() == [3] and sp[1].stride_per_key() == [1] and sp[1].values().tolist() == [4, 5]
    moved = kjt.to(torch.device("cpu"), dtype=torch.float64) if False else kjt.to(torch.device("cpu"))
    assert moved.stride_per_key() == [3, 1] and moved.values().tolist() == kjt.values().tolist()
    with pytest.raises(Exception):
        KeyedJaggedTensor(keys=keys, values=torch.tensor([1]), lengths=torch.tensor([1])).inverse_indices()
    # several ranks per key: the stride of a key is the sum over ranks
    k2 = KeyedJaggedTensor(keys=["a"], values=torch.arange(4), lengths=torch.tensor([1, 1, 2, 0, 0]), stride_per_key_per_rank=[[2, 3]])
    assert k2.stride_per_key() == [5] and k2.stride_per_key_per_rank() == [[2, 3]]


def test_kjt_dist_round_trip_over_virtual_ranks():
    """dist_labels / dist_splits / dist_tensors of W senders, exchanged by hand, rebuilt with dist_init: the receiver of key
    group g sees every sender's batch for those keys, sender-major inside each key."""
    for seed in range(12):
        rng = random.Random(300 + seed)
        W = rng.randint(2, 4)
        n_keys = rng.randint(W, W + 3)
        keys = [f"k{i}" for i in range(n_keys)]
        weighted = rng.random() < 0.5
        uneven = rng.random() < 0.5
        # key group of every destination rank
        cuts = sorted(rng.choices(range(0, n_keys + 1), k=W - 1))
        key_splits = [b - a for a, b in zip([0] + cuts, cuts + [n_keys])]
        models, kjts, Bs = [], [], []
        for r in range(W):
            B = rng.randint(1, 4) if uneven or r == 0 else Bs[0]
            Bs.append(B)
            model = {k: [[(rng.randrange(100), round(rng.random() + 0.5, 3)) for _ in range(rng.randint(0, 3))] for _ in range(B)] for k in keys}
            models.append(model)
            kjts.append(_build(keys, model, weighted, False, False))
        labels = kjts[0].dist_labels()
        assert labels[:2] == ["lengths", "values"] and (("weights" in labels) == weighted)
        splits = [k.dist_splits(key_splits) for k in kjts]  # per sender: per tensor: per destination
        tensors = [k.dist_tensors() for k in kjts]
        for dst in range(W):
            recv_tensors, recv_splits = [], []
            for ti in range(len(labels)):
                chunks, sizes = [], []
                for src in range(W):
                    sp = splits[src][ti]
                    lo = sum(sp[:dst])
                    chunks.append(tensors[src][ti][lo : lo + sp[dst]])
                    sizes.append(sp[dst])
                recv_tensors.append(torch.cat(chunks))
                recv_splits.append(sizes)
            my_keys = keys[sum(key_splits[:dst]) : sum(key_splits[: dst + 1])]
            got = KeyedJaggedTensor.dist_init(keys=my_keys, tensors=recv_tensors, variable_stride_per_key=False, num_workers=W, recat=None if not my_keys else _recat(len(my_keys), W, Bs),
                                              stride_per_rank=Bs, stagger=1)
            merged = {k: [bag for src in range(W) for bag in models[src][k]] for k in my_keys}
            assert got.keys() == my_keys
            if my_keys:
                assert got.lengths().tolist() == [len(b) for k in my_keys for b in merged[k]], f"seed {seed} dst {dst} Bs {Bs}"
                assert got.values().tolist() == [v for k in my_keys for b in merged[k] for v, _ in b], f"seed {seed} dst {dst} Bs {Bs}"
                if weighted:
                    torch.testing.assert_close(got.weights(), torch.tensor([w for k in my_keys for b in merged[k] for _, w in b], dtype=torch.float32))
                assert got.stride() == sum(Bs)


def _recat(n_local_keys: int, W: int, Bs: List[int]) -> torch.Tensor:
    """Received order is sender-major (sender, key); the KJT wants key-major (key, sender). Even batches permute rows of the
    [W * F, B] length matrix, uneven ones permute variable segments - same index list."""
    return torch.tensor([s * n_local_keys + k for k in range(n_local_keys) for s in range(W)], dtype=torch.int32)


def test_kjt_pytree_and_list_flatten():
    import torch.utils._pytree as pytree

    rng = random.Random(11)
    keys, model = _random_model(rng, False, n_keys=3)
    kjt = _build(keys, model, True, False, False)
    leaves, spec = pytree.tree_flatten(kjt)
    assert all(isinstance(x, torch.Tensor) or x is None for x in leaves)
    back = pytree.tree_unflatten(leaves, spec)
    _check(back, keys, model, True, "pytree")
    nested = {"x": [kjt, torch.ones(2)], "y": kjt["k0"]}
    leaves, spec = pytree.tree_flatten(nested)
    again = pytree.tree_unflatten(leaves, spec)
    assert kjt_is_equal(again["x"][0], kjt) and again["y"].values().tolist() == kjt["k0"].values().tolist()
    other = _build(keys[:2], model, False, True, False)
    flat, ctx = flatten_kjt_list([kjt, other])
    out = unflatten_kjt_list(flat, ctx)
    assert kjt_is_equal(out[0], kjt) and kjt_is_equal(out[1], other)
    kt = KeyedTensor(keys=["a", "b"], length_per_key=[2, 3], values=torch.randn(4, 5))
    leaves, spec = pytree.tree_flatten(kt)
    kt2 = pytree.tree_unflatten(leaves, spec)
    assert kt2.keys() == ["a", "b"] and kt2.length_per_key() == [2, 3] and torch.equal(kt2.values(), kt.values())


# ---- JaggedTensor ---------------------------------------------------------------------------------------------------------------------------------
def test_jagged_tensor_dense_round_trips():
    for seed in range(20):
        rng = random.Random(400 + seed)
        bags = [[float(rng.randrange(100)) for _ in range(rng.randint(0, 5))] for _ in range(rng.randint(1, 6))]
        wts = [[rng.random() for _ in b] for b in bags]
        jt = JaggedTensor.from_dense([torch.tensor(b) for b in bags], weights=[torch.tensor(w) for w in wts])
        assert jt.lengths().tolist() == [len(b) for b in bags]
        assert [t.tolist() for t in jt.to_dense()] == bags
        for a, b in zip(jt.to_dense_weights(), wts):
            torch.testing.assert_close(a, torch.tensor(b))
        L = max(len(b) for b in bags)
        for want in (None, L + 2, max(L - 1, 1)):
            width = L if want is None else want
            pad = jt.to_padded_dense(want, padding_value=-1.0)
            assert pad.shape == (len(bags), width)
            for i, b in enumerate(bags):
                row = (b + [-1.0] * width)[:width]
                assert pad[i].tolist() == row, (seed, want)
            padw = jt.to_padded_dense_weights(want, padding_value=0.5)
            for i, w in enumerate(wts):
                torch.testing.assert_close(padw[i], torch.tensor((w + [0.5] * width)[:width], dtype=padw.dtype))
        # from_dense_lengths: a padded matrix + lengths
        if L > 0:
            dense = torch.tensor([(b + [0.0] * L)[:L] for b in bags])
            jt2 = JaggedTensor.from_dense_lengths(dense, torch.tensor([len(b) for b in bags]))
            assert jt2.values().tolist() == [x for b in bags for x in b]
            assert jt2.lengths().tolist() == [len(b) for b in bags]
    e = JaggedTensor.empty(is_weighted=True)
    assert e.values().numel() == 0 and e.weights().numel() == 0 and e.lengths().numel() == 0 and e.offsets().numel() in (0, 1)
    jt = JaggedTensor(values=torch.arange(3), offsets=torch.tensor([0, 1, 3]))
    assert jt.lengths().tolist() == [1, 2] and jt.lengths_or_none() is not None and jt.weights_or_none() is None
    with pytest.raises(Exception):
        jt.weights()
    assert "JaggedTensor" in str(jt)
    assert jt.to(torch.device("cpu")).values().tolist() == [0, 1, 2]


def test_jagged_tensor_2d_values():
    """Sequence embeddings: values [N, D]."""
    vals = torch.arange(12.0).view(6, 2)
    jt = JaggedTensor(values=vals, lengths=torch.tensor([2, 0, 3, 1]))
    dense = jt.to_dense()
    assert [d.shape[0] for d in dense] == [2, 0, 3, 1] and torch.equal(dense[2], vals[2:5])
    pad = jt.to_padded_dense(3)
    assert pad.shape == (4, 3, 2) and torch.equal(pad[0, :2], vals[:2]) and float(pad[1].abs().sum()) == 0 and torch.equal(pad[2], vals[2:5])


# ---- KeyedTensor ----------------------------------------------------------------------------------------------------------------------------------
def test_keyed_tensor_accessors_and_regroup_random():
    for seed in range(25):
        rng = random.Random(500 + seed)
        B = rng.randint(1, 4)
        kts, store = [], {}
        kid = 0
        for _ in range(rng.randint(1, 3)):
            names, dims = [], []
            for _ in range(rng.randint(1, 4)):
                names.append(f"e{kid}")
                dims.append(rng.choice([1, 2, 4, 8]))
                kid += 1
            vals = torch.randn(B, sum(dims))
            kt = KeyedTensor(keys=names, length_per_key=dims, values=vals)
            kts.append(kt)
            off = 0
            for n, d in zip(names, dims):
                store[n] = vals[:, off : off + d]
                off += d
            assert kt.keys() == names and kt.length_per_key() == dims and kt.key_dim() == 1
            assert kt.offset_per_key() == [sum(dims[:i]) for i in range(len(dims) + 1)]
            for n in names:
                assert torch.equal(kt[n], store[n])
            assert list(kt.to_dict().keys()) == names
        allk = list(store)
        groups = [[rng.choice(allk) for _ in range(rng.randint(1, 4))] for _ in range(rng.randint(1, 3))]  # duplicates inside / across groups allowed
        want = [torch.cat([store[k] for k in g], dim=1) for g in groups]
        for fn in (KeyedTensor.regroup, regroup_kts, permute_multi_embedding):
            got = fn(kts, groups)
            assert len(got) == len(want)
            for a, b in zip(got, want):
                assert torch.equal(a, b), (seed, fn)
        d = KeyedTensor.regroup_as_dict(kts, groups, [f"g{i}" for i in range(len(groups))])
        assert list(d.keys()) == [f"g{i}" for i in range(len(groups))] and all(torch.equal(d[f"g{i}"], want[i]) for i in range(len(groups)))


def test_keyed_tensor_regroup_backward_and_from_tensor_list():
    a = torch.randn(3, 6, requires_grad=True)
    b = torch.randn(3, 4, requires_grad=True)
    kts = [KeyedTensor(keys=["x", "y"], length_per_key=[2, 4], values=a), KeyedTensor(keys=["z"], length_per_key=[4], values=b)]
    out = KeyedTensor.regroup(kts, [["z", "x"], ["y", "x"]])
    (out[0].sum() * 2 + out[1].sum() * 3).backward()
    ga = torch.cat([torch.full((3, 2), 5.0), torch.full((3, 4), 3.0)], dim=1)  # x is in both groups
    torch.testing.assert_close(a.grad, ga)
    torch.testing.assert_close(b.grad, torch.full((3, 4), 2.0))
    kt = KeyedTensor.from_tensor_list(["p", "q"], [torch.ones(2, 3), torch.zeros(2, 1)])
    assert kt.length_per_key() == [3, 1] and kt.values().shape == (2, 4) and float(kt["q"].sum()) == 0
    kt0 = KeyedTensor.from_tensor_list(["p", "q"], [torch.ones(3, 2), torch.zeros(1, 2)], key_dim=0, cat_dim=0)
    assert kt0.key_dim() == 0 and kt0["p"].shape == (3, 2) and kt0["q"].shape == (1, 2)
    assert "KeyedTensor" in str(kt)
    assert kt.to(torch.device("cpu")).keys() == ["p", "q"]


def test_construct_jagged_tensors_column_blocks_dedup_and_inference_form():
    """Sequence lookups come back as one [sum L, D] tensor: per-feature JaggedTensors, column-wise sharded tables as several blocks that
    are concatenated (in permuted order), de-duplicated ids expanded by the reverse index, ids in the weights slot on request."""
    from torchrec_b200.modules.utils import (construct_jagged_tensors, construct_jagged_tensors_inference, construct_modulelist_from_single_module,
                                             init_mlp_weights_xavier_uniform, reset_module_states_post_sharding)

    # features as the lookup saw them: f0 once, f1 twice (two column blocks of one table)
    lengths = torch.tensor([2, 0, 1, 1, 2, 0, 1, 2, 0])
    values = torch.tensor([10, 11, 12, 20, 21, 21, 20, 21, 21])
    feats = KeyedJaggedTensor(keys=["f0", "f1", "f1"], values=values, lengths=lengths)
    emb = torch.arange(9 * 2, dtype=torch.float32).view(9, 2)
    out = construct_jagged_tensors(emb, feats, ["f0", "f1", "f1"], need_indices=True, features_to_permute_indices={"f1": [1, 0]})
    assert set(out) == {"f0", "f1"}
    assert out["f0"].lengths().tolist() == [2, 0, 1] and torch.equal(out["f0"].values(), emb[:3]) and out["f0"].weights().tolist() == [10, 11, 12]
    assert out["f1"].values().shape == (3, 4) and torch.equal(out["f1"].values(), torch.cat([emb[6:9], emb[3:6]], dim=1))  # block order swapped
    assert out["f1"].lengths().tolist() == [1, 2, 0]
    plain = construct_jagged_tensors(emb, feats, ["f0", "f1", "f1"])
    assert plain["f0"].weights_or_none() is None and torch.equal(plain["f1"].values(), torch.cat([emb[3:6], emb[6:9]], dim=1))
    # de-duplicated lookup: 3 unique rows, reverse index per original id
    uniq = torch.tensor([[1.0, 1.0], [2.0, 2.0], [3.0, 3.0]])
    rev = torch.tensor([2, 0, 0, 1])
    orig = KeyedJaggedTensor(keys=["a", "b"], values=torch.tensor([7, 5, 5, 6]), lengths=torch.tensor([1, 1, 2, 0]))
    for gather in (False, True):
        got = construct_jagged_tensors(uniq, orig, ["a", "b"], original_features=orig, reverse_indices=rev, use_gather_select=gather)
        assert got["a"].values().tolist() == [[3.0, 3.0], [1.0, 1.0]] and got["b"].values().tolist() == [[1.0, 1.0], [2.0, 2.0]] and got["b"].lengths().tolist() == [2, 0]
    # inference form: plain tensors, padding rows cut
    padded = torch.cat([emb, torch.zeros(3, 2)])
    inf = construct_jagged_tensors_inference(padded, lengths.view(3, 3), values, ["f0", "f1", "f1"], need_indices=True, remove_padding=True)
    assert torch.equal(inf["f0"].values(), emb[:3]) and torch.equal(inf["f1"].values(), torch.cat([emb[3:6], emb[6:9]], dim=1)) and inf["f1"].weights().tolist() == [20, 21, 21]
    # module helpers
    mlp = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.ReLU())
    grid = construct_modulelist_from_single_module(mlp, (2, 3))
    assert len(grid) == 2 and len(grid[0]) == 3 and grid[0][0][0].weight.data_ptr() != grid[1][2][0].weight.data_ptr()
    assert float(grid[1][1][0].bias.abs().sum()) == 0.0 and not torch.equal(grid[0][0][0].weight, grid[0][1][0].weight)
    init_mlp_weights_xavier_uniform(torch.nn.ReLU())  # no-op on other layers

    from torchrec_b200.modules.regroup import KTRegroupAsDict

    re = KTRegroupAsDict([["x"], ["y"]], ["g0", "g1"])
    kt = KeyedTensor(keys=["x", "y"], length_per_key=[2, 1], values=torch.randn(2, 3))
    re([kt])
    assert re._is_inited
    reset_module_states_post_sharding(torch.nn.Sequential(re))
    assert not re._is_inited
    assert torch.equal(re([kt])["g1"], kt["y"])
