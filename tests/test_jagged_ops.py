"""The jagged / sparse op surface (``ops/jagged.py``, CPU implementations - the fallbacks of the CUDA kernels and the references the
GPU numerics tests compare against) checked against plain python loops on random inputs."""
import random
from typing import List

import pytest
import torch

from torchrec_b200.ops import jagged as J


def _rand_lengths(rng: random.Random, n: int, hi: int = 4) -> List[int]:
    return [0 if rng.random() < 0.25 else rng.randint(1, hi) for _ in range(n)]


def test_cumsums_offsets_range_invert_permute():
    x = torch.tensor([3, 0, 2, 5])
    assert J.asynchronous_complete_cumsum(x).tolist() == [0, 3, 3, 5, 10]
    assert J.asynchronous_inclusive_cumsum(x).tolist() == [3, 3, 5, 10]
    assert J.asynchronous_exclusive_cumsum(x).tolist() == [0, 3, 3, 5]
    assert J.asynchronous_complete_cumsum(torch.tensor([[1, 2], [3, 4]])).tolist() == [[0, 1, 3], [0, 3, 7]]
    assert J.asynchronous_complete_cumsum(torch.zeros(0, dtype=torch.long)).tolist() == [0]
    assert J.offsets_range(torch.tensor([0, 3, 3, 5]), 10).tolist() == [0, 1, 2, 0, 1, 0, 1, 2, 3, 4]
    assert J.offsets_range(torch.tensor([0, 0]), 0).numel() == 0
    for seed in range(10):
        p = torch.randperm(seed + 1, generator=torch.Generator().manual_seed(seed))
        inv = J.invert_permute(p)
        assert torch.equal(p[inv], torch.arange(seed + 1)) and torch.equal(inv[p], torch.arange(seed + 1))


def test_segment_sum_csr_random():
    for seed in range(15):
        rng = random.Random(seed)
        bs = rng.randint(1, 3)
        segs = _rand_lengths(rng, rng.randint(1, 6), 3)
        csr = [0]
        for s in segs:
            csr.append(csr[-1] + s)
        vals = [rng.random() for _ in range(csr[-1] * bs)]
        got = J.segment_sum_csr(bs, torch.tensor(csr), torch.tensor(vals, dtype=torch.float32))
        want = [sum(vals[csr[i] * bs : csr[i + 1] * bs]) for i in range(len(segs))]
        torch.testing.assert_close(got, torch.tensor(want, dtype=torch.float32))


@pytest.mark.parametrize("weighted", [False, True])
def test_permute_2d_and_1d_random(weighted):
    for seed in range(25):
        rng = random.Random(50 + seed)
        F, B = rng.randint(1, 5), rng.randint(1, 4)
        lens = [_rand_lengths(rng, B) for _ in range(F)]
        bags = [[[rng.randrange(100) for _ in range(lens[f][b])] for b in range(B)] for f in range(F)]
        values = torch.tensor([v for f in range(F) for b in range(B) for v in bags[f][b]], dtype=torch.int64)
        weights = values.float() * 0.5 if weighted else None
        perm = [rng.randrange(F) for _ in range(rng.randint(0, F + 2))]
        want_vals = [v for f in perm for b in range(B) for v in bags[f][b]]
        for total in (None, len(want_vals)):
            pl, pv, pw = J.permute_2D_sparse_data(torch.tensor(perm, dtype=torch.int32), torch.tensor(lens).view(F, B), values, weights, total)
            assert pl.tolist() == [lens[f] for f in perm]
            assert pv.tolist() == want_vals
            assert (pw is None) == (not weighted)
            if weighted:
                assert pw.tolist() == [0.5 * v for v in want_vals]
        fl, fv, _ = J.permute_2D_sparse_data_input1D(torch.tensor(perm, dtype=torch.int32), torch.tensor(lens).view(-1), values, B, weights, None)
        assert fl.tolist() == [x for f in perm for x in lens[f]] and fv.tolist() == want_vals
        # 1D: segments of different sizes
        seg = _rand_lengths(rng, rng.randint(1, 6), 5)
        data = [[rng.randrange(100) for _ in range(s)] for s in seg]
        p1 = [rng.randrange(len(seg)) for _ in range(rng.randint(0, len(seg) + 2))]
        v1 = torch.tensor([x for d in data for x in d], dtype=torch.int64)
        ol, ov, ow = J.permute_1D_sparse_data(torch.tensor(p1, dtype=torch.int32), torch.tensor(seg), v1, v1.float() if weighted else None, None)
        assert ol.tolist() == [seg[i] for i in p1] and ov.tolist() == [x for i in p1 for x in data[i]]
        if weighted:
            assert ow.tolist() == [float(x) for i in p1 for x in data[i]]
        # 2-D values (rows travel whole)
        v2 = torch.stack([v1, v1 + 1000], dim=1)
        _, ov2, _ = J.permute_1D_sparse_data(torch.tensor(p1, dtype=torch.int32), torch.tensor(seg), v2, None, None)
        assert ov2.shape == (len(ov), 2) and ov2[:, 0].tolist() == ov.tolist()


def test_expand_into_jagged_permute_random():
    for seed in range(15):
        rng = random.Random(80 + seed)
        seg = _rand_lengths(rng, rng.randint(1, 6), 4)
        n = len(seg)
        perm = list(range(n))
        rng.shuffle(perm)
        in_off = [0]
        for s in seg:
            in_off.append(in_off[-1] + s)
        out_len = [seg[p] for p in perm]
        out_off = [0]
        for s in out_len:
            out_off.append(out_off[-1] + s)
        got = J.expand_into_jagged_permute(torch.tensor(perm), torch.tensor(in_off), torch.tensor(out_off), out_off[-1])
        want = [in_off[p] + j for p in perm for j in range(seg[p])]
        assert got.tolist() == want


def _naive_bucketize(lens, bags, F, B, W, block, pos_tables, keep_orig):
    """-> (new_lengths [W*F*B], new ids, source order) in (bucket, feature, sample) order, stable inside a bag."""
    out = [[[] for _ in range(F * B)] for _ in range(W)]
    flat_pos = 0
    for f in range(F):
        for b in range(B):
            for j, v in enumerate(bags[f][b]):
                if pos_tables is not None:
                    t = pos_tables[f]
                    bucket = max(i for i in range(W) if t[i] <= v) if v >= t[0] else 0
                    bucket = min(bucket, W - 1)
                    local = v - t[bucket]
                else:
                    blk = v // block[f]
                    if blk < W:
                        bucket, local = blk, v - blk * block[f]
                    else:
                        bucket, local = v % W, v // W
                out[bucket][f * B + b].append((v if keep_orig else local, flat_pos, j))
                flat_pos += 1
    new_lengths = [len(out[w][i]) for w in range(W) for i in range(F * B)]
    items = [x for w in range(W) for i in range(F * B) for x in out[w][i]]
    return new_lengths, [x[0] for x in items], [x[1] for x in items], [x[2] for x in items]


@pytest.mark.parametrize("mode", ["uniform", "pos_tables", "keep_orig"])
def test_block_bucketize_sparse_features_random(mode):
    for seed in range(25):
        rng = random.Random(120 + seed)
        F, B, W = rng.randint(1, 4), rng.randint(1, 4), rng.randint(1, 4)
        rows = [rng.randint(W, 60) for _ in range(F)]
        block = [(r + W - 1) // W for r in rows]
        lens = [_rand_lengths(rng, B) for _ in range(F)]
        overflow = mode == "uniform" and rng.random() < 0.3  # ids past the last block wrap round-robin
        bags = [[[rng.randrange(rows[f] * (3 if overflow else 1)) for _ in range(lens[f][b])] for b in range(B)] for f in range(F)]
        pos_tables = None
        if mode == "pos_tables":  # uneven row blocks: block w of feature f covers [t[w], t[w+1])
            pos_tables = []
            for f in range(F):
                cuts = sorted(rng.sample(range(1, rows[f]), W - 1)) if W > 1 and rows[f] > W else list(range(1, W))
                pos_tables.append([0] + cuts + [rows[f]])
        indices = torch.tensor([v for f in range(F) for b in range(B) for v in bags[f][b]], dtype=torch.int64)
        weights = indices.float() + 0.5
        nl, ni, nw, npos, unb = J.block_bucketize_sparse_features(
            torch.tensor([x for f in range(F) for x in lens[f]], dtype=torch.int32), indices, bucketize_pos=True, sequence=True, block_sizes=torch.tensor(block), my_size=W,
            weights=weights, block_bucketize_pos=None if pos_tables is None else [torch.tensor(t) for t in pos_tables], keep_orig_idx=(mode == "keep_orig"))
        want_len, want_ids, src, want_pos = _naive_bucketize(lens, bags, F, B, W, block, pos_tables, mode == "keep_orig")
        ctx = f"seed {seed} F {F} B {B} W {W}"
        assert nl.tolist() == want_len, ctx
        assert ni.tolist() == want_ids, ctx
        assert nw.tolist() == [float(indices[s]) + 0.5 for s in src], ctx
        assert npos.tolist() == want_pos, ctx
        # unbucketize permute: where every original element went
        assert [src[int(u)] for u in unb.tolist()] == list(range(len(src))), ctx
        # without the optional outputs
        nl2, ni2, nw2, npos2, unb2 = J.block_bucketize_sparse_features(torch.tensor([x for f in range(F) for x in lens[f]], dtype=torch.int32), indices, False, False,
                                                                       torch.tensor(block), W, block_bucketize_pos=None if pos_tables is None else [torch.tensor(t) for t in pos_tables],
                                                                       keep_orig_idx=(mode == "keep_orig"))
        assert nl2.tolist() == want_len and ni2.tolist() == want_ids and nw2 is None and npos2 is None and unb2 is None


def test_jagged_dense_conversions_random():
    for seed in range(20):
        rng = random.Random(160 + seed)
        N = rng.randint(1, 6)
        lens = _rand_lengths(rng, N, 5)
        off = [0]
        for x in lens:
            off.append(off[-1] + x)
        D = rng.choice([None, 1, 3])
        vals = torch.randn(off[-1]) if D is None else torch.randn(off[-1], D)
        max_len = rng.randint(1, 6)
        pad = J.jagged_to_padded_dense(vals, [torch.tensor(off)], [max_len], -2.0)
        assert pad.shape[:2] == (N, max_len)
        for i in range(N):
            for j in range(max_len):
                if j < lens[i]:
                    assert torch.equal(pad[i, j], vals[off[i] + j])
                else:
                    assert bool((pad[i, j] == -2.0).all())
        if D is None:
            assert torch.equal(J.jagged_1d_to_dense(vals, torch.tensor(off), max_len, -2.0), pad)
        else:
            assert torch.equal(J.jagged_2d_to_dense(vals, torch.tensor(off), max_len), J.jagged_to_padded_dense(vals, [torch.tensor(off)], [max_len], 0.0))
        # dense -> jagged of a wide enough matrix is the inverse
        wide = J.jagged_to_padded_dense(vals, [torch.tensor(off)], [max(max(lens), 1)], 0.0)
        back, offs = J.dense_to_jagged(wide, [torch.tensor(off)])
        assert torch.equal(back, vals) and offs[0].tolist() == off
        back2, _ = J.dense_to_jagged(wide, [torch.tensor(off)], total_L=off[-1])
        assert torch.equal(back2, vals)


def test_index_selects_random():
    for seed in range(20):
        rng = random.Random(200 + seed)
        # jagged_index_select_2d: pick bags
        N = rng.randint(1, 6)
        lens = _rand_lengths(rng, N, 4)
        rows = [[[rng.random(), rng.random()] for _ in range(n)] for n in lens]
        vals = torch.tensor([r for bag in rows for r in bag], dtype=torch.float32).view(-1, 2)
        idx = [rng.randrange(N) for _ in range(rng.randint(0, N + 2))]
        out, out_len = J.jagged_index_select_2d(vals, torch.tensor(lens), torch.tensor(idx, dtype=torch.int64))
        assert out_len.tolist() == [lens[i] for i in idx]
        want = [r for i in idx for r in rows[i]]
        torch.testing.assert_close(out, torch.tensor(want, dtype=torch.float32).view(-1, 2))
        # keyed_jagged_index_select_dim1: the same batch positions of every key
        F, B = rng.randint(1, 4), rng.randint(1, 5)
        kl = [_rand_lengths(rng, B) for _ in range(F)]
        kb = [[[rng.randrange(100) for _ in range(kl[f][b])] for b in range(B)] for f in range(F)]
        kv = torch.tensor([v for f in range(F) for b in range(B) for v in kb[f][b]], dtype=torch.int64)
        flat_l = torch.tensor([x for f in range(F) for x in kl[f]])
        offs = J.asynchronous_complete_cumsum(flat_l)
        sel = [rng.randrange(B) for _ in range(rng.randint(1, B + 1))]
        want_v = [v for f in range(F) for b in sel for v in kb[f][b]]
        for w, total in ((None, None), (kv.float(), len(want_v))):
            res = J.keyed_jagged_index_select_dim1(kv, flat_l, offs, torch.tensor(sel), B, w, total)
            assert res[0].tolist() == want_v and res[1].tolist() == [kl[f][b] for f in range(F) for b in sel]
            assert len(res) == (2 if w is None else 3)
            if w is not None:
                assert res[2].tolist() == [float(v) for v in want_v]
        # batch_index_select_dim0: row gathers from several tables stored flat
        shapes = [(rng.randint(1, 5), rng.randint(1, 3)) for _ in range(rng.randint(1, 3))]
        tabs = [torch.randn(r, c) for r, c in shapes]
        n_idx = [rng.randint(1, 4) for _ in shapes]
        ids = [[rng.randrange(r) for _ in range(n)] for (r, _), n in zip(shapes, n_idx)]
        got = J.batch_index_select_dim0(torch.cat([t.reshape(-1) for t in tabs]), torch.tensor([i for x in ids for i in x]), n_idx, [r for r, _ in shapes], [c for _, c in shapes])
        want_b = torch.cat([t[torch.tensor(i)].reshape(-1) for t, i in zip(tabs, ids)])
        assert torch.equal(got, want_b)
        # group_index_select_dim0 (mixed widths) incl. gradients
        ins = [torch.randn(rng.randint(2, 5), rng.choice([2, 2, 3]), requires_grad=True) for _ in range(rng.randint(1, 4))]
        gi = [torch.tensor([rng.randrange(x.shape[0]) for _ in range(rng.randint(1, 5))]) for x in ins]
        outs = J.group_index_select_dim0(ins, gi)
        for o, x, i in zip(outs, ins, gi):
            assert torch.equal(o, x.detach()[i])
        sum((o * (k + 1)).sum() for k, o in enumerate(outs)).backward()
        for k, (x, i) in enumerate(zip(ins, gi)):
            want_g = torch.zeros_like(x)
            want_g.index_add_(0, i, torch.full((i.numel(), x.shape[1]), float(k + 1)))
            torch.testing.assert_close(x.grad, want_g)


def test_jagged_unique_indices_random():
    """Per-table unique ids: features of one table share a hash-size range; reverse index maps every id to its unique slot."""
    for seed in range(20):
        rng = random.Random(240 + seed)
        T = rng.randint(1, 3)
        feats_per_table = [rng.randint(1, 2) for _ in range(T)]
        rows = [rng.randint(3, 12) for _ in range(T)]
        F, B = sum(feats_per_table), rng.randint(1, 4)
        table_of_feat = [t for t in range(T) for _ in range(feats_per_table[t])]
        # hash_size_cumsum is per feature (features of one table share the table's base); hash_size_offsets: first feature of every table
        base, tb = [], 0
        for t in range(T):
            base += [tb] * feats_per_table[t]
            tb += rows[t]
        hash_size_cumsum = torch.tensor(base + [tb])
        hso = [0]
        for n in feats_per_table:
            hso.append(hso[-1] + n)
        lens = [_rand_lengths(rng, B) for _ in range(F)]
        bags = [[[rng.randrange(rows[table_of_feat[f]]) for _ in range(lens[f][b])] for b in range(B)] for f in range(F)]
        ids = [v for f in range(F) for b in range(B) for v in bags[f][b]]
        feat_of_id = [f for f in range(F) for b in range(B) for _ in bags[f][b]]
        offsets = J.asynchronous_complete_cumsum(torch.tensor([x for f in range(F) for x in lens[f]]))
        if seed % 2:  # the reference's layout: one entry per feature, the feature count of a table on its first feature
            per_feat = [n if j == 0 else 0 for n in feats_per_table for j in range(n)]
            hso_arg = [0]
            for n in per_feat:
                hso_arg.append(hso_arg[-1] + n)
        else:
            hso_arg = hso
        out_len, out_off, uniq, inv = J.jagged_unique_indices(hash_size_cumsum, torch.tensor(hso_arg), offsets, torch.tensor(ids, dtype=torch.int64))
        # every id is recovered through the reverse index, unique ids are unique per table
        want_per_table = [sorted({v for v, f in zip(ids, feat_of_id) if table_of_feat[f] == t}) for t in range(T)]
        assert uniq.tolist() == [v for t in range(T) for v in want_per_table[t]], seed
        assert [uniq.tolist()[i] for i in inv.tolist()] == ids, seed
        assert int(out_len.sum()) == uniq.numel() and out_off.tolist()[-1] == uniq.numel()
        # the unique ids of table t sit in the first bag of its first feature
        for t in range(T):
            assert int(out_len[hso[t] * B]) == len(want_per_table[t]), seed


def test_permute_pooled_embs_and_nbit_round_trip():
    for seed in range(10):
        rng = random.Random(280 + seed)
        dims = [rng.choice([1, 2, 4]) for _ in range(rng.randint(1, 5))]
        off = [0]
        for d in dims:
            off.append(off[-1] + d)
        x = torch.randn(3, off[-1], requires_grad=True)
        perm = list(range(len(dims)))
        rng.shuffle(perm)
        out = J.permute_pooled_embs(x, off, perm)
        want = torch.cat([x.detach()[:, off[p] : off[p + 1]] for p in perm], dim=1)
        assert torch.equal(out, want)
        (out * torch.arange(out.shape[1]).float()).sum().backward()
        col = 0
        for p in perm:
            for j in range(dims[p]):
                assert float(x.grad[0, off[p] + j]) == float(col)
                col += 1
    for bits in (2, 4, 8):
        w = torch.randn(6, 16)
        q = J.fused_nbit_rowwise_quantize(w, bits)
        assert q.dtype == torch.uint8 and q.shape == (6, 16 * bits // 8 + 4)
        back = J.fused_nbit_rowwise_dequantize(q, bits, 16)
        step = (w.max(1).values - w.min(1).values) / (2 ** bits - 1)
        assert float(((back - w).abs() - 0.51 * step[:, None]).max()) < 2e-2, bits
