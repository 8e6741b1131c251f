"""CUDA kernels of the jagged-op and quantized-communication families against their PyTorch reference implementations (the CPU path of
the same functions): block_bucketize (even / uneven blocks, overflow ids, positions, unbucketize), jagged <-> padded dense, segment sums,
pooled-embedding column permutation, FP8 / INT8 / MX4 wire codecs."""
import pytest
import torch

from torchrec_b200.ops import jagged as J

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _jagged(F, B, maxL, hi, seed, dtype=torch.int64):
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(0, maxL + 1, (F * B,), generator=g)
    vals = torch.randint(0, hi, (int(lengths.sum()),), generator=g).to(dtype)
    return lengths, vals, torch.rand(vals.numel(), generator=g)


@pytest.mark.parametrize("my_size", [2, 8])
@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
@pytest.mark.parametrize("uneven", [False, True])
def test_block_bucketize_matches_reference(my_size, idx_dtype, uneven):
    F, B = 3, 37
    rows = [1000, 64, 333]
    lengths, vals, w = _jagged(F, B, 6, 10, 1, idx_dtype)
    # ids per feature in that feature's range, a few past the last block (round-robin rule) for the even layout
    off = torch.cat([torch.zeros(1, dtype=torch.int64), lengths.cumsum(0)])
    g = torch.Generator().manual_seed(9)
    for f in range(F):
        a, b = int(off[f * B]), int(off[(f + 1) * B])
        vals[a:b] = torch.randint(0, rows[f], (b - a,), generator=g).to(idx_dtype)
    block_sizes = torch.tensor([(r + my_size - 1) // my_size for r in rows], dtype=idx_dtype)
    pos = None
    if uneven:
        pos = []
        for r in rows:
            cuts = torch.sort(torch.randint(0, r, (my_size - 1,), generator=g)).values
            pos.append(torch.cat([torch.zeros(1, dtype=torch.int64), cuts, torch.tensor([r])]).to(idx_dtype))
    else:
        vals[:4] = torch.tensor([block_sizes[0] * my_size + 5, block_sizes[0] * my_size + 6, 3, 7]).to(idx_dtype)[: min(4, vals.numel())]
    ref = J.block_bucketize_sparse_features(lengths, vals, True, True, block_sizes, my_size, w, block_bucketize_pos=pos)
    got = J.block_bucketize_sparse_features(lengths.to(DEV), vals.to(DEV), True, True, block_sizes.to(DEV), my_size, w.to(DEV),
                                            block_bucketize_pos=None if pos is None else [p.to(DEV) for p in pos])
    for name, r, gt in zip(["lengths", "indices", "weights", "pos", "unbucketize"], ref, got):
        assert torch.equal(gt.cpu(), r) if name != "weights" else torch.allclose(gt.cpu(), r), name
    got2 = J.block_bucketize_sparse_features(lengths.to(DEV), vals.to(DEV), False, False, block_sizes.to(DEV), my_size,
                                             block_bucketize_pos=None if pos is None else [p.to(DEV) for p in pos])
    assert got2[2] is None and got2[3] is None and got2[4] is None and torch.equal(got2[1].cpu(), ref[1])


@pytest.mark.parametrize("D", [0, 1, 16])
def test_jagged_padded_dense_round_trip(D):
    lengths, _, _ = _jagged(1, 50, 7, 10, 2)
    off = torch.cat([torch.zeros(1, dtype=torch.int64), lengths.cumsum(0)])
    total = int(off[-1])
    vals = torch.randn(total) if D == 0 else torch.randn(total, D)
    ref = J.jagged_to_padded_dense(vals, [off], [5], -1.5)
    got = J.jagged_to_padded_dense(vals.to(DEV), [off.to(DEV)], [5], -1.5)
    torch.testing.assert_close(got.cpu(), ref)
    if D > 0:
        dense = torch.randn(50, 8, D)
        ref_j, _ = J.dense_to_jagged(dense, [off.clamp(max=10**9)], total_L=None)
        lengths_c = (off[1:] - off[:-1]).clamp(max=8)
        off_c = torch.cat([torch.zeros(1, dtype=torch.int64), lengths_c.cumsum(0)])
        ref_j, _ = J.dense_to_jagged(dense, [off_c])
        got_j, _ = J.dense_to_jagged(dense.to(DEV), [off_c.to(DEV)], total_L=int(off_c[-1]))
        torch.testing.assert_close(got_j.cpu(), ref_j)


def test_segment_sum_and_permute_pooled():
    lengths, _, _ = _jagged(6, 20, 9, 10, 3)
    csr = torch.tensor([0, 1, 3, 6], dtype=torch.int64)
    for vals in (lengths.float(), lengths, lengths.int()):
        ref = J.segment_sum_csr(20, csr, vals)
        got = J.segment_sum_csr(20, csr.to(DEV), vals.to(DEV))
        assert torch.allclose(got.cpu().double(), ref.double())
    x = torch.randn(33, 40).bfloat16()
    offs, perm = [0, 8, 24, 32, 40], [2, 0, 3, 1]
    ref = J.permute_pooled_embs(x, offs, perm)
    got = J.permute_pooled_embs(x.to(DEV), offs, perm)
    assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("kind", ["FP8", "INT8", "MX4"])
def test_wire_codecs_cuda_vs_reference(kind):
    from torchrec_b200.parallel.qcomm_codec import CommType, get_qcomm_codec

    codec = get_qcomm_codec(getattr(CommType, kind), None, 64)
    x = torch.randn(37 * 64) * torch.logspace(-3, 2, 37 * 64)
    x[5 * 64 : 6 * 64] = 0.0  # an all-zero row / group
    enc_ref = codec.encode(x)
    enc = codec.encode(x.to(DEV))
    assert enc.numel() == codec.calc_quantized_size(x.numel()) == enc_ref.numel()
    dec_ref = codec.decode(enc_ref)
    dec = codec.decode(enc)
    # same wire format: CPU bytes decode on the GPU and vice versa
    torch.testing.assert_close(codec.decode(enc_ref.to(DEV)).cpu(), dec_ref, rtol=3e-5, atol=1e-6)  # fma vs mul + add
    torch.testing.assert_close(codec.decode(enc.cpu()), dec.cpu(), rtol=3e-5, atol=1e-6)
    # and both quantizers land within one quantization step of the input
    rows = x.view(-1, 64 if kind != "MX4" else 32)
    step = {"FP8": rows.abs().amax(1, keepdim=True) * 2 ** -3, "INT8": (rows.amax(1, keepdim=True) - rows.amin(1, keepdim=True)) / 255,
            "MX4": rows.abs().amax(1, keepdim=True) * 0.5}[kind]
    err = (dec.cpu().view(rows.shape) - rows).abs()
    assert bool((err <= step + 1e-6).all()), float((err - step).max())
    err_ref = (dec_ref.view(rows.shape) - rows).abs()
    assert bool((err_ref <= step + 1e-6).all())
