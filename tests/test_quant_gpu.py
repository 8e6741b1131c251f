"""Quantized inference lookup kernel (csrc/tbe_quant.cu) on the GPU: every row format - FP32 / FP16 / BF16, row-wise INT8 / INT4 / INT2 with fused
fp16 scale + bias, block-scaled FP8 (e4m3 + fp16 scale per 32 elements) - against the fp32 lookup over the DEQUANTISED tables."""
import pytest
import torch

from torchrec_b200.ops.quant_tbe import QuantTableBatchedEmbeddingBags, dequantize_rows, quantize_rows, row_bytes
from torchrec_b200.types import DataType

pytestmark = pytest.mark.gpu

FORMATS = [DataType.FP32, DataType.FP16, DataType.BF16, DataType.INT8, DataType.INT4, DataType.INT2, DataType.FP8]


def _ids(F, B, rows, maxL, seed):
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(0, maxL + 1, (F * B,), generator=g)
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lengths.cumsum(0)])
    idx = torch.cat([torch.randint(0, rows[f], (int(lengths[f * B : (f + 1) * B].sum()),), generator=g) for f in range(F)])
    if idx.numel() > 2:
        idx[0], idx[1] = -1, rows[0] + 5  # invalid ids contribute zero
    return idx, offsets, torch.rand(idx.numel(), generator=g)


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("dim", [32, 128, 256])
def test_quantized_pooled_lookup_matches_dequantised_reference(fmt, dim):
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    specs = [("a", 97, dim, fmt), ("b", 400, dim, fmt), ("c", 33, dim, fmt)]
    fmap = [0, 1, 1, 2]
    B = 37
    gpu = QuantTableBatchedEmbeddingBags(specs, fmap, pooling_mode=0, device=dev)
    cpu = QuantTableBatchedEmbeddingBags(specs, fmap, pooling_mode=0)
    for t, (_, r, d, _) in enumerate(specs):
        w = torch.randn(r, d) * 0.3
        gpu.assign_from_float(t, w)
        cpu.assign_from_float(t, w)
        assert torch.equal(gpu.split_embedding_weights()[t].cpu(), cpu.split_embedding_weights()[t])  # same bytes from the CPU and GPU quantizers
    idx, off, psw = _ids(4, B, [97, 400, 400, 33], 6, 11)
    for pooling, weighted in ((0, False), (0, True), (1, False)):
        gpu.pooling_mode = cpu.pooling_mode = pooling
        out = gpu(idx.to(dev), off.to(dev), psw.to(dev) if weighted else None)
        ref = cpu(idx, off, psw if weighted else None)
        torch.testing.assert_close(out.cpu(), ref, rtol=2e-3, atol=2e-3)
    gpu.pooling_mode = cpu.pooling_mode = 0
    gpu.output_dtype = torch.bfloat16
    out = gpu(idx.to(dev), off.to(dev))
    torch.testing.assert_close(out.float().cpu(), cpu(idx, off), rtol=2e-2, atol=3e-2)


@pytest.mark.parametrize("fmt", [DataType.INT8, DataType.INT4, DataType.FP8, DataType.FP16])
def test_quantized_sequence_lookup(fmt):
    dev = torch.device("cuda:0")
    dim = 64
    specs = [("a", 50, dim, fmt), ("b", 70, dim, fmt)]
    gpu = QuantTableBatchedEmbeddingBags(specs, [0, 1], pooling_mode=2, device=dev)
    cpu = QuantTableBatchedEmbeddingBags(specs, [0, 1], pooling_mode=2)
    for t, (_, r, d, _) in enumerate(specs):
        w = torch.randn(r, d)
        gpu.assign_from_float(t, w)
        cpu.assign_from_float(t, w)
    idx, off, _ = _ids(2, 9, [50, 70], 4, 5)
    torch.testing.assert_close(gpu(idx.to(dev), off.to(dev)).cpu(), cpu(idx, off), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("fmt", [DataType.INT8, DataType.INT4, DataType.INT2, DataType.FP8])
def test_quantisation_error_bounds(fmt):
    """Round trip error of the row formats: what the serving tables lose against the trained fp32 rows."""
    w = torch.randn(64, 128, device="cuda")
    q = quantize_rows(w, fmt)
    assert q.shape == (64, row_bytes(128, fmt))
    err = (dequantize_rows(q, 128, fmt) - w).abs().max().item()
    span = (w.max(1).values - w.min(1).values).max().item()
    bound = {DataType.INT8: span / 255 * 0.6, DataType.INT4: span / 15 * 0.6, DataType.INT2: span / 3 * 0.6, DataType.FP8: w.abs().max().item() * 2 ** -3}[fmt]
    assert err <= bound + 1e-3, (err, bound)
