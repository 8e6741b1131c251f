"""Fused CTR head kernels vs plain PyTorch fp32 references."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,K", [(1000, 256), (33, 64), (4096, 1024)])
@pytest.mark.parametrize("mask", [False, True])
def test_rowdot_matches_linear(B, K, mask):
    from torchrec_b200.ops.head import RowDotFn

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    pre = torch.randn(B, K, device=dev)
    x = (torch.relu(pre) if mask else pre).to(torch.bfloat16)
    w = torch.randn(1, K, device=dev, requires_grad=True)
    b = torch.randn(1, device=dev, requires_grad=True)
    xg = x.clone().requires_grad_()
    if mask:
        xg._trb_relu_out = True
    out = RowDotFn.apply(xg, w, b)
    ref_x = x.float().requires_grad_()
    ref = ref_x @ w.detach().t() + b.detach()
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-3)
    g = torch.randn(B, 1, device=dev)
    gx, gw, gb = torch.autograd.grad(out, (xg, w, b), g)
    rgx = g * w.detach()
    if mask:
        rgx = rgx * (x.float() > 0)
        assert getattr(gx, "_trb_masked", False)
    torch.testing.assert_close(gx.float(), rgx, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(gw, (g * x.float()).sum(0, keepdim=True), rtol=1e-3, atol=1e-2)
    torch.testing.assert_close(gb, g.sum().reshape(1), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("ldtype", [torch.float32, torch.int64])
def test_bce_fwd_bwd(ldtype):
    from torchrec_b200.ops.head import bce_with_logits_mean

    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    z = (torch.randn(5000, device=dev) * 4).requires_grad_()
    y = torch.randint(0, 2, (5000,), device=dev).to(ldtype)
    loss = bce_with_logits_mean(z, y)
    z2 = z.detach().clone().requires_grad_()
    ref = torch.nn.functional.binary_cross_entropy_with_logits(z2, y.float())
    torch.testing.assert_close(loss, ref, rtol=1e-5, atol=1e-6)
    (loss * 3).backward()
    (ref * 3).backward()
    torch.testing.assert_close(z.grad, z2.grad, rtol=1e-4, atol=1e-8)
