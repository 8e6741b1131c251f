"""Sync (functional-collective) variants of the embedding collectives + custom ops + the pluggable all-to-all hook: same values and
gradients as the default request / wait implementation, 2 gloo ranks."""
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def _run(ctx):
    import torch.distributed as dist

    from torchrec_b200.parallel import comm_ops as C
    from torchrec_b200.parallel import comm_ops_sync as S

    W, me = ctx.world_size, ctx.rank
    pg = dist.group.WORLD
    torch.manual_seed(10 + me)
    B, dims = [3, 3], [4, 6]

    def run(fn):
        x = torch.randn(sum(B), dims[me], requires_grad=True)
        out = fn(x)
        (out * torch.arange(out.numel(), dtype=torch.float32).view_as(out)).sum().backward()
        return out.detach().clone(), x.grad.clone(), x.detach().clone()

    torch.manual_seed(10 + me)
    o_async, g_async, x0 = run(lambda x: C.alltoall_pooled(x, B, dims, group=pg).wait())
    C.set_use_sync_collectives(True)
    try:
        torch.manual_seed(10 + me)
        o_sync, g_sync, x1 = run(lambda x: C.alltoall_pooled(x, B, dims, group=pg).wait())
        assert torch.equal(x0, x1)
        torch.testing.assert_close(o_sync, o_async)
        torch.testing.assert_close(g_sync, g_async)
        # reduce-scatter / all-gather pair through the custom ops
        torch.manual_seed(3 + me)
        y = torch.randn(W * 2, 5, requires_grad=True)
        rs = C.reduce_scatter_base_pooled(y, group=pg).wait()
        rs.sum().backward()
        g_rs = y.grad.clone()
        C.set_use_sync_collectives(False)
        y2 = y.detach().clone().requires_grad_()
        rs2 = C.reduce_scatter_base_pooled(y2, group=pg).wait()
        rs2.sum().backward()
        torch.testing.assert_close(rs, rs2)
        torch.testing.assert_close(g_rs, y2.grad)
        C.set_use_sync_collectives(True)
        z = torch.randn(2, 5, requires_grad=True)
        ag = C.all_gather_base_pooled(z, group=pg).wait()
        assert ag.shape == (2 * W, 5)
        ag.pow(2).sum().backward()
        assert z.grad is not None and z.grad.shape == z.shape
        # uneven reduce-scatter
        splits = [1, 3]
        v = torch.randn(4, 5, requires_grad=True)
        rv = C.reduce_scatter_v_pooled(v, splits, group=pg).wait()
        assert rv.shape == (splits[me], 5)
        rv.sum().backward()
        ref_all = [torch.zeros_like(v) for _ in range(W)]
        dist.all_gather(ref_all, v.detach())
        off = sum(splits[:me])
        torch.testing.assert_close(rv.detach(), sum(t[off : off + splits[me]] for t in ref_all))
    finally:
        C.set_use_sync_collectives(False)
    # shape-only (fake) implementations: what a tracer sees
    from torch._subclasses.fake_tensor import FakeTensorMode

    with FakeTensorMode():
        f = torch.empty(8, 6)
        assert torch.ops.torchrec_b200.reduce_scatter_tensor(f, "sum", 2, pg.group_name, True).shape == (4, 6)
        assert torch.ops.torchrec_b200.all_gather_into_tensor(f, 0, 2, pg.group_name, True).shape == (16, 6)
        assert torch.ops.torchrec_b200._split_1d_cat_2d(torch.empty(3 * 10), 3, [4, 6]).shape == (3, 10)
    # pluggable transport: a counting All2AllSingle sees the forward exchange
    calls = []

    class Counting(S.DefaultAll2AllSingle):
        def allocate(self, numel, dtype, device):
            calls.append(("alloc", numel))
            return super().allocate(numel, dtype, device)

        def all_to_all_single(self, output, input, o, i, async_op=True):
            calls.append(("a2a", sum(o)))
            return super().all_to_all_single(output, input, o, i, async_op)

    x = x0.clone().requires_grad_()
    out = C.alltoall_pooled(x, B, dims, group=pg, comm=Counting(pg)).wait()
    torch.testing.assert_close(out.detach(), o_async)
    assert [c[0] for c in calls] == ["alloc", "a2a"] and calls[0][1] == B[me] * sum(dims)


def test_sync_collectives_and_custom_ops():
    run_multi_process(_run, world_size=2, backend="gloo")
