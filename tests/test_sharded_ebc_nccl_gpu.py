"""Same golden-model check as test_sharded_ebc_gloo.py on 2 GPUs: once through the fused NVLink
kernels (default on one host) and once through the portable NCCL transport."""
import os

import pytest
import torch

from tests.test_sharded_ebc_gloo import _run
from torchrec_b200.utils.multiprocess import run_multi_process

pytestmark = pytest.mark.gpu


def _need2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def _run_transport(ctx, sharding, weighted, transport):
    os.environ["TRB_TRANSPORT"] = transport
    _run(ctx, sharding=sharding, weighted=weighted)
    if transport == "p2p":
        from torchrec_b200.parallel.p2p import PeerGroup

        assert len(PeerGroup._CACHE) > 0, "fused NVLink path was not used"


@pytest.mark.parametrize("sharding", ["tw", "rw", "cw", "mixed", "twrw"])
def test_fused_nvlink_path(sharding):
    _need2()
    run_multi_process(_run_transport, world_size=2, backend="nccl", sharding=sharding, weighted=False, transport="p2p")


@pytest.mark.parametrize("sharding", ["tw", "rw"])
def test_nccl_transport(sharding):
    _need2()
    run_multi_process(_run_transport, world_size=2, backend="nccl", sharding=sharding, weighted=True, transport="nccl")
