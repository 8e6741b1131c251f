"""Run ONE GEMM shape a few times (target for `ncu -k regex:gemm_bf16_tcgen05 --launch-skip 2 -c 1`).
Usage: python tools/gemm_one.py M N K [tile_n] [bias_relu]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchrec_b200.ops import gemm as G  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
tile_n = int(sys.argv[4]) if len(sys.argv) > 4 else 0
bias_relu = int(sys.argv[5]) if len(sys.argv) > 5 else 1
dev = torch.device("cuda:0")
a = torch.randn(M, K, device=dev).bfloat16()
w = torch.randn(N, K, device=dev).bfloat16()
bias = torch.zeros(N, device=dev)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(5):
    if bias_relu:
        G.gemm_bf16(a, w, bias=bias, act=G.ACT_RELU, out=out, tile_n=tile_n)
    else:
        G.gemm_bf16(a, w, out=out, tile_n=tile_n)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
