"""Hardware model constants of the planner — B200 / NVLink-5 defaults.

The reference models an A100-era box (HBM 32 GB @ 897 GB/s, 600 GB/s intra-node, 12.5 GB/s cross-node,
planner/constants.py:16-36); here the defaults describe one 8xB200 HGX node: 180 GB HBM3e per GPU, the
MEASURED copy bandwidth of this pool (6.6 TB/s, MEASURED_PEAKS.json), 900 GB/s per direction NVLink 5
through NVSwitch (770 GB/s measured peer copy) with every peer at full bandwidth, 400 Gb/s NICs across
nodes.
"""
from typing import Dict, Optional

from ..embedding_types import EmbeddingComputeKernel

MAX_SIZE: int = (1 << 63) - 1

INTRA_NODE_BANDWIDTH: float = 770 * 1024 * 1024 * 1024 / 1000  # bytes / ms  (measured NVLink-5 peer copy, per direction)
CROSS_NODE_BANDWIDTH: float = 50 * 1024 * 1024 * 1024 / 1000  # bytes / ms  (400 Gb/s NIC)

MIN_CW_DIM: int = 128
POOLING_FACTOR: float = 1.0

BIGINT_DTYPE: int = 8

HBM_CAP: int = 180 * 1024 * 1024 * 1024  # 180 GB HBM3e
DDR_CAP: int = 2048 * 1024 * 1024 * 1024
SSD_CAP: int = 16 * 1024 * 1024 * 1024 * 1024
DDR_MEM_BW: float = 400 * 1024 * 1024 * 1024 / 1000  # bytes / ms
HBM_MEM_BW: float = 6588 * 1024 * 1024 * 1024 / 1000  # bytes / ms (measured STREAM-style copy)
SSD_MEM_BW: float = 7 * 1024 * 1024 * 1024 / 1000
HBM_TO_DDR_MEM_BW: float = 55 * 1024 * 1024 * 1024 / 1000  # PCIe Gen5 x16
UVM_CACHING_RATIO: float = 0.2
BATCH_SIZE: int = 512

BATCHED_COPY_PERF_FACTOR: float = 2.455  # empirical fwd/bwd asymmetry of scatter-style copies
FULL_BLOCK_EMB_DIM: int = 128  # one warp moves a 512 B (128 x fp32) row per request
HALF_BLOCK_PENALTY: float = 1.15
QUARTER_BLOCK_PENALTY: float = 1.75
BWD_COMPUTE_MULTIPLIER: float = 2  # sort + reduce + optimizer write-back
WEIGHTED_KERNEL_MULTIPLIER: float = 1.1
DP_ELEMENTWISE_KERNELS_PERF_FACTOR: float = 9.22


def kernel_bw_lookup(compute_device: str, compute_kernel: str, hbm_mem_bw: float, ddr_mem_bw: float, hbm_to_ddr_mem_bw: float,
                     caching_ratio: Optional[float] = None, prefetch_pipeline: bool = False) -> Optional[float]:
    """Effective bandwidth (bytes/ms) a lookup kernel sees for a compute kernel / placement."""
    caching_ratio = caching_ratio if caching_ratio else UVM_CACHING_RATIO
    lookup: Dict = {
        ("cpu", EmbeddingComputeKernel.DENSE.value): 0.5 * ddr_mem_bw,
        ("cpu", EmbeddingComputeKernel.FUSED.value): 1 * ddr_mem_bw,
        ("cpu", EmbeddingComputeKernel.QUANT.value): 1 * ddr_mem_bw,
        ("cuda", EmbeddingComputeKernel.DENSE.value): 0.5 * hbm_mem_bw,
        ("cuda", EmbeddingComputeKernel.FUSED.value): 1 * hbm_mem_bw,
        ("cuda", EmbeddingComputeKernel.FUSED_UVM.value): hbm_to_ddr_mem_bw / 10,
        ("cuda", EmbeddingComputeKernel.FUSED_UVM_CACHING.value): (caching_ratio * hbm_mem_bw + (1 - caching_ratio) * hbm_to_ddr_mem_bw) / 10,
        ("cuda", EmbeddingComputeKernel.QUANT.value): 1 * hbm_mem_bw,
        ("cuda", EmbeddingComputeKernel.QUANT_UVM.value): hbm_to_ddr_mem_bw / 10,
        ("cuda", EmbeddingComputeKernel.QUANT_UVM_CACHING.value): (caching_ratio * hbm_mem_bw + (1 - caching_ratio) * hbm_to_ddr_mem_bw) / 10,
        ("cuda", EmbeddingComputeKernel.KEY_VALUE.value): hbm_to_ddr_mem_bw,
        ("cuda", EmbeddingComputeKernel.SSD_VIRTUAL_TABLE.value): SSD_MEM_BW,
        ("cuda", EmbeddingComputeKernel.DRAM_VIRTUAL_TABLE.value): hbm_to_ddr_mem_bw,
    }
    if prefetch_pipeline and compute_device == "cuda" and compute_kernel == EmbeddingComputeKernel.FUSED_UVM_CACHING.value:
        return lookup.get(("cuda", EmbeddingComputeKernel.FUSED.value))
    return lookup.get((compute_device, compute_kernel))
