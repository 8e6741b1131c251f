"""Measured costs of THIS framework's kernels on B200, the numbers the planner's perf model is built from.

Every entry names its measurement (all on 8 x B200 HGX boxes of this pool, CUDA events, max over ranks; write-ups under ``profiles/``):

* ``profiles/scaling_r2.md``      isolated NVLink phases at N = 2 and N = 8 (`bench.py --measure-comm`), single-process peer microbenchmark
                                   (`tools/peer_bench.py`), N = 2 timeline
* ``profiles/step_breakdown_*``   per-kernel time of the 1-GPU DLRM step
* ``MEASURED_PEAKS.json``         STREAM-style copy bandwidth of the pool (6.59 TB/s)

Nothing here is inherited from the reference's A100-era constants (``planner/constants.py:16-45`` there: 2.455 batched-copy factor, 9.22
data-parallel factor, /10 UVM rule ...): a factor that was never measured on this hardware is not in the model.
"""
from __future__ import annotations

from typing import Dict

GB = 1024 * 1024 * 1024

# ---- NVLink 5 through NVSwitch ------------------------------------------------------------------------------------------------
# one direction, one peer, kernel stores == copy engine (fused lookup + dist 680 GB/s, gradient push 670 GB/s, torch peer copy 671 GB/s)
PEER_STORE_GBPS: float = 670.0
# all-to-all among W ranks with both directions busy, as a fraction of PEER_STORE_GBPS: W = 2 fused forward 619 / push 630 GB/s;
# W = 8 fused forward 479 / push 547 GB/s (rank that sends the most). W = 4 was not measured: interpolated.
ALL_TO_ALL_EFFICIENCY: Dict[int, float] = {1: 1.0, 2: 0.93, 4: 0.85, 8: 0.765}


def all_to_all_gbps(world_size: int) -> float:
    """Per-direction GB/s one rank sustains inside an all-to-all of ``world_size`` ranks of one NVLink domain."""
    if world_size in ALL_TO_ALL_EFFICIENCY:
        eff = ALL_TO_ALL_EFFICIENCY[world_size]
    else:
        known = sorted(ALL_TO_ALL_EFFICIENCY)
        lo = max((k for k in known if k <= world_size), default=known[0])
        hi = min((k for k in known if k >= world_size), default=known[-1])
        eff = ALL_TO_ALL_EFFICIENCY[lo] if lo == hi else ALL_TO_ALL_EFFICIENCY[lo] + (ALL_TO_ALL_EFFICIENCY[hi] - ALL_TO_ALL_EFFICIENCY[lo]) * (world_size - lo) / (hi - lo)
    return PEER_STORE_GBPS * eff


# ---- table-batched lookup kernels -------------------------------------------------------------------------------------------------
# forward (tbe_pooled_fwd_chunk, one-hot fp32 rows -> bf16): 4 tables x 262144 rows, 805 MB of row reads + pooled writes in 135 us
# = 5.96 TB/s of algorithmic bytes against the 6.59 TB/s copy peak (hot rows of small tables are served by the 126 MB L2)
LOOKUP_FWD_EFFICIENCY: float = 0.90
# backward (tbe_bwd_chunk + span kernels; keys + radix sort run on an auxiliary stream, off the critical path): 852 k ids x (256 B gradient
# row + 512 B row read + 512 B row write) = 1.09 GB in 260 + 65 us = 3.4 TB/s
LOOKUP_BWD_EFFICIENCY: float = 0.51
# rows narrower than one 512 B warp request: dim 64 / dim 32 forward kernels against dim 128 (ab runs of round 1, same ids)
HALF_ROW_PENALTY: float = 1.15
QUARTER_ROW_PENALTY: float = 1.75


def fused_overlap(world_size: int) -> float:
    """Share of min(lookup, transfer) that the fused lookup + output-dist kernel hides: T = lookup + transfer - overlap * min(...).
    W = 2: 110 us lookup, 174 us transfer, 176 us fused -> 0.98; W = 8: 135 / 427 / 490 us -> 0.53."""
    if world_size <= 2:
        return 0.98
    if world_size >= 8:
        return 0.53
    return 0.98 + (0.53 - 0.98) * (world_size - 2) / 6.0


# gradient dist: captured inside the dense backward graph beside the deferred weight-gradient GEMMs; what stays exposed of the push at
# W = 8 (348 us standalone, N = 8 step 2.595 -> 2.354 ms when it moved into the graph together with the other overlaps)
GRAD_PUSH_EXPOSED: float = 0.45
# device-side barrier, 8 ranks (trb_barrier_kernel): 11 us, three per step
BARRIER_MS: float = 0.011
# input dist (kjt_route: lengths + scan + peer write): 35 us for 26 x 32768 one-hot ids, hidden by the pipeline
INPUT_DIST_MS_PER_MILLION_IDS: float = 0.041
