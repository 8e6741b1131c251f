"""Move trained (sharded) state dicts into serving shape. Parity: reference ``inference/state_dict_transform.py:18-108``."""
from __future__ import annotations

from typing import Dict, List, Union

import torch
import torch.distributed as dist
from torch.distributed._shard.sharded_tensor import Shard, ShardedTensor


def state_dict_gather(src: Dict[str, Union[torch.Tensor, ShardedTensor]], dst: Dict[str, torch.Tensor]) -> None:
    """Fill ``dst`` (full, unsharded tensors; only rank 0's copy is written for sharded entries) from the per-rank ``src``."""
    for key, dst_tensor in dst.items():
        src_tensor = src[key]
        if isinstance(src_tensor, ShardedTensor):
            src_tensor.gather(out=dst_tensor if dist.get_rank() == 0 else None, dtype=dst_tensor.dtype)
        elif isinstance(src_tensor, torch.Tensor):
            dst_tensor.copy_(src_tensor)
        else:
            raise ValueError(f"Unsupported tensor {key} type {type(src_tensor)}")


def state_dict_all_gather_keys(state_dict: Dict[str, Union[torch.Tensor, ShardedTensor]], pg: dist.ProcessGroup) -> List[str]:
    """Sorted union of the keys every rank holds (ranks own different shards / tables)."""
    names = list(state_dict.keys())
    all_names: List[List[str]] = [None] * dist.get_world_size(pg)  # type: ignore[list-item]
    dist.all_gather_object(all_names, names, pg)
    return sorted({n for local in all_names for n in local})


def state_dict_to_device(state_dict: Dict[str, Union[torch.Tensor, ShardedTensor]], pg: dist.ProcessGroup, device: torch.device) -> Dict[str, Union[torch.Tensor, ShardedTensor]]:
    """Copy a state dict to ``device``; ShardedTensors are rebuilt from their moved local shards (collective over ``pg``)."""
    ret: Dict[str, Union[torch.Tensor, ShardedTensor]] = {}
    for key in state_dict_all_gather_keys(state_dict, pg):
        if key not in state_dict:
            continue
        t = state_dict[key]
        if isinstance(t, ShardedTensor):
            shards = [Shard.from_tensor_and_offsets(tensor=s.tensor.to(device), shard_offsets=s.metadata.shard_offsets, rank=dist.get_rank(pg)) for s in t.local_shards()]
            ret[key] = ShardedTensor._init_from_local_shards(shards, t.metadata().size, process_group=pg)
        elif isinstance(t, torch.Tensor):
            ret[key] = t.to(device)
        else:
            raise ValueError(f"Unsupported tensor {key} type {type(t)}")
    return ret
