"""Structural validation of KeyedJaggedTensors (reference torchrec/sparse/jagged_tensor_validator.py:20-304).
Run on the first batches of a job: catches malformed inputs before they become out-of-bounds gathers on the GPU."""
from typing import Dict, List, Optional

import torch

from .jagged_tensor import KeyedJaggedTensor


def validate_keyed_jagged_tensor(kjt: KeyedJaggedTensor, configs: Optional[List] = None) -> bool:
    """Raises ``ValueError`` describing the first inconsistency; returns True when ids are within the configs' ranges
    (False when some feature has out-of-range ids and ``configs`` is given)."""
    _validate_lengths_and_offsets(kjt)
    _validate_keys(kjt)
    _validate_weights(kjt)
    _validate_vbe_properties(kjt)
    if configs is not None:
        return _validate_feature_range(kjt, configs)
    return True


def collect_problems(kjt: KeyedJaggedTensor, configs: Optional[List] = None) -> List[str]:
    """Every inconsistency instead of the first one (data-pipeline debugging): one line per failed check, ``[]`` for a clean batch.
    Out-of-range ids are reported per feature with their count and extreme values."""
    problems: List[str] = []
    for check in (_validate_lengths_and_offsets, _validate_keys, _validate_weights, _validate_vbe_properties):
        try:
            check(kjt)
        except ValueError as e:
            problems.append(str(e))
    if configs is not None and not problems:
        problems.extend(out_of_range_report(kjt, configs))
    return problems


def out_of_range_report(kjt: KeyedJaggedTensor, configs: List) -> List[str]:
    limits: Dict[str, int] = {f: c.num_embeddings for c in configs for f in c.feature_names}
    lines: List[str] = []
    for key, vals in zip(kjt.keys(), torch.split(kjt.values(), kjt.length_per_key())):
        if key not in limits or vals.numel() == 0:
            continue
        bad = (vals < 0) | (vals >= limits[key])
        n = int(bad.sum())
        if n:
            lines.append(f"feature {key}: {n} of {vals.numel()} ids outside [0, {limits[key]}) (min {int(vals.min())}, max {int(vals.max())})")
    return lines


def validate_on_device(kjt: KeyedJaggedTensor, configs: Optional[List] = None) -> torch.Tensor:
    """The numeric checks as ONE device-side reduction, no host sync: returns an int32 bit mask tensor on the KJT's device
    (bit 0 negative length, bit 1 offsets not monotonic, bit 2 lengths / offsets / values sizes disagree, bit 3 id out of range).
    Meant to be read back lazily (``mask.item()`` a few steps later, or folded into a metric) inside a training loop."""
    values = kjt.values()
    dev = values.device
    mask = torch.zeros((), dtype=torch.int32, device=dev)
    lengths, offsets = kjt.lengths_or_none(), kjt.offsets_or_none()
    if lengths is not None and lengths.numel():
        mask = mask | (lengths.min() < 0).to(torch.int32)
        mask = mask | ((lengths.sum() != values.numel()).to(torch.int32) << 2)
    if offsets is not None and offsets.numel() > 1:
        mask = mask | ((offsets[1:] < offsets[:-1]).any().to(torch.int32) << 1)
        mask = mask | (((offsets[-1] != values.numel()) | (offsets[0] != 0)).to(torch.int32) << 2)
    if configs is not None and values.numel():
        limits: Dict[str, int] = {f: c.num_embeddings for c in configs for f in c.feature_names}
        lpk = kjt.length_per_key()
        hi = torch.repeat_interleave(torch.tensor([limits.get(k, 1 << 62) for k in kjt.keys()], dtype=torch.int64, device=dev),
                                     torch.tensor(lpk, dtype=torch.int64, device=dev), output_size=values.numel())
        mask = mask | (((values < 0) | (values.long() >= hi)).any().to(torch.int32) << 3)
    return mask


def _validate_lengths_and_offsets(kjt: KeyedJaggedTensor) -> None:
    lengths, offsets = kjt.lengths_or_none(), kjt.offsets_or_none()
    if lengths is None and offsets is None:
        raise ValueError("lengths and offsets cannot be both empty")
    if lengths is not None:
        if lengths.dim() != 1:
            raise ValueError(f"lengths must be 1-D, got {lengths.dim()}-D")
        if lengths.numel() and int(lengths.min()) < 0:
            raise ValueError("lengths must be non-negative")
        if int(lengths.sum()) != kjt.values().numel():
            raise ValueError(f"Sum of lengths must equal the number of values, but got {int(lengths.sum())} and {kjt.values().numel()}")
    if offsets is not None:
        if offsets.dim() != 1:
            raise ValueError(f"offsets must be 1-D, got {offsets.dim()}-D")
        if offsets.numel() == 0:
            raise ValueError("offsets cannot be empty")
        if int(offsets[0]) != 0:
            raise ValueError(f"Expected first offset to be 0, but got {int(offsets[0])}")
        if offsets.numel() > 1 and bool((offsets[1:] < offsets[:-1]).any()):
            raise ValueError("offsets is not equal to the cumulative sum of lengths (must be non-decreasing)")
        if int(offsets[-1]) != kjt.values().numel():
            raise ValueError(f"The last element of offsets must equal the number of values, but got {int(offsets[-1])} and {kjt.values().numel()}")
    if lengths is not None and offsets is not None:
        if lengths.numel() + 1 != offsets.numel():
            raise ValueError(f"Expected lengths size to be 1 more than offsets size, but got lengths size: {lengths.numel()} and offsets size: {offsets.numel()}")
        if not torch.equal(torch.cumsum(lengths, 0).to(offsets.dtype), offsets[1:]):
            raise ValueError("offsets is not equal to the cumulative sum of lengths")


def _validate_keys(kjt: KeyedJaggedTensor) -> None:
    keys = kjt.keys()
    if len(set(keys)) != len(keys):
        raise ValueError(f"keys must be unique, but got {keys}")
    n = kjt.lengths_or_none().numel() if kjt.lengths_or_none() is not None else kjt.offsets().numel() - 1
    if len(keys) == 0:
        if n != 0:
            raise ValueError("Expected empty lengths / offsets when there are no keys")
        return
    if kjt.variable_stride_per_key():
        if n != sum(kjt.stride_per_key()):
            raise ValueError(f"lengths size {n} does not match the sum of strides per key {sum(kjt.stride_per_key())}")
    elif n % len(keys) != 0:
        raise ValueError(f"lengths size must be divisible by keys size, but got {n} and {len(keys)}")


def _validate_weights(kjt: KeyedJaggedTensor) -> None:
    w = kjt.weights_or_none()
    if w is not None and w.numel() != kjt.values().numel():
        raise ValueError(f"weights size must equal to values size, but got {w.numel()} and {kjt.values().numel()}")


def _validate_vbe_properties(kjt: KeyedJaggedTensor) -> None:
    """Variable-batch-per-feature metadata: ``stride_per_key_per_rank`` is [keys][ranks] of non-negative batch sizes whose per-key sums
    are the strides; ``inverse_indices`` = (keys, [keys, full_batch]) maps every sample of the full batch to a row of the (deduplicated)
    per-key batch, so its entries index below that key's stride (reference :191-262)."""
    spkpr = kjt.stride_per_key_per_rank() if kjt.variable_stride_per_key() else []
    inv = kjt.inverse_indices_or_none()
    keys = kjt.keys()
    if spkpr:
        if len(spkpr) != len(keys):
            raise ValueError(f"stride_per_key_per_rank has {len(spkpr)} rows for {len(keys)} keys")
        widths = {len(r) for r in spkpr}
        if len(widths) > 1:
            raise ValueError(f"stride_per_key_per_rank rows must have one entry per rank, got row sizes {sorted(widths)}")
        if any(b < 0 for r in spkpr for b in r):
            raise ValueError("stride_per_key_per_rank entries must be non-negative")
    if inv is None:
        return
    inv_keys, inv_t = inv
    if list(inv_keys) != list(keys):
        raise ValueError(f"inverse_indices keys {list(inv_keys)} do not match the KJT keys {list(keys)}")
    if inv_t.dim() != 2 or inv_t.shape[0] != len(keys):
        raise ValueError(f"inverse_indices must be [num_keys, full_batch], got {tuple(inv_t.shape)} for {len(keys)} keys")
    if not kjt.variable_stride_per_key():
        raise ValueError("inverse_indices given for a KJT without stride_per_key_per_rank")
    if inv_t.numel():
        strides = torch.tensor(kjt.stride_per_key(), dtype=inv_t.dtype, device=inv_t.device).unsqueeze(1)
        if bool((inv_t < 0).any()) or bool((inv_t >= strides).any()):
            raise ValueError("inverse_indices entries must index inside the per-key batch (0 <= index < stride of the key)")


def _validate_feature_range(kjt: KeyedJaggedTensor, configs: List) -> bool:
    limits: Dict[str, int] = {f: c.num_embeddings for c in configs for f in c.feature_names}
    ok = True
    lpk = kjt.length_per_key()
    for key, vals in zip(kjt.keys(), torch.split(kjt.values(), lpk)):
        if key in limits and vals.numel():
            lo, hi = int(vals.min()), int(vals.max())
            if lo < 0 or hi >= limits[key]:
                ok = False
    return ok
