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
    if configs is not None:
        return _validate_feature_range(kjt, configs)
    return True


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
