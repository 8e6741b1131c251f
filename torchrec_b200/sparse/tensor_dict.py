"""TensorDict interop (reference torchrec/sparse/tensor_dict.py:17): accept a TensorDict-like mapping of jagged features
wherever a KeyedJaggedTensor is expected. ``tensordict`` itself is optional (not in this image): anything exposing
``keys()`` and per-key objects with ``values``/``lengths`` (or ``_values``/``_lengths`` / nested tensors) works."""
from typing import Any, List, Optional

import torch

from .jagged_tensor import JaggedTensor, KeyedJaggedTensor


def _jagged_parts(x: Any):
    if isinstance(x, JaggedTensor):
        return x.values(), x.lengths()
    if isinstance(x, torch.Tensor) and x.is_nested:
        parts = x.unbind()
        return torch.cat([p.reshape(-1) for p in parts]) if parts else x.new_zeros(0), torch.tensor([p.numel() for p in parts], dtype=torch.int64)
    for v, l in (("values", "lengths"), ("_values", "_lengths")):
        if hasattr(x, v) and hasattr(x, l):
            vv, ll = getattr(x, v), getattr(x, l)
            return (vv() if callable(vv) else vv), (ll() if callable(ll) else ll)
    if isinstance(x, (list, tuple)):
        return torch.cat([torch.as_tensor(p).reshape(-1) for p in x]), torch.tensor([torch.as_tensor(p).numel() for p in x], dtype=torch.int64)
    raise TypeError(f"cannot interpret {type(x)} as a jagged feature")


def maybe_td_to_kjt(features: Any, keys: Optional[List[str]] = None) -> KeyedJaggedTensor:
    """KeyedJaggedTensor passthrough; TensorDict / mapping of jagged features -> KJT (key order = ``keys`` or mapping order)."""
    if isinstance(features, KeyedJaggedTensor):
        return features
    if not hasattr(features, "keys"):
        raise TypeError(f"expected a KeyedJaggedTensor or a TensorDict-like mapping, got {type(features)}")
    ks = list(keys) if keys is not None else [k for k in features.keys()]
    vals, lens = [], []
    for k in ks:
        v, l = _jagged_parts(features[k])
        vals.append(v)
        lens.append(l)
    return KeyedJaggedTensor(keys=ks, values=torch.cat(vals), lengths=torch.cat(lens))
