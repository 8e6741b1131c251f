"""Lazy-module helpers: deferred ``apply`` and a LazyModule mixin whose ``forward`` may take keyword arguments.

Parity: torchrec/modules/lazy_extension.py:24-257 (``lazy_apply``, ``LazyModuleExtensionMixin``). The reference re-implements
``Module._call_impl`` to thread kwargs through the parameter-inference pre-hook; current PyTorch passes kwargs to
``with_kwargs=True`` pre-hooks natively, so the mixin here only adds the two behaviours that are still missing upstream:
a guarded ``apply`` and the post-first-forward function queue."""
from __future__ import annotations

from typing import Callable, List

import torch
from torch.nn.modules.lazy import LazyModuleMixin

_QUEUE = "_trb_lazy_apply_fns"
_HOOK = "_trb_lazy_apply_hook"


def _has_uninitialized(module: torch.nn.Module) -> bool:
    for t in list(module.parameters(recurse=True)) + list(module.buffers(recurse=True)):
        if isinstance(t, torch.nn.parameter.UninitializedTensorMixin):
            return True
    return False


def _drain(module: torch.nn.Module, *_: object) -> None:
    fns: List[Callable[[torch.nn.Module], None]] = module.__dict__.pop(_QUEUE, [])
    hook = module.__dict__.pop(_HOOK, None)
    if hook is not None:
        hook.remove()
    for fn in fns:
        torch.nn.Module.apply(module, fn)


def lazy_apply(module: torch.nn.Module, fn: Callable[[torch.nn.Module], None]) -> torch.nn.Module:
    """Queue ``fn`` to be ``apply``-ed to ``module`` (and all sub-modules) right after its FIRST forward, i.e. once every lazy
    parameter has a shape. Works on lazy and non-lazy modules; functions run once, in registration order."""
    queue = module.__dict__.setdefault(_QUEUE, [])
    queue.append(fn)
    if _HOOK not in module.__dict__:
        module.__dict__[_HOOK] = module.register_forward_hook(_drain)
    return module


class LazyModuleExtensionMixin(LazyModuleMixin):
    """``LazyModuleMixin`` with (1) keyword arguments forwarded to ``initialize_parameters`` and (2) an ``apply`` that refuses
    to touch shape-less parameters (use :func:`lazy_apply` to defer the function instead)."""

    def apply(self, fn: Callable[[torch.nn.Module], None]) -> torch.nn.Module:
        if hasattr(self, "_initialize_hook") or _has_uninitialized(self):  # type: ignore[arg-type]
            raise RuntimeError(
                f"{type(self).__name__} has uninitialized parameters: apply(fn) would see tensors without a shape. "
                "Call lazy_apply(module, fn) to run fn after the first forward pass."
            )
        return super().apply(fn)  # type: ignore[misc]

    def _infer_parameters(self, module, args, kwargs=None):  # type: ignore[override]
        kwargs = kwargs or {}
        module.initialize_parameters(*args, **kwargs)
        if module.has_uninitialized_params():
            raise RuntimeError(f"module {self._get_name()} has not been fully initialized")  # type: ignore[attr-defined]
        module._initialize_hook.remove()
        module._load_hook.remove()
        delattr(module, "_initialize_hook")
        delattr(module, "_load_hook")
        if module.cls_to_become is not None:
            module.__class__ = module.cls_to_become
