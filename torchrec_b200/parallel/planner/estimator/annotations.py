"""Decorators that turn a plain class into a hardware perf config.

Reference: ``torchrec/distributed/planner/estimator/annotations.py`` - hardware class decorators :241-455, formula method decorators
(``forward_compute`` :458, ``backward_compute`` :554, ``prefetch_compute`` :630, ``input_dist_comms`` :651, ``fwd_comms`` :743, ``bwd_comms`` :835,
``output_write_size`` :927), behaviour switches :1028-1187, coefficient providers :1190-1286 and the ``get_*`` resolvers.

    @hbm_mem_bw(6.5e6)
    @intra_host_bw(7.7e5)
    class MyGpu(HardwarePerfConfig):
        @forward_compute(sharding_type="row_wise")
        def rw_fwd(self, ctx):            # overrides the built-in row-wise forward formula
            return ctx.lookup_size / ctx.device_bw
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Type, TypeVar, Union

__all__ = [
    "hbm_mem_bw", "ddr_mem_bw", "ssd_mem_bw", "hbm_to_ddr_mem_bw", "intra_host_bw", "inter_host_bw", "device_bw", "use_min_dim_for_lookup", "use_block_usage_penalty",
    "use_bytes_for_input_read_size", "input_data_type_size", "supported_sharding_types", "forward_compute", "backward_compute", "prefetch_compute", "input_dist_comms",
    "fwd_comms", "bwd_comms", "output_write_size", "fwd_coefficient", "bwd_coefficient", "prefetch_coefficient", "get_custom_method", "get_forward_compute",
    "get_backward_compute", "get_prefetch_compute", "get_input_dist_comms", "get_fwd_comms", "get_bwd_comms", "get_output_write_size", "get_fwd_coefficient",
    "get_bwd_coefficient", "get_prefetch_coefficient",
]

T = TypeVar("T")
_ROLE = "_estimator_role"
_TYPES = "_estimator_sharding_types"


# ---- class decorators: hardware numbers -------------------------------------------------------------------------------
def _set(attr: str, value: Any) -> Callable[[Type[T]], Type[T]]:
    def deco(cls: Type[T]) -> Type[T]:
        setattr(cls, attr, value)
        return cls

    return deco


def hbm_mem_bw(value: float) -> Callable[[Type[T]], Type[T]]:
    return _set("hbm_mem_bw", value)


def ddr_mem_bw(value: float) -> Callable[[Type[T]], Type[T]]:
    return _set("ddr_mem_bw", value)


def ssd_mem_bw(value: float) -> Callable[[Type[T]], Type[T]]:
    return _set("ssd_mem_bw", value)


def hbm_to_ddr_mem_bw(value: float) -> Callable[[Type[T]], Type[T]]:
    return _set("hbm_to_ddr_mem_bw", value)


def intra_host_bw(value: float) -> Callable[[Type[T]], Type[T]]:
    return _set("intra_host_bw", value)


def inter_host_bw(value: float) -> Callable[[Type[T]], Type[T]]:
    return _set("inter_host_bw", value)


def device_bw(compute_device: str, compute_kernel: str, value: Union[float, Callable[..., float]]) -> Callable[[Type[T]], Type[T]]:
    """Pin the effective lookup bandwidth of one (device, compute kernel) pair; ``value`` may be ``f(config, caching_ratio, prefetch_pipeline)``."""

    def deco(cls: Type[T]) -> Type[T]:
        table = dict(getattr(cls, "_device_bw_overrides", {}))
        table[(compute_device, compute_kernel)] = value
        cls._device_bw_overrides = table  # type: ignore[attr-defined]
        return cls

    return deco


def use_min_dim_for_lookup(value: bool = True) -> Callable[[Type[T]], Type[T]]:
    return _set("use_min_dim_for_lookup", value)


def use_block_usage_penalty(value: bool = True) -> Callable[[Type[T]], Type[T]]:
    return _set("use_block_usage_penalty", value)


def use_bytes_for_input_read_size(value: bool = True) -> Callable[[Type[T]], Type[T]]:
    return _set("use_bytes_for_input_read_size", value)


def input_data_type_size(value: float) -> Callable[[Type[T]], Type[T]]:
    return _set("input_data_type_size", value)


def supported_sharding_types(*sharding_types: str) -> Callable[[Type[T]], Type[T]]:
    return _set("supported_sharding_types", list(sharding_types))


# ---- method decorators: formulas ------------------------------------------------------------------------------------------
def _formula(role: str):
    def decorator(fn: Optional[Callable[..., Any]] = None, *, sharding_type: Union[None, str, List[str]] = None):
        def mark(f: Callable[..., Any]) -> Callable[..., Any]:
            setattr(f, _ROLE, role)
            setattr(f, _TYPES, None if sharding_type is None else ([sharding_type] if isinstance(sharding_type, str) else list(sharding_type)))
            return f

        return mark(fn) if callable(fn) else mark

    decorator.__name__ = role
    return decorator


forward_compute = _formula("forward_compute")
backward_compute = _formula("backward_compute")
prefetch_compute = _formula("prefetch_compute")
input_dist_comms = _formula("input_dist_comms")
fwd_comms = _formula("fwd_comms")
bwd_comms = _formula("bwd_comms")
output_write_size = _formula("output_write_size")
fwd_coefficient = _formula("fwd_coefficient")
bwd_coefficient = _formula("bwd_coefficient")
prefetch_coefficient = _formula("prefetch_coefficient")


def _matches_sharding_type(method: Callable[..., Any], sharding_type: str) -> bool:
    types = getattr(method, _TYPES, None)
    return types is None or sharding_type in types


def _get_annotated_method(config: Any, role: str, sharding_type: str) -> Optional[Callable[..., Any]]:
    """The config's method for ``role``: a type-specific one wins over a catch-all; None when the config defines neither."""
    specific = generic = None
    for name in dir(type(config)):
        m = getattr(type(config), name, None)
        if callable(m) and getattr(m, _ROLE, None) == role:
            if getattr(m, _TYPES, None) is None:
                generic = generic or m
            elif _matches_sharding_type(m, sharding_type):
                specific = specific or m
    chosen = specific or generic
    return None if chosen is None else chosen.__get__(config, type(config))


def get_custom_method(config: Any, role: str, sharding_type: str) -> Optional[Callable[..., Any]]:
    return _get_annotated_method(config, role, sharding_type)


def get_forward_compute(config: Any, sharding_type: str) -> Optional[Callable[..., float]]:
    return _get_annotated_method(config, "forward_compute", sharding_type)


def get_backward_compute(config: Any, sharding_type: str) -> Optional[Callable[..., float]]:
    return _get_annotated_method(config, "backward_compute", sharding_type)


def get_prefetch_compute(config: Any, sharding_type: str = "") -> Optional[Callable[..., float]]:
    return _get_annotated_method(config, "prefetch_compute", sharding_type)


def get_input_dist_comms(config: Any, sharding_type: str) -> Optional[Callable[..., float]]:
    return _get_annotated_method(config, "input_dist_comms", sharding_type)


def get_fwd_comms(config: Any, sharding_type: str) -> Optional[Callable[..., float]]:
    return _get_annotated_method(config, "fwd_comms", sharding_type)


def get_bwd_comms(config: Any, sharding_type: str) -> Optional[Callable[..., float]]:
    return _get_annotated_method(config, "bwd_comms", sharding_type)


def get_output_write_size(config: Any, sharding_type: str) -> Optional[Callable[..., float]]:
    return _get_annotated_method(config, "output_write_size", sharding_type)


def get_fwd_coefficient(config: Any, sharding_type: str) -> Optional[Callable[..., Any]]:
    return _get_annotated_method(config, "fwd_coefficient", sharding_type)


def get_bwd_coefficient(config: Any, sharding_type: str) -> Optional[Callable[..., Any]]:
    return _get_annotated_method(config, "bwd_coefficient", sharding_type)


def get_prefetch_coefficient(config: Any, sharding_type: str = "") -> Optional[Callable[..., Any]]:
    return _get_annotated_method(config, "prefetch_coefficient", sharding_type)
