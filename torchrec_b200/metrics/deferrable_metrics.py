"""A metrics dict whose values may still be in flight (reference metrics/deferrable_metrics.py:31-234).

``RecMetricModule.compute`` on the training thread returns immediately with a ``DeferrableMetrics`` backed by a future (device -> host
copies and the cross-rank sync run elsewhere); consumers either ``subscribe`` a callback or ``resolve()`` (blocking) when they need the
numbers. Behaves like a read-only mapping once resolved; synchronous access before that resolves on the spot (and says so once)."""
from __future__ import annotations

import logging
from concurrent.futures import Future
from typing import Any, Callable, Dict, Iterator, Mapping, Optional, Union

import torch

logger: logging.Logger = logging.getLogger(__name__)


def device_supports_async(device: torch.device) -> bool:
    return torch.device(device).type == "cuda"


def transfer_tensors_to_cpu(tensors: Dict[str, Any], non_blocking: bool = True) -> "tuple[Dict[str, Any], Optional[torch.cuda.Event]]":
    """Start D2H copies of every CUDA tensor into pinned buffers; returns (host dict, event to wait on or None)."""
    values = tensors
    out: Dict[str, Any] = {}
    any_cuda = False
    for k, v in values.items():
        if isinstance(v, torch.Tensor) and v.is_cuda:
            buf = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            buf.copy_(v.detach(), non_blocking=non_blocking)
            out[k] = buf
            any_cuda = True
        else:
            out[k] = v
    ev = None
    if any_cuda:
        ev = torch.cuda.Event()
        ev.record()
    return out, ev


class DeferrableMetrics(Mapping):
    def __init__(self, inner: Optional[Union[Dict[str, Any], "Future[Dict[str, Any]]"]] = None, values: Optional[Union[Dict[str, Any], "Future[Dict[str, Any]]"]] = None) -> None:
        values = inner if inner is not None else values  # ``inner``: the reference's name of the argument
        self._values: Dict[str, Any] = {}
        self._future: Optional["Future[Dict[str, Any]]"] = None
        self._warned = False
        if isinstance(values, Future):
            self._future = values
        elif values:
            self._values.update(values)

    # ---- asynchronous consumption -------------------------------------------------------------------------------------------------
    def subscribe(self, callback: Callable[[Dict[str, Any]], None], on_error: Optional[Callable[[Exception], None]] = None) -> None:
        """``callback(metrics)`` once everything is available (immediately if it already is)."""
        if self._future is None:
            callback(dict(self._values))
            return

        def _on_complete(f: "Future[Dict[str, Any]]") -> None:
            try:
                res = f.result()
            except Exception as e:  # noqa: BLE001
                if on_error is not None:
                    on_error(e)
                else:
                    logger.exception("deferred metric computation failed")
                return
            merged = dict(res)
            merged.update(self._values)
            callback(merged)

        self._future.add_done_callback(_on_complete)

    def resolve(self, timeout: Optional[float] = None) -> Dict[str, Any]:
        if self._future is not None:
            res = self._future.result(timeout=timeout)
            merged = dict(res)
            merged.update(self._values)
            self._values, self._future = merged, None
        return dict(self._values)

    def is_resolved(self) -> bool:
        return self._future is None or self._future.done()

    def update(self, other: Union[Dict[str, Any], "DeferrableMetrics"]) -> None:  # type: ignore[override]
        if isinstance(other, DeferrableMetrics):
            other = other.resolve() if other._future is None or other._future.done() else other
        if isinstance(other, DeferrableMetrics):  # still pending: chain the futures
            mine, theirs = self._future, other._future
            chained: "Future[Dict[str, Any]]" = Future()

            def _finish(_f: Any = None) -> None:
                if chained.done():
                    return
                try:
                    a = mine.result() if mine is not None else {}
                    b = theirs.result() if theirs is not None else {}
                    if (mine is None or mine.done()) and (theirs is None or theirs.done()):
                        merged = dict(a)
                        merged.update(b)
                        merged.update(other._values)
                        chained.set_result(merged)
                except Exception as e:  # noqa: BLE001
                    chained.set_exception(e)

            for f in (mine, theirs):
                if f is not None:
                    f.add_done_callback(_finish)
            self._future = chained
        else:
            self._values.update(other)

    # ---- mapping protocol (synchronous access resolves) --------------------------------------------------------------------------------
    def _sync(self) -> Dict[str, Any]:
        if self._future is not None:
            if not self._warned and not self._future.done():
                logger.warning("DeferrableMetrics accessed synchronously before completion: blocking on the metric computation")
                self._warned = True
            self.resolve()
        return self._values

    def __bool__(self) -> bool:
        return self._future is not None or bool(self._values)

    def __setitem__(self, key: str, value: Any) -> None:
        self._values[key] = value

    def __getitem__(self, key: str) -> Any:
        return self._sync()[key]

    def __iter__(self) -> Iterator[str]:
        return iter(self._sync())

    def __len__(self) -> int:
        return len(self._sync())

    def __repr__(self) -> str:
        state = "resolved" if self.is_resolved() else "pending"
        return f"DeferrableMetrics({state}, {len(self._values)} local values)"
