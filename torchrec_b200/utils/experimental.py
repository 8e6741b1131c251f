"""``@experimental`` marker for unstable APIs (reference torchrec/utils/experimental.py:20-88): warns once per decorated object."""
from __future__ import annotations

import functools
import warnings
from typing import Any, Callable, Optional, TypeVar, Union

__all__ = ["experimental"]
T = TypeVar("T")
_WARNED: set = set()


def experimental(obj: Optional[Union[Callable[..., Any], type]] = None, *, feature: Optional[str] = None, since: Optional[str] = None):
    """Decorate a function or class whose API may change; the first call / instantiation emits a ``UserWarning``."""

    def decorator(target: Any) -> Any:
        name = feature or getattr(target, "__qualname__", str(target))
        msg = f"`{name}` is *experimental*" + (f" (since {since})" if since else "") + " and may change or be removed without notice."

        def warn_once() -> None:
            if name not in _WARNED:
                _WARNED.add(name)
                warnings.warn(msg, UserWarning, stacklevel=3)

        if isinstance(target, type):
            orig_init = target.__init__

            @functools.wraps(orig_init)
            def new_init(self, *args: Any, **kwargs: Any) -> None:
                warn_once()
                orig_init(self, *args, **kwargs)

            target.__init__ = new_init  # type: ignore[method-assign]
            return target

        @functools.wraps(target)
        def wrapper(*args: Any, **kwargs: Any) -> Any:
            warn_once()
            return target(*args, **kwargs)

        return wrapper

    return decorator(obj) if obj is not None else decorator
