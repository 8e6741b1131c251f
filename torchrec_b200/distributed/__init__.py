"""``torchrec_b200.distributed`` - the reference's name for ``torchrec_b200.parallel``.

A user of the reference writes ``from torchrec.distributed.model_parallel import DistributedModelParallel``; here the same line with ``torchrec_b200``
works: this package installs an import hook that resolves ``torchrec_b200.distributed[.x.y]`` to THE SAME module object as
``torchrec_b200.parallel[.x.y]`` (an alias, not a second copy - classes keep one identity, ``isinstance`` works across both spellings).
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import sys

_SRC = "torchrec_b200.parallel"
_DST = __name__  # torchrec_b200.distributed


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target: str) -> None:
        self._target = target

    def create_module(self, spec):  # noqa: ANN001
        mod = importlib.import_module(self._target)
        self._real_spec = mod.__spec__
        return mod

    def exec_module(self, module) -> None:  # noqa: ANN001
        # already executed under its real name; the import machinery just stamped the ALIAS spec on it - put the real one back so the module keeps
        # one identity (``__spec__.parent == __package__``, relative imports and pickling by qualified name keep working)
        module.__spec__ = self._real_spec


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):  # noqa: ANN001
        if not fullname.startswith(_DST + "."):
            return None
        real = _SRC + fullname[len(_DST):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except ModuleNotFoundError:
            return None
        return importlib.util.spec_from_loader(fullname, _AliasLoader(real))


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())

_parallel = importlib.import_module(_SRC)
# the package itself: same attributes as ``parallel`` (top-level re-exports) plus lazy sub-modules through the finder
from ..parallel import *  # noqa: F401,F403,E402


def __getattr__(name: str):  # sub-modules and late attributes of ``parallel``
    try:
        return getattr(_parallel, name)
    except AttributeError:
        try:
            return importlib.import_module(f"{_DST}.{name}")
        except ModuleNotFoundError:
            raise AttributeError(f"module {_DST!r} has no attribute {name!r}") from None
