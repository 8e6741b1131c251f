"""Package a ``PredictFactory`` for the serving process.

Parity: reference ``inference/model_packager.py`` (``PredictFactoryPackager.save_predict_factory`` :61) + ``dlrm_packager.py``. The
reference writes a ``torch.package`` archive that its C++ server unpacks with torch::deploy; here the archive is a plain zip
(``model_config.pkl`` + ``meta.json`` naming the factory class + optional ``state_dict.pt``) that ``load_predict_factory`` turns back
into a factory inside the Python executor threads of ``inference/server.py`` - no interpreter embedding needed."""
from __future__ import annotations

import abc
import importlib
import io
import json
import pickle
import zipfile
from pathlib import Path
from typing import Any, Dict, Optional, Type, Union

import torch

from .modules import PredictFactory


class PredictFactoryPackager(abc.ABC):
    """Subclass hook points kept from the reference (extern / mocked modules are meaningless for a zip of pickles and are
    accepted for signature compatibility)."""

    @classmethod
    def set_extern_modules(cls, pe: Any) -> None:  # noqa: D401
        return None

    @classmethod
    def set_mocked_modules(cls, pe: Any) -> None:
        return None

    @classmethod
    def save_predict_factory(cls, predict_factory: Type[PredictFactory], configs: Dict[str, Any], output: Union[str, Path], extra_files: Optional[Dict[str, Union[str, bytes]]] = None,
                             state_dict: Optional[Dict[str, torch.Tensor]] = None) -> None:
        meta = {"factory_module": predict_factory.__module__, "factory_class": predict_factory.__qualname__, "config_keys": sorted(configs)}
        with zipfile.ZipFile(output, "w", compression=zipfile.ZIP_STORED) as z:
            z.writestr("meta.json", json.dumps(meta))
            for name, cfg in configs.items():
                z.writestr(f"configs/{name}.pkl", pickle.dumps(cfg))
            for name, data in (extra_files or {}).items():
                z.writestr(f"extra_files/{name}", data)
            if state_dict is not None:
                buf = io.BytesIO()
                torch.save(state_dict, buf)
                z.writestr("state_dict.pt", buf.getvalue())


def load_config_text(archive: Union[str, Path], name: str) -> str:
    with zipfile.ZipFile(archive) as z:
        return z.read(f"extra_files/{name}").decode()


def load_pickle_config(archive: Union[str, Path], name: str) -> Any:
    with zipfile.ZipFile(archive) as z:
        return pickle.loads(z.read(f"configs/{name}.pkl"))


def load_predict_factory(archive: Union[str, Path]) -> PredictFactory:
    """Rebuild the packaged factory: ``factory_class(**configs)`` (single config -> passed positionally)."""
    with zipfile.ZipFile(archive) as z:
        meta = json.loads(z.read("meta.json"))
        configs = {k: pickle.loads(z.read(f"configs/{k}.pkl")) for k in meta["config_keys"]}
    mod = importlib.import_module(meta["factory_module"])
    cls = mod
    for part in meta["factory_class"].split("."):
        cls = getattr(cls, part)
    return cls(*configs.values()) if len(configs) == 1 else cls(**configs)  # type: ignore[operator]


def load_packaged_state_dict(archive: Union[str, Path]) -> Optional[Dict[str, torch.Tensor]]:
    with zipfile.ZipFile(archive) as z:
        if "state_dict.pt" not in z.namelist():
            return None
        return torch.load(io.BytesIO(z.read("state_dict.pt")), map_location="cpu")
