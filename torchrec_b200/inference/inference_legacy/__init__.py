"""``torchrec.inference.inference_legacy`` import paths (the reference keeps its first-generation Python serving helpers there):
``client``, ``model_packager``, ``state_dict_transform`` and ``modules`` are the same module objects as ``torchrec_b200.inference.<name>``."""
import importlib
import sys

for _name in ("client", "model_packager", "state_dict_transform", "modules", "dlrm_predict", "dlrm_packager"):
    _mod = importlib.import_module(f"torchrec_b200.inference.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    globals()[_name] = _mod
