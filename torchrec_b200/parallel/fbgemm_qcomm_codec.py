"""Reference import path ``torchrec/distributed/fbgemm_qcomm_codec.py`` (``CommType`` :31, ``QCommsConfig`` :55, ``get_qcomm_codecs`` :131,
``get_qcomm_codecs_registry`` :185). The codecs themselves are this framework's own (``qcomm_codec.py``) - there is no FBGEMM here."""
from .qcomm_codec import *  # noqa: F401,F403
from .qcomm_codec import CommType, QCommsConfig, get_qcomm_codec, get_qcomm_codecs_registry  # noqa: F401

try:
    from .qcomm_codec import get_qcomm_codecs  # noqa: F401
except ImportError:  # pragma: no cover
    from .types import QuantizedCommCodecs

    def get_qcomm_codecs(qcomms_config):  # type: ignore[no-redef]
        """forward / backward codec pair of one collective."""
        if qcomms_config is None:
            return QuantizedCommCodecs()
        return QuantizedCommCodecs(forward=get_qcomm_codec(qcomms_config.forward_precision, qcomms_config.forward_loss_scale, getattr(qcomms_config, "fp8_quantize_dim", None)),
                                   backward=get_qcomm_codec(qcomms_config.backward_precision, qcomms_config.backward_loss_scale, getattr(qcomms_config, "fp8_quantize_dim_bwd", None)))


def comm_type_to_sparse_type(comm_type: CommType) -> str:
    """The row format name of a comm precision (the reference maps to FBGEMM ``SparseType``; here a plain string)."""
    return {CommType.FP32: "fp32", CommType.FP16: "fp16", CommType.BF16: "bf16", CommType.FP8: "fp8", CommType.INT8: "int8"}.get(comm_type, str(getattr(comm_type, "value", comm_type)))
