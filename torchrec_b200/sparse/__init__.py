from .jagged_tensor import (  # noqa: F401
    ComputeJTDictToKJT,
    ComputeKJTToJTDict,
    JaggedTensor,
    KeyedJaggedTensor,
    KeyedTensor,
    flatten_kjt_list,
    kjt_is_equal,
    permute_multi_embedding,
    regroup_kts,
    unflatten_kjt_list,
)
