"""Dynamic embedding: unbounded id spaces over a fixed-size HBM cache table backed by a parameter server
(reference torchrec/csrc/dynamic_embedding + contrib/dynamic_embedding). Native core in csrc/dynemb."""
from .dataloader import DataLoader, wrap  # noqa: F401
from .id_transformer import IDTransformer  # noqa: F401
from .id_transformer_collection import IDTransformerCollection  # noqa: F401
from .ps import PS, load_io_plugin  # noqa: F401
from .id_transformer_group import IDTransformerGroup  # noqa: F401,E402
from .tensor_list import TensorList  # noqa: F401,E402
