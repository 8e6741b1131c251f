from .object_pools import KeyedJaggedTensorPool  # noqa: F401
