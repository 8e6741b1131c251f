from .object_pools import ObjectPool, TensorPool  # noqa: F401
