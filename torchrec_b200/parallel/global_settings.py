"""Process-wide switches (reference distributed/global_settings.py:13-32)."""
PROPOGATE_DEVICE: bool = False


def set_propogate_device(val: bool) -> None:
    """When set, sharded inference modules keep the device of their inputs instead of forcing the shard device (name kept as in
    the reference, typo included)."""
    global PROPOGATE_DEVICE
    PROPOGATE_DEVICE = bool(val)


def get_propogate_device() -> bool:
    return PROPOGATE_DEVICE
