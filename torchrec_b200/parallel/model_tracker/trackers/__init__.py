from .raw_id_tracker import RawIdTracker  # noqa: F401
