"""Re-export at the reference's path (torchrec/utils/percentile_logger.py)."""
from ..parallel.logger import PercentileLogger  # noqa: F401
