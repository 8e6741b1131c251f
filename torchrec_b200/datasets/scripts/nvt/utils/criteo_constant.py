"""Column names and table sizes of the Criteo click logs (reference ``datasets/scripts/nvt/utils/criteo_constant.py``)."""
from typing import Dict, List

FREQUENCY_THRESHOLD = 3
INT_FEATURE_COUNT = 13
CAT_FEATURE_COUNT = 26
DAYS = 24
DEFAULT_LABEL_NAME = "label"
DEFAULT_INT_NAMES: List[str] = [f"int_{idx}" for idx in range(INT_FEATURE_COUNT)]
DEFAULT_CAT_NAMES: List[str] = [f"cat_{idx}" for idx in range(CAT_FEATURE_COUNT)]
DEFAULT_COLUMN_NAMES: List[str] = [DEFAULT_LABEL_NAME, *DEFAULT_INT_NAMES, *DEFAULT_CAT_NAMES]
# MLPerf DLRM v2 table sizes (40 M cap)
NUM_EMBEDDINGS_PER_FEATURE: List[int] = [40000000, 39060, 17295, 7424, 20265, 3, 7122, 1543, 63, 40000000, 3067956, 405282, 10, 2209, 11938, 155, 4, 976, 14, 40000000,
                                         40000000, 40000000, 590152, 12973, 108, 36]
NUM_EMBEDDINGS_PER_FEATURE_DICT: Dict[str, int] = dict(zip(DEFAULT_CAT_NAMES, NUM_EMBEDDINGS_PER_FEATURE))
