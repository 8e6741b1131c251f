from .delta_store import DeltaStore, DeltaStoreTrec, RawIdTrackerStore, compute_unique_rows  # noqa: F401
from .model_delta_tracker import ModelDeltaTracker as _AbstractTracker, ModelDeltaTrackerTrec, UPDATE_MODE_MAP  # noqa: F401
from .types import (DeltaTrackerConfig, IndexedLookup, ModelTrackerConfigs, RawIdTrackerConfig, RawIndexedLookup, Trackers, TrackingMode, UniqueRows,  # noqa: F401
                    UpdateMode)

# DistributedModelParallel instantiates ``ModelDeltaTracker(model, **config)``: the concrete tracker
ModelDeltaTracker = ModelDeltaTrackerTrec
