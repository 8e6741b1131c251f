"""Synthetic Criteo files for tests (reference datasets/test_utils/criteo_test_utils.py:23-155): tiny TSV days and npy triplets with the
real column layout (label, 13 integer features, 26 hex categorical features)."""
from __future__ import annotations

import contextlib
import csv
import os
import random
import tempfile
from typing import Any, Dict, Generator, List, Optional

import numpy as np

from ..criteo import CAT_FEATURE_COUNT, INT_FEATURE_COUNT


class CriteoTest:
    """Mix-in with the helper generators (usable from pytest functions or unittest classes)."""

    INT_FEATURE_COUNT = INT_FEATURE_COUNT
    CAT_FEATURE_COUNT = CAT_FEATURE_COUNT
    LABEL_VAL_RANGE = (0, 1)
    INT_VAL_RANGE = (0, 100)
    CAT_VAL_RANGE = (0, 1000)

    @classmethod
    @contextlib.contextmanager
    def _create_dataset_tsv(cls, num_rows: int = 10000, train: bool = True, filename: str = "criteo", seed: int = 0) -> Generator[str, None, None]:
        rng = random.Random(seed)
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, filename)
            with open(path, "w", newline="") as f:
                w = csv.writer(f, delimiter="\t")
                for _ in range(num_rows):
                    row: List[Any] = []
                    if train:
                        row.append(rng.randint(*cls.LABEL_VAL_RANGE))
                    row += [rng.randint(*cls.INT_VAL_RANGE) if rng.random() > 0.05 else "" for _ in range(cls.INT_FEATURE_COUNT)]
                    row += [("%08x" % rng.randint(*cls.CAT_VAL_RANGE)) if rng.random() > 0.05 else "" for _ in range(cls.CAT_FEATURE_COUNT)]
                    w.writerow(row)
            yield path

    def _validate_sample(self, sample: Dict[str, Any], train: bool = True) -> None:
        if train:
            assert sample["label"] in range(self.LABEL_VAL_RANGE[0], self.LABEL_VAL_RANGE[1] + 1)
        for i in range(self.INT_FEATURE_COUNT):
            v = sample[f"int_{i}"]
            assert v == 0 or self.INT_VAL_RANGE[0] <= v <= self.INT_VAL_RANGE[1]
        for i in range(self.CAT_FEATURE_COUNT):
            v = sample[f"cat_{i}"]
            assert v == "" or (isinstance(v, str) and len(v) == 8) or isinstance(v, int)

    @classmethod
    @contextlib.contextmanager
    def _create_dataset_npys(cls, num_rows: int = 10000, filenames: Optional[List[str]] = None, generate_dense: bool = True, generate_sparse: bool = True,
                             generate_labels: bool = True, dense: Optional[np.ndarray] = None, sparse: Optional[np.ndarray] = None, labels: Optional[np.ndarray] = None,
                             seed: int = 0) -> Generator[List[str], None, None]:
        rng = np.random.default_rng(seed)
        filenames = filenames or ["criteo"]
        with tempfile.TemporaryDirectory() as tmp:
            paths: List[str] = []
            for name in filenames:
                base = os.path.join(tmp, name)
                if generate_dense:
                    np.save(base + "_dense.npy", dense if dense is not None else rng.random((num_rows, cls.INT_FEATURE_COUNT), dtype=np.float32))
                    paths.append(base + "_dense.npy")
                if generate_sparse:
                    np.save(base + "_sparse.npy", sparse if sparse is not None else rng.integers(cls.CAT_VAL_RANGE[0], cls.CAT_VAL_RANGE[1] + 1,
                                                                                              size=(num_rows, cls.CAT_FEATURE_COUNT), dtype=np.int32))
                    paths.append(base + "_sparse.npy")
                if generate_labels:
                    np.save(base + "_labels.npy", labels if labels is not None else rng.integers(0, 2, size=(num_rows, 1), dtype=np.int32))
                    paths.append(base + "_labels.npy")
            yield paths
