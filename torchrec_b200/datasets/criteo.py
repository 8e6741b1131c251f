"""Criteo click-logs datasets: TSV datapipe, binary (npy) preprocessing utilities and the in-memory binary loader
(reference torchrec/datasets/criteo.py:59-1020, scripts/npy_preproc_criteo.py).

The binary path is the production one: per-day ``*_dense.npy`` (float32 [N,13], log(x+3) transformed), ``*_sparse.npy``
(int32/int64 [N,26], hashed or contiguous ids) and ``*_labels.npy`` (int32 [N,1]). ``InMemoryBinaryCriteoIterDataPipe``
assigns every rank a contiguous row range over the days and emits ready ``Batch`` objects (one id per feature -> lengths are
all ones and are built once)."""
from __future__ import annotations

import math
import os
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Tuple, Union

import numpy as np
import torch
from torch.utils.data import IterableDataset

from ..sparse.jagged_tensor import KeyedJaggedTensor
from .utils import Batch, LoadFiles, ReadLinesFromCSV, safe_cast

FREQUENCY_THRESHOLD = 3
INT_FEATURE_COUNT = 13
CAT_FEATURE_COUNT = 26
DAYS = 24
DEFAULT_LABEL_NAME = "label"
DEFAULT_INT_NAMES: List[str] = [f"int_{idx}" for idx in range(INT_FEATURE_COUNT)]
DEFAULT_CAT_NAMES: List[str] = [f"cat_{idx}" for idx in range(CAT_FEATURE_COUNT)]
DEFAULT_COLUMN_NAMES: List[str] = [DEFAULT_LABEL_NAME, *DEFAULT_INT_NAMES, *DEFAULT_CAT_NAMES]
TOTAL_TRAINING_SAMPLES = 4195197692  # days 0-22 of Criteo 1TB
COLUMN_TYPE_CASTERS: List[Callable[[Union[int, str]], Union[int, str]]] = [
    lambda val: safe_cast(val, int, 0),
    *(lambda val: safe_cast(val, int, 0) for _ in range(INT_FEATURE_COUNT)),
    *(lambda val: safe_cast(val, str, "") for _ in range(CAT_FEATURE_COUNT)),
]
# Criteo 1TB per-feature cardinalities (MLPerf DLRM, days 0-23)
CRITEO_1TB_CARDINALITIES: List[int] = [
    45833188, 36746, 17245, 7413, 20243, 3, 7114, 1441, 62, 29275261, 1572176, 345138, 10, 2209, 11267, 128, 4, 974, 14, 48937457, 11316796, 40094537, 452104,
    12606, 104, 35]


def _default_row_mapper(example: List[str]) -> Dict[str, Union[int, str]]:
    return {DEFAULT_COLUMN_NAMES[idx]: COLUMN_TYPE_CASTERS[idx](val) for idx, val in enumerate(example)}


class CriteoIterDataPipe:
    """Iterate the rows of Criteo TSV files as dicts (``row_mapper`` may change the representation)."""

    def __init__(self, paths: Iterable[str], *, row_mapper: Optional[Callable[[List[str]], Any]] = _default_row_mapper, **open_kw: Any) -> None:
        self.paths = paths
        self.row_mapper = row_mapper
        self.open_kw = open_kw

    def __iter__(self) -> Iterator[Any]:
        datapipe = ReadLinesFromCSV(LoadFiles(self.paths, mode="r", **self.open_kw), delimiter="\t")
        for row in datapipe:
            yield self.row_mapper(row) if self.row_mapper else row


def criteo_terabyte(paths: Iterable[str], *, row_mapper=_default_row_mapper, **open_kw: Any) -> CriteoIterDataPipe:
    return CriteoIterDataPipe(paths, row_mapper=row_mapper, **open_kw)


def criteo_kaggle(path: str, *, row_mapper=_default_row_mapper, **open_kw: Any) -> CriteoIterDataPipe:
    return CriteoIterDataPipe((path,), row_mapper=row_mapper, **open_kw)


class BinaryCriteoUtils:
    """TSV -> npy conversion, row-range bookkeeping and shuffling for the binary dataset."""

    @staticmethod
    def tsv_to_npys(in_file: str, out_dense_file: str, out_sparse_file: str, out_labels_file: str, dataset_name: str = "criteo_1tb", path_manager_key: str = "") -> None:
        """One TSV day -> three npy files. Dense: missing -> 0; sparse: hex string -> int32 (missing -> 0)."""
        dense, sparse, labels = [], [], []
        with open(in_file, "r") as f:
            for line in f:
                row = line.rstrip("\n").split("\t")
                if dataset_name == "criteo_kaggle" and len(row) == 1 + INT_FEATURE_COUNT + CAT_FEATURE_COUNT - 1:
                    row = ["0"] + row  # kaggle test split has no label
                labels.append(safe_cast(row[0], int, 0))
                dense.append([safe_cast(v, int, 0) for v in row[1 : 1 + INT_FEATURE_COUNT]])
                sparse.append([int(v, 16) if v else 0 for v in row[1 + INT_FEATURE_COUNT : 1 + INT_FEATURE_COUNT + CAT_FEATURE_COUNT]])
        np.save(out_dense_file, np.array(dense, dtype=np.float32).reshape(-1, INT_FEATURE_COUNT))
        np.save(out_sparse_file, (np.array(sparse, dtype=np.int64).reshape(-1, CAT_FEATURE_COUNT) & 0xFFFFFFFF).astype(np.int64).astype(np.uint32).view(np.int32))
        np.save(out_labels_file, np.array(labels, dtype=np.int32).reshape(-1, 1))

    @staticmethod
    def get_shape_from_npy(path: str, path_manager_key: str = "") -> Tuple[int, ...]:
        return tuple(np.load(path, mmap_mode="r").shape)

    @staticmethod
    def get_file_row_ranges_and_remainder(lengths: List[int], rank: int, world_size: int, start_row: int = 0, last_row: Optional[int] = None
                                          ) -> Tuple[Dict[int, Tuple[int, int]], int]:
        """Contiguous, near-equal row share of ``rank`` over the concatenated files: file idx -> (first row, last row) inclusive."""
        total = sum(lengths)
        last_row = total - 1 if last_row is None else last_row
        n = last_row - start_row + 1
        per, rem = divmod(n, world_size)
        lo = start_row + rank * per + min(rank, rem)
        hi = lo + per + (1 if rank < rem else 0) - 1
        out: Dict[int, Tuple[int, int]] = {}
        base = 0
        for i, ln in enumerate(lengths):
            f_lo, f_hi = max(lo, base), min(hi, base + ln - 1)
            if f_lo <= f_hi:
                out[i] = (f_lo - base, f_hi - base)
            base += ln
        return out, rem

    @staticmethod
    def load_npy_range(fname: str, start_row: int, num_rows: int, path_manager_key: str = "", mmap_mode: bool = False) -> np.ndarray:
        arr = np.load(fname, mmap_mode="r")
        if start_row + num_rows > arr.shape[0]:
            raise ValueError(f"requested rows [{start_row}, {start_row + num_rows}) of {fname} which has {arr.shape[0]} rows")
        sl = arr[start_row : start_row + num_rows]
        return sl if mmap_mode else np.ascontiguousarray(sl)

    @staticmethod
    def sparse_to_contiguous(in_files: List[str], output_dir: str, frequency_threshold: int = FREQUENCY_THRESHOLD, columns: int = CAT_FEATURE_COUNT,
                             path_manager_key: str = "", output_file_suffix: str = "_contig_freq.npy") -> None:
        """Re-index every sparse column to contiguous ids; ids seen fewer than ``frequency_threshold`` times map to 1, missing to 0."""
        counts: List[Dict[int, int]] = [dict() for _ in range(columns)]
        for f in in_files:
            arr = np.load(f)
            for c in range(columns):
                u, n = np.unique(arr[:, c], return_counts=True)
                d = counts[c]
                for k, v in zip(u.tolist(), n.tolist()):
                    d[k] = d.get(k, 0) + v
        maps: List[Dict[int, int]] = []
        for c in range(columns):
            m, nxt = {}, 2
            for k, v in counts[c].items():
                if v >= frequency_threshold:
                    m[k] = nxt
                    nxt += 1
            maps.append(m)
        os.makedirs(output_dir, exist_ok=True)
        for f in in_files:
            arr = np.load(f)
            out = np.empty_like(arr, dtype=np.int64)
            for c in range(columns):
                m = maps[c]
                out[:, c] = np.fromiter((m.get(int(v), 1) for v in arr[:, c]), dtype=np.int64, count=arr.shape[0])
            np.save(os.path.join(output_dir, os.path.basename(f).replace(".npy", "") + output_file_suffix), out)

    @staticmethod
    def shuffle(input_dir_labels_and_dense: str, input_dir_sparse: str, output_dir_shuffled: str, rows_per_day: Dict[int, int], output_dir_full_set: Optional[str] = None,
                days: int = DAYS, int_columns: int = INT_FEATURE_COUNT, sparse_columns: int = CAT_FEATURE_COUNT, path_manager_key: str = "", random_seed: int = 0) -> None:
        """Globally shuffle days [0, days-1) (the last day stays the in-order eval set) and re-split into days of the original sizes."""
        train_days = list(range(days - 1))
        dense = np.concatenate([np.load(os.path.join(input_dir_labels_and_dense, f"day_{d}_dense.npy")) for d in train_days])
        labels = np.concatenate([np.load(os.path.join(input_dir_labels_and_dense, f"day_{d}_labels.npy")) for d in train_days])
        sparse = np.concatenate([np.load(os.path.join(input_dir_sparse, f"day_{d}_sparse.npy")) for d in train_days])
        perm = np.random.default_rng(random_seed).permutation(dense.shape[0])
        os.makedirs(output_dir_shuffled, exist_ok=True)
        lo = 0
        for d in train_days:
            idx = perm[lo : lo + rows_per_day[d]]
            lo += rows_per_day[d]
            np.save(os.path.join(output_dir_shuffled, f"day_{d}_dense.npy"), dense[idx])
            np.save(os.path.join(output_dir_shuffled, f"day_{d}_sparse.npy"), sparse[idx])
            np.save(os.path.join(output_dir_shuffled, f"day_{d}_labels.npy"), labels[idx])


class InMemoryBinaryCriteoIterDataPipe(IterableDataset):
    """Rank-sharded in-memory loader of the preprocessed npy days -> ``Batch`` stream.

    Args mirror the reference: ``stage`` ("train"|"val"|"test": val/test split the LAST day in halves), dense/sparse/label
    file lists, ``batch_size`` (per rank), ``rank``/``world_size``, ``drop_last``, ``shuffle_batches``,
    ``shuffle_training_set`` (row-level), ``hashes`` (per-feature modulo), ``mmap_mode``."""

    def __init__(self, stage: str, dense_paths: List[str], sparse_paths: List[str], labels_paths: List[str], batch_size: int, rank: int, world_size: int,
                 drop_last: Optional[bool] = False, shuffle_batches: bool = False, shuffle_training_set: bool = False, shuffle_training_set_random_seed: int = 0,
                 mmap_mode: bool = False, hashes: Optional[List[int]] = None, path_manager_key: str = "", sparse_labels: Optional[List[str]] = None) -> None:
        self.stage, self.batch_size, self.rank, self.world_size = stage, batch_size, rank, world_size
        self.dense_paths, self.sparse_paths, self.labels_paths = dense_paths, sparse_paths, labels_paths
        self.drop_last, self.shuffle_batches, self.mmap_mode = drop_last, shuffle_batches, mmap_mode
        self.shuffle_training_set, self.seed = shuffle_training_set, shuffle_training_set_random_seed
        self.hashes = np.array(hashes, dtype=np.int64).reshape(1, -1) if hashes is not None else None
        self.keys: List[str] = sparse_labels or DEFAULT_CAT_NAMES
        self._load_data_for_rank()
        self.num_rows_per_file = [a.shape[0] for a in self.dense_arrs]
        n = sum(self.num_rows_per_file)
        self.num_batches = n // batch_size if drop_last else math.ceil(n / batch_size)
        self._lengths = torch.ones(batch_size * len(self.keys), dtype=torch.int32)
        self._offsets = torch.arange(0, batch_size * len(self.keys) + 1, dtype=torch.int32)
        self._lpk = [batch_size] * len(self.keys)

    def _load_data_for_rank(self) -> None:
        lengths = [BinaryCriteoUtils.get_shape_from_npy(p)[0] for p in self.dense_paths]
        start_row, last_row = 0, None
        if self.stage in ("val", "test"):
            # the last day is split into a validation half and a test half
            total, last = sum(lengths), lengths[-1]
            half = total - last + last // 2
            if self.stage == "val":
                start_row, last_row = total - last, half - 1
            else:
                start_row = half
        ranges, _ = BinaryCriteoUtils.get_file_row_ranges_and_remainder(lengths, self.rank, self.world_size, start_row, last_row)
        self.dense_arrs, self.sparse_arrs, self.labels_arrs = [], [], []
        for i, (lo, hi) in sorted(ranges.items()):
            n = hi - lo + 1
            self.dense_arrs.append(BinaryCriteoUtils.load_npy_range(self.dense_paths[i], lo, n, mmap_mode=self.mmap_mode))
            self.sparse_arrs.append(BinaryCriteoUtils.load_npy_range(self.sparse_paths[i], lo, n, mmap_mode=self.mmap_mode))
            self.labels_arrs.append(BinaryCriteoUtils.load_npy_range(self.labels_paths[i], lo, n, mmap_mode=self.mmap_mode))
        if not self.mmap_mode and self.hashes is not None:
            self.sparse_arrs = [a.astype(np.int64) % self.hashes for a in self.sparse_arrs]
        if self.shuffle_training_set and self.stage == "train" and self.dense_arrs:
            d, s, l = (np.concatenate(x) for x in (self.dense_arrs, self.sparse_arrs, self.labels_arrs))
            perm = np.random.default_rng(self.seed).permutation(d.shape[0])
            self.dense_arrs, self.sparse_arrs, self.labels_arrs = [d[perm]], [s[perm]], [l[perm]]

    def _np_arrays_to_batch(self, dense: np.ndarray, sparse: np.ndarray, labels: np.ndarray) -> Batch:
        if self.shuffle_batches:
            perm = torch.randperm(dense.shape[0]).numpy()
            dense, sparse, labels = dense[perm], sparse[perm], labels[perm]
        if self.mmap_mode and self.hashes is not None:
            sparse = sparse.astype(np.int64) % self.hashes
        B = dense.shape[0]
        full = B == self.batch_size
        F = len(self.keys)
        values = torch.from_numpy(np.array(sparse.T, dtype=np.int64).reshape(-1))
        kjt = KeyedJaggedTensor(keys=self.keys, values=values, lengths=self._lengths if full else torch.ones(B * F, dtype=torch.int32),
                                offsets=self._offsets if full else torch.arange(0, B * F + 1, dtype=torch.int32), stride=B,
                                length_per_key=self._lpk if full else [B] * F)
        return Batch(dense_features=torch.from_numpy(np.array(dense, dtype=np.float32)), sparse_features=kjt,
                     labels=torch.from_numpy(np.array(labels.reshape(-1))))

    def __iter__(self) -> Iterator[Batch]:
        buf: List[Tuple[np.ndarray, np.ndarray, np.ndarray]] = []
        have = 0

        def emit(n: int) -> Batch:
            nonlocal buf, have
            parts, take = [], n
            while take > 0:
                d, s, l = buf[0]
                if d.shape[0] <= take:
                    parts.append(buf.pop(0))
                    take -= d.shape[0]
                else:
                    parts.append((d[:take], s[:take], l[:take]))
                    buf[0] = (d[take:], s[take:], l[take:])
                    take = 0
            have -= n
            return self._np_arrays_to_batch(*(np.concatenate(x) if len(x) > 1 else x[0] for x in zip(*parts)))

        for d, s, l in zip(self.dense_arrs, self.sparse_arrs, self.labels_arrs):
            buf.append((d, s, l))
            have += d.shape[0]
            while have >= self.batch_size:
                yield emit(self.batch_size)
        if have > 0 and not self.drop_last:
            yield emit(have)

    def __len__(self) -> int:
        return self.num_batches
