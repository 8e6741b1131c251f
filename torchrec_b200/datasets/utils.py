"""Dataset utilities (reference torchrec/datasets/utils.py:28)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Iterable, Iterator, List, Optional, Tuple

import torch

from ..sparse.jagged_tensor import KeyedJaggedTensor
from ..streamable import Pipelineable


@dataclass
class Batch(Pipelineable):
    """One training batch: dense float features [B, n_dense], sparse KJT, labels [B]."""

    dense_features: torch.Tensor
    sparse_features: KeyedJaggedTensor
    labels: torch.Tensor

    def to(self, device: torch.device, non_blocking: bool = False) -> "Batch":
        return Batch(
            dense_features=self.dense_features.to(device=device, non_blocking=non_blocking),
            sparse_features=self.sparse_features.to(device=device, non_blocking=non_blocking),
            labels=self.labels.to(device=device, non_blocking=non_blocking),
        )

    def record_stream(self, stream: torch.Stream) -> None:
        if self.dense_features.is_cuda:
            self.dense_features.record_stream(stream)
            self.labels.record_stream(stream)
        self.sparse_features.record_stream(stream)

    def pin_memory(self) -> "Batch":
        return Batch(
            dense_features=self.dense_features.pin_memory(),
            sparse_features=self.sparse_features.pin_memory(),
            labels=self.labels.pin_memory(),
        )

    def nbytes(self) -> int:
        n = self.dense_features.numel() * self.dense_features.element_size() + self.labels.numel() * self.labels.element_size()
        kjt = self.sparse_features
        for t in (kjt.values(), kjt.lengths_or_none(), kjt.offsets_or_none(), kjt.weights_or_none()):
            if t is not None:
                n += t.numel() * t.element_size()
        return n


class Limit:
    """Iterate at most ``limit`` items of an iterable."""

    def __init__(self, datapipe: Iterable, limit: int) -> None:
        self.datapipe = datapipe
        self.limit = limit

    def __iter__(self) -> Iterator:
        for i, x in enumerate(self.datapipe):
            if i >= self.limit:
                break
            yield x


def safe_cast(val: Any, dest_type: Callable, default: Any) -> Any:
    try:
        return dest_type(val)
    except (ValueError, TypeError):
        return default


def rand_split_train_val(datapipe: Iterable, train_perc: float, random_seed: int = 0) -> Tuple[Iterable, Iterable]:
    """Deterministic pseudo-random split of a datapipe into train / val streams."""
    if not 0.0 < train_perc < 1.0:
        raise ValueError("train_perc must be in (0, 1)")

    class _Split:
        def __init__(self, train: bool) -> None:
            self.train = train

        def __iter__(self):
            import random

            rng = random.Random(random_seed)
            for x in datapipe:
                is_train = rng.random() < train_perc
                if is_train == self.train:
                    yield x

    return _Split(True), _Split(False)


def idx_split_train_val(datapipe: Iterable, train_perc: float, decimal_places: int = 3, key_fn: Callable = lambda x: x):
    if not 0.0 < train_perc < 1.0:
        raise ValueError("train_perc must be in (0, 1)")
    shift = 10**decimal_places
    thr = int(train_perc * shift)

    class _Split:
        def __init__(self, train: bool) -> None:
            self.train = train

        def __iter__(self):
            for x in datapipe:
                if ((key_fn(x) % shift) < thr) == self.train:
                    yield x

    return _Split(True), _Split(False)


class LoadFiles:
    """(path, open file) pairs of the given paths."""

    def __init__(self, datapipe: Iterable[str], mode: str = "b", length: int = -1, **open_kw: Any) -> None:
        self.datapipe, self.mode, self.open_kw = datapipe, mode, open_kw

    def __iter__(self) -> Iterator[Tuple[str, Any]]:
        for path in self.datapipe:
            mode = self.mode if "r" in self.mode else "r" + self.mode
            with open(path, mode, **self.open_kw) as f:
                yield path, f


class ReadLinesFromCSV:
    """Rows (list of str) of the CSV/TSV files yielded by ``LoadFiles``."""

    def __init__(self, datapipe: Iterable[Tuple[str, Any]], skip_first_line: bool = False, **fmtparams: Any) -> None:
        self.datapipe, self.skip_first_line, self.fmtparams = datapipe, skip_first_line, fmtparams

    def __iter__(self) -> Iterator[List[str]]:
        import csv

        for _, f in self.datapipe:
            rd = csv.reader(f, **self.fmtparams)
            if self.skip_first_line:
                next(rd, None)
            yield from rd


class ParallelReadConcat:
    """Concatenate datapipes, giving every DataLoader worker (and rank) its own subset of them."""

    def __init__(self, *datapipes: Iterable, dp_selector: Optional[Callable] = None) -> None:
        self.datapipes = datapipes
        self.dp_selector = dp_selector or _default_dp_selector

    def __iter__(self) -> Iterator[Any]:
        for dp in self.dp_selector(self.datapipes):
            yield from dp


def _default_dp_selector(datapipes):
    from torch.utils.data import get_worker_info

    info = get_worker_info()
    if info is None or info.num_workers <= 1:
        return datapipes
    if info.num_workers > len(datapipes):
        raise ValueError(f"Number of workers {info.num_workers} exceeds number of datapipes ({len(datapipes)})!")
    return datapipes[info.id :: info.num_workers]


def train_filter(key_fn: Callable[[int], int], train_perc: float, decimal_places: int, idx: int) -> bool:
    shift = 10**decimal_places
    return (key_fn(idx) % shift) < round(train_perc * shift)


def val_filter(key_fn: Callable[[int], int], train_perc: float, decimal_places: int, idx: int) -> bool:
    return not train_filter(key_fn, train_perc, decimal_places, idx)
