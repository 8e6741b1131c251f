"""Synthetic recsys data (reference torchrec/datasets/random.py:125).

``RandomRecDataset`` yields ``Batch`` objects with Criteo-shaped random data. Batches are
pre-generated with one vectorised ``torch.randint`` per tensor (no per-feature Python loop in
the steady state) and may be pinned for asynchronous H2D copies.
"""
from __future__ import annotations

import itertools
from typing import Iterator, List, Optional, Union

import torch
from torch.utils.data import IterableDataset

from ..sparse.jagged_tensor import KeyedJaggedTensor
from .utils import Batch


class _RandomRecBatch:
    def __init__(
        self,
        keys: List[str],
        batch_size: int,
        hash_sizes: List[int],
        ids_per_features: List[int],
        num_dense: int,
        manual_seed: Optional[int] = None,
        num_generated_batches: int = 10,
        num_batches: Optional[int] = None,
        min_ids_per_features: Optional[List[int]] = None,
        pin_memory: bool = False,
        index_dtype: torch.dtype = torch.int64,
        lengths_dtype: torch.dtype = torch.int32,
    ) -> None:
        self.keys = keys
        self.keys_length = len(keys)
        self.batch_size = batch_size
        self.hash_sizes = hash_sizes
        self.ids_per_features = ids_per_features
        self.min_ids_per_features = min_ids_per_features if min_ids_per_features is not None else ids_per_features
        self.num_dense = num_dense
        self.num_batches = num_batches
        self.num_generated_batches = num_generated_batches
        self.pin_memory = pin_memory
        self.index_dtype = index_dtype
        self.lengths_dtype = lengths_dtype
        self.generator = None
        if manual_seed is not None:
            self.generator = torch.Generator()
            self.generator.manual_seed(manual_seed)
        self._generated_batches: List[Batch] = [self._generate_batch() for _ in range(num_generated_batches)]
        self.batch_index = 0

    def __iter__(self) -> "_RandomRecBatch":
        self.batch_index = 0
        return self

    def __next__(self) -> Batch:
        if self.batch_index == self.num_batches:
            raise StopIteration
        batch = self._generated_batches[self.batch_index % len(self._generated_batches)] if self.num_generated_batches >= 0 else self._generate_batch()
        self.batch_index += 1
        return batch

    def _generate_batch(self) -> Batch:
        B, F = self.batch_size, self.keys_length
        g = self.generator
        lo = torch.tensor(self.min_ids_per_features, dtype=torch.int64).view(F, 1)
        hi = torch.tensor(self.ids_per_features, dtype=torch.int64).view(F, 1)
        if bool((lo == hi).all()):
            lengths2d = hi.expand(F, B).contiguous()
        else:
            u = torch.rand(F, B, generator=g)
            lengths2d = (lo + (u * (hi - lo + 1).float()).long()).clamp(max=hi)
        lengths = lengths2d.reshape(-1)
        per_key = lengths2d.sum(1)
        total = int(per_key.sum())
        hs = torch.repeat_interleave(torch.tensor(self.hash_sizes, dtype=torch.int64), per_key)
        values = (torch.rand(total, generator=g, dtype=torch.float64) * hs.double()).long()
        values = torch.minimum(values, hs - 1).to(self.index_dtype)
        sparse = KeyedJaggedTensor(
            keys=self.keys, values=values, lengths=lengths.to(self.lengths_dtype), stride=B,
            length_per_key=per_key.tolist(),
        )
        dense = torch.randn(B, self.num_dense, generator=g)
        labels = torch.randint(low=0, high=2, size=(B,), generator=g).to(torch.float32)
        batch = Batch(dense_features=dense, sparse_features=sparse, labels=labels)
        if self.pin_memory and torch.cuda.is_available():
            batch = batch.pin_memory()
        return batch


class RandomRecDataset(IterableDataset):
    """Random dataset of ``Batch``es.

    Args follow the reference: ``keys``, ``batch_size``, ``hash_size``/``hash_sizes``,
    ``ids_per_feature``/``ids_per_features``, ``num_dense``, ``manual_seed``, ``num_batches``,
    ``num_generated_batches``, ``min_ids_per_feature(s)``.
    """

    def __init__(
        self,
        keys: List[str],
        batch_size: int,
        hash_size: Optional[int] = None,
        hash_sizes: Optional[List[int]] = None,
        ids_per_feature: Optional[int] = None,
        ids_per_features: Optional[List[int]] = None,
        num_dense: int = 50,
        manual_seed: Optional[int] = None,
        num_batches: Optional[int] = None,
        num_generated_batches: int = 10,
        min_ids_per_feature: Optional[int] = None,
        min_ids_per_features: Optional[List[int]] = None,
        pin_memory: bool = False,
        index_dtype: torch.dtype = torch.int64,
    ) -> None:
        super().__init__()
        if hash_sizes is None:
            hash_size = hash_size if hash_size is not None else 100
            hash_sizes = [hash_size] * len(keys)
        assert len(hash_sizes) == len(keys), "length of hash_sizes must be equal to the number of keys"
        if ids_per_features is None:
            ids_per_feature = ids_per_feature if ids_per_feature is not None else 2
            ids_per_features = [ids_per_feature] * len(keys)
        assert len(ids_per_features) == len(keys), "length of ids_per_features must be equal to the number of keys"
        if min_ids_per_features is None:
            if min_ids_per_feature is not None:
                min_ids_per_features = [min_ids_per_feature] * len(keys)
            else:
                min_ids_per_features = list(ids_per_features)
        self.batch_generator = _RandomRecBatch(
            keys=keys, batch_size=batch_size, hash_sizes=hash_sizes, ids_per_features=ids_per_features, num_dense=num_dense,
            manual_seed=manual_seed, num_batches=None, num_generated_batches=num_generated_batches,
            min_ids_per_features=min_ids_per_features, pin_memory=pin_memory, index_dtype=index_dtype,
        )
        self.num_batches: int = num_batches if num_batches is not None else 2**62

    def __iter__(self) -> Iterator[Batch]:
        return itertools.islice(iter(self.batch_generator), self.num_batches)

    def __len__(self) -> int:
        return self.num_batches
