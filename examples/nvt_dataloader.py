"""Binary Criteo dataloader in the layout NVTabular preprocessing produces (reference ``examples/nvt_dataloader/nvt_binary_dataloader.py``, ``train_torchrec.py``).

The reference converts Criteo TSV -> parquet -> three flat binary files per day with NVTabular (``torchrec/datasets/scripts/nvt/``) and then reads fixed-size
record blocks: ``label.bin`` (B x int32 -> float), ``dense.bin`` (B x 13 x fp16/fp32), ``sparse.bin`` (B x 26 x int32 ids, already hashed into table range).
No NVTabular here (no GPU dataframe library in the image), but the FORMAT is just three arrays - this module writes and reads it with numpy memmaps:

* ``write_binary_dataset(dir, n, ...)``  - synthetic data in that layout (or convert your own arrays with ``save_arrays``),
* ``NvtBinaryDataset``                   - map-style dataset of whole batches: one contiguous read per file per batch, pinned, rank-strided,
* ``main()``                             - trains a small DLRM from the files through the sparse-dist pipeline.

Every sample has exactly ONE id per feature, so a batch's KJT lengths are all ones and ``values`` is the transposed id block - no per-sample python work.
"""
import argparse
import os
import sys
import tempfile
from typing import Iterator, Optional

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from torchrec_b200.datasets.utils import Batch  # noqa: E402
from torchrec_b200.sparse import KeyedJaggedTensor  # noqa: E402

NUM_DENSE, NUM_SPARSE = 13, 26
KEYS = [f"cat_{i}" for i in range(NUM_SPARSE)]


def save_arrays(directory: str, labels: np.ndarray, dense: np.ndarray, sparse: np.ndarray, dense_dtype: str = "float16") -> None:
    os.makedirs(directory, exist_ok=True)
    labels.astype(np.int32).tofile(os.path.join(directory, "label.bin"))
    dense.astype(dense_dtype).tofile(os.path.join(directory, "dense.bin"))
    sparse.astype(np.int32).tofile(os.path.join(directory, "sparse.bin"))
    with open(os.path.join(directory, "meta.txt"), "w") as f:
        f.write(f"{labels.shape[0]} {dense_dtype}\n")


def write_binary_dataset(directory: str, num_samples: int, hash_sizes, seed: int = 0) -> None:
    rng = np.random.default_rng(seed)
    sparse = np.stack([rng.integers(0, h, num_samples) for h in hash_sizes], axis=1)
    dense = np.log1p(rng.exponential(2.0, (num_samples, NUM_DENSE)))
    # a learnable signal: the label depends on the parity of two features and one dense column
    logit = (sparse[:, 0] % 2) * 1.5 - (sparse[:, 1] % 2) * 1.5 + (dense[:, 0] - dense[:, 0].mean())
    labels = (rng.random(num_samples) < 1.0 / (1.0 + np.exp(-logit))).astype(np.int32)
    save_arrays(directory, labels, dense, sparse)


class NvtBinaryDataset(torch.utils.data.Dataset):
    """``ds[i]`` = batch ``i`` of THIS rank (global batch ``i * world + rank``): ``Batch(dense [B, 13] fp32, sparse KJT, labels [B])``."""

    def __init__(self, directory: str, batch_size: int, rank: int = 0, world_size: int = 1, drop_last: bool = True, pin_memory: bool = False) -> None:
        n, dense_dtype = open(os.path.join(directory, "meta.txt")).read().split()
        self.n, self.B, self.rank, self.world = int(n), batch_size, rank, world_size
        self.labels = np.memmap(os.path.join(directory, "label.bin"), dtype=np.int32, mode="r", shape=(self.n,))
        self.dense = np.memmap(os.path.join(directory, "dense.bin"), dtype=dense_dtype, mode="r", shape=(self.n, NUM_DENSE))
        self.sparse = np.memmap(os.path.join(directory, "sparse.bin"), dtype=np.int32, mode="r", shape=(self.n, NUM_SPARSE))
        total = self.n // batch_size if drop_last else -(-self.n // batch_size)
        self.num_batches = total // world_size
        self.pin = pin_memory and torch.cuda.is_available()
        self._lengths = torch.ones(NUM_SPARSE * batch_size, dtype=torch.int32)

    def __len__(self) -> int:
        return self.num_batches

    def __getitem__(self, i: int) -> Batch:
        if not 0 <= i < self.num_batches:
            raise IndexError(i)
        lo = (i * self.world + self.rank) * self.B
        hi = min(lo + self.B, self.n)
        dense = torch.from_numpy(np.ascontiguousarray(self.dense[lo:hi]).astype(np.float32))
        labels = torch.from_numpy(np.ascontiguousarray(self.labels[lo:hi]).astype(np.float32))
        ids = torch.from_numpy(np.ascontiguousarray(self.sparse[lo:hi].T).astype(np.int64)).reshape(-1)  # feature-major: [26 * B]
        lengths = self._lengths if hi - lo == self.B else torch.ones(NUM_SPARSE * (hi - lo), dtype=torch.int32)
        b = Batch(dense_features=dense, sparse_features=KeyedJaggedTensor(keys=KEYS, values=ids, lengths=lengths, stride=hi - lo), labels=labels)
        return b.pin_memory() if self.pin else b

    def __iter__(self) -> Iterator[Batch]:
        for i in range(self.num_batches):
            yield self[i]


class NvtSplitBinaryDataset(NvtBinaryDataset):
    """The per-column layout ``torchrec_b200.datasets.scripts.nvt.split_binary_dataset`` writes (the reference's NVT dataloader layout):
    ``numerical.bin`` float32 [N, 13], ``label.bin`` float32 [N], ``cat_<i>.bin`` int32 [N] for i in 0..25. One contiguous read per
    column file per batch; the ids arrive feature-major, which is the KJT value order."""

    def __init__(self, directory: str, batch_size: int, rank: int = 0, world_size: int = 1, drop_last: bool = True, pin_memory: bool = False) -> None:
        self.n = os.path.getsize(os.path.join(directory, "label.bin")) // 4
        self.B, self.rank, self.world = batch_size, rank, world_size
        self.labels = np.memmap(os.path.join(directory, "label.bin"), dtype=np.float32, mode="r", shape=(self.n,))
        self.dense = np.memmap(os.path.join(directory, "numerical.bin"), dtype=np.float32, mode="r", shape=(self.n, NUM_DENSE))
        self.cats = [np.memmap(os.path.join(directory, f"cat_{i}.bin"), dtype=np.int32, mode="r", shape=(self.n,)) for i in range(NUM_SPARSE)]
        total = self.n // batch_size if drop_last else -(-self.n // batch_size)
        self.num_batches = total // world_size
        self.pin = pin_memory and torch.cuda.is_available()
        self._lengths = torch.ones(NUM_SPARSE * batch_size, dtype=torch.int32)

    def __getitem__(self, i: int) -> Batch:
        if not 0 <= i < self.num_batches:
            raise IndexError(i)
        lo = (i * self.world + self.rank) * self.B
        hi = min(lo + self.B, self.n)
        dense = torch.from_numpy(np.array(self.dense[lo:hi], dtype=np.float32))
        labels = torch.from_numpy(np.array(self.labels[lo:hi], dtype=np.float32))
        ids = torch.from_numpy(np.concatenate([c[lo:hi] for c in self.cats]).astype(np.int64))  # feature-major: [26 * B]
        lengths = self._lengths if hi - lo == self.B else torch.ones(NUM_SPARSE * (hi - lo), dtype=torch.int32)
        b = Batch(dense_features=dense, sparse_features=KeyedJaggedTensor(keys=KEYS, values=ids, lengths=lengths, stride=hi - lo), labels=labels)
        return b.pin_memory() if self.pin else b


def main(steps: int = 60, batch_size: int = 256, directory: Optional[str] = None) -> float:
    from torchrec_b200.models.dlrm import DLRM, DLRMTrain
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection

    torch.manual_seed(0)
    hash_sizes = [1000] * NUM_SPARSE
    tmp = None
    if directory is None:
        tmp = tempfile.TemporaryDirectory()
        directory = tmp.name
        write_binary_dataset(directory, steps * batch_size, hash_sizes)
    ds = NvtBinaryDataset(directory, batch_size)
    ebc = EmbeddingBagCollection([EmbeddingBagConfig(name=f"t_{k}", embedding_dim=16, num_embeddings=h, feature_names=[k]) for k, h in zip(KEYS, hash_sizes)])
    model = DLRMTrain(DLRM(ebc, NUM_DENSE, [32, 16], [32, 1]))
    opt = torch.optim.Adagrad(model.parameters(), lr=0.05)
    first = last = 0.0
    for step, batch in enumerate(ds):
        loss, _ = model(batch)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step < 5:
            first += float(loss) / 5
        if step >= len(ds) - 5:
            last += float(loss) / 5
    if tmp is not None:
        tmp.cleanup()
    print(f"loss {first:.4f} -> {last:.4f} over {len(ds)} batches read from {directory}")
    return last / max(first, 1e-9)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default=None)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--batch-size", type=int, default=256)
    a = ap.parse_args()
    main(a.steps, a.batch_size, a.dir)
