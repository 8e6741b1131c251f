"""``criteo_parquet`` -> ``criteo_preproc/{train,validation,test}``: categories are hashed into their table (missing -> 0), dense
features become log(x + 3) with missing -> 0 (fill -2, add 2, log1p - the reference's NVTabular workflow: FillMissing >> Categorify(num_buckets)
and FillMissing >> +2 >> LogOp). Train = all days but the last, validation / test = the halves of the last day; ``--shuffle_train``
shuffles the rows inside every output file. Reference: ``datasets/scripts/nvt/process_criteo_parquet.py``.

Categorify with ``num_buckets`` maps a category to ``hash(category) % num_buckets``; here the hash is a 64-bit multiplicative mix of the
integer id (deterministic across runs and machines - no fit pass over the training data is needed)."""
import argparse
import os
import shutil
import time
from typing import List

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

from .utils.criteo_constant import DAYS, DEFAULT_CAT_NAMES, DEFAULT_INT_NAMES, DEFAULT_LABEL_NAME, NUM_EMBEDDINGS_PER_FEATURE_DICT
from .utils.dask import setup_dask


def hash_bucket(ids: np.ndarray, num_buckets: int) -> np.ndarray:
    x = ids.astype(np.uint64)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xFF51AFD7ED558CCD)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xC4CEB9FE1A85EC53)
    x ^= x >> np.uint64(33)
    return (x % np.uint64(num_buckets)).astype(np.int64)


def transform_table(tbl: pa.Table) -> pa.Table:
    cols, names = [], []
    for c in DEFAULT_CAT_NAMES:
        arr = tbl.column(c).combine_chunks()
        valid = np.asarray(arr.is_valid())
        ids = np.asarray(arr.fill_null(0))
        out = np.where(valid, hash_bucket(ids, NUM_EMBEDDINGS_PER_FEATURE_DICT[c]), 0)
        cols.append(pa.array(out.astype(np.int64)))
        names.append(c)
    for c in DEFAULT_INT_NAMES:
        x = np.asarray(tbl.column(c).combine_chunks().fill_null(-2)).astype(np.float64)
        cols.append(pa.array(np.log1p(np.maximum(x + 2.0, 0.0)).astype(np.float32)))
        names.append(c)
    cols.append(pa.array(np.asarray(tbl.column(DEFAULT_LABEL_NAME).combine_chunks().fill_null(0)).astype(np.float32)))
    names.append(DEFAULT_LABEL_NAME)
    return pa.Table.from_arrays(cols, names=names)


def process_file(args) -> int:
    src, dst, shuffle, seed = args
    pf = pq.ParquetFile(src)
    n = 0
    writer = None
    try:
        for rg in range(pf.num_row_groups):
            out = transform_table(pf.read_row_group(rg))
            if shuffle:
                out = out.take(pa.array(np.random.default_rng(seed + rg).permutation(out.num_rows)))
            if writer is None:
                writer = pq.ParquetWriter(dst, out.schema)
            writer.write_table(out)
            n += out.num_rows
    finally:
        if writer is not None:
            writer.close()
    return n


def process(base_path: str, shuffle_train: bool = False, days: int = DAYS, client=None) -> str:
    input_path = os.path.join(base_path, "criteo_parquet")
    assert os.path.exists(input_path), f"Criteo parquet path {input_path} does not exist"
    output_path = os.path.join(base_path, "criteo_preproc")
    if os.path.exists(output_path):
        shutil.rmtree(output_path)
    jobs: List = []
    for split, files in (("train", [f"day_{d}.parquet" for d in range(days - 1)]), ("validation", [f"day_{days - 1}.part0.parquet"]), ("test", [f"day_{days - 1}.part1.parquet"])):
        os.makedirs(os.path.join(output_path, split))
        for i, f in enumerate(files):
            jobs.append((os.path.join(input_path, f), os.path.join(output_path, split, f"part_{i}.parquet"), shuffle_train and split == "train", 1000 * i))
    if client is None:
        for j in jobs:
            process_file(j)
    else:
        list(client.map(process_file, jobs))
    return output_path


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Preprocess criteo dataset")
    parser.add_argument("--base_path", "-b", dest="base_path", help="Base path")
    parser.add_argument("--shuffle_train", "-s", dest="shuffle_train", default=False, action="store_true", help="shuffle the rows of the training files")
    parser.add_argument("--days", type=int, default=DAYS)
    return parser.parse_args(argv)


def main(argv=None) -> None:
    args = parse_args(argv)
    start = time.time()
    client = setup_dask(os.path.join(args.base_path, "dask_workdir"))
    try:
        process(args.base_path, args.shuffle_train, args.days, client)
    finally:
        client.shutdown()
    print(f"Processing took {time.time() - start:.2f} sec")


if __name__ == "__main__":
    main()
