"""``criteo_preproc/{train,validation,test}/*.parquet`` -> ``<dst_dir>/{train,validation,test}_data.bin``: row-major records of 40
4-byte fields - label (int32), 13 dense features (float32), 26 category ids (int32). Per-file intermediates are written in parallel and
concatenated in file-name order. Reference: ``datasets/scripts/nvt/convert_parquet_to_binary.py``."""
import argparse
import glob
import os
import shutil
import time

import numpy as np
import pyarrow.parquet as pq

from .utils.criteo_constant import DEFAULT_CAT_NAMES, DEFAULT_COLUMN_NAMES, DEFAULT_INT_NAMES, DEFAULT_LABEL_NAME
from .utils.dask import setup_dask

RECORD_DTYPE = np.dtype([(DEFAULT_LABEL_NAME, np.int32)] + [(c, np.float32) for c in DEFAULT_INT_NAMES] + [(c, np.int32) for c in DEFAULT_CAT_NAMES])
assert RECORD_DTYPE.itemsize == 4 * len(DEFAULT_COLUMN_NAMES)


def process_file(args) -> int:
    src, dst_dir = args
    pf = pq.ParquetFile(src)
    n = 0
    with open(os.path.join(dst_dir, os.path.basename(src) + ".bin"), "wb") as out:
        for rg in range(pf.num_row_groups):
            tbl = pf.read_row_group(rg, columns=DEFAULT_COLUMN_NAMES)
            rec = np.empty(tbl.num_rows, dtype=RECORD_DTYPE)
            for c in DEFAULT_COLUMN_NAMES:
                rec[c] = np.asarray(tbl.column(c).combine_chunks()).astype(RECORD_DTYPE[c])
            out.write(rec.tobytes())
            n += tbl.num_rows
    return n


def convert(src_dir: str, intermediate_dir: str, dst_dir: str, client=None) -> None:
    os.makedirs(dst_dir, exist_ok=True)
    for split in ("train", "test", "validation"):
        files = sorted(glob.glob(os.path.join(src_dir, split, "*.parquet")))
        inter = os.path.join(intermediate_dir, split)
        os.makedirs(inter, exist_ok=True)
        jobs = [(f, inter) for f in files]
        if client is None:
            for j in jobs:
                process_file(j)
        else:
            list(client.map(process_file, jobs))
        with open(os.path.join(dst_dir, f"{split}_data.bin"), "wb") as out:
            for f in files:
                with open(os.path.join(inter, os.path.basename(f) + ".bin"), "rb") as part:
                    shutil.copyfileobj(part, out, 64 << 20)


def main(argv=None) -> None:
    parser = argparse.ArgumentParser()
    parser.add_argument("--src_dir", type=str)
    parser.add_argument("--intermediate_dir", type=str)
    parser.add_argument("--dst_dir", type=str)
    parser.add_argument("--parallel_jobs", default=20, type=int)
    args = parser.parse_args(argv)
    start = time.time()
    client = setup_dask(None, args.parallel_jobs)
    try:
        convert(args.src_dir, args.intermediate_dir, args.dst_dir, client)
    finally:
        client.shutdown()
    print(f"Processing took {time.time() - start:.2f} sec")


if __name__ == "__main__":
    main()
