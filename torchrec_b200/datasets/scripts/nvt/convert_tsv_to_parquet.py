"""Criteo day files (tab separated: label, 13 integers, 26 hexadecimal categories, empty = missing) -> one parquet file per day under
``<output_base_path>/criteo_parquet``; the last day is split into ``day_23.part0`` (first half: validation) and ``day_23.part1``
(second half: test). Integers become nullable int32, the hex strings nullable int64 (reference: int32 after NVTabular's "hex" reader;
int64 keeps ids above 2^31). Reference: ``datasets/scripts/nvt/convert_tsv_to_parquet.py``."""
import argparse
import os
import shutil
import time
from typing import List, Optional

import numpy as np
import pyarrow as pa
import pyarrow.csv as pacsv
import pyarrow.parquet as pq

from .utils.criteo_constant import DAYS, DEFAULT_CAT_NAMES, DEFAULT_COLUMN_NAMES, DEFAULT_INT_NAMES, DEFAULT_LABEL_NAME
from .utils.dask import setup_dask


def _hex_to_int64(col: pa.ChunkedArray) -> pa.Array:
    arr = col.combine_chunks()
    valid = np.asarray(arr.is_valid())
    strs = arr.to_pylist()
    out = np.zeros(len(strs), dtype=np.int64)
    for i, s in enumerate(strs):
        if s:
            out[i] = int(s, 16)
        else:
            valid[i] = False
    return pa.array(out, mask=~valid)


def convert_file(src: str, dst: str, first_line: int = 0, n_lines: Optional[int] = None, block_size: int = 64 << 20) -> int:
    """One TSV file (or the line range [first_line, first_line + n_lines) of it) -> one parquet file; returns the number of rows."""
    read = pacsv.ReadOptions(column_names=DEFAULT_COLUMN_NAMES, block_size=block_size)
    parse = pacsv.ParseOptions(delimiter="\t")
    types = {DEFAULT_LABEL_NAME: pa.int32(), **{c: pa.int32() for c in DEFAULT_INT_NAMES}, **{c: pa.string() for c in DEFAULT_CAT_NAMES}}
    conv = pacsv.ConvertOptions(column_types=types, null_values=[""], strings_can_be_null=True)
    schema = pa.schema([(DEFAULT_LABEL_NAME, pa.int32())] + [(c, pa.int32()) for c in DEFAULT_INT_NAMES] + [(c, pa.int64()) for c in DEFAULT_CAT_NAMES])
    written, seen = 0, 0
    last = None if n_lines is None else first_line + n_lines
    with pq.ParquetWriter(dst, schema) as writer:
        for batch in pacsv.open_csv(src, read_options=read, parse_options=parse, convert_options=conv):
            lo, hi = seen, seen + batch.num_rows
            seen = hi
            if hi <= first_line:
                continue
            if last is not None and lo >= last:
                break
            a, b = max(first_line - lo, 0), (min(last, hi) - lo if last is not None else batch.num_rows)
            tbl = pa.Table.from_batches([batch.slice(a, b - a)])
            cols = [tbl.column(DEFAULT_LABEL_NAME).combine_chunks()] + [tbl.column(c).combine_chunks() for c in DEFAULT_INT_NAMES] + \
                   [_hex_to_int64(tbl.column(c)) for c in DEFAULT_CAT_NAMES]
            writer.write_table(pa.Table.from_arrays(cols, schema=schema))
            written += b - a
    return written


def _count_lines(path: str) -> int:
    n = 0
    with open(path, "rb") as f:
        while True:
            buf = f.read(16 << 20)
            if not buf:
                return n
            n += buf.count(b"\n")


def _job(args) -> int:
    return convert_file(*args)


def convert_tsv_to_parquet(input_path: str, output_base_path: str, days: int = DAYS, client=None) -> List[str]:
    output_path = os.path.join(output_base_path, "criteo_parquet")
    if os.path.exists(output_path):
        shutil.rmtree(output_path)
    os.makedirs(output_path)
    last = os.path.join(input_path, f"day_{days - 1}")
    n = _count_lines(last)
    valid_size = n // 2
    first_size = n - valid_size  # the first part takes the extra line of an odd count
    jobs = [(os.path.join(input_path, f"day_{d}"), os.path.join(output_path, f"day_{d}.parquet"), 0, None) for d in range(days - 1)]
    jobs += [(last, os.path.join(output_path, f"day_{days - 1}.part0.parquet"), 0, first_size),
             (last, os.path.join(output_path, f"day_{days - 1}.part1.parquet"), first_size, valid_size)]
    if client is None:
        for j in jobs:
            _job(j)
    else:
        list(client.map(_job, jobs))
    return [j[1] for j in jobs]


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Convert criteo tsv to parquet")
    parser.add_argument("--input_path", "-i", dest="input_path", help="Input path containing tsv files")
    parser.add_argument("--output_base_path", "-o", dest="output_base_path", help="Output base path")
    parser.add_argument("--days", type=int, default=DAYS, help="number of day files (day_0 .. day_<days-1>)")
    return parser.parse_args(argv)


def main(argv=None) -> None:
    args = parse_args(argv)
    assert os.path.exists(args.input_path), f"Input path {args.input_path} does not exist"
    client = setup_dask(os.path.join(args.output_base_path, "dask_workdir"))
    start = time.time()
    try:
        convert_tsv_to_parquet(args.input_path, args.output_base_path, args.days, client)
    finally:
        client.shutdown()
    print(f"Conversion from tsv to parquet took {time.time() - start:.2f} sec")


if __name__ == "__main__":
    main()
