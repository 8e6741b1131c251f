"""CLI: globally shuffle the training days of the preprocessed Criteo npy set (reference datasets/scripts/shuffle_preproc_criteo.py).
The last day is left in order (it is the eval / test set)."""
import argparse
import os
import sys
from typing import Dict, List

import numpy as np

from torchrec_b200.datasets.criteo import BinaryCriteoUtils

DAYS = 24


def parse_args(argv: List[str]) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Criteo npy shuffle script.")
    p.add_argument("--input_dir_labels_and_dense", type=str, required=True, help="directory with day_*_dense.npy and day_*_labels.npy")
    p.add_argument("--input_dir_sparse", type=str, required=True, help="directory with day_*_sparse.npy (or their *_contig_freq.npy versions)")
    p.add_argument("--output_dir_shuffled", type=str, required=True)
    p.add_argument("--random_seed", type=int, default=0)
    p.add_argument("--days", type=int, default=DAYS)
    return p.parse_args(argv)


def count_rows(rows_per_file: Dict[int, int], path: str, day: int) -> None:
    rows_per_file[day] = int(np.load(os.path.join(path, f"day_{day}_labels.npy"), mmap_mode="r").shape[0])


def main(argv: List[str]) -> None:
    a = parse_args(argv)
    rows: Dict[int, int] = {}
    for d in range(a.days):
        count_rows(rows, a.input_dir_labels_and_dense, d)
    print(f"rows per day: {rows}")
    BinaryCriteoUtils.shuffle(a.input_dir_labels_and_dense, a.input_dir_sparse, a.output_dir_shuffled, rows, days=a.days, random_seed=a.random_seed)
    print("Done shuffling.")


if __name__ == "__main__":
    main(sys.argv[1:])
