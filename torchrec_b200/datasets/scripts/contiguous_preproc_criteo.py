"""CLI: re-index the sparse npy days to contiguous ids (reference datasets/scripts/contiguous_preproc_criteo.py).
``python -m torchrec_b200.datasets.scripts.contiguous_preproc_criteo --input_dir D --output_dir O [--frequency_threshold N] [--days 24]``"""
import argparse
import os
import sys
from typing import List

from torchrec_b200.datasets.criteo import BinaryCriteoUtils

DAYS = 24


def parse_args(argv: List[str]) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Criteo sparse -> contiguous preprocessing script.")
    p.add_argument("--input_dir", type=str, required=True, help="directory with day_{0..days-1}_sparse.npy")
    p.add_argument("--output_dir", type=str, required=True)
    p.add_argument("--frequency_threshold", type=int, default=0, help="ids seen fewer times than this map to index 1 (0: keep all)")
    p.add_argument("--days", type=int, default=DAYS)
    return p.parse_args(argv)


def main(argv: List[str]) -> None:
    a = parse_args(argv)
    files = [os.path.join(a.input_dir, f"day_{i}_sparse.npy") for i in range(a.days)]
    missing = [f for f in files if not os.path.exists(f)]
    if missing:
        raise ValueError(f"missing sparse files in {a.input_dir}: {missing[:3]}{'...' if len(missing) > 3 else ''}")
    print(f"Processing {len(files)} files from {a.input_dir}; outputs go to {a.output_dir}.")
    BinaryCriteoUtils.sparse_to_contiguous(files, a.output_dir, frequency_threshold=int(a.frequency_threshold))
    print("Done processing.")


if __name__ == "__main__":
    main(sys.argv[1:])
