"""CLI: Criteo TSV days -> npy (dense / sparse / labels) (reference datasets/scripts/npy_preproc_criteo.py)."""
import argparse
import os
import sys
from typing import List

from torchrec_b200.datasets.criteo import BinaryCriteoUtils


def main(argv: List[str]) -> None:
    p = argparse.ArgumentParser(description="Criteo tsv -> npy preprocessing")
    p.add_argument("--input_dir", required=True)
    p.add_argument("--output_dir", required=True)
    p.add_argument("--dataset_name", default="criteo_1tb", choices=["criteo_1tb", "criteo_kaggle"])
    a = p.parse_args(argv)
    os.makedirs(a.output_dir, exist_ok=True)
    for f in sorted(os.listdir(a.input_dir)):
        src = os.path.join(a.input_dir, f)
        if not os.path.isfile(src) or f.endswith(".npy"):
            continue
        base = os.path.join(a.output_dir, f)
        print(f"processing {src}")
        BinaryCriteoUtils.tsv_to_npys(src, base + "_dense.npy", base + "_sparse.npy", base + "_labels.npy", a.dataset_name)


if __name__ == "__main__":
    main(sys.argv[1:])
