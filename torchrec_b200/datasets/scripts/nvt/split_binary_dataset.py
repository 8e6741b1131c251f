"""Row-major ``{train,test,validation}_data.bin`` -> one directory per split with column files: ``numerical.bin`` (float32 [N, 13]),
``label.bin`` (float32 [N]), ``cat_<i>.bin`` (int32 [N]) - the layout the NVT-style binary dataloader memory-maps (``examples/
nvt_dataloader.py``). Streams ``batch_size`` records at a time. Reference: ``datasets/scripts/nvt/split_binary_dataset.py``."""
import argparse
import math
import os
import time
from typing import Sequence

import numpy as np

from .utils.criteo_constant import CAT_FEATURE_COUNT, DEFAULT_INT_NAMES, NUM_EMBEDDINGS_PER_FEATURE


def split_binary_file(binary_file_path: str, output_dir: str, categorical_feature_sizes: Sequence[int], batch_size: int, source_data_type: str = "int32") -> int:
    n_int, n_cat = len(DEFAULT_INT_NAMES), len(categorical_feature_sizes)
    record_width = 1 + n_int + n_cat
    bytes_per_entry = record_width * np.dtype(source_data_type).itemsize
    n_records = os.path.getsize(binary_file_path) // bytes_per_entry
    os.makedirs(output_dir, exist_ok=True)
    streams = []
    try:
        src = open(binary_file_path, "rb")  # noqa: SIM115
        streams.append(src)
        numerical = open(os.path.join(output_dir, "numerical.bin"), "wb")  # noqa: SIM115
        label = open(os.path.join(output_dir, "label.bin"), "wb")  # noqa: SIM115
        cats = [open(os.path.join(output_dir, f"cat_{i}.bin"), "wb") for i in range(n_cat)]  # noqa: SIM115
        streams += [numerical, label, *cats]
        for _ in range(int(math.ceil(n_records / batch_size))):
            raw = np.frombuffer(src.read(bytes_per_entry * batch_size), dtype=np.int32).reshape(-1, record_width)
            numerical.write(np.ascontiguousarray(raw[:, 1 : 1 + n_int]).view(np.float32).tobytes())
            label.write(raw[:, 0].astype(np.float32).tobytes())
            for i in range(n_cat):
                cats[i].write(np.ascontiguousarray(raw[:, 1 + n_int + i]).astype(np.int32).tobytes())
    finally:
        for s in streams:
            s.close()
    return n_records


def split_dataset(dataset_dir: str, output_dir: str, batch_size: int) -> None:
    os.makedirs(output_dir, exist_ok=True)
    for split in ("test", "train", "validation"):
        split_binary_file(os.path.join(dataset_dir, f"{split}_data.bin"), os.path.join(output_dir, split), NUM_EMBEDDINGS_PER_FEATURE[:CAT_FEATURE_COUNT], batch_size)


def main(argv=None) -> None:
    parser = argparse.ArgumentParser()
    parser.add_argument("--input_path", type=str, required=True)
    parser.add_argument("--output_path", type=str, required=True)
    parser.add_argument("--batch_size", type=int, required=True)
    args = parser.parse_args(argv)
    start = time.time()
    split_dataset(args.input_path, args.output_path, args.batch_size)
    print(f"Processing took {time.time() - start:.2f} sec")


if __name__ == "__main__":
    main()
