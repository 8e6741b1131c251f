"""Worker-pool set-up. The reference starts a dask-cuda cluster for NVTabular; here the fan-out over files is a process pool
(``setup_dask`` keeps its name and returns the pool; ``client.map`` / ``client.shutdown`` are what the scripts use)."""
import multiprocessing
import os
from concurrent.futures import ProcessPoolExecutor
from typing import Optional


def setup_dask(dask_workdir: Optional[str] = None, n_workers: Optional[int] = None) -> ProcessPoolExecutor:
    if dask_workdir:
        os.makedirs(dask_workdir, exist_ok=True)
    # spawn: the parent may hold torch / arrow threads, forking those can deadlock a worker
    return ProcessPoolExecutor(max_workers=n_workers or max(1, min(16, (os.cpu_count() or 2) - 1)), mp_context=multiprocessing.get_context("spawn"))
