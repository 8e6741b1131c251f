"""In-tree build of the native (host C++) runtime libraries: dynamic embedding, inference server core.

Each sub-directory ``csrc/<name>/`` becomes ``csrc/_lib/libtrb_<name>.so`` (g++ -O2 -std=c++17, C ABI, ctypes-loaded,
no torch headers -> seconds per library). ``python -m torchrec_b200.csrc.build`` / ``__graft_entry__.build()``."""
from __future__ import annotations

import os
import subprocess
from typing import Dict, List

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_lib")
LIBS: Dict[str, List[str]] = {"dynemb": ["-lpthread", "-ldl"], "serving": ["-lpthread"]}


def lib_path(name: str) -> str:
    return os.path.join(OUT, f"libtrb_{name}.so")


def build_one(name: str, force: bool = False, verbose: bool = False) -> str:
    src_dir = os.path.join(HERE, name)
    srcs = sorted(os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith(".cpp"))
    deps = srcs + [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith(".h")]
    lib = lib_path(name)
    os.makedirs(OUT, exist_ok=True)
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= max(os.path.getmtime(d) for d in deps):
        return lib
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-o", lib, *srcs, *LIBS.get(name, [])]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for {name}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[trb200 build] {os.path.basename(lib)} ok")
    return lib


def build_native_test(name: str = "dynemb", verbose: bool = False) -> str:
    """Compile ``csrc/tests/<name>_test.cpp`` together with the library's sources into ``csrc/_lib/<name>_test`` (a plain executable:
    the image has neither gtest nor google-benchmark). ``<exe>`` runs the checks, ``<exe> --bench`` the micro-benchmark."""
    src_dir = os.path.join(HERE, name)
    test_src = os.path.join(HERE, "tests", f"{name}_test.cpp")
    srcs = sorted(os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith(".cpp") and f != "c_api.cpp")
    exe = os.path.join(OUT, f"{name}_test")
    os.makedirs(OUT, exist_ok=True)
    deps = srcs + [test_src] + [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith(".h")]
    if os.path.exists(exe) and os.path.getmtime(exe) >= max(os.path.getmtime(d) for d in deps):
        return exe
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, test_src, *srcs, *LIBS.get(name, [])]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for {name}_test:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[trb200 build] {os.path.basename(exe)} ok")
    return exe


def build(force: bool = False, verbose: bool = False) -> List[str]:
    return [build_one(n, force, verbose) for n in LIBS if os.path.isdir(os.path.join(HERE, n))]


def load(name: str):
    """ctypes handle of a native library, building it on first use."""
    import ctypes

    path = lib_path(name)
    if not os.path.exists(path) or os.environ.get("TRB_REBUILD_NATIVE"):
        build_one(name)
    else:
        try:
            build_one(name)  # rebuild when sources are newer (no-op otherwise)
        except (RuntimeError, FileNotFoundError):
            pass
    return ctypes.CDLL(path)


if __name__ == "__main__":
    build(verbose=True)
