"""Build the in-tree C-ABI library with nvcc for sm_100a (cross-compiles without a GPU).

The library lives next to its sources (``ppo_b200/csrc/libppo_b200.so``) so that it travels with
the repository snapshot and shows up as an in-tree mapping of the process that loads it.
"""
import hashlib
import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(CSRC, "libppo_b200.so")
_STAMP = os.path.join(CSRC, ".build_stamp")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas=-v",
    "-cudart", "static",
    "-shared",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint():
    h = hashlib.sha256()
    root = os.path.dirname(os.path.dirname(CSRC))
    files = _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    files.append(os.path.join(root, "include", "ppo_b200.h"))
    for path in files:
        h.update(path.encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into libppo_b200.so.  Returns the library path."""
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(_STAMP):
        with open(_STAMP) as fh:
            if fh.read().strip() == fp:
                return LIB_PATH
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build ppo_b200/csrc/libppo_b200.so")
    # one translation unit per nvcc process, in parallel; objects live in csrc/build/ (git-ignored)
    obj_dir = os.path.join(CSRC, "build")
    os.makedirs(obj_dir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + compile_flags + ["-c", "-o", obj, src]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        return obj, " ".join(cmd) + "\n" + proc.stdout + proc.stderr, proc.returncode

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, _sources()))
    log = "".join(text for _, text, _ in results)
    failed = [text for _, text, rc in results if rc != 0]
    if not failed:
        cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + \
              [obj for obj, _, _ in results]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + proc.stdout + proc.stderr
        if proc.returncode != 0:
            failed = [proc.stdout + proc.stderr]
    with open(os.path.join(CSRC, "build.log"), "w") as fh:
        fh.write(log)
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(failed)[-8000:])
    if verbose:
        print(log)
    with open(_STAMP, "w") as fh:
        fh.write(fp)
    return LIB_PATH


def ensure_built():
    """Library path, building it when missing or stale and a compiler is available."""
    if nvcc_path() is not None:
        return build()
    if os.path.exists(LIB_PATH):
        return LIB_PATH
    raise RuntimeError("ppo_b200/csrc/libppo_b200.so is missing and nvcc is not available to build it")
