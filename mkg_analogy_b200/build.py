"""In-tree build of libb200mart.so (nvcc, sm_100a only).

Each csrc/*.cu is compiled to an object file in parallel (objects are cached by content hash under
csrc/.obj/), then linked into one shared library.  The .so is git-ignored but travels to the GPU box with
the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, ".obj")
LIB = os.path.join(HERE, "libb200mart.so")
STAMP = os.path.join(HERE, ".libb200mart.stamp")
HEADER = os.path.join(HERE, "..", "include", "b200mart.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(".cuh") or f.endswith(".h"):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(f.encode() + fh.read())
    with open(HEADER, "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h


def _obj_digest(src, hd):
    h = hd.copy()
    with open(src, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()[:24]


def _digest():
    hd = _headers_digest()
    return hashlib.sha256("".join(_obj_digest(s, hd) for s in sources()).encode()).hexdigest()


def is_current():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _digest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library. No-op when sources are unchanged."""
    if not force and is_current():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hd = _headers_digest()

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + "." + _obj_digest(src, hd) + ".o")
        if os.path.exists(obj) and not force:
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}{res.stderr}")
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    objs = [o for o, _ in results]
    if verbose:
        sys.stderr.write("".join(log for _, log in results))
    # drop stale objects
    keep = set(objs)
    for f in os.listdir(OBJ):
        p = os.path.join(OBJ, f)
        if p not in keep:
            os.remove(p)
    res = subprocess.run([nvcc, "-shared", "-o", LIB] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link of libb200mart.so failed:\n" + res.stdout + res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
