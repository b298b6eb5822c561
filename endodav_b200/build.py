"""Build recipe for the sm_100a C-ABI library (in-tree, so the .so travels to the GPU box).

The kernels are split over several translation units (csrc/*.cu) that nvcc compiles in parallel;
objects go to csrc/build/ and are relinked into csrc/libendodav_b200.so.  An object is rebuilt when
its source or ANY header of csrc/ or include/ is newer (the headers are shared templates)."""
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libendodav_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_header_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(nvcc, src, obj):
    cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    started = time.time()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode == 0 and os.path.exists(obj):
        os.utime(obj, (started, started))   # a source edited WHILE nvcc was running must still look newer than the object
    return src, cmd, res


def build(force=False, verbose=False):
    """Compile csrc/*.cu into csrc/libendodav_b200.so with nvcc (cross-compiles without a GPU)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr = _newest_header_mtime()
    todo, objs = [], []
    for s in sources():
        o = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(hdr, os.path.getmtime(os.path.join(CSRC, s))):
            todo.append((s, o))
    if not todo and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(o) for o in objs):
        return LIB
    log = os.path.join(CSRC, "build.log")
    failed = []
    with open(log, "w") as f:
        with ThreadPoolExecutor(max_workers=max(1, min(len(todo) or 1, os.cpu_count() or 1))) as ex:
            for src, cmd, res in ex.map(lambda so: _compile(nvcc, *so), todo):
                f.write("### %s\n%s\n%s%s\n" % (src, " ".join(cmd), res.stdout, res.stderr))
                if verbose or res.returncode != 0:
                    sys.stderr.write(res.stdout + res.stderr)
                if res.returncode != 0:
                    failed.append(src)
        if not failed:
            cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB]
            res = subprocess.run(cmd, capture_output=True, text=True)
            f.write("### link\n%s\n%s%s\n" % (" ".join(cmd), res.stdout, res.stderr))
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                failed.append("link")
    if failed:
        raise RuntimeError("nvcc failed for %s (see %s)" % (", ".join(failed), log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
