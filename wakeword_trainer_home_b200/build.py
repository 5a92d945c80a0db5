"""In-tree build of libwwfeat.so for sm_100a (nvcc cross-compiles without a GPU).

The feature kernel is a template over n_fft; each n_fft family is its own translation unit
(csrc/wwf_feat_inst.cu with -DWWF_INST_NFFT=...), compiled in parallel with the host / reverb /
auxiliary unit (csrc/wwfeat.cu), then linked into one shared object.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libwwfeat.so")
N_FFTS = (256, 400, 512, 1024, 2048)
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "wwfeat.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    return r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """``out`` / ``defines`` build an experimental variant next to the product library (A/B runs load it
    through the WWF_LIB environment variable)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if out == OUT and not force and not _stale():
        return OUT
    OUT_ = out
    os.makedirs(os.path.dirname(OUT_), exist_ok=True)
    OBJ = tempfile.mkdtemp(prefix="wwfeat_obj_")          # objects stay out of the tree (only the .so ships)
    extra = (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines]
    jobs = [([nvcc] + FLAGS + extra + ["-c", os.path.join(CSRC, "wwfeat.cu"), "-o", os.path.join(OBJ, "wwfeat.o")])]
    for n in N_FFTS:
        jobs.append([nvcc] + FLAGS + extra + [f"-DWWF_INST_NFFT={n}", "-c", os.path.join(CSRC, "wwf_feat_inst.cu"),
                                               "-o", os.path.join(OBJ, f"feat_{n}.o")])
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
        results = list(ex.map(_run, jobs))
    log = "".join(out for _, out in results)
    if any(rc != 0 for rc, _ in results):
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libwwfeat.so")
    objs = [os.path.join(OBJ, "wwfeat.o")] + [os.path.join(OBJ, f"feat_{n}.o") for n in N_FFTS]
    rc, msg = _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT_] + objs)
    if rc != 0:
        sys.stderr.write(msg)
        raise RuntimeError("nvcc failed linking libwwfeat.so")
    shutil.rmtree(OBJ, ignore_errors=True)
    if verbose:
        print(log)
    return OUT_


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=os.path.abspath(outs[0]) if outs else OUT, defines=defs))
