"""Build libphyss_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m physs_gp_b200.build [--force] [--jobs N]

One object per translation unit under csrc/, compiled in parallel, linked into
physs_gp_b200/libphyss_b200.so (git-ignored, but it travels to the GPU box with the repo
snapshot).  Objects are cached under csrc/_obj keyed by a hash of the source + headers + flags.
"""
import argparse
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libphyss_b200.so")
LIB_BIG = os.path.join(HERE, "libphyss_b200_big.so")     # large-block path: links cuBLAS + cuSOLVER
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", INCLUDE,
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _headers_digest():
    h = hashlib.sha256()
    for root in (CSRC, INCLUDE):
        for fn in sorted(os.listdir(root)):
            if fn.endswith((".h", ".cuh")):
                with open(os.path.join(root, fn), "rb") as f:
                    h.update(fn.encode())
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(args):
    src, obj, stamp, digest, verbose = args
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    return src, r.stderr


def build_library(force=False, jobs=None, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hd = _headers_digest()
    sources = sorted(fn for fn in os.listdir(CSRC) if fn.endswith(".cu"))
    todo, objs = [], []
    for fn in sources:
        src = os.path.join(CSRC, fn)
        obj = os.path.join(OBJ, fn[:-3] + ".o")
        stamp = obj + ".stamp"
        with open(src, "rb") as f:
            digest = hashlib.sha256(f.read() + hd.encode()).hexdigest()
        objs.append(obj)
        fresh = (not force and os.path.exists(obj) and os.path.exists(stamp)
                 and open(stamp).read() == digest)
        if not fresh:
            todo.append((src, obj, stamp, digest, verbose))
    logs = []
    if todo:
        jobs = jobs or min(len(todo), os.cpu_count() or 4)
        with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
            for src, log in ex.map(_compile_one, todo):
                logs.append((src, log))
    if todo or not os.path.exists(LIB):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    build_big(force=force)
    return LIB, logs


def build_big(force=False):
    """libphyss_b200_big.so from csrc/big/physs_big.cu (one translation unit, -lcublas -lcusolver)."""
    src = os.path.join(CSRC, "big", "physs_big.cu")
    hdr = os.path.join(INCLUDE, "physs_b200_big.h")
    stamp = LIB_BIG + ".stamp"
    with open(src, "rb") as f, open(hdr, "rb") as g:
        digest = hashlib.sha256(f.read() + g.read() + " ".join(NVCC_FLAGS).encode()).hexdigest()
    if (not force and os.path.exists(LIB_BIG) and os.path.exists(stamp) and open(stamp).read() == digest):
        return LIB_BIG
    cmd = [_nvcc()] + NVCC_FLAGS + ["-shared", src, "-o", LIB_BIG, "-lcublas", "-lcusolver"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_BIG


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    ap.add_argument("--ptxas-v", action="store_true", help="print ptxas -v resource usage")
    a = ap.parse_args()
    lib, logs = build_library(force=a.force, jobs=a.jobs, verbose=a.ptxas_v)
    if a.ptxas_v:
        for src, log in logs:
            sys.stdout.write("== %s\n%s\n" % (src, log))
    print(lib)


if __name__ == "__main__":
    main()
