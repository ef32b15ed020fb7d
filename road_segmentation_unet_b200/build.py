"""Build librsu_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The shared object is git-ignored but travels to the GPU box with the repo snapshot, so the
normal flow is: build here (nvcc cross-compiles without a GPU), run there.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librsu_b200.so")
SOURCES = ["host_common.cu", "conv_gemm.cu", "conv_gemm2.cu", "conv_halo.cu", "conv_halo2.cu", "wgrad_gemm.cu", "wgrad_gemm2.cu", "wgrad_halo.cu",
           "conv_first.cu", "elementwise.cu", "geometry.cu", "peer_sgd.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    with open(os.path.join(HERE, "..", "include", "rsu_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _up_to_date(stamp, digest):
    if os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as f:
            return f.read().strip() == digest
    return False


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ and link librsu_b200.so; skipped when up to date.  Several
    processes may call this at once (one per GPU under torchrun): a file lock lets one of them
    build while the others wait and then find the library up to date."""
    stamp = LIB + ".stamp"
    digest = _digest()
    if not force and _up_to_date(stamp, digest):
        return LIB
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(stamp, digest):
                return LIB
            return _build_locked(stamp, digest, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(stamp, digest, verbose):
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
            "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
