"""In-tree build of libu96stereo.so (nvcc, sm_100a only) -- no JIT cache, no fallback."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib", "libu96stereo.so")
SOURCES = ["u96_stereo.cu", "rect.cu", "xsobel.cu", "bm.cu", "bm_fast_cs1.cu", "bm_fast_cs2.cu", "bm_fast_cs4.cu", "bm_fused.cu", "reproject.cu", "uvc.cu", "gftt.cu", "postfilter.cu", "microbench.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]
OBJDIR = os.path.join(PKG, "lib", "obj")


def source_sha16(prefix="bm"):
    """SHA-256 (first 16 hex digits) over the sources of the BM kernels (csrc/bm*.cu, bm*.cuh, common.cuh): identifies the code a
    BM profile was taken with.  (nvcc does not produce bit-identical libraries from identical sources, so the hash of the .so would
    not survive a rebuild; the handle / rect / post-filter sources do not change what the BM kernel is.)"""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        path = os.path.join(CSRC, f)
        if os.path.isfile(path) and f.endswith((".cu", ".cuh")) and (f.startswith(prefix) or f == "common.cuh"):
            h.update(f.encode()); h.update(open(path, "rb").read())
    return h.hexdigest()[:16]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, "..", "include", "u96_stereo.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [os.path.join(PKG, "..", "include", "u96_stereo.h")]
    hdr_t = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(OBJDIR, src[:-3] + ".o")
        path = os.path.join(CSRC, src)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(path), hdr_t):
            subprocess.check_call([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj])
        return obj

    from concurrent.futures import ThreadPoolExecutor      # translation units compile side by side (the BM templates dominate)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
