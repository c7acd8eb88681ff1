"""Build libvireo_b200.so in-tree with nvcc for sm_100a.

    python -m vireo_b200.build [--force] [--verbose]

The shared library lands next to this file (vireo_b200/libvireo_b200.so) so that it travels to the
GPU box with the repository snapshot; it is git-ignored.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in ("vb_stage.cu", "vb_em.cu", "vb_seg.cu")]
DEPS = SRC + [os.path.join(HERE, "csrc", "vb_common.cuh"), os.path.join(HERE, "csrc", "vb_stream.cuh"),
              os.path.join(HERE, "csrc", "vb_tail.cuh"), os.path.join(os.path.dirname(HERE), "include", "vireo_b200.h")]
LIB = os.path.join(HERE, "libvireo_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libvireo_b200.so cannot be built")


# sanitizer variants: same sources, one macro (see vb_seg.cu); built on demand, loaded through VIREO_B200_LIB
VARIANTS = {"plainfill": ["-DVB_SEG_PLAIN_FILL", "-DVB_SEG_CANARY"], "canary": ["-DVB_SEG_CANARY"],
            # timing diagnostic of the segment kernels (results are garbage by design): the table loads never execute
            "nolds": ["-DVB_SEG_DIAG_NOLDS"],
            # window waits that spin on try_wait instead of sleeping between polls
            "spin": ["-DVB_SEG_SLEEP_P=0", "-DVB_SEG_SLEEP_C=0"]}


def up_to_date(lib=LIB):
    if not os.path.exists(lib):
        return False
    t = os.path.getmtime(lib)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False, variant=None):
    """One object per source (compiled in parallel, re-used while the source and the headers are older), then the
    link -- a full build is as long as its slowest file."""
    lib = LIB if variant is None else LIB.replace(".so", "_%s.so" % variant)
    if not force and up_to_date(lib):
        return lib
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "_build", variant or "default")
    os.makedirs(objdir, exist_ok=True)
    headers = [d for d in DEPS if d not in SRC]
    flags = ["-O3", "-std=c++17", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC", *(VARIANTS[variant] if variant else []),
             "-Xptxas", "-v" if verbose else "-O3"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if (not force and os.path.exists(obj)
                and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in [src] + headers)):
            return obj, 0, ""
        res = subprocess.run([nvcc_path(), *flags, "-c", "-o", obj, src], capture_output=True, text=True)
        return obj, res.returncode, res.stdout + res.stderr

    with ThreadPoolExecutor(len(SRC)) as ex:
        done = list(ex.map(compile_one, SRC))
    for obj, rc, log in done:
        if verbose or rc != 0:
            sys.stderr.write(log)
        if rc != 0:
            raise RuntimeError("nvcc failed (exit %d) on %s:\n%s" % (rc, obj, log[-4000:]))
    res = subprocess.run([nvcc_path(), *ARCH, "-shared", "-o", lib, *[o for o, _, _ in done], "-ldl"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed (exit %d):\n%s" % (res.returncode, res.stderr[-4000:]))
    return lib


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=var[0] if var else None))
