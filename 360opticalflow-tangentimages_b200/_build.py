"""Build libtif_b200.so (sm_100a) in-tree with nvcc.  No torch headers are involved: the library is
plain CUDA runtime + the C ABI of include/tif_b200.h, so it cross-compiles on a box without a GPU."""
import fcntl
import hashlib
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(REPO_ROOT, "include")
BUILD_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libtif_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC]
# translation unit -> extra flags.  The plan restates numpy float64 expressions one rounding per
# operation, so it must not contract a*b+c into an FMA.
UNITS = {
    "tif_plan.cu": ["-fmad=false"],
    "tif_exec.cu": [],
    "tif_warp.cu": [],
    "tif_rotate.cu": [],
    "tif_metrics.cu": [],
    "tif_vis.cu": ["-fmad=false"],       # numpy evaluates u*u + v*v and the wheel interpolation without contraction
}


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libtif_b200.so cannot be built")
    return exe


def _digest(paths, extra):
    """Content hash of the sources and flags (mtimes do not survive the copy to the GPU box)."""
    h = hashlib.sha1()
    for path in paths:
        h.update(open(path, "rb").read())
    h.update(repr(extra).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    """Compile every CUDA unit for sm_100a and link libtif_b200.so next to this file.

    Safe to call from several processes at once (torchrun ranks): the build is serialised with a file
    lock, skipped when the content hash of sources + flags matches the stamp of the existing library,
    and the library is replaced atomically."""
    override = os.environ.get("TIF_B200_LIB")
    if override:            # an experiment variant built ahead of time (build_variant): use it as it is
        if not os.path.exists(override):
            raise RuntimeError("TIF_B200_LIB=%s does not exist" % override)
        return override
    os.makedirs(BUILD_DIR, exist_ok=True)
    defs = os.environ.get("TIF_NVCC_DEFS", "").split()        # experiment knobs, e.g. -DTIF_STITCH_MIN_BLOCKS=6
    sources = [os.path.join(CSRC, u) for u in UNITS] + [os.path.join(CSRC, "tif_common.cuh"), os.path.join(CSRC, "tif_bilinear.cuh"),
                                                        os.path.join(INCLUDE, "tif_b200.h")]
    want = _digest(sources, (ARCH, COMMON[:4], sorted(UNITS.items()), defs))
    stamp = os.path.join(BUILD_DIR, "libtif_b200.stamp")
    with open(os.path.join(BUILD_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == want:
                return LIB_PATH
            nvcc = _nvcc()
            objects = []
            for unit, extra in UNITS.items():
                src = os.path.join(CSRC, unit)
                obj = os.path.join(BUILD_DIR, unit.replace(".cu", ".o"))
                objects.append(obj)
                cmd = [nvcc] + ARCH + COMMON + extra + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
                res = subprocess.run(cmd, capture_output=True, text=True)
                if verbose or res.returncode != 0:
                    print(" ".join(cmd))
                    print(res.stdout + res.stderr)
                if res.returncode != 0:
                    raise RuntimeError("nvcc failed for %s" % unit)
            tmp = LIB_PATH + ".tmp.%d" % os.getpid()
            cmd = [nvcc] + ARCH + ["-shared", "-o", tmp] + objects + ["-lcudart"]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                print(res.stdout + res.stderr)
                raise RuntimeError("nvcc link failed")
            os.replace(tmp, LIB_PATH)
            with open(stamp, "w") as f:
                f.write(want)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def build_variant(name, defs, units=("tif_exec.cu", "tif_plan.cu", "tif_warp.cu")):
    """A/B experiments: compile `units` with extra -D knobs and link build/variants/<name>.so, reusing the objects of
    the regular build for the other units.  Select it at run time with TIF_B200_LIB=<path> (see _lib.LIB_PATH), so the
    GPU box spends no time compiling."""
    build_library()
    vdir = os.path.join(BUILD_DIR, "variants", name)
    os.makedirs(vdir, exist_ok=True)
    nvcc = _nvcc()
    objects = []
    for unit, extra in UNITS.items():
        obj = os.path.join(BUILD_DIR, unit.replace(".cu", ".o"))
        if unit in units:
            obj = os.path.join(vdir, unit.replace(".cu", ".o"))
            cmd = [nvcc] + ARCH + COMMON + extra + list(defs) + ["-c", os.path.join(CSRC, unit), "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                print(res.stdout + res.stderr)
                raise RuntimeError("nvcc failed for %s (%s)" % (unit, name))
        objects.append(obj)
    out = os.path.join(BUILD_DIR, "variants", name + ".so")
    res = subprocess.run([nvcc] + ARCH + ["-shared", "-o", out] + objects + ["-lcudart"], capture_output=True, text=True)
    if res.returncode != 0:
        print(res.stdout + res.stderr)
        raise RuntimeError("nvcc link failed (%s)" % name)
    return out


if __name__ == "__main__":
    import sys
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2].split()))
    else:
        print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
