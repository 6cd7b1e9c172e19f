"""Build the C-ABI shared library (lib/liblgdwt_b200.so) with plain nvcc for sm_100a.

No torch headers are involved, so a full rebuild takes seconds.  Never add --use_fast_math: the blend kernels
rely on libdevice expf and IEEE div/sqrt to stay bit-compatible with the reference build.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIBNAME = os.environ.get("LGDWT_LIBNAME", "liblgdwt_b200.so")  # tuning variants: another name + LGDWT_EXTRA_FLAGS

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
] + os.environ.get("LGDWT_EXTRA_FLAGS", "").split()
if LIBNAME != "liblgdwt_b200.so":
    OBJDIR = os.path.join(HERE, "build", LIBNAME)


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(path, extra):
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for dep in extra + [path]:
        with open(dep, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _compile(src, verbose):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJDIR, src[:-3] + ".o")
    headers = [os.path.join(CSRC, h) for h in sorted(os.listdir(CSRC)) if h.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "lgdwt_b200.h"))
    stamp = _stamp(path, headers)
    stamp_file = obj + ".stamp"
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, ""
    cmd = ["nvcc"] + NVCC_FLAGS + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return obj, r.stderr if verbose else ""


def build(verbose=False, force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    srcs = _sources()
    objs, logs = [], []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for obj, log in ex.map(lambda s: _compile(s, verbose), srcs):
            objs.append(obj)
            logs.append(log)
    lib = os.path.join(LIBDIR, LIBNAME)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(lib) or os.path.getmtime(lib) < newest:
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        sys.stderr.write("\n".join(logs))
    return lib


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
