"""Runs a reference script UNCHANGED (runpy, as `python script.py args...` would) and afterwards reports which
shared objects of this repository / of the reference install the process has mapped, so that a test can assert that
the run really went through the intended native library and not through the other arm's.

    python tests/run_unchanged.py REPORT.json SCRIPT.py [script args...]
"""
import json
import os
import runpy
import sys


def main():
    report, script = sys.argv[1], os.path.abspath(sys.argv[2])
    sys.argv = [script] + sys.argv[3:]
    sys.path.insert(0, os.path.dirname(script))   # what `python script.py` does
    status = "ok"
    try:
        runpy.run_path(script, run_name="__main__")
    except SystemExit as e:
        if e.code not in (None, 0):
            status = "exit %r" % (e.code,)
    finally:
        libs = set()
        with open("/proc/self/maps") as f:
            for line in f:
                path = line.split()[-1]
                if path.endswith(".so") and ("/baseline/_ref/" in path or "/oracle/_ref/" in path or
                                             "liblgdwt_b200" in path):
                    libs.add(path)
        mods = {name: getattr(sys.modules[name], "__file__", None)
                for name in ("diff_gaussian_rasterization", "simple_knn._C", "pytorch_wavelets", "fused_ssim", "plyfile")
                if name in sys.modules}
        with open(report, "w") as f:
            json.dump({"status": status, "native_libraries": sorted(libs), "modules": mods}, f, indent=1)
    if status != "ok":
        sys.exit(1)


if __name__ == "__main__":
    main()
