"""Builds the CUDA C-ABI library in-tree: pysurfinv_b200/libsurfdisp_b200.so (sm_100a only)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", "surfdisp_kernels.cu")]
DEPS = SRC + [os.path.join(HERE, "csrc", "surfdisp_core.cuh"), os.path.join(HERE, "csrc", "sd_libm.cuh"), os.path.join(HERE, "csrc", "surfdisp_mc.cuh"),
              os.path.join(os.path.dirname(HERE), "include", "surfdisp_b200.h")]
OUT = os.path.join(HERE, "libsurfdisp_b200.so")


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def build(force=False, verbose=False, extra=()):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-o", OUT] + list(extra) + SRC
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed")
    with open(os.path.join(HERE, "csrc", "ptxas_report.txt"), "w") as f:
        # (without the compile times: the report is tracked and should only change with the code)
        f.write("".join(l for l in r.stdout.splitlines(True) if "Compile time" not in l))
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
