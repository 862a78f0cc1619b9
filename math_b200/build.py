"""In-tree build of libstanmath_cuda.so (nvcc, -gencode arch=compute_100a,code=sm_100a)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False, jobs=None):
    jobs = jobs or os.cpu_count() or 4
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), f"-j{jobs}"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True)
    if verbose or res.returncode:
        print(res.stdout)
    if res.returncode:
        raise RuntimeError("building libstanmath_cuda.so failed")
    return os.path.join(HERE, "lib", "libstanmath_cuda.so")
