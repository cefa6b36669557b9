"""Build tests/emu/libss_emu.so (CPU emulation of the kernels) with g++.  Test infrastructure."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "ss_emu.cu")
OUT = os.path.join(HERE, "libss_emu.so")
DEPS = [SRC, os.path.join(ROOT, "sonicsim_b200", "csrc", "ss_core.cuh"),
        os.path.join(ROOT, "sonicsim_b200", "csrc", "ss_phases.cuh"),
        os.path.join(ROOT, "sonicsim_b200", "csrc", "ss_loud.cuh")]


def build(force=False):
    if not force and os.path.isfile(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cuda_inc = os.environ.get("CUDA_INC", "/usr/local/cuda/include")
    cmd = ["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-I" + cuda_inc, "-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed: %s\n%s" % (" ".join(cmd), res.stderr))
    return OUT


if __name__ == "__main__":
    print(build(force=True))
