"""Build tests/emu/libhyprec_emu.so: the SIMT kernels compiled for the CPU (test infrastructure)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libhyprec_emu.so")


def build():
    src = os.path.join(HERE, "emu_kernels.cpp")
    csrc = os.path.join(os.path.dirname(os.path.dirname(HERE)), "hyprec_b200", "csrc")
    deps = [src, os.path.join(HERE, "emu_cuda.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", HERE, "-o", LIB, src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("g++ failed:\n" + proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build())
