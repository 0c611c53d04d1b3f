"""TEST INFRASTRUCTURE: build the CPU SIMT-emulated replica of libsalt_b200 (see cuda_shim.h)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libsalt_b200_emul.so")


def build(force=False):
    csrc = os.path.join(HERE, "..", "..", "salt_b200", "csrc")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(HERE, f) for f in ("cuda_shim.h", "emul_lib.cpp")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-fvisibility=hidden",
                           "-Wno-unused-function", "-o", OUT, os.path.join(HERE, "emul_lib.cpp")])
    return OUT


HOST_OUT = os.path.join(HERE, "libsalt_host_emul.so")


def build_host(force=False):
    """the host-side C layer linked against the emulated engine (CPU tests of the host logic)"""
    import sys
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from salt_b200 import build as b
    return b.build_host(force=force, engine=build(), out=HOST_OUT)


if __name__ == "__main__":
    print(build(True))
