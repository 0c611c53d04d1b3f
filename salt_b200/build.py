"""Build libsalt_b200.so (sm_100a only) in-tree with nvcc.  `python -m salt_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsalt_b200.so")
SOURCES = ["engine.cu", "verify.cu", "ssw.cu", "samtail.cu", "mixref.cu", "transport.cu", "seed.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "--use_fast_math"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "salt_b200.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + FLAGS + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd)))
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed on " + src)
    if force or procs or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-Xlinker", "--exclude-libs,ALL"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    build_host(force or bool(procs))
    return OUT


HOST_OUT = os.path.join(HERE, "libsalt_host.so")


def build_host(force=False, engine=OUT, out=HOST_OUT):
    """libsalt_host.so: the host-side C layer (include/salt_host.h), gcc, linked against the engine."""
    src = os.path.join(HERE, "host", "salt_host.c")
    hdrs = [os.path.join(HERE, "..", "include", f) for f in ("salt_host.h", "salt_b200.h")]
    if force or _stale(out, [src, engine] + hdrs):
        d, f = os.path.split(engine)
        subprocess.check_call([os.environ.get("CC", "gcc"), "-O2", "-std=gnu11", "-Wall", "-Wextra", "-fPIC", "-shared",
                               "-o", out, src, "-L" + d, "-l:" + f, "-Wl,-rpath," + d, "-Wl,-rpath,$ORIGIN", "-lpthread"])
    l0 = os.path.join(os.path.dirname(out), "libsalt_level0" + ("_emul" if out.endswith("_emul.so") else "") + ".so")
    src0 = os.path.join(HERE, "host", "level0_shim.c")
    if force or _stale(l0, [src0, engine] + hdrs):
        d, f = os.path.split(engine)
        subprocess.check_call([os.environ.get("CC", "gcc"), "-O2", "-std=gnu11", "-Wall", "-Wextra", "-fPIC", "-shared",
                               "-fvisibility=hidden", "-o", l0, src0, "-L" + d, "-l:" + f, "-Wl,-rpath," + d, "-Wl,-rpath,$ORIGIN", "-lpthread"])
    build_aln(force, engine=engine, hostlib=out)
    return out


def build_aln(force=False, engine=OUT, hostlib=HOST_OUT):
    """salt_aln: the aligner as a program of its own (host/salt_aln.c) over libsalt_host.so + libsalt_b200.so; next to the host
    library it links (salt_b200/salt_aln, or tests/emul/salt_aln_emul over the emulated engine)."""
    src = os.path.join(HERE, "host", "salt_aln.c")
    emul = hostlib.endswith("_emul.so")
    out = os.path.join(os.path.dirname(hostlib), "salt_aln_emul" if emul else "salt_aln")
    hdrs = [os.path.join(HERE, "..", "include", f) for f in ("salt_host.h", "salt_b200.h")]
    if force or _stale(out, [src, engine, hostlib] + hdrs):
        d, f = os.path.split(engine)
        dh, fh = os.path.split(hostlib)
        subprocess.check_call([os.environ.get("CC", "gcc"), "-O2", "-std=gnu11", "-Wall", "-Wextra", "-I" + os.path.join(HERE, "..", "include"),
                               "-o", out, src, "-L" + dh, "-l:" + fh, "-L" + d, "-l:" + f, "-Wl,-rpath," + d, "-Wl,-rpath," + dh,
                               "-Wl,-rpath,$ORIGIN", "-lpthread", "-lz"])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
