"""In-tree build of libmrblip_b200.so (nvcc, sm_100a only).  `python -m mr_blip_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmrblip_b200.so")
SOURCES = ["abi.cu", "gemm.cu", "gemm2.cu", "attention.cu", "attention_tc.cu", "attention_tc_bwd.cu", "wgrad_tc.cu", "elementwise.cu", "dropout.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-cudart", "shared"]


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force=False, verbose=True, variant="", extra_flags=()):
    """variant: suffix of an experimental second library (own object dir), e.g. build(variant="_hint", extra_flags=["-DMRB_WAIT_HINT_NS=20000"])
    -> libmrblip_b200_hint.so, loaded when MRB_LIB_VARIANT=_hint (A/B measurements on one box)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if variant:
        return _build_variant(nvcc, variant, list(extra_flags), verbose)
    objs, dirty = [], force
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(src, obj) or any(_newer(h, obj) for h in hdrs):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
            dirty = True
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed on %s" % s)
    if dirty or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-cudart", "shared", "-o", LIB] + objs + ["-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


def _build_variant(nvcc, variant, extra_flags, verbose):
    lib = os.path.join(HERE, "libmrblip_b200%s.so" % variant)
    odir = os.path.join(HERE, "build", variant.strip("_") or "variant")
    os.makedirs(odir, exist_ok=True)
    procs, objs = [], []
    for s in SOURCES:
        obj = os.path.join(odir, s.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((s, subprocess.Popen([nvcc] + NVCC_FLAGS + extra_flags + ["-c", os.path.join(CSRC, s), "-o", obj],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed on %s" % s)
    subprocess.check_call([nvcc, "-shared", "-cudart", "shared", "-o", lib] + objs + ["-Xlinker", "-rpath=/usr/local/cuda/lib64"])
    if verbose:
        print("built", lib)
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        build(variant=sys.argv[i + 1], extra_flags=sys.argv[i + 2:])
    else:
        build(force="--force" in sys.argv)
