"""In-tree build of libmrblip_b200.so (nvcc, sm_100a only).  `python -m mr_blip_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmrblip_b200.so")
SOURCES = ["abi.cu", "gemm.cu", "gemm2.cu", "attention.cu", "attention_tc.cu", "attention_tc_bwd.cu", "attention_vit.cu", "wgrad_tc.cu", "elementwise.cu", "dropout.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-cudart", "shared"]


def _digest(src, hdrs, flags):
    """Content hash of everything one object file is made from (source, every header of csrc/, compiler flags): an object is
    reused only when the hash stored next to it matches -- file times play no role (a checkout or a copied tree resets them)."""
    import hashlib
    h = hashlib.sha256(" ".join(flags).encode())
    for f in [src] + sorted(hdrs):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()


def _stale(obj, digest):
    try:
        return (not os.path.exists(obj)) or open(obj + ".sha256").read().strip() != digest
    except OSError:
        return True


def build(force=False, verbose=True, variant="", extra_flags=()):
    """variant: suffix of an experimental second library (own object dir), e.g. build(variant="_hint", extra_flags=["-DMRB_WAIT_HINT_NS=20000"])
    -> libmrblip_b200_hint.so, loaded when MRB_LIB_VARIANT=_hint (A/B measurements on one box)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if variant:
        return _build_variant(nvcc, variant, list(extra_flags), verbose)
    force = force or os.environ.get("MRB_FORCE_BUILD", "0") == "1"
    objs, dirty = [], force
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(obj)
        digest = _digest(src, hdrs, NVCC_FLAGS)
        if force or _stale(obj, digest):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            if os.path.exists(obj + ".sha256"):
                os.remove(obj + ".sha256")
            procs.append((s, obj, digest, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
            dirty = True
    for s, obj, digest, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed on %s" % s)
        open(obj + ".sha256", "w").write(digest)
    if verbose and not procs:
        print("every object matches the content hash of its sources: nothing to compile", flush=True)
    if dirty or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
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
