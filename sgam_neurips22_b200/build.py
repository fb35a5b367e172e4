"""Build recipe for libsgam_b200.so (nvcc, sm_100a only, in-tree).

    python -m sgam_neurips22_b200.build [--force]

The shared library is written next to the sources (sgam_neurips22_b200/libsgam_b200.so), is git-ignored and
travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
SO = os.path.join(HERE, "libsgam_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# per-file flags: the bit-exact kernels must not have a*b+c contracted behind their back
SOURCES = {
    "abi.cu": [],
    "splat.cu": ["--fmad=false"],
    "vq.cu": ["--fmad=false"],
    "tsdf.cu": ["--fmad=false"],
    "pcd.cu": ["--fmad=false"],
    "net_simt.cu": [],
    "net_tc.cu": [],
    "net_tc2.cu": [],
    "net_tc3.cu": [],
    "net_attn.cu": [],
    "net_tc_prep.cu": [],
}


def _stale(out, deps):
    return (not os.path.exists(out)) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "sgam_b200.h"))
    objs = []
    for src, extra in SOURCES.items():
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = ["nvcc", *ARCH, *COMMON, *extra, "-c", path, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log = r.stdout + r.stderr
            with open(obj + ".log", "w") as f:
                f.write(" ".join(cmd) + "\n" + log)
            if r.returncode != 0:
                sys.stderr.write(log)
                raise RuntimeError(f"nvcc failed on {src}")
            if verbose:
                print(log)
    if force or _stale(SO, objs):
        cmd = ["nvcc", *ARCH, "-shared", "-o", SO, *objs, "-lcuda"]
        subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
