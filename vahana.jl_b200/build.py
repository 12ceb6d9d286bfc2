"""Builds the CUDA engine in-tree for sm_100a: vahana.jl_b200/csrc/build/libvahana_b200.so.
nvcc cross-compiles without a GPU; the .so travels to the GPU box with the gpurun snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(CSRC, "build")
OUT = os.path.join(OUT_DIR, "libvahana_b200.so")
SOURCES = [os.path.join(CSRC, "engine", "engine.cu"), os.path.join(CSRC, "transitions", "builtin.cu"),
           os.path.join(CSRC, "transitions", "builtin_tests.cu"),
           os.path.join(CSRC, "workloads", "generators.cu")]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("VB_NVCC_EXTRA", "").split()   # e.g. -DVB_BLK_MINCTAS=8 for tuning runs


def _deps():
    deps = list(SOURCES)
    for root in (os.path.join(CSRC, "engine"), os.path.join(CSRC, "transitions"), os.path.join(CSRC, "workloads"),
                 os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh", ".inc", ".cu")):
                deps.append(os.path.join(root, f))
    return deps


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    from concurrent.futures import ThreadPoolExecutor
    newest = max(os.path.getmtime(d) for d in _deps())

    def compile_one(src):
        obj = os.path.join(OUT_DIR, os.path.basename(src) + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < newest:
            cmd = ["nvcc", *FLAGS, "-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            with open(obj + ".log", "w") as f:
                f.write(r.stdout + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
            if verbose:
                print(r.stderr[-2000:])
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:     # the translation units compile concurrently
        objs = list(ex.map(compile_one, SOURCES))
    if force or not os.path.exists(OUT) or any(os.path.getmtime(o) > os.path.getmtime(OUT) for o in objs):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs, "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
