"""The plugin story end to end: a model library compiled outside the engine (tests/ext_model/ext_model.cu: include the device API,
define functors, VB_REGISTER_TRANSITION / VB_REGISTER_MAP) is built with nvcc for sm_100a, loaded with vb_load_model_library and its
transitions are found by name.  Compiling and registering need no GPU; applying does (GPU variant: tests/test_zzr_ext_model_gpu.py)."""
import os
import subprocess

import vahana_b200 as vh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "ext_model", "ext_model.cu")
OUT = os.path.join(ROOT, "tests", "ext_model", "build", "libext_model.so")


def build_ext_model():
    engine_dir = os.path.dirname(vh.DEFAULT_LIB)
    deps = [SRC, os.path.join(ROOT, "include", "vahana_device.cuh"), os.path.join(ROOT, "include", "vahana_model.h")]
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        cmd = ["nvcc", "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "--expt-relaxed-constexpr", "-lineinfo",
               "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"), SRC, "-o", OUT,
               "-L", engine_dir, "-lvahana_b200", "-Xlinker", "-rpath", "-Xlinker", engine_dir]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    return OUT


def test_external_model_library_builds_and_registers():
    if not os.path.exists(vh.DEFAULT_LIB):
        import __graft_entry__
        __graft_entry__.build()
    be = vh.load_backend()                                   # the CUDA engine library loads without a GPU; only vb_init needs one
    assert be.name == "cuda-sm100a"
    assert not be.has_transition("ext_add_one_plus_degree", "AMortal")
    be.load_model_library(build_ext_model())
    assert be.has_transition("ext_add_one_plus_degree", "AMortal") and not be.has_transition("ext_add_one_plus_degree", "AImm")
    # the model library only needs the engine's registration entry points
    nm = subprocess.run(["nm", "-D", "--undefined-only", OUT], capture_output=True, text=True).stdout
    assert "vb_register_transition" in nm and "vb_register_map" in nm
    try:
        be.load_model_library(os.path.join(ROOT, "tests", "ext_model", "does_not_exist.so"))
        raise SystemExit("expected an error")
    except RuntimeError as e:
        assert "dlopen" in str(e)
