"""Build recipes for the oracle (TEST INFRASTRUCTURE, not product code).

* build_oracle(): gcc -> oracle/_build/liboracle.so from oracle/quant_oracle.c
* build_ref():    compiles the REFERENCE's own quant_cuda extension from the
  sources where they lie under /root/reference (never copied into this repo)
  into oracle/_ref/quant_cuda*.so.  Only possible in the build container
  (/root/reference does not exist on the GPU box; the built .so travels with
  the gpurun snapshot because oracle/_ref/ is git-ignored but not
  gpurun-ignored).  The reference build system (setup.py install) is not run;
  the two source files are compiled directly with torch.utils.cpp_extension.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
REF_DIR = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/AdaQP/util/quantization/src"


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_oracle(verbose: bool = False) -> str:
    os.makedirs(BUILD_DIR, exist_ok=True)
    src = os.path.join(HERE, "quant_oracle.c")
    out = os.path.join(BUILD_DIR, "liboracle.so")
    if _stale(out, [src]):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
               "-o", out, src, "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return out


def ref_available() -> bool:
    if not os.path.isdir(REF_DIR):
        return False
    return any(f.startswith("quant_cuda") and f.endswith(".so") for f in os.listdir(REF_DIR))


def build_ref(verbose: bool = False) -> str | None:
    """Compile the reference quant_cuda for sm_100a into oracle/_ref/."""
    if not os.path.isdir(REF_SRC):
        return None
    if ref_available():
        return REF_DIR
    os.makedirs(REF_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils import cpp_extension
    tmp = os.path.join(BUILD_DIR, "ref_build")
    os.makedirs(tmp, exist_ok=True)
    cpp_extension.load(
        name="quant_cuda",
        sources=[os.path.join(REF_SRC, "quantization.cc"),
                 os.path.join(REF_SRC, "quantization_cuda_kernel.cu")],
        extra_cuda_cflags=["--expt-extended-lambda"],
        build_directory=tmp,
        verbose=verbose,
        is_python_module=False,
    )
    import shutil
    for f in os.listdir(tmp):
        if f.endswith(".so"):
            shutil.copy2(os.path.join(tmp, f), os.path.join(REF_DIR, f))
    return REF_DIR


def load_ref():
    """Import the reference-built quant_cuda module (GPU box or container)."""
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    for f in sorted(os.listdir(REF_DIR)):
        if f.startswith("quant_cuda") and f.endswith(".so"):
            spec = importlib.util.spec_from_file_location("quant_cuda", os.path.join(REF_DIR, f))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    raise ImportError("oracle/_ref/quant_cuda*.so not built")


if __name__ == "__main__":
    print(build_oracle(verbose=True))
    if "--ref" in sys.argv:
        print(build_ref(verbose=True))
