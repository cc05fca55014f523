"""The C-ABI library loads on a machine without a GPU, exports every symbol include/sharp_b200.h declares, and fails
loudly (no CPU fallback) when a compute call is attempted without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import sharp_b200
from sharp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sharp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sharp_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    decl = declared_symbols()
    assert len(decl) >= 30
    assert sorted(_lib.EXPORTS) == decl


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    hdr = open(os.path.join(ROOT, "include", "sharp_b200.h")).read()
    assert lib.sharp_abi_version() == int(re.search(r"#define SHARP_B200_ABI_VERSION (\d+)", hdr).group(1)) == 4


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "sharp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)   # declarations only
    assert "torch" not in src.lower() and "at::" not in src and "std::" not in src
    assert 'extern "C"' in src


def test_kernel_classes_are_named():
    lib = _lib.load()
    names = [lib.sharp_prof_name(i).decode() for i in range(lib.sharp_prof_kernels())]
    assert "rp_project" in names and "corrdist" in names and "hclust" in names and all(names)


def test_product_does_not_import_the_oracle():
    # a product path that routes through oracle/ would void every parity claim
    pkg = os.path.join(ROOT, "sharp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "sharp_oracle" not in txt and "import orc" not in txt and "oracle/" not in txt, f


@pytest.mark.skipif(_lib.load().sharp_device_count() > 0, reason="a CUDA device is present")
def test_compute_fails_loudly_without_a_device():
    assert sharp_b200.device_count() == 0
    with pytest.raises(sharp_b200.SharpError) as ei:
        sharp_b200.Context(0)
    assert ei.value.code == _lib.E_CUDA and "no CPU fallback" in str(ei.value)
    with pytest.raises(sharp_b200.SharpError):
        sharp_b200.SHARP(np.ones((10, 20)), rN_seed=1)


def test_missing_library_is_an_import_error(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", os.path.join(ROOT, "sharp_b200", "does_not_exist.so"))
    with pytest.raises(ImportError) as ei:
        _lib.load()
    assert "no CPU fallback" in str(ei.value)


def test_host_only_entry_points_run_without_a_device():
    # the R-compatible generators are pure host code
    out = np.empty(10, dtype=np.int64)
    assert _lib.load().sharp_r_sample_perm(C.c_int64(42), C.c_int64(10), out.ctypes.data_as(C.POINTER(C.c_int64))) == 0
    assert list(out) == [1, 5, 10, 8, 2, 4, 6, 9, 7, 3]
    r = _lib.r_ranm(300, 12, 7)
    assert r["p"][-1] == len(r["i"]) == len(r["x"])


def test_r_glue_compiles_against_the_r_api():
    """r_glue/sharp_r_glue.c (the .Call stubs a maintainer adds to the R package) cannot be built
    without R; its syntax, its use of the C ABI (every sharp_* call must match include/sharp_b200.h) and its use of the
    R C API (tests/r_stub: declarations only, taken from the published R API) are checked with the host compiler."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no host C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([gcc, "-fsyntax-only", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
                        "-I" + os.path.join(root, "tests", "r_stub"), "-I" + os.path.join(root, "include"),
                        os.path.join(root, "r_glue", "sharp_r_glue.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(root, "r_glue", "sharp_r_glue.c")).read()
    for stub in ("sharp_R_run", "sharp_R_run_parts", "sharp_R_rp_project", "sharp_R_opt_hclust", "sharp_R_wmetac",
                 "sharp_R_smetac", "sharp_R_centroids", "sharp_R_smetac_centroids", "sharp_R_rm_upload"):
        assert '{"%s"' % stub in src, stub + " is not registered in call_methods"
