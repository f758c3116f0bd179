"""CPU-side checks of the C-ABI boundary: the library loads, exports every declared symbol, and the
product path fails loudly (no fallback) when there is no B200."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def L():
    from infinicube_b200 import _lib, build
    if not _lib.LIB_PATH.exists():
        build.build()
    return _lib.lib()


def test_exports_every_declared_symbol(L):
    from infinicube_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_header_has_no_torch_types_and_cites_reference():
    text = (ROOT / "include" / "infinicube_b200.h").read_text()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # declarations only: no torch / ATen / c10 types in signatures
    assert not re.search(r"torch|at::|c10::|Tensor", code)
    assert 'extern "C"' in text
    for cite in ("videogen/inference.py", "utils/fvdb_utils.py", "camera/base.py", "utils/buffer_utils.py"):
        assert cite in text


def test_error_strings(L):
    assert L.ic_error_string(0) == b"ok"
    assert b"no fallback" in L.ic_error_string(-3)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_fails_loudly_without_device(L):
    from infinicube_b200 import _lib
    assert L.ic_device_check() == -3
    with pytest.raises(_lib.ICError):
        _lib.require_device()
    cfg = _lib.DitConfig(1536, 8960, 12, 1, 16, 16, 4096, 256, 512, 32, 1e-6, 8, 32, 32, 0, 8, 1, 0)
    h = C.c_void_p()
    assert L.ic_dit_create(C.byref(cfg), C.byref(h)) == -3
    ep = _lib.GemmEpilogue()
    assert L.ic_gemm_bf16(C.c_void_p(16), 8, C.c_void_p(16), 8, 8, 8, 8, C.byref(ep), None) == -3
    from infinicube_b200.raster import VoxelGrid
    with pytest.raises(_lib.ICError):
        VoxelGrid(torch.zeros(4, 3))
    from infinicube_b200.videogen import WanVideoGenerator
    with pytest.raises(_lib.ICError):
        WanVideoGenerator("x.safetensors", synthetic_weights=True, use_wan_1pt3b=True)


def test_product_never_imports_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle/|raster_oracle|wan_dit_oracle|wan_vae_oracle|umt5_oracle|knn_oracle|mesh_oracle", re.M)
    offenders = []
    for f in (ROOT / "infinicube_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h") and pat.search(f.read_text().replace("oracle/raster_oracle.c", "")):
            offenders.append(str(f))
    assert not offenders, offenders


def test_sass_is_blackwell_native():
    """tcgen05 / TMA must be in the shipped SASS (B200_PROFILING.md: UTC*MMA, LDTM, UTMALDG)."""
    from infinicube_b200 import _lib
    r = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = r.stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16" not in sass  # no legacy mma.sync path
