"""CPU: the C-ABI library loads and exports exactly what include/egx.h declares; the product
path refuses to run without CUDA (no fallback)."""
import os
import re
import subprocess

import pytest
import torch

from emotiongestures_b200 import TED, Transformer, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "egx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(egx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from emotiongestures_b200 import build
    build.build()
    lib = _lib.load_library()
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/egx.h but not exported"
        assert s in _lib.PROTOTYPES, f"{s} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == syms
    assert lib.egx_version() == 1


def test_exports_are_plain_c_and_nothing_else_leaks():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert exported == set(_header_symbols())


def test_sass_contains_blackwell_tensor_and_tma_ops():
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):      # tcgen05.mma, cp.async.bulk.tensor, tcgen05.ld
        assert mnemonic in sass, mnemonic
    assert "HMMA.16" not in sass                           # no legacy mma.sync path


def test_no_cpu_fallback():
    from emotiongestures_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(TED, "cpu")
    if not torch.cuda.is_available():
        m = Transformer.from_config(TED).eval()
        with pytest.raises(RuntimeError):
            m(torch.zeros(1, 128, 70), torch.zeros(1, 60, dtype=torch.int64), torch.zeros(1, 4, 126))


def test_create_fails_cleanly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ctypes as C
    lib = _lib.load_library()
    cfg = _lib.EgxCfg(34, 4, 126, 256, 1024, 3, 8, 64, 64, 128, 70, 60, 1)
    h = C.c_void_p()
    assert lib.egx_create(C.byref(cfg), 0, C.byref(h)) != 0 and not h.value
    assert lib.egx_last_error(None) == b"null handle"
    assert lib.egx_workspace_bytes(None, 4) == 0 and lib.egx_launch_count(None) == 0


def test_missing_library_is_loud(tmp_path):
    with pytest.raises(RuntimeError, match="no fallback"):
        _lib.load_library(str(tmp_path / "libegx.so"))
