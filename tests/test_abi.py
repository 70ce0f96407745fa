"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly the symbols include/tactilesr_b200.h declares
(no compute calls -- there is no GPU here); the header is valid C; the product package never imports the oracle."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tactilesr_b200.h")


def _header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsr_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from tactilesr_b200.csrc import build
    lib_path = build.build()
    assert os.path.exists(lib_path)
    from tactilesr_b200 import _lib
    L = _lib.lib()
    declared = _header_symbols()
    assert declared, "no declarations parsed from the header"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES.keys()) == declared, (set(_lib.SIGNATURES) ^ set(declared))
    assert L.tsr_version() >= 100
    assert isinstance(L.tsr_last_error(), bytes)


def test_header_is_valid_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "tactilesr_b200.h"\nint main(void) { return (int)sizeof(&tsr_version) == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(c)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP (B200_PROFILING.md)."""
    from tactilesr_b200 import _lib
    r = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    for mnem in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP"):
        assert mnem in r.stdout, mnem
    assert "HGMMA" not in r.stdout
    # the fp32 (<= 1e-5) mode: packed fp32 FMAs (fma.rn.f32x2 -> FFMA2) and register-free global -> shared copies
    # (cp.async -> LDGSTS) in the conv / weight-gradient / head / tail kernels
    assert r.stdout.count("FFMA2") > 1000
    assert "LDGSTS" in r.stdout


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tactilesr_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dp, f)


def test_cpu_tensors_are_rejected_not_silently_computed():
    import torch
    from tactilesr_b200 import TsrError
    from tactilesr_b200.model import TactileSR, tPSFNet
    with pytest.raises(TsrError):
        TactileSR()(torch.zeros(1, 3, 4, 4))
    with pytest.raises(TsrError):
        tPSFNet(1.4, None, device="cpu")(torch.zeros(1, 3, 4, 4), torch.zeros(1, 1, 100, 100))


def test_unsupported_configurations_raise_instead_of_computing_garbage():
    """Shapes / BatchNorm variants the fused kernels do not implement must fail loudly (they would read the wrong memory)."""
    import torch
    from tactilesr_b200 import TsrError, engine as E
    bn = torch.nn.BatchNorm2d(64, momentum=None)
    with pytest.raises(TsrError):
        E.BNReLUOp(E.View.of(E.Buf("y", 64)), bn, E.View.of(E.Buf("a", 64)))
    with pytest.raises(TsrError):
        E.BNReLUOp(E.View.of(E.Buf("y", 64)), torch.nn.BatchNorm2d(64, affine=False), E.View.of(E.Buf("a", 64)))
