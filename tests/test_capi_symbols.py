"""The C-ABI shared library loads here (no GPU) and exports every symbol include/ssme_b200.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(ssme_b200_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    import ssme_b200 as sb
    lib = sb.load_library()
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "libssme_b200.so does not export %s" % n
    assert b"sm_100a" in lib.ssme_b200_build_info()


def test_no_cpu_fallback_without_a_gpu():
    """Without a CUDA device the product fails loudly instead of computing on the host."""
    import ssme_b200 as sb
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(sb.SsmeB200Error) as ei:
        sb.ParticleFilterBackend(sb.FilterConfig(num_particles=64))
    assert "no CPU fallback" in str(ei.value)


def test_argument_validation_happens_before_device_use():
    import ssme_b200 as sb
    with pytest.raises(ValueError):
        sb.ParticleFilterBackend(sb.FilterConfig(num_particles=0))
    with pytest.raises(ValueError):
        sb.ParticleFilterBackend(sb.FilterConfig(model=7))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under ssme_b200/ or include/ may reference it."""
    bad = []
    for base in ("ssme_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            if os.path.basename(dp) in ("build", "lib", "__pycache__"):
                continue
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, fn)).read()
                    if re.search(r"(import|from)\s+oracle|#include\s+[\"<][^\">]*oracle/|libssme_oracle", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
