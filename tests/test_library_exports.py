"""CPU-side checks of the product library: it loads, exports every symbol include/xmapper_b200.h declares, and refuses to
run without a CUDA device (no CPU fallback)."""
import os
import re

import pytest

from mapper_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "xmapper_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|void|int64_t|const char\*)\s+(xm_[a-z_]+)\s*\(", text, flags=re.M)))


def test_header_symbols_are_exported():
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = capi.load_library()
    syms = declared_symbols()
    assert set(capi.EXPORTS) == set(syms)
    for s in syms:
        assert hasattr(L, s), s


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.XmError):
        capi.XMapper(synth.DEFAULT_PARAMS, device=0)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under mapper_b200/ may import, link or open it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mapper_b200")):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "xm_oracle" not in text and "libxmoracle" not in text and "oracle/" not in text, os.path.join(dirpath, f)
                assert "import xm_emu" not in text and "libxmemu" not in text, os.path.join(dirpath, f)
