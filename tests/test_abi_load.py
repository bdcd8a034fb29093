"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/bioranges.h declares,
and refuses to run (loudly, no fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "bioranges.h")).read()
    return sorted(set(re.findall(r"BRG_API[^;(]*?\b(brg_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    import bio_ranges_b200 as brg

    assert _header_symbols() == sorted(brg._cabi.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    import bio_ranges_b200 as brg

    L = brg._cabi.lib()
    for name in _header_symbols():
        assert hasattr(L, name), f"libbioranges.so does not export {name}"
    assert L.brg_version() >= 100


def test_product_does_not_touch_oracle():
    """The product package must not import, load or call anything under oracle/."""
    pkg = os.path.join(ROOT, "datafusion-bio-functions_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "brgo_" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import bio_ranges_b200 as brg

    with pytest.raises(brg.BioRangesError) as ei:
        brg.Context(0)
    assert ei.value.status == 2  # BRG_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)
