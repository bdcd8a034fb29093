"""B200-native interval-overlap engine behind the API surface of datafusion-bio-function-ranges.

Layout of the package (only what the hot path needs):
  csrc/            hand-written CUDA for sm_100a + the C ABI (include/bioranges.h) -> libbioranges.so
  _cabi.py         ctypes binding of the C ABI (what a Rust `extern "C"` block would bind)
  engine.py        Context / Index: thin object wrappers over the ABI (numpy host buffers, torch device tensors)

The product path fails loudly when libbioranges.so is missing or no sm_100 GPU is present; nothing here imports
or calls oracle/.
"""
from . import _cabi
from ._cabi import (JOIN_ORDERED, KEY_ABSENT, PAIRS_ANTI, PAIRS_INNER, PAIRS_INNER_ANY_ORDER, PAIRS_SEMI, WANT_ALL, WANT_COUNT, WANT_COVERAGE, WANT_NEAREST,
                    WANT_PAIRS, BioRangesError)
from .engine import Context, Index, default_context

__all__ = ["Context", "Index", "default_context", "BioRangesError", "KEY_ABSENT", "WANT_ALL", "WANT_COUNT",
           "WANT_COVERAGE", "WANT_NEAREST", "WANT_PAIRS", "PAIRS_INNER", "PAIRS_SEMI", "PAIRS_ANTI", "PAIRS_INNER_ANY_ORDER", "JOIN_ORDERED", "_cabi"]
