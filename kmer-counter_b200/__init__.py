"""kmer-counter_b200 -- B200-native drop-in for the counting path of jsdjayanga/kmer-counter.

Import name: ``kmer_counter_b200`` (the directory keeps the project's hyphenated
name; ``kmer_counter_b200.py`` at the repository root aliases it).

Layers
  _lib.py      ctypes binding of libkc_b200.so, the C ABI of include/kc_api.h
  engine.py    thin object wrappers (Counter, Run)
  refapi.py    host-side mirror of the reference's interface for this path:
               PrepareGPU / processKMers / FreeGPU, FileDump, KMerFileMerger,
               KMerFileMergeHandler, KMerPrinter -- same names and argument meaning
  multigpu.py  one process per GPU: key-range ownership + all-to-all of run slices
  synth.py     deterministic synthetic reads (host side, numpy)
"""
from . import _lib
from ._lib import KC_COUNT_AUTO, KC_COUNT_HASH, KC_COUNT_SORT, build
from .engine import Counter, KcError, Run, key_words, record_size

__all__ = ["Counter", "Run", "KcError", "key_words", "record_size", "build",
           "KC_COUNT_AUTO", "KC_COUNT_SORT", "KC_COUNT_HASH"]
