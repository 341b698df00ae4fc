"""The C-ABI library loads and exports every symbol include/acb200.h declares (no compute calls)."""
import os
import re

from pyahocorasick_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "acb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(acb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    L = N.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/acb200.h but not exported"
    assert sorted(N.EXPORTED_SYMBOLS) == names
    assert L.acb_abi_version() == N.ABI_VERSION == 2


def test_host_calls_work_without_a_gpu_and_scan_fails_loudly():
    import ctypes
    import numpy as np
    import pytest
    L = N.lib()
    t = L.acb_trie_new(1)
    prev = ctypes.c_int32()
    assert L.acb_trie_add_word(t, b"abc", 3, 0, ctypes.byref(prev)) == 0 and prev.value == -1
    assert L.acb_trie_add_word(t, b"abc", 3, 5, ctypes.byref(prev)) == 0 and prev.value == 0
    assert L.acb_trie_add_word(t, b"", 0, 1, ctypes.byref(prev)) == 0 and prev.value == -2
    assert L.acb_trie_kind(t) == 1 and L.acb_trie_count(t) == 1
    fv = N.FlatView()
    assert L.acb_trie_flat_view(t, ctypes.byref(fv)) == N.ACB_ESTATE          # not built yet
    built = ctypes.c_int32()
    assert L.acb_trie_make_automaton(t, ctypes.byref(built)) == 0 and built.value == 1
    assert L.acb_trie_make_automaton(t, ctypes.byref(built)) == 0 and built.value == 0
    assert L.acb_trie_flat_view(t, ctypes.byref(fv)) == 0 and fv.n_states == 4 and fv.n_keys == 6
    import torch
    if not torch.cuda.is_available():
        tb = ctypes.c_void_p()
        rc = L.acb_table_upload(t, 0, ctypes.byref(tb))
        assert rc == N.ACB_ECUDA and N.last_error()                            # no silent CPU fallback
        with pytest.raises(N.NativeError):
            N.check(rc)
    L.acb_trie_free(t)
    assert L.acb_trie_new(3) is None and "letter_bytes" in N.last_error()
