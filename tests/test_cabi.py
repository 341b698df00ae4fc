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
    assert L.acb_abi_version() == N.ABI_VERSION == 5


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
    st = ctypes.c_int32(7)
    assert L.acb_table_set_long_state(None, 0) == N.ACB_EINVAL                 # iter_long streaming state: no table, no state
    assert L.acb_table_get_long_state(None, ctypes.byref(st)) == N.ACB_EINVAL
    L.acb_trie_free(t)
    assert L.acb_trie_new(3) is None and "letter_bytes" in N.last_error()


def test_pinned_record_holder_releases_once_after_the_last_view():
    """automaton._PinnedRecords: numpy views keep the taken buffer alive; it goes back to the library exactly once"""
    import ctypes
    import gc
    import numpy as np
    from pyahocorasick_b200 import automaton as am

    calls = []

    class FakeLib:
        def acb_release_records(self, p, cap):
            calls.append((p, cap))

    raw = (ctypes.c_int32 * 24)(*range(24))
    a = np.asarray(am._PinnedRecords(FakeLib(), ctypes.addressof(raw), 5, 8))
    assert a.dtype == N.MATCH_DTYPE and a.shape == (5,)
    assert a["hay_id"].tolist() == [0, 3, 6, 9, 12] and a["key_id"].tolist() == [2, 5, 8, 11, 14]
    v, w = a["end_index"], a[1:3]
    del a
    gc.collect()
    assert calls == []
    del v
    gc.collect()
    assert calls == [] and w["end_index"].tolist() == [4, 7]
    del w
    gc.collect()
    assert calls == [(ctypes.addressof(raw), 8)]


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/acb200.h is a C ABI: it compiles as C99 (-pedantic) and a C program links against libacb200.so"""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text(
        '#include "acb200.h"\n'
        "int main(void) {\n"
        "    acb_trie *t = acb_trie_new(1);\n"
        "    int32_t prev = 0, built = 0, kid = -1, pre = 0;\n"
        "    if (!t || acb_abi_version() != ACB_ABI_VERSION) return 1;\n"
        '    if (acb_trie_add_word(t, (const uint8_t *)"he", 2, 0, &prev) != ACB_OK) return 2;\n'
        "    if (acb_trie_make_automaton(t, &built) != ACB_OK || !built || acb_trie_kind(t) != ACB_AHOCORASICK) return 3;\n"
        '    if (acb_trie_find(t, (const uint8_t *)"he", 2, &kid, &pre) != ACB_OK || kid != 0) return 4;\n'
        "    acb_trie_free(t);\n"
        "    return 0;\n"
        "}\n")
    libdir = os.path.dirname(N.LIB_PATH)
    exe = tmp_path / "t"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-lacb200", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True)
    assert subprocess.run([str(exe)]).returncode == 0
