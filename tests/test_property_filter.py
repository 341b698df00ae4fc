"""Property test of the host-built tables (hypothesis): for arbitrary key sets and haystacks the emulated filter
path (stage 1/2/3 bitmaps, anchor table, UNIQUE compares, MULTI trie walks -- tests/emul.py restates the kernel
on the real flattened tables) and the emulated DFA path both equal the oracle.  Byte alphabets include 0x00 and
0xFF, keys run from 1 to 40 bytes (beyond the 20 bytes an anchor entry can carry), many keys share grams, and
the letter widths 1, 2 (bytes-flavour sequences) and 4 (unicode) are all covered."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import emul
import oracle
import pyahocorasick_b200 as pkg

ALPHABETS = [b"ab", b"\x00\xff", b"abcdefgh", bytes(range(256))]


@st.composite
def byte_case(draw):
    al = draw(st.sampled_from(ALPHABETS))
    sym = st.sampled_from(list(al))
    lo = draw(st.integers(1, 6))
    hi = draw(st.integers(lo, 40))
    keys = draw(st.lists(st.lists(sym, min_size=lo, max_size=hi).map(bytes), min_size=1, max_size=40, unique=True))
    n_hay = draw(st.integers(1, 5))
    hays = [bytearray(draw(st.lists(sym, min_size=0, max_size=90).map(bytes))) for _ in range(n_hay)]
    for h in hays:                                   # plant a few keys so that matches exist
        if len(h) and draw(st.booleans()):
            k = draw(st.sampled_from(keys))
            p = draw(st.integers(0, len(h)))
            h[p:p + len(k)] = k
    return keys, [bytes(h) for h in hays]


def _oracle_records(keys, hays):
    O = oracle.OracleAutomaton()
    for i, k in enumerate(keys):
        O.add_word(k, i)
    O.make_automaton()
    out = []
    for h, hay in enumerate(hays):
        out += [(h, e, v) for e, v in O.find_all(hay)]
    return out


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(byte_case())
def test_emulated_kernels_equal_the_oracle_bytes(case):
    keys, hays = case
    A = pkg.flavour("bytes").Automaton(pkg.STORE_INTS)
    for i, k in enumerate(keys):
        A.add_word(k, i)
    A.make_automaton()
    f = A.flat()
    assert f["gram_bytes"] + f["stride"] - 1 <= f["min_key_bytes"]
    flat = np.frombuffer(b"".join(hays), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum([len(h) for h in hays])]).astype(np.int64)
    want = _oracle_records(keys, hays)
    assert emul.emul_filter(f, flat, off, 0) == want
    assert emul.emul_dfa(f, flat, off, 0) == want


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.lists(st.text(alphabet="ab\xe9\u0142\U0001f600", min_size=1, max_size=9), min_size=1, max_size=12, unique=True),
       st.text(alphabet="ab\xe9\u0142\U0001f600", min_size=0, max_size=60))
def test_emulated_kernels_equal_the_oracle_unicode(keys, hay):
    A = pkg.flavour("unicode").Automaton(pkg.STORE_INTS)
    for i, k in enumerate(keys):
        A.add_word(k, i)
    A.make_automaton()
    f = A.flat()
    assert f["letter_bytes"] == 4
    flat = np.frombuffer(hay.encode("utf-32-le"), dtype=np.uint8)
    off = np.array([0, flat.size], dtype=np.int64)
    want = _oracle_records(keys, [hay])
    assert emul.emul_filter(f, flat, off, 0) == want
    assert emul.emul_dfa(f, flat, off, 0) == want


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.lists(st.lists(st.sampled_from([0, 1, 255, 256, 65535]), min_size=1, max_size=7).map(tuple), min_size=1, max_size=10, unique=True),
       st.lists(st.sampled_from([0, 1, 255, 256, 65535]), min_size=0, max_size=50).map(tuple))
def test_emulated_kernels_equal_the_oracle_sequences(keys, hay):
    mod = pkg.flavour("bytes")
    A = mod.Automaton(mod.STORE_INTS, mod.KEY_SEQUENCE)
    for i, k in enumerate(keys):
        A.add_word(k, i)
    A.make_automaton()
    f = A.flat()
    assert f["letter_bytes"] == 2
    flat = np.frombuffer(np.asarray(hay, dtype="<u2").tobytes(), dtype=np.uint8)
    off = np.array([0, flat.size], dtype=np.int64)
    want = _oracle_records(keys, [hay])
    assert emul.emul_filter(f, flat, off, 0) == want
    assert emul.emul_dfa(f, flat, off, 0) == want
