"""White-box checks of the host core: flattened goto/fail/output tables and the gram filter."""
import numpy as np
import pytest

import emul
import oracle
import pyahocorasick_b200 as ac
from pyahocorasick_b200 import synth

B = ac.flavour("bytes")


def _brute_fail(keys):
    """fail(s) by definition: longest proper suffix of the state's string that is a trie node."""
    pref = {b""}
    for k in keys:
        for i in range(1, len(k) + 1):
            pref.add(k[:i])
    out = {}
    for p in pref:
        if not p:
            continue
        f = b""
        for i in range(1, len(p)):
            if p[i:] in pref:
                f = p[i:]
                break
        out[p] = f
    return out


@pytest.mark.parametrize("seed", range(5))
def test_tables_match_definitions(seed):
    rng = np.random.Generator(np.random.PCG64(100 + seed))
    keys = synth.draw_keys(rng, np.frombuffer(b"abc", dtype=np.uint8), 40, 1, 7)
    A = synth.build_automaton(keys)
    f = A.flat()
    S = f["n_states"]
    # reconstruct every state's string by walking goto from the root
    inv = {int(c): b for b, c in enumerate(f["byte_class"].tolist()) if c}
    strings = {0: b""}
    for s in range(S):                       # BFS numbering: parents come first
        for c in range(1, f["n_classes"]):
            t = int(f["goto_cm"][c, s])
            if t >= 0:
                assert t > s and t not in strings
                strings[t] = strings[s] + bytes([inv[c]])
    assert len(strings) == S
    lens = [len(strings[s]) for s in range(S)]
    assert lens == sorted(lens)              # depth-ordered ids
    bf = _brute_fail(keys)
    ids = {v: k for k, v in strings.items()}
    assert f["fail"][0] == -1
    for s in range(1, S):
        assert strings[int(f["fail"][s])] == bf[strings[s]]
    # key_of / key_len / CSR outputs: every key that is a suffix of the state's string, longest first
    for s in range(S):
        want = [i for i, k in sorted(enumerate(keys), key=lambda t: -len(t[1])) if strings[s].endswith(k)]
        got = f["out_idx"][f["out_ptr"][s]:f["out_ptr"][s + 1]].tolist()
        assert got == want
        assert f["key_of"][s] == (keys.index(strings[s]) if strings[s] in keys else -1)
    assert f["key_len"].tolist() == [len(k) for k in keys]
    assert f["min_key_bytes"] == min(map(len, keys)) and f["max_key_bytes"] == max(map(len, keys))
    assert A.get_stats()["nodes_count"] == S


CASES = [
    (b"ab", 30, 1, 6), (b"ab", 30, 3, 9), (b"ACGT", 400, 12, 20), (b"ACGT", 300, 20, 20),
    (synth.ALNUM.tobytes(), 500, 4, 16), (synth.ALNUM.tobytes(), 500, 5, 9), (synth.ALNUM.tobytes(), 300, 9, 30),
    (bytes(range(256)), 300, 2, 4), (synth.ALNUM.tobytes(), 200, 17, 40),
]


@pytest.mark.parametrize("alpha,nk,lo,hi", CASES)
def test_filter_never_loses_a_match(alpha, nk, lo, hi):
    """emulated filter kernel == emulated DFA kernel == the oracle, on fixed-stride and ragged batches."""
    rng = np.random.Generator(np.random.PCG64(nk * 31 + lo))
    al = np.frombuffer(alpha, dtype=np.uint8)
    keys = synth.draw_keys(rng, al, nk, lo, hi)
    A = synth.build_automaton(keys)
    f = A.flat()
    assert f["gram_bytes"] + f["stride"] - 1 <= f["min_key_bytes"]
    O = oracle.OracleAutomaton()
    for i, k in enumerate(keys):
        O.add_word(k, i)
    O.make_automaton()
    nh, hl = 24, 96
    hay = synth.random_haystacks(rng, al, nh, hl)
    synth.plant(rng, hay, keys, np.arange(nh))
    flat = hay.reshape(-1)
    off = np.arange(nh + 1, dtype=np.int64) * hl
    want = [tuple(r) for r in O.scan_batch_bytes(flat, off).tolist()]
    assert want
    assert emul.emul_filter(f, flat, None, hl) == want
    assert emul.emul_dfa(f, flat, None, hl) == want
    cuts = np.sort(rng.integers(0, flat.size + 1, size=17))
    roff = np.concatenate([[0, 0], cuts, [flat.size]]).astype(np.int64)
    want2 = [tuple(r) for r in O.scan_batch_bytes(flat, roff).tolist()]
    assert emul.emul_filter(f, flat, roff, 0) == want2
    assert emul.emul_dfa(f, flat, roff, 0) == want2


def test_filter_choice_for_baseline_configs():
    w = synth.make("C2", scale=0.001)
    f = synth.build_automaton(w.keys).flat()
    assert (f["gram_bytes"], f["stride"], f["log2_bits1"]) == (4, 1, 20)
    assert f["filter_flags"] == emul.FILTER_PAIR          # 10 k grams of 4 bytes at stride 1: one word per two positions
    assert f["log2_bits2"] == 17 and f["bitmap1"].size == (1 << 15) + (1 << 12)
    bits = np.unpackbits(f["bitmap1"].view(np.uint8))
    n1 = 1 << 20
    assert 0.015 < bits[:n1].mean() < 0.02      # level 1: one bit per gram and role in 2^20 bits
    assert 0.12 < bits[n1:].mean() < 0.16       # level 2: two bits per gram (blocked Bloom, k = 2, keyed by the anchor tag) in 2^17 bits
    a = f["anchors"]
    used = a[a[:, 0] != 0]
    assert 9000 < len(used) <= 10000 and len(used) * 4 <= len(a)         # one anchor per distinct 4-byte prefix
    assert (used[:, 1].astype(np.int32) >= 0).mean() > 0.95               # almost all UNIQUE


def test_filter_choice_for_the_dense_key_sets():
    """BASELINE configs 5 and 3 keep the SINGLE placement (acb_stream_kernel): 100 k four-byte grams would fill a
    pair filter's level 1 to 19 % per role; the C4 key set is C2's (PAIR, acb_pair_kernel)"""
    rng = np.random.Generator(np.random.PCG64(1005))
    keys = synth.draw_keys(rng, synth.ALNUM, 100_000, 4, 16)                 # config 5's key set
    s5 = synth.build_automaton(keys).filter_shape()
    assert (s5["gram_bytes"], s5["stride"], s5["log2_bits1"]) == (4, 1, 20)
    assert s5["filter_flags"] == emul.FILTER_WIDE and s5["log2_bits2"] == 0 and s5["log2_bits3"] >= 20      # tag bitmap in L2
    rng = np.random.Generator(np.random.PCG64(3))
    dna = synth.draw_keys(rng, synth.DNA, 100_000, 20, 20)                   # config 3's shape
    s3 = synth.build_automaton(dna).filter_shape()
    assert s3["stride"] == 8 and s3["gram_bytes"] == 13 and not (s3["filter_flags"] & emul.FILTER_PAIR)
    s4 = synth.build_automaton(synth.make("C4", scale=0.001).keys).filter_shape()
    assert s4["filter_flags"] == emul.FILTER_PAIR and s4["log2_bits2"] == 17


def test_state_machine_and_removal():
    A = B.Automaton()
    assert A.kind == ac.EMPTY and A.make_automaton() is False
    assert A.add_word(b"", 1) is False and A.kind == ac.EMPTY
    assert A.add_word(b"he", 1) and A.add_word(b"hers", 2) and not A.add_word(b"he", 3)
    assert A.kind == ac.TRIE and len(A) == 2 and A.get(b"he") == 3
    assert A.make_automaton() is None and A.kind == ac.AHOCORASICK and A.make_automaton() is False
    assert A.add_word(b"she", 4) and A.kind == ac.TRIE              # demoted (src/trie.c:60)
    A.make_automaton()
    assert A.remove_word(b"hers") and not A.remove_word(b"hers") and A.kind == ac.TRIE
    assert not A.exists(b"hers") and A.match(b"he") and not A.match(b"her") and A.longest_prefix(b"hexx") == 2
    A.make_automaton()
    assert A.get_stats()["nodes_count"] == 6                        # root, h, he, s, sh, she
    assert A.pop(b"she") == 4 and sorted(A.keys()) == [b"he"]
    with pytest.raises(KeyError):
        A.pop(b"she")
    A.clear()
    assert A.kind == ac.EMPTY and len(A) == 0
    with pytest.raises(ValueError):
        B.Automaton(-42)
    with pytest.raises(ValueError):
        B.Automaton(ac.STORE_ANY, -42)


def test_pickle_round_trip_keeps_keys_values_and_kind():
    import pickle
    for mod, conv in ((B, lambda s: s.encode()), (ac.flavour("unicode"), lambda s: s)):
        A = mod.Automaton()
        for i, w in enumerate(["he", "her", "hers", "she"]):
            A.add_word(conv(w), (i, w))
        A.make_automaton()
        C = pickle.loads(pickle.dumps(A))
        assert type(C) is type(A) and C.kind == ac.AHOCORASICK and len(C) == 4
        assert sorted(C.items(), key=repr) == sorted(A.items(), key=repr)
        assert C.flat()["fail"].tolist() == A.flat()["fail"].tolist()
        T = mod.Automaton(ac.STORE_LENGTH)
        T.add_word(conv("abc"))
        D = pickle.loads(pickle.dumps(T))
        assert D.kind == ac.TRIE and D.get(conv("abc")) == 3 and D.store == ac.STORE_LENGTH


def test_tag_bitmap_is_built_for_dense_key_sets_and_never_loses_a_match(monkeypatch):
    """key sets the shared-memory filter cannot hold (100k keys of C5; forced here for a small set too) get the tag
    bitmap in global memory; the emulated filter path with it still equals the oracle"""
    w5 = synth.make("C5", scale=0.1)
    f5 = synth.build_automaton(w5.keys).flat()
    assert f5["log2_bits3"] >= 16 and f5["bitmap3"].size == 1 << (f5["log2_bits3"] - 5)
    monkeypatch.setenv("ACB_FORCE_TAGMAP", "1")
    rng = np.random.Generator(np.random.PCG64(5))
    keys = synth.draw_keys(rng, synth.ALNUM, 300, 3, 9)
    A = synth.build_automaton(keys)
    f = A.flat()
    assert f["log2_bits3"] >= 16
    hay = synth.random_haystacks(rng, synth.ALNUM, 40, 200)
    synth.plant(rng, hay, keys, np.arange(40))
    O = oracle.OracleAutomaton()
    for i, k in enumerate(keys):
        O.add_word(k, i)
    O.make_automaton()
    off = np.arange(41, dtype=np.int64) * 200
    want = [tuple(r) for r in O.scan_batch_bytes(hay.reshape(-1), off).tolist()]
    assert emul.emul_filter(f, hay.reshape(-1), None, 200) == want
