"""Host logic of the drop-in (argument parsing, state machine, iterator/set() bookkeeping,
value stores, flatten + filter construction) replayed over the committed golden vectors.

No GPU here: Automaton._scan_flat is routed through tests/emul.py, a pure-Python restatement
of the two device kernels operating on the tables the host core produced.  The same scenarios
run against the real kernels in tests/test_gpu_parity.py (-m gpu).
"""
import pytest

import emul
import pyahocorasick_b200 as ac
from golden_driver import all_scenarios, run_ops

SC = all_scenarios()


@pytest.mark.parametrize("algo", ["filter", "dfa"])
@pytest.mark.parametrize("sc", SC, ids=[s["name"] for s in SC])
def test_golden_via_emulated_device(sc, algo, monkeypatch):
    emul.install(monkeypatch, algo)
    bad = run_ops(ac.flavour(sc["flavour"]), sc, record=False)
    assert not bad, bad[:3]


def test_unicode_narrow_and_wide_paths_agree_with_oracle(monkeypatch):
    """unicode flavour: latin-1 haystacks take the 1-byte-per-letter automaton, others the 4-byte one;
    streams may switch between the two from chunk to chunk."""
    import oracle
    emul.install(monkeypatch, "filter")
    U = ac.flavour("unicode")
    words = ["he", "hers", "é", "éa", "中文", "a中", "he\U0001F629"]
    A = U.Automaton()
    O = oracle.OracleAutomaton()
    for i, w in enumerate(words):
        A.add_word(w, i)
        O.add_word(w, i)
    A.make_automaton()
    O.make_automaton()
    texts = ["hers é éa he", "xx中文a中 hers", "he\U0001F629é", "", "plain ascii hers"]
    for t in texts:
        assert list(A.iter(t)) == O.find_all(t)
        assert list(A.iter_long(t)) == O.iter_long(t)
    m = A.find_all_batch(texts)                       # mixed batch -> all wide
    assert m.per_haystack(len(texts)) == [O.find_all(t) for t in texts]
    m = A.find_all_batch([texts[0], texts[4]])        # all latin-1 -> narrow
    assert m.per_haystack(2) == [O.find_all(texts[0]), O.find_all(texts[4])]
    # a stream that alternates narrow and wide chunks, with keys straddling the boundaries
    chunks = ["..h", "ers a", "中", "文 é", "a h", "e\U0001F629"]
    it = A.iter("")
    got = []
    for c in chunks:
        it.set(c)
        got += list(it)
    assert got == O.find_all("".join(chunks))
    # only non-latin keys: the narrow automaton does not exist
    B = U.Automaton()
    B.add_word("中文", 1)
    B.make_automaton()
    assert list(B.iter("latin only")) == [] and list(B.iter("x中文")) == [(2, 1)]


def test_one_automaton_is_safe_to_share_between_threads(monkeypatch):
    """the native calls drop the GIL, so searches and key-set changes on one Automaton are serialised by a
    per-object lock (the reference holds the GIL for a whole search).  Hammer one automaton from several threads."""
    import threading
    import emul
    import pyahocorasick_b200 as pkg
    emul.install(monkeypatch)
    A = pkg.flavour("bytes").Automaton()
    for w in (b"he", b"her", b"hers", b"she"):
        A.add_word(w, w)
    A.make_automaton()
    want = list(A.iter(b"_sherhershe_"))
    errors = []

    def search():
        try:
            for _ in range(30):
                assert list(A.iter(b"_sherhershe_")) == want
        except Exception as e:                                   # pragma: no cover
            errors.append(e)

    def churn():
        try:
            for i in range(30):
                B = A                                            # same object: add and remove a key, rebuild
                B.add_word(b"zz%d" % i, b"zz")
                B.make_automaton()
                B.remove_word(b"zz%d" % i)
                B.make_automaton()
        except Exception as e:                                   # pragma: no cover
            errors.append(e)

    ts = [threading.Thread(target=search) for _ in range(3)] + [threading.Thread(target=churn)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not [e for e in errors if not isinstance(e, (ValueError, AttributeError))], errors   # stale iterators may raise, nothing may crash
