"""AutomatonSearchIterLong.set() (src/AutomatonSearchIterLong.c:156-212) against the reference extension:
random keys, random chunks, set() after a random number of next() calls (before a match, between matches,
after exhaustion, after extra calls past exhaustion), with and without reset.  On CPU the device is the
emulation of tests/emul.py; the gpu-marked twin runs the real ACB_ALGO_LONG kernel."""
import numpy as np
import pytest

import emul
import oracle
import pyahocorasick_b200 as pkg

needs_ref = pytest.mark.skipif(not oracle.ref_available("bytes"), reason="needs oracle/_ref")


def _fuzz(fl, trials, seed):
    ref, mod = oracle.ref_module(fl), pkg.flavour(fl)
    rng = np.random.default_rng(seed)
    al = "abc"

    def word(lo, hi):
        s = "".join(al[int(j)] for j in rng.integers(0, len(al), size=int(rng.integers(lo, hi))))
        return s.encode() if fl == "bytes" else s

    for _ in range(trials):
        keys = list({word(1, 5) for _ in range(int(rng.integers(1, 8)))})
        A, R = mod.Automaton(), ref.Automaton()
        for i, k in enumerate(keys):
            A.add_word(k, i), R.add_word(k, i)
        A.make_automaton(), R.make_automaton()
        chunks = [word(0, 12) for _ in range(4)]
        ia, ir = A.iter_long(chunks[0]), R.iter_long(chunks[0])
        got, want = [], []
        for c in chunks[1:] + [None]:
            for _ in range(int(rng.integers(0, 6))):
                for it, acc in ((ir, want), (ia, got)):
                    try:
                        acc.append(next(it))
                    except StopIteration:
                        acc.append("stop")
            if c is None:
                break
            reset = bool(rng.integers(0, 4) == 0)
            ir.set(c, reset), ia.set(c, reset)
        got += list(ia)
        want += list(ir)
        assert got == want, (fl, keys, chunks)


@needs_ref
@pytest.mark.parametrize("fl", ["bytes", "unicode"])
def test_iter_long_set_matches_the_reference_emulated(fl, monkeypatch):
    emul.install(monkeypatch, "filter")
    _fuzz(fl, 150, 11)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("fl", ["bytes", "unicode"])
def test_iter_long_set_matches_the_reference_on_gpu(fl):
    _fuzz(fl, 25, 12)


def test_iter_long_set_documented_example_and_errors(monkeypatch):
    """the straddling key is found, positions continue, and the argument rules are the reference's"""
    emul.install(monkeypatch, "filter")
    mod = pkg.flavour("bytes")
    A = mod.Automaton()
    for i, k in enumerate([b"he", b"here", b"her", b"abcd"]):
        A.add_word(k, i)
    A.make_automaton()
    it = A.iter_long(b"xxhe")
    assert list(it) == [(3, 0)]
    it.set(b"re is abcd")                     # the walk restarted after "he": nothing straddles
    assert list(it) == [(4 + 9, 3)]
    it.set(b"..ab")
    assert list(it) == []
    it.set(b"cd!")                            # "abcd" straddles the seam: 14 + 4 + 1
    assert list(it) == [(19, 3)]
    it.set(b"here", True)
    assert list(it) == [(3, 1)]
    with pytest.raises(IndexError):
        it.set()
    with pytest.raises(TypeError):
        it.set("text")
    A.add_word(b"zzz", 9)
    with pytest.raises(ValueError, match="has changed"):
        next(it)
