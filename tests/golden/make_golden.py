#!/usr/bin/env python3
"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs oracle/_ref, i.e. /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

Every expected value in tests/golden/*.json is produced by the reference C
extension (bytes and unicode flavours) executing the scenario -- nothing is
typed in by hand.  Scenarios restate the reference's own hot-path tests
(tests/test_unit.py:529-857, tests/test_basic.py:18-50, tests/test_issue_10.py,
_53.py, _56.py, _8.py) plus seeded random differential cases the reference
suite lacks (last paragraph).

Scenario format (JSON):
  {"name", "flavour": "bytes"|"unicode", "store": 10|20|30, "key_type": 100|200,
   "words": [[key, value|null], ...], "make": true|false,
   "ops": [{"op": ..., ...args..., "expect": ... | "raises": "ExcName"}]}
Keys / haystacks are encoded as {"b": hex} (bytes), {"s": [codepoints]} (str)
or {"t": [ints]} (tuple, KEY_SEQUENCE).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402

STORE_INTS, STORE_LENGTH, STORE_ANY = 10, 20, 30
KEY_STRING, KEY_SEQUENCE = 100, 200


# ---------------------------------------------------------------------------
# encoding helpers (shared with tests/golden_driver.py)
# ---------------------------------------------------------------------------
def enc(x):
    if isinstance(x, (bytes, bytearray)):
        return {"b": bytes(x).hex()}
    if isinstance(x, str):
        return {"s": [ord(c) for c in x]}
    if isinstance(x, tuple):
        return {"t": list(x)}
    if x is None:
        return None
    raise TypeError(type(x))


def dec(d):
    if d is None:
        return None
    if "b" in d:
        return bytes.fromhex(d["b"])
    if "s" in d:
        return "".join(chr(c) for c in d["s"])
    return tuple(d["t"])


def conv(flavour, s):
    """pytestingutils.conv of the reference: identity / utf-8 encode."""
    return s if flavour == "unicode" else s.encode("utf-8")


# ---------------------------------------------------------------------------
# run a scenario against an ahocorasick-like module, recording results
# ---------------------------------------------------------------------------
def run_ops(mod, sc, record: bool):
    """Execute sc against module `mod`.  record=True fills expect/raises; record=False
    returns a list of (op, got, want) mismatches."""
    args = [sc["store"]]
    if sc.get("key_type", KEY_STRING) != KEY_STRING:
        args.append(sc["key_type"])
    A = mod.Automaton(*args)
    for key, val in sc["words"]:
        k = dec(key)
        if sc["store"] == STORE_ANY:
            A.add_word(k, val)
        elif sc["store"] == STORE_INTS and val is not None:
            A.add_word(k, val)
        else:
            A.add_word(k)
    if sc.get("make", True):
        A.make_automaton()
    bad = []
    for op in sc["ops"]:
        got, exc = None, None
        try:
            got = _exec(A, op)
        except Exception as e:  # noqa: BLE001 - recording the type is the point
            exc = type(e).__name__
        if record:
            if exc is not None:
                op["raises"] = exc
            else:
                op["expect"] = got
        else:
            if "raises" in op:
                if exc != op["raises"]:
                    bad.append((op, exc or got, op["raises"]))
            elif exc is not None or got != op["expect"]:
                bad.append((op, exc or got, op["expect"]))
    return bad


def _pairs(it):
    return [[int(i), v] for i, v in it]


def _exec(A, op):
    kind = op["op"]
    if kind == "iter":
        kw = dict(op.get("kw", {}))
        return _pairs(A.iter(dec(op["hay"]), *op.get("args", []), **kw))
    if kind == "find_all":
        res = []
        r = A.find_all(dec(op["hay"]), lambda i, v: res.append([int(i), v]), *op.get("args", []))
        assert r is None
        return res
    if kind == "iter_long":
        return _pairs(A.iter_long(dec(op["hay"]), *op.get("args", [])))
    if kind == "find_all_notbuilt":
        return A.find_all(dec(op["hay"]), lambda i, v: None)
    if kind == "iter_set":
        it = A.iter(dec(op["init"]), *op.get("args", []))
        out = []
        if op.get("drain_first"):
            out.append(_pairs(it))
        for chunk, reset in op["chunks"]:
            if reset is None:
                it.set(dec(chunk))
            else:
                it.set(dec(chunk), reset)
            out.append(_pairs(it))
        return out
    if kind == "iter_set_partial":
        # consume `take` items from the first haystack, then set() and drain
        it = A.iter(dec(op["init"]))
        out = [[list(map(_jsonable, next(it))) for _ in range(op["take"])]]
        it.set(dec(op["chunk"]), op.get("reset", False))
        out.append(_pairs(it))
        return out
    if kind == "kind":
        return int(A.kind)
    if kind == "len":
        return len(A)
    raise ValueError(kind)


def _jsonable(x):
    return int(x) if isinstance(x, (int, np.integer)) else x


# ---------------------------------------------------------------------------
# scenarios
# ---------------------------------------------------------------------------
def hotpath_scenarios():
    S = []
    for fl in ("bytes", "unicode"):
        c = lambda s, fl=fl: enc(conv(fl, s))  # noqa: E731
        base_words = [[c(w), i] for i, w in enumerate("he her hers she".split())]
        text = "_sherhershe_"
        # tests/test_unit.py:593-696 + :699-807
        S.append(dict(name=f"unit_search_{fl}", flavour=fl, store=STORE_ANY, words=base_words, ops=[
            dict(op="iter", hay=c(text)),
            dict(op="find_all", hay=c(text)),
            dict(op="iter", hay=c(text[4:9])),
            dict(op="iter", hay=c(text), args=[4, 9]),
            dict(op="find_all", hay=c(text[4:9])),
            dict(op="find_all", hay=c(text), args=[4, 9]),
            dict(op="find_all", hay=c(text), args=[0, len(text) + 5]),
            dict(op="find_all", hay=c(text), args=[-len(text) - 1, 3]),
            dict(op="find_all", hay=c(text), args=[0]),
            dict(op="find_all", hay=c(text), args=[-3, 4]),
            dict(op="find_all", hay=c(text), args=[0, -1]),
            dict(op="find_all", hay=c(text), args=[-9, -2]),
            dict(op="find_all", hay=c(text), args=[len(text), len(text)]),
            dict(op="find_all", hay=c(text), args=[3, 3]),
            dict(op="find_all", hay=c(text), args=[5, 2]),
            dict(op="iter", hay=c(text), args=[5, 2]),
            dict(op="iter", hay=c(text), args=[0, -1]),
            dict(op="iter", hay=c(text), args=[-1, 6]),
            dict(op="iter", hay=c(text), kw=dict(start=2)),
            dict(op="iter", hay=c(text), kw=dict(end=7)),
            dict(op="iter", hay=c("")),
            dict(op="find_all", hay=c("")),
            dict(op="iter_set", init=c(""), chunks=[[c(p), None] for p in "_sh erhe rshe _".split()]),
            dict(op="iter_set", init=c(""), chunks=[[c(p), True] for p in ["he", "she"]]),
            dict(op="iter_set", init=c("_she"), drain_first=True, chunks=[[c("rs"), False], [c("he"), False], [c(""), False], [c("rs_she"), True]]),
            dict(op="iter_set", init=c("xxhers"), args=[2], drain_first=True, chunks=[[c("he"), False]]),
            dict(op="iter_set_partial", init=c("_sherhershe_"), take=3, chunk=c("rshe")),
            dict(op="iter_set_partial", init=c("_sherhershe_"), take=1, chunk=c("rshe")),
            dict(op="iter_set_partial", init=c("_sherhershe_"), take=2, chunk=c("she"), reset=True),
            dict(op="kind"), dict(op="len"),
        ]))
        # not built: tests/test_unit.py:596-605, :702-711
        S.append(dict(name=f"unit_notbuilt_{fl}", flavour=fl, store=STORE_ANY, make=False, words=base_words, ops=[
            dict(op="find_all_notbuilt", hay=c(text)), dict(op="iter", hay=c(text)), dict(op="kind")]))
        S.append(dict(name=f"unit_empty_{fl}", flavour=fl, store=STORE_ANY, make=True, words=[], ops=[
            dict(op="find_all_notbuilt", hay=c(text)), dict(op="iter", hay=c(text)), dict(op="kind")]))
        # ignore_white_space: tests/test_unit.py:810-857
        ws = "_sh e rher she_"
        S.append(dict(name=f"unit_ignore_ws_{fl}", flavour=fl, store=STORE_ANY, words=base_words, ops=[
            dict(op="iter", hay=c(ws), kw=dict(ignore_white_space=True)),
            dict(op="iter", hay=c(ws), kw=dict(ignore_white_space=True, start=12)),
            dict(op="iter", hay=c(ws), kw=dict(ignore_white_space=False)),
            dict(op="iter", hay=c(ws), kw=dict(ignore_white_space=2)),
            dict(op="iter", hay=c("s\th\ne\r\x0b\x0cr s  he"), kw=dict(ignore_white_space=True)),
            dict(op="iter", hay=c("  she  "), kw=dict(ignore_white_space=True, start=1, end=6)),
            dict(op="iter", hay=c("   "), kw=dict(ignore_white_space=True)),
        ]))
        # tests/test_basic.py:18-50 (+ unicode twin :101-133): duplicate key keeps the last value
        bw = "he e hers his she hi him man he".split()
        S.append(dict(name=f"basic_{fl}", flavour=fl, store=STORE_ANY, words=[[c(w), i] for i, w in enumerate(bw)], ops=[
            dict(op="iter", hay=c("he rshershidamanza "), kw=dict(start=2, end=8)),
            dict(op="find_all", hay=c("he rshershidamanza "), args=[2, 11]),
            dict(op="iter", hay=c("he rshershidamanza ")),
            dict(op="len"),
        ]))
        # tests/test_issue_10.py
        S.append(dict(name=f"issue10_{fl}", flavour=fl, store=STORE_ANY, words=[[c("S"), 1]], ops=[
            dict(op="iter", hay=c("SSS"), args=[0, 3]), dict(op="iter", hay=c("SSS"), args=[0, 2]),
            dict(op="iter", hay=c("SSS"), args=[1]), dict(op="iter", hay=c("SSS"), args=[3, 3])]))
        # tests/test_issue_53.py
        S.append(dict(name=f"issue53_{fl}", flavour=fl, store=STORE_ANY, words=[[c("wounded"), 7]], ops=[
            dict(op="iter", hay=c("Winning \U0001F629 so gutted, can't do anything for 4 weeks... Myth. #wounded")),
            dict(op="iter", hay=c("Winning so gutted, can't do anything for 4 weeks... Myth. #wounded"))]))
        # tests/test_issue_56.py
        S.append(dict(name=f"issue56_{fl}", flavour=fl, store=STORE_ANY,
                      words=[[c(w), i] for i, w in enumerate(("poke", "go", "pokegois", "egoist"))], ops=[
            dict(op="iter", hay=c("pokego pokego  pokegoist")), dict(op="find_all", hay=c("pokego pokego  pokegoist"))]))
        # tests/test_issue_8.py:50-83 (multi-byte UTF-8 / non-latin letters)
        pl = ["a", "wy", "ważyć", "aż", "waży", "ż", "ć"]
        S.append(dict(name=f"issue8_{fl}", flavour=fl, store=STORE_ANY, words=[[c(w), i] for i, w in enumerate(pl)], ops=[
            dict(op="iter", hay=c("wyważyć")), dict(op="find_all", hay=c("wyważyć")),
            dict(op="iter", hay=c("zażółć gęślą jaźń wyważyć"))]))
        # STORE_INTS default values count+1 and STORE_LENGTH: tests/test_unit.py:987-1074
        S.append(dict(name=f"store_ints_{fl}", flavour=fl, store=STORE_INTS,
                      words=[[c(w), None] for w in "he her hers she he".split()], ops=[
            dict(op="iter", hay=c(text)), dict(op="find_all", hay=c(text)), dict(op="len")]))
        S.append(dict(name=f"store_ints_explicit_{fl}", flavour=fl, store=STORE_INTS,
                      words=[[c("he"), 2 ** 31 - 1], [c("she"), -5], [c("her"), 2 ** 40 + 5], [c("hers"), 0]], ops=[
            dict(op="iter", hay=c(text)), dict(op="find_all", hay=c(text))]))
        S.append(dict(name=f"store_length_{fl}", flavour=fl, store=STORE_LENGTH,
                      words=[[c(w), None] for w in "he her hers she".split()], ops=[
            dict(op="iter", hay=c(text)), dict(op="find_all", hay=c(text))]))
        # pathological overlaps: every suffix is a key
        S.append(dict(name=f"overlap_a_{fl}", flavour=fl, store=STORE_ANY,
                      words=[[c("a" * k), k] for k in (1, 2, 3, 4, 7)], ops=[
            dict(op="iter", hay=c("a" * 9)), dict(op="iter", hay=c("aabaaaabaaaaaaa")),
            dict(op="iter", hay=c("a" * 9), args=[2, 7])]))
        S.append(dict(name=f"overlap_mixed_{fl}", flavour=fl, store=STORE_ANY,
                      words=[[c(w), i] for i, w in enumerate(["abcd", "bcd", "cd", "d", "abcde", "bc", "cdx", "xab"])], ops=[
            dict(op="iter", hay=c("xabcdexabcdx")), dict(op="find_all", hay=c("xabcdexabcdx"), args=[1, 11])]))
        # regression "SAMSUNG-GT-C3303": tests/test_unit.py:1102-1115 shape (keys sharing long prefixes)
        S.append(dict(name=f"shared_prefix_{fl}", flavour=fl, store=STORE_ANY,
                      words=[[c(w), i] for i, w in enumerate(["GT-C3303", "SAMSUNG-GT-C3303K/"])], ops=[
            dict(op="iter", hay=c("SAMSUNG-GT-C3303i/1.0 NetFront/3.5 Profile/MIDP-2.0")),
            dict(op="iter", hay=c("SAMSUNG-GT-C3303K/1.0 SAMSUNG-GT-C330 GT-C3303"))]))
    # iter_long: tests/test_unit.py:1491-1520, tests/test_issue_133.py, docs/automaton_iter_long.rst:41-44
    for fl in ("bytes", "unicode"):
        c = lambda s, fl=fl: enc(conv(fl, s))  # noqa: E731
        S.append(dict(name=f"iter_long_{fl}", flavour=fl, store=STORE_ANY,
                      words=[[c(w), i] for i, w in enumerate("he here her".split())], ops=[
            dict(op="iter_long", hay=c("he here her")), dict(op="iter_long", hay=c("he here her"), args=[2]),
            dict(op="iter_long", hay=c("he here her"), args=[3, 9]), dict(op="iter_long", hay=c("")),
            dict(op="iter_long", hay=c("hehehere herehe h")), dict(op="iter", hay=c("he here her"))]))
        S.append(dict(name=f"iter_long_133a_{fl}", flavour=fl, store=STORE_ANY,
                      words=[[c("b"), 0], [c("abc"), 1]], ops=[dict(op="iter_long", hay=c("abb")), dict(op="iter_long", hay=c("ababcabb"))]))
        S.append(dict(name=f"iter_long_133b_{fl}", flavour=fl, store=STORE_ANY,
                      words=[[c(w), i] for i, w in enumerate(["b", "c", "abd"])], ops=[dict(op="iter_long", hay=c("abc")), dict(op="iter_long", hay=c("abdabcab"))]))
        S.append(dict(name=f"iter_long_133c_{fl}", flavour=fl, store=STORE_ANY,
                      words=[[c(w), i] for i, w in enumerate(["知识产权", "国家知识产权局"])], ops=[dict(op="iter_long", hay=c("国家知识产权")), dict(op="iter_long", hay=c("国家知识产权局知识产权"))]))
        S.append(dict(name=f"iter_long_overlap_{fl}", flavour=fl, store=STORE_ANY,
                      words=[[c(w), i] for i, w in enumerate(["a", "ab", "abc", "bcd", "cde", "abcdef", "f", "ef"])], ops=[
            dict(op="iter_long", hay=c("abcdefabcdeabcdabcabaf")), dict(op="iter_long", hay=c("xxabcdexefab")),
            dict(op="iter_long", hay=c("aaaaabababcabcdabcdeabcdef"))]))
    S.append(dict(name="iter_long_sequence", flavour="unicode", store=STORE_ANY, key_type=KEY_SEQUENCE,
                  words=[[enc((1, 2)), 0], [enc((1, 2, 3)), 1], [enc((1, 2, 3, 4)), 2]], ops=[
        dict(op="iter_long", hay=enc((0, 1, 2, 3, 4, 0, 0, 1, 2, 0, 1, 3, 1, 2, 3, 0)))]))
    S.append(dict(name="iter_long_notbuilt", flavour="bytes", store=STORE_ANY, make=False,
                  words=[[enc(b"he"), 0]], ops=[dict(op="iter_long", hay=enc(b"he"))]))
    # bytes >= 0x80 (src/utils.c:199-202 sign extension) -- bytes flavour only
    S.append(dict(name="high_bytes", flavour="bytes", store=STORE_ANY,
                  words=[[enc(b"\xff\x80"), 0], [enc(b"\x80"), 1], [enc(b"\x00\x00"), 2], [enc(b"\x7f\x80\xff"), 3]], ops=[
        dict(op="iter", hay=enc(b"\x00\xff\x80\x00\x00\x00\x7f\x80\xff\x80")),
        dict(op="find_all", hay=enc(bytes(range(256)) * 2))]))
    # astral + mixed-width letters -- unicode flavour only
    S.append(dict(name="astral", flavour="unicode", store=STORE_ANY,
                  words=[[enc("\U0001F629"), 0], [enc("a\U0001F629b"), 1], [enc("ż中"), 2], [enc("中文"), 3], [enc("ab"), 4]], ops=[
        dict(op="iter", hay=enc("xa\U0001F629b中文ż中文ab")),
        dict(op="iter", hay=enc("ab latin only ab")),
        dict(op="iter", hay=enc("中文 bmp only 中文ab"))]))
    # KEY_SEQUENCE: tests/test_unit.py:1224-1259 shape
    for fl in ("bytes", "unicode"):
        hi = 65535 if fl == "bytes" else 2 ** 32 - 1
        S.append(dict(name=f"sequence_{fl}", flavour=fl, store=STORE_ANY, key_type=KEY_SEQUENCE,
                      words=[[enc((1, 2, 3)), 0], [enc((2, 3)), 1], [enc((3,)), 2], [enc((hi, 0, hi)), 3], [enc((1, 2, 3, 4, 5)), 4], [enc((300, 2)), 5]], ops=[
            dict(op="iter", hay=enc((0, 1, 2, 3, 4, 5, hi, 0, hi, 300, 2, 3))),
            dict(op="find_all", hay=enc((1, 2, 3, 1, 2, 3, 4, 5)), args=[1, 7]),
            dict(op="iter", hay=enc(())),
        ]))
    # wrong argument types: tests/test_unit.py:802-807, :1261-1267
    for fl in ("bytes", "unicode"):
        wrong = enc("text") if fl == "bytes" else enc(b"text")
        S.append(dict(name=f"wrong_type_{fl}", flavour=fl, store=STORE_ANY,
                      words=[[enc(conv(fl, "he")), 0]], ops=[dict(op="iter", hay=wrong), dict(op="iter", hay=None),
                                                            dict(op="find_all", hay=wrong)]))
    return S


def random_scenarios():
    """Seeded differential cases: random keys + haystacks, expected output = reference iter()."""
    S = []
    specs = [
        # name, flavour, alphabet, n_words, len range, n_hay, hay_len
        ("rand_ab", "bytes", b"ab", 12, (1, 6), 6, 60),
        ("rand_abc_long", "bytes", b"abc", 40, (2, 9), 4, 200),
        ("rand_dna", "bytes", b"ACGT", 60, (4, 8), 6, 150),
        ("rand_alnum", "bytes", b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789", 80, (1, 5), 6, 256),
        ("rand_allbytes", "bytes", bytes(range(256)), 200, (1, 3), 4, 512),
        ("rand_hibytes", "bytes", bytes([0, 1, 127, 128, 129, 254, 255]), 30, (1, 5), 6, 100),
        ("rand_min4", "bytes", b"abcdef", 50, (4, 16), 6, 256),
        ("rand_uni_latin", "unicode", "abcé", 20, (1, 5), 5, 80),
        ("rand_uni_bmp", "unicode", "ab中文ż", 25, (1, 5), 5, 80),
        ("rand_uni_astral", "unicode", "a\U0001F629\U00010000￿b", 25, (1, 4), 5, 80),
    ]
    for si, (name, fl, alpha, nw, (lo, hi), nh, hl) in enumerate(specs):
        rng = np.random.Generator(np.random.PCG64(4242 + si))
        is_b = fl == "bytes"

        def draw(n):
            ix = rng.integers(0, len(alpha), size=n)
            return bytes(alpha[i] for i in ix) if is_b else "".join(alpha[i] for i in ix)

        words = []
        for i in range(nw):
            words.append([enc(draw(int(rng.integers(lo, hi + 1)))), i])
        ops = []
        for h in range(nh):
            hay = draw(hl)
            # plant a few keys so matches are not vanishingly rare
            for _ in range(3):
                w = dec(words[int(rng.integers(0, nw))][0])
                p = int(rng.integers(0, max(1, hl - len(w))))
                hay = hay[:p] + w + hay[p + len(w):]
            ops.append(dict(op="iter", hay=enc(hay)))
            ops.append(dict(op="iter_long", hay=enc(hay)))
            if h % 3 == 0:
                a = int(rng.integers(0, hl // 2))
                b = int(rng.integers(hl // 2, hl + 1))
                ops.append(dict(op="iter", hay=enc(hay), args=[a, b]))
                ops.append(dict(op="find_all", hay=enc(hay), args=[a, b]))
            if h % 3 == 1:
                cuts = sorted(int(x) for x in rng.integers(0, hl, size=3))
                parts = [hay[:cuts[0]], hay[cuts[0]:cuts[1]], hay[cuts[1]:cuts[2]], hay[cuts[2]:]]
                ops.append(dict(op="iter_set", init=enc(parts[0]), drain_first=True, chunks=[[enc(p), False] for p in parts[1:]]))
        S.append(dict(name=name, flavour=fl, store=STORE_ANY, words=words, ops=ops))
    return S


def main():
    mods = {fl: oracle.ref_module(fl) for fl in ("bytes", "unicode")}
    for fname, scs in (("golden_hotpath.json", hotpath_scenarios()), ("golden_random.json", random_scenarios())):
        for sc in scs:
            sc.setdefault("key_type", KEY_STRING)
            sc.setdefault("make", True)
            run_ops(mods[sc["flavour"]], sc, record=True)
        with open(os.path.join(HERE, fname), "w") as f:
            json.dump(dict(generator="tests/golden/make_golden.py",
                           reference="WojciechMula/pyahocorasick v2.2.0 (oracle/_ref, unmodified)",
                           scenarios=scs), f, indent=1, sort_keys=True)
        print(fname, len(scs), "scenarios", sum(len(s["ops"]) for s in scs), "ops")


if __name__ == "__main__":
    main()
