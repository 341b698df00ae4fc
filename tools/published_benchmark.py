#!/usr/bin/env python3
"""The reference's only published benchmark shape, through the drop-in (and, optionally, through the reference itself).

The shape (reference etc/benchmarks/benchmark.py:40-41, 52-58, 89-94, 116; numbers in etc/benchmarks/results/*.txt):
N random words of 3..32 characters over [a-zA-Z0-9], each stored with itself as value; four timed stages -- add
every word, make_automaton, 2 N lookups with get(), and ONE search: iter() over one random string of 1 000 000
characters, counting matches.  Published for N = 1 000 000 on a Xeon E3-1505M v6: add 1.040 s, build 6.015 s, lookup
1.307 s, search 0.279 s.

    python tools/published_benchmark.py [--words 1000000] [--flavour bytes|unicode] [--reference]

Prints one JSON line.  `filter` says how the gram filter came out for this key set (which part saturates)."""
import argparse
import json
import os
import random
import string
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(mod, words, text, conv):
    t = {}
    A = mod.Automaton()
    t0 = time.perf_counter()
    for w in words:
        A.add_word(w, w)
    t["add_words_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    A.make_automaton()
    t["build_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(2):
        for w in words:
            A.get(w)
    t["lookup_2n_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    n = sum(1 for _ in A.iter(text))                 # first search: includes the upload of the tables to the device
    t["first_search_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    n2 = sum(1 for _ in A.iter(text))
    t["search_s"] = time.perf_counter() - t0
    assert n == n2
    t["matches"] = n
    return A, t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--words", type=int, default=1_000_000)
    ap.add_argument("--flavour", default="bytes", choices=["bytes", "unicode"])
    ap.add_argument("--reference", action="store_true", help="also time the compiled reference (oracle/_ref) on the same input")
    a = ap.parse_args()
    rng = random.Random(0)
    chars = string.ascii_letters + string.digits
    seen = set()
    while len(seen) < a.words:
        seen.add("".join(rng.choice(chars) for _ in range(rng.randint(3, 32))))
    conv = (lambda s: s.encode()) if a.flavour == "bytes" else (lambda s: s)
    words = [conv(w) for w in seen]
    text = conv("".join(rng.choice(chars) for _ in range(1_000_000)))
    import pyahocorasick_b200 as pkg
    A, ours = run(pkg.flavour(a.flavour), words, text, conv)
    f = A.flat()
    bits = __import__("numpy").unpackbits(f["bitmap1"].view("uint8"))
    out = {"shape": f"{a.words} words of 3..32 chars over [a-zA-Z0-9], one 1 000 000-char haystack, {a.flavour} flavour",
           "ours": ours,
           "filter": {"gram_bytes": f["gram_bytes"], "stride": f["stride"], "placement": {0: "single", 1: "single (wide)", 2: "pair"}[f["filter_flags"]],
                      "shared_memory_bits_log2": f["log2_bits1"], "shared_memory_fill": float(bits.mean()),
                      "tag_bitmap_bits_log2": f["log2_bits3"], "states": f["n_states"], "classes": f["n_classes"],
                      "goto_table_bytes": int(f["n_states"]) * int(f["n_classes"]) * 4},
           "published_xeon_e3_1505m_v6": {"add_words_s": 1.040, "build_s": 6.015, "lookup_2n_s": 1.307, "search_s": 0.279, "words": 1_000_000}}
    if a.reference:
        import oracle
        if oracle.ref_available(a.flavour):
            _, ref = run(oracle.ref_module(a.flavour), words, text, conv)
            assert ref["matches"] == ours["matches"], (ref["matches"], ours["matches"])
            out["reference_here"] = ref
    print(json.dumps(out))


if __name__ == "__main__":
    main()
