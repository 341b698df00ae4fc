#!/usr/bin/env python3
"""Golden vectors for the reference's two serialisations, produced by the
UNMODIFIED reference extension (oracle/_ref, compiled from /root/reference by oracle/Makefile):
for a handful of automata, the argument tuple of `__reduce__` and the bytes of a `save` file, plus what the
reference answers after reading them back.  Run here (the reference does not exist on the GPU box):

    python tests/golden/make_serialized.py        # rewrites tests/golden/golden_serialized.json
"""
import base64
import json
import os
import pickle
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

WORDS = ["he", "her", "hers", "she", "his", "x\xe9y", "abc", "abd", "b", "\xff"]
HAY = "_sherhershe_x\xe9yabdhis\xffb"
STORES = {"STORE_ANY": 30, "STORE_INTS": 20, "STORE_LENGTH": 10}


def conv(fl, s):
    return s.encode("latin-1") if fl == "bytes" else s


def main():
    out = []
    for fl in ("bytes", "unicode"):
        ref = oracle.ref_module(fl)
        for sname, store in STORES.items():
            for built in (False, True):
                for removed in (False, True):
                    A = ref.Automaton(store)
                    for i, w in enumerate(WORDS):
                        k = conv(fl, w)
                        if store == 30:
                            A.add_word(k, [i, w])
                        elif store == 20:
                            A.add_word(k, i * 1000003 - 5)
                        else:
                            A.add_word(k)
                    if removed:
                        A.remove_word(conv(fl, "hers"))
                        A.remove_word(conv(fl, "abc"))
                    if built:
                        A.make_automaton()
                    args = A.__reduce__()[1]
                    with tempfile.TemporaryDirectory() as d:
                        p = os.path.join(d, "f")
                        if store == 30:
                            A.save(p, pickle.dumps)
                        else:
                            A.save(p)
                        blob = open(p, "rb").read()
                    vals = sorted(json.dumps(v) for v in A.values())
                    sc = dict(flavour=fl, store=sname, built=built, removed=removed, kind=A.kind, count=len(A),
                              reduce_chunks=[base64.b64encode(c).decode() for c in args[0]],
                              reduce_tail=[args[1], args[2], args[3], args[4], args[5]],
                              reduce_values=args[6], save_file=base64.b64encode(blob).decode(),
                              values_sorted=vals)
                    if fl == "unicode":            # keys() of the bytes build is not usable; its values are
                        sc["items"] = sorted([k, json.dumps(v)] for k, v in A.items())
                    if built:
                        sc["iter"] = [[e, json.dumps(v)] for e, v in A.iter(conv(fl, HAY))]
                    out.append(sc)
    path = os.path.join(HERE, "golden_serialized.json")
    with open(path, "w") as fh:
        json.dump(dict(words=WORDS, hay=HAY, scenarios=out), fh, indent=0)
    print(len(out), "scenarios ->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
