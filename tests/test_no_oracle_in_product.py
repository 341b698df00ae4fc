"""The oracle is test infrastructure: nothing under pyahocorasick_b200/ may reference it,
and the product has no CPU search fallback."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_never_touches_the_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "pyahocorasick_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".c")):
                src = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"^\s*(import|from)\s+oracle\b|oracle/|liboracle|ac_oracle|tests[./]emul", src, flags=re.M):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
