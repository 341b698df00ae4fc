"""Load the committed golden scenarios and replay them against any ahocorasick-like module.

The scenarios and their expected values were produced by the unmodified reference
(tests/golden/make_golden.py); this module only reads the JSON, so it works on the
GPU box where /root/reference does not exist.
"""
import importlib.util
import json
import os

_G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(_G, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)

dec, enc, run_ops = make_golden.dec, make_golden.enc, make_golden.run_ops


def load(which):
    with open(os.path.join(_G, f"golden_{which}.json")) as f:
        return json.load(f)["scenarios"]


def all_scenarios():
    return load("hotpath") + load("random")
