#!/usr/bin/env python3
"""print a one-line summary of a bench.py JSON line read from stdin (diagnostics)"""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
txt = sys.stdin.read().strip().splitlines()
try:
    d = json.loads(txt[-1])
    r = d["roofline"]
    print(f"{tag} value={d['value']:.0f}GB/s frac={r['frac']:.3f} kernel_ms={r['kernel_ms']:.4f} matches={d['matches_per_step']} e2e={(d.get('e2e') or {}).get('value')}")
except Exception as e:
    print(tag, "NO JSON:", e, txt[-3:] if txt else "")
