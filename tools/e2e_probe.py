"""diagnostic: where the end-to-end time of find_all_batch goes (run with ACB_TRACE=1)"""
import cProfile, pstats, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from pyahocorasick_b200 import synth
w = synth.make("C2", 1.0)
A = synth.build_automaton(w.keys)
pinned = torch.empty(w.haystacks.shape, dtype=torch.uint8, pin_memory=True)
pinned.numpy()[...] = w.haystacks
host = pinned.numpy()
for i in range(3):
    t = time.perf_counter(); m = A.find_all_batch(host); dt = time.perf_counter() - t
    print("pinned", i, f"{dt*1e3:.2f} ms", len(m), flush=True)
pr = cProfile.Profile(); pr.enable()
for i in range(5):
    m = A.find_all_batch(host)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
