#!/usr/bin/env python3
"""bench.py -- haystack GB/s of the batched Aho-Corasick search on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic haystacks.  At N=1 the
workload is BASELINE.json configs[1] (C2: 10k random 4-16 B alnum keys, 1M x 256 B haystacks,
one key planted per haystack).  Under torchrun each rank scans its own C2-sized shard (weak
scaling) and the per-rank match counts are all-gathered over NCCL inside the timed region.

  value     haystack bytes / s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e       same metric through Automaton.find_all_batch() with a pinned HOST batch: H2D, kernel,
            D2H of count + records and the reference-order sort all inside the timed region
  roofline  the filter kernel alone: (haystack bytes + 12 B/match) / mean event-timed launch,
            against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the reference's own C extension (oracle/_ref) looping iter() on a bounded sample,
            single thread (the reference holds the GIL), rank 0 only

--impl reference times the reference extension on the host cores (one process per core, each
looping iter() over its slice of a bounded sample) and prints the same JSON shape.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD_DESC = {
    "C2": "C2: 10k random [4-16]B alnum keys, 1M x 256B haystacks, 1 planted key per haystack, seed 1001",
    "C3": "C3: 100k DNA 20-mers, 10M x 150B reads, 10% planted, seed 1003",
    "C4": "C4: 10k keys (C2 set), 64 x 16MiB haystacks, planted ~1 per 256B, seed 1004",
    "C5": "C5: 100k random [4-16]B alnum keys, 1M x 256B haystacks per GPU, seed 1005",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="C2")
    ap.add_argument("--scale", type=float, default=None, help="shrink the batch (debug only; invalidates the number)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="auto", choices=["auto", "filter", "dfa"])
    ap.add_argument("--cpu-sample", type=int, default=1_000_000, help="haystacks in the cpu_baseline sample (~10 s of CPU for C2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--variant", default="planted", choices=["planted", "sparse"], help="sparse = pure random haystacks (diagnostic)")
    return ap.parse_args()


def default_scale(cfg):
    return {"C2": 1.0, "C3": 1.0, "C4": 1.0, "C5": 0.125}[cfg]


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- reference arm / cpu baseline
def _ref_worker(args):
    keys, rows = args
    import oracle
    ref = oracle.ref_module("bytes")
    R = ref.Automaton(ref.STORE_INTS)
    for i, k in enumerate(keys):
        R.add_word(k, i)
    R.make_automaton()
    hs = [r.tobytes() for r in rows]
    t0 = time.perf_counter()
    n = 0
    for h in hs:
        for _ in R.iter(h):
            n += 1
    return time.perf_counter() - t0, n


def time_reference(keys, rows, procs):
    """Loop the reference's iter() over `rows` (uint8 [n, stride]) split across `procs` processes.
    Returns (seconds = slowest worker, matches, kind)."""
    import oracle
    if not oracle.ref_available("bytes"):
        # the reference did not travel: fall back to the C restatement (kind 'port')
        O = oracle.OracleAutomaton()
        for i, k in enumerate(keys):
            O.add_word(k, i)
        O.make_automaton()
        off = np.arange(rows.shape[0] + 1, dtype=np.int64) * rows.shape[1]
        t0 = time.perf_counter()
        n = len(O.scan_batch_bytes(rows.reshape(-1), off))
        return time.perf_counter() - t0, n, "port", 1
    if procs <= 1:
        dt, n = _ref_worker((keys, rows))
        return dt, n, "reference", 1
    import multiprocessing as mp
    parts = np.array_split(rows, procs)
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(_ref_worker, [(keys, p) for p in parts])
    return max(r[0] for r in res), sum(r[1] for r in res), "reference", procs


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.
    One process per core (the reference holds the GIL: processes are the only way to use the cores),
    each looping Automaton.iter() over its slice of a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pyahocorasick_b200 import synth
    cfg = args.config
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    full = {"C2": 1_000_000, "C3": 10_000_000, "C5": 1_000_000}.get(cfg, 1_000_000)
    per_core = {"C3": 2048}.get(cfg, 4096)                 # haystacks per core per step: ~1 MB, ~0.05-0.2 s of CPU
    step_rows = min(per_core * cores, full)
    if cfg == "C4":
        w = synth.make(cfg, scale=step_rows * 4096 / (64 * 16 * 1024 * 1024))
        rows_all = w.haystacks.reshape(-1, 4096)[:step_rows]
    else:
        w = synth.make(cfg, scale=step_rows / {"C2": 1e6, "C3": 1e7, "C5": 8e6}[cfg])
        rows_all = w.haystacks[:step_rows]
    step_rows = rows_all.shape[0]
    times, matches = [], 0
    kind, used = "reference", cores
    for s in range(args.warmup + args.steps):              # every step scans the same bounded sample
        dt, m, kind, used = time_reference(w.keys, rows_all, cores)
        if s >= args.warmup:
            times.append(dt)
            matches += m
    total_t = sum(times)
    nbytes = rows_all.size * len(times)
    val = nbytes / total_t / 1e9
    line = {
        "impl": "reference", "metric": "haystack GB/s", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[cfg], "sample": f"{step_rows} haystacks x {rows_all.shape[1]} B per step (bounded sample of the same generator)"},
        "matches_per_s": matches / total_t,
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": used, "kind": kind,
                         "sample": f"{step_rows} x {rows_all.shape[1]} B per step, one process per core looping Automaton.iter()"},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- our arm
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from pyahocorasick_b200 import _native as N
    from pyahocorasick_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = args.config
    scale = args.scale if args.scale is not None else default_scale(cfg)

    # ---- workload (every rank: same keys, its own shard of haystacks) -------------------------
    w = synth.make(cfg, scale=scale, planted=(args.variant == "planted"))
    if world > 1 and rank > 0:                         # different haystack bytes per rank, same shape
        rng = np.random.Generator(np.random.PCG64(9000 + rank))
        hay = synth.random_haystacks(rng, synth.DNA if cfg == "C3" else synth.ALNUM, *w.haystacks.shape)
        synth.plant(rng, hay, w.keys, np.arange(hay.shape[0]) if cfg != "C3" else np.nonzero(rng.random(hay.shape[0]) < 0.1)[0])
        w.haystacks = hay
    A = synth.build_automaton(w.keys)
    L = N.lib()
    tb = A._ensure_table(local)
    n_hay, stride = w.haystacks.shape
    total = int(w.haystacks.size)

    pinned = torch.empty(w.haystacks.shape, dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[...] = w.haystacks
    d_hay = pinned.cuda(non_blocking=True)
    d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    cap = max(4 * n_hay, 1 << 20)
    d_out = torch.empty((cap, 3), dtype=torch.int32, device="cuda")
    gathered = torch.zeros(world, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    algo = N.ALGOS[args.algo]

    def step():
        d_cnt.zero_()
        N.check(L.acb_scan_device(tb, d_hay.data_ptr(), total, None, n_hay, stride, d_out.data_ptr(), cap, d_cnt.data_ptr(), stream, algo))
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_cnt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    n_matches = int(d_cnt.item())

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = L.acb_launch_count()
    # ---- timed region: K steps, CUDA events on the launching stream -------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev0.record()
    for i in range(args.steps):
        d_cnt.zero_()
        kev[i][0].record()
        N.check(L.acb_scan_device(tb, d_hay.data_ptr(), total, None, n_hay, stride, d_out.data_ptr(), cap, d_cnt.data_ptr(), stream, algo))
        kev[i][1].record()
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_cnt)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    launches = L.acb_launch_count() - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * total / (ms_step * 1e-3) / 1e9
    total_matches = int(gathered.sum().item()) if world > 1 else n_matches

    # ---- e2e: the public API with a pinned host batch -----------------------------------------
    e2e = None
    if not args.no_e2e:
        host = pinned.numpy()
        for _ in range(2):
            m = A.find_all_batch(host, algo=args.algo)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(e2e_steps):
            m = A.find_all_batch(host, algo=args.algo)
            if world > 1:
                c = torch.tensor([len(m)], dtype=torch.int64, device="cuda")
                dist.all_gather_into_tensor(gathered, c)
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * total / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": world * total,
               "d2h_bytes_per_step": world * (8 + 12 * len(m)), "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "includes": "H2D of the batch from pinned host memory, kernel, D2H of count+records, reference-order sort"}
        assert len(m) == n_matches, (len(m), n_matches)

    # clocks under load: the timed region is only a few ms long, so keep the same step running for another
    # ~0.5 s (untimed) while nvidia-smi samples SM clock and throttle reasons every 50 ms
    t_end = time.perf_counter() + 0.5
    while time.perf_counter() < t_end:
        for _ in range(50):
            step()
        torch.cuda.synchronize()
    clk = clocks.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    algo_bytes = total + 12 * n_matches
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(cfg)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "acb_filter_kernel (+ acb_verify_kernel over the spill list, 4 us when empty)" if args.algo != "dfa" else "acb_dfa_kernel",
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": algo_bytes}

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ns = min(args.cpu_sample, n_hay)
        rows = w.haystacks[:ns]
        dt, m, kind, used = time_reference(w.keys, rows, 1)
        cpu = {"value": rows.size / dt / 1e9, "unit": "GB/s", "cores": used, "kind": kind,
               "sample": f"first {ns} haystacks of the batch ({rows.size / 1e6:.1f} MB), Automaton.iter() loop, {dt:.1f} s",
               "matches_per_s": m / dt, "host_cores_available": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": "haystack GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[cfg] + ("" if args.variant == "planted" else " [SPARSE variant: no planted keys]") + ("" if scale == default_scale(cfg) else f" [SCALED x{scale}: not a valid bench number]"),
                       "n_haystacks_per_gpu": n_hay, "haystack_bytes": stride, "n_keys": len(w.keys),
                       "l2": f"batch {total / 1e6:.0f} MB per GPU > 126 MB L2, no flush needed" if total > 126e6 else "batch smaller than L2",
                       "algo": args.algo, "parallelism": f"batch-sharded x{world}, NCCL all-gather of match counts" if world > 1 else "single GPU"},
            "matches_per_s": total_matches / (ms_step * 1e-3), "matches_per_step": total_matches,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
