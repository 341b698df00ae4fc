"""The N>1 path on CPU: world_size-2 gloo processes shard a batch, exchange match counts and (optionally)
gather the records.  The device scan is replaced by tests/emul.py in the workers (no GPU here); on a
multi-GPU box the same code runs with NCCL through bench.py --gpus N."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, n_hay, q):
    try:
        _worker_body(rank, world, port, n_hay, q)
    except Exception as e:                       # surface the failure instead of letting the parent time out
        q.put((rank, False, repr(e), []))


def _worker_body(rank, world, port, n_hay, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import emul
    import oracle
    from pyahocorasick_b200 import automaton as am
    from pyahocorasick_b200 import distributed as D
    from pyahocorasick_b200 import synth
    am.Automaton._scan_flat = emul.install(None, "filter")          # CPU stand-in for the kernels
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.Generator(np.random.PCG64(77))
    keys = synth.draw_keys(rng, synth.ALNUM, 300, 4, 9)
    hay = synth.random_haystacks(rng, synth.ALNUM, n_hay, 64)
    synth.plant(rng, hay, keys, np.arange(n_hay))
    A = synth.build_automaton(keys)
    sm = D.scan_sharded(A, hay)
    allrec = D.gather_records(sm)
    O = oracle.OracleAutomaton()
    for i, k in enumerate(keys):
        O.add_word(k, i)
    O.make_automaton()
    want = O.scan_batch_bytes(hay.reshape(-1), np.arange(n_hay + 1, dtype=np.int64) * 64)
    lo, hi = D.shard_bounds(n_hay, world, rank)
    mine = want[(want[:, 0] >= lo) & (want[:, 0] < hi)]
    ok = (len(sm) == len(mine) and np.array_equal(np.stack([sm.hay_id, sm.end_index, sm.key_id], axis=1), mine.astype(np.int64))
          and int(sm.counts[rank]) == len(mine) and sm.total == len(want) and np.array_equal(allrec, want.astype(np.int64))
          and int(sm.offsets[rank]) == int((want[:, 0] < lo).sum()))
    q.put((rank, bool(ok), int(sm.total), sm.counts.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_hay", [1, 37, 64])
def test_two_rank_sharded_scan(n_hay):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_hay) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_hay, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]


def test_shard_bounds_cover_everything():
    from pyahocorasick_b200.distributed import shard_bounds
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_strong_scaling_shards_tile_the_global_batch():
    """bench.py --gpus N (strong scaling): every rank builds only its rows of the one global batch; the shards of any
    world size concatenate to the same batch, plants included"""
    from pyahocorasick_b200 import distributed as D
    from pyahocorasick_b200 import synth
    n = 70_000 * 2
    full = synth.make_rows("C2", 0, n)
    for world in (2, 3):
        parts = [synth.make_rows("C2", *D.shard_bounds(n, world, r)) for r in range(world)]
        assert np.array_equal(np.concatenate([p.haystacks for p in parts]), full.haystacks)
        lo = [D.shard_bounds(n, world, r)[0] for r in range(world)]
        assert np.array_equal(np.concatenate([p.planted_hay + lo[r] for r, p in enumerate(parts)]), full.planted_hay)
        assert sum(len(p.planted_hay) for p in parts) == n
