"""Deterministic synthetic workloads of BASELINE.json / SURVEY.md section 8(d).

All draws come from numpy.random.Generator(PCG64(seed)); keys are distinct; haystack
batches are contiguous uint8[n, stride].  `planted` variants copy one key into each
haystack (C2/C4/C5) or into 10 % of the reads (C3) and return where, so tests can check
that every planted occurrence is reported without running a CPU oracle at full size.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

ALNUM = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789", dtype=np.uint8)
DNA = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class Workload:
    name: str
    keys: List[bytes]
    haystacks: np.ndarray                 # uint8 [n, stride]
    planted_hay: Optional[np.ndarray]     # int64 [m]   haystack of each planted occurrence
    planted_end: Optional[np.ndarray]     # int64 [m]   end_index of each planted occurrence
    planted_key: Optional[np.ndarray]     # int64 [m]   key id (= insertion index)

    @property
    def n_hay(self):
        return self.haystacks.shape[0]

    @property
    def total_bytes(self):
        return int(self.haystacks.size)


def draw_keys(rng, alphabet: np.ndarray, n: int, lo: int, hi: int) -> List[bytes]:
    """n distinct keys with length U{lo..hi}; re-drawn until the count is exact."""
    seen = set()
    out: List[bytes] = []
    while len(out) < n:
        m = n - len(out)
        lens = rng.integers(lo, hi + 1, size=m)
        sym = alphabet[rng.integers(0, len(alphabet), size=int(lens.sum()))]
        pos = 0
        for ln in lens.tolist():
            k = sym[pos:pos + ln].tobytes()
            pos += ln
            if k not in seen:
                seen.add(k)
                out.append(k)
    return out


def random_haystacks(rng, alphabet: np.ndarray, n: int, stride: int) -> np.ndarray:
    idx = rng.integers(0, len(alphabet), size=n * stride, dtype=np.uint8)
    return alphabet[idx].reshape(n, stride)


def plant(rng, hay: np.ndarray, keys: List[bytes], rows: np.ndarray):
    """Copy one uniformly chosen key at a uniform offset into each listed row (vectorised by key length)."""
    n, stride = hay.shape
    kid = rng.integers(0, len(keys), size=len(rows))
    klen = np.fromiter((len(k) for k in keys), dtype=np.int64, count=len(keys))
    ln = klen[kid]
    off = (rng.random(len(rows)) * (stride - ln + 1)).astype(np.int64)
    flat = hay.reshape(-1)
    maxlen = int(klen.max())
    keymat = np.zeros((len(keys), maxlen), dtype=np.uint8)
    for i, k in enumerate(keys):
        keymat[i, :len(k)] = np.frombuffer(k, dtype=np.uint8)
    for L in np.unique(ln).tolist():
        sel = np.nonzero(ln == L)[0]
        base = rows[sel] * stride + off[sel]
        dst = base[:, None] + np.arange(L)[None, :]
        flat[dst] = keymat[kid[sel], :L]
    return rows.astype(np.int64), (off + ln - 1).astype(np.int64), kid.astype(np.int64)


def make(name: str, scale: float = 1.0, planted: bool = True) -> Workload:
    """name in {C1, C2, C3, C4, C5}; scale shrinks the haystack batch (and C3/C5 key sets)
    for tests -- scale=1.0 is the BASELINE.json size."""
    if name == "C1":
        keys = [b"he", b"her", b"hers", b"she"]
        tile = np.frombuffer(b"_sherhershe_", dtype=np.uint8)
        hay = np.resize(tile, 1024).reshape(1, 1024).copy()
        return Workload(name, keys, hay, None, None, None)
    if name == "C2":
        rng = np.random.Generator(np.random.PCG64(1001))
        keys = draw_keys(rng, ALNUM, 10_000, 4, 16)
        n = max(1, int(round(1_000_000 * scale)))
        hay = random_haystacks(rng, ALNUM, n, 256)
        pl = plant(rng, hay, keys, np.arange(n)) if planted else (None, None, None)
        return Workload(name, keys, hay, *pl)
    if name == "C3":
        rng = np.random.Generator(np.random.PCG64(1003))
        nk = max(100, int(round(100_000 * min(1.0, scale * 10))))
        keys = draw_keys(rng, DNA, nk, 20, 20)
        n = max(1, int(round(10_000_000 * scale)))
        hay = random_haystacks(rng, DNA, n, 150)
        if planted:
            rows = np.nonzero(rng.random(n) < 0.10)[0]
            pl = plant(rng, hay, keys, rows)
        else:
            pl = (None, None, None)
        return Workload(name, keys, hay, *pl)
    if name == "C4":
        rng = np.random.Generator(np.random.PCG64(1004))
        keys = draw_keys(np.random.Generator(np.random.PCG64(1001)), ALNUM, 10_000, 4, 16)   # C2's key set
        nlong = 64
        size = max(4096, int(round(16 * 1024 * 1024 * scale)) // 256 * 256)
        hay = random_haystacks(rng, ALNUM, nlong, size)
        if planted:
            # one key per 256 B on average: plant into a [nlong*size/256, 256] view, which also puts
            # keys right up to every 256 B boundary; straddlers are added below
            view = hay.reshape(-1, 256)
            pr, pe, pk = plant(rng, view, keys, np.arange(view.shape[0]))
            per = size // 256
            ph, pend = pr // per, (pr % per) * 256 + pe
            # straddle every 16 KiB boundary (the filter kernel's work unit) and every 512 B warp step
            flat = hay.reshape(-1)
            sh, se, sk = [], [], []
            for h in range(nlong):
                for b in range(16384, size, 16384):
                    kid = int(rng.integers(0, len(keys)))
                    k = np.frombuffer(keys[kid], dtype=np.uint8)
                    cut = int(rng.integers(1, len(k)))
                    st = b - cut
                    flat[h * size + st:h * size + st + len(k)] = k
                    sh.append(h); se.append(st + len(k) - 1); sk.append(kid)
            # the straddlers may have overwritten parts of earlier plants: keep only plants whose bytes survived
            ph = np.concatenate([ph, np.asarray(sh, dtype=np.int64)])
            pend = np.concatenate([pend, np.asarray(se, dtype=np.int64)])
            pk = np.concatenate([pk, np.asarray(sk, dtype=np.int64)])
            klen = np.fromiter((len(k) for k in keys), dtype=np.int64, count=len(keys))
            ok = np.ones(len(ph), dtype=bool)
            starts = ph * size + pend - klen[pk] + 1
            # verify survival exactly (cheap: vectorised compare per key length)
            for L in np.unique(klen[pk]).tolist():
                sel = np.nonzero(klen[pk] == L)[0]
                got = flat[starts[sel][:, None] + np.arange(L)[None, :]]
                want = np.stack([np.frombuffer(keys[int(k)], dtype=np.uint8) for k in pk[sel]])
                ok[sel] = (got == want).all(axis=1)
            pl = (ph[ok], pend[ok], pk[ok])
        else:
            pl = (None, None, None)
        return Workload(name, keys, hay, *pl)
    if name == "C5":
        rng = np.random.Generator(np.random.PCG64(1005))
        nk = max(100, int(round(100_000 * min(1.0, scale * 10))))
        keys = draw_keys(rng, ALNUM, nk, 4, 16)
        n = max(1, int(round(8_000_000 * scale)))
        hay = random_haystacks(rng, ALNUM, n, 256)
        pl = plant(rng, hay, keys, np.arange(n)) if planted else (None, None, None)
        return Workload(name, keys, hay, *pl)
    raise ValueError(name)


ROW_BLOCK = 65536


def make_rows(name: str, lo: int, hi: int, planted: bool = True) -> Workload:
    """Rows [lo, hi) of the SHARDED form of a fixed-stride workload (C2 / C5), for multi-GPU strong scaling.

    The global batch is defined block by block -- ROW_BLOCK rows each, every block its own PCG64 stream seeded by
    (seed, block index) -- so a rank builds only the blocks its shard touches, and all ranks agree on the global batch
    whatever the world size is.  (The bytes differ from make(name): that one draws the whole batch from one stream.)
    Keys are make(name)'s.  planted_hay is relative to `lo`."""
    seed, alphabet, nk, klo, khi, stride = {"C2": (1001, ALNUM, 10_000, 4, 16, 256), "C5": (1005, ALNUM, 100_000, 4, 16, 256)}[name]
    keys = draw_keys(np.random.Generator(np.random.PCG64(seed)), alphabet, nk, klo, khi)
    parts, ph, pe, pk = [], [], [], []
    for b in range(lo // ROW_BLOCK, (max(hi, lo + 1) - 1) // ROW_BLOCK + 1):
        rng = np.random.Generator(np.random.PCG64([seed, b]))
        hay = random_haystacks(rng, alphabet, ROW_BLOCK, stride)
        pl = plant(rng, hay, keys, np.arange(ROW_BLOCK)) if planted else None
        r0, r1 = max(lo, b * ROW_BLOCK) - b * ROW_BLOCK, min(hi, (b + 1) * ROW_BLOCK) - b * ROW_BLOCK
        parts.append(hay[r0:r1])
        if pl is not None:
            keep = (pl[0] >= r0) & (pl[0] < r1)
            ph.append(pl[0][keep] + b * ROW_BLOCK - lo); pe.append(pl[1][keep]); pk.append(pl[2][keep])
    hay = np.ascontiguousarray(np.concatenate(parts, axis=0)) if parts else np.empty((0, stride), dtype=np.uint8)
    if planted:
        return Workload(name, keys, hay, np.concatenate(ph), np.concatenate(pe), np.concatenate(pk))
    return Workload(name, keys, hay, None, None, None)


def build_automaton(keys: List[bytes], module=None):
    """STORE_INTS automaton, value = insertion index."""
    if module is None:
        from . import flavour
        module = flavour("bytes")
    A = module.Automaton(module.STORE_INTS)
    for i, k in enumerate(keys):
        A.add_word(k, i)
    A.make_automaton()
    return A
