"""Test-only, pure-Python emulation of the DEVICE algorithm (csrc/acb_device.cu) on the
flattened tables produced by the host core.  It exists so that the CPU test-suite can
exercise the host logic (filter construction, flattening, the Python API layer) without
a GPU; it is never imported by the product.

emul_filter() restates acb_stream_kernel: the gram-bitmap probe at every `stride`-th byte (single or pair placement),
then the anchor table (UNIQUE anchors compare the key, MULTI anchors walk the trie).  emul_dfa() restates
acb_dfa_kernel (goto / fail / CSR outputs).  Both return records sorted the way
acb_scan_host sorts them.
"""
from __future__ import annotations

import numpy as np

M32 = 0xFFFFFFFF
S1 = (0x9E3779B1, 0x85EBCA77, 0xC2B2AE3D, 0x27D4EB2F)
S2 = (0x165667B1, 0xD3A2646D, 0xFD7046C5, 0xB55A4F09)


def multipliers(g, stage):
    base = S1 if stage == 1 else S2
    nw = (g + 3) // 4
    out = []
    for k in range(4):
        m = base[k]
        if k >= nw:
            m = 0
        elif k == nw - 1:
            unused = 4 * nw - g
            m = (m << (8 * unused)) & M32
        out.append(m)
    return out


def hash_bytes(buf, q, g, mul):
    """zero-filled past the end of buf, exactly like the kernel's guarded loads"""
    h = 0
    nw = (g + 3) // 4
    n = len(buf)
    for k in range(nw):
        w = 0
        for b in range(4):
            i = 4 * k + b
            if i < g and q + i < n:
                w |= int(buf[q + i]) << (8 * b)
        h = (h + w * mul[k]) & M32
    return h


def hash_bytes_wide(buf, q, g, mul):
    """acb_hash_bytes_wide: the same sum with 64-bit products (low half == hash_bytes)"""
    h = 0
    nw = (g + 3) // 4
    n = len(buf)
    for k in range(nw):
        w = 0
        for b in range(4):
            i = 4 * k + b
            if i < g and q + i < n:
                w |= int(buf[q + i]) << (8 * b)
        h = (h + w * mul[k]) & 0xFFFFFFFFFFFFFFFF
    return h


FILTER_WIDE = 1
FILTER_PAIR = 2


def _bit(bm, idx):
    return (int(bm[idx >> 5]) >> (idx & 31)) & 1


def _bounds(q, offsets, stride, n_hay):
    if offsets is None:
        h = q // stride
        return h, h * stride, h * stride + stride
    h = int(np.searchsorted(offsets, q, side="right")) - 1
    return h, int(offsets[h]), int(offsets[h + 1])


def _sorted(recs, key_len):
    recs.sort(key=lambda r: (r[0], r[1], -int(key_len[r[2]])))
    return recs


def _walk(f, buf, start, h, hs, he, recs):
    L = f["letter_bytes"]
    cls, gto, key_of = f["byte_class"], f["goto_cm"], f["key_of"]
    st = 0
    for i in range(start, he):
        nx = int(gto[cls[buf[i]], st])
        if nx < 0:
            break
        st = nx
        k = int(key_of[st])
        if k >= 0:
            recs.append((h, (i - hs + 1) // L - 1, k))


def _entry_bytes(e, n):
    raw = b"".join(int(w).to_bytes(4, "little") for w in e[3:8])
    return raw[:n]


PAIR_M = 0x9E3779B1


def pair_place(G, role, log2_bits):
    """acb_pair_place (csrc/acb_hash.h): (word, bit) of gram G (little-endian u32) in level 1, in one role"""
    mulp = (PAIR_M << 8) & M32
    common = G if role else (G >> 8)
    hc = (common * mulp) & M32
    amount = ((G * mulp) >> 32) if role else G
    return hc >> (37 - log2_bits), 1 << (31 - (amount & 31))


def pair_place2(tag, log2_bits2):
    """acb_pair_place2: (word, bits) of an anchor tag in level 2 (word counted from the start of level 2)"""
    return tag >> (37 - log2_bits2), (1 << ((tag >> (32 - log2_bits2)) & 31)) | (1 << ((tag >> (27 - log2_bits2)) & 31))


def _u32_at(buf, q):
    """little-endian word at q, zero filled past the end of buf (the kernel may see other bytes there: they can
    only add survivors that the anchor compare rejects)"""
    n = len(buf)
    return sum(int(buf[q + i]) << (8 * i) for i in range(4) if q + i < n)


def _passes_bitmap(f, buf, q):
    """the shared-memory bitmap test of acb_stream_kernel at probe position q"""
    g, l1, flags = f["gram_bytes"], f["log2_bits1"], f["filter_flags"]
    n_words = 1 << (l1 - 5)
    if flags & FILTER_PAIR:
        assert g == 4 and f["stride"] == 1 and f["letter_bytes"] == 1
        role = q & 1                                   # x even: role 0 of pair (x, x+1); x odd: role 1 of (x-1, x)
        G = _u32_at(buf, q)
        word1, bit1 = pair_place(G, role, l1)
        if not int(f["bitmap1"][word1]) & bit1:
            return False
        tag = ((G * multipliers(4, 2)[0]) & M32) | 1
        word2, bits2 = pair_place2(tag, f["log2_bits2"])
        return (int(f["bitmap1"][n_words + word2]) & bits2) == bits2
    mul1 = multipliers(g, 1)
    hw = hash_bytes_wide(buf, q, g, mul1)
    h1 = hw & M32
    assert bool(flags & FILTER_WIDE) == (g % 4 == 0)
    bit_a = ((hw >> 32) & 31) if flags & FILTER_WIDE else ((h1 >> (32 - l1)) & 31)
    w1 = int(f["bitmap1"][(h1 * n_words) >> 32])
    return bool((w1 >> bit_a) & (w1 >> (h1 & 31)) & 1)


def emul_filter(f, buf, offsets=None, stride_bytes=0):
    """gram bitmap -> anchor table (UNIQUE: direct key compare, MULTI: trie walk)"""
    L, g, s = f["letter_bytes"], f["gram_bytes"], f["stride"]
    mul2 = multipliers(g, 2)
    lA = f["log2_anchor_slots"]
    anchors = f["anchors"]
    amask = (1 << lA) - 1
    total = len(buf)
    raw = bytes(bytearray(buf))
    n_hay = (len(offsets) - 1) if offsets is not None else total // stride_bytes
    recs = []
    if f["n_keys"] == 0:
        return recs
    for q in range(0, total, s):
        if not _passes_bitmap(f, buf, q):
            continue
        if q + g > total:
            continue
        tag = hash_bytes(buf, q, g, mul2) | 1
        l3 = f["log2_bits3"]
        if l3 and not _bit(f["bitmap3"], ((tag * 0x9E3779B1) & M32) >> (32 - l3)):
            continue
        slot = tag >> (32 - lA)
        bounds = None
        while True:
            e = anchors[slot]
            if int(e[0]) == 0:
                break
            if int(e[0]) == tag:
                if bounds is None:
                    bounds = _bounds(q, offsets, stride_bytes, n_hay)
                h, hs, he = bounds
                kid = int(np.int32(np.uint32(e[1])))
                j, ln = int(e[2]) & 0xFF, (int(e[2]) >> 8) & 0xFF
                start = q - j
                if start >= hs:
                    if kid >= 0:
                        if start + ln <= he and raw[start:start + ln] == _entry_bytes(e, ln):
                            recs.append((h, (start + ln - hs) // L - 1, kid))
                    elif raw[q:q + ln] == _entry_bytes(e, ln):
                        _walk(f, buf, start, h, hs, he, recs)
                if (int(e[2]) >> 16) & 1:
                    break                      # last entry carrying this tag
            slot = (slot + 1) & amask
    return _sorted(recs, f["key_len"])


def emul_dfa(f, buf, offsets=None, stride_bytes=0, span=64):
    L = f["letter_bytes"]
    total = len(buf)
    n_hay = (len(offsets) - 1) if offsets is not None else total // stride_bytes
    cls, gto, fail = f["byte_class"], f["goto_cm"], f["fail"]
    out_ptr, out_idx, key_len = f["out_ptr"], f["out_idx"], f["key_len"]
    maxb = f["max_key_bytes"]
    recs = []
    for a in range(0, total, span):
        b = min(a + span, total)
        h, hs, he = _bounds(a, offsets, stride_bytes, n_hay)
        i = max(hs, a - maxb)
        st = 0
        while i < b:
            while i >= he:
                h += 1
                hs = he
                he = hs + stride_bytes if offsets is None else int(offsets[h + 1])
                st = 0
            c = cls[buf[i]]
            nx = int(gto[c, st])
            while nx < 0 and st != 0:
                st = int(fail[st])
                nx = int(gto[c, st])
            st = 0 if nx < 0 else nx
            if i >= a and st != 0 and (i + 1 - hs) % L == 0:
                for o in range(int(out_ptr[st]), int(out_ptr[st + 1])):
                    k = int(out_idx[o])
                    recs.append((h, (i - hs + 1) // L - 1, k))
            i += 1
    return _sorted(recs, key_len)


def emul_long(f, buf, offsets=None, stride_bytes=0, init_state=0, want_state=False):
    """acb_long_kernel: the reference's iter_long state machine (src/AutomatonSearchIterLong.c:89-153) on the
    flattened tables, one haystack at a time, letter by letter.  init_state: the state haystack 0 starts in
    (acb_table_set_long_state); want_state: also return the state it ended in (acb_table_get_long_state)."""
    L = f["letter_bytes"]
    total = len(buf)
    n_hay = (len(offsets) - 1) if offsets is not None else total // stride_bytes
    cls, gto, lfail, key_of = f["byte_class"], f["goto_cm"], f["letter_fail"], f["key_of"]
    recs = []

    for h in range(n_hay):
        hs = int(offsets[h]) if offsets is not None else h * stride_bytes
        he = int(offsets[h + 1]) if offsets is not None else hs + stride_bytes
        n = (he - hs) // L

        def step(st, i):
            for b in range(L):
                st = int(gto[cls[buf[hs + i * L + b]], st])
                if st < 0:
                    return -1
            return st

        state, index, last_node, last_index = (init_state if h == 0 else 0), -1, -1, -1
        while True:
            if last_node >= 0:
                recs.append((h, last_index, int(key_of[last_node])))
                state, index, last_node, last_index = 0, last_index, -1, -1
            index += 1
            emit = False
            while index < n:
                nx = step(state, index)
                if nx >= 0:
                    if key_of[nx] >= 0:
                        last_node, last_index = nx, index
                    else:
                        fl = int(lfail[nx])
                        if fl > 0 and key_of[fl] >= 0:
                            last_node, last_index, emit = fl, index, True
                            break
                    state = nx
                    index += 1
                else:
                    if last_node >= 0:
                        emit = True
                        break
                    while True:
                        state = int(lfail[state])
                        if state < 0:
                            state = 0
                            index += 1
                            break
                        if step(state, index) >= 0:
                            break
            if not emit and last_node < 0:
                break
        if h == 0:
            final_state = state
    if want_state:
        return _sorted(recs, f["key_len"]), (final_state if n_hay else init_state)
    return _sorted(recs, f["key_len"])


def install(monkeypatch_or_none, algo="filter"):
    """Route Automaton._scan_flat through the emulation (CPU tests of the host logic only)."""
    from pyahocorasick_b200 import _native as N
    from pyahocorasick_b200 import automaton as am

    default_algo = algo

    def fake_scan_flat(self, flat, offsets, n_hay, stride_bytes, algo="auto", sort=True, device=None, narrow=False, long_state=None):
        f = self.flat(narrow=narrow)
        if f is None:
            return np.empty(0, dtype=N.MATCH_DTYPE)
        if algo == "auto":
            algo = default_algo
        fn = {"dfa": emul_dfa, "long": emul_long}.get(algo, emul_filter)
        if long_state is not None:
            recs, self._long_state_out = emul_long(f, np.asarray(flat, dtype=np.uint8), offsets, stride_bytes, init_state=long_state, want_state=True)
        else:
            recs = fn(f, np.asarray(flat, dtype=np.uint8), offsets, stride_bytes)
        out = np.empty(len(recs), dtype=N.MATCH_DTYPE)
        for i, r in enumerate(recs):
            out[i] = r
        return out

    if monkeypatch_or_none is not None:
        monkeypatch_or_none.setattr(am.Automaton, "_scan_flat", fake_scan_flat)
    return fake_scan_flat
