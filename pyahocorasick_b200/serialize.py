"""The reference's two serialisations, read and written by the drop-in.

* pickle -- ``Automaton.__reduce__`` returns ``(Automaton, (bytes_list, kind, store, key_type, count,
  longest_word, values))`` (src/Automaton_pickle.c:215-262) and the constructor accepts that 7-tuple back
  (src/Automaton.c:106-147 -> ``automaton_unpickle`` src/Automaton_pickle.c:330-456).
* ``Automaton.save(path[, serializer])`` / ``ahocorasick.load(path, deserializer)`` -- a header, one record per
  node, a footer (src/custompickle/custompickle.h:5-24, save/automaton_save.c:39-138,
  load/module_automaton_load.c:46-280).

Both carry the same node records; ``libacb200`` produces and parses them natively
(``acb_trie_export_nodes`` / ``acb_trie_import_nodes``, include/acb200.h).  Files and pickles are
interchangeable with a reference build of the SAME flavour (the letter width is baked into the records:
2 bytes in the bytes build, 4 in the unicode build) on an LP64 little-endian machine -- the reference dumps raw
structs, so that restriction is the reference's own.  Failure links are not read back: loading enters the keys
and, for a file of kind AHOCORASICK, runs ``make_automaton`` -- the links are a function of the key set.
"""
from __future__ import annotations

import ctypes
import struct

import numpy as np

from . import _native as N

EMPTY, TRIE, AHOCORASICK = 0, 1, 2
STORE_INTS, STORE_LENGTH, STORE_ANY = 10, 20, 30
KEY_STRING, KEY_SEQUENCE = 100, 200

MAGICK = b"pyahocorasick002"                     # src/custompickle/custompickle.c:5-8
HEADER = struct.Struct("<16siii4xQi4x")          # CustompickleHeader: magick, kind, store, key_type, words_count, longest_word
FOOTER = struct.Struct("<Q16s")                  # CustompickleFooter: nodes_count, magick
CHUNK_BYTES = 16 * 1024 * 1024                   # src/Automaton_pickle.c:196
ACB_NODES_PICKLE, ACB_NODES_SAVE = 0, 1
_M64 = (1 << 64) - 1


def _width(A) -> int:
    return 4 if A._L == 4 else 2                 # TRIE_LETTER_TYPE, src/common.h:51-67


def _vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# ------------------------------------------------------------------ export
def export_nodes(A):
    """(records uint8[], rec_off int64[N+1], eow_key int32[N]); `output` carries the value unless STORE_ANY"""
    vals = None
    if A._store != STORE_ANY:
        vals = np.array([(int(v) & _M64) if v is not None else 0 for v in A._values], dtype=np.uint64).view(np.int64)
    nv = 0 if vals is None else len(vals)
    need, n = ctypes.c_int64(0), ctypes.c_int64(0)
    N.check(A._lib.acb_trie_export_nodes(A._trie, _width(A), _vp(vals), nv, None, 0, ctypes.byref(need), ctypes.byref(n),
                                         None, None, 0))
    out = np.zeros(max(1, need.value), dtype=np.uint8)
    rec_off = np.zeros(n.value + 1, dtype=np.int64)
    eow = np.full(max(1, n.value), -1, dtype=np.int32)
    N.check(A._lib.acb_trie_export_nodes(A._trie, _width(A), _vp(vals), nv, _vp(out), need.value, ctypes.byref(need),
                                         ctypes.byref(n), _vp(rec_off), _vp(eow), n.value))
    return out[:need.value], rec_off, eow[:n.value]


def reduce_args(A) -> tuple:
    """the argument tuple of __reduce__ (src/Automaton_pickle.c:199-262); () for an automaton without keys"""
    if len(A) == 0:
        return ()
    rec, rec_off, eow = export_nodes(A)
    n = len(eow)
    total = int(rec_off[n])
    chunks = []
    if total <= CHUNK_BYTES:
        chunks.append(struct.pack("<q", n) + rec.tobytes())
    else:                                        # equal-sized arrays, whole records only, the last one shrunk
        room = CHUNK_BYTES - 8
        i = 0
        while i < n:
            j = int(np.searchsorted(rec_off, rec_off[i] + room, side="right")) - 1
            j = max(j, i + 1)
            body = rec[int(rec_off[i]):int(rec_off[j])].tobytes()
            pad = b"" if j == n else b"\0" * (room - len(body))
            chunks.append(struct.pack("<q", j - i) + body + pad)
            i = j
    values = [A._values[k] for k in eow.tolist() if k >= 0] if A._store == STORE_ANY else None
    return (chunks, A.kind, A._store, A._key_type, len(A), int(A._lib.acb_trie_longest_word(A._trie)), values)


# ------------------------------------------------------------------ import
def _key_object(A, raw: bytes):
    if A._key_type == KEY_SEQUENCE:
        return tuple(np.frombuffer(raw, dtype="<u4" if A._L == 4 else "<u2").tolist())
    if A._L == 4:
        return raw.decode("utf-32-le", "surrogatepass")
    return raw


def _import(A, buf: np.ndarray, n_nodes: int, mode: int, store_any: bool):
    """parse the records in buf into A's (empty) trie; -> (values int64[n_keys], blob_off int64[n_keys], consumed)"""
    lib = A._lib
    w = _width(A)
    nk, need, used = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
    scratch = lib.acb_trie_new(A._L)             # sizes first: the import itself needs an empty trie
    if not scratch:
        raise MemoryError(N.last_error())
    try:
        N.check(lib.acb_trie_import_nodes(scratch, _vp(buf), len(buf), n_nodes, w, mode, int(store_any), None, None, 0,
                                          ctypes.byref(nk), ctypes.byref(used), None, 0, None, ctypes.byref(need)))
    finally:
        lib.acb_trie_free(scratch)
    n_keys = nk.value
    vals = np.zeros(max(1, n_keys), dtype=np.int64)
    blob = np.full(max(1, n_keys), -1, dtype=np.int64)
    kbytes = np.zeros(max(1, need.value), dtype=np.uint8)
    koff = np.zeros(n_keys + 1, dtype=np.int64)
    N.check(lib.acb_trie_import_nodes(A._trie, _vp(buf), len(buf), n_nodes, w, mode, int(store_any), _vp(vals), _vp(blob),
                                      n_keys, ctypes.byref(nk), ctypes.byref(used), _vp(kbytes), need.value, _vp(koff),
                                      ctypes.byref(need)))
    raw = kbytes.tobytes()
    ko = koff.tolist()
    A._key_objs = [_key_object(A, raw[ko[k]:ko[k + 1]]) for k in range(n_keys)]
    A._key_ids = {A._hashable(k): i for i, k in enumerate(A._key_objs)}
    return vals[:n_keys], blob[:n_keys], used.value


def _int_values(vals: np.ndarray) -> list:
    """`output.integer` as the reference hands it out: Py_BuildValue("i") truncation"""
    return vals.astype(np.int32).tolist()        # int64 -> int32 keeps the low 32 bits, sign included


def _finish(A, kind: int, count: int, longest: int, cache_path=None):
    if kind == AHOCORASICK:
        A._make_automaton_cached(cache_path)         # the flat-table cache: a second load of the same keys skips the build
    # counters as written in the file; they equal what entering the keys produced unless the writer had
    # removed words (longest_word never shrinks, src/Automaton.c:285-286)
    A._version = 0


def from_reduce_args(A, args) -> None:
    """the 7-tuple branch of the constructor, src/Automaton.c:106-147"""
    try:
        bytes_list, kind, store, key_type, count, longest, values = args
        kind, store, key_type, count, longest = (int(x) for x in (kind, store, key_type, count, longest))
    except (TypeError, ValueError):
        raise ValueError("Unable to load from pickle.") from None
    A._check_store(store)
    if kind not in (EMPTY, TRIE, AHOCORASICK):
        raise ValueError("kind value must be one of ahocorasick.EMPTY, TRIE or AHOCORASICK")
    A._check_key_type(key_type)
    if type(bytes_list) is not list:
        raise TypeError("Expected list")
    A._configure(store, key_type)
    if kind == EMPTY:
        return
    w = _width(A)
    bodies, total = [], 0
    for k, chunk in enumerate(bytes_list):       # automaton_unpickle__validate_bytes_list, :270-303
        if type(chunk) is not bytes:
            raise ValueError(f"Item #{k} on the bytes list is not a bytes object")
        if len(chunk) < 8:
            raise ValueError(f"Data truncated [parsing header of node #0]: chunk #{k}")
        cnt = struct.unpack_from("<q", chunk)[0]
        if cnt <= 0:
            raise ValueError(f"Nodes count for item #{k} on the bytes list is not positive ({cnt})")
        body = np.frombuffer(chunk, dtype=np.uint8)[8:]
        span = ctypes.c_int64(0)
        N.check(A._lib.acb_node_records_span(_vp(body), len(body), cnt, w, ctypes.byref(span)))
        bodies.append(body[:span.value])
        total += cnt
    buf = np.ascontiguousarray(np.concatenate(bodies)) if len(bodies) != 1 else np.ascontiguousarray(bodies[0])
    vals, _, _ = _import(A, buf, total, ACB_NODES_PICKLE, store == STORE_ANY)
    if store == STORE_ANY:
        if values is None or len(values) < len(vals):
            raise IndexError("list index out of range")          # PyList_GetItem in :425
        A._values = list(values[:len(vals)])
    else:
        A._values = _int_values(vals)
    _finish(A, kind, count, longest)


# ------------------------------------------------------------------ save / load
def _parse_save_load_args(store: int, args):
    """src/custompickle/pyhelpers.c:4-59"""
    if store == STORE_ANY:
        if len(args) != 2:
            raise ValueError("expected exactly two arguments")
    elif len(args) != 1:
        raise ValueError("expected exactly one argument")
    path = args[0]
    if not isinstance(path, str):
        raise TypeError("the first argument must be a string")
    callback = None
    if store == STORE_ANY:
        callback = args[1]
        if not callable(callback):
            raise TypeError("the second argument must be a callable object")
    return path, callback


def save(A, *args) -> None:
    path, serializer = _parse_save_load_args(A._store, args)
    kind = A.kind
    header = HEADER.pack(MAGICK, kind, A._store, A._key_type, len(A), int(A._lib.acb_trie_longest_word(A._trie)))
    with open(path, "wb") as fh:
        fh.write(header)
        n = 0
        if kind != EMPTY:
            rec, rec_off, eow = export_nodes(A)
            n = len(eow)
            raw = rec.tobytes()
            off = rec_off.tolist()
            keys = eow.tolist()
            any_store = A._store == STORE_ANY
            for i in range(n):                   # automaton_save_node, :85-138: address, record, children, value
                a, b = off[i], off[i + 1]
                fh.write(struct.pack("<Q", i + 1))
                if any_store and keys[i] >= 0:
                    blob = serializer(A._values[keys[i]])
                    if type(blob) is not bytes:
                        raise TypeError("serializer must return bytes object")
                    fh.write(struct.pack("<Q", len(blob)) + raw[a + 8:b] + blob)
                else:
                    fh.write(raw[a:b])
        fh.write(FOOTER.pack(n, MAGICK))


def load(cls, *args):
    """module-level load(path, deserializer): the new automaton starts as STORE_ANY, so both arguments are
    always required (module_automaton_load.c:16-27)"""
    path, deserializer = _parse_save_load_args(STORE_ANY, args)
    with open(path, "rb") as fh:
        data = fh.read()
    if len(data) < HEADER.size + FOOTER.size:
        raise OSError("file too short for a header and a footer")
    magick, kind, store, key_type, count, longest = HEADER.unpack_from(data, 0)
    n_nodes, magick2 = FOOTER.unpack_from(data, len(data) - FOOTER.size)
    if magick != MAGICK or store not in (STORE_LENGTH, STORE_INTS, STORE_ANY) or kind not in (EMPTY, TRIE, AHOCORASICK) \
            or key_type not in (KEY_STRING, KEY_SEQUENCE):
        raise ValueError("invalid header")
    if magick2 != MAGICK:
        raise ValueError("invalid footer")
    A = cls(store, key_type)
    if kind == EMPTY:
        return A
    body = np.frombuffer(data, dtype=np.uint8)[HEADER.size:len(data) - FOOTER.size]
    vals, blob, _ = _import(A, np.ascontiguousarray(body), n_nodes, ACB_NODES_SAVE, store == STORE_ANY)
    if store == STORE_ANY:
        sizes = (vals.view(np.uint64)).tolist()
        offs = blob.tolist()
        base = HEADER.size
        A._values = [deserializer(data[base + o:base + o + s]) for o, s in zip(offs, sizes)]
    else:
        A._values = _int_values(vals)
    _finish(A, kind, count, longest, cache_path=path + ".acb200")
    return A
