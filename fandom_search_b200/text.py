"""Host-side text helpers (native, through the C ABI): tokeniser, string ids, edit distance.

Mirrors what the reference takes from spaCy's tokeniser-only pipeline (search.py:43-63,
164-166) and from python-Levenshtein (search.py:14,190).  The tokenisation rule is
"split on ASCII whitespace", which is what the oracle's spaCy stand-in does and what the
reference's cleaned corpora look like (ao3.py:55 collapses all whitespace to single spaces);
spaCy's punctuation rules are NOT reproduced (no spaCy in the build image) -- a production
deployment that has spaCy passes `tokenizer=` to the pipeline instead.
"""
import numpy as np

from . import _native as nt


def tokenize_spans(data):
    """(starts, ends) byte offsets of the whitespace-separated tokens of `data` (bytes)."""
    lib = nt.load()
    n = len(data)
    cap = n // 2 + 1
    starts = np.empty(cap, dtype=np.int64)
    ends = np.empty(cap, dtype=np.int64)
    cnt = lib.fs_tokenize_ws(data, n, nt.ptr(starts), nt.ptr(ends), cap)
    return starts[:cnt], ends[:cnt]


def tokenize(text):
    """List of token strings of `text` (no whitespace tokens are produced)."""
    data = text.encode("utf-8")
    starts, ends = tokenize_spans(data)
    return [data[s:e].decode("utf-8") for s, e in zip(starts.tolist(), ends.tolist())]


_ID_CACHE = {}


def string_id(text):
    """spaCy-compatible 64-bit id of a string: MurmurHash64A(utf8, seed 1)."""
    h = _ID_CACHE.get(text)
    if h is None:
        b = text.encode("utf-8")
        h = int(nt.load().fs_murmurhash64a(b, len(b), 1))
        if len(_ID_CACHE) < 1 << 20:
            _ID_CACHE[text] = h
    return h


def levenshtein(a, b):
    ab = a.encode("utf-8")
    bb = b.encode("utf-8")
    return int(nt.load().fs_levenshtein_utf8(ab, len(ab), bb, len(bb)))
