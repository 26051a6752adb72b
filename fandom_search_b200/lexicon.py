"""Lexicon: token text -> embedding-row id, plus the reference's OOV pseudo-vector rule.

Host-side mirror of what spaCy's vocabulary gives the reference on the hot path:
`Token.has_vector` / `Token.vector` (keyed on the verbatim ORTH text, search.py:74-75) and
the out-of-vocabulary rule of mk_vectors (search.py:76-83):

    vectors[i] = 0
    vectors[i][hash(w) % cols] = 1.0; vectors[i][hash(w * 2) % cols] = 1.0; vectors[i][hash(w * 3) % cols] = 1.0

(`w * 2` is string repetition).  The reference uses Python's per-process randomised `hash`;
this module does the same by default, so a drop-in run behaves like the reference run in the
same interpreter.  `py_hash_seed0` reproduces `hash(str)` under PYTHONHASHSEED=0 (SipHash-1-3
with a zero key over the string's canonical PEP-393 buffer) so that results are repeatable
without controlling the environment -- tests and golden fixtures use it.
"""
import sys

import numpy as np

_MASK = 0xFFFFFFFFFFFFFFFF


def _rotl(x, b):
    return ((x << b) | (x >> (64 - b))) & _MASK


def siphash13(data, k0=0, k1=0):
    """SipHash-1-3 of bytes (the str/bytes hash of CPython >= 3.11)."""
    v0 = k0 ^ 0x736f6d6570736575
    v1 = k1 ^ 0x646f72616e646f6d
    v2 = k0 ^ 0x6c7967656e657261
    v3 = k1 ^ 0x7465646279746573

    def rnd(v0, v1, v2, v3):
        v0 = (v0 + v1) & _MASK
        v1 = _rotl(v1, 13)
        v1 ^= v0
        v0 = _rotl(v0, 32)
        v2 = (v2 + v3) & _MASK
        v3 = _rotl(v3, 16)
        v3 ^= v2
        v0 = (v0 + v3) & _MASK
        v3 = _rotl(v3, 21)
        v3 ^= v0
        v2 = (v2 + v1) & _MASK
        v1 = _rotl(v1, 17)
        v1 ^= v2
        v2 = _rotl(v2, 32)
        return v0, v1, v2, v3

    n = len(data)
    end = n - (n % 8)
    for i in range(0, end, 8):
        m = int.from_bytes(data[i:i + 8], "little")
        v3 ^= m
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
        v0 ^= m
    b = (n & 0xFF) << 56
    b |= int.from_bytes(data[end:], "little") if n % 8 else 0
    v3 ^= b
    v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    v0 ^= b
    v2 ^= 0xFF
    for _ in range(3):
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    return (v0 ^ v1 ^ v2 ^ v3) & _MASK


def _pep393_bytes(s):
    """The buffer CPython hashes for a str: latin-1 / UCS-2 / UCS-4 by widest code point."""
    if not s:
        return b""
    m = max(map(ord, s))
    if m < 256:
        return s.encode("latin-1")
    if m < 65536:
        return s.encode("utf-16-le" if sys.byteorder == "little" else "utf-16-be", "surrogatepass")
    return s.encode("utf-32-le" if sys.byteorder == "little" else "utf-32-be", "surrogatepass")


def py_hash_seed0(s):
    """hash(s) of CPython 3.11+ run with PYTHONHASHSEED=0 (verified in tests against a
    subprocess)."""
    if not s:
        return 0
    h = siphash13(_pep393_bytes(s), 0, 0)
    if h >= 1 << 63:
        h -= 1 << 64
    if h == -1:
        h = -2
    return h


def oov_indices(word, cols, hash_fn=hash):
    """The three hot positions of search.py:81-83 (may coincide)."""
    return (hash_fn(word) % cols, hash_fn(word * 2) % cols, hash_fn(word * 3) % cols)


class Lexicon:
    """keys -> rows of a float32 [R, d] table, plus a growing registry of OOV strings."""

    def __init__(self, keys, rows, table, hash_fn=None):
        self.table = np.ascontiguousarray(table, dtype=np.float32)
        self.dim = int(self.table.shape[1])
        self.n_rows = int(self.table.shape[0])
        self.key_to_row = {str(k): int(r) for k, r in zip(keys, rows)}
        self.hash_fn = hash if hash_fn is None else hash_fn
        self._oov_id = {}      # text -> id >= n_rows
        self._oov_hot = []     # per OOV: (i0, i1, i2)

    @classmethod
    def from_npz(cls, path, hash_fn=None):
        with np.load(path, allow_pickle=False) as z:
            return cls(z["keys"], z["rows"], z["table"], hash_fn=hash_fn)

    @classmethod
    def from_spacy(cls, nlp, hash_fn=None):  # pragma: no cover - needs a real spaCy install
        """Export a loaded spaCy pipeline's vectors table (production path; spaCy and
        en_core_web_md are not available in the build image)."""
        vectors = nlp.vocab.vectors
        keys, rows = [], []
        for key, row in vectors.key2row.items():
            keys.append(nlp.vocab.strings[key])
            rows.append(row)
        return cls(keys, rows, np.asarray(vectors.data, dtype=np.float32), hash_fn=hash_fn)

    # -- lookups ------------------------------------------------------------
    @property
    def n_oov(self):
        return len(self._oov_hot)

    def has_vector(self, text):
        return text in self.key_to_row

    def row_id(self, text):
        """Embedding-row id of a token text; OOV strings get a stable id >= n_rows."""
        r = self.key_to_row.get(text)
        if r is not None:
            return r
        r = self._oov_id.get(text)
        if r is None:
            r = self.n_rows + len(self._oov_hot)
            self._oov_id[text] = r
            self._oov_hot.append(oov_indices(text, self.dim, self.hash_fn))
        return r

    def row_ids(self, words):
        get = self.key_to_row.get
        out = np.array([get(w, -1) for w in words], dtype=np.int32).reshape(-1)
        for i in np.nonzero(out < 0)[0]:
            out[i] = self.row_id(words[i])
        return out

    def batch_oov(self, strings, n_fixed):
        """Row ids and 3-hot rows for the unique out-of-vocabulary strings of ONE batch, without
        registering them: a string the registry already knows below `n_fixed` (an OOV word of the
        indexed script) keeps that id -- identical windows must have identical row ids -- every
        other string gets a batch-local id n_fixed, n_fixed + 1, ... and its row in the returned
        float32 [n_new, d] matrix.  The registry therefore holds script-side OOV words only, however
        many fanworks (and misspellings) a run sees."""
        ids = np.empty(len(strings), dtype=np.int32)
        hot = []
        known = self._oov_id
        for k, w in enumerate(strings):
            r = known.get(w)
            if r is not None and r < n_fixed:
                ids[k] = r
            else:
                ids[k] = n_fixed + len(hot)
                hot.append(oov_indices(w, self.dim, self.hash_fn))
        extra = None
        if hot:
            extra = np.zeros((len(hot), self.dim), dtype=np.float32)
            h = np.array(hot, dtype=np.int64).reshape(len(hot), 3)
            rows = np.arange(len(hot))
            for c in range(3):
                extra[rows, h[:, c]] = 1.0
        return ids, extra

    def oov_rows(self, start=0, stop=None):
        """float32 [stop-start, d] 3-hot rows of the OOV registry slice."""
        hot = self._oov_hot[start:stop]
        m = np.zeros((len(hot), self.dim), dtype=np.float32)
        for i, idx in enumerate(hot):
            m[i, list(idx)] = 1.0
        return m

    def oov_rows_at(self, ids):
        """float32 [len(ids), d] 3-hot rows of the OOV registry entries `ids` (registry indices,
        i.e. row id - n_rows), built with three vectorised scatters."""
        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        m = np.zeros((len(ids), self.dim), dtype=np.float32)
        if len(ids):
            hot = np.array([self._oov_hot[int(i)] for i in ids], dtype=np.int64).reshape(len(ids), 3)
            rows = np.arange(len(ids))
            for k in range(3):
                m[rows, hot[:, k]] = 1.0
        return m

    def vector(self, text):
        """float32 row exactly as mk_vectors would fill it (search.py:73-83)."""
        r = self.row_id(text)
        if r < self.n_rows:
            return self.table[r]
        return self.oov_rows(r - self.n_rows, r - self.n_rows + 1)[0]
