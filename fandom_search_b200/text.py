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


# ---------------------------------------------------------------------------------------------
# rule tokeniser: an APPROXIMATION of spaCy's English tokeniser (search.py:43-44 loads
# en_core_web_md with everything but the tokeniser disabled).  spaCy is not installed in the build
# image, so these rules are written from its documented behaviour (whitespace split, then prefix /
# suffix / infix punctuation and the common English contractions) and are NOT validated against it:
# CSV parity with a reference run holds for pre-tokenised text (tokens separated by spaces, which is
# what the whitespace tokeniser assumes); for raw prose pass the real thing as `tokenizer=`.
# ---------------------------------------------------------------------------------------------
import re  # noqa: E402

_PREFIX = re.compile(r"""^(?:\.\.\.|--+|[\[\](){}<>"'`“”‘’«»¡¿$£€#*_~|/\\,;:!?…-])""")
_SUFFIX = re.compile(r"""(?:\.\.\.|--+|[\[\](){}<>"'`“”‘’«»,;:!?…%*_~|/\\-]|(?<=[0-9a-zA-Z)\]"'’”])\.)$""")
_ABBREV = re.compile(r"^(?:[A-Za-z]\.){2,}$|^(?:Mr|Mrs|Ms|Dr|Prof|St|Jr|Sr|vs|etc|e\.g|i\.e)\.$")
_CONTRACTION = re.compile(r"^(.+?)(n't|n’t|'s|’s|'re|’re|'ve|’ve|'ll|’ll|'d|’d|'m|’m)$", re.IGNORECASE)
_INFIX = re.compile(r"(?<=[A-Za-z])(--+|—|–|\.\.\.|…|/)(?=[A-Za-z])")


def _split_word(word, out):
    prefixes, suffixes = [], []
    while word:
        if _ABBREV.match(word):
            break
        m = _PREFIX.match(word)
        if m and len(word) > len(m.group(0)):
            prefixes.append(m.group(0))
            word = word[len(m.group(0)):]
            continue
        m = _SUFFIX.search(word)
        if m and len(word) > len(m.group(0)):
            suffixes.append(m.group(0))
            word = word[:-len(m.group(0))]
            continue
        break
    out.extend(prefixes)
    if word:
        m = _CONTRACTION.match(word)
        if m:
            if m.group(2).lower() in ("n't", "n’t") and m.group(1).lower() == "ca":
                out.extend([m.group(1), m.group(2)])          # can't -> ca n't
            else:
                out.extend([m.group(1), m.group(2)])
        else:
            pieces = _INFIX.split(word)
            out.extend(p for p in pieces if p)
    out.extend(reversed(suffixes))


def tokenize_rules(text):
    """Whitespace split + punctuation / contraction rules (see the note above): 'Hello, world!' ->
    Hello , world ! ; "don't" -> do n't ; "it's" -> it 's ; 'U.S.' and 'Mr.' stay whole."""
    out = []
    for word in text.split():
        _split_word(word, out)
    return out


def glued_punctuation_share(words, sample=2000):
    """Share of (a sample of) tokens that start or end with punctuation although they contain
    letters or digits: raw prose through the whitespace tokeniser looks like this ('Hello,' /
    '"Why?'), pre-tokenised text does not."""
    n = bad = 0
    for w in words[:sample]:
        if len(w) < 2 or not any(c.isalnum() for c in w):
            continue
        n += 1
        if not (w[0].isalnum() and w[-1].isalnum()):
            bad += 1
    return bad / n if n else 0.0


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


# ---------------------------------------------------------------------------------------------
# native host pipeline (csrc/host_pipeline.cpp)
# ---------------------------------------------------------------------------------------------
import ctypes  # noqa: E402
import os  # noqa: E402


def _view(ptr, count, dtype):
    if count == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    ctype = {np.dtype(np.int64): ctypes.c_int64, np.dtype(np.int32): ctypes.c_int32,
             np.dtype(np.uint8): ctypes.c_uint8, np.dtype(np.uint32): ctypes.c_uint32,
             np.dtype(np.uint16): ctypes.c_uint16}[np.dtype(dtype)]
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=(count,))


class Vocab:
    """Native key -> embedding-row id map built from a Lexicon."""

    def __init__(self, lexicon):
        lib = nt.load()
        keys = list(lexicon.key_to_row)
        enc = [k.encode("utf-8") for k in keys]
        off = np.zeros(len(enc) + 1, dtype=np.int64)
        np.cumsum([len(e) for e in enc], out=off[1:])
        rows = np.array([lexicon.key_to_row[k] for k in keys], dtype=np.int32)
        blob = b"".join(enc)
        self._lib = lib
        self._h = lib.fs_vocab_create(blob, nt.ptr(off), nt.ptr(rows), len(enc))
        if not self._h:
            raise nt.NativeError(nt.FS_E_INVALID, lib.fs_last_error().decode())

    def lookup(self, word):
        b = word.encode("utf-8")
        return int(self._lib.fs_vocab_lookup(self._h, b, len(b)))

    def encode_files(self, paths, threads=None):
        if threads is None:
            # leave two cores to the thread that drives the GPU and to the post-processing thread:
            # with every core tokenising, the kernel launches of the search call queue up behind them;
            # under torchrun the node's cores are shared between the ranks (LOCAL_WORLD_SIZE)
            threads = host_threads()
        arr = (ctypes.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
        h = self._lib.fs_batch_encode_files(self._h, arr, len(paths), threads)
        if not h:
            raise nt.NativeError(nt.FS_E_INVALID, self._lib.fs_last_error().decode())
        return Batch(self._lib, h, list(paths))

    def __del__(self):
        try:
            if self._h:
                self._lib.fs_vocab_destroy(self._h)
                self._h = None
        except Exception:
            pass


class Batch:
    """One cluster of files, read + tokenised + encoded natively.  Arrays are zero-copy views of
    native memory and stay valid while this object is alive."""

    def __init__(self, lib, handle, paths):
        self._lib = lib
        self._h = handle
        self.paths = paths
        n_files = int(lib.fs_batch_info(handle, 0))
        n_tok = int(lib.fs_batch_info(handle, 1))
        n_oov = int(lib.fs_batch_info(handle, 2))
        n_text = int(lib.fs_batch_info(handle, 3))
        arr = lambda which: lib.fs_batch_array(handle, which)
        self.text = _view(arr(0), n_text, np.uint8)
        self.tok_off = _view(arr(2), n_files + 1, np.int64)
        self.tok = _view(arr(3), n_tok, np.int32)
        self.tok_start = _view(arr(4), n_tok, np.int64)
        self.tok_end = _view(arr(5), n_tok, np.int64)
        self.tok_start32 = _view(arr(9), n_tok, np.uint32)     # compact form for the device-side records
        self.tok_len16 = _view(arr(10), n_tok, np.uint16)
        self.oov_start = _view(arr(6), n_oov, np.int64)
        self.oov_end = _view(arr(7), n_oov, np.int64)
        status = _view(arr(8), n_files, np.int32)
        bad = np.nonzero(status)[0]
        if len(bad):
            raise FileNotFoundError("cannot read %s" % paths[int(bad[0])])

    @classmethod
    def from_token_lists(cls, lists):
        """Same structure from already-tokenised works (custom tokeniser path)."""
        self = cls.__new__(cls)
        self._lib = None
        self._h = None
        enc = [[w.encode("utf-8") for w in ws] for ws in lists]
        flat = [b for ws in enc for b in ws]
        lens = np.array([len(b) for b in flat], dtype=np.int64)
        self.tok_start = np.zeros(len(flat), dtype=np.int64)
        if len(flat):
            self.tok_start[1:] = np.cumsum(lens[:-1] + 1)
        self.tok_end = self.tok_start + lens
        self.tok_start32 = self.tok_start.astype(np.uint32)
        self.tok_len16 = np.minimum(lens, 65535).astype(np.uint16)
        self.text = np.frombuffer(b" ".join(flat), dtype=np.uint8)
        self.tok_off = np.zeros(len(lists) + 1, dtype=np.int64)
        np.cumsum([len(ws) for ws in lists], out=self.tok_off[1:])
        self.tok = np.full(len(flat), -1, dtype=np.int32)
        self.oov_start = np.zeros(0, dtype=np.int64)
        self.oov_end = np.zeros(0, dtype=np.int64)
        return self

    def token_text(self, pos):
        return self.text[int(self.tok_start[pos]):int(self.tok_end[pos])].tobytes().decode("utf-8")

    def oov_strings(self):
        return [self.text[int(s):int(e)].tobytes().decode("utf-8")
                for s, e in zip(self.oov_start.tolist(), self.oov_end.tolist())]

    def close(self):
        if self._h:
            self.text = self.tok = self.tok_start = self.tok_end = self.tok_off = None
            self.tok_start32 = self.tok_len16 = None
            self._lib.fs_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def host_threads(reserve=2, limit=16):
    """Host threads one rank may use for a native stage: the cores this process may run on, shared
    between the ranks of the node (LOCAL_WORLD_SIZE under torchrun), minus `reserve` cores for the
    threads that drive the GPU and write the CSVs."""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4)
    try:
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    except ValueError:
        local_world = 1
    share = cores // local_world
    if share <= 6:
        reserve = min(reserve, 1)      # few cores per rank: the GPU-driving thread sleeps in its waits
    return max(1, min(limit, share) - reserve)


def records_best(matches, tie, window, topk, batch, script_blob, script_off, threads=None):
    """Native search.py:182-226 core.  Returns dict of arrays (work, word, window_ix, match_ix,
    distance, lev) for the winning record of every matched fan word, sorted by (work, word)."""
    lib = nt.load()
    n = len(matches)
    matches = np.ascontiguousarray(matches)
    tie_arr = None if tie is None else np.ascontiguousarray(tie, dtype=np.int32)
    cap = max(64, 6 * n)
    out = None
    while True:
        out = {"work": np.empty(cap, np.int32), "word": np.empty(cap, np.int32),
               "window_ix": np.empty(cap, np.int32), "match_ix": np.empty(cap, np.int32),
               "distance": np.empty(cap, np.float64), "lev": np.empty(cap, np.int32)}
        text = np.ascontiguousarray(batch.text)
        rows = lib.fs_records_best_mt(
            nt.ptr(matches) if n else None, nt.ptr(tie_arr) if tie_arr is not None and n else None, n,
            window, topk, nt.ptr(text) if len(text) else None, nt.ptr(np.ascontiguousarray(batch.tok_start)),
            nt.ptr(np.ascontiguousarray(batch.tok_end)), nt.ptr(np.ascontiguousarray(batch.tok_off)),
            len(batch.tok_off) - 1, script_blob, nt.ptr(script_off), len(script_off) - 1,
            nt.ptr(out["work"]), nt.ptr(out["word"]), nt.ptr(out["window_ix"]), nt.ptr(out["match_ix"]),
            nt.ptr(out["distance"]), nt.ptr(out["lev"]), cap,
            host_threads(limit=8) if threads is None else threads)
        if rows < 0 and rows > -(1 << 62):
            cap = -rows
            continue
        if rows < 0:
            raise nt.NativeError(nt.FS_E_INVALID, lib.fs_last_error().decode())
        return {k: v[:rows] for k, v in out.items()}


def records_format_csv(best, filenames, batch, script_blob, script_off, script_orth, char_blob, char_off,
                       char_none, scene, scene_none, word_base=0):
    """CSV text (bytes) of the records in `best` (records_best output): search.py:206-217 rows as
    csv.writer writes them (search.py:331-334)."""
    lib = nt.load()
    rows = len(best["work"])
    if rows == 0:
        return b""
    names = [os.fspath(f).encode("utf-8") if not isinstance(f, bytes) else f for f in filenames]
    noff = np.zeros(len(names) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in names], out=noff[1:])
    text = np.ascontiguousarray(batch.text)
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    arrs = [c(best["work"], np.int32), c(best["word"], np.int32), c(best["window_ix"], np.int32),
            c(best["match_ix"], np.int32), c(best["distance"], np.float64), c(best["lev"], np.int32)]
    tok_start, tok_end, tok_off = (c(batch.tok_start, np.int64), c(batch.tok_end, np.int64),
                                   c(batch.tok_off, np.int64))
    script_off = c(script_off, np.int64)
    script_orth = c(script_orth, np.uint64)
    char_off, char_none = c(char_off, np.int64), c(char_none, np.uint8)
    scene, scene_none = c(scene, np.int64), c(scene_none, np.uint8)
    out = ctypes.c_void_p()
    n = lib.fs_records_format_csv(rows, *[nt.ptr(a) for a in arrs], b"".join(names), nt.ptr(noff),
                                  nt.ptr(text), nt.ptr(tok_start),
                                  nt.ptr(tok_end), nt.ptr(tok_off), script_blob, nt.ptr(script_off),
                                  nt.ptr(script_orth), char_blob, nt.ptr(char_off), nt.ptr(char_none),
                                  nt.ptr(scene), nt.ptr(scene_none), word_base, ctypes.byref(out))
    if n < 0:
        raise nt.NativeError(int(n), lib.fs_last_error().decode())
    try:
        return ctypes.string_at(out, n)
    finally:
        lib.fs_free(out)
