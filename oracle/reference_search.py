"""TEST INFRASTRUCTURE -- CPU oracle for the reuse-search hot path.  Never shipped, never on
the product path: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this package.

What it is.  A numpy restatement of the algorithm of /root/reference/search.py (the
`ao3.py search` stage), each function citing the lines it follows.  The arithmetic of that
stage lives in three un-pinned, un-vendored third-party packages (nearpy, spaCy +
en_core_web_md, python-Levenshtein; /root/reference/requirements.txt:1-9) that are not
installed here; their published behaviour is restated in oracle/shims/ and marked [recalled].

PARITY STATUS: *unpinned against the third-party numerics* -- the reference ships no tests,
fixtures or golden outputs for this path, and nearpy/spaCy cannot be
run here.  What IS pinned: tests/golden/*.csv were produced by the UNMODIFIED
/root/reference/search.py executed over oracle/shims (script: oracle/make_golden.py), and this
restatement is checked against them (tests/test_oracle_golden.py), as are the shims'
known-answer vectors (MurmurHash64A("coffee") = 3197928453018144401, Levenshtein pairs).

Two engines produce the same records:
  engine="nearpy"  the reference's own per-window loop over the nearpy stand-in -- the cost
                   structure of the real thing; this is what the CPU baseline times;
  engine="dense"   vectorised float64 (blocked matmul); fast checker for larger cases.
Two modes: mode="lsh" (seeded random-hyperplane index, what the reference does up to its
unseeded randomness) and mode="exhaustive" (every window compared; superset of any LSH run).
"""
import csv
import datetime
import os
import random
import re
import sys
import zlib

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")

FIELDS = ['FAN_WORK_FILENAME', 'FAN_WORK_WORD_INDEX', 'FAN_WORK_WORD', 'FAN_WORK_ORTH_ID',
          'ORIGINAL_SCRIPT_WORD_INDEX', 'ORIGINAL_SCRIPT_WORD', 'ORIGINAL_SCRIPT_ORTH_ID',
          'ORIGINAL_SCRIPT_CHARACTER', 'ORIGINAL_SCRIPT_SCENE', 'BEST_MATCH_DISTANCE',
          'BEST_LEVENSHTEIN_DISTANCE', 'BEST_COMBINED_DISTANCE']   # search.py:20-33


def _shim(name):
    """Import one of the stand-in packages by path without touching sys.path globally."""
    import importlib.util
    key = "_oracle_shim_" + name
    if key in sys.modules:
        return sys.modules[key]
    path = os.path.join(_SHIMS, name, "__init__.py")
    spec = importlib.util.spec_from_file_location(key, path, submodule_search_locations=[
        os.path.join(_SHIMS, name)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod


def murmurhash64a(data, seed=1):
    return _shim("spacy").tokens.murmurhash64a(data, seed)


def levenshtein(a, b):
    return _shim("Levenshtein").distance(a, b)


def tokenize(text):
    """The spaCy stand-in's tokeniser: split on ASCII whitespace."""
    return [w.decode('utf-8') for w in text.encode('utf-8').split()]


class OracleLexicon(object):
    def __init__(self, path, oov_hash=hash):
        with numpy.load(path, allow_pickle=False) as z:
            self.table = numpy.asarray(z['table'], dtype=numpy.float32)
            self.key_to_row = {str(k): int(r) for k, r in zip(z['keys'], z['rows'])}
        self.dim = self.table.shape[1]
        self.oov_hash = oov_hash
        self._ids = {}

    def string_id(self, text):
        h = self._ids.get(text)
        if h is None:
            h = self._ids[text] = murmurhash64a(text.encode('utf-8'), 1)
        return h

    def vectors(self, words):
        """mk_vectors, search.py:65-84."""
        rows = len(words)
        cols = self.dim if rows else 0
        vectors = numpy.empty((rows, cols), dtype=float)
        for i, w in enumerate(words):
            r = self.key_to_row.get(w)
            if r is not None:
                vectors[i] = self.table[r]
            else:
                vectors[i] = 0
                vectors[i][self.oov_hash(w) % cols] = 1.0
                vectors[i][self.oov_hash(w * 2) % cols] = 1.0
                vectors[i][self.oov_hash(w * 3) % cols] = 1.0
        return vectors


def load_markup_script(filename, lexicon):
    """search.py:290-329 -> list of [lower_, lower id, scene, character] (no header row)."""
    line_rex = re.compile('LINE<<(?P<line>[^>]*)>>')
    scene_rex = re.compile('SCENE_NUMBER<<(?P<scene>[^>]*)>>')
    char_rex = re.compile('CHARACTER_NAME<<(?P<character>[^>]*)>>')
    rows = []
    current_scene = None
    current_scene_count = 0
    error_fix = False
    current_char = None
    with open(filename, encoding='utf-8') as ip:
        for line in ip:
            if scene_rex.search(line):
                current_scene_count += 1
                digits = ''.join(c for c in scene_rex.search(line).group('scene') if c.isdigit())
                try:
                    current_scene = int(digits)
                except ValueError:
                    error_fix = True
                if error_fix:
                    current_scene = current_scene_count
            elif char_rex.search(line):
                current_char = char_rex.search(line).group('character')
            elif line_rex.search(line):
                for w in tokenize(line_rex.search(line).group('line')):
                    lw = w.lower()
                    rows.append([lw, lexicon.string_id(lw), current_scene, current_char])
    return rows


def windows_of(vectors, w):
    """search.py:94-95 / 170-173: rolling concatenation -> [T-w+1, w*d] (empty if T < w)."""
    n = vectors.shape[0] - w + 1
    if n <= 0:
        return numpy.zeros((0, w * (vectors.shape[1] if vectors.ndim == 2 else 0)))
    return numpy.array([vectors[i:i + w, :].ravel() for i in range(n)])


def unit_rows(m):
    """nearpy unitvec per row: v / ||v||, unchanged when the norm is 0 [recalled]."""
    norms = numpy.sqrt((m * m).sum(axis=1))
    safe = numpy.where(norms > 0.0, norms, 1.0)
    return m / safe[:, None]


def exhaustive_gemm_pairs(script_vectors, fan_vectors, w=6, threshold=0.1, dtype=numpy.float32,
                          block=2048):
    """The "best honest CPU" comparator of SURVEY 8(d): the author's own exhaustive formula
    (_deprecated.py:1-26, 1 - (A@B)/(|A||B|)) as one BLAS GEMM per block of fan windows, in
    `dtype`.  Returns (fan window, script window) index arrays of the pairs under the threshold."""
    def fast_windows(v):
        v = numpy.ascontiguousarray(v, dtype=dtype)
        if v.shape[0] < w:
            return numpy.zeros((0, w * v.shape[1]), dtype=dtype)
        view = numpy.lib.stride_tricks.sliding_window_view(v, (w, v.shape[1]))
        return unit_rows(view.reshape(v.shape[0] - w + 1, w * v.shape[1])).astype(dtype)

    sw = script_vectors if getattr(script_vectors, 'ndim', 0) == 2 and script_vectors.shape[1] == \
        w * fan_vectors.shape[1] else fast_windows(script_vectors)
    fw = fast_windows(fan_vectors)
    fi, si = [], []
    for b in range(0, fw.shape[0], block):
        d = 1.0 - fw[b:b + block] @ sw.T
        i, j = numpy.nonzero(d < threshold)
        fi.append(i + b)
        si.append(j)
    if not fi:
        return numpy.zeros(0, numpy.int64), numpy.zeros(0, numpy.int64)
    return numpy.concatenate(fi), numpy.concatenate(si)


def lsh_seed(base, hash_name):
    return (int(base) + zlib.crc32(hash_name.encode('utf-8'))) % (2 ** 32)


class OracleIndex(object):
    """AnnIndexSearch, search.py:130-226."""

    def __init__(self, script_filename, lexicon, window_size=6, number_of_hashes=15,
                 hash_dimensions=14, distance_threshold=0.1, mode="exhaustive", seed=0,
                 engine="dense", unique=True):
        assert mode in ("exhaustive", "lsh") and engine in ("dense", "nearpy")
        self.lexicon = lexicon
        self.window_size = window_size
        self.distance_threshold = distance_threshold
        self.mode = mode
        self.engine_kind = engine
        self.unique = unique
        rows = load_markup_script(script_filename, lexicon)
        self.word_lowercase = [r[0] for r in rows]
        self.orth_id = [r[1] for r in rows]
        self.scene = [r[2] for r in rows]
        self.character = [r[3] for r in rows]
        self.match_str = [' '.join(self.word_lowercase[i:i + window_size])
                          for i in range(max(len(rows) - window_size + 1, 0))]   # search.py:123
        orig_vectors = lexicon.vectors(self.word_lowercase)        # script side is lower-cased (:151)
        orig_win = windows_of(orig_vectors, window_size)
        self.n_script_windows = orig_win.shape[0]
        self.windows_processed = 0
        self.normals = None
        if mode == "lsh":
            # 15 x RandomBinaryProjections('rbp{i}', 14), search.py:112-116, seeded for repeatability
            self.normals = [numpy.random.RandomState(lsh_seed(seed, 'rbp{}'.format(i)))
                            .randn(hash_dimensions, orig_win.shape[1] if orig_win.size else
                                   window_size * lexicon.dim)
                            for i in range(number_of_hashes)]
        if engine == "nearpy":
            nearpy = _shim("nearpy")
            env_backup = {k: os.environ.get(k) for k in
                          ('NEARPY_SHIM_SEED', 'NEARPY_SHIM_EXHAUSTIVE', 'NEARPY_SHIM_UNIQUE')}
            os.environ['NEARPY_SHIM_SEED'] = str(seed)
            os.environ['NEARPY_SHIM_EXHAUSTIVE'] = '1' if mode == "exhaustive" else '0'
            os.environ['NEARPY_SHIM_UNIQUE'] = '1' if unique else '0'
            try:
                hashes = [nearpy.hashes.RandomBinaryProjections('rbp{}'.format(i), hash_dimensions)
                          for i in range(number_of_hashes)]
                self.nearpy_engine = nearpy.Engine(window_size * lexicon.dim, lshashes=hashes,
                                                   distance=nearpy.distances.CosineDistance())
            finally:
                for k, v in env_backup.items():
                    if v is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v
            for ix, row in enumerate(orig_win):
                self.nearpy_engine.store_vector(row, (ix, self.match_str[ix]))
        else:
            self.script_unit = unit_rows(orig_win) if orig_win.size else orig_win
            if mode == "lsh":
                self.script_keys = self._keys(orig_win)

    def _keys(self, win):
        """[n, number_of_hashes] integer bucket keys: bit b of table t = (normals_t[b] . v > 0)."""
        keys = numpy.zeros((win.shape[0], len(self.normals)), dtype=numpy.int64)
        for t, nrm in enumerate(self.normals):
            bits = (win @ nrm.T) > 0.0
            keys[:, t] = (bits * (1 << numpy.arange(nrm.shape[0]))).sum(axis=1)
        return keys

    # -- candidate lists: per fan window, [(match_ix, distance)] as neighbours() + :182-184 --
    def _neighbours_dense(self, fan_win):
        thr = self.distance_threshold
        out = {}
        self.last_all_pairs = []
        if fan_win.shape[0] == 0 or self.n_script_windows == 0:
            return out
        fan_unit = unit_rows(fan_win)
        fan_keys = self._keys(fan_win) if self.mode == "lsh" else None
        block = max(1, (1 << 24) // max(self.n_script_windows, 1))
        for b0 in range(0, fan_unit.shape[0], block):
            d = 1.0 - fan_unit[b0:b0 + block] @ self.script_unit.T
            ii, jj = numpy.nonzero(d < thr)
            for i, j in zip(ii.tolist(), jj.tolist()):
                fan_ix = b0 + i
                if fan_keys is not None:
                    same = numpy.nonzero(fan_keys[fan_ix] == self.script_keys[j])[0]
                    if same.size == 0:
                        continue
                    tables = same.tolist() if not self.unique else [int(same[0])]
                else:
                    tables = [0]
                for t in tables:
                    out.setdefault(fan_ix, []).append((d[i, j], t, j))
        res = {}
        # (kept for the parity tests: every pair under the threshold, before the top-10 cut)
        self.last_all_pairs = [(fan_ix, j, dist) for fan_ix in sorted(out) for dist, t, j in out[fan_ix]]
        for fan_ix, cands in out.items():
            cands.sort()                       # (distance, first table, insertion order) == stable sort
            res[fan_ix] = [(j, dist) for dist, t, j in cands[:10]]   # NearestFilter(10)
        return res

    def _neighbours_nearpy(self, fan_win):
        res = {}
        for fan_ix, row in enumerate(fan_win):                  # search.py:176-184
            results = self.nearpy_engine.neighbours(row)
            results = [(match_ix, distance) for vec, (match_ix, match_str), distance in results
                       if distance < self.distance_threshold]
            if results:
                res[fan_ix] = results
        return res

    def search_words(self, fan, filename):
        """search.py:168-226 given the token texts of one work."""
        w = self.window_size
        fan_vectors = self.lexicon.vectors(fan)
        fan_win = windows_of(fan_vectors, w)
        self.windows_processed += fan_win.shape[0]
        if self.engine_kind == "nearpy":
            neigh = self._neighbours_nearpy(fan_win)
        else:
            neigh = self._neighbours_dense(fan_win)
        duplicate_records = {}
        for fan_ix in sorted(neigh):
            for match_ix, distance in neigh[fan_ix]:
                fan_context = '[' + ', '.join(fan[fan_ix:fan_ix + w]) + ']'    # str(list of Token)
                lev_d = levenshtein(self.match_str[match_ix], fan_context)
                for window_ix in range(w):
                    fan_word_ix = fan_ix + window_ix
                    orig_word_ix = match_ix + window_ix
                    duplicate_records.setdefault((filename, fan_word_ix), []).append(
                        [filename, fan_word_ix, fan[fan_word_ix],
                         self.lexicon.string_id(fan[fan_word_ix]), orig_word_ix,
                         self.word_lowercase[orig_word_ix], self.orth_id[orig_word_ix],
                         self.character[orig_word_ix], self.scene[orig_word_ix],
                         distance, lev_d, distance * lev_d])
        self.last_candidates = duplicate_records
        best = [min(dset, key=lambda r: r[11]) for dset in duplicate_records.values()]
        return sorted(best, key=lambda r: (r[0], r[1]))

    def search(self, filename):
        with open(filename, encoding='utf8') as f:
            fan = tokenize(f.read())
        return self.search_words(fan, filename)

    def all_pairs_words(self, fan):
        """Every (fan_ix, match_ix, distance) with distance < threshold, BEFORE the top-10 cut of
        NearestFilter: the full float64 match set of search.py:176-184 under exhaustive candidates
        (dense engine only) -- what the GPU path's match list must equal pair for pair."""
        assert self.engine_kind == "dense" and self.mode == "exhaustive"
        self._neighbours_dense(windows_of(self.lexicon.vectors(fan), self.window_size))
        return list(self.last_all_pairs)

    def pairs_words(self, fan):
        """All (fan_ix, match_ix, distance) under the threshold after the top-10 filter."""
        fan_win = windows_of(self.lexicon.vectors(fan), self.window_size)
        neigh = (self._neighbours_nearpy if self.engine_kind == "nearpy"
                 else self._neighbours_dense)(fan_win)
        return [(i, j, d) for i in sorted(neigh) for j, d in neigh[i]]


def write_records(records, filename):
    with open(filename, 'w', encoding='utf-8') as out:      # search.py:331-334
        csv.writer(out).writerows(records)


def analyze(fan_work_directory, script, lexicon, out_dir='.', skip_works=-1, num_works=-1,
            window_size=6, number_of_hashes=15, hash_dimensions=14, distance_threshold=0.1,
            chunk_size=500, **index_kwargs):
    """search.py:336-399, single process.  Returns (aggregate file name, records)."""
    subsample_start = 0 if skip_works < 0 else skip_works
    subsample_end = None if num_works < 0 else num_works + subsample_start
    fan_works = [os.path.join(fan_work_directory, f) for f in os.listdir(fan_work_directory)]
    random.seed(4815162342)
    random.shuffle(fan_works)
    fan_works = fan_works[subsample_start:subsample_end]
    clusters = [fan_works[i:i + chunk_size] for i in range(0, len(fan_works), chunk_size)]
    index = OracleIndex(script, lexicon, window_size, number_of_hashes, hash_dimensions,
                        distance_threshold, **index_kwargs)
    accumulated = [FIELDS]
    for i, cluster in enumerate(clusters):
        records = [r for fn in cluster for r in index.search(fn)]
        write_records(records, os.path.join(out_dir, 'match-{}gram-batch-{}.csv'.format(window_size, i)))
        accumulated.extend(records)
    k = 0
    name = os.path.join(out_dir, 'match-{}gram-{:%Y%m%d}.csv'.format(window_size, datetime.date.today()))
    while os.path.exists(name):
        k += 1
        name = os.path.join(out_dir, 'match-{}gram-{:%Y%m%d}-{}.csv'.format(
            window_size, datetime.date.today(), k))
    write_records(accumulated, name)
    return name, accumulated
