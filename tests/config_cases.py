"""TEST INFRASTRUCTURE: the configuration-size parity cases (BASELINE.json configs C1/C2, d = 768,
and an adversarial near-threshold corpus).

Inputs are regenerated from fixed seeds wherever they are needed (the build container, the GPU
box); the EXPECTED outputs -- every (work, fan window, script window, float64 distance) under the
threshold and the CSV rows of search.py:188-226 -- come from the CPU oracle
(oracle.reference_search.OracleIndex, mode="exhaustive", engine="dense") and are committed under
tests/golden/config/ by oracle/make_config_golden.py.  An input digest stored with each fixture
guards against generator drift.
"""
import hashlib
import os

import numpy as np

from fandom_search_b200 import synth

GOLDEN_CONFIG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config")


class Case(object):
    """One corpus: lexicon (.npz), markup script, fanwork files."""

    def __init__(self, name, lex, script_ids, works, chunk_size=500):
        self.name = name
        self.lex = lex                    # object with .save(path) and .words
        self.script_ids = script_ids
        self.works = works                # list of word-id arrays
        self.chunk_size = chunk_size

    def digest(self):
        h = hashlib.sha256()
        h.update(np.asarray(self.script_ids, dtype=np.int64).tobytes())
        for w in self.works:
            h.update(np.asarray(w, dtype=np.int64).tobytes())
        h.update(np.ascontiguousarray(self.lex.table).tobytes())
        return h.hexdigest()

    def write(self, root):
        """Writes lexicon.npz, script.txt and fanworks/%07d.txt under `root`; returns
        (lexicon path, script path, list of fanwork paths in name order)."""
        os.makedirs(os.path.join(root, "fanworks"), exist_ok=True)
        lex_path = self.lex.save(os.path.join(root, "lexicon.npz"))
        script_path = os.path.join(root, "script.txt")
        synth.write_markup_script(self.lex, self.script_ids, script_path)
        files = []
        for k, ids in enumerate(self.works):
            fn = os.path.join(root, "fanworks", "%07d.txt" % k)
            with open(fn, "w", encoding="utf-8") as f:
                f.write(" ".join(self.lex.words[np.asarray(ids, dtype=np.int64)].tolist()))
            files.append(fn)
        return lex_path, script_path, files


def _synth_case(name, vocab, dim, oov_frac, n_script, n_works, chunk_size=500):
    lex = synth.SynthLexicon(vocab=vocab, dim=dim, oov_frac=oov_frac, seed=1001)
    script = synth.make_script_tokens(lex, n_script)
    works = [synth.make_fanwork_tokens(lex, script, k)[0] for k in range(n_works)]
    return Case(name, lex, script, works, chunk_size)


def c2_64():
    """C2 shape: the bench.py workload itself (same lexicon, script and fanwork seeds): the first
    64 works of cluster 0 against the 25 000-token script, d = 300."""
    return _synth_case("c2_64", 50000, 300, 0.0, 25000, 64)


def c1_500():
    """C1 whole: 500 works x ~5k words vs one 10k-word script (BASELINE.json configs[0]), with 2 %
    of the vocabulary out of the lexicon (3-hot OOV rule)."""
    return _synth_case("c1_500", 50000, 300, 0.02, 10000, 500)


def d768_24():
    """Wide embeddings: 24 works vs the 25 000-token script at d = 768."""
    return _synth_case("d768_24", 20000, 768, 0.01, 25000, 24)


class _AdvLexicon(object):
    def __init__(self, words, table):
        self.words = np.array(words)
        self.table = np.ascontiguousarray(table, dtype=np.float32)
        self.keys = self.words
        self.rows = np.arange(len(words), dtype=np.int32)

    def save(self, path):
        np.savez(path, keys=self.keys, rows=self.rows, table=self.table)
        return path


ADV_DELTAS = (1e-4, 3e-4, 1e-3, 3e-3, 1e-2)


def adversarial(n_planted=400, dim=300, window=6):
    """Hundreds of fan windows planted at cosine distance 0.1 +- {1e-4 ... 1e-2} from a script
    window, over an embedding table whose row norms span orders of magnitude (log-normal, sigma 1.2)
    and with every planted window rescaled as a whole by 10^U(-1, 1): the fp8 pre-filter must hand
    every pair just inside the threshold to the float64 decision and that decision must agree with
    the oracle on both sides of 0.1.  The planted rows are lexicon entries of their own
    ("adv00017_3"), so the whole case runs through text files like any corpus."""
    rng = np.random.default_rng(4242)
    base = synth.SynthLexicon(vocab=6000, dim=dim, oov_frac=0.0, seed=77)
    table = base.table_all * rng.lognormal(0.0, 1.2, size=(base.vocab, 1)).astype(np.float32)
    words = list(base.words)
    script = synth.make_script_tokens(base, 4000, seed=91)
    starts = rng.choice(np.arange(10, len(script) - 20, 9), n_planted, replace=False)
    rows = [table]
    works, cur = [], []
    planted = []
    for k, j in enumerate(starts.tolist()):
        s = table[script[j:j + window]].astype(np.float64).ravel()
        s_unit = s / np.linalg.norm(s)
        noise = rng.standard_normal(s.shape[0])
        noise -= (noise @ s_unit) * s_unit
        noise /= np.linalg.norm(noise)
        delta = 0.1 + (1 if k % 2 == 0 else -1) * ADV_DELTAS[(k // 2) % len(ADV_DELTAS)]
        cos = 1.0 - delta
        f = (cos * s_unit + np.sqrt(1.0 - cos * cos) * noise) * np.linalg.norm(s) * 10.0 ** rng.uniform(-1, 1)
        frows = f.reshape(window, dim).astype(np.float32)
        ids = []
        for r in range(window):
            ids.append(len(words))
            words.append("adv%05d_%d" % (k, r))
        rows.append(frows)
        gap = base.sample_words(rng, int(rng.integers(8, 40)))
        cur.extend(gap.tolist())
        planted.append((len(works), len(cur), int(j), delta))
        cur.extend(ids)
        if len(cur) > 1500:
            cur.extend(base.sample_words(rng, 12).tolist())
            works.append(np.array(cur, dtype=np.int64))
            cur = []
    if cur:
        cur.extend(base.sample_words(rng, 12).tolist())
        works.append(np.array(cur, dtype=np.int64))
    lex = _AdvLexicon(words, np.concatenate(rows, axis=0))
    case = Case("adversarial", lex, script, works)
    case.planted = planted
    return case


CASES = {"c2_64": c2_64, "c1_500": c1_500, "d768_24": d768_24, "adversarial": adversarial}
