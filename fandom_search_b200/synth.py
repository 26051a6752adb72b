"""Deterministic synthetic lexicon / markup script / fanwork corpora.

The reference ships no sample data (its fanworks/, scripts/, results/ directories are empty),
so every test and benchmark input is generated here from fixed seeds:

  vocabulary   V tokens "w%05d", Zipf(s=1) unigram frequencies, a fraction flagged OOV
               (absent from the lexicon -> exercises the 3-hot rule of search.py:79-83)
  embeddings   [V, d] float32, clustered: `cluster` words share a centre,
               e_v = 0.8 c + 0.6 n  (within-cluster cosine ~0.64, across ~0)
  script       markup with SCENE_NUMBER<<n>>, CHARACTER_NAME<<NAME>>, LINE<<...>> lines
               matching the regexes of search.py:291-293
  fanworks     Zipf background text with planted reuse spans copied from the script:
               1/2 verbatim, 1/4 one within-cluster substitution per 6 tokens (hit),
               1/4 one cross-cluster substitution per 6 tokens (miss)
"""
import os

import numpy as np

CHARACTER_NAMES = ["LUKE", "LEIA", "HAN", "VADER", "BEN", "THREEPIO", "TARKIN", "OWEN"]


class SynthLexicon:
    def __init__(self, vocab=50000, dim=300, cluster=10, oov_frac=0.02, cased_frac=0.0, seed=1001):
        rng = np.random.default_rng(seed)
        self.vocab = int(vocab)
        self.dim = int(dim)
        self.words = np.array(["w%05d" % i for i in range(vocab)])
        ranks = np.arange(1, vocab + 1, dtype=np.float64)
        p = 1.0 / ranks
        self.prob = p / p.sum()
        self.cdf = np.cumsum(self.prob)
        self.is_oov = np.zeros(vocab, dtype=bool)
        n_oov = int(round(vocab * oov_frac))
        if n_oov:
            self.is_oov[rng.choice(vocab, n_oov, replace=False)] = True
        rng2 = np.random.default_rng(seed + 1)
        n_clusters = (vocab + cluster - 1) // cluster
        centres = rng2.standard_normal((n_clusters, dim)).astype(np.float32)
        noise = rng2.standard_normal((vocab, dim)).astype(np.float32)
        # cluster membership is a random permutation so that clusters mix frequency ranks
        perm = rng2.permutation(vocab)
        self.cluster_of = np.empty(vocab, dtype=np.int64)
        self.cluster_of[perm] = np.arange(vocab) // cluster
        self.table_all = (0.8 * centres[self.cluster_of] + 0.6 * noise).astype(np.float32)
        self.members = [[] for _ in range(n_clusters)]
        for w, c in enumerate(self.cluster_of):
            self.members[c].append(w)
        # lexicon keys: every in-vocabulary word; optionally a capitalised variant
        # ("W00012") that shares the lower-case row for half of them and has its own row
        # for the rest (spaCy maps several keys onto one vectors row, and ORTH lookups
        # are case-sensitive)
        in_vocab = np.nonzero(~self.is_oov)[0]
        keys = [self.words[w] for w in in_vocab]
        rows = list(range(len(in_vocab)))
        table = [self.table_all[in_vocab]]
        self.row_of_word = np.full(vocab, -1, dtype=np.int64)
        self.row_of_word[in_vocab] = np.arange(len(in_vocab))
        n_cased = int(round(len(in_vocab) * cased_frac))
        self.cased = set()
        if n_cased:
            cased = rng.choice(in_vocab, n_cased, replace=False)
            extra_rows = []
            for k, w in enumerate(cased):
                self.cased.add(int(w))
                keys.append(self.words[w].upper())
                if k % 2 == 0:
                    rows.append(int(self.row_of_word[w]))
                else:
                    rows.append(len(in_vocab) + len(extra_rows))
                    extra_rows.append(rng2.standard_normal(dim).astype(np.float32))
            if extra_rows:
                table.append(np.stack(extra_rows))
        self.keys = np.array(keys)
        self.rows = np.array(rows, dtype=np.int32)
        self.table = np.concatenate(table, axis=0).astype(np.float32)

    def save(self, path):
        np.savez(path, keys=self.keys, rows=self.rows, table=self.table)
        return path

    def sample_words(self, rng, n):
        return np.searchsorted(self.cdf, rng.random(n), side="right").clip(0, self.vocab - 1)


def make_script_tokens(lex, n_tokens, seed=1003):
    rng = np.random.default_rng(seed)
    return lex.sample_words(rng, n_tokens)


def write_markup_script(lex, tokens, path, seed=1003):
    """Emit `tokens` (word ids) as a markup script; returns the per-token (scene, character)."""
    rng = np.random.default_rng(seed + 17)
    lines = []
    scene = 0
    meta = []
    pos = 0
    line_no = 0
    next_char_in = 0
    cur_char = None
    cur_scene = None
    n = len(tokens)
    while pos < n:
        if line_no % 100 == 0:
            scene += 1
            cur_scene = scene
            lines.append("SCENE_NUMBER<<%d>>" % scene)
            lines.append("SCENE_DESCRIPTION<<INT. SYNTHETIC SET %d>>" % scene)
        if next_char_in <= 0:
            cur_char = CHARACTER_NAMES[int(rng.integers(0, len(CHARACTER_NAMES)))]
            lines.append("CHARACTER_NAME<<%s>>" % cur_char)
            next_char_in = int(rng.integers(1, 5))
        ln = int(np.clip(rng.integers(6, 19), 1, n - pos))
        words = [lex.words[w] for w in tokens[pos:pos + ln]]
        lines.append("LINE<<%s>>" % " ".join(words))
        meta.extend([(cur_scene, cur_char)] * ln)
        pos += ln
        line_no += 1
        next_char_in -= 1
        if line_no % 7 == 0:
            lines.append("DIRECTION<<they move>>")
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(lines) + "\n")
    return meta


def make_fanwork_tokens(lex, script_tokens, work_id, mean_len=5000, sd_len=1000, min_len=50,
                        max_len=20000, spans_mean=3.0, seed_base=2000, window=6):
    """Word ids of one synthetic fanwork with planted reuse. Returns (ids, planted list)."""
    rng = np.random.default_rng(seed_base + work_id)
    length = int(np.clip(round(rng.normal(mean_len, sd_len)), min_len, max_len))
    ids = lex.sample_words(rng, length)
    planted = []
    n_spans = int(rng.poisson(spans_mean))
    ns = len(script_tokens)
    for _ in range(n_spans):
        span = int(rng.integers(window, 31))
        if span > length or span > ns:
            continue
        src = int(rng.integers(0, ns - span + 1))
        dst = int(rng.integers(0, length - span + 1))
        chunk = np.array(script_tokens[src:src + span])
        kind = rng.random()
        if kind >= 0.5:
            within = kind < 0.75
            for s in range(0, span, window):
                p = s + int(rng.integers(0, min(window, span - s)))
                w = int(chunk[p])
                if within:
                    mates = [m for m in lex.members[lex.cluster_of[w]] if m != w]
                    if mates:
                        chunk[p] = mates[int(rng.integers(0, len(mates)))]
                else:
                    chunk[p] = int(rng.integers(0, lex.vocab))
        ids[dst:dst + span] = chunk
        planted.append((dst, src, span, "verbatim" if kind < 0.5 else ("near" if kind < 0.75 else "far")))
    return ids, planted


def fanwork_text(lex, ids, cased_rng=None, case_prob=0.0):
    words = lex.words[ids]
    if cased_rng is not None and case_prob > 0:
        flip = cased_rng.random(len(ids)) < case_prob
        words = np.where(flip, np.char.upper(words), words)
    return " ".join(words.tolist())


def write_corpus(lex, script_tokens, out_dir, n_works, mean_len=5000, sd_len=1000, min_len=50,
                 max_len=20000, spans_mean=3.0, seed_base=2000, case_prob=0.0, first_id=0):
    """Write `n_works` fanworks as %07d.txt into out_dir. Returns total window count (w=6)."""
    os.makedirs(out_dir, exist_ok=True)
    windows = 0
    for k in range(first_id, first_id + n_works):
        ids, _ = make_fanwork_tokens(lex, script_tokens, k, mean_len, sd_len, min_len, max_len,
                                     spans_mean, seed_base)
        crng = np.random.default_rng(seed_base + 7919 * (k + 1)) if case_prob > 0 else None
        with open(os.path.join(out_dir, "%07d.txt" % k), "w", encoding="utf-8") as f:
            f.write(fanwork_text(lex, ids, crng, case_prob))
        windows += max(len(ids) - 5, 0)
    return windows


def synth_csr_batch(lex, script_tokens, work_ids, **kw):
    """Token-row-id CSR batch for a list of work ids, bypassing text (for benchmarks).
    OOV words get ids >= len(lex.table) with the extra rows returned alongside."""
    toks = []
    offs = [0]
    for k in work_ids:
        ids, _ = make_fanwork_tokens(lex, script_tokens, k, **kw)
        toks.append(ids)
        offs.append(offs[-1] + len(ids))
    words = np.concatenate(toks) if toks else np.zeros(0, dtype=np.int64)
    return words, np.array(offs, dtype=np.int64)
