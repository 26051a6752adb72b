"""Device-side records: top-10 per fan window, Levenshtein, six records per pair and
the per-word argmin of search.py:182-226 on the GPU (fs_search_submit_rows) against the native host
implementation of the same step (fs_records_best_mt) on the same match list, against the reference's
golden CSVs, and against the CPU oracle on a quotation-heavy corpus (half of every work is planted
reuse) where the surviving pairs are dense."""
import glob
import os

import numpy as np
import pytest

from fandom_search_b200 import _native as nt
from fandom_search_b200 import search, synth
from fandom_search_b200 import text as _text
from fandom_search_b200.lexicon import Lexicon, py_hash_seed0
from tests.util import compare_records, normalise, read_csv

pytestmark = pytest.mark.gpu

FIELDS = ('work', 'word', 'window_ix', 'match_ix', 'distance', 'lev')


def _both_ways(index, files):
    """(device rows, host rows) of one cluster: the same search, records made on either side."""
    for attempt in range(3):
        # (a cluster whose matches or rows outgrow the default buffers is handed back as raw matches
        # once; the wrapper remembers the sizes and the next submit has room)
        prep = index.prepare(files)
        found = index.collect_prepared(prep, index.submit_prepared(prep, rows=True))
        if isinstance(found[0], search.DeviceRows):
            break
    assert isinstance(found[0], search.DeviceRows), "the device did not finish the records"
    dev = found[0]
    prep2 = index.prepare(files)
    matches, first_table = index.search_prepared(prep2)
    blob, soff = index._script_text()
    host = _text.records_best(matches, first_table, index.window_size, 10, prep2['batch'], blob, soff)
    return prep, dev, prep2, host, matches


def _assert_same(dev, host):
    dev = dev.best
    assert len(dev['work']) == len(host['work'])
    for k in FIELDS:
        assert np.array_equal(dev[k], host[k]), k          # bit for bit, float64 distance included


@pytest.mark.parametrize("mode", ["exhaustive", "lsh"])
def test_device_records_equal_host_records_on_the_golden_corpus(golden_dir, mode, monkeypatch):
    search.set_pipeline(search.Pipeline(
        Lexicon.from_npz(os.path.join(golden_dir, "lexicon.npz"), hash_fn=py_hash_seed0)))
    if mode == "lsh":
        monkeypatch.setenv("FANDOM_SEARCH_MODE", "lsh")
        monkeypatch.setenv("FANDOM_SEARCH_LSH_SEED", "7")
    try:
        index = search.AnnIndexSearch(os.path.join(golden_dir, "script.txt"), 6, 15, 14, 0.1)
        assert index.engine.index.device_records
        files = sorted(glob.glob(os.path.join(golden_dir, "fanworks", "*.txt")))
        prep, dev, prep2, host, matches = _both_ways(index, files)
        _assert_same(dev, host)
        assert len(dev) > 300
        # the work quoting the 11-fold repeated script 6-gram: the top-10 cut was made
        per_window = np.bincount(matches['fan_pos'] if mode == "exhaustive" else
                                 matches['fan_pos'][((matches['flags'] >> 8) & 0xFF) > 0])
        assert per_window.max() >= 11
        # and the whole drop-in path (search_many uses the device records) gives the reference's rows
        got = normalise([r for s in index.search_many(files) for r in s])
        want = read_csv(os.path.join(golden_dir, "golden_exhaustive.csv" if mode == "exhaustive"
                                     else "golden_lsh_seed7.csv"))
        compare_records(got, want, tol=1e-12)
    finally:
        search.set_pipeline(None)


def test_quotation_heavy_corpus(tmp_path):
    """Half of every work is reuse planted from the script (verbatim, near and far copies): the
    surviving pairs are dense, windows overlap heavily, every fan word sees up to six candidate
    windows.  Device records == host records == CPU oracle."""
    from oracle import reference_search as ora
    lex = synth.SynthLexicon(vocab=4000, dim=300, oov_frac=0.03, cased_frac=0.1, seed=31)
    lex_path = lex.save(str(tmp_path / "lexicon.npz"))
    script = synth.make_script_tokens(lex, 3000, seed=5)
    script[500:506] = script[100:106]                      # repeated 6-grams: several candidates per window
    script[900:912] = script[100:112]
    script_path = str(tmp_path / "script.txt")
    synth.write_markup_script(lex, script, script_path)
    fan_dir = tmp_path / "fanworks"
    fan_dir.mkdir()
    files = []
    for k in range(24):
        ids, planted = synth.make_fanwork_tokens(lex, script, k, mean_len=1500, sd_len=300, min_len=200,
                                                 max_len=3000, spans_mean=40.0, seed_base=7000)
        crng = np.random.default_rng(k)
        fn = str(fan_dir / ("%07d.txt" % k))
        with open(fn, "w", encoding="utf-8") as f:
            f.write(synth.fanwork_text(lex, ids, crng, 0.05))
        files.append(fn)
    search.set_pipeline(search.Pipeline(Lexicon.from_npz(lex_path, hash_fn=py_hash_seed0)))
    try:
        index = search.AnnIndexSearch(script_path, 6, 15, 14, 0.1)
        prep, dev, prep2, host, matches = _both_ways(index, files)
        _assert_same(dev, host)
        n_tok = int(prep2['offs'][-1])
        assert len(dev) > 0.25 * n_tok                     # dense: more than a quarter of all fan words carry a record
        got = normalise([r for s in index.records_prepared(prep, dev) for r in s])
        oracle = ora.OracleIndex(script_path, ora.OracleLexicon(lex_path, oov_hash=py_hash_seed0),
                                 mode="exhaustive", engine="dense")
        want = normalise([r for fn in files for r in oracle.search(fn)])
        compare_records(got, want, tol=1e-12)
    finally:
        search.set_pipeline(None)


def _tiny_case(tmp_path, fan_words, script_words, extra_keys=()):
    """A lexicon whose keys are the given words (random rows) plus extra_keys = (key, same-row-as),
    one script line and one fanwork."""
    rng = np.random.default_rng(3)
    aliased = {k for k, _ in extra_keys}
    words = sorted(set(script_words) | set(w for w in fan_words if w not in aliased))
    keys = list(words)
    rows = list(range(len(words)))
    for key, same_as in extra_keys:
        keys.append(key)
        rows.append(words.index(same_as))
    table = rng.standard_normal((len(words), 300)).astype(np.float32)
    lex_path = str(tmp_path / "lexicon.npz")
    np.savez(lex_path, keys=np.array(keys), rows=np.array(rows, dtype=np.int32), table=table)
    script_path = str(tmp_path / "script.txt")
    with open(script_path, "w", encoding="utf-8") as f:
        f.write("SCENE_NUMBER<<1>>\nCHARACTER_NAME<<A>>\nLINE<<%s>>\n" % " ".join(script_words))
    fan_path = str(tmp_path / "fan.txt")
    with open(fan_path, "w", encoding="utf-8") as f:
        f.write(" ".join(fan_words))
    return lex_path, script_path, fan_path


def test_unicode_tokens_and_levenshtein_over_code_points(tmp_path):
    script_words = ["café", "日本語", "naïve", "😀x", "plain", "zürich", "añb", "end", "of", "line", "here", "now"]
    fan_words = ["before", "Café", "日本語", "naïve", "😀x", "plain", "zürich", "añb", "END", "of", "line", "after", "x"]
    lex_path, script_path, fan_path = _tiny_case(
        tmp_path, fan_words, script_words,
        extra_keys=(("Café", "café"), ("END", "end"), ("before", "café"), ("after", "now"), ("x", "now")))
    search.set_pipeline(search.Pipeline(Lexicon.from_npz(lex_path, hash_fn=py_hash_seed0)))
    try:
        index = search.AnnIndexSearch(script_path, 6, 15, 14, 0.1)
        prep, dev, prep2, host, matches = _both_ways(index, [fan_path])
        _assert_same(dev, host)
        assert len(dev) >= 9 and dev.best['lev'].min() >= 7
        # the case-folded keys share rows with the lower-case ones: identical vectors, distance 0, but
        # the verbatim fan text differs from the lower-cased script text -> larger edit distance
        lev_of = dict(zip(dev.best['word'].tolist(), dev.best['lev'].tolist()))
        assert len(set(lev_of.values())) > 1
        from oracle import reference_search as ora
        oracle = ora.OracleIndex(script_path, ora.OracleLexicon(lex_path, oov_hash=py_hash_seed0),
                                 mode="exhaustive", engine="dense")
        got = normalise(index.search_many([fan_path])[0])
        compare_records(got, normalise(oracle.search(fan_path)), tol=1e-12)
    finally:
        search.set_pipeline(None)


def test_texts_too_long_for_the_device_fall_back_to_the_host_records(tmp_path):
    long_tok = "x" * 70000                       # a 70 000-byte token: beyond the device's 16-bit lengths
    mid = ["w%02d" % k + "y" * 60 for k in range(6)]         # 6 x 64 characters: both strings > 250 code points
    script_words = ["a", "b", "c", "d", "e", "f"] + mid + ["g", "h"]
    fan_words = ["q", "a", "b", "c", "d", "e", long_tok, "r"] + mid + ["s"]
    lex_path, script_path, fan_path = _tiny_case(
        tmp_path, fan_words, script_words,
        extra_keys=((long_tok, "f"), ("q", "g"), ("r", "h"), ("s", "g")))
    search.set_pipeline(search.Pipeline(Lexicon.from_npz(lex_path, hash_fn=py_hash_seed0)))
    try:
        index = search.AnnIndexSearch(script_path, 6, 15, 14, 0.1)
        prep = index.prepare([fan_path])
        found = index.collect_prepared(prep, index.submit_prepared(prep, rows=True))
        assert not isinstance(found[0], search.DeviceRows)       # FS_OVERFLOW_TEXT -> raw matches
        got = normalise(index.records_prepared(prep, *found)[0])
        from oracle import reference_search as ora
        oracle = ora.OracleIndex(script_path, ora.OracleLexicon(lex_path, oov_hash=py_hash_seed0),
                                 mode="exhaustive", engine="dense")
        want = normalise(oracle.search(fan_path))
        compare_records(got, want, tol=1e-12)
        assert max(r[10] for r in got) > 60000
        assert got == normalise(index.search_many([fan_path])[0])
    finally:
        search.set_pipeline(None)


def test_reuse_histogram_from_device_rows_equals_format_data(golden_dir, tmp_path, monkeypatch):
    """analyze(reuse_histogram=True): the table the device accumulates from its winning rows ==
    the thresholded group-by of the reference's `ao3.py format` (ao3.py:351-363,407-411, restated
    with the same pandas calls) applied to the reference's golden match CSV."""
    import argparse
    import pandas as pd
    from fandom_search_b200.aggregate import COLUMN_NAMES
    search.set_pipeline(search.Pipeline(
        Lexicon.from_npz(os.path.join(golden_dir, "lexicon.npz"), hash_fn=py_hash_seed0)))
    try:
        listing = open(os.path.join(golden_dir, "listing.txt")).read().split()
        real_listdir = os.listdir
        monkeypatch.setattr(os, "listdir", lambda d: list(listing) if str(d) == "fanworks" else real_listdir(d))
        monkeypatch.chdir(tmp_path)
        os.symlink(os.path.join(golden_dir, "fanworks"), "fanworks")
        os.symlink(os.path.join(golden_dir, "script.txt"), "script.txt")
        args = argparse.Namespace(fan_works="fanworks", script="script.txt", skip_works=-1, num_works=-1)
        search.analyze(args, chunk_size=16, reuse_histogram=True)
        got = pd.read_csv(glob.glob("match-6gram-2*-reuse.csv")[0], index_col='ORIGINAL_SCRIPT_WORD_INDEX')

        def format_data_counts(match_table):
            # --- the reference's own statements (ao3.py:351-363, 407-411)
            matches = pd.read_csv(match_table)
            name = 'Frequency of Reuse (Exact Matches)'
            matches_thresh = matches.assign(**{name: matches.BEST_COMBINED_DISTANCE <= 0})
            thresholds = [0.05, 0.1, 0.15, 0.2, 0.25, 0.3, 0.35, 0.4, 0.45, 0.5]
            threshname = ['Frequency of Reuse (0-{})'.format(str(t)) for t in thresholds]
            for thresh, name in zip(thresholds, threshname):
                matches_thresh = matches_thresh.assign(**{name: matches.BEST_COMBINED_DISTANCE <= thresh})
            threshname = ['Frequency of Reuse (Exact Matches)'] + threshname
            counts = matches_thresh.groupby('ORIGINAL_SCRIPT_WORD_INDEX').aggregate({n: 'sum' for n in threshname})
            return counts.reindex(got.index, fill_value=0), threshname

        # (1) exactly what format_data computes from the match CSV of THIS run -- without the CSV round trip
        own = [f for f in glob.glob("match-6gram-2*.csv") if not f.endswith("-reuse.csv")][0]
        want, threshname = format_data_counts(own)
        assert threshname == COLUMN_NAMES
        for col in threshname:
            assert (got[col].to_numpy() == want[col].to_numpy().astype(np.int64)).all(), col
        assert list(got['ORIGINAL_SCRIPT_WORD']) == list(search.AnnIndexSearch("script.txt", 6, 15, 14, 0.1).word_lowercase)
        # (2) against the reference's golden CSV: between exact-reuse alternates (identical window
        # vectors, combined distance +-1e-16: which of them wins is float noise in the reference
        # itself, tests/util.py) a count may sit on another script word and on either side of "<= 0";
        # every threshold above the noise must give the same totals
        ref, _ = format_data_counts(os.path.join(golden_dir, "golden_exhaustive.csv"))
        for col in threshname[1:]:
            assert int(got[col].sum()) == int(ref[col].sum()), col
        moved = int((got[threshname[1]].to_numpy() != ref[threshname[1]].to_numpy()).sum())
        assert moved <= 2 * 58            # (the 58 exact-reuse alternates of this corpus, tests/test_gpu_parity.py)
        assert got[threshname[0]].sum() > 0 and got[threshname[1]].sum() > got[threshname[0]].sum()
        assert (got[threshname].to_numpy()[:, 1:] >= got[threshname].to_numpy()[:, :-1]).all()
    finally:
        search.set_pipeline(None)
