"""The oracle restatement vs the golden CSVs produced by the unmodified reference search.py
(oracle/make_golden.py).  CPU only."""
import os
import subprocess
import sys

import pytest

from fandom_search_b200.lexicon import py_hash_seed0, siphash13
from oracle import reference_search as ora
from tests.util import compare_records, normalise, read_csv


@pytest.fixture(scope="module")
def lexicon(golden_dir):
    return ora.OracleLexicon(os.path.join(golden_dir, "lexicon.npz"), oov_hash=py_hash_seed0)


def _run(golden_dir, lexicon, **kw):
    index = ora.OracleIndex(os.path.join(golden_dir, "script.txt"), lexicon, **kw)
    fan_dir = os.path.join(golden_dir, "fanworks")
    records = []
    for name in sorted(os.listdir(fan_dir)):
        records.extend(index.search(os.path.join("fanworks", name) if False else os.path.join(fan_dir, name)))
    return index, normalise(records)


@pytest.mark.parametrize("engine", ["dense", "nearpy"])
def test_exhaustive_matches_reference_golden(golden_dir, lexicon, engine):
    index, got = _run(golden_dir, lexicon, mode="exhaustive", engine=engine)
    want = read_csv(os.path.join(golden_dir, "golden_exhaustive.csv"))
    compare_records(got, want, tol=1e-12)
    assert index.windows_processed == sum(
        max(len(open(os.path.join(golden_dir, "fanworks", f)).read().split()) - 5, 0)
        for f in os.listdir(os.path.join(golden_dir, "fanworks")))


@pytest.mark.parametrize("engine", ["dense", "nearpy"])
def test_seeded_lsh_matches_reference_golden(golden_dir, lexicon, engine):
    _, got = _run(golden_dir, lexicon, mode="lsh", seed=7, engine=engine)
    want = read_csv(os.path.join(golden_dir, "golden_lsh_seed7.csv"))
    compare_records(got, want, tol=1e-12)


def test_lsh_is_subset_of_exhaustive(golden_dir):
    ex = {(os.path.basename(r[0]), r[1]) for r in read_csv(os.path.join(golden_dir, "golden_exhaustive.csv"))}
    ls = {(os.path.basename(r[0]), r[1]) for r in read_csv(os.path.join(golden_dir, "golden_lsh_seed7.csv"))}
    assert ls <= ex and len(ls) < len(ex)


def test_batches_concatenate_to_aggregate(golden_dir):
    for name in ("golden_exhaustive", "golden_lsh_seed7"):
        agg = read_csv(os.path.join(golden_dir, name + ".csv"))
        parts = []
        for i in range(3):
            parts.extend(read_csv(os.path.join(golden_dir, "%s.batch%d.csv" % (name, i)), header=False))
        assert agg == parts


def test_shim_known_answers():
    assert ora.murmurhash64a(b"coffee", 1) == 3197928453018144401   # spaCy's documented id
    assert ora.levenshtein("kitten", "sitting") == 3
    assert ora.levenshtein("a b c d e f", "[a, b, c, d, e, f]") == 7
    assert ora.levenshtein("", "abc") == 3 and ora.levenshtein("flaw", "lawn") == 2


def test_python_hash_seed0_emulation():
    words = ["a", "w00012", "w00012w00012", "", "héllo", "日本語", "😀x", "abcdefgh", "abcdefghi"]
    code = "import sys\nfor s in sys.argv[1:]: print(hash(s))"
    out = subprocess.check_output([sys.executable, "-c", code] + words,
                                  env=dict(os.environ, PYTHONHASHSEED="0"), text=True)
    want = [int(x) for x in out.split()]
    assert [py_hash_seed0(w) for w in words] == want
    assert siphash13(b"", 0, 0) != 0


def test_config_size_fixture_is_what_the_oracle_computes(tmp_path):
    """tests/golden/config/adversarial.* (oracle/make_config_golden.py) replayed here: same inputs
    (digest), same pairs and distances from the oracle's exhaustive float64 engine."""
    import numpy as np
    from tests import config_cases
    case = config_cases.adversarial()
    z = np.load(os.path.join(config_cases.GOLDEN_CONFIG, "adversarial.npz"))
    assert case.digest() == str(z['digest'])
    lex_path, script_path, files = case.write(str(tmp_path))
    index = ora.OracleIndex(script_path, ora.OracleLexicon(lex_path, oov_hash=py_hash_seed0),
                            mode="exhaustive", engine="dense")
    got = []
    for k, fn in enumerate(files):
        index.search(fn)
        got.extend((k, i, j, d) for i, j, d in index.last_all_pairs)
    want = list(zip(z['work'].tolist(), z['fan'].tolist(), z['script'].tolist(), z['distance'].tolist()))
    assert [g[:3] for g in got] == [w[:3] for w in want]
    assert max(abs(g[3] - w[3]) for g, w in zip(got, want)) < 1e-13
    inside = {(w, f, j) for w, f, j, delta in case.planted if delta < 0.1}
    assert inside <= {g[:3] for g in got} and len(inside) == 200
    assert not ({(w, f, j) for w, f, j, delta in case.planted if delta > 0.1} & {g[:3] for g in got})
    for name in ("c2_64", "d768_24", "c1_500"):
        assert os.path.exists(os.path.join(config_cases.GOLDEN_CONFIG, name + ".npz"))
        assert os.path.exists(os.path.join(config_cases.GOLDEN_CONFIG, name + ".csv.gz"))
