"""CPU tests of the host side of the drop-in search.py (tokenise -> CSR row ids -> OOV
extras -> records -> CSV) with the device replaced by tests/numpy_index.NumpyIndex.
The GPU parity tests (tests/test_gpu_parity.py) run the same flow on the real device."""
import argparse
import glob
import os

import pytest

import fandom_search_b200.engine as engine_mod
from fandom_search_b200 import search
from fandom_search_b200.lexicon import Lexicon, py_hash_seed0
from oracle import reference_search as ora
from tests.numpy_index import NumpyIndex
from tests.util import compare_records, normalise, read_csv


@pytest.fixture()
def cpu_device(monkeypatch, golden_dir):
    monkeypatch.setattr(engine_mod, "DeviceIndex", NumpyIndex)
    search.set_pipeline(search.Pipeline(
        Lexicon.from_npz(os.path.join(golden_dir, "lexicon.npz"), hash_fn=py_hash_seed0)))
    yield
    search.set_pipeline(None)


def test_load_markup_script_matches_oracle(cpu_device, golden_dir):
    rows = search.load_markup_script(os.path.join(golden_dir, "script.txt"))
    assert rows[0] == ['LOWERCASE', 'SPACY_ORTH_ID', 'SCENE', 'CHARACTER']
    lex = ora.OracleLexicon(os.path.join(golden_dir, "lexicon.npz"), oov_hash=py_hash_seed0)
    assert rows[1:] == ora.load_markup_script(os.path.join(golden_dir, "script.txt"), lex)
    assert rows[-1][2] == 2 and rows[-1][3] == 'LEIA'     # scene-number fallback to the running count


def test_search_many_matches_reference_golden(cpu_device, golden_dir):
    idx = search.AnnIndexSearch(os.path.join(golden_dir, "script.txt"), 6, 15, 14, 0.1)
    files = sorted(glob.glob(os.path.join(golden_dir, "fanworks", "*.txt")))
    sets = idx.search_many(files)
    got = normalise([r for s in sets for r in s])
    want = read_csv(os.path.join(golden_dir, "golden_exhaustive.csv"))
    compare_records(got, want, tol=1e-12)
    assert idx.windows_processed > 0
    # single-file entry point == the batched one
    k = files.index(os.path.join(golden_dir, "fanworks", "0000003.txt"))
    assert normalise(idx.search(files[k])) == normalise(sets[k])
    # edge cases: 5-token, empty and exactly-6-token works
    assert sets[files.index(os.path.join(golden_dir, "fanworks", "0000037.txt"))] == []
    assert sets[files.index(os.path.join(golden_dir, "fanworks", "0000038.txt"))] == []
    assert len(sets[files.index(os.path.join(golden_dir, "fanworks", "0000039.txt"))]) == 6


def test_analyze_writes_reference_files(cpu_device, golden_dir, tmp_path, monkeypatch):
    listing = open(os.path.join(golden_dir, "listing.txt")).read().split()
    real_listdir = os.listdir
    monkeypatch.setattr(os, "listdir", lambda d: list(listing) if str(d) == "fanworks" else real_listdir(d))
    monkeypatch.chdir(tmp_path)
    os.symlink(os.path.join(golden_dir, "fanworks"), "fanworks")
    os.symlink(os.path.join(golden_dir, "script.txt"), "script.txt")
    args = argparse.Namespace(fan_works="fanworks", script="script.txt", skip_works=-1, num_works=-1)
    search.analyze(args, chunk_size=16)
    aggs = glob.glob("match-6gram-2*.csv")
    assert len(aggs) == 1
    got = read_csv(aggs[0])
    want = read_csv(os.path.join(golden_dir, "golden_exhaustive.csv"))
    compare_records(got, want, tol=1e-12, basename=False)
    assert [(r[0], r[1]) for r in got] == [(r[0], r[1]) for r in want]      # same row order
    header = open(aggs[0], newline='').readline()
    assert header == ",".join(search.new_record_structure['fields']) + "\r\n"
    for i in range(3):
        b = read_csv("match-6gram-batch-%d.csv" % i, header=False)
        wb = read_csv(os.path.join(golden_dir, "golden_exhaustive.batch%d.csv" % i), header=False)
        assert [(r[0], r[1]) for r in b] == [(r[0], r[1]) for r in wb]
    # a second run the same day must not overwrite the aggregate (search.py:390-397)
    search.analyze(args, chunk_size=16)
    assert len(glob.glob("match-6gram-2*.csv")) == 2
    # -s / -n sub-sampling (search.py:345-358)
    args2 = argparse.Namespace(fan_works="fanworks", script="script.txt", skip_works=4, num_works=3)
    for f in glob.glob("match-6gram-*.csv"):
        os.remove(f)
    search.analyze(args2, chunk_size=16)
    sub = read_csv(glob.glob("match-6gram-2*.csv")[0])
    import random
    order = [os.path.join("fanworks", f) for f in listing]
    random.seed(4815162342)
    random.shuffle(order)
    assert {r[0] for r in sub} <= set(order[4:7])


def test_mk_vectors_and_chunks(cpu_device):
    toks = search.get_spacy_model()("w00001 W00001 notaword w00002")
    v = search.mk_vectors(toks)
    assert v.shape == (4, 300) and v.dtype == float
    assert sorted(set(v[2].tolist())) == [0.0, 1.0] and 1 <= v[2].sum() <= 3
    hot = {py_hash_seed0("notaword") % 300, py_hash_seed0("notaword" * 2) % 300,
           py_hash_seed0("notaword" * 3) % 300}
    assert set(v[2].nonzero()[0].tolist()) == hot
    assert search.mk_vectors([]).shape == (0, 0)
    long_text = " ".join(["w00001"] * 20000)          # 139999 chars -> two chunks
    chunks = list(search.sp_parse_chunks(long_text))
    assert len(chunks) == 2 and sum(len(c) for c in chunks) == 20000


def test_custom_tokenizer_path_matches_golden(monkeypatch, golden_dir):
    """A user-supplied tokeniser (e.g. spaCy's) bypasses the native encoder; same records."""
    monkeypatch.setattr(engine_mod, "DeviceIndex", NumpyIndex)
    lex = Lexicon.from_npz(os.path.join(golden_dir, "lexicon.npz"), hash_fn=py_hash_seed0)
    search.set_pipeline(search.Pipeline(lex, tokenizer=lambda text: text.split()))
    try:
        idx = search.AnnIndexSearch(os.path.join(golden_dir, "script.txt"), 6, 15, 14, 0.1)
        files = sorted(glob.glob(os.path.join(golden_dir, "fanworks", "*.txt")))
        got = normalise([r for s in idx.search_many(files) for r in s])
        compare_records(got, read_csv(os.path.join(golden_dir, "golden_exhaustive.csv")), tol=1e-12)
    finally:
        search.set_pipeline(None)


def test_native_vocab_and_batch(golden_dir, tmp_path):
    from fandom_search_b200 import text
    lex = Lexicon.from_npz(os.path.join(golden_dir, "lexicon.npz"), hash_fn=py_hash_seed0)
    vocab = text.Vocab(lex)
    for k, r in list(lex.key_to_row.items())[:50]:
        assert vocab.lookup(k) == r
    assert vocab.lookup("definitely-not-a-word") == -1 and vocab.lookup("") == -1
    (tmp_path / "a.txt").write_text("w00001  zzz W00001\nzzz\tqqq é日本", encoding="utf-8")
    (tmp_path / "b.txt").write_text("", encoding="utf-8")
    (tmp_path / "c.txt").write_text("qqq", encoding="utf-8")
    b = vocab.encode_files([str(tmp_path / n) for n in ("a.txt", "b.txt", "c.txt")], threads=3)
    assert b.tok_off.tolist() == [0, 6, 6, 7]
    words = [b.token_text(i) for i in range(7)]
    assert words == ["w00001", "zzz", "W00001", "zzz", "qqq", "é日本", "qqq"]
    want_oov = []
    for w in words:
        if w not in lex.key_to_row and w not in want_oov:
            want_oov.append(w)
    assert b.oov_strings() == want_oov and want_oov[0] == "zzz"
    tok = b.tok.tolist()
    for w, t in zip(words, tok):
        assert t == (lex.key_to_row[w] if w in lex.key_to_row else -(1 + want_oov.index(w)))
    with pytest.raises(FileNotFoundError):
        vocab.encode_files([str(tmp_path / "missing.txt")])


def test_native_vocab_short_and_long_keys(tmp_path):
    """The vocabulary table keeps keys of <= 8 bytes inline and longer ones by hash: every length,
    keys that are prefixes of each other, and near-misses must resolve like a Python dict."""
    import types
    import numpy as np
    from fandom_search_b200 import text
    rng = np.random.default_rng(77)
    alphabet = "abcdé日x"
    keys = {}
    for n in range(1, 21):
        for _ in range(60):
            w = "".join(alphabet[i] for i in rng.integers(0, len(alphabet), n))
            keys.setdefault(w, len(keys) * 3 + 1)
    for w in ("a", "aa", "aaaaaaaa", "aaaaaaaaa", "aaaaaaa", "aaaaaaab"):
        keys.setdefault(w, len(keys) * 3 + 1)
    vocab = text.Vocab(types.SimpleNamespace(key_to_row=keys))
    for w, r in keys.items():
        assert vocab.lookup(w) == r, w
    misses = [w + "a" for w in keys if w + "a" not in keys] + [w[:-1] for w in keys if w[:-1] not in keys]
    for w in misses:
        assert vocab.lookup(w) == -1, w
    # through the file tokeniser (8-byte loads at the very end of the text buffer included)
    words = list(keys)[::7] + misses[::11]
    words = [w for w in words if w]
    (tmp_path / "t.txt").write_text(" ".join(words), encoding="utf-8")
    b = vocab.encode_files([str(tmp_path / "t.txt")], threads=2)
    oov = b.oov_strings()
    for w, t in zip(words, b.tok.tolist()):
        assert t == (keys[w] if w in keys else -(1 + oov.index(w))), w


def test_native_oov_numbering_many_and_long_words(tmp_path):
    """Unique out-of-vocabulary strings are numbered in order of first appearance over the batch
    (files in order): several thousand of them (the table grows), lengths 1..30, words that share
    their first 8 bytes, across three files tokenised by different threads."""
    import types
    import numpy as np
    from fandom_search_b200 import text
    rng = np.random.default_rng(5)
    known = {"k%d" % i: i for i in range(50)}
    pool = []
    for i in range(6000):
        n = int(rng.integers(1, 31))
        pool.append(("q%dé" % i + "abcdefgh" * 4)[:n] if i % 3 else "samehead" + str(i))
    pool = [w for w in dict.fromkeys(pool) if w not in known]
    files = []
    all_words = []
    for f in range(3):
        words = [pool[int(j)] if rng.random() < 0.7 else "k%d" % int(rng.integers(0, 50))
                 for j in rng.integers(0, len(pool), 9000)]
        (tmp_path / ("f%d.txt" % f)).write_text(" ".join(words), encoding="utf-8")
        files.append(str(tmp_path / ("f%d.txt" % f)))
        all_words += words
    vocab = text.Vocab(types.SimpleNamespace(key_to_row=known))
    b = vocab.encode_files(files, threads=3)
    want_oov = list(dict.fromkeys(w for w in all_words if w not in known))
    assert len(want_oov) > 2048
    assert b.oov_strings() == want_oov
    index = {w: i for i, w in enumerate(want_oov)}
    want_tok = [known[w] if w in known else -(1 + index[w]) for w in all_words]
    assert b.tok.tolist() == want_tok


def test_multi_script_pass_equals_separate_runs(cpu_device, golden_dir, tmp_path, monkeypatch):
    """N4: indexing two scripts side by side and searching once == two single-script runs."""
    lines = open(os.path.join(golden_dir, "script.txt"), encoding="utf-8").read().splitlines()
    cut = len(lines) // 2
    a, b = tmp_path / "alpha.txt", tmp_path / "beta.txt"
    a.write_text("\n".join(lines[:cut]) + "\n", encoding="utf-8")
    b.write_text("\n".join(lines[cut:]) + "\n", encoding="utf-8")
    files = sorted(glob.glob(os.path.join(golden_dir, "fanworks", "*.txt")))
    both = search.AnnIndexSearch([str(a), str(b)], 6, 15, 14, 0.1)
    multi = both.search_many_scripts(files)
    for k, path in enumerate((a, b)):
        single = search.AnnIndexSearch(str(path), 6, 15, 14, 0.1)
        want = normalise([r for s in single.search_many(files) for r in s])
        got = normalise([r for s in multi[k] for r in s])
        assert len(want) > 0
        compare_records(got, want, tol=1e-12)
    # the driver writes one set of files per script
    monkeypatch.chdir(tmp_path)
    args = argparse.Namespace(fan_works=os.path.join(golden_dir, "fanworks"), script=None,
                              skip_works=-1, num_works=-1)
    search.analyze_scripts(args, [str(a), str(b)], chunk_size=16)
    assert len(glob.glob("match-6gram-alpha-batch-*.csv")) == 3
    assert len(glob.glob("match-6gram-beta-2*.csv")) == 1
    agg = read_csv(glob.glob("match-6gram-alpha-2*.csv")[0])
    single = search.AnnIndexSearch(str(a), 6, 15, 14, 0.1)
    compare_records(agg, normalise([r for s in single.search_many(files) for r in s]), tol=1e-12)


def test_native_csv_rows_match_python_csv_writer():
    """fs_records_format_csv == csv.writer(...).writerows on the same records (search.py:331-334),
    with quoting, None fields, unicode and float repr corner cases."""
    import numpy as np
    from fandom_search_b200 import text as T, search as S
    words = ['plain', 'com,ma', 'quo"te', '"', ',', 'üñí', 'a"b"c', "it's", 'semi;colon', '日本語', 'x']
    lists = [[words[(i*3+j) % len(words)] for j in range(9)] for i in range(4)]
    batch = T.Batch.from_token_lists(lists)
    filenames = ['dir/a,b.txt', 'dir/"q".txt', 'dir/ü.txt', 'plain.txt']
    script_words = ['may', 'the', 'fo,rce', 'be', '"with"', 'you', 'always', 'ok']
    enc = [w.encode() for w in script_words]
    soff = np.zeros(len(enc)+1, np.int64); np.cumsum([len(e) for e in enc], out=soff[1:])
    blob = b''.join(enc)
    orth = np.array([T.string_id(w) for w in script_words], np.uint64)
    chars = ['LUKE', None, 'HAN, SOLO', 'C"3PO', None, 'LEIA', 'OBI\nWAN', 'R2']
    scenes = [1, None, 3, 40000000000, None, 6, 7, 8]
    cenc = [b'' if c is None else c.encode() for c in chars]
    coff = np.zeros(len(cenc)+1, np.int64); np.cumsum([len(e) for e in cenc], out=coff[1:])
    cnone = np.array([c is None for c in chars], np.uint8)
    snone = np.array([s is None for s in scenes], np.uint8)
    scene = np.array([0 if s is None else s for s in scenes], np.int64)
    rng = np.random.default_rng(0)
    rows = 30
    best = {'work': np.sort(rng.integers(0, 4, rows)).astype(np.int32), 'word': rng.integers(0, 9, rows).astype(np.int32),
            'window_ix': rng.integers(0, 3, rows).astype(np.int32), 'match_ix': rng.integers(0, 6, rows).astype(np.int32),
            'distance': np.concatenate([[0.0, 1e-17, 2.220446049250313e-16, 0.1], rng.random(rows-4)*0.1]), 'lev': rng.integers(0, 40, rows).astype(np.int32)}
    got = T.records_format_csv(best, filenames, batch, blob, soff, orth, b''.join(cenc), coff, cnone, scene, snone, 0)
    recs = []
    for i in range(rows):
        w = int(best['work'][i]); g = int(best['match_ix'][i] + best['window_ix'][i])
        fw = lists[w][int(best['word'][i])]
        d = float(best['distance'][i]); l = int(best['lev'][i])
        recs.append([filenames[w], int(best['word'][i]), fw, T.string_id(fw), g, script_words[g], int(orth[g]), chars[g], scenes[g], d, l, d*l])
    want = S.format_records(recs).encode()
    assert got == want

def test_native_float_repr_matches_python():
    import ctypes, random, struct
    from fandom_search_b200 import _native as nt
    lib = nt.load()
    buf = ctypes.create_string_buffer(64)
    random.seed(5)
    cases = [0.0, -0.0, 1.0, 0.1, 1e-4, 1e-5, 9.999e-5, 1e15, 1e16, 1e17, 123456789012345678.0, 5e-324,
             2.2250738585072014e-308, 1.7976931348623157e308, -2.220446049250313e-16,
             0.05000000000000004, 1e22, 1e23, float('inf'), float('-inf')]
    for _ in range(20000):
        cases.append(struct.unpack('d', struct.pack('Q', random.getrandbits(64)))[0])
        cases.append(random.uniform(0, 0.1) * random.randint(0, 60))
    for x in cases:
        if x != x:
            continue
        n = lib.fs_format_py_float(x, buf, 64)
        assert buf.value.decode() == repr(x), (repr(x), buf.value)
        assert n == len(repr(x))


def test_levenshtein_bit_vector_and_fallback_match_the_oracle():
    """fs_levenshtein_utf8 (Myers/Hyyro bit-vector up to 64 code points, two-row DP beyond) against
    the oracle's plain DP on 30 k random pairs: ASCII, accents, CJK, astral plane, lengths 0-90."""
    import random
    from fandom_search_b200 import text as T
    from oracle import reference_search as ora
    random.seed(3)
    alph = "ab cdé日本,[]😀xyz"
    bad=0
    for it in range(30000):
        la=random.randint(0,90); lb=random.randint(0,90)
        if it%3==0: la=min(la,20); lb=min(lb,20)
        a=''.join(random.choice(alph) for _ in range(la))
        if random.random()<0.5:
            # b = mutated a
            b=list(a)
            for _ in range(random.randint(0,6)):
                r=random.random()
                if r<0.33 and b: b.pop(random.randrange(len(b)))
                elif r<0.66: b.insert(random.randint(0,len(b)), random.choice(alph))
                elif b: b[random.randrange(len(b))]=random.choice(alph)
            b=''.join(b)
        else:
            b=''.join(random.choice(alph) for _ in range(lb))
        x=T.levenshtein(a,b); y=ora.levenshtein(a,b)
        if x!=y:
            bad+=1
    assert bad == 0

def test_records_best_against_python_restatement():
    """fs_records_best (sorted single pass with a ring of open positions) == search.py:182-226 restated
    in plain Python: top-10 per window, Levenshtein, six records per pair, first minimal record wins."""
    import numpy as np
    from fandom_search_b200 import text as T, _native as nt
    from oracle import reference_search as ora
    rng = np.random.default_rng(5)
    vocab=["w%03d"%i for i in range(300)]+["é%d"%i for i in range(20)]
    for trial in range(40):
        nw=int(rng.integers(1,6))
        works=[[vocab[i] for i in rng.integers(0,len(vocab),int(rng.integers(6,60)))] for _ in range(nw)]
        batch=T.Batch.from_token_lists(works)
        script=[vocab[i] for i in rng.integers(0,len(vocab),80)]
        enc=[w.encode() for w in script]; soff=np.zeros(len(enc)+1,np.int64); np.cumsum([len(e) for e in enc],out=soff[1:]); blob=b''.join(enc)
        tokoff=np.asarray(batch.tok_off)
        n=int(rng.integers(0,120))
        m=np.zeros(n,dtype=nt.MATCH_DTYPE)
        recs=[]
        for i in range(n):
            w=int(rng.integers(0,nw)); L=len(works[w])
            fp=int(rng.integers(0,L-5)); sp=int(rng.integers(0,75))
            m[i]=(tokoff[w]+fp, sp, float(rng.choice([0.0,0.01,0.05,0.05,0.09])), w, 0)
        # unique (fan_pos, script_pos)
        _,ui=np.unique(np.stack([m['fan_pos'],m['script_pos']],1),axis=0,return_index=True); m=m[np.sort(ui)]
        best=T.records_best(m,None,6,10,batch,blob,soff)
        # python restatement
        byw={}
        for r in m: byw.setdefault(int(r['fan_pos']),[]).append(r)
        want={}
        for fp in sorted(byw):
            cands=sorted(byw[fp], key=lambda r:(r['distance'], r['script_pos']))[:10]
            w=int(cands[0]['work']); loc=fp-int(tokoff[w])
            for r in cands:
                sp=int(r['script_pos'])
                ctx="["+", ".join(works[w][loc:loc+6])+"]"; ms=" ".join(script[sp:sp+6])
                lev=ora.levenshtein(ms,ctx); comb=float(r['distance'])*lev
                for k in range(6):
                    key=(w,loc+k)
                    if key not in want or comb<want[key][0]: want[key]=(comb,float(r['distance']),k,sp,lev)
        got={(int(a),int(b)):(float(d)*int(l),float(d),int(k),int(mi),int(l)) for a,b,k,mi,d,l in zip(best['work'],best['word'],best['window_ix'],best['match_ix'],best['distance'],best['lev'])}
        assert list(got)==sorted(got), "order"
        assert got==want, (trial, len(got), len(want))


def _raw_and_pretokenised(golden_dir, tmp_path):
    """A fanwork written as raw prose (punctuation glued to the words, a contraction, a newline)
    and the same token stream written pre-tokenised (single spaces)."""
    from fandom_search_b200 import text
    script_words = [r[0] for r in search.load_markup_script(os.path.join(golden_dir, "script.txt"))[1:]]
    q = script_words[20:34]
    raw = ('Before, "%s %s %s %s %s %s %s!" she said; it isn\'t\n(%s %s %s %s %s %s %s)... After--end.'
           % tuple(q[:14]))
    raw_path = tmp_path / "raw.txt"
    raw_path.write_text(raw, encoding="utf-8")
    toks = text.tokenize_rules(raw)
    pre_path = tmp_path / "pre.txt"
    pre_path.write_text(" ".join(toks), encoding="utf-8")
    return str(raw_path), str(pre_path), toks


def test_custom_tokenizer_with_punctuation(monkeypatch, golden_dir, tmp_path):
    """tokenizer= (here the built-in rule tokeniser; in production spaCy's): the file goes through
    sp_parse_chunks and the is_space filter like in the reference (search.py:164-166), punctuation
    becomes tokens of its own, and the records equal those of the same token stream pre-tokenised."""
    from fandom_search_b200 import text
    monkeypatch.setattr(engine_mod, "DeviceIndex", NumpyIndex)
    lex = Lexicon.from_npz(os.path.join(golden_dir, "lexicon.npz"), hash_fn=py_hash_seed0)
    script = os.path.join(golden_dir, "script.txt")
    try:
        search.set_pipeline(search.Pipeline(lex))
        raw_path, pre_path, toks = _raw_and_pretokenised(golden_dir, tmp_path)
        assert "," in toks and "n't" in toks and '"' in toks and "..." in toks and "--" in toks
        want = search.AnnIndexSearch(script, 6, 15, 14, 0.1).search(pre_path)
        # a tokeniser that also yields whitespace tokens (as spaCy does for newlines): they are dropped
        noisy = lambda s: [t for w in text.tokenize_rules(s) for t in (w, " ")] + ["\n"]
        for tok in (text.tokenize_rules, "rules", noisy):
            search.set_pipeline(search.Pipeline(lex, tokenizer=tok))
            got = search.AnnIndexSearch(script, 6, 15, 14, 0.1).search(raw_path)
            assert len(got) == len(want) > 0
            for g, w in zip(normalise(got), normalise(want)):
                assert g[1:] == w[1:]                          # everything but the file name
        # the default whitespace tokeniser on the raw prose: punctuation stays glued, fewer matches,
        # and the pipeline says so
        search.set_pipeline(search.Pipeline(lex))
        with pytest.warns(RuntimeWarning, match="punctuation"):
            glued = search.AnnIndexSearch(script, 6, 15, 14, 0.1).search(raw_path)
        assert len(glued) < len(want)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("error")                     # pre-tokenised text: no warning
            search.set_pipeline(search.Pipeline(lex))
            search.AnnIndexSearch(script, 6, 15, 14, 0.1).search(pre_path)
    finally:
        search.set_pipeline(None)


def test_rule_tokeniser_examples():
    from fandom_search_b200 import text
    assert text.tokenize_rules('Hello, world!') == ['Hello', ',', 'world', '!']
    assert text.tokenize_rules("I don't know; it's \"fine\"...") == \
        ['I', 'do', "n't", 'know', ';', 'it', "'s", '"', 'fine', '"', '...']
    assert text.tokenize_rules("The U.S. and Mr. Smith (really) can't go--now.") == \
        ['The', 'U.S.', 'and', 'Mr.', 'Smith', '(', 'really', ')', 'ca', "n't", 'go', '--', 'now', '.']
    assert text.tokenize_rules("  spaced\tout\n\ntext  ") == ['spaced', 'out', 'text']
    assert text.tokenize_rules("") == [] and text.tokenize_rules("...") == ['...']
    assert text.glued_punctuation_share('Hello, world! said "Bob" and left.'.split()) > 0.5
    assert text.glued_punctuation_share('w00012 w00013 the of'.split()) == 0.0


def test_fan_side_oov_words_are_not_registered(cpu_device, golden_dir):
    """Only the script's out-of-vocabulary words live in the lexicon's registry: searching works full
    of unknown words must not grow it (a run over millions of fanworks sees millions of them)."""
    idx = search.AnnIndexSearch(os.path.join(golden_dir, "script.txt"), 6, 15, 14, 0.1)
    lex = search.get_spacy_model().lexicon
    n_script_oov = lex.n_oov
    assert n_script_oov > 0
    files = sorted(glob.glob(os.path.join(golden_dir, "fanworks", "*.txt")))
    first = normalise([r for s in idx.search_many(files) for r in s])
    assert lex.n_oov == n_script_oov
    again = normalise([r for s in idx.search_many(files[::-1]) for r in s])      # other batch order, same rows
    assert sorted(map(tuple, again)) == sorted(map(tuple, first)) and lex.n_oov == n_script_oov
    ids, extra = lex.batch_oov(["zzz-unknown", "zzz-unknown-2"], lex.n_rows + n_script_oov)
    assert ids.tolist() == [lex.n_rows + n_script_oov, lex.n_rows + n_script_oov + 1] and extra.shape == (2, lex.dim)
    assert 1 <= extra[0].sum() <= 3 and lex.n_oov == n_script_oov
