"""Drop-in replacement for the reference's `search.py` call surface, on the B200 path.

`ao3.py` reaches this module at `search.analyze` (ao3.py:519), `search.validate_cmd`
(ao3.py:511) and `search.load_markup_script` (ao3.py:365,434); names, argument meaning, file
naming and the CSV schema follow /root/reference/search.py.  What changes is underneath:

  reference (search.py:381-386)                   here
  ------------------------------------------      -------------------------------------------
  Pool(4).map(multi_search_wrapper, cluster)  ->  AnnIndexSearch.search_many(cluster):
    per work: spaCy tokens -> float64 windows       tokens -> CSR row ids -> ONE C-ABI call
    per window: nearpy LSH neighbours()             (gather + tcgen05 window contraction +
    threshold, Levenshtein, explode, dedup          float64 rescoring on the GPU), then the
                                                    same Levenshtein/explode/dedup on the few
                                                    surviving pairs

The default search is exhaustive (a superset of what the reference's LSH finds; identical for
identical windows).  FANDOM_SEARCH_MODE=lsh reproduces a seeded random-hyperplane index
exactly (FANDOM_SEARCH_LSH_SEED), see `lsh.py`.  There is no CPU fallback: without the CUDA
library or a B200 the search raises.

Environment knobs (the CLI itself is unchanged):
  FANDOM_SEARCH_LEXICON   path of the lexicon .npz (stand-in for spaCy's en_core_web_md table)
  FANDOM_SEARCH_MODE      exhaustive (default) | lsh
  FANDOM_SEARCH_LSH_SEED  int seed of the emulated hyperplanes (mode lsh)
  FANDOM_SEARCH_DEVICE    CUDA device ordinal (default: LOCAL_RANK or 0)
  FANDOM_SEARCH_OOV_HASH  python (default: builtin hash, as the reference) | seed0
"""
import csv
import datetime
import io
import os
import random
import re
import sys

import numpy

from . import _native as nt
from . import text as _text
from .lexicon import Lexicon, py_hash_seed0

_SPACY_MODEL = None
_ANN_INDEX = None
# wall-clock marks of the last analyze() of this process (time.perf_counter): start, index_ready,
# collected[(cluster, t, windows so far)], searched, end -- read by bench.py / tools/pipeline_bench.py
ANALYZE_STATS = {}

# search.py:20-37 -- the CSV schema is the drop-in contract
new_record_structure = {
    'fields': ['FAN_WORK_FILENAME',
               'FAN_WORK_WORD_INDEX',
               'FAN_WORK_WORD',
               'FAN_WORK_ORTH_ID',
               'ORIGINAL_SCRIPT_WORD_INDEX',
               'ORIGINAL_SCRIPT_WORD',
               'ORIGINAL_SCRIPT_ORTH_ID',
               'ORIGINAL_SCRIPT_CHARACTER',
               'ORIGINAL_SCRIPT_SCENE',
               'BEST_MATCH_DISTANCE',
               'BEST_LEVENSHTEIN_DISTANCE',
               'BEST_COMBINED_DISTANCE'],
    'types': [str, int, str, int, int, str, int, str, int, float, int, float],
}


class Token(object):
    """What the hot path needs of a spaCy Token (search.py:166,194-195,327)."""
    __slots__ = ('text',)

    def __init__(self, text):
        self.text = text

    @property
    def is_space(self):
        # spaCy's Token.is_space: the token consists of whitespace characters only
        return self.text.isspace()

    @property
    def orth_(self):
        return self.text

    @property
    def orth(self):
        return _text.string_id(self.text)

    @property
    def lower_(self):
        return self.text.lower()

    @property
    def lower(self):
        return _text.string_id(self.text.lower())

    def __str__(self):
        return self.text

    __repr__ = __str__


class Pipeline(object):
    """Tokeniser + lexicon; stands where the reference holds the spaCy model."""

    def __init__(self, lexicon, tokenizer=None):
        self.lexicon = lexicon
        if tokenizer == 'rules':
            tokenizer = _text.tokenize_rules
        self.tokenizer = tokenizer or _text.tokenize
        self._warned_glued = False

    def check_tokens(self, words, where):
        """The default tokeniser splits on whitespace only (what pre-tokenised / cleaned corpora
        need).  Raw prose then keeps its punctuation glued to the words ('Hello,' / "don't"), those
        tokens miss the lexicon, become 3-hot OOV vectors, and the matches silently diverge from a
        reference run with spaCy: say so, once, loudly."""
        if self._warned_glued or self.tokenizer is not _text.tokenize:
            return
        share = _text.glued_punctuation_share(words)
        if share > 0.02:
            self._warned_glued = True
            import warnings
            warnings.warn(
                "%s: %.0f %% of the tokens carry leading/trailing punctuation -- this looks like raw prose, "
                "and the default tokeniser splits on whitespace only (no spaCy here). Results will differ from "
                "the reference's spaCy tokenisation; pass a spaCy-compatible `tokenizer=` to search.Pipeline "
                "(or FANDOM_SEARCH_TOKENIZER=rules for the built-in approximation)." % (where, 100 * share),
                RuntimeWarning, stacklevel=3)

    def __call__(self, text):
        return [Token(w) for w in self.tokenizer(text)]


def set_pipeline(pipeline):
    """Install the tokeniser/lexicon explicitly (instead of FANDOM_SEARCH_LEXICON)."""
    global _SPACY_MODEL
    _SPACY_MODEL = pipeline


def get_spacy_model():
    # search.py:40-45 (lazy module-level singleton)
    global _SPACY_MODEL
    if _SPACY_MODEL is None:
        path = os.environ.get('FANDOM_SEARCH_LEXICON')
        if not path:
            raise RuntimeError(
                "no lexicon configured: set FANDOM_SEARCH_LEXICON to a lexicon .npz exported "
                "from the spaCy vectors table (Lexicon.from_spacy) or call search.set_pipeline()")
        _SPACY_MODEL = Pipeline(Lexicon.from_npz(path, hash_fn=_default_oov_hash()),
                                tokenizer=os.environ.get('FANDOM_SEARCH_TOKENIZER') or None)
    return _SPACY_MODEL


def _default_oov_hash():
    """The hash of the out-of-vocabulary rule (search.py:79-83).  The reference uses the builtin str
    hash, randomised per INTERPRETER: its four workers are forked from one process and share it.
    Ranks under torchrun are separate interpreters -- with the builtin hash every rank would build
    different OOV pseudo-vectors and an N-GPU run would differ from a 1-GPU run.  So: seed0 (the
    PYTHONHASHSEED=0 hash, identical everywhere) when asked for, or when WORLD_SIZE > 1 and the
    environment does not pin PYTHONHASHSEED itself; the builtin hash otherwise."""
    choice = os.environ.get('FANDOM_SEARCH_OOV_HASH')
    if choice == 'seed0':
        return py_hash_seed0
    if choice in (None, ''):
        world = int(os.environ.get('WORLD_SIZE', '1') or 1)
        pinned = os.environ.get('PYTHONHASHSEED', '') not in ('', 'random')
        if world > 1 and not pinned:
            return py_hash_seed0
    return None


def sp_parse_chunks(txt, size=100000):
    # search.py:47-63: texts of 100000+ characters are cut at spaces into <=100k pieces.
    # (`size` is ignored there too.)  With a whitespace tokeniser the pieces tokenise to the
    # same stream as the whole text; kept for call-surface parity.
    model = get_spacy_model()
    if len(txt) < 100000:
        yield model(txt)
        return
    start = 0
    while start < len(txt):
        end = start + 100000
        if end > len(txt):
            end = len(txt)
        else:
            while txt[end] != ' ':   # IndexError at end == len(txt), as in the reference
                end -= 1
        yield model(txt[start:end])
        start = end + 1


def mk_vectors(sp_txt):
    # search.py:65-84, host restatement for API parity (the search itself gathers on the GPU)
    lex = get_spacy_model().lexicon
    rows = len(sp_txt)
    cols = lex.dim if rows else 0
    vectors = numpy.empty((rows, cols), dtype=float)
    for i, word in enumerate(sp_txt):
        vectors[i] = lex.vector(str(word))
    return vectors


class _PinnedTokens(object):
    """A few reusable page-locked int32 buffers for the CSR token ids of a cluster: the C-ABI call
    copies them host -> device, and from pageable memory that copy (10 MB per cluster) costs ~1.5 ms
    of the 22 ms a cluster takes; from page-locked memory it is a plain DMA.  Without a CUDA device
    (the CPU tests) `take` returns an ordinary numpy copy."""

    def __init__(self, keep=4):
        self._free = []
        self._keep = keep
        self._lock = __import__('threading').Lock()

    def take(self, src):
        try:
            import torch
            if not torch.cuda.is_available():
                raise RuntimeError
        except Exception:
            return numpy.array(src, dtype=numpy.int32), None
        n = len(src)
        with self._lock:
            owner = None
            for k, t in enumerate(self._free):
                if t.numel() >= n:
                    owner = self._free.pop(k)
                    break
        if owner is None:
            owner = torch.empty(max(n + n // 8, 1 << 20), dtype=torch.int32, pin_memory=True)
        view = owner.numpy()[:n]
        numpy.copyto(view, src)
        return view, owner

    def give_back(self, owner):
        if owner is None:
            return
        with self._lock:
            if len(self._free) < self._keep:
                self._free.append(owner)


_PINNED_TOKENS = _PinnedTokens()


def _device_ordinal():
    from .parallel import device_ordinal
    return device_ordinal()


class DeviceRows(object):
    """Winning records of one cluster as the device returned them (fs_row array, sorted by
    (work, word)), presented as the arrays fs_records_best yields."""

    def __init__(self, rows):
        self.rows = rows
        self.best = {k: numpy.ascontiguousarray(rows[k]) for k in
                     ('work', 'word', 'window_ix', 'match_ix', 'distance', 'lev')}

    def __len__(self):
        return len(self.rows)


class ScriptEngine(object):
    """Returned by build_lsh_engine: the device-resident script index (replaces the nearpy
    Engine of search.py:118-123).  The per-window `neighbours(v)` seam of nearpy is replaced
    by the batched `DeviceIndex.search_host`."""

    def __init__(self, device_index, lsh=None):
        self.index = device_index
        self.lsh = lsh

    def neighbours(self, v):
        raise NotImplementedError(
            "the per-window nearpy seam (search.py:178) is replaced by the batched GPU search; "
            "use AnnIndexSearch.search / search_many")

    def store_vector(self, v, data=None):
        raise NotImplementedError("the script index is immutable once built")


def build_lsh_engine(orig, window_size, number_of_hashes, hash_dimensions,
                     distance_threshold=0.1, script_offsets=None):
    # search.py:86-124: script tokens -> device index.  `orig` is the sequence of script
    # tokens (lower-cased words); `script_offsets` (optional) are the CSR boundaries when several
    # scripts are indexed together -- windows never straddle a boundary.
    from .engine import DeviceIndex
    lex = get_spacy_model().lexicon
    words = [str(t) for t in orig]
    script_tok = lex.row_ids(words)
    n_sx = lex.n_oov
    index = DeviceIndex(lex.table, script_tok, script_off=script_offsets, extra=lex.oov_rows(0, n_sx),
                        window=window_size, threshold=distance_threshold, device=_device_ordinal())
    lsh = None
    if os.environ.get('FANDOM_SEARCH_MODE', 'exhaustive') == 'lsh':
        from .lsh import LshEmulation
        seed = int(os.environ.get('FANDOM_SEARCH_LSH_SEED', '0'))
        lsh = LshEmulation(number_of_hashes, hash_dimensions, window_size * lex.dim, seed)
        lsh.install(index)
    engine = ScriptEngine(index, lsh)
    engine.script_tok = script_tok
    engine.n_script_extra = n_sx
    # single script: the records (top-10, Levenshtein, per-word argmin; search.py:182-226) are made on
    # the device right behind the search -- FANDOM_SEARCH_DEVICE_RECORDS=0 keeps them on the host
    single = script_offsets is None or len(script_offsets) == 2
    if (single and hasattr(index, 'set_script_text') and len(words) > 0
            and os.environ.get('FANDOM_SEARCH_DEVICE_RECORDS', '1') != '0'):
        enc = [w.encode('utf-8') for w in words]
        off = numpy.zeros(len(enc) + 1, dtype=numpy.int64)
        numpy.cumsum([len(e) for e in enc], out=off[1:])
        index.set_script_text(b''.join(enc), off)
    return engine


def multi_search_wrapper(work):
    # search.py:126-128
    result = _ANN_INDEX.search(work)
    return result


class AnnIndexSearch(object):
    def __init__(self, original_script_filename, window_size,
                 number_of_hashes, hash_dimensions, distance_threshold):
        # search.py:131-154.  `original_script_filename` may also be a LIST of markup scripts
        # (SURVEY 8f row N4): they are indexed side by side (one wider script matrix, windows
        # never straddle scripts) and the corpus is searched against all of them in one pass;
        # see search_many_scripts / analyze_scripts.
        multi = not isinstance(original_script_filename, (str, bytes, os.PathLike))
        self.script_filenames = list(original_script_filename) if multi else [original_script_filename]
        rows, offsets = [], [0]
        for fn in self.script_filenames:
            rows.extend(load_markup_script(fn)[1:])
            offsets.append(len(rows))
        if rows:
            (self.word_lowercase, self.orth_id, self.scene, self.character) = zip(*rows)
        else:
            self.word_lowercase = self.orth_id = self.scene = self.character = ()
        self.script_offsets = numpy.array(offsets, dtype=numpy.int64)
        self.word_index = tuple(range(len(self.word_lowercase)))
        self.window_size = window_size
        self.distance_threshold = distance_threshold
        self.spacy_model = get_spacy_model()
        self.engine = build_lsh_engine(self.word_lowercase, window_size, number_of_hashes,
                                       hash_dimensions, distance_threshold,
                                       script_offsets=self.script_offsets)
        self.reset_stats()

    def reset_stats(self):
        self._windows_processed = 0

    @property
    def windows_processed(self):
        return self._windows_processed

    # -- host side of one batch --------------------------------------------
    def _tokenize_file(self, filename):
        # search.py:164-166: the text goes through sp_parse_chunks (>= 100k characters: pieces cut at
        # spaces, search.py:47-63) and the is_space tokens are dropped, exactly as the reference does
        # with whatever tokeniser the pipeline holds.  Returns the token TEXTS.
        with open(filename, encoding='utf8') as fan_file:
            fan = fan_file.read()
        return [str(t) for ch in sp_parse_chunks(fan) for t in ch if not t.is_space]

    def search(self, filename):
        # search.py:163-226 for one work
        return self.search_many([filename])[0]

    def search_many(self, filenames):
        """Record lists (one per file, each sorted) for a cluster of works."""
        prep = self.prepare(filenames)
        if len(self.script_filenames) != 1:
            return self.run_prepared(prep)
        return self.records_prepared(prep, *self.collect_prepared(prep, self.submit_prepared(prep, rows=True)))

    def prepare(self, filenames):
        """Host stage before the GPU: read + tokenise + row ids -> CSR batch (search.py:164-169).
        Native and multi-threaded for the default whitespace tokeniser; safe to run in a
        background thread while the previous cluster is being searched."""
        lex = self.spacy_model.lexicon
        n_fixed = lex.n_rows + self.engine.n_script_extra
        if self.spacy_model.tokenizer is _text.tokenize:
            if getattr(self.spacy_model, '_vocab', None) is None:
                self.spacy_model._vocab = _text.Vocab(lex)
            batch = self.spacy_model._vocab.encode_files(filenames)
            if not self.spacy_model._warned_glued and len(batch.tok):
                n = min(int(batch.tok_off[-1]), 2000)
                self.spacy_model.check_tokens([batch.token_text(i) for i in range(n)], filenames[0])
            tok, pin = _PINNED_TOKENS.take(batch.tok)              # private, page-locked copy
            offs = numpy.array(batch.tok_off, dtype=numpy.int64)
            extra = None
            if len(batch.oov_start):
                # only the OOV positions are touched (a few % of the tokens).  The batch's unique OOV
                # strings get the script's registered id when the script holds the same word, else a
                # batch-local row -- nothing is added to the lexicon's registry (lexicon.batch_oov)
                oov_ids, extra = lex.batch_oov(batch.oov_strings(), n_fixed)
                pos = numpy.flatnonzero(tok < 0)
                tok[pos] = oov_ids[-tok[pos] - 1]
            return {'filenames': list(filenames), 'batch': batch, 'tok': tok, 'offs': offs, 'extra': extra,
                    'pin': pin}
        fans = [self._tokenize_file(fn) for fn in filenames]
        batch = _text.Batch.from_token_lists(fans)
        if fans and not self.spacy_model._warned_glued:
            self.spacy_model.check_tokens(fans[0][:2000], filenames[0])
        words = [w for f in fans for w in f]
        get = lex.key_to_row.get
        tok = numpy.array([get(w, -1) for w in words], dtype=numpy.int32).reshape(-1)
        offs = numpy.array(batch.tok_off, dtype=numpy.int64)
        # out-of-vocabulary words: the script's registered id or a batch-local row, never registered
        extra = None
        pos = numpy.flatnonzero(tok < 0)
        if len(pos):
            uniq = {}
            for i in pos.tolist():
                uniq.setdefault(words[i], len(uniq))
            oov_ids, extra = lex.batch_oov(list(uniq), n_fixed)
            tok[pos] = oov_ids[[uniq[words[i]] for i in pos.tolist()]]
        return {'filenames': list(filenames), 'batch': batch, 'tok': tok, 'offs': offs, 'extra': extra}

    def run_prepared(self, prep):
        return self.records_prepared(prep, *self.search_prepared(prep))

    def search_prepared(self, prep):
        """GPU stage: one C-ABI call for the cluster (search.py:169-184 for every work)."""
        return self.collect_prepared(prep, self.submit_prepared(prep))

    def submit_prepared(self, prep, rows=False):
        """Enqueue the cluster on the GPU and return at once (fs_search_submit; two clusters may be
        in flight, so the next one is already queued while this one is searched).  rows=True: the
        records are made on the device too (fs_search_submit_rows) when the index supports it."""
        index = self.engine.index
        batch = prep['batch']
        if (rows and getattr(index, 'device_records', False) and getattr(batch, 'tok_start32', None) is not None
                and len(batch.text) < (1 << 32)):
            return index.search_submit_rows(prep['tok'], prep['offs'], prep['extra'], batch.text,
                                            batch.tok_start32, batch.tok_len16,
                                            lsh_filter=self.engine.lsh is not None)
        return index.search_submit(prep['tok'], prep['offs'], prep['extra'])

    def collect_prepared(self, prep, ticket):
        """Wait for a submitted cluster: (matches, first LSH table or None) -- or, for a cluster
        submitted with rows=True, (DeviceRows, None): the winning rows, made on the device."""
        index = self.engine.index
        if 'cap_rows' in ticket:
            rows, counters = index.search_collect_rows(ticket)
            if rows is not None:
                ticket.clear()
                if prep.get('pin') is not None:
                    _PINNED_TOKENS.give_back(prep.pop('pin'))
                    prep['tok'] = None
                self._windows_processed += int(counters[nt.FS_CNT_WINDOWS])
                return DeviceRows(rows), None
            # the device could not finish the records of this cluster: raw matches, host records
        matches, counters = index.search_collect(ticket)
        ticket.clear()
        if prep.get('pin') is not None:
            # the token ids are on the device now: the page-locked buffer goes back to the pool
            _PINNED_TOKENS.give_back(prep.pop('pin'))
            prep['tok'] = None
        self._windows_processed += int(counters[nt.FS_CNT_WINDOWS])
        first_table = None
        if self.engine.lsh is not None:
            # keep only the pairs the emulated LSH index would have compared (lsh.py)
            first_table = ((matches['flags'] >> nt.FS_MATCH_LSH_SHIFT) & 0xFF).astype(numpy.int32)
            keep = first_table > 0
            matches, first_table = matches[keep], first_table[keep]
        return matches, first_table

    def records_prepared(self, prep, matches, first_table=None):
        """Host stage after the GPU (search.py:188-226)."""
        if len(self.script_filenames) != 1:
            raise ValueError("several scripts are indexed: use records_prepared_scripts")
        return self._records(prep['filenames'], prep['batch'], matches, first_table)

    def _best(self, batch, matches, first_table):
        """Winning records as arrays: straight from the device (DeviceRows) or made natively on the
        host from the match list (fs_records_best_mt)."""
        if isinstance(matches, DeviceRows):
            return matches.best
        blob, soff = self._script_text()
        return _text.records_best(matches, first_table, self.window_size, 10, batch, blob, soff)

    def records_prepared_scripts(self, prep, matches, first_table=None):
        """One record-set list per indexed script: what separate reference runs, one per script,
        would each have produced for these works (top-10, argmin and word indices are per script)."""
        sid = numpy.searchsorted(self.script_offsets, matches['script_pos'], side='right') - 1
        out = []
        for k in range(len(self.script_filenames)):
            sel = sid == k
            out.append(self._records(prep['filenames'], prep['batch'], matches[sel],
                                     None if first_table is None else first_table[sel],
                                     word_base=int(self.script_offsets[k])))
        return out

    def search_many_scripts(self, filenames):
        prep = self.prepare(filenames)
        return self.records_prepared_scripts(prep, *self.search_prepared(prep))

    def _script_text(self):
        if getattr(self, '_script_blob', None) is None:
            enc = [w.encode('utf-8') for w in self.word_lowercase]
            off = numpy.zeros(len(enc) + 1, dtype=numpy.int64)
            numpy.cumsum([len(e) for e in enc], out=off[1:])
            self._script_blob = b''.join(enc)
            self._script_off = off
        return self._script_blob, self._script_off

    def _script_columns(self):
        # script-side CSV columns as flat arrays for fs_records_format_csv
        if getattr(self, '_script_cols', None) is None:
            n = len(self.word_lowercase)
            orth = numpy.array(self.orth_id, dtype=numpy.uint64) if n else numpy.zeros(0, numpy.uint64)
            enc = [b'' if c is None else str(c).encode('utf-8') for c in self.character]
            coff = numpy.zeros(n + 1, dtype=numpy.int64)
            numpy.cumsum([len(e) for e in enc], out=coff[1:])
            cnone = numpy.array([c is None for c in self.character], dtype=numpy.uint8)
            snone = numpy.array([sc is None for sc in self.scene], dtype=numpy.uint8)
            scene = numpy.array([0 if sc is None else sc for sc in self.scene], dtype=numpy.int64)
            self._script_cols = (orth, b''.join(enc), coff, cnone, scene, snone)
        return self._script_cols

    def records_text_prepared(self, prep, matches, first_table=None, word_base=0, on_best=None):
        """CSV text (utf-8 bytes) of the cluster's records: byte for byte what
        write_records(records_prepared(...)) puts in the batch file, formatted natively
        (fs_records_format_csv) without building per-row Python objects."""
        filenames, batch = prep['filenames'], prep['batch']
        if len(matches) == 0:
            return b''
        blob, soff = self._script_text()
        best = self._best(batch, matches, first_table)
        if on_best is not None:
            on_best(matches, best)
        orth, cblob, coff, cnone, scene, snone = self._script_columns()
        return _text.records_format_csv(best, filenames, batch, blob, soff, orth, cblob, coff, cnone,
                                        scene, snone, word_base)

    def _records(self, filenames, batch, matches, first_table=None, word_base=0):
        # search.py:182-226 on the surviving pairs: top-10 per window, Levenshtein, six records
        # per pair, per-word argmin (native, fs_records_best); here only the row formatting.
        out = [[] for _ in filenames]
        if len(matches) == 0:
            return out
        best = self._best(batch, matches, first_table)
        tok_off = batch.tok_off
        work = best['work'].tolist()
        word = best['word'].tolist()
        window_ix = best['window_ix'].tolist()
        match_ix = best['match_ix'].tolist()
        distance = best['distance'].tolist()
        lev = best['lev'].tolist()
        for i in range(len(work)):
            wi = work[i]
            fan_word = batch.token_text(int(tok_off[wi]) + word[i])
            orig_word_ix = match_ix[i] + window_ix[i]
            out[wi].append([filenames[wi],
                            word[i],
                            fan_word,                        # orth_
                            _text.string_id(fan_word),       # orth
                            orig_word_ix - word_base,        # index inside its own script
                            self.word_lowercase[orig_word_ix],
                            self.orth_id[orig_word_ix],
                            self.character[orig_word_ix],
                            self.scene[orig_word_ix],
                            distance[i],
                            lev[i],
                            distance[i] * lev[i]])
        return out


def validate_markup_script(filename, interactive=False,
                           _unbalanced_l=re.compile('<<[^>]*<<'),
                           _unbalanced_r=re.compile('>>[^<]*>>'),
                           _tags=re.compile(r'>>\s*([^<]*)\s*<<')):
    # search.py:228-285 (not on the hot path; kept so `ao3.py validate` keeps working)
    with open(filename, encoding='utf-8') as ip:
        script = ip.read()
    print('Checking script for markup errors.')
    print()
    problems = False

    def report(title, rex, group=0, keep=lambda s: True):
        found = False
        for m in rex.finditer(script):
            shown = m.group(group).strip()
            if not keep(shown):
                continue
            if not found:
                print(title)
                found = True
            print('  On line {}'.format(script[:m.start(group) + 1].count('\n') + 1))
            print('    {}'.format(shown))
        if found:
            print()
        return found

    problems |= report('Unbalanced left tag delimiters:', _unbalanced_l)
    problems |= report('Unbalanced right tag delimiters:', _unbalanced_r)
    expected = {'LINE', 'DIRECTION', 'SCENE_NUMBER', 'SCENE_DESCRIPTION', 'CHARACTER_NAME'}
    problems |= report('Unexpected tag labels:', _tags, 1, lambda s: s not in expected)
    if not problems:
        print('No markup errors found.')
        return True
    if interactive:
        print('Errors were found in the script markup. Do you want to continue? (Default is no.)')
        print()
        r = ''
        while r.lower() not in ('y', 'yes', 'n', 'no'):
            r = input('Enter y for yes or n for no: ')
            if not r.strip():
                r = 'n'
        return r.lower() in ('y', 'yes')
    return False


def validate_cmd(args):
    return validate_markup_script(args.script)


def load_markup_script(filename,
                       _line_rex=re.compile('LINE<<(?P<line>[^>]*)>>'),
                       _scene_rex=re.compile('SCENE_NUMBER<<(?P<scene>[^>]*)>>'),
                       _char_rex=re.compile('CHARACTER_NAME<<(?P<character>[^>]*)>>')):
    # search.py:290-329: one row [lower_, lower id, scene, character] per LINE token
    model = get_spacy_model()
    rows = [['LOWERCASE', 'SPACY_ORTH_ID', 'SCENE', 'CHARACTER']]
    scene = None
    scenes_seen = 0
    scene_fallback = False     # sticky once a scene number fails to parse (search.py:313-317)
    character = None
    with open(filename, encoding='utf-8') as ip:
        for line in ip:
            m = _scene_rex.search(line)
            if m:
                scenes_seen += 1
                digits = ''.join(c for c in m.group('scene') if c.isdigit())
                try:
                    scene = int(digits)
                except ValueError:
                    scene_fallback = True
                    print("Error in Scene markup: {}".format(line))
                if scene_fallback:
                    scene = scenes_seen
                continue
            m = _char_rex.search(line)
            if m:
                character = m.group('character')
                continue
            m = _line_rex.search(line)
            if m:
                for t in model(m.group('line')):
                    if not t.is_space:
                        rows.append([t.lower_, t.lower, scene, character])
    return rows


def write_records(records, filename):
    # search.py:331-334 (csv default dialect: \r\n, minimal quoting, None -> empty)
    with open(filename, 'w', encoding='utf-8') as out:
        wr = csv.writer(out)
        wr.writerows(records)


def format_records(records):
    """The exact text write_records would put in the file (rows end in \\r\\n)."""
    buf = io.StringIO(newline='')
    csv.writer(buf).writerows(records)
    return buf.getvalue()


def _write_text(text, filename):
    if isinstance(text, str):
        text = text.encode('utf-8')
    with open(filename, 'wb') as out:
        out.write(text)


def analyze_scripts(args, scripts, window_size=6, number_of_hashes=15, hash_dimensions=14,
                    distance_threshold=0.1, chunk_size=500):
    """One pass over the fanwork folder for SEVERAL markup scripts (the
    reference runs `ao3.py search` once per film, workflow/*.py).  Writes, per script, the files
    a separate run would have written, with the script's file stem in the name:
    match-{w}gram-{stem}-batch-{i}.csv and match-{w}gram-{stem}-YYYYMMDD[-k].csv."""
    fan_work_directory = args.fan_works
    subsample_start = 0 if args.skip_works < 0 else args.skip_works
    subsample_end = None if args.num_works < 0 else args.num_works + subsample_start
    fan_works = [os.path.join(fan_work_directory, f) for f in os.listdir(fan_work_directory)]
    random.seed(4815162342)
    random.shuffle(fan_works)
    fan_works = fan_works[subsample_start:subsample_end]
    fan_clusters = [fan_works[i:i + chunk_size] for i in range(0, len(fan_works), chunk_size)]
    stems = [os.path.splitext(os.path.basename(sc))[0] for sc in scripts]
    if len(set(stems)) != len(stems):
        raise ValueError("script file stems must be distinct: %s" % stems)
    rank, world = _dist_env()
    ann_index = AnnIndexSearch(list(scripts), window_size, number_of_hashes, hash_dimensions,
                               distance_threshold)
    per_script = [dict() for _ in scripts]
    for i, fan_cluster in enumerate(fan_clusters):
        if i % world != rank:
            continue
        print('Processing cluster {} ({}-{})'.format(i, chunk_size * i, chunk_size * (i + 1)))
        for k, record_sets in enumerate(ann_index.search_many_scripts(fan_cluster)):
            records = [r for r_set in record_sets for r in r_set]
            write_records(records, 'match-{}gram-{}-batch-{}.csv'.format(window_size, stems[k], i))
            per_script[k][i] = records
    if world > 1:
        from .parallel import gather_cluster_records
        per_script = [gather_cluster_records(d, rank, world) for d in per_script]
        if rank != 0:
            return
    for k, stem in enumerate(stems):
        accumulated = [new_record_structure['fields']]
        for i in sorted(per_script[k]):
            accumulated.extend(per_script[k][i])
        n = 0
        name = 'match-{}gram-{}-{:%Y%m%d}.csv'.format(window_size, stem, datetime.date.today())
        while os.path.exists(name):
            n += 1
            name = 'match-{}gram-{}-{:%Y%m%d}-{}.csv'.format(window_size, stem, datetime.date.today(), n)
        write_records(accumulated, name)


def _file_size(path):
    try:
        return os.path.getsize(path)
    except OSError:
        return 0


def _dist_env():
    world = int(os.environ.get('WORLD_SIZE', '1') or 1)
    rank = int(os.environ.get('RANK', '0') or 0)
    return rank, world


def _concatenate_files(target, header, parts, threads=4):
    """target = header + the bytes of every file of `parts`, in order (search.py:367, 388: the dated
    aggregate is the header row and the rows of every cluster).  The pieces are copied inside the kernel
    (os.copy_file_range: page cache to page cache, no pass through Python buffers) by a few threads, each
    into its own byte range of the pre-sized target: the aggregate of a million works is gigabytes, and
    rank 0 writes it while every other rank waits."""
    sizes = [os.stat(p).st_size for p in parts]            # FileNotFoundError names the missing piece
    offsets = [len(header)]
    for n in sizes:
        offsets.append(offsets[-1] + n)
    with open(target, 'wb') as out:
        out.write(header)
        out.truncate(offsets[-1])
        out.flush()
        dst = out.fileno()

        def copy(k):
            with open(parts[k], 'rb') as src:
                done, fd = 0, src.fileno()
                while done < sizes[k]:
                    n = 0
                    if hasattr(os, 'copy_file_range'):
                        try:
                            n = os.copy_file_range(fd, dst, sizes[k] - done, done, offsets[k] + done)
                        except OSError:                     # another file system, an old kernel: plain copy
                            n = 0
                    if n == 0:
                        buf = os.pread(fd, min(sizes[k] - done, 1 << 22), done)
                        if not buf:
                            raise IOError("%s shrank while it was being copied" % parts[k])
                        n = os.pwrite(dst, buf, offsets[k] + done)
                    done += n

        if threads > 1 and len(parts) > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=threads) as pool:
                list(pool.map(copy, range(len(parts))))
        else:
            for k in range(len(parts)):
                copy(k)


def analyze(args,
            window_size=6,
            number_of_hashes=15,
            hash_dimensions=14,
            distance_threshold=0.1,
            chunk_size=500,
            reuse_histogram=None):
    # search.py:336-399.  Listing, seeded shuffle, sub-sampling, clustering and file names are
    # the reference's; the per-cluster pool.map becomes one batched GPU search.  Under
    # torchrun (WORLD_SIZE > 1) cluster i is searched by rank i % WORLD_SIZE on its own GPU,
    # each rank writes its own batch files and rank 0 writes the aggregate in cluster order.
    import time
    stats = {'start': time.perf_counter(), 'collected': []}
    ANALYZE_STATS.clear()
    ANALYZE_STATS.update(stats)
    fan_work_directory = args.fan_works
    original_script_markup = args.script
    subsample_start = 0 if args.skip_works < 0 else args.skip_works
    subsample_end = None if args.num_works < 0 else args.num_works + subsample_start

    fan_works = [os.path.join(fan_work_directory, f) for f in os.listdir(fan_work_directory)]
    random.seed(4815162342)
    random.shuffle(fan_works)
    fan_works = fan_works[subsample_start:subsample_end]

    start = 0
    fan_clusters = [fan_works[i:i + chunk_size] for i in range(start, len(fan_works), chunk_size)]
    filename_base = 'match-{}gram{{}}'.format(window_size)
    batch_filename = filename_base.format('-batch-{}.csv')

    rank, world = _dist_env()
    if world > 1:
        from .parallel import init_process_group
        init_process_group()
    ann_index = AnnIndexSearch(original_script_markup, window_size, number_of_hashes,
                               hash_dimensions, distance_threshold)
    global _ANN_INDEX
    _ANN_INDEX = ann_index
    ANALYZE_STATS['index_ready'] = time.perf_counter()

    # Optional (reuse_histogram=True / a path, or FANDOM_SEARCH_REUSE_HISTOGRAM): the
    # per-script-word reuse counts of `ao3.py format` (ao3.py:351-363,407-411), accumulated on the GPU
    # from the winning rows while the search runs, summed over the ranks with one all_reduce and
    # written next to the aggregate as match-{w}gram-YYYYMMDD[-k]-reuse.csv
    if reuse_histogram is None:
        reuse_histogram = os.environ.get('FANDOM_SEARCH_REUSE_HISTOGRAM') or None
    hist = None
    if reuse_histogram:
        from .aggregate import ReuseHistogram
        hist = ReuseHistogram(len(ann_index.word_lowercase), device=_device_ordinal())
        if getattr(ann_index.engine.index, 'device_records', False):
            hist.attach(ann_index.engine.index)

    def count_host_records(found0, best):
        # clusters whose records the device made are counted by the device itself
        if hist is not None and not isinstance(found0, DeviceRows):
            hist.add_best(best)

    # Three overlapped stages per cluster: (1) native read + tokenise + encode of the NEXT
    # cluster, (2) the GPU search of this one, (3) records + batch CSV of the PREVIOUS one.
    # ctypes releases the GIL inside the native calls, so plain threads are enough.
    from concurrent.futures import ThreadPoolExecutor
    if world > 1:
        # cluster -> rank: balanced on the bytes of text per cluster, the same table on every rank
        from .parallel import assign_clusters
        if len(fan_clusters) >= 8 * world and not os.environ.get('FANDOM_SEARCH_BALANCE'):
            # many clusters per rank: round robin leaves a tail of at most one cluster in eight or more --
            # not worth a stat() of every file on every rank (1 M files: seconds)
            owner = assign_clusters([1] * len(fan_clusters), world, policy='roundrobin')
        else:
            sizes = [sum(_file_size(f) for f in c) for c in fan_clusters]
            owner = assign_clusters(sizes, world)
    else:
        owner = [0] * len(fan_clusters)
    mine = [(i, c) for i, c in enumerate(fan_clusters, start=start) if owner[i - start] == rank]

    def finish(i, prep, found):
        # rows are formatted ONCE, natively; the aggregate is assembled from the batch files
        _write_text(ann_index.records_text_prepared(prep, *found, on_best=count_host_records),
                    batch_filename.format(i))
        return i

    import collections
    # The GPU-driving thread re-takes the GIL after every native call; with the default 5 ms switch
    # interval it would wait that long behind the Python parts of the two helper threads.
    switch_interval = sys.getswitchinterval()
    sys.setswitchinterval(2e-4)
    failure = None
    try:
        with ThreadPoolExecutor(max_workers=1) as prep_pool, ThreadPoolExecutor(max_workers=2) as post_pool:
            pending = prep_pool.submit(ann_index.prepare, mine[0][1]) if mine else None
            on_gpu = collections.deque()         # clusters submitted to the GPU, oldest first (<= 2)
            in_flight = collections.deque()      # records/CSV stages running (a few clusters' buffers alive)

            def collect_oldest():
                i0, prep0, ticket0 = on_gpu.popleft()
                found = ann_index.collect_prepared(prep0, ticket0)
                ANALYZE_STATS['collected'].append((i0, time.perf_counter(), ann_index.windows_processed))
                in_flight.append(post_pool.submit(finish, i0, prep0, found))
                while len(in_flight) > 3:
                    in_flight.popleft().result()

            for k, (i, fan_cluster) in enumerate(mine):
                print('Processing cluster {} ({}-{})'.format(i, chunk_size * i, chunk_size * (i + 1)))
                prep = pending.result()
                if k == 0:
                    ANALYZE_STATS['first_prepared'] = time.perf_counter()
                pending = prep_pool.submit(ann_index.prepare, mine[k + 1][1]) if k + 1 < len(mine) else None
                # the GPU always holds the next cluster: submit this one BEFORE waiting for the previous
                on_gpu.append((i, prep, ann_index.submit_prepared(prep, rows=True)))
                if k == 0:
                    ANALYZE_STATS['first_submitted'] = time.perf_counter()
                del prep
                if len(on_gpu) == 2:
                    collect_oldest()
            while on_gpu:
                collect_oldest()
            while in_flight:
                in_flight.popleft().result()
        ANALYZE_STATS['searched'] = time.perf_counter()
    except Exception as exc:     # noqa: BLE001 -- reported to every rank below, then re-raised
        if world <= 1:
            raise
        failure = exc
    finally:
        sys.setswitchinterval(switch_interval)
    if world > 1:
        # a rank that failed (unreadable file, out of memory, ...) tells the others instead of leaving
        # them at the barrier until the collective times out
        from .parallel import any_rank_failed
        if any_rank_failed(failure is not None):
            if failure is not None:
                raise failure
            raise RuntimeError("another rank failed while searching its clusters; see its log")

    if hist is not None:
        hist.detach()
    if world > 1:
        # every rank has written its own batch files (same directory, one node): rank 0 only has
        # to wait for them -- no record ever crosses ranks
        from .parallel import barrier
        barrier()
        if hist is not None:
            hist.all_reduce()       # the one collective of the path: [n_script_words, 11] int64
        if rank != 0:
            ANALYZE_STATS['end'] = time.perf_counter()
            return

    i = 0
    today_str = '-{:%Y%m%d}.csv'.format(datetime.date.today())
    name_check = filename_base.format(today_str)
    while os.path.exists(name_check):
        i += 1
        today_str = '-{:%Y%m%d}-{}.csv'.format(datetime.date.today(), i)
        name_check = filename_base.format(today_str)
    # header row (search.py:367) + the rows of every cluster in cluster order (search.py:388),
    # streamed from the batch files: memory stays flat however large the corpus is
    if hist is not None:
        hist_name = reuse_histogram if isinstance(reuse_histogram, str) and reuse_histogram not in ('1', 'true', 'True') \
            else name_check[:-4] + '-reuse.csv'
        hist.write_csv(hist_name, ann_index.word_lowercase)
    header = format_records([new_record_structure['fields']]).encode('utf-8')
    parts = [batch_filename.format(ci) for ci in range(start, start + len(fan_clusters))]
    try:
        _concatenate_files(name_check, header, parts)
    except FileNotFoundError as exc:
        raise RuntimeError(
            "%s is missing: under torchrun every rank writes the batch files of its clusters into "
            "the CURRENT DIRECTORY and rank 0 assembles the aggregate from them -- all ranks must "
            "run on one node (or share that directory)" % exc.filename) from None
    ANALYZE_STATS['end'] = time.perf_counter()
