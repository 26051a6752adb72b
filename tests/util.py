"""Helpers shared by the CPU and GPU tests: CSV reading and tolerance-aware record compare."""
import csv
import os

FLOAT_COLS = (9, 11)
INT_COLS = (1, 3, 4, 6, 10)


def parse_row(r):
    out = list(r)
    for c in INT_COLS:
        out[c] = int(out[c])
    for c in FLOAT_COLS:
        out[c] = float(out[c])
    out[7] = out[7] if out[7] != '' else None
    out[8] = int(out[8]) if out[8] not in ('', None) else None
    return out


def read_csv(path, header=None):
    with open(path, newline='', encoding='utf-8') as f:
        rows = list(csv.reader(f))
    if header is None:
        header = bool(rows) and rows[0] and rows[0][0] == 'FAN_WORK_FILENAME'
    if header:
        rows = rows[1:]
    return [parse_row(r) for r in rows]


def normalise(records):
    """Records as produced in memory (None / ints / floats) -> same typing as parse_row."""
    out = []
    for r in records:
        r = list(r)
        r[1], r[3], r[4], r[6], r[10] = int(r[1]), int(r[3]), int(r[4]), int(r[6]), int(r[10])
        r[9], r[11] = float(r[9]), float(r[11])
        r[8] = None if r[8] is None else int(r[8])
        out.append(r)
    return out


def compare_records(got, want, tol=1e-9, tie_tol=1e-12, basename=True):
    """Row sets keyed by (filename, fan word index) must be identical; string/int columns
    equal; float columns within `tol`.

    One exception, restricted to EXACT-REUSE rows: the reference's per-word argmin
    (search.py:224-225) is decided by float noise when several windows cover the word with a
    combined distance of ~ +-1e-16 (identical window vectors, distance = 1 - dot(u, u); SURVEY
    7.3-3) -- nothing a re-implementation (or another BLAS build under the reference itself) can
    reproduce.  Only when BOTH the wanted and the produced row have |BEST_MATCH_DISTANCE| <
    `tie_tol` (the window vectors are identical; BEST_COMBINED_DISTANCE is that noise times the
    Levenshtein distance) may the winning window differ (script word index / word / character /
    scene / Levenshtein columns).  Every other row must agree in every column; the tolerance of
    BEST_COMBINED_DISTANCE = distance * lev scales with lev.  Returns the number of such
    exact-reuse alternates."""
    def key(r):
        fn = os.path.basename(r[0]) if basename else r[0]
        return (fn, r[1])
    g = {key(r): r for r in got}
    w = {key(r): r for r in want}
    assert len(g) == len(got), "duplicate keys in result"
    missing = sorted(set(w) - set(g))
    extra = sorted(set(g) - set(w))
    assert not missing and not extra, "missing %s extra %s" % (missing[:5], extra[:5])
    ties = 0
    for k, wr in w.items():
        gr = g[k]
        assert gr[2] == wr[2] and gr[3] == wr[3], (gr, wr)
        strict = (gr[4:9] == wr[4:9] and gr[10] == wr[10]
                  and abs(gr[9] - wr[9]) <= tol and abs(gr[11] - wr[11]) <= tol * max(1, wr[10]))
        if strict:
            continue
        exact_reuse = abs(wr[9]) < tie_tol and abs(gr[9]) < tie_tol
        assert exact_reuse, (gr, wr)
        ties += 1
    return ties
