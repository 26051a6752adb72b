"""GPU parity tests: every CUDA stage, called through the C ABI, against the float64 numpy
checker (tests/numpy_index.py), the oracle and the golden CSVs of the unmodified reference.

Bar: bit-exact for integer/index work (gather, hash-join, match sets, CSV string/int columns);
|delta distance| <= 1e-12 for the float64 rescoring (tolerance stated here; the reference's own
BLAS order is not reproducible bit-for-bit)."""
import argparse
import glob
import os

import numpy as np
import pytest

from fandom_search_b200 import _native as nt
from fandom_search_b200 import search, synth
from fandom_search_b200.lexicon import Lexicon, py_hash_seed0
from tests.numpy_index import NumpyIndex
from tests.util import compare_records, normalise, read_csv

pytestmark = pytest.mark.gpu

DIST_TOL = 1e-12


def _device_index(*a, bits=16, **kw):
    """fp16 operands unless asked otherwise (the library default is fp8, see _f8_index and
    test_library_defaults): the stage-level expectations below are written for fp16."""
    from fandom_search_b200.engine import DeviceIndex
    idx = DeviceIndex(*a, **kw)
    if bits is not None:
        # ... and for operand rows that keep every embedding column (the library default drops the
        # lowest-energy ones, see test_prefilter_columns_*)
        idx.set_option(nt.FS_OPT_PREFILTER_DIMS, 0)
        if idx.operand_bits != bits:
            idx.set_option(nt.FS_OPT_OPERAND_BITS, bits)
    return idx


def _case(seed, vocab=800, dim=300, n_script=700, works=(200, 3, 0, 6, 397, 150), n_extra_s=3,
          n_extra_f=4, plant=True, clustered=True):
    rng = np.random.default_rng(seed)
    if clustered:
        centres = rng.standard_normal((vocab // 8 + 1, dim)).astype(np.float32)
        table = (0.8 * centres[np.arange(vocab) // 8] + 0.6 * rng.standard_normal((vocab, dim))).astype(np.float32)
    else:
        table = rng.standard_normal((vocab, dim)).astype(np.float32)
    sx = np.zeros((n_extra_s, dim), np.float32)
    fx = np.zeros((n_extra_f, dim), np.float32)
    for m in (sx, fx):
        for r in range(m.shape[0]):
            m[r, rng.integers(0, dim, 3)] = 1.0
    script = rng.integers(0, vocab + n_extra_s, n_script).astype(np.int32)
    off = np.concatenate([[0], np.cumsum(works)]).astype(np.int64)
    tok = rng.integers(0, vocab + n_extra_s + n_extra_f, int(off[-1])).astype(np.int32)
    if plant:
        for w in range(len(works)):
            a, b = int(off[w]), int(off[w + 1])
            if b - a >= 40:
                src = int(rng.integers(0, n_script - 30))
                ln = int(rng.integers(6, 25))
                dst = int(rng.integers(a, b - ln))
                tok[dst:dst + ln] = script[src:src + ln]
                # near copy: one within-cluster substitution
                src2 = int(rng.integers(0, n_script - 12))
                dst2 = int(rng.integers(a, b - 12))
                tok[dst2:dst2 + 12] = script[src2:src2 + 12]
                p = dst2 + 3
                if tok[p] < vocab:
                    tok[p] = (tok[p] // 8) * 8 + (tok[p] + 1) % 8
                    tok[p] = min(tok[p], vocab - 1)
    return table, sx, fx, script, tok, off


def _pairs(m):
    return set(zip(m['fan_pos'].tolist(), m['script_pos'].tolist()))


@pytest.mark.parametrize("pair", [0, 1, 2])          # 2 = CTA pair + resident fan tile
@pytest.mark.parametrize("diag", [1, 2, 3, 6])
@pytest.mark.parametrize("seed,dim", [(1, 300), (2, 64), (3, 768), (4, 100)])
def test_search_equals_float64_reference(seed, dim, diag, pair):
    table, sx, fx, script, tok, off = _case(seed, dim=dim)
    ref = NumpyIndex(table, script, extra=sx)
    want, wc = ref.search_host(tok, off, fx)
    idx = _device_index(table, script, extra=sx)
    idx.set_option(nt.FS_OPT_DIAG, diag)
    idx.set_option(nt.FS_OPT_CTA_PAIR, 1 if pair else 0)
    idx.set_option(nt.FS_OPT_A_RESIDENT, 1 if pair == 2 else 0)
    got, gc = idx.search_host(tok, off, fx)
    assert _pairs(got) == _pairs(want) and len(got) == len(want)
    assert len(want) > 0
    wd = {(a, b): (d, w, f) for a, b, d, w, f in zip(want['fan_pos'].tolist(), want['script_pos'].tolist(),
                                                      want['distance'].tolist(), want['work'].tolist(),
                                                      want['flags'].tolist())}
    for a, b, d, w, f in zip(got['fan_pos'].tolist(), got['script_pos'].tolist(), got['distance'].tolist(),
                             got['work'].tolist(), got['flags'].tolist()):
        rd, rw, rf = wd[(a, b)]
        assert abs(d - rd) <= DIST_TOL and w == rw and f == rf
    assert gc[nt.FS_CNT_WINDOWS] == wc[nt.FS_CNT_WINDOWS]
    assert gc[nt.FS_CNT_CANDIDATES] >= gc[nt.FS_CNT_MATCHES] == len(want)
    assert idx.n_script_windows == ref.n_script_windows
    idx.close()


@pytest.mark.parametrize("bits", [16, 8])
@pytest.mark.parametrize("window", [3, 4, 5, 7, 8])
def test_other_window_sizes(window, bits):
    """window_size is a keyword of analyze (search.py:337); the kernel picks E = 3, 2 or 1."""
    table, sx, fx, script, tok, off = _case(20 + window, dim=100)
    ref = NumpyIndex(table, script, extra=sx, window=window)
    want, wc = ref.search_host(tok, off, fx)
    idx = _device_index(table, script, extra=sx, window=window, bits=bits)
    assert idx.diag == (3 if window % 3 == 0 else 2 if window % 2 == 0 else 1)   # dim 100 < 416
    got, gc = idx.search_host(tok, off, fx)
    assert _pairs(got) == _pairs(want) and len(want) > 0
    assert gc[nt.FS_CNT_WINDOWS] == wc[nt.FS_CNT_WINDOWS]
    pj, _ = idx.exact_join_host(tok, off)
    wj, _ = ref.exact_join_host(tok, off)
    assert sorted(map(tuple, pj.tolist())) == sorted(map(tuple, wj.tolist()))
    idx.close()


@pytest.mark.parametrize("diag", [3, 6])
def test_unpacked_shuffles_give_the_same_matches(diag):
    """FS_OPT_PACKED_SHUFFLE=0 (full fp32 row shuffles) and the default fp16x2-packed shuffles
    must end in the same float64-decided match set."""
    table, sx, fx, script, tok, off = _case(31)
    ref = NumpyIndex(table, script, extra=sx)
    want, _ = ref.search_host(tok, off, fx)
    for pack in (0, 1, 2):
        idx = _device_index(table, script, extra=sx)
        idx.set_option(nt.FS_OPT_DIAG, diag)
        idx.set_option(nt.FS_OPT_PACKED_SHUFFLE, pack)
        got, _ = idx.search_host(tok, off, fx)
        assert _pairs(got) == _pairs(want)
        idx.close()


def test_gather_is_bit_exact_and_norms_match():
    import torch
    table, sx, fx, script, tok, off = _case(5)
    idx = _device_index(table, script, extra=sx)
    tok_t, off_t, fx_t = idx.to_device(tok, off, fx)
    emb, thr = idx.stage_embed(tok_t, off_t, fx_t)
    torch.cuda.synchronize()
    allrows = np.concatenate([table, sx, fx], axis=0)
    scale = np.float32(idx.scale)
    want16 = np.zeros((len(tok), idx.dim_pad), np.float16)
    want16[:, :table.shape[1]] = (allrows[tok] * scale).astype(np.float16)
    got16 = emb.cpu().numpy()
    assert got16.dtype == np.float16 and np.array_equal(got16.view(np.uint16), want16.view(np.uint16))
    # per window: (norm of the scaled window, norm of its fp16 rounding error), NaN where the window
    # leaves its work
    thr = thr.cpu().numpy()
    assert thr.shape[1] == 4 and np.all(thr[:, 3] == 0)
    assert np.all(thr[np.isfinite(thr[:, 2]), 2] == 0)      # nothing dropped: every column is kept
    thr = thr[:, :2]
    x = allrows.astype(np.float64) * float(scale)
    back = (allrows * scale).astype(np.float16).astype(np.float64)
    sq = (x[tok] ** 2).sum(axis=1)
    er = ((x[tok] - back[tok]) ** 2).sum(axis=1)
    valid = np.zeros(len(tok), bool)
    wantn = np.zeros(len(tok))
    wante = np.zeros(len(tok))
    for a, b in zip(off[:-1], off[1:]):
        for i in range(int(a), int(b) - 5):
            valid[i] = True
            wantn[i] = np.sqrt(sq[i:i + 6].sum())
            wante[i] = np.sqrt(er[i:i + 6].sum())
    assert np.all(np.isnan(thr[~valid])) and np.all(np.isfinite(thr[valid]))
    np.testing.assert_allclose(thr[valid, 0], wantn[valid], rtol=2e-6)
    np.testing.assert_allclose(thr[valid, 1], wante[valid], rtol=2e-4)
    assert np.all(thr[valid, 1] < 1e-3 * thr[valid, 0])       # fp16: ~3e-4 relative
    idx.close()


def test_library_defaults():
    """6-gram windows: fp8 operands, E = 6, CTA pairs, resident fan tile, fp16x2 epilogue."""
    table, sx, fx, script, tok, off = _case(5)
    idx = _device_index(table, script, extra=sx, bits=None)
    assert idx.operand_bits == 8 and idx.diag == 6 and idx.cta_pair == 1 and idx.info(7) == 1 and idx.info(8) == 2
    assert idx.info(12) == 103 and idx.info(15) == 1          # ... in 128-column tiles, two CTA pairs per TPC         # grouped stages + early accumulator release + one-pass epilogue
    assert idx.kept_dims == 256 and idx.dim_pad == 256      # two 128-byte chunks of the 300 columns
    want, _ = NumpyIndex(table, script, extra=sx).search_host(tok, off, fx)
    got, _ = idx.search_host(tok, off, fx)
    assert _pairs(got) == _pairs(want)
    idx.close()
    table, sx, fx, script, tok, off = _case(3, dim=768)
    idx = _device_index(table, script, extra=sx, bits=None)
    assert idx.operand_bits == 8 and idx.diag == 6 and idx.kept_dims == 640 and idx.info(15) == 0
    idx.close()


@pytest.mark.parametrize("pair", [0, 1, 2])
@pytest.mark.parametrize("diag,shifts", [(1, 1), (1, 2), (1, 3), (1, 6), (2, 1), (2, 3), (3, 1), (3, 2), (6, 1)])
def test_tensor_core_dots_match_fp16_contraction(diag, shifts, pair):
    import torch
    table, sx, fx, script, tok, off = _case(6, plant=False, clustered=False)
    idx = _device_index(table, script, extra=sx)
    idx.set_option(nt.FS_OPT_DIAG, diag)
    idx.set_option(nt.FS_OPT_CTA_PAIR, 1 if pair else 0)
    idx.set_option(nt.FS_OPT_A_RESIDENT, 1 if pair == 2 else 0)
    idx.set_option(nt.FS_OPT_SHIFTS_PER_STAGE, shifts)
    tok_t, off_t, fx_t = idx.to_device(tok, off, fx)
    dots = idx.stage_dots(tok_t, off_t, fx_t).cpu().numpy()
    allrows = np.concatenate([table, sx, fx], axis=0)
    e16 = (allrows * np.float32(idx.scale)).astype(np.float16).astype(np.float32)
    ef = np.zeros((len(tok) + 6, table.shape[1]), np.float32)
    es = np.zeros((len(script) + 6, table.shape[1]), np.float32)
    ef[:len(tok)] = e16[tok]
    es[:len(script)] = e16[script]
    g = ef.astype(np.float64) @ es.astype(np.float64).T
    want = sum(g[k:k + len(tok), k:k + len(script)] for k in range(6))
    # fp32 accumulation of 1800 fp16 products (+ fp16-packed shuffles for E = 3, 6): tolerance
    # 2e-3 of the largest magnitude
    assert np.abs(dots - want).max() <= 2e-3 * np.abs(want).max()
    idx.close()


@pytest.mark.parametrize("diag", [2, 3, 6])
def test_half_precision_epilogue_dots(diag):
    """E = 3, 6 with the diagonal summed in fp16x2 arithmetic (pack level 2): looser tolerance,
    bounded by 2^-9 * sum of the six partial-dot magnitudes."""
    import torch
    table, sx, fx, script, tok, off = _case(6, plant=False, clustered=False)
    for pair in (0, 1, 2):
        idx = _device_index(table, script, extra=sx)
        idx.set_option(nt.FS_OPT_DIAG, diag)
        idx.set_option(nt.FS_OPT_PACKED_SHUFFLE, 2)
        idx.set_option(nt.FS_OPT_CTA_PAIR, 1 if pair else 0)
        idx.set_option(nt.FS_OPT_A_RESIDENT, 1 if pair == 2 else 0)
        tok_t, off_t, fx_t = idx.to_device(tok, off, fx)
        dots = idx.stage_dots(tok_t, off_t, fx_t).cpu().numpy()
        allrows = np.concatenate([table, sx, fx], axis=0)
        e16 = (allrows * np.float32(idx.scale)).astype(np.float16).astype(np.float64)
        ef = np.zeros((len(tok) + 6, table.shape[1])); ef[:len(tok)] = e16[tok]
        es = np.zeros((len(script) + 6, table.shape[1])); es[:len(script)] = e16[script]
        g = ef @ es.T
        parts = [g[k:k + len(tok), k:k + len(script)] for k in range(6)]
        want = sum(parts)
        bound = 2.0 ** -9 * sum(np.abs(q) for q in parts) + 1e-4 * np.abs(want).max()
        # rows whose window leaves the tile grid are not dumped (zeros): compare where dumped
        mask = dots != 0
        assert mask.mean() > 0.9 and np.all(np.abs(dots - want)[mask] <= bound[mask])
        idx.close()


# ---------------------------------------------------------------------------------------------
# fp8 e4m3 operands (FS_OPT_OPERAND_BITS = 8): measured rounding error in the pre-filter threshold
# ---------------------------------------------------------------------------------------------
def _f8_index(table, script, sx, diag=None, **opts):
    idx = _device_index(table, script, extra=sx, bits=8)
    if diag is not None:
        idx.set_option(nt.FS_OPT_DIAG, diag)
    for k, v in opts.items():
        idx.set_option(k, v)
    return idx


def _e4m3(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.float8_e4m3fn)


@pytest.mark.parametrize("pack", [1, 2])
@pytest.mark.parametrize("diag", [1, 2, 3, 6])
@pytest.mark.parametrize("seed,dim", [(1, 300), (2, 64), (3, 768), (4, 100), (9, 50)])
def test_fp8_search_equals_float64_reference(seed, dim, diag, pack):
    table, sx, fx, script, tok, off = _case(seed, dim=dim)
    want, wc = NumpyIndex(table, script, extra=sx).search_host(tok, off, fx)
    idx = _f8_index(table, script, sx, diag)
    idx.set_option(nt.FS_OPT_PACKED_SHUFFLE, pack)
    assert idx.operand_bits == 8 and idx.dim_pad % 32 == 0
    got, gc = idx.search_host(tok, off, fx)
    assert _pairs(got) == _pairs(want) and len(got) == len(want) and len(want) > 0
    wd = {(a, b): d for a, b, d in zip(want['fan_pos'].tolist(), want['script_pos'].tolist(), want['distance'].tolist())}
    for a, b, d in zip(got['fan_pos'].tolist(), got['script_pos'].tolist(), got['distance'].tolist()):
        assert abs(d - wd[(a, b)]) <= DIST_TOL
    assert gc[nt.FS_CNT_WINDOWS] == wc[nt.FS_CNT_WINDOWS]
    # switching back re-converts the index to fp16
    idx.set_option(nt.FS_OPT_OPERAND_BITS, 16)
    got16, _ = idx.search_host(tok, off, fx)
    assert _pairs(got16) == _pairs(want)
    idx.close()


def test_fp8_gather_is_bit_exact_and_thresholds_hold_the_measured_error():
    import torch
    table, sx, fx, script, tok, off = _case(5)
    idx = _f8_index(table, script, sx)
    tok_t, off_t, fx_t = idx.to_device(tok, off, fx)
    emb, thr = idx.stage_embed(tok_t, off_t, fx_t)
    torch.cuda.synchronize()
    allrows = np.concatenate([table, sx, fx], axis=0)
    scale = np.float32(idx.scale)
    # scale = sqrt(55000 / window) / largest row norm of the table and the script extras
    big = np.sqrt((np.concatenate([table, sx]).astype(np.float64) ** 2).sum(axis=1).max())
    np.testing.assert_allclose(scale, np.sqrt(55000.0 / 6) / big, rtol=1e-5)
    want8 = torch.zeros((len(tok), idx.dim_pad), dtype=torch.uint8)
    q = _e4m3(allrows[tok] * scale)
    want8[:, :table.shape[1]] = q.view(torch.uint8)
    assert torch.equal(emb.cpu(), want8)
    # per window: (|f|, |f - q(f)|) from the measured rounding error of every e4m3 row
    x = allrows.astype(np.float64) * float(scale)
    back = _e4m3(allrows * scale).float().numpy().astype(np.float64)
    sq, er = (x ** 2).sum(axis=1), ((x - back) ** 2).sum(axis=1)
    thr = thr.cpu().numpy()
    rel = []
    for a, b in zip(off[:-1], off[1:]):
        for i in range(int(a), int(b) - 5):
            f_sq, f_er = sq[tok[i:i + 6]].sum(), er[tok[i:i + 6]].sum()
            assert abs(thr[i, 0] - np.sqrt(f_sq)) <= 2e-6 * np.sqrt(f_sq)
            assert abs(thr[i, 1] - np.sqrt(f_er)) <= 2e-4 * np.sqrt(f_er) + 1e-6
            rel.append(np.sqrt(f_er / f_sq))
    assert 0.01 < max(rel) < 0.06                    # e4m3: ~3 % relative per window
    idx.close()


@pytest.mark.parametrize("diag", [1, 2, 3, 6])
def test_fp8_dots_match_the_e4m3_contraction(diag):
    import torch
    table, sx, fx, script, tok, off = _case(6, plant=False, clustered=False, works=(900, 3, 0, 6, 1400, 700))
    idx = _f8_index(table, script, sx, diag)
    tok_t, off_t, fx_t = idx.to_device(tok, off, fx)
    dots = idx.stage_dots(tok_t, off_t, fx_t).cpu().numpy()
    allrows = np.concatenate([table, sx, fx], axis=0)
    e8 = _e4m3(allrows * np.float32(idx.scale)).float().numpy().astype(np.float64)
    ef = np.zeros((len(tok) + 6, table.shape[1])); ef[:len(tok)] = e8[tok]
    es = np.zeros((len(script) + 6, table.shape[1])); es[:len(script)] = e8[script]
    g = ef @ es.T
    parts = [g[k:k + len(tok), k:k + len(script)] for k in range(6)]
    want = sum(parts)
    # products of e4m3 values are exact; fp32 accumulation + fp16x2 epilogue sums (2^-9 for E = 6)
    bound = 2.0 ** -9 * sum(np.abs(q) for q in parts) + 1e-4 * np.abs(want).max()
    assert np.all(np.abs(dots - want) <= bound)
    idx.close()


@pytest.mark.parametrize("diag", [2, 3, 6])
def test_fp8_candidates_are_a_superset(diag):
    import torch
    table, sx, fx, script, tok, off = _case(7, works=(900, 3, 0, 6, 1400, 700))
    ref = NumpyIndex(table, script, extra=sx)
    d, fpos = ref.distances(tok, off, fx)
    idx = _f8_index(table, script, sx, diag)
    tok_t, off_t, fx_t = idx.to_device(tok, off, fx)
    cand, cnt = idx.stage_candidates(tok_t, off_t, fx_t)
    n = int(cnt.cpu()[nt.FS_CNT_CANDIDATES])
    cand = cand.cpu().numpy()[:n]
    got = set(map(tuple, cand.tolist()))
    assert len(got) == len(cand)
    ii, jj = np.nonzero(d < 0.1)
    must = set(zip(fpos[ii].tolist(), ref.spos[jj].tolist()))
    assert must <= got
    row_of = {int(p): i for i, p in enumerate(fpos)}
    col_of = {int(p): j for j, p in enumerate(ref.spos)}
    worst = max(d[row_of[a], col_of[b]] for a, b in got)
    assert worst < 0.1 + 0.15            # the fp8 slack is wide but finite
    assert len(got) <= 4 * len(must) + 64
    idx.close()


@pytest.mark.parametrize("diag", [3, 6])
def test_fp8_resident_fan_tile(diag):
    for seed, dim, works in ((1, 300, (200, 3, 0, 6, 397, 150)), (4, 100, (900, 3, 0, 6, 1400, 700))):
        table, sx, fx, script, tok, off = _case(seed, dim=dim, works=works)
        want, _ = NumpyIndex(table, script, extra=sx).search_host(tok, off, fx)
        for grid in (0, 4):
            idx = _f8_index(table, script, sx, diag)
            idx.set_option(nt.FS_OPT_A_RESIDENT, 1)
            idx.set_option(nt.FS_OPT_GRID_LIMIT, grid)
            got, _ = idx.search_host(tok, off, fx)
            assert _pairs(got) == _pairs(want) and len(got) == len(want)
            idx.close()


@pytest.mark.parametrize("dim", [300, 200, 100])
def test_grouped_stages_and_early_release(dim):
    """FS_OPT_TILE_GROUP: one barrier pair and one block of MMAs per script tile (bit 0), accumulator
    handed back before the sums (bit 1), both chunks of a warp in one pass with fp32 row maxima
    (bit 2).  dim 300/200/100 = 3/2/1 chunks per row with 2/3/4 K-steps
    in the last one; a 4-CTA grid makes every CTA wrap its ring of tile groups many times."""
    table, sx, fx, script, tok, off = _case(11, dim=dim, works=(900, 3, 0, 6, 1400, 700))
    want, _ = NumpyIndex(table, script, extra=sx).search_host(tok, off, fx)
    dots = {}
    for group in (0, 1, 2, 3, 4, 7, 39, 20, 23, 28, 31, 55, 103, 71):
        for grid in (0, 4):
            idx = _f8_index(table, script, sx, 6)
            idx.set_option(nt.FS_OPT_A_RESIDENT, 1)
            idx.set_option(nt.FS_OPT_TILE_GROUP, group)
            idx.set_option(nt.FS_OPT_GRID_LIMIT, grid)
            assert idx.info(12) == group
            got, _ = idx.search_host(tok, off, fx)
            assert _pairs(got) == _pairs(want) and len(got) == len(want)
            tok_t, off_t, fx_t = idx.to_device(tok, off, fx)
            dots[group, grid] = idx.stage_dots(tok_t, off_t, fx_t).cpu().numpy()
            idx.close()
    for key, d in dots.items():
        assert np.array_equal(d, dots[0, 0]), key      # the same products in the same order


def test_fp8_needs_cta_pairs():
    table, sx, fx, script, tok, off = _case(1, dim=64)
    idx = _f8_index(table, script, sx)
    idx.set_option(nt.FS_OPT_CTA_PAIR, 0)
    with pytest.raises(nt.NativeError):
        idx.search_host(tok, off, fx)
    idx.close()


@pytest.mark.parametrize("diag", [2, 3, 6])
@pytest.mark.parametrize("pack", [1, 2])
def test_half_precision_epilogue_search(pack, diag):
    for seed, dim in ((1, 300), (2, 64), (4, 100)):
        table, sx, fx, script, tok, off = _case(seed, dim=dim)
        want, _ = NumpyIndex(table, script, extra=sx).search_host(tok, off, fx)
        for pair in (0, 1, 2):
            idx = _device_index(table, script, extra=sx)
            idx.set_option(nt.FS_OPT_DIAG, diag)
            idx.set_option(nt.FS_OPT_PACKED_SHUFFLE, pack)
            idx.set_option(nt.FS_OPT_CTA_PAIR, 1 if pair else 0)
            idx.set_option(nt.FS_OPT_A_RESIDENT, 1 if pair == 2 else 0)
            got, _ = idx.search_host(tok, off, fx)
            assert _pairs(got) == _pairs(want)
            idx.close()


@pytest.mark.parametrize("pair", [0, 1, 2])
@pytest.mark.parametrize("diag", [1, 2, 3, 6])
def test_candidates_are_a_superset_within_slack(diag, pair):
    import torch
    table, sx, fx, script, tok, off = _case(7)
    ref = NumpyIndex(table, script, extra=sx)
    d, fpos = ref.distances(tok, off, fx)
    idx = _device_index(table, script, extra=sx)
    idx.set_option(nt.FS_OPT_DIAG, diag)
    idx.set_option(nt.FS_OPT_CTA_PAIR, 1 if pair else 0)
    idx.set_option(nt.FS_OPT_A_RESIDENT, 1 if pair == 2 else 0)
    tok_t, off_t, fx_t = idx.to_device(tok, off, fx)
    cand, cnt = idx.stage_candidates(tok_t, off_t, fx_t)
    n = int(cnt.cpu()[nt.FS_CNT_CANDIDATES])
    cand = cand.cpu().numpy()[:n]
    got = set(map(tuple, cand.tolist()))
    assert len(got) == len(cand)          # overlapping tiles must not emit a pair twice
    row_of = {int(p): i for i, p in enumerate(fpos)}
    col_of = {int(p): j for j, p in enumerate(ref.spos)}
    ii, jj = np.nonzero(d < 0.1)
    must = set(zip(fpos[ii].tolist(), ref.spos[jj].tolist()))
    assert must <= got
    for a, b in got:                      # nothing far from the threshold gets through
        assert d[row_of[a], col_of[b]] < 0.1 + 2 * 4.0e-3
    idx.close()


def test_hash_join_is_bit_exact():
    table, sx, fx, script, tok, off = _case(8, n_script=900)
    script[300:306] = script[100:106]          # duplicate script 6-grams -> several hits per window
    script[500:506] = script[100:106]
    tok[20:30] = script[98:108]
    ref = NumpyIndex(table, script, extra=sx)
    want, _ = ref.exact_join_host(tok, off)
    idx = _device_index(table, script, extra=sx)
    got, cnt = idx.exact_join_host(tok, off)
    assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, want.tolist()))
    assert cnt[nt.FS_CNT_EXACT] == len(want) > 10
    # every exact pair is also found by the distance path, flagged exact, at distance ~0
    m, _ = idx.search_host(tok, off, fx)
    exact = {(a, b) for a, b, f in zip(m['fan_pos'].tolist(), m['script_pos'].tolist(), m['flags'].tolist())
             if f & nt.FS_MATCH_EXACT}
    assert exact == set(map(tuple, want.tolist()))
    assert np.all(np.abs(m['distance'][(m['flags'] & 1) == 1]) < 1e-12)
    idx.close()


@pytest.mark.parametrize("window", [6, 4])
def test_hash_join_with_heavy_collisions(window):
    # a 3-word vocabulary: most script windows repeat many times (chains of equal keys in the
    # table) and most fan windows hit several of them; window 4 runs the run-time-length kernel
    rng = np.random.default_rng(77)
    table = rng.standard_normal((3, 300)).astype(np.float32)
    script = rng.integers(0, 3, 1500).astype(np.int32)
    works = (700, 5, 0, 3000, 64)
    off = np.concatenate([[0], np.cumsum(works)]).astype(np.int64)
    tok = rng.integers(0, 3, int(off[-1])).astype(np.int32)
    ref = NumpyIndex(table, script, window=window)
    want, _ = ref.exact_join_host(tok, off)
    idx = _device_index(table, script, window=window)
    got, cnt = idx.exact_join_host(tok, off, cap=len(want) + 10)
    assert cnt[nt.FS_CNT_EXACT] == len(want) > 5000
    assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, want.tolist()))
    idx.close()


@pytest.mark.parametrize("n_script", [40000, 140000])
def test_hash_join_with_a_long_script(n_script):
    """Scripts of 40 k / 140 k tokens: the probe's filter takes 64 KB / 128 KB of shared memory (opt-in
    above 48 KB, fewer resident blocks); pairs equal the dictionary join."""
    rng = np.random.default_rng(n_script)
    table = rng.standard_normal((2000, 32)).astype(np.float32)
    script = rng.integers(0, 2000, n_script).astype(np.int32)
    works = (3000, 1, 2500)
    off = np.concatenate([[0], np.cumsum(works)]).astype(np.int64)
    tok = rng.integers(0, 2000, int(off[-1])).astype(np.int32)
    for k in range(40):                                         # planted quotes, some across the end of a work
        src = int(rng.integers(0, n_script - 30))
        dst = int(rng.integers(0, len(tok) - 30))
        tok[dst:dst + 20] = script[src:src + 20]
    ref = NumpyIndex(table, script)
    want, _ = ref.exact_join_host(tok, off)
    idx = _device_index(table, script)
    got, cnt = idx.exact_join_host(tok, off)
    assert cnt[nt.FS_CNT_EXACT] == len(want) > 300
    assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, want.tolist()))
    idx.close()


def test_ragged_and_empty_batches():
    table, sx, fx, script, _, _ = _case(9)
    idx = _device_index(table, script, extra=sx)
    ref = NumpyIndex(table, script, extra=sx)
    for works in ([], [0], [5], [6], [0, 0, 7, 0], [5, 5, 5]):
        off = np.concatenate([[0], np.cumsum(works)]).astype(np.int64)
        tok = np.resize(script[10:40], int(off[-1])).astype(np.int32)
        got, gc = idx.search_host(tok, off)
        want, wc = ref.search_host(tok, off)
        assert _pairs(got) == _pairs(want)
        assert gc[nt.FS_CNT_WINDOWS] == wc[nt.FS_CNT_WINDOWS] == sum(max(w - 5, 0) for w in works)
        pj, _ = idx.exact_join_host(tok, off)
        wj, _ = ref.exact_join_host(tok, off)
        assert sorted(map(tuple, pj.tolist())) == sorted(map(tuple, wj.tolist()))
    idx.close()


def test_multiple_scripts_do_not_straddle():
    table, sx, fx, script, tok, off = _case(10, n_script=600)
    soff = np.array([0, 200, 203, 600], np.int64)
    # a fan span copied across the script boundary must NOT match as one window
    tok[50:62] = script[194:206]
    ref = NumpyIndex(table, script, script_off=soff, extra=sx)
    idx = _device_index(table, script, script_off=soff, extra=sx)
    got, _ = idx.search_host(tok, off, fx)
    want, _ = ref.search_host(tok, off, fx)
    assert _pairs(got) == _pairs(want)
    assert all(not (195 <= b < 200) and not (198 <= b < 203) for _, b in _pairs(got))
    assert idx.n_script_windows == ref.n_script_windows == 195 + 0 + 392
    idx.close()


def test_overflow_is_reported_and_retried():
    table, sx, fx, script, tok, off = _case(11)
    ref = NumpyIndex(table, script, extra=sx)
    want, _ = ref.search_host(tok, off, fx)
    idx = _device_index(table, script, extra=sx)
    out = np.empty(2, dtype=nt.MATCH_DTYPE)
    cnt = np.zeros(nt.FS_CNT_COUNT, np.int64)
    st = idx._lib.fs_search_csr_host(idx._h, nt.ptr(tok), len(tok), nt.ptr(off), len(off) - 1,
                                     nt.ptr(fx), fx.shape[0], nt.ptr(out), 2, nt.ptr(cnt))
    assert st == nt.FS_E_OVERFLOW and cnt[nt.FS_CNT_MATCHES] == len(want)
    assert b"overflow" in idx._lib.fs_last_error()
    idx.close()
    idx2 = _device_index(table, script, extra=sx)
    idx2.reserve(len(tok), 4)                       # tiny candidate buffer
    assert idx2._lib.fs_index_get_info(idx2._h, 3) == 4
    got, _ = idx2.search_host(tok, off, fx, cap=3)  # wrapper grows both buffers and retries
    assert _pairs(got) == _pairs(want)
    idx2.close()


def test_device_entry_point_equals_host_entry_point():
    import torch
    table, sx, fx, script, tok, off = _case(12)
    idx = _device_index(table, script, extra=sx)
    host, hc = idx.search_host(tok, off, fx)
    tok_t, off_t, fx_t = idx.to_device(tok, off, fx)
    out_t = torch.empty(24 * 4096, dtype=torch.uint8, device="cuda")
    cnt_t = torch.zeros(nt.FS_CNT_COUNT, dtype=torch.int64, device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        idx.search_dev(tok_t, off_t, fx_t, out_t, cnt_t, stream=stream)
    stream.synchronize()
    cnt = cnt_t.cpu().numpy()
    dev = np.frombuffer(out_t.cpu().numpy().tobytes(), dtype=nt.MATCH_DTYPE)[:cnt[nt.FS_CNT_MATCHES]]
    assert _pairs(dev) == _pairs(host) and np.array_equal(cnt, hc)
    ms, n = idx.timing_read()
    assert n >= 2 and ms > 0
    idx.close()


def test_invalid_arguments_fail_loudly():
    table, sx, fx, script, tok, off = _case(13)
    lib = nt.load()
    import ctypes
    h = ctypes.c_void_p()
    bad = lib.fs_index_create(ctypes.byref(h), 0, nt.ptr(table), table.shape[0], table.shape[1], None, 0,
                              nt.ptr(script), len(script), nt.ptr(np.array([0, 5], np.int64)), 1, 6, 0.1)
    assert bad == nt.FS_E_INVALID and b"script_off" in lib.fs_last_error()
    with pytest.raises(nt.NativeError):
        _device_index(table, script, device=99)


def _golden_pipeline(golden_dir):
    search.set_pipeline(search.Pipeline(
        Lexicon.from_npz(os.path.join(golden_dir, "lexicon.npz"), hash_fn=py_hash_seed0)))


def test_golden_csv_end_to_end(golden_dir, tmp_path, monkeypatch):
    """ao3.py search equivalent on the committed corpus == CSV of the unmodified reference."""
    _golden_pipeline(golden_dir)
    try:
        listing = open(os.path.join(golden_dir, "listing.txt")).read().split()
        real_listdir = os.listdir
        monkeypatch.setattr(os, "listdir", lambda d: list(listing) if str(d) == "fanworks" else real_listdir(d))
        monkeypatch.chdir(tmp_path)
        os.symlink(os.path.join(golden_dir, "fanworks"), "fanworks")
        os.symlink(os.path.join(golden_dir, "script.txt"), "script.txt")
        args = argparse.Namespace(fan_works="fanworks", script="script.txt", skip_works=-1, num_works=-1)
        search.analyze(args, chunk_size=16)
        got = read_csv(glob.glob("match-6gram-2*.csv")[0])
        want = read_csv(os.path.join(golden_dir, "golden_exhaustive.csv"))
        compare_records(got, want, tol=DIST_TOL, basename=False)
        assert [(r[0], r[1]) for r in got] == [(r[0], r[1]) for r in want]
        for i in range(3):
            b = read_csv("match-6gram-batch-%d.csv" % i, header=False)
            wb = read_csv(os.path.join(golden_dir, "golden_exhaustive.batch%d.csv" % i), header=False)
            assert [(r[0], r[1]) for r in b] == [(r[0], r[1]) for r in wb]
    finally:
        search.set_pipeline(None)


def test_full_size_cluster_properties():
    """One BASELINE-size cluster (500 works x ~5k tokens vs a 25k-token script, d=300):
    size-independent properties instead of an O(N^2) CPU check."""
    lex = synth.SynthLexicon(vocab=50000, dim=300, oov_frac=0.0, seed=1001)
    script = synth.make_script_tokens(lex, 25000).astype(np.int32)
    words, off = synth.synth_csr_batch(lex, script, range(500))
    tok = words.astype(np.int32)
    idx = _device_index(lex.table_all, script, bits=None)       # the library defaults: fp8, E = 6, ...
    assert idx.operand_bits == 8 and idx.diag == 6
    m, cnt = idx.search_host(tok, off, cap=1 << 20)
    assert cnt[nt.FS_CNT_WINDOWS] == int(np.maximum(np.diff(off) - 5, 0).sum())
    # (1) every hash-join pair is found by the distance path, flagged exact, |distance| ~ 0
    pj, _ = idx.exact_join_host(tok, off, cap=1 << 20)
    exact = set(map(tuple, pj.tolist()))
    found = {(a, b): (d, f) for a, b, d, f in zip(m['fan_pos'].tolist(), m['script_pos'].tolist(),
                                                  m['distance'].tolist(), m['flags'].tolist())}
    assert len(exact) > 1000 and exact <= set(found)
    assert all(found[p][1] & 1 and abs(found[p][0]) < 1e-12 for p in exact)
    assert {p for p, v in found.items() if v[1] & 1} == exact
    # (2) all reported distances are below the threshold, windows lie inside their work
    assert np.all(m['distance'] < 0.1)
    assert np.all(m['fan_pos'] + 6 <= off[m['work'] + 1]) and np.all(m['fan_pos'] >= off[m['work']])
    # (3) spot check 200 reported pairs against a float64 recomputation
    t64 = lex.table_all.astype(np.float64)
    rng = np.random.default_rng(0)
    for k in rng.choice(len(m), 200, replace=False):
        f = t64[tok[m['fan_pos'][k]:m['fan_pos'][k] + 6]].ravel()
        s = t64[script[m['script_pos'][k]:m['script_pos'][k] + 6]].ravel()
        d = 1.0 - (f / np.linalg.norm(f)) @ (s / np.linalg.norm(s))
        assert abs(d - m['distance'][k]) <= DIST_TOL
    # (4) idempotence: a second pass returns the identical set (order aside)
    m2, cnt2 = idx.search_host(tok, off, cap=1 << 20)
    a = np.sort(m, order=['fan_pos', 'script_pos'])
    b = np.sort(m2, order=['fan_pos', 'script_pos'])
    assert a.tobytes() == b.tobytes() and np.array_equal(cnt, cnt2)
    # (5) every other kernel family ends in the same float64-decided match set: fp8 E = 3 (boundary
    # pass), fp16 E = 3, fp16 dense -- only the candidate counts differ
    key = np.sort(m['fan_pos'].astype(np.int64) * 100000 + m['script_pos'])
    for opts in ((nt.FS_OPT_DIAG, 3), (nt.FS_OPT_OPERAND_BITS, 16), (nt.FS_OPT_DIAG, 1)):
        idx.set_option(*opts)
        mo, co = idx.search_host(tok, off, cap=1 << 20)
        assert np.array_equal(key, np.sort(mo['fan_pos'].astype(np.int64) * 100000 + mo['script_pos'])), opts
        assert co[nt.FS_CNT_CANDIDATES] >= co[nt.FS_CNT_MATCHES] == len(m)
    idx.close()


def test_lsh_emulation_reproduces_seeded_reference_golden(golden_dir, tmp_path, monkeypatch):
    """FANDOM_SEARCH_MODE=lsh with the seed of the golden run == CSV of the unmodified reference
    over a seeded 15x14-bit random-hyperplane index (rows the LSH missed are missing here too)."""
    _golden_pipeline(golden_dir)
    monkeypatch.setenv("FANDOM_SEARCH_MODE", "lsh")
    monkeypatch.setenv("FANDOM_SEARCH_LSH_SEED", "7")
    try:
        listing = open(os.path.join(golden_dir, "listing.txt")).read().split()
        real_listdir = os.listdir
        monkeypatch.setattr(os, "listdir", lambda d: list(listing) if str(d) == "fanworks" else real_listdir(d))
        monkeypatch.chdir(tmp_path)
        os.symlink(os.path.join(golden_dir, "fanworks"), "fanworks")
        os.symlink(os.path.join(golden_dir, "script.txt"), "script.txt")
        args = argparse.Namespace(fan_works="fanworks", script="script.txt", skip_works=-1, num_works=-1)
        search.analyze(args, chunk_size=16)
        got = read_csv(glob.glob("match-6gram-2*.csv")[0])
        want = read_csv(os.path.join(golden_dir, "golden_lsh_seed7.csv"))
        exhaustive = read_csv(os.path.join(golden_dir, "golden_exhaustive.csv"))
        assert len(want) < len(exhaustive)
        compare_records(got, want, tol=DIST_TOL, basename=False)
        assert [(r[0], r[1]) for r in got] == [(r[0], r[1]) for r in want]
    finally:
        search.set_pipeline(None)


def test_reuse_histogram_matches_format_groupby(golden_dir):
    """fs_reuse_histogram_dev == the thresholded group-by of ao3.py format_data (ao3.py:351-363,
    407-411) applied to the reference's golden match CSV."""
    from fandom_search_b200.aggregate import THRESHOLDS, ReuseHistogram
    rows = read_csv(os.path.join(golden_dir, "golden_exhaustive.csv"))
    n_words = 1 + max(r[4] for r in rows)
    want = np.zeros((n_words, len(THRESHOLDS)), np.int64)
    for r in rows:                                   # matches.BEST_COMBINED_DISTANCE <= t, summed per word
        for k, t in enumerate(THRESHOLDS):
            want[r[4], k] += r[11] <= t
    hist = ReuseHistogram(n_words)
    half = len(rows) // 2
    hist.add_records(rows[:half])
    hist.add_records(rows[half:])
    got = hist.result()
    assert np.array_equal(got, want) and got.sum() > 0
    assert np.all(np.diff(got, axis=1) >= 0)         # cumulative in the threshold


def test_multi_script_pass_on_device(golden_dir, tmp_path):
    """N4 on the GPU: two scripts indexed side by side == two single-script searches."""
    _golden_pipeline(golden_dir)
    try:
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
            compare_records(got, want, tol=DIST_TOL)
    finally:
        search.set_pipeline(None)


@pytest.mark.parametrize("bits", [8, 16])
def test_heterogeneous_row_norms(bits):
    """Embedding rows whose norms span three orders of magnitude, near-zero rows, a few huge
    elements, windows of one repeated maximum-norm token (largest possible partial sums): the
    fp8/fp16 pre-filter must still hand every true match to the float64 rescoring."""
    rng = np.random.default_rng(77)
    table, sx, fx, script, tok, off = _case(77, dim=300, works=(400, 3, 0, 6, 500, 300))
    table = table * rng.lognormal(0.0, 1.5, size=(table.shape[0], 1)).astype(np.float32)
    table[5] *= 1e-6                    # near-zero row
    table[6] = 0.0                      # zero row
    table[7, :4] *= 50.0                # a few dominant elements
    big = int(np.argmax((table.astype(np.float64) ** 2).sum(axis=1)))
    script[100:112] = big               # twelve times the largest row in a row
    script[200:206] = [5, 6, 5, 6, 5, 6]
    tok[50:62] = big
    tok[70:76] = [5, 6, 5, 6, 5, 6]
    tok[80:86] = 6                      # an all-zero window
    # a per-batch (fan-side) row 1000 x longer than any row of the index, parallel to a table row that
    # stands alone in a script window: cosine 1, far outside the operand range chosen for the index
    fx = fx.copy()
    fx[0] = 1000.0 * table[big]
    id_fx0 = table.shape[0] + sx.shape[0]
    script[300:306] = [6, 6, big, 6, 6, 6]
    tok[90:96] = [6, 6, id_fx0, 6, 6, 6]
    ref = NumpyIndex(table, script, extra=sx)
    want, _ = ref.search_host(tok, off, fx)
    idx = _device_index(table, script, extra=sx, bits=bits)
    got, _ = idx.search_host(tok, off, fx)
    assert _pairs(got) == _pairs(want) and len(want) > 20
    assert (90, 300) in _pairs(want)
    for diag, pack in ((6, 2), (2, 2), (3, 1), (1, 0)):
        idx.set_option(nt.FS_OPT_DIAG, diag)
        idx.set_option(nt.FS_OPT_PACKED_SHUFFLE, pack)
        got, _ = idx.search_host(tok, off, fx)
        assert _pairs(got) == _pairs(want)
    idx.close()


# ---------------------------------------------------------------------------------------------
# pre-filter columns (FS_OPT_PREFILTER_DIMS): operand rows keep the highest-energy columns only
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bits,diag", [(8, 6), (8, 3), (16, 3), (16, 1)])
@pytest.mark.parametrize("seed,dim,keeps", [(1, 300, (-1, 288, 256, 200, 64)), (3, 768, (-1, 640, 512)),
                                            (4, 100, (96, 32))])
def test_prefilter_columns_keep_the_match_set(seed, dim, keeps, bits, diag):
    """Dropping embedding columns from the tensor-core operands widens every window's threshold by
    |f_drop||s_drop| (measured per window): the float64-decided match set must not change."""
    table, sx, fx, script, tok, off = _case(seed, dim=dim)
    want, _ = NumpyIndex(table, script, extra=sx).search_host(tok, off, fx)
    idx = _device_index(table, script, extra=sx, bits=bits)
    idx.set_option(nt.FS_OPT_DIAG, diag)
    assert idx.kept_dims == dim
    for keep in keeps:
        idx.set_option(nt.FS_OPT_PREFILTER_DIMS, keep)
        assert idx.kept_dims == (keep if keep > 0 else idx.kept_dims) and idx.kept_dims <= dim
        if keep == -1:
            assert idx.kept_dims < dim or dim <= 128
            assert 0.75 < idx.info(14) / 1e6 <= 1.0
        got, gc = idx.search_host(tok, off, fx)
        assert _pairs(got) == _pairs(want) and len(got) == len(want) and len(want) > 0, keep
        assert gc[nt.FS_CNT_CANDIDATES] >= len(want)
    idx.set_option(nt.FS_OPT_PREFILTER_DIMS, 0)
    assert idx.kept_dims == dim
    got, _ = idx.search_host(tok, off, fx)
    assert _pairs(got) == _pairs(want)
    idx.close()


def test_prefilter_columns_are_the_high_energy_ones_and_bounds_measure_the_rest():
    import torch
    table, sx, fx, script, tok, off = _case(5)
    rng = np.random.default_rng(3)
    col_scale = rng.permutation(np.linspace(0.2, 3.0, table.shape[1])).astype(np.float32)
    table = table * col_scale                        # clearly distinct column energies
    idx = _f8_index(table, script, sx)
    idx.set_option(nt.FS_OPT_PREFILTER_DIMS, 192)
    assert idx.kept_dims == 192 and idx.dim_pad == 192
    energy = (np.concatenate([table, sx]).astype(np.float64) ** 2).sum(axis=0)
    perm = np.argsort(-energy, kind="stable")
    np.testing.assert_allclose(idx.info(14) / 1e6, energy[perm[:192]].sum() / energy.sum(), atol=2e-6)
    tok_t, off_t, fx_t = idx.to_device(tok, off, fx)
    emb, thr = idx.stage_embed(tok_t, off_t, fx_t)
    torch.cuda.synchronize()
    allrows = np.concatenate([table, sx, fx], axis=0)
    scale = np.float32(idx.scale)
    kept = allrows[:, perm[:192]] * scale
    q = _e4m3(kept[tok])
    assert torch.equal(emb.cpu(), q.view(torch.uint8))
    x = allrows.astype(np.float64) * float(scale)
    back = _e4m3(kept).float().numpy().astype(np.float64)
    sq = (x ** 2).sum(axis=1)
    er = ((x[:, perm[:192]] - back) ** 2).sum(axis=1)
    dr = (x[:, perm[192:]] ** 2).sum(axis=1)
    thr = thr.cpu().numpy()
    assert thr.shape[1] == 4
    for a, b in zip(off[:-1], off[1:]):
        for i in range(int(a), int(b) - 5):
            ids = tok[i:i + 6]
            np.testing.assert_allclose(thr[i, 0], np.sqrt(sq[ids].sum()), rtol=3e-6)
            np.testing.assert_allclose(thr[i, 1], np.sqrt(er[ids].sum()), rtol=3e-4, atol=1e-6)
            np.testing.assert_allclose(thr[i, 2], np.sqrt(dr[ids].sum()), rtol=3e-5, atol=1e-6)
        for i in range(max(int(a), int(b) - 5), int(b)):
            assert np.all(np.isnan(thr[i, :3]))
    want, _ = NumpyIndex(table, script, extra=sx).search_host(tok, off, fx)
    got, _ = idx.search_host(tok, off, fx)
    assert _pairs(got) == _pairs(want)
    idx.close()


def test_fused_gather_equals_the_materialised_fan_matrix():
    """FS_OPT_FUSED_GATHER: the default kernel fetches its fan rows from the operand-row table by token id
    (TMA tile::gather4) -- candidates, matches and counters are those of the run that writes the fan
    operand matrix first; batches with extra rows, ragged works and a batch after a larger one (stale rows
    in the table's tail) included."""
    import torch
    table, sx, fx, script, tok, off = _case(31, works=(700, 3, 0, 6, 2397, 150, 40))
    table2, _, fx2, _, tok2, off2 = _case(32, works=(64, 64), n_extra_f=2)
    no_extra = np.where(tok < table.shape[0] + len(sx), tok, 0).astype(np.int32)
    idx = _device_index(table, script, extra=sx, bits=None)      # library defaults: fp8, E = 6, 128-column kernel
    assert idx.info(15) == 1
    runs = {}
    for fused in (1, 0, 1):
        idx.set_option(nt.FS_OPT_FUSED_GATHER, fused)
        got = [idx.search_host(tok, off, fx)]
        assert idx.info(16) == fused
        got.append(idx.search_host(tok2, off2, fx2))
        got.append(idx.search_host(no_extra, off))
        tok_t, off_t, fx_t = idx.to_device(tok, off, fx)
        cand, cnt = idx.stage_candidates(tok_t, off_t, fx_t)
        torch.cuda.synchronize()
        pairs = set(map(tuple, cand.cpu().numpy()[:int(cnt.cpu()[nt.FS_CNT_CANDIDATES])].tolist()))
        if fused in runs:                                        # a path is reproducible
            assert runs[('cand', fused)] == pairs
        runs[fused], runs[('cand', fused)] = got, pairs
    for (m0, c0), (m1, c1) in zip(runs[0], runs[1]):
        assert np.array_equal(np.sort(m0, order=['fan_pos', 'script_pos']), np.sort(m1, order=['fan_pos', 'script_pos']))
        assert np.array_equal(c0, c1)
    assert runs[('cand', 0)] == runs[('cand', 1)] and len(runs[('cand', 0)]) > 0
    assert len(runs[0][0][0]) > 0
    # other configurations fall back to the materialised matrix by themselves
    idx.set_option(nt.FS_OPT_DIAG, 3)
    idx.search_host(tok, off, fx)
    assert idx.info(16) == 0
    idx.close()


def test_fused_gather_on_ragged_shapes_around_the_tile_steps():
    """Batches whose token counts straddle the fan-tile step (108 rows, pairs of tiles: 216) and the
    gather granularity (4 rows), with and without extra rows: fused and materialised runs agree."""
    table, sx, fx, script, _, _ = _case(34, works=(50,))
    rng = np.random.default_rng(34)
    n_ids = table.shape[0] + len(sx) + len(fx)
    idx = _device_index(table, script, extra=sx, bits=None)
    assert idx.info(15) == 1
    total = 0
    for n_tok in (6, 7, 105, 107, 108, 109, 111, 113, 215, 216, 217, 323, 324, 325, 431, 432, 433, 1000, 2701):
        cuts = np.sort(rng.choice(np.arange(1, n_tok), size=min(3, n_tok - 1), replace=False)) if n_tok > 6 else []
        off = np.concatenate([[0], cuts, [n_tok]]).astype(np.int64)
        tok = rng.integers(0, n_ids, n_tok).astype(np.int32)
        src = int(rng.integers(0, len(script) - 40))
        ln = min(n_tok, 30)
        tok[:ln] = script[src:src + ln]                      # some reuse in every batch
        res = []
        for fused in (1, 0):
            idx.set_option(nt.FS_OPT_FUSED_GATHER, fused)
            m, c = idx.search_host(tok, off, fx)
            assert idx.info(16) == fused
            res.append((np.sort(m, order=['fan_pos', 'script_pos']), c))
        assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]), n_tok
        total += len(res[0][0])
    assert total > 50
    idx.close()


def test_fused_gather_falls_back_when_a_batch_has_too_many_extra_rows():
    """More out-of-vocabulary rows than the table's tail holds (65536): the batch is searched through the
    materialised fan matrix, with the same result as a batch that fits."""
    table, sx, _, script, tok, off = _case(33, dim=64, works=(300, 200), n_extra_f=0)
    rng = np.random.default_rng(5)
    n_fixed = table.shape[0] + len(sx)
    few = np.zeros((8, 64), np.float32)
    few[np.arange(8), rng.integers(0, 64, 8)] = 1.0
    many = np.zeros((70000, 64), np.float32)
    many[:8] = few
    tok = tok.copy()
    tok[10:18] = n_fixed + np.arange(8)                          # both batches use the same eight extra rows
    idx = _device_index(table, script, extra=sx, bits=None)
    if idx.info(15) != 1:
        pytest.skip("the 128-column kernel does not run this configuration")
    a, ca = idx.search_host(tok, off, few)
    assert idx.info(16) == 1
    b, cb = idx.search_host(tok, off, many)
    assert idx.info(16) == 0
    assert np.array_equal(np.sort(a, order=['fan_pos', 'script_pos']), np.sort(b, order=['fan_pos', 'script_pos']))
    assert np.array_equal(ca, cb)
    idx.close()


def test_two_batches_in_flight_equal_the_blocking_call():
    """fs_search_submit / fs_search_collect: two clusters queued back to back, collected in either
    order, give the matches and counters of fs_search_csr_host; a third submit is refused."""
    table, sx, fx, script, tok, off = _case(12)
    table2, _, fx2, _, tok2, off2 = _case(13)
    idx = _device_index(table, script, extra=sx, bits=None)
    want_a, ca = idx.search_host(tok, off, fx)
    want_b, cb = idx.search_host(tok2, off2, fx2)
    for order in ((0, 1), (1, 0)):
        tickets = [idx.search_submit(tok, off, fx), idx.search_submit(tok2, off2, fx2)]
        with pytest.raises(nt.NativeError):
            idx.search_submit(tok, off, fx)
        res = {}
        for k in order:
            res[k] = idx.search_collect(tickets[k])
        for k, (want, cw) in enumerate(((want_a, ca), (want_b, cb))):
            got, cg = res[k]
            assert np.array_equal(np.sort(got, order=['fan_pos', 'script_pos']),
                                  np.sort(want, order=['fan_pos', 'script_pos']))
            assert np.array_equal(cg, cw)
        with pytest.raises(nt.NativeError):
            idx.search_collect(tickets[0])          # a ticket is collected once
    # overflow of a submitted batch is reported at collect and retried by the wrapper
    idx2 = _device_index(table, script, extra=sx, bits=None)
    idx2.reserve(len(tok), 4)
    t = idx2.search_submit(tok, off, fx, cap=3)
    got, cg = idx2.search_collect(t)
    assert _pairs(got) == _pairs(want_a)
    idx.close()
    idx2.close()


def test_device_entry_point_flags_overflow_on_the_device():
    import torch
    table, sx, fx, script, tok, off = _case(12)
    idx = _device_index(table, script, extra=sx)
    host, hc = idx.search_host(tok, off, fx)
    assert hc[nt.FS_CNT_OVERFLOW] == 0 and len(host) > 4
    tok_t, off_t, fx_t = idx.to_device(tok, off, fx)
    out_t = torch.empty(24 * 2, dtype=torch.uint8, device="cuda")             # room for 2 matches only
    cnt_t = torch.zeros(nt.FS_CNT_COUNT, dtype=torch.int64, device="cuda")
    idx.search_dev(tok_t, off_t, fx_t, out_t, cnt_t)
    torch.cuda.synchronize()
    cnt = cnt_t.cpu().numpy()
    assert cnt[nt.FS_CNT_OVERFLOW] == nt.FS_OVERFLOW_MATCHES and cnt[nt.FS_CNT_MATCHES] == len(host)
    idx2 = _device_index(table, script, extra=sx)
    idx2.reserve(len(tok), 4)                                                 # candidate buffer of 4 pairs
    out_t = torch.empty(24 * 4096, dtype=torch.uint8, device="cuda")
    idx2.search_dev(tok_t, off_t, fx_t, out_t, cnt_t)
    torch.cuda.synchronize()
    assert int(cnt_t.cpu()[nt.FS_CNT_OVERFLOW]) & nt.FS_OVERFLOW_CANDIDATES
    idx.close()
    idx2.close()


def test_indexes_on_two_devices_in_one_process():
    """The distance kernel needs > 48 KB of dynamic shared memory, an attribute of the (function,
    device) pair: a second index on another GPU of the same process must work too."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    table, sx, fx, script, tok, off = _case(12)
    want, _ = NumpyIndex(table, script, extra=sx).search_host(tok, off, fx)
    for dev in (0, 1, 0):
        idx = _device_index(table, script, extra=sx, bits=None, device=dev)
        got, _ = idx.search_host(tok, off, fx)
        assert _pairs(got) == _pairs(want)
        idx.close()
