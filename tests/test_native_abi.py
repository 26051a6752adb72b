"""CPU checks of the C-ABI library: it loads, exports every symbol include/fandom_search.h
declares, its host-side helpers agree with the oracle's, and the GPU entry points fail
loudly (no CPU fallback) when no device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

from fandom_search_b200 import _native as nt
from fandom_search_b200 import text
from oracle import reference_search as ora

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "fandom_search.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = nt.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libfandom_search.so does not export %s" % n
    assert set(names) == set(nt.SIGNATURES), set(names) ^ set(nt.SIGNATURES)
    assert lib.fs_abi_version() == 2


def test_struct_layout_matches_header():
    assert nt.MATCH_DTYPE.itemsize == 24
    assert [nt.MATCH_DTYPE.fields[k][1] for k in ("fan_pos", "script_pos", "distance", "work", "flags")] == [0, 4, 8, 16, 20]
    assert nt.PAIR_DTYPE.itemsize == 8


def test_host_helpers_agree_with_oracle_shims():
    rng = np.random.default_rng(0)
    alphabet = "abcdefg ,[]XYZé日😀"
    for _ in range(300):
        a = "".join(rng.choice(list(alphabet), rng.integers(0, 40)))
        b = "".join(rng.choice(list(alphabet), rng.integers(0, 40)))
        assert text.levenshtein(a, b) == ora.levenshtein(a, b)
        assert text.string_id(a) == ora.murmurhash64a(a.encode("utf-8"), 1)
    assert text.string_id("coffee") == 3197928453018144401
    assert text.levenshtein("a b c d e f", "[a, b, c, d, e, f]") == 7
    s = "  the quick\tbrown\n\nfox  jumps\r\nover\x0bthe\x0clazy dog "
    assert text.tokenize(s) == ora.tokenize(s) == s.split()
    assert text.tokenize("") == [] and text.tokenize("   ") == []
    assert text.tokenize("naïve café 日本語 x") == ["naïve", "café", "日本語", "x"]


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fandom_search_b200.engine import DeviceIndex
    table = np.zeros((4, 8), np.float32)
    with pytest.raises(nt.NativeError) as e:
        DeviceIndex(table, np.array([0, 1, 2, 3, 0, 1, 2], np.int32))
    assert e.value.status == nt.FS_E_NODEVICE and "no CPU fallback" in str(e.value)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fandom_search_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "numpy_index" not in src, f
