"""TEST INFRASTRUCTURE: expected outputs of the configuration-size parity cases
(tests/config_cases.py), computed by the CPU oracle and committed under tests/golden/config/.

    python oracle/make_config_golden.py [case ...]       # default: every case

Per case it writes
  <case>.npz      input digest; every (work, fan window, script window, float64 distance) under
                  the threshold BEFORE the top-10 cut (the full match set of search.py:176-184 with
                  exhaustive candidates)
  <case>.csv.gz   the CSV rows of search.py:188-226 in aggregate order (file name = basename)

The oracle runs in float64 on the host cores (minutes for C1-whole); the GPU box replays the
inputs from the same seeds and compares with these files (tests/test_gpu_config_parity.py).
"""
import csv
import gzip
import io
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from fandom_search_b200.lexicon import py_hash_seed0  # noqa: E402
from oracle import reference_search as ora            # noqa: E402
from tests import config_cases                        # noqa: E402


def run_case(name):
    case = config_cases.CASES[name]()
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory(prefix="fs_cfg_") as tmp:
        lex_path, script_path, files = case.write(tmp)
        index = ora.OracleIndex(script_path, ora.OracleLexicon(lex_path, oov_hash=py_hash_seed0),
                                mode="exhaustive", engine="dense")
        pw, pf, ps, pd = [], [], [], []
        rows = []
        for k, fn in enumerate(files):
            recs = index.search(fn)                      # search.py:163-226 for this work
            for fan_ix, match_ix, dist in index.last_all_pairs:
                pw.append(k)
                pf.append(fan_ix)
                ps.append(match_ix)
                pd.append(dist)
            for r in recs:
                rows.append([os.path.basename(r[0])] + list(r[1:]))
            if (k + 1) % 50 == 0:
                print("  %s: %d/%d works, %.0f s" % (name, k + 1, len(files), time.perf_counter() - t0), flush=True)
    os.makedirs(config_cases.GOLDEN_CONFIG, exist_ok=True)
    np.savez_compressed(os.path.join(config_cases.GOLDEN_CONFIG, name + ".npz"),
                        digest=np.array(case.digest()), work=np.array(pw, np.int32), fan=np.array(pf, np.int32),
                        script=np.array(ps, np.int32), distance=np.array(pd, np.float64),
                        n_script_windows=np.array(index.n_script_windows),
                        windows=np.array(index.windows_processed))
    buf = io.StringIO(newline='')
    csv.writer(buf).writerows(rows)
    with gzip.GzipFile(os.path.join(config_cases.GOLDEN_CONFIG, name + ".csv.gz"), "wb", mtime=0) as f:
        f.write(buf.getvalue().encode("utf-8"))
    print("%s: %d works, %d windows, %d pairs under the threshold, %d rows, %.0f s"
          % (name, len(files), index.windows_processed, len(pw), len(rows), time.perf_counter() - t0))


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(config_cases.CASES)):
        run_case(name)
