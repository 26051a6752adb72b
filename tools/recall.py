"""Recall of the reference's approximate index against the exhaustive GPU search.

    python tools/recall.py [--works 100] [--seeds 5]

The reference finds a (fan window, script window) pair only if the two windows share a bucket in
one of 15 random-hyperplane tables of 14 bits (search.py:112-116,178); the GPU path compares
every pair.  For several hyperplane seeds this tool runs the exhaustive search once with the
LSH-emulation flags switched on (csrc/lsh.cu) and reports, per distance bin, the fraction of
exhaustive matches a reference run with that seed would have seen, next to the closed form
1 - (1 - (1 - theta/pi)^14)^15, theta = arccos(1 - distance)  (SURVEY 7.2a)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from fandom_search_b200 import _native as nt
from fandom_search_b200 import synth
from fandom_search_b200.engine import DeviceIndex
from fandom_search_b200.lsh import LshEmulation


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--works", type=int, default=100)
    ap.add_argument("--seeds", type=int, default=5)
    ap.add_argument("--script-tokens", type=int, default=10000)
    args = ap.parse_args()
    lex = synth.SynthLexicon(vocab=50000, dim=300, oov_frac=0.0, seed=1001)
    script = synth.make_script_tokens(lex, args.script_tokens).astype(np.int32)
    words, off = synth.synth_csr_batch(lex, script, range(args.works), spans_mean=12.0)
    tok = words.astype(np.int32)
    idx = DeviceIndex(lex.table_all, script)
    bins = [0.0, 1e-9, 0.02, 0.04, 0.06, 0.08, 0.1]
    rows = []
    for seed in range(args.seeds):
        LshEmulation(15, 14, 6 * 300, seed).install(idx)
        m, cnt = idx.search_host(tok, off)
        seen = ((m["flags"] >> nt.FS_MATCH_LSH_SHIFT) & 0xFF) > 0
        d = np.maximum(m["distance"], 0.0)
        rows.append([(int(((d >= lo) & (d < hi)).sum()), int((seen & (d >= lo) & (d < hi)).sum()))
                     for lo, hi in zip(bins[:-1], bins[1:])])
        total, found = len(m), int(seen.sum())
        print(json.dumps({"seed": seed, "exhaustive_pairs": total, "seen_by_lsh": found,
                          "recall": found / max(total, 1)}), flush=True)
    rows = np.array(rows)
    for b, (lo, hi) in enumerate(zip(bins[:-1], bins[1:])):
        n, f = rows[:, b, 0].sum(), rows[:, b, 1].sum()
        mid = 0.5 * (lo + hi) if lo > 0 else 0.0
        theta = np.arccos(1.0 - mid)
        theory = 1.0 - (1.0 - (1.0 - theta / np.pi) ** 14) ** 15
        print(json.dumps({"distance_bin": [lo, hi], "pairs": int(n), "measured_recall": f / max(n, 1),
                          "closed_form_at_bin_centre": theory}), flush=True)


if __name__ == "__main__":
    main()
