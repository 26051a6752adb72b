"""Per-stage timing of the HBM-bound kernels: token gather (+ window norms) and exact hash-join.

    python tools/stage_bench.py [--tokens 20000000]

CUDA events on the launching stream, 3 warm-up + 10 timed launches, inputs larger than L2.
Algorithmic bytes: gather 4 B read + d_pad*2 B written per token; hash-join 4 B
(int32 row id; the survey's 8 B assumed 64-bit keys) read per fan window."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fandom_search_b200 import _native as nt
from fandom_search_b200 import synth
from fandom_search_b200.engine import DeviceIndex


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=20_000_000)
    ap.add_argument("--dim", type=int, default=300)
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    lex = synth.SynthLexicon(vocab=50000, dim=args.dim, oov_frac=0.0, seed=1001)
    script = synth.make_script_tokens(lex, 25000).astype(np.int32)
    idx = DeviceIndex(lex.table_all, script)
    n_works = args.tokens // 5000
    rng = np.random.default_rng(1)
    # Zipf-distributed ids like the real workload, with verbatim script quotes planted
    tok = lex.sample_words(rng, args.tokens).astype(np.int32)
    for w in range(0, n_works, 3):
        src = int(rng.integers(0, 25000 - 30))
        tok[w * 5000 + 100:w * 5000 + 120] = script[src:src + 20]
    off = (np.arange(n_works + 1, dtype=np.int64) * 5000)
    off[-1] = args.tokens
    tok_t, off_t, _ = idx.to_device(tok, off)
    idx.reserve(args.tokens, 1 << 20)
    res = {"tokens": args.tokens, "dim": args.dim, "dim_pad": idx.dim_pad, "operand_bits": idx.operand_bits,
           "hbm_peak_gbs": peaks["hbm_gbs"]}

    ms = timeit(lambda: idx.stage_embed(tok_t, off_t))
    row_bytes = idx.dim_pad * idx.operand_bits // 8        # fp8: 1 B per element, fp16: 2
    bytes_alg = args.tokens * (4 + row_bytes)
    res["embed"] = {"ms": ms, "algorithmic_gbs": bytes_alg / ms / 1e6, "frac_of_hbm_peak": bytes_alg / ms / 1e6 / peaks["hbm_gbs"],
                    "tokens_per_s": args.tokens / ms * 1e3, "note": "gather + window norms + torch.empty of outputs"}

    # write-only reference: how fast can this GPU stream zeros into a buffer of the same size?
    buf = torch.empty(args.tokens * row_bytes, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: buf.zero_())
    res["memset_same_bytes"] = {"ms": ms, "gbs": buf.numel() / ms / 1e6}
    del buf
    out_t = torch.empty((1 << 22, 2), dtype=torch.int32, device="cuda")
    cnt_t = torch.zeros(nt.FS_CNT_COUNT, dtype=torch.int64, device="cuda")
    ms = timeit(lambda: idx.exact_join_dev(tok_t, off_t, out_t, cnt_t))
    n_pairs = int(cnt_t.cpu()[nt.FS_CNT_EXACT])
    windows = int(np.maximum(np.diff(off) - 5, 0).sum())
    res["hash_join"] = {"ms": ms, "algorithmic_gbs": windows * 4 / ms / 1e6, "frac_of_hbm_peak": windows * 4 / ms / 1e6 / peaks["hbm_gbs"],
                        "windows_per_s": windows / ms * 1e3, "pairs": n_pairs}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
