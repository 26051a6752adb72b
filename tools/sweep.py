"""Kernel sweep (BASELINE.json configs[4]): fan windows x script windows x embedding dim.

    python tools/sweep.py [--quick] > gpurun_out/sweep.jsonl

Token ids are drawn directly (no text), inputs resident in HBM, the distance kernel is timed
with CUDA events on its launching stream (fs_timing_read); reports windows/s, TFLOP/s by the
nominal dense formula 2*(6*d)*Ns and by executed flops (d padded to a multiple of 16)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fandom_search_b200 import _native as nt
from fandom_search_b200.engine import DeviceIndex


def run_case(nf, ns, d, reps, rng, vocab=50000, works_len=5000, diag=1, pair=0, pack=None, clocks=False, bits=8, group=None,
             total_windows=None):
    """total_windows: search that many fan windows by re-launching over the SAME device buffer of nf
    windows (the 1e8 / 1e9 points of BASELINE.json configs[4]: a fixed device buffer, many launches)."""
    table = rng.standard_normal((vocab, d), dtype=np.float32)
    script = rng.integers(0, vocab, ns + 5).astype(np.int32)
    idx = DeviceIndex(table, script, window=6, threshold=0.1)
    if diag is None:
        # library defaults (what search.py runs): fp8, E = 6, CTA pairs, resident fan tile where it fits
        diag, pair, bits = idx.diag, 2 if idx.info(7) else idx.cta_pair, idx.operand_bits
    else:
        if pair == 0:
            bits = 16      # fp8 operands exist for CTA pairs only
        idx.set_option(nt.FS_OPT_OPERAND_BITS, bits)
        idx.set_option(nt.FS_OPT_DIAG, diag)
        idx.set_option(nt.FS_OPT_CTA_PAIR, 1 if pair else 0)
        idx.set_option(nt.FS_OPT_A_RESIDENT, 1 if pair == 2 else 0)
    if pack is not None:
        idx.set_option(nt.FS_OPT_PACKED_SHUFFLE, pack)
    if group is not None:
        idx.set_option(nt.FS_OPT_TILE_GROUP, group)
    n_works = max(1, nf // works_len)
    lens = np.full(n_works, (nf + 5 * n_works) // n_works + 1, dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    g = torch.Generator(device="cuda").manual_seed(3000)
    tok_t = torch.randint(0, vocab, (int(off[-1]),), generator=g, device="cuda", dtype=torch.int32)
    off_t = torch.from_numpy(off).cuda()
    out_t = torch.empty(24 << 16, dtype=torch.uint8, device="cuda")
    cnt_t = torch.zeros(nt.FS_CNT_COUNT, dtype=torch.int64, device="cuda")
    idx.reserve(int(off[-1]), 1 << 16)
    idx.search_dev(tok_t, off_t, None, out_t, cnt_t)   # warm-up
    idx.search_dev(tok_t, off_t, None, out_t, cnt_t)
    torch.cuda.synchronize()
    sampler = None
    if clocks:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from bench import ClockSampler
        sampler = ClockSampler(0)
        sampler.start()
    if total_windows:
        reps = max(1, int(round(total_windows / nf)))
    idx.timing_reset()
    ms, n = 0.0, 0
    done = 0
    while done < reps:       # (the library keeps the events of the last 256 launches)
        burst = min(200, reps - done)
        idx.timing_reset()
        for _ in range(burst):
            idx.search_dev(tok_t, off_t, None, out_t, cnt_t)
        torch.cuda.synchronize()
        b_ms, b_n = idx.timing_read()
        ms += b_ms
        n += b_n
        done += burst
    clk = sampler.stop() if sampler else None
    windows = int(cnt_t.cpu()[nt.FS_CNT_WINDOWS])
    per = ms / n * 1e-3
    m_step = 108 if diag == 6 else 129 - diag      # E = 6: overlapping lane quarters
    exec_factor = (6 // diag) * (128.0 * 256.0) / (m_step * (257 - diag))
    res = {"launches": n, "total_fan_windows": windows * n, "kept_dims": idx.kept_dims, "bits": bits, "candidates": int(cnt_t.cpu()[nt.FS_CNT_CANDIDATES]), "matches": int(cnt_t.cpu()[nt.FS_CNT_MATCHES]),
           "diag": diag, "pair": pair, "pack": pack, "group": group, "fan_windows": windows, "script_windows": idx.n_script_windows, "dim": d, "dim_pad": idx.dim_pad,
           "kernel_ms": ms / n, "windows_per_s": windows / per, "clocks": clk,
           "tflops_dense_nominal": 2.0 * 6 * d * idx.n_script_windows * windows / per / 1e12,
           "tflops_executed": 2.0 * exec_factor * idx.dim_pad * idx.n_script_windows * windows / per / 1e12}
    idx.close()
    del tok_t, out_t
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--defaults", action="store_true", help="the C5 grid with the library's default kernel")
    ap.add_argument("--diag", action="store_true", help="compare the diagonal-sum factors at C2 size")
    ap.add_argument("--bits", type=int, default=8, help="operand bits for --one (8 = fp8 e4m3, 16 = fp16)")
    ap.add_argument("--f8", action="store_true", help="fp8 e4m3 operands vs fp16 at C2 size, all diagonal factors")
    ap.add_argument("--one", type=int, nargs=4, metavar=("DIAG", "NF", "NS", "D"), help="run a single case")
    ap.add_argument("--pair", type=int, default=0)
    ap.add_argument("--pack", type=int, default=None)
    ap.add_argument("--group", type=int, default=None, help="FS_OPT_TILE_GROUP for --one (default: library default)")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--big", action="store_true",
                    help="the 1e8 and 1e9 fan-window points of the C5 grid, default kernel, through a fixed 1e7-window device buffer")
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    if args.one:
        diag, nf, ns, d = args.one
        print(json.dumps(run_case(nf, ns, d, args.reps, rng, diag=diag, pair=args.pair, pack=args.pack,
                                  bits=args.bits, group=args.group, clocks=True)), flush=True)
        return
    if args.big:
        for d in (300, 768):
            for ns in (1000, 10000, 100000):
                for total in (100_000_000, 1_000_000_000):
                    print(json.dumps(run_case(10_000_000, ns, d, 1, rng, diag=None, total_windows=total, clocks=True)), flush=True)
        return
    if args.defaults:
        for d in (300, 768):
            for ns in (1000, 10000, 100000):
                for nf in (100_000, 1_000_000, 10_000_000):
                    if nf * ns * d > 3.1e15:
                        continue
                    print(json.dumps(run_case(nf, ns, d, 3, rng, diag=None)), flush=True)
        return
    if args.f8:
        for bits, diag, d, pack, pair in ((16, 3, 300, 2, 1), (16, 3, 300, 2, 2), (8, 3, 300, 2, 1), (8, 3, 300, 2, 2),
                                          (8, 6, 300, 2, 1), (8, 6, 300, 2, 2), (8, 2, 300, 2, 1), (8, 6, 768, 2, 1),
                                          (8, 3, 768, 2, 1), (16, 6, 768, 2, 1), (8, 3, 300, 2, 2)):
            print(json.dumps(run_case(2_500_000, 25000, d, 20, rng, diag=diag, pair=pair, pack=pack,
                                      clocks=True, bits=bits)), flush=True)
        return
    if args.diag:
        for pair in (1, 2):
            for diag in (1, 2, 3, 6):
                for (nf, ns, d) in ((2_500_000, 25000, 300),) + (((2_500_000, 25000, 768),) if pair == 1 else ()):
                    print(json.dumps(run_case(nf, ns, d, 3, rng, diag=diag, pair=pair)), flush=True)
        return
    if args.quick:
        cases = [(2_500_000, 25000, 300, 3), (2_500_000, 25000, 320, 3), (2_500_000, 25000, 304, 3)]
    else:
        cases = []
        for d in (300, 768):
            for ns in (1000, 10000, 100000):
                for nf in (100_000, 1_000_000, 10_000_000):
                    if nf * ns * d > 3.1e15:
                        continue
                    cases.append((nf, ns, d, 2))
        cases += [(2_500_000, 25000, 300, 3), (2_500_000, 25000, 320, 3), (100_000_000, 1000, 300, 1)]
    for nf, ns, d, reps in cases:
        print(json.dumps(run_case(nf, ns, d, reps, rng)), flush=True)


if __name__ == "__main__":
    main()
