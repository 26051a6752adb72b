"""clock64 timeline of CTA 0 of the distance kernel (debug build only).

    FS_NVCC_EXTRA=-DFS_TIMELINE python -m fandom_search_b200.build --force
    python tools/timeline.py > gpurun_out/timeline.txt        # on the GPU box

Rows: role tile slot... ; roles 0 = TMA producer (slot 0 loop top, 1+2c stage of chunk c free,
2+2c its TMA issued), 1 = MMA issuer (0 loop top, 1 accumulator stage free, 2+2c chunk c landed,
3+2c chunk c issued and committed), 2+w = epilogue warp w (0 loop top, 1 accumulator full,
2 all columns read/tested, 3 stage released).  Stamps are relative to the issuer's first."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fandom_search_b200 import _native as nt
from fandom_search_b200.engine import DeviceIndex

ROLES, TILES, SLOTS = 18, 64, 12


def main():
    rng = np.random.default_rng(0)
    vocab, d, ns, nf = 50000, 300, 25000, 2_500_000
    table = rng.standard_normal((vocab, d), dtype=np.float32)
    script = rng.integers(0, vocab, ns + 5).astype(np.int32)
    idx = DeviceIndex(table, script, window=6, threshold=0.1)
    c2 = "--c2" in sys.argv[1:]
    if c2:
        # the bench.py workload (Zipf text with planted reuse) instead of uniform random tokens
        from fandom_search_b200 import synth
        idx.close()
        lex = synth.SynthLexicon(vocab=50000, dim=300, oov_frac=0.0, seed=1001)
        script = synth.make_script_tokens(lex, 25000).astype(np.int32)
        idx = DeviceIndex(lex.table_all, script, window=6, threshold=0.1)
    for a in sys.argv[1:]:
        if a.startswith("--"):
            continue
        k, v = a.split("=")
        idx.set_option(getattr(nt, k), int(v))
    if c2:
        words, off = synth.synth_csr_batch(lex, script, range(500))
        tok_t = torch.from_numpy(words.astype(np.int32)).cuda()
        off_t = torch.from_numpy(off).cuda()
    else:
        n_works = nf // 5000
        lens = np.full(n_works, 5006, dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        g = torch.Generator(device="cuda").manual_seed(3000)
        tok_t = torch.randint(0, vocab, (int(off[-1]),), generator=g, device="cuda", dtype=torch.int32)
        off_t = torch.from_numpy(off).cuda()
    out_t = torch.empty(24 << 16, dtype=torch.uint8, device="cuda")
    cnt_t = torch.zeros(nt.FS_CNT_COUNT, dtype=torch.int64, device="cuda")
    idx.reserve(int(off[-1]), 1 << 16)
    for _ in range(12 if "--hot" in sys.argv[1:] else 3):      # --hot: more launches first (clocks settle under the cap)
        idx.search_dev(tok_t, off_t, None, out_t, cnt_t)
    torch.cuda.synchronize()
    idx.timing_reset()
    idx.search_dev(tok_t, off_t, None, out_t, cnt_t)
    torch.cuda.synchronize()
    ms, n = idx.timing_read()
    lib = nt.load()
    buf = np.zeros((ROLES, TILES, SLOTS), dtype=np.int64)
    rc = lib.fs_debug_timeline(buf.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    t0 = buf[1, 0, 0]
    print("# kernel_ms %.3f" % (ms / n))
    for r in range(ROLES):
        for t in range(8, 40):
            print(r, t, " ".join(str(int(x - t0)) if x else "-" for x in buf[r, t]))


if __name__ == "__main__":
    main()
