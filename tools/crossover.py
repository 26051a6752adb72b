"""E = 3 vs E = 6 on the C2 workload (Zipf tokens + planted reuse) for several embedding widths:
where the default diagonal factor should switch (python tools/crossover.py [dims...])."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fandom_search_b200 import _native as nt
from fandom_search_b200 import synth
from fandom_search_b200.engine import DeviceIndex


def main():
    dims = [int(a) for a in sys.argv[1:]] or [300, 384, 512, 640, 768]
    for d in dims:
        lex = synth.SynthLexicon(vocab=50000, dim=d, oov_frac=0.0, seed=1001)
        script = synth.make_script_tokens(lex, 25000).astype(np.int32)
        tok, off = synth.synth_csr_batch(lex, script, range(500))
        tok = tok.astype(np.int32)
        idx = DeviceIndex(lex.table_all, script)
        tok_t, off_t, _ = idx.to_device(tok, off)
        out_t = torch.empty(24 << 20, dtype=torch.uint8, device="cuda")
        cnt_t = torch.zeros(nt.FS_CNT_COUNT, dtype=torch.int64, device="cuda")
        idx.reserve(len(tok), 1 << 20)
        res = {"dim": d, "default_diag": idx.diag}
        for e in (3, 6):
            idx.set_option(nt.FS_OPT_DIAG, e)
            for _ in range(3):
                idx.search_dev(tok_t, off_t, None, out_t, cnt_t)
            torch.cuda.synchronize()
            idx.timing_reset()
            for _ in range(15):
                idx.search_dev(tok_t, off_t, None, out_t, cnt_t)
            torch.cuda.synchronize()
            ms, n = idx.timing_read()
            windows = int(cnt_t.cpu()[nt.FS_CNT_WINDOWS])
            res["E%d_Mwindows_per_s" % e] = round(windows / (ms / n) / 1e3, 2)
        print(json.dumps(res), flush=True)
        idx.close()


if __name__ == "__main__":
    main()
