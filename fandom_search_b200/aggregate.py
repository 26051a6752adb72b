"""Per-script-word reuse histogram on the device.

`ao3.py format` re-reads the (possibly multi-GB) match CSV and counts, for every
ORIGINAL_SCRIPT_WORD_INDEX, the rows whose BEST_COMBINED_DISTANCE is <= 0 ("exact matches") and
<= 0.05, 0.1, ..., 0.5 (ao3.py:351-363, 407-411).  ReuseHistogram accumulates the same table
cluster by cluster from the winning records, on the GPU, so the aggregate can be produced
without the CSV round-trip."""
import ctypes

import numpy as np

from . import _native as nt

# ao3.py:353-363: exact matches (<= 0) then 0.05 .. 0.5
THRESHOLDS = [0.0, 0.05, 0.1, 0.15, 0.2, 0.25, 0.3, 0.35, 0.4, 0.45, 0.5]
COLUMN_NAMES = (['Frequency of Reuse (Exact Matches)'] +
                ['Frequency of Reuse (0-{})'.format(str(t)) for t in THRESHOLDS[1:]])


class ReuseHistogram:
    def __init__(self, n_script_words, device=0, thresholds=THRESHOLDS):
        import torch
        self._torch = torch
        self.device = torch.device("cuda", device)
        self.n_words = int(n_script_words)
        self.thresholds = list(thresholds)
        self._thr = torch.tensor(self.thresholds, dtype=torch.float64, device=self.device)
        self.counts = torch.zeros((self.n_words, len(self.thresholds)), dtype=torch.int64,
                                  device=self.device)

    def add(self, word_ix, combined):
        """word_ix int32 [n] (ORIGINAL_SCRIPT_WORD_INDEX), combined float64 [n]."""
        torch = self._torch
        word_ix = np.ascontiguousarray(word_ix, dtype=np.int32)
        combined = np.ascontiguousarray(combined, dtype=np.float64)
        if len(word_ix) == 0:
            return
        w = torch.from_numpy(word_ix).to(self.device)
        c = torch.from_numpy(combined).to(self.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        nt.check(nt.load().fs_reuse_histogram_dev(stream, nt.ptr(w), nt.ptr(c), len(word_ix),
                                                  nt.ptr(self._thr), len(self.thresholds),
                                                  self.n_words, nt.ptr(self.counts)))

    def add_records(self, records):
        if records:
            self.add([r[4] for r in records], [r[11] for r in records])

    def result(self):
        return self.counts.cpu().numpy()
