"""Per-script-word reuse histogram on the device.

`ao3.py format` re-reads the (possibly multi-GB) match CSV and counts, for every
ORIGINAL_SCRIPT_WORD_INDEX, the rows whose BEST_COMBINED_DISTANCE is <= 0 ("exact matches") and
<= 0.05, 0.1, ..., 0.5 (ao3.py:351-363, 407-411: boolean columns summed by a pandas group-by,
re-indexed over every script word with 0).  ReuseHistogram accumulates the same table while the
search runs: attached to a DeviceIndex it is fed by the device itself from the winning rows of
every cluster (fs_index_set_reuse_histogram -- the rows are counted before they are copied out, no
CSV round trip); clusters whose records were made on the host are added from the host arrays
(`add`).  Under torchrun the per-rank tables are summed with ONE all_reduce of the [n_words, 11]
int64 table."""
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
        self._index = None

    def attach(self, device_index):
        """From now on the device counts the winning rows of every search_submit_rows batch of
        `device_index` into this table."""
        if device_index.n_script_tok != self.n_words:
            raise ValueError("one histogram row per script word is expected")
        self._torch.cuda.synchronize(self.device)
        nt.check(device_index._lib.fs_index_set_reuse_histogram(
            device_index._h, nt.ptr(self.counts), nt.ptr(self._thr), len(self.thresholds)))
        self._index = device_index

    def detach(self):
        if self._index is not None and self._index._h is not None:
            nt.check(self._index._lib.fs_index_set_reuse_histogram(self._index._h, None, None, 0))
        self._index = None

    def add(self, word_ix, combined):
        """word_ix int32 [n] (ORIGINAL_SCRIPT_WORD_INDEX), combined float64 [n] -- host arrays."""
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
        torch.cuda.current_stream(self.device).synchronize()     # w, c may be released now

    def add_best(self, best):
        """Winning records as arrays (fs_records_best output / DeviceRows.best)."""
        self.add(best['match_ix'].astype(np.int64) + best['window_ix'],
                 best['distance'] * best['lev'].astype(np.float64))

    def add_records(self, records):
        if records:
            self.add([r[4] for r in records], [r[11] for r in records])

    def all_reduce(self):
        """Sum the tables of all ranks (one collective; a no-op in a single process)."""
        import torch.distributed as dist
        self._torch.cuda.synchronize(self.device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.counts, op=dist.ReduceOp.SUM)
        return self

    def result(self):
        self._torch.cuda.synchronize(self.device)
        return self.counts.cpu().numpy()

    def write_csv(self, path, script_words):
        """The count columns of `ao3.py format`'s output (ao3.py:407-428): one row per script word,
        index ORIGINAL_SCRIPT_WORD_INDEX, the eleven 'Frequency of Reuse' columns, and the word."""
        import csv
        table = self.result()
        with open(path, 'w', encoding='utf-8', newline='') as out:
            wr = csv.writer(out)
            wr.writerow(['ORIGINAL_SCRIPT_WORD_INDEX'] + COLUMN_NAMES + ['ORIGINAL_SCRIPT_WORD'])
            for i, row in enumerate(table.tolist()):
                wr.writerow([i] + row + [script_words[i]])
