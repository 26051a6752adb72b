"""TEST INFRASTRUCTURE: float64 numpy statement of the device search on embedding-row ids.

Same call shape as fandom_search_b200.engine.DeviceIndex.search_host / exact_join_host, so it
serves (a) as the checker for the GPU parity tests and (b) as a stand-in device for the CPU-only
tests of the host logic (tokenise -> CSR -> records -> CSV, multi-rank sharding).  It follows
/root/reference search.py:94-95,123,170-184 (windows, unit vectors, 1 - dot, < threshold)."""
import numpy as np

from fandom_search_b200 import _native as nt


class NumpyIndex:
    def __init__(self, table, script_tok, script_off=None, extra=None, window=6, threshold=0.1,
                 device=0):
        self.table = np.asarray(table, dtype=np.float32)
        self.dim = self.table.shape[1]
        self.n_base = self.table.shape[0]
        self.script_extra = (np.zeros((0, self.dim), np.float32) if extra is None
                             else np.asarray(extra, np.float32).reshape(-1, self.dim))
        self.n_script_extra = self.script_extra.shape[0]
        self.script_tok = np.asarray(script_tok, dtype=np.int64)
        self.script_off = (np.array([0, len(self.script_tok)], np.int64) if script_off is None
                           else np.asarray(script_off, np.int64))
        self.window = window
        self.threshold = threshold
        self.sw, self.spos = self._windows(self.script_tok, self.script_off, None)
        self.n_script_windows = len(self.spos)
        self.sm_count = 0

    def _rows(self, ids, extra):
        parts = [self.table, self.script_extra]
        if extra is not None:
            parts.append(np.asarray(extra, np.float32).reshape(-1, self.dim))
        allrows = np.concatenate(parts, axis=0)
        ids = np.asarray(ids, dtype=np.int64)
        ok = (ids >= 0) & (ids < allrows.shape[0])
        out = np.zeros((len(ids), self.dim), np.float64)
        out[ok] = allrows[ids[ok]]
        return out

    def _windows(self, tok, off, extra):
        w = self.window
        vec = self._rows(tok, extra)
        pos = [i for a, b in zip(off[:-1], off[1:]) for i in range(int(a), int(b) - w + 1)]
        pos = np.array(pos, dtype=np.int64)
        if len(pos) == 0:
            return np.zeros((0, w * self.dim)), pos
        win = np.stack([vec[i:i + w].ravel() for i in pos])
        nrm = np.sqrt((win * win).sum(axis=1))
        win = win / np.where(nrm > 0, nrm, 1.0)[:, None]
        return win, pos

    def distances(self, tok, off, extra=None):
        fw, fpos = self._windows(np.asarray(tok, np.int64), np.asarray(off, np.int64), extra)
        if len(fpos) == 0 or len(self.spos) == 0:
            return np.zeros((len(fpos), len(self.spos))), fpos
        return 1.0 - fw @ self.sw.T, fpos

    def search_host(self, tok, off, extra=None, cap=None, out=None):
        tok = np.asarray(tok, np.int64)
        off = np.asarray(off, np.int64)
        d, fpos = self.distances(tok, off, extra)
        ii, jj = np.nonzero(d < self.threshold)
        m = np.zeros(len(ii), dtype=nt.MATCH_DTYPE)
        m['fan_pos'] = fpos[ii]
        m['script_pos'] = self.spos[jj]
        m['distance'] = d[ii, jj]
        m['work'] = np.searchsorted(off, fpos[ii], side='right') - 1
        same = np.array([np.array_equal(tok[a:a + self.window], self.script_tok[b:b + self.window])
                         for a, b in zip(m['fan_pos'], m['script_pos'])], dtype=bool)
        m['flags'] = same.astype(np.uint32) * nt.FS_MATCH_EXACT
        counters = np.zeros(nt.FS_CNT_COUNT, np.int64)
        counters[nt.FS_CNT_MATCHES] = len(m)
        counters[nt.FS_CNT_CANDIDATES] = len(m)
        counters[nt.FS_CNT_WINDOWS] = len(fpos)
        return m, counters

    def search_submit(self, tok, off, extra=None, cap=None):
        return {'result': self.search_host(tok, off, extra)}

    def search_collect(self, ticket, out=None):
        return ticket['result']

    def exact_join_host(self, tok, off, cap=None):
        tok = np.asarray(tok, np.int64)
        off = np.asarray(off, np.int64)
        w = self.window
        table = {}
        for j in self.spos.tolist():
            table.setdefault(tuple(self.script_tok[j:j + w].tolist()), []).append(j)
        pairs = []
        for a, b in zip(off[:-1], off[1:]):
            for i in range(int(a), int(b) - w + 1):
                for j in table.get(tuple(tok[i:i + w].tolist()), ()):
                    pairs.append((i, j))
        out = np.zeros(len(pairs), dtype=nt.PAIR_DTYPE)
        if pairs:
            arr = np.array(pairs)
            out['fan_pos'], out['script_pos'] = arr[:, 0], arr[:, 1]
        counters = np.zeros(nt.FS_CNT_COUNT, np.int64)
        counters[nt.FS_CNT_EXACT] = len(pairs)
        return out, counters

    def close(self):
        pass
