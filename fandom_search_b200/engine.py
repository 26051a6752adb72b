"""Device-resident script index + batched window search (thin host layer over the C ABI).

PyTorch is used only for device memory and streams; every kernel is in csrc/.
Replaces, for one cluster of fanworks, the reference's
``pool.map(multi_search_wrapper, fan_cluster)`` up to the threshold test
(/root/reference search.py:381-386 -> :163-184).
"""
import ctypes
import os

import numpy as np

from . import _native as nt


def _torch():
    import torch
    return torch


class DeviceIndex:
    """Script-side index on one B200 (AnnIndexSearch.__init__, search.py:131-154)."""

    def __init__(self, table, script_tok, script_off=None, extra=None, window=6, threshold=0.1,
                 device=0):
        lib = nt.load()
        self._lib = lib
        self._h = None
        table = np.ascontiguousarray(table, dtype=np.float32)
        if table.ndim != 2:
            raise ValueError("table must be [rows, dim]")
        script_tok = np.ascontiguousarray(script_tok, dtype=np.int32)
        if script_off is None:
            script_off = np.array([0, script_tok.shape[0]], dtype=np.int64)
        script_off = np.ascontiguousarray(script_off, dtype=np.int64)
        if extra is None:
            extra = np.zeros((0, table.shape[1]), dtype=np.float32)
        extra = np.ascontiguousarray(extra, dtype=np.float32).reshape(-1, table.shape[1])
        self.dim = int(table.shape[1])
        self.n_base = int(table.shape[0])
        self.n_script_extra = int(extra.shape[0])
        self.window = int(window)
        self.threshold = float(threshold)
        self.device = int(device)
        self.n_script_tok = int(script_tok.shape[0])
        h = ctypes.c_void_p()
        nt.check(lib.fs_index_create(ctypes.byref(h), device, nt.ptr(table), table.shape[0],
                                     table.shape[1], nt.ptr(extra) if extra.shape[0] else None,
                                     extra.shape[0], nt.ptr(script_tok) if script_tok.size else None,
                                     script_tok.shape[0], nt.ptr(script_off), script_off.shape[0] - 1,
                                     window, threshold))
        self._h = h
        self.dim_pad = int(lib.fs_index_get_info(h, 1))
        self.n_script_windows = int(lib.fs_index_get_info(h, 0))
        self.sm_count = int(lib.fs_index_get_info(h, 2))
        self.scale = float(lib.fs_index_scale(h))
        # experiment knobs (the defaults are chosen by the library, see csrc/api.cu)
        for env, opt in (("FANDOM_SEARCH_DIAG", nt.FS_OPT_DIAG), ("FANDOM_SEARCH_CTA_PAIR", nt.FS_OPT_CTA_PAIR),
                         ("FANDOM_SEARCH_A_RESIDENT", nt.FS_OPT_A_RESIDENT),
                         ("FANDOM_SEARCH_PACKED_SHUFFLE", nt.FS_OPT_PACKED_SHUFFLE),
                         ("FANDOM_SEARCH_OPERAND_BITS", nt.FS_OPT_OPERAND_BITS),
                         ("FANDOM_SEARCH_PREFILTER_DIMS", nt.FS_OPT_PREFILTER_DIMS),
                         ("FANDOM_SEARCH_FUSED_GATHER", nt.FS_OPT_FUSED_GATHER),
                         ("FANDOM_SEARCH_TILE_GROUP", nt.FS_OPT_TILE_GROUP)):
            v = os.environ.get(env)
            if v not in (None, ""):
                self.set_option(opt, int(v))

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if self._h is not None:
            self._lib.fs_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- knobs ------------------------------------------------------------
    def set_option(self, option, value):
        nt.check(self._lib.fs_index_set_option(self._h, option, value))
        if option in (nt.FS_OPT_OPERAND_BITS, nt.FS_OPT_PREFILTER_DIMS):      # the index was re-converted
            self.dim_pad = int(self._lib.fs_index_get_info(self._h, 1))
            self.scale = float(self._lib.fs_index_scale(self._h))

    @property
    def operand_bits(self):
        """16: fp16 operands, 8: fp8 e4m3 operands of the distance kernel."""
        return int(self._lib.fs_index_get_info(self._h, 11))

    def info(self, what):
        """fs_index_get_info(what) (see include/fandom_search.h)."""
        return int(self._lib.fs_index_get_info(self._h, what))

    @property
    def kept_dims(self):
        """Embedding columns the pre-filter's operand rows keep (FS_OPT_PREFILTER_DIMS)."""
        return int(self._lib.fs_index_get_info(self._h, 13))

    @property
    def diag(self):
        """Diagonal-sum factor E of the distance kernel (tensor cores do window/E shifts)."""
        return int(self._lib.fs_index_get_info(self._h, 5))

    @property
    def cta_pair(self):
        return int(self._lib.fs_index_get_info(self._h, 6))

    def set_lsh(self, normals, n_tables, n_bits):
        """Switch on LSH emulation: normals float64 [n_tables*n_bits, window*dim] (None = off)."""
        if normals is None or n_tables == 0:
            nt.check(self._lib.fs_index_set_lsh(self._h, None, 0, 0))
            return
        normals = np.ascontiguousarray(normals, dtype=np.float64)
        if normals.shape != (n_tables * n_bits, self.window * self.dim):
            raise ValueError("normals must be [n_tables*n_bits, window*dim]")
        nt.check(self._lib.fs_index_set_lsh(self._h, nt.ptr(normals), n_tables, n_bits))

    def reserve(self, max_tokens, max_candidates):
        nt.check(self._lib.fs_index_reserve(self._h, max_tokens, max_candidates))

    def timing_reset(self):
        nt.check(self._lib.fs_timing_reset(self._h))

    def timing_read(self):
        ms = ctypes.c_double()
        n = ctypes.c_int64()
        nt.check(self._lib.fs_timing_read(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    # -- host-buffer entry points (what the drop-in search.py calls) -------
    @staticmethod
    def _host_batch(tok, off, extra, dim):
        tok = np.ascontiguousarray(tok, dtype=np.int32)
        off = np.ascontiguousarray(off, dtype=np.int64)
        if extra is None:
            extra = np.zeros((0, dim), dtype=np.float32)
        extra = np.ascontiguousarray(extra, dtype=np.float32).reshape(-1, dim)
        return tok, off, extra

    def search_host(self, tok, off, extra=None, cap=None, out=None):
        """tok/off: host CSR batch.  Returns (matches[MATCH_DTYPE], counters[int64 x FS_CNT_COUNT])."""
        tok, off, extra = self._host_batch(tok, off, extra, self.dim)
        if cap is None:
            cap = max(4096, tok.shape[0] // 8)
        counters = np.zeros(nt.FS_CNT_COUNT, dtype=np.int64)
        while True:
            if out is None or out.shape[0] < cap:
                out = np.empty(cap, dtype=nt.MATCH_DTYPE)
            st = self._lib.fs_search_csr_host(
                self._h, nt.ptr(tok) if tok.size else None, tok.shape[0], nt.ptr(off),
                off.shape[0] - 1, nt.ptr(extra) if extra.shape[0] else None, extra.shape[0],
                nt.ptr(out), cap, nt.ptr(counters))
            if st == nt.FS_E_OVERFLOW:
                if counters[nt.FS_CNT_OVERFLOW] & nt.FS_OVERFLOW_CANDIDATES:
                    self.reserve(tok.shape[0], int(counters[nt.FS_CNT_CANDIDATES]) * 5 // 4 + 1024)
                    # the match count of an overflowed candidate list is a lower bound
                    cap = max(cap, int(counters[nt.FS_CNT_CANDIDATES]))
                else:
                    cap = int(counters[nt.FS_CNT_MATCHES]) + 1024
                out = None
                continue
            nt.check(st)
            return out[:counters[nt.FS_CNT_MATCHES]], counters

    def search_submit(self, tok, off, extra=None, cap=None):
        """Asynchronous form of search_host (fs_search_submit): enqueues the batch and returns a
        ticket at once; up to two batches may be in flight.  The ticket keeps the host arrays alive
        (the copies run asynchronously from them; pass page-locked arrays for true DMA)."""
        tok, off, extra = self._host_batch(tok, off, extra, self.dim)
        if cap is None:
            cap = max(4096, tok.shape[0] // 8)
        t = ctypes.c_int32(-1)
        nt.check(self._lib.fs_search_submit(
            self._h, nt.ptr(tok) if tok.size else None, tok.shape[0], nt.ptr(off), off.shape[0] - 1,
            nt.ptr(extra) if extra.shape[0] else None, extra.shape[0], cap, ctypes.byref(t)))
        return {'ticket': t.value, 'tok': tok, 'off': off, 'extra': extra, 'cap': cap}

    def search_collect(self, ticket, out=None):
        """Wait for one submitted batch: (matches[MATCH_DTYPE], counters).  A batch whose buffers
        overflowed is searched again synchronously with larger ones (rare: the first clusters of a run)."""
        cap = ticket['cap']
        if out is None or out.shape[0] < cap:
            out = np.empty(cap, dtype=nt.MATCH_DTYPE)
        counters = np.zeros(nt.FS_CNT_COUNT, dtype=np.int64)
        st = self._lib.fs_search_collect(self._h, ticket['ticket'], nt.ptr(out), cap, nt.ptr(counters))
        if st == nt.FS_E_OVERFLOW:
            grow = max(int(counters[nt.FS_CNT_CANDIDATES]), int(counters[nt.FS_CNT_MATCHES])) * 5 // 4 + 1024
            if counters[nt.FS_CNT_OVERFLOW] & nt.FS_OVERFLOW_CANDIDATES:
                self.reserve(ticket['tok'].shape[0], grow)
            return self.search_host(ticket['tok'], ticket['off'], ticket['extra'], cap=max(cap, grow))
        nt.check(st)
        return out[:counters[nt.FS_CNT_MATCHES]], counters

    # -- search + records on the device ------------------------------
    def set_script_text(self, blob, word_off):
        """Register the lower-cased script words (utf-8 blob + offsets, one word per script token)
        for the device-side Levenshtein of fs_search_submit_rows."""
        word_off = np.ascontiguousarray(word_off, dtype=np.int64)
        nt.check(self._lib.fs_index_set_script_text(self._h, blob, nt.ptr(word_off), word_off.shape[0] - 1))
        self._script_text_set = True

    @property
    def device_records(self):
        return bool(getattr(self, "_script_text_set", False))

    def search_submit_rows(self, tok, off, extra, text, tok_start32, tok_len16, lsh_filter=False,
                           cap=None, cap_rows=None):
        """fs_search_submit_rows: like search_submit, plus the verbatim token texts of the batch;
        the matching search_collect_rows returns the WINNING ROWS (top-10 per window, Levenshtein,
        per-word argmin done on the GPU)."""
        tok, off, extra = self._host_batch(tok, off, extra, self.dim)
        text = np.ascontiguousarray(text, dtype=np.uint8)
        tok_start32 = np.ascontiguousarray(tok_start32, dtype=np.uint32)
        tok_len16 = np.ascontiguousarray(tok_len16, dtype=np.uint16)
        hint = getattr(self, "_cap_hint", (0, 0))
        if cap is None:
            cap = max(4096, tok.shape[0] // 8, hint[0])
        if cap_rows is None:
            cap_rows = max(4096, tok.shape[0] // 4, hint[1])
        t = ctypes.c_int32(-1)
        nt.check(self._lib.fs_search_submit_rows(
            self._h, nt.ptr(tok) if tok.size else None, tok.shape[0], nt.ptr(off), off.shape[0] - 1,
            nt.ptr(extra) if extra.shape[0] else None, extra.shape[0],
            nt.ptr(text) if text.size else None, text.shape[0], nt.ptr(tok_start32) if tok.size else None,
            nt.ptr(tok_len16) if tok.size else None, 1 if lsh_filter else 0, cap, cap_rows, ctypes.byref(t)))
        return {'ticket': t.value, 'tok': tok, 'off': off, 'extra': extra, 'cap': cap, 'cap_rows': cap_rows,
                'text': (text, tok_start32, tok_len16)}

    def search_collect_rows(self, ticket):
        """(rows[ROW_DTYPE] sorted by (work, word), counters), or (None, None) when the device could
        not finish the records (a buffer overflowed, or a window text was too long for the device
        Levenshtein): the ticket is then still open for search_collect (raw matches)."""
        cap = ticket['cap_rows']
        out = np.empty(cap, dtype=nt.ROW_DTYPE)
        counters = np.zeros(nt.FS_CNT_COUNT, dtype=np.int64)
        st = self._lib.fs_search_collect_rows(self._h, ticket['ticket'], nt.ptr(out), cap, nt.ptr(counters))
        if st == nt.FS_E_OVERFLOW:
            self._cap_hint = (int(counters[nt.FS_CNT_MATCHES]) * 5 // 4 + 1024,
                              int(counters[nt.FS_CNT_ROWS]) * 5 // 4 + 1024)
            return None, counters
        nt.check(st)
        return out[:counters[nt.FS_CNT_ROWS]], counters

    def exact_join_host(self, tok, off, cap=None):
        tok, off, _ = self._host_batch(tok, off, None, self.dim)
        if cap is None:
            cap = max(4096, tok.shape[0] // 8)
        counters = np.zeros(nt.FS_CNT_COUNT, dtype=np.int64)
        while True:
            out = np.empty(cap, dtype=nt.PAIR_DTYPE)
            st = self._lib.fs_exact_join_host(self._h, nt.ptr(tok) if tok.size else None,
                                              tok.shape[0], nt.ptr(off), off.shape[0] - 1,
                                              nt.ptr(out), cap, nt.ptr(counters))
            if st == nt.FS_E_OVERFLOW:
                cap = int(counters[nt.FS_CNT_EXACT]) + 1024
                continue
            nt.check(st)
            return out[:counters[nt.FS_CNT_EXACT]], counters

    # -- device-buffer entry points (torch tensors on this device) ---------
    def _stream(self, stream):
        torch = _torch()
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        return ctypes.c_void_p(stream.cuda_stream)

    def to_device(self, tok, off, extra=None):
        torch = _torch()
        dev = torch.device("cuda", self.device)
        tok, off, extra = self._host_batch(tok, off, extra, self.dim)
        tok_t = torch.from_numpy(tok).to(dev)
        off_t = torch.from_numpy(off).to(dev)
        extra_t = torch.from_numpy(extra).to(dev) if extra.shape[0] else None
        return tok_t, off_t, extra_t

    def search_dev(self, tok_t, off_t, extra_t, out_t, counters_t, stream=None):
        """Stream-ordered search; out_t: uint8 tensor of cap*24 bytes, counters_t: int64[4]."""
        cap = out_t.numel() * out_t.element_size() // nt.MATCH_DTYPE.itemsize
        nt.check(self._lib.fs_search_csr_dev(
            self._h, self._stream(stream), nt.ptr(tok_t), tok_t.numel(), nt.ptr(off_t),
            off_t.numel() - 1, nt.ptr(extra_t), 0 if extra_t is None else extra_t.shape[0],
            nt.ptr(out_t), cap, nt.ptr(counters_t)))

    def exact_join_dev(self, tok_t, off_t, out_t, counters_t, stream=None):
        cap = out_t.numel() * out_t.element_size() // nt.PAIR_DTYPE.itemsize
        nt.check(self._lib.fs_exact_join_dev(
            self._h, self._stream(stream), nt.ptr(tok_t), tok_t.numel(), nt.ptr(off_t),
            off_t.numel() - 1, nt.ptr(out_t), cap, nt.ptr(counters_t)))

    # -- stage-level entry points (parity tests, per-kernel timing) --------
    def stage_embed(self, tok_t, off_t, extra_t=None, stream=None):
        torch = _torch()
        n = tok_t.numel()
        if self.operand_bits == 8:      # raw e4m3 bytes (view as torch.float8_e4m3fn)
            emb = torch.empty((n, self.dim_pad), dtype=torch.uint8, device=tok_t.device)
        else:
            emb = torch.empty((n, self.dim_pad), dtype=torch.float16, device=tok_t.device)
        # (|window|, |rounding error of the kept columns|, |dropped columns|, 0)
        thr = torch.empty((n, 4), dtype=torch.float32, device=tok_t.device)
        nt.check(self._lib.fs_stage_embed_dev(
            self._h, self._stream(stream), nt.ptr(tok_t), n, nt.ptr(off_t), off_t.numel() - 1,
            nt.ptr(extra_t), 0 if extra_t is None else extra_t.shape[0], nt.ptr(emb), nt.ptr(thr)))
        return emb, thr

    def stage_dots(self, tok_t, off_t, extra_t=None, stream=None):
        torch = _torch()
        n = tok_t.numel()
        ld = self.n_script_tok
        dots = torch.zeros((n, ld), dtype=torch.float32, device=tok_t.device)
        nt.check(self._lib.fs_stage_dots_dev(
            self._h, self._stream(stream), nt.ptr(tok_t), n, nt.ptr(off_t), off_t.numel() - 1,
            nt.ptr(extra_t), 0 if extra_t is None else extra_t.shape[0], nt.ptr(dots), ld))
        return dots

    def stage_candidates(self, tok_t, off_t, extra_t=None, cap=1 << 20, stream=None):
        torch = _torch()
        out = torch.empty((cap, 2), dtype=torch.int32, device=tok_t.device)
        counters = torch.zeros((nt.FS_CNT_COUNT,), dtype=torch.int64, device=tok_t.device)
        nt.check(self._lib.fs_stage_candidates_dev(
            self._h, self._stream(stream), nt.ptr(tok_t), tok_t.numel(), nt.ptr(off_t),
            off_t.numel() - 1, nt.ptr(extra_t), 0 if extra_t is None else extra_t.shape[0],
            nt.ptr(out), cap, nt.ptr(counters)))
        return out, counters
