"""FASTQ ingest through the library (wq_fastq_*): make_fqparser + read_chunk! + fill! of the reference
(src/reads.jl:2-23, src/record.jl:13-19).  The reader inflates, splits 4-line records and packs them into
wq_read_batch buffers on the library's host threads; Python only holds the handle."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi, lib
from .flat import ReadBatch


class FastqReader:
    def __init__(self, path: str | None = None, text: bytes | None = None, qualoffset: int = 33):
        self.L = lib.load()
        self.h = C.c_void_p()
        self._text = text                      # must outlive the reader
        if path is not None:
            lib.check(self.L.wq_fastq_open(path.encode(), int(qualoffset), C.byref(self.h)))
        else:
            lib.check(self.L.wq_fastq_from_memory(text, len(text), int(qualoffset), C.byref(self.h)))

    def next_batch_c(self, max_reads: int) -> abi.ReadBatchC:
        """The next packed batch as the C struct (buffers owned by the reader, valid until the call after the next)."""
        b = abi.ReadBatchC()
        lib.check(self.L.wq_fastq_next(self.h, int(max_reads), C.byref(b)))
        return b

    def next_batch(self, max_reads: int):
        """(ReadBatch copy, names) of the next batch, or None at the end of the input."""
        b = self.next_batch_c(max_reads)
        n = int(b.n_reads)
        if n == 0:
            return None
        arr = lambda p, cnt, dt: np.ctypeslib.as_array(p, shape=(cnt,)).astype(dt, copy=True)
        rb = ReadBatch(arr(b.len, n, "u1"), arr(b.seq_off, n, "<u8"), arr(b.seq2, max(1, int(b.seq2_words)), "<u8"),
                       arr(b.qual_off, n, "<u8"), arr(b.qual, max(1, int(b.qual_bytes)), "u1"))
        off, ln, names = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)(), C.c_char_p()
        pn = C.c_void_p()
        lib.check(self.L.wq_fastq_names(self.h, C.byref(off), C.byref(ln), C.byref(pn)))
        raw = C.string_at(pn, int(off[n - 1]) + int(ln[n - 1])) if n else b""
        return rb, [raw[int(off[i]):int(off[i]) + int(ln[i])].decode() for i in range(n)]

    def close(self):
        if self.h:
            self.L.wq_fastq_close(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
