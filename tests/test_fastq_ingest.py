"""FASTQ ingest (wq_fastq_*): packing equals ReadBatch.from_strings with FASTX's fill_ambiguous = DNA_A
(src/reads.jl:9) and fill!'s UInt8 quality arithmetic (src/record.jl:13-19).  Host-side code: runs without a GPU."""
import gzip
import os

import numpy as np
import pytest

import whippet_jl_b200 as wq
from whippet_jl_b200 import lib
from whippet_jl_b200.fastq import FastqReader


def make_records(n, seed, ragged=True):
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(n):
        L = int(rng.integers(0 if i % 97 == 5 else 1, 256)) if ragged else 101
        seq = "".join(rng.choice(list("ACGTacgtNRYn."), p=[.22, .22, .22, .22, .02, .02, .02, .02, .01, .01, .005, .005, .01], size=L))
        qual = "".join(chr(int(q)) for q in rng.integers(33, 75, size=L))
        recs.append((f"read{i}/1 extra words", seq, qual))
    return recs


def to_text(recs, crlf=False, final_newline=True):
    nl = "\r\n" if crlf else "\n"
    t = nl.join(f"@{n}{nl}{s}{nl}+{nl}{q}" for n, s, q in recs)
    return (t + (nl if final_newline else "")).encode()


def expect_batch(recs):
    return wq.ReadBatch.from_strings([r[1] for r in recs], [r[2] for r in recs], 33, fill="A")


def check(rb, names, recs):
    want = expect_batch(recs)
    assert rb.n == len(recs) and names == [r[0] for r in recs]
    assert np.array_equal(rb.len, want.len)
    for i in range(rb.n):
        L = int(rb.len[i])
        nw = (L + 31) // 32
        a = rb.seq2[int(rb.seq_off[i]):int(rb.seq_off[i]) + nw]
        b = want.seq2[int(want.seq_off[i]):int(want.seq_off[i]) + nw]
        assert np.array_equal(a, b), i
        assert np.array_equal(rb.qual[int(rb.qual_off[i]):int(rb.qual_off[i]) + L],
                              want.qual[int(want.qual_off[i]):int(want.qual_off[i]) + L]), i


@pytest.mark.parametrize("crlf,final_newline", [(False, True), (True, True), (False, False)])
def test_memory_batches_equal_python_packer(crlf, final_newline):
    recs = make_records(3000, 7)
    rd = FastqReader(text=to_text(recs, crlf, final_newline))
    got = 0
    for size in (1, 999, 5000):
        out = rd.next_batch(size)
        assert out is not None
        rb, names = out
        check(rb, names, recs[got:got + rb.n])
        assert rb.n == min(size, len(recs) - got)
        got += rb.n
    assert got == len(recs) and rd.next_batch(10) is None
    rd.close()


def test_gzip_file_and_large_batch(tmp_path):
    recs = make_records(60000, 8, ragged=False)        # larger than one 8 MB inflate chunk, threaded conversion
    path = os.path.join(tmp_path, "r.fastq.gz")
    with gzip.open(path, "wb", compresslevel=1) as f:
        f.write(to_text(recs))
    rd = FastqReader(path=path)
    rb, names = rd.next_batch(1 << 20)
    check(rb, names, recs)
    assert rd.next_batch(1) is None
    rd.close()


def test_empty_and_blank_inputs():
    for text, n in ((b"", 0), (b"\n\r\n", 0), (b"@a\nAC\n+\nII\n\n\n", 1), (b"@a\n\n+\n\n@b\nA\n+a\nI", 2)):
        rd = FastqReader(text=text)
        out = rd.next_batch(10)
        assert (out is None and n == 0) or (out is not None and out[0].n == n), text
        assert rd.next_batch(10) is None
        rd.close()


def test_phred64_and_errors(tmp_path):
    rd = FastqReader(text=b"@a\nACGT\n+\nhhhh\n", qualoffset=64)
    rb, _ = rd.next_batch(4)
    assert rb.qual[:4].tolist() == [40] * 4
    rd.close()
    for bad, msg in ((b"@a\nACGT\n+\nhh\n", "lengths differ"), (b"a\nACGT\n+\nhhhh\n", "malformed"), (b"@a\nACGT\n+\n", "truncated"),
                     (b"@a\n" + b"A" * 256 + b"\n+\n" + b"h" * 256 + b"\n", "longer than 255")):
        rd = FastqReader(text=bad)
        with pytest.raises(lib.WhippetB200Error, match=msg):
            rd.next_batch(4)
        rd.close()
    with pytest.raises(lib.WhippetB200Error, match="cannot open"):
        FastqReader(path=os.path.join(tmp_path, "missing.fastq"))


@pytest.mark.gpu
def test_process_fastq_equals_process_reads(tmp_path):
    from wq_inputs import synth
    idx = synth.make_index(n_genes=300, K=9, seed=51, paralog_frac=0.1)
    rb = synth.simulate_reads(idx, 50000, 101, seed=52, sub_rate=0.01)
    # render the packed batch back to FASTQ text
    recs = []
    for i in range(rb.n):
        L = int(rb.len[i])
        w = rb.seq2[int(rb.seq_off[i]):int(rb.seq_off[i]) + (L + 31) // 32]
        codes = [(int(w[j // 32]) >> (2 * (j % 32))) & 3 for j in range(L)]
        q = rb.qual[int(rb.qual_off[i]):int(rb.qual_off[i]) + L]
        recs.append((f"r{i}", "".join("ACGT"[c] for c in codes), "".join(chr(int(x) + 33) for x in q)))
    path = os.path.join(tmp_path, "reads.fastq.gz")
    with gzip.open(path, "wb", compresslevel=1) as f:
        f.write(to_text(recs))
    p = wq.AlignParam()
    a, b = wq.GraphLibQuant(idx), wq.GraphLibQuant(idx)
    sa = a.process_reads(p, rb)
    sb = b.process_fastq(p, path, batch_reads=12000)            # several batches, parse overlapped with alignment
    assert (sa.total, sa.mapped, sa.multi, sa.sum_readlen) == (sb.total, sb.mapped, sb.multi, sb.sum_readlen)
    for x, y in zip(a.counts(), b.counts()):
        assert np.array_equal(x, y)
    assert a.keyed(0) == b.keyed(0) and a.keyed(1) == b.keyed(1)
    a.close(); b.close()


@pytest.mark.gpu
def test_process_fastq_paired(tmp_path):
    from data_gen import pe_case
    idx, p, f, r = pe_case(4, 300, 76, 5000, True, 30000)

    def render(rb, tag):
        recs = []
        for i in range(rb.n):
            L = int(rb.len[i])
            w = rb.seq2[int(rb.seq_off[i]):int(rb.seq_off[i]) + (L + 31) // 32]
            codes = [(int(w[j // 32]) >> (2 * (j % 32))) & 3 for j in range(L)]
            q = rb.qual[int(rb.qual_off[i]):int(rb.qual_off[i]) + L]
            recs.append((f"p{i}/{tag}", "".join("ACGT"[c] for c in codes), "".join(chr(int(x) + 33) for x in q)))
        return recs
    paths = []
    for rb, tag in ((f, 1), (r, 2)):
        path = os.path.join(tmp_path, f"m{tag}.fastq.gz")
        with gzip.open(path, "wb", compresslevel=1) as fh:
            fh.write(to_text(render(rb, tag)))
        paths.append(path)
    a, b = wq.GraphLibQuant(idx), wq.GraphLibQuant(idx)
    sa = a.process_paired_reads(p, f, r)
    sb = b.process_fastq(p, paths[0], paths[1], batch_reads=7000)
    assert (sa.total, sa.mapped, sa.multi, sa.sum_readlen) == (sb.total, sb.mapped, sb.multi, sb.sum_readlen) and sa.mapped > 1000
    for x, y in zip(a.counts(), b.counts()):
        assert np.array_equal(x, y)
    assert a.keyed(0) == b.keyed(0) and a.keyed(1) == b.keyed(1)
    a.close(); b.close()


def _render(rb, fmt):
    recs = []
    for i in range(rb.n):
        L = int(rb.len[i])
        w = rb.seq2[int(rb.seq_off[i]):int(rb.seq_off[i]) + (L + 31) // 32]
        codes = [(int(w[j // 32]) >> (2 * (j % 32))) & 3 for j in range(L)]
        q = rb.qual[int(rb.qual_off[i]):int(rb.qual_off[i]) + L]
        recs.append((fmt.format(i=i), "".join("ACGT"[c] for c in codes), "".join(chr(int(x) + 33) for x in q)))
    return recs


@pytest.mark.gpu
def test_process_fastq_streams_sam(tmp_path):
    """sam = true of process_reads! / process_paired_reads! (src/reads.jl:49-52, 63-67, 110-121): header, then the records
    of every batch in read order, written while the next batch is parsed; equal to wq_sam_header + wq_sam_batch over the
    whole input, with the counts unchanged.  Single end (plain file, many batches) and paired (gzip)."""
    from data_gen import pe_case
    idx, p, f, r = pe_case(6, 300, 101, 5000, True, 30000)
    rng = np.random.default_rng(8)
    idx.gene_strand = rng.integers(0, 2, idx.n_genes).astype("u1")
    idx.gene_seqname = [f"chr{int(x)}" for x in rng.integers(1, 4, idx.n_genes)]
    n1, n2 = [f"p{i}/1 x" for i in range(f.n)], [f"p{i}/2 x" for i in range(f.n)]
    paths = []
    for rb, tag in ((f, 1), (r, 2)):
        path = os.path.join(tmp_path, f"s{tag}.fastq" + (".gz" if tag == 2 else ""))
        text = to_text(_render(rb, "p{i}/%d x" % tag))
        with (gzip.open(path, "wb", compresslevel=1) if tag == 2 else open(path, "wb")) as fh:
            fh.write(text)
        paths.append(path)
    gz1 = os.path.join(tmp_path, "s1.fastq.gz")
    with gzip.open(gz1, "wb", compresslevel=1) as fh:
        fh.write(open(paths[0], "rb").read())
    for paired in (False, True):
        a, b = wq.GraphLibQuant(idx), wq.GraphLibQuant(idx)
        for q in (a, b):
            q.set_gene_info(idx.gene_seqname, idx.gene_strand)
        pp = p if paired else wq.AlignParam()
        if paired:
            res = a.ungapped_align(pp, f, r)
            want = a.sam_header() + a.sam_batch(f, n1, res, mate=r, mate_names=n2)
            a.reset(); sa = a.process_paired_reads(pp, f, r)
        else:
            res = a.ungapped_align(pp, f)
            want = a.sam_header() + a.sam_batch(f, n1, res)
            a.reset(); sa = a.process_reads(pp, f)
        out = os.path.join(tmp_path, f"out{int(paired)}.sam")
        fd = os.open(out, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
        try:
            sb = b.process_fastq(pp, gz1 if paired else paths[0], paths[1] if paired else None, batch_reads=4321, sam_fd=fd)
        finally:
            os.close(fd)
        got = open(out).read()
        assert got == want and got.count("\n") > 20000
        assert (sa.total, sa.mapped, sa.multi, sa.sum_readlen) == (sb.total, sb.mapped, sb.multi, sb.sum_readlen)
        for x, y in zip(a.counts(), b.counts()):
            assert np.array_equal(x, y)
        assert a.keyed(0) == b.keyed(0) and a.keyed(1) == b.keyed(1)
        c = wq.GraphLibQuant(idx)                        # without lib.info the call refuses before touching the counts
        with pytest.raises(lib.WhippetB200Error, match="wq_index_set_gene_info"):
            c.process_fastq(pp, paths[0], sam_fd=1)
        a.close(); b.close(); c.close()


def _batches(path_or_text, batch, from_memory=False):
    """All batches of a reader as (lens, seq words per read, qual bytes per read, names)."""
    import ctypes as C
    from whippet_jl_b200 import abi, lib
    L = lib.load()
    h = C.c_void_p()
    if from_memory:
        lib.check(L.wq_fastq_from_memory(path_or_text, len(path_or_text), 33, C.byref(h)))
    else:
        lib.check(L.wq_fastq_open(str(path_or_text).encode(), 33, C.byref(h)))
    out = []
    while True:
        b = abi.ReadBatchC()
        lib.check(L.wq_fastq_next(h, batch, C.byref(b)))
        if b.n_reads == 0:
            break
        n = b.n_reads
        lens = np.ctypeslib.as_array(b.len, (n,)).copy()
        seq = np.ctypeslib.as_array(b.seq2, (int(b.seq2_words),)).copy()
        qual = np.ctypeslib.as_array(b.qual, (int(b.qual_bytes),)).copy()
        so = np.ctypeslib.as_array(b.seq_off, (n,)).copy()
        qo = np.ctypeslib.as_array(b.qual_off, (n,)).copy()
        out.append((lens, seq, qual, so, qo))
    L.wq_fastq_close(h)
    return out


def test_large_plain_file_blocked_gzip_and_plain_gzip_agree(tmp_path):
    """The three input forms of one text -- mapped plain file (parallel newline scan over a window of many MB), blocked gzip
    (BGZF: parallel inflate) and ordinary gzip (serial, the library's own inflate) -- give the same batches as the in-memory reader, across batch
    boundaries that fall inside scan pieces and inflate windows."""
    import gzip
    from wq_inputs import synth
    idx = synth.make_index(n_genes=60, K=9, seed=3)
    rb = synth.simulate_reads(idx, 120_000, 101, seed=4)
    text = synth.fastq_bytes(rb)
    assert len(text) > 20 * (1 << 20)
    plain, bg, gz = tmp_path / "r.fastq", tmp_path / "r.bgzf.fastq.gz", tmp_path / "r.fastq.gz"
    plain.write_bytes(text)
    bg.write_bytes(synth.bgzf_compress(text))
    with gzip.open(gz, "wb", compresslevel=1) as fh:
        fh.write(text)
    assert gzip.decompress(bg.read_bytes()) == text             # a BGZF file is a valid multi-member gzip file
    want = _batches(text, 50_000, from_memory=True)
    assert sum(len(b[0]) for b in want) == rb.n
    # the packed form equals the batch the text was rendered from
    lens = np.concatenate([b[0] for b in want])
    assert np.array_equal(lens, rb.len) and np.array_equal(np.concatenate([b[1] for b in want]), rb.seq2)
    for path in (plain, bg, gz):
        got = _batches(path, 50_000)
        assert len(got) == len(want), path
        for g, w in zip(got, want):
            assert all(np.array_equal(a, b) for a, b in zip(g, w)), path


def test_ordinary_gzip_own_inflate_zlib_switch_and_damaged_files(tmp_path, monkeypatch):
    """Ordinary gzip goes through the library's own inflate (csrc/wq_inflate.h; tests/test_inflate_cpu.py checks it against zlib): the
    batches equal those through zlib's gzread (WQ_GZ_ZLIB=1) for a best-compression, two-member file; a file cut short or with a
    flipped bit is an error of the reader (the member's CRC-32 and length are verified), not wrong reads."""
    import gzip
    from whippet_jl_b200.lib import WhippetB200Error
    from wq_inputs import synth
    idx = synth.make_index(n_genes=40, K=9, seed=13)
    rb = synth.simulate_reads(idx, 30_000, 101, seed=14)
    text = synth.fastq_bytes(rb)
    cut = text.index(b"\n@", len(text) // 2) + 1                # a record boundary
    gz = tmp_path / "two.fastq.gz"
    gz.write_bytes(gzip.compress(text[:cut], 9) + gzip.compress(text[cut:], 9))
    want = _batches(text, 7000, from_memory=True)
    got = _batches(gz, 7000)
    monkeypatch.setenv("WQ_GZ_ZLIB", "1")
    via_zlib = _batches(gz, 7000)
    monkeypatch.delenv("WQ_GZ_ZLIB")
    for other in (got, via_zlib):
        assert len(other) == len(want)
        for g, w in zip(other, want):
            assert all(np.array_equal(a, b) for a, b in zip(g, w))
    whole = gzip.compress(text, 6)
    short = tmp_path / "short.fastq.gz"
    short.write_bytes(whole[:len(whole) * 2 // 3])
    with pytest.raises(WhippetB200Error):
        _batches(short, 7000)
    flipped = bytearray(whole)
    flipped[len(whole) // 2] ^= 0x40
    bad = tmp_path / "flipped.fastq.gz"
    bad.write_bytes(bytes(flipped))
    with pytest.raises(WhippetB200Error):
        _batches(bad, 7000)


def test_two_readers_side_by_side(tmp_path):
    """wq_process_fastq reads the two files of a paired input on two threads (an ordinary .fastq.gz is inflated serially, so one file
    after the other would halve the rate): two readers running at the same time -- gzip and blocked gzip, several batches, many
    repetitions so that the page-locked block cache is taken and given back concurrently -- return what they return alone."""
    import gzip
    import threading
    from wq_inputs import synth
    idx = synth.make_index(n_genes=40, K=9, seed=23)
    f, r = synth.simulate_reads(idx, 40_000, 101, seed=24, paired=True)
    t1, t2 = synth.fastq_bytes(f), synth.fastq_bytes(r)
    p1, p2 = tmp_path / "m1.fastq.gz", tmp_path / "m2.fastq.gz"
    p1.write_bytes(gzip.compress(t1, 4))
    p2.write_bytes(synth.bgzf_compress(t2))
    alone = [_batches(p1, 9000), _batches(p2, 9000)]
    for _ in range(6):
        got = [None, None]
        def work(k, path):
            got[k] = _batches(path, 9000)
        th = [threading.Thread(target=work, args=(k, path)) for k, path in enumerate((p1, p2))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for k in range(2):
            assert len(got[k]) == len(alone[k])
            for g, w in zip(got[k], alone[k]):
                assert all(np.array_equal(a, b) for a, b in zip(g, w)), k


def test_truncated_and_malformed_files_are_errors(tmp_path):
    from whippet_jl_b200.lib import WhippetB200Error
    good = b"@a\nACGT\n+\nIIII\n@b\nAC\n+\nII\n"
    for bad in (good[:-4], good.replace(b"+\nII", b"-\nII"), good.replace(b"IIII", b"III")):
        with pytest.raises(WhippetB200Error):
            _batches(bad, 10, from_memory=True)
    # blank lines at the end of the file are not records
    assert sum(len(b[0]) for b in _batches(good + b"\n\r\n\n", 10, from_memory=True)) == 2
    assert sum(len(b[0]) for b in _batches(good[:-1], 10, from_memory=True)) == 2          # last line without a newline
