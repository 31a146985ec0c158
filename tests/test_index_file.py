"""The library's native on-disk index (wq_index_write / wq_index_load, csrc/wq_index_file.h).

CPU: the writer is host-only, so the file is checked here section by section against the flat GraphLib arrays (occurrence
blocks against the BWT's running counts, suffix array, 2-bit text + ambiguity mask, node records, k-mer tables, annotation
paths, lib.info) and damaged files must be refused with a message.  GPU: a handle loaded from the file gives the same
alignments, counts, TPM and events as one made from the arrays (which the other parity tests compare with the oracle)."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import whippet_jl_b200 as wq
from whippet_jl_b200 import lib

SEC = dict(OCC=1, SA=2, TEXT2=3, AMB=4, NODES=5, GENES=6, BUCKET=7, LPTR=8, RPTR=9, LENT=10, RENT=11, NODE_BASE=12,
           NODEOFFSET=13, NODECOORD=14, NODELEN=15, ISO_BASE=16, ISO_NODE_PTR=17, ISO_NODES=18, NAME_OFF=19, NAMES=20, STRAND=21)
HDR = struct.Struct("<8sIIQiIIIQQQ4IIIII4Q")
NODE_DT = np.dtype([("abs_start", "<u4"), ("len", "<u4"), ("gene", "<u4"), ("node", "<u4"), ("kleft", "<u4"), ("kright", "<u4"),
                    ("et_left", "u1"), ("et_right", "u1"), ("pad", "<u2"), ("pad2", "<u4")])


def parse(path):
    raw = open(path, "rb").read()
    f = HDR.unpack_from(raw, 0)
    hd = dict(magic=f[0], version=f[1], header_bytes=f[2], file_bytes=f[3], K=f[4], G=f[5], N=f[6], I=f[7], n=f[8], sentinel=f[9],
              nslots=f[10], C=f[11:15], has_amb=f[15], has_info=f[16], n_sections=f[17], rec_sizes=f[18])
    secs = {}
    for i in range(hd["n_sections"]):
        sid, elem, off, cnt, ck = struct.unpack_from("<IIQQQ", raw, HDR.size + 32 * i)
        body = raw[off:off + elem * cnt]
        pad = body + b"\0" * (-len(body) % 8)
        assert int(np.frombuffer(pad, "<u8").sum(dtype=np.uint64)) == ck, sid          # checksum: sum of the 64-bit words
        assert off % 4096 == 0
        secs[sid] = (elem, cnt, body)
    return raw, hd, secs


def small_index():
    from wq_inputs import synth
    idx = synth.make_index(n_genes=120, K=9, seed=71, paralog_frac=0.1)
    rng = np.random.default_rng(3)
    idx.gene_strand = rng.integers(0, 2, idx.n_genes).astype("u1")
    idx.gene_seqname = [f"chr{int(x)}" for x in rng.integers(1, 6, idx.n_genes)]
    return idx


def check_file(idx, path, info):
    raw, hd, secs = parse(path)
    n, G, N = idx.text_len, idx.n_genes, idx.n_nodes
    assert hd["magic"] == b"WQXIDX\x01\x00" and hd["version"] == 1 and hd["file_bytes"] == len(raw)
    assert (hd["K"], hd["G"], hd["N"], hd["I"], hd["n"], hd["sentinel"]) == (idx.kmer, G, N, idx.n_isoforms, n, idx.sentinel)
    assert hd["nslots"] == 4 ** idx.kmer and list(hd["C"]) == [int(x) for x in idx.count]
    arr = lambda name, dt: np.frombuffer(secs[SEC[name]][2], dt)
    # occurrence blocks: cnt[4] = symbols of each kind before the block, then the two bit planes of its 64 symbols
    occ = np.frombuffer(secs[SEC["OCC"]][2], np.dtype([("cnt", "<u4", (4,)), ("lo", "<u8"), ("hi", "<u8")]))
    assert len(occ) == n // 64 + 1
    bwt = idx.bwt.astype(np.int64)
    onehot = np.zeros((n + 1, 4), np.int64)
    onehot[np.arange(1, n + 1), bwt] = 1
    run = np.cumsum(onehot, axis=0)
    starts = np.arange(len(occ)) * 64
    assert np.array_equal(occ["cnt"][: (n + 63) // 64], run[starts[: (n + 63) // 64]])
    bits = np.zeros(len(occ) * 64, np.uint8)
    bits[:n] = bwt & 1
    assert np.array_equal(np.packbits(bits, bitorder="little").view("<u8"), occ["lo"])
    bits[:n] = bwt >> 1
    assert np.array_equal(np.packbits(bits, bitorder="little").view("<u8"), occ["hi"])
    assert np.array_equal(arr("SA", "<u4"), idx.sa)
    # 2-bit text and ambiguity mask
    code = np.select([idx.seq4 == 1, idx.seq4 == 2, idx.seq4 == 4, idx.seq4 == 8], [0, 1, 2, 3], default=0).astype(np.uint8)
    t2 = arr("TEXT2", "<u8")
    got = ((t2[np.arange(n) // 32] >> (2 * (np.arange(n) % 32)).astype(np.uint64)) & np.uint64(3)).astype(np.uint8)
    assert np.array_equal(got, code)
    amb = ~np.isin(idx.seq4, [1, 2, 4, 8])
    assert bool(hd["has_amb"]) == bool(amb.any())
    if amb.any():
        m = arr("AMB", "<u4")
        assert np.array_equal(((m[np.arange(n) // 32] >> (np.arange(n) % 32).astype(np.uint32)) & 1).astype(bool), amb)
    # node records
    nodes = np.frombuffer(secs[SEC["NODES"]][2], NODE_DT)
    gene_of = np.repeat(np.arange(G), np.diff(idx.node_base))
    assert np.array_equal(nodes["gene"], gene_of + 1) and np.array_equal(nodes["len"], idx.nodelen)
    assert np.array_equal(nodes["abs_start"], idx.gene_offset[gene_of] + idx.nodeoffset - 1)
    eslot = np.arange(N) + gene_of
    assert np.array_equal(nodes["et_left"], idx.edgetype[eslot]) and np.array_equal(nodes["et_right"], idx.edgetype[eslot + 1])
    assert np.array_equal(nodes["kleft"], idx.edgeleft[eslot].astype("<u4")) and np.array_equal(nodes["kright"], idx.edgeright[eslot].astype("<u4"))
    bucket = arr("BUCKET", "<u4")
    pos = np.arange(len(bucket), dtype=np.int64) * 64
    assert np.array_equal(bucket, np.searchsorted(nodes["abs_start"].astype(np.int64), pos, side="right"))
    # k-mer tables: CSR heads verbatim, entries as (gene, global node index)
    for side, ptr, ent in (("L", idx.left_ptr, idx.left_nodes), ("R", idx.right_ptr, idx.right_nodes)):
        assert np.array_equal(arr(side + "PTR", "<u4"), ptr)
        e = arr(side + "ENT", "<u4").reshape(-1, 2)[: len(ent) // 2]
        pairs = ent.reshape(-1, 2)
        assert np.array_equal(e[:, 0], pairs[:, 0]) and np.array_equal(e[:, 1], idx.node_base[pairs[:, 0] - 1] + pairs[:, 1] - 1)
    for name in ("NODE_BASE", "NODEOFFSET", "NODECOORD", "NODELEN", "ISO_BASE", "ISO_NODE_PTR", "ISO_NODES"):
        assert np.array_equal(arr(name, "<u4"), getattr(idx, name.lower())), name
    assert bool(hd["has_info"]) == info
    if info:
        off = arr("NAME_OFF", "<u8")
        blob = secs[SEC["NAMES"]][2]
        assert [blob[int(off[g]):int(off[g + 1])].decode() for g in range(G)] == [str(x) for x in idx.gene_seqname]
        assert np.array_equal(arr("STRAND", "u1"), np.asarray(idx.gene_strand, "u1"))
    return raw, hd


def test_written_file_matches_the_flat_index(tmp_path, fixture_index):
    p1 = os.path.join(tmp_path, "fixture.wqx")
    wq.write_index(fixture_index, p1)
    check_file(fixture_index, p1, info=True)
    idx = small_index()
    idx.seq4 = idx.seq4.copy()
    idx.seq4[[5, 700, 701, 40000 % idx.text_len]] = 15                  # ambiguity codes -> the mask section exists
    p2 = os.path.join(tmp_path, "synth.wqx")
    wq.write_index(idx, p2)
    check_file(idx, p2, info=True)
    p3 = os.path.join(tmp_path, "noinfo.wqx")
    wq.write_index(idx, p3, with_info=False)
    check_file(idx, p3, info=False)


def test_damaged_files_are_refused(tmp_path, fixture_index):
    L = lib.load()
    good = os.path.join(tmp_path, "good.wqx")
    wq.write_index(fixture_index, good)
    raw = open(good, "rb").read()

    def load_error(data):
        p = os.path.join(tmp_path, "bad.wqx")
        open(p, "wb").write(data)
        h = C.c_void_p()
        rc = L.wq_index_load(os.fsencode(p), 0, C.byref(h))
        assert rc != 0 and not h
        return L.wq_last_error().decode()

    assert "bad magic" in load_error(b"JLS" + raw[3:])
    assert "version" in load_error(raw[:8] + struct.pack("<I", 9) + raw[12:])
    assert "truncated" in load_error(raw[:-4096])
    assert "too short" in load_error(raw[:40])
    flipped = bytearray(raw)
    flipped[HDR.size + 32 * 4 + 8 + 2] ^= 1                             # the offset of a section
    assert "section" in load_error(bytes(flipped))
    body = bytearray(raw)
    _, hd, _ = parse(good)
    off = struct.unpack_from("<IIQQQ", raw, HDR.size + 32 * 4)[2]        # a small section: always verified
    body[off] ^= 0x40
    assert "checksum" in load_error(bytes(body))
    h = C.c_void_p()
    assert L.wq_index_load(os.fsencode(os.path.join(tmp_path, "missing.wqx")), 0, C.byref(h)) != 0
    assert "cannot open" in L.wq_last_error().decode()
    # a writer given an inconsistent GraphLib refuses, too
    bad = wq.FlatIndex.load(os.path.join(os.path.dirname(__file__), "golden", "test_index_flat.npz"))
    bad.iso_nodes = bad.iso_nodes.copy()
    bad.iso_nodes[0] = 99
    with pytest.raises(lib.WhippetB200Error, match="outside its gene"):
        wq.write_index(bad, os.path.join(tmp_path, "x.wqx"))


def test_load_fails_loudly_without_gpu(tmp_path, fixture_index):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = os.path.join(tmp_path, "f.wqx")
    wq.write_index(fixture_index, p)
    with pytest.raises(lib.WhippetB200Error, match="no CUDA device"):
        wq.GraphLibQuant(fixture_index, index_file=p)


@pytest.mark.gpu
def test_loaded_handle_equals_created_handle(tmp_path, fixture_index, test_param):
    from wq_inputs import synth
    from data_gen import fixture_random_reads
    from helpers_cmp import results_equal
    cases = [(fixture_index, test_param, fixture_random_reads(fixture_index, 20000))]
    idx = small_index()
    idx.seq4 = idx.seq4.copy()
    cases.append((idx, wq.AlignParam(), synth.simulate_reads(idx, 60000, 101, seed=5, sub_rate=0.01)))
    for k, (ix, p, rb) in enumerate(cases):
        path = os.path.join(tmp_path, f"i{k}.wqx")
        wq.write_index(ix, path)
        a, b = wq.GraphLibQuant(ix), wq.GraphLibQuant(ix, index_file=path)
        a.set_gene_info(ix.gene_seqname, ix.gene_strand)
        ra, rb_ = a.ungapped_align(p, rb), b.ungapped_align(p, rb)
        assert results_equal(ra, rb_, rb.n) == []
        names = [f"r{i}" for i in range(rb.n)]
        assert a.sam_header() == b.sam_header() and a.sam_batch(rb, names, ra) == b.sam_batch(rb, names, rb_)
        a.reset(); b.reset()
        sa, sb = a.process_reads(p, rb), b.process_reads(p, rb)
        assert (sa.total, sa.mapped, sa.multi, sa.sum_readlen) == (sb.total, sb.mapped, sb.multi, sb.sum_readlen) and sa.mapped > 0
        for x, y in zip(a.counts(), b.counts()):
            assert np.array_equal(x, y)
        assert a.keyed(0) == b.keyed(0) and a.keyed(1) == b.keyed(1)
        ta, tb = a.gene_em(readlen=int(sa.mean_readlen)), b.gene_em(readlen=int(sb.mean_readlen))
        for x, y in zip(ta, tb):
            assert np.array_equal(np.asarray(x), np.asarray(y))
        for x, y in zip(a.assign_ambig(), b.assign_ambig()):
            assert np.array_equal(x, y)
        ea, eb = a.process_events(int(sa.mean_readlen)), b.process_events(int(sb.mean_readlen))
        assert len(ea) == len(eb) and all(x == y for x, y in zip(ea, eb))
        a.close(); b.close()
