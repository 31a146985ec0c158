"""SAM emission (wq_sam_batch / wq_sam_header) against a test-side restatement of src/io.jl:69-271 applied to the
oracle's alignments, and against the reference's own known answers: the names of the ten reads of
test/runtests.jl:290-330 encode the CIGAR string and the leftmost coordinate each must be written with."""
import json
import os

import numpy as np
import pytest

import whippet_jl_b200 as wq
from conftest import GOLDEN
from data_gen import fixture_random_reads, pe_case
from helpers import write_sam_one, write_sam_reads
from oracle import Oracle


def batch_text(rb):
    seqs, quals = [], []
    for i in range(rb.n):
        L = int(rb.len[i])
        w = rb.seq2[int(rb.seq_off[i]):int(rb.seq_off[i]) + (L + 31) // 32]
        seqs.append("".join("ACGT"[(int(w[j // 32]) >> (2 * (j % 32))) & 3] for j in range(L)))
        quals.append([int(x) for x in rb.qual[int(rb.qual_off[i]):int(rb.qual_off[i]) + L]])
    return seqs, quals


def test_known_answer_reads_sam_fields(fixture_index, test_param):
    """POS and CIGAR of the SAM line of the best alignment equal the values encoded in the read names."""
    kat = json.load(open(os.path.join(GOLDEN, "kat_reads.json")))["reads"]
    rb = wq.ReadBatch.from_strings([r["seq"] for r in kat], [r["qual"] for r in kat], 33, fill="C")
    res = Oracle(fixture_index).align_se(test_param, rb)
    seqs, quals = batch_text(rb)
    sam = write_sam_reads(fixture_index, [r["name"] for r in kat], seqs, quals, res)
    lines = [l.split("\t") for l in sam.splitlines()]
    primary = {l[0]: l for l in lines if not int(l[1]) & 0x100}
    assert len(primary) == 10
    for r in kat:
        cigar, positions = [x for x in r["name"].split("%") if x][:2]
        l = primary[r["name"]]
        assert l[5] == cigar and int(l[3]) == int(positions.split(",")[0]), r["name"]
        assert l[2] == "chr0" and len(l[9]) == len(r["seq"]) == len(l[10])


def test_sam_formatter_on_cpu(fixture_index, test_param):
    """The library's SAM formatter (csrc/wq_sam.h, compiled into the host test harness) applied to the oracle's
    alignments equals the restatement, single-end on the fixture (both gene strands, supplementary entries) and paired-end
    on a synthetic index with random gene strands."""
    from emu.host_emu import HostSam
    rb = fixture_random_reads(fixture_index, 5000)
    names = [f"r{i} desc" for i in range(rb.n)]
    res = Oracle(fixture_index).align_se(test_param, rb)
    hs = HostSam(fixture_index)
    seqs, quals = batch_text(rb)
    got, want = hs.batch(rb, names, res), write_sam_reads(fixture_index, names, seqs, quals, res)
    assert got == want and got.count("\n") > 1000
    assert {0, 16, 256, 272} <= {int(l.split("\t")[1]) for l in got.splitlines()}
    assert hs.header().startswith("@HD\tVN:1.0\tSO:unsorted\n@SQ\tSN:chr0\tLN:")
    hs.close()
    idx, p, f, r = pe_case(3, 200, 101, 5000, True, 8000)
    rng = np.random.default_rng(5)
    idx.gene_strand = rng.integers(0, 2, idx.n_genes).astype("u1")
    idx.gene_seqname = [f"chr{int(x)}" for x in rng.integers(1, 5, idx.n_genes)]
    n1, n2 = [f"p{i}/1" for i in range(f.n)], [f"p{i}/2" for i in range(f.n)]
    s1, q1 = batch_text(f)
    s2, q2 = batch_text(r)
    res = Oracle(idx).align_pe(p, f, r)
    hs = HostSam(idx)
    got = hs.batch(f, n1, res, mate=r, mate_names=n2)
    want = write_sam_reads(idx, n1, s1, q1, res, mate=(n2, s2, q2))
    assert got == want and got.count("\n") > 2000
    hs.close()


@pytest.mark.gpu
def test_sam_fixture_and_header(fixture_index, test_param):
    rb = fixture_random_reads(fixture_index, 20000)
    names = [f"r{i} desc {i}" for i in range(rb.n)]
    q = wq.GraphLibQuant(fixture_index)
    q.set_gene_info(fixture_index.gene_seqname, fixture_index.gene_strand)
    res = q.ungapped_align(test_param, rb)
    got = q.sam_batch(rb, names, res)
    seqs, quals = batch_text(rb)
    want = write_sam_reads(fixture_index, names, seqs, quals, Oracle(fixture_index).align_se(test_param, rb))
    assert got == want and got.count("\n") > 5000
    flags = {int(l.split("\t")[1]) for l in got.splitlines()}
    assert {0, 16, 256, 272} <= flags                 # both gene strands, regular and supplementary entries
    # write_sam_header (src/io.jl:261-276) compares the stored len + 10000 with the next gene's raw len, so the first gene
    # of a sequence usually decides: kissing's largest nodecoord is 21 -> LN:10021
    refs = {}
    for g in range(fixture_index.n_genes):
        nb0, nb1 = int(fixture_index.node_base[g]), int(fixture_index.node_base[g + 1])
        ln = int(fixture_index.nodecoord[nb0:nb1].max())
        nm = fixture_index.gene_seqname[g]
        if nm not in refs or refs[nm] < ln:
            refs[nm] = ln + 10000
    assert q.sam_header() == "@HD\tVN:1.0\tSO:unsorted\n" + "".join(f"@SQ\tSN:{k}\tLN:{refs[k]}\n" for k in sorted(refs))
    q.close()


@pytest.mark.gpu
def test_sam_synthetic_single_and_paired():
    from wq_inputs import synth
    idx, p, f, r = pe_case(3, 300, 101, 5000, True, 40000)
    rng = np.random.default_rng(5)
    idx.gene_strand = rng.integers(0, 2, idx.n_genes).astype("u1")
    idx.gene_seqname = [f"chr{int(x)}" for x in rng.integers(1, 5, idx.n_genes)]
    q = wq.GraphLibQuant(idx)
    q.set_gene_info(idx.gene_seqname, idx.gene_strand)
    o = Oracle(idx)
    n1, n2 = [f"p{i}/1" for i in range(f.n)], [f"p{i}/2" for i in range(f.n)]
    s1, q1 = batch_text(f)
    s2, q2 = batch_text(r)
    res = q.ungapped_align(p, f, r)
    got = q.sam_batch(f, n1, res, mate=r, mate_names=n2)
    want = write_sam_reads(idx, n1, s1, q1, o.align_pe(p, f, r), mate=(n2, s2, q2))
    assert got == want and got.count("\n") > 10000
    pse = wq.AlignParam()
    res = q.ungapped_align(pse, f)
    got = q.sam_batch(f, n1, res, qualoffset=64)
    want = write_sam_reads(idx, n1, s1, q1, o.align_se(pse, f), qualoffset=64)
    assert got == want and any("N" in l.split("\t")[5] for l in got.splitlines())      # spliced CIGARs present
    q.close()
