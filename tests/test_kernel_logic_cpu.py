"""CPU-side check of the kernel bodies (tests/emu drives csrc/wq_kernels.cuh single-threaded) against
the oracle.  This is a development aid for a container without a GPU: the shipped library has no CPU
path, and the parity tests proper are the -m gpu ones."""
import json
import os

import numpy as np
import pytest

import whippet_jl_b200 as wq
from wq_inputs import synth
from conftest import GOLDEN
from data_gen import PE_CASES, fixture_random_reads, pe_case, ragged_reads, refseq_subset, tiny_node_index
from emu.emu import Emu
from helpers_cmp import results_equal
from oracle import Oracle


def compare_all(idx, p, rb):
    o, e = Oracle(idx), Emu(idx)
    ro, re = o.align_se(p, rb), e.align_se(p, rb)
    assert results_equal(ro, re, rb.n) == []
    o.quant_reset()
    o.process_reads(p, rb)
    st, c = o.stats(), e.counters()
    for k in ("total", "mapped", "multi", "sum_readlen"):
        assert st[k] == c[k], k
    assert c["path_overflow"] == 0 and c["edge_full"] == 0 and c["keyed_full"] == 0
    assert np.array_equal(o.node_counts(), e.node_counts().astype(float))
    ok, ov = o.edges()
    ek, ev = e.edges()
    assert np.array_equal(ok, ek) and np.array_equal(ov, ev.astype(float))
    el, em = e.keyed()
    assert o.keyed(0) == {k: float(v) for k, v in el.items()}
    assert o.keyed(1) == {k: float(v) for k, v in em.items()}
    return ro


def test_fixture_known_answer_reads(fixture_index, test_param):
    kat = json.load(open(os.path.join(GOLDEN, "kat_reads.json")))["reads"]
    rb = wq.ReadBatch.from_strings([r["seq"] for r in kat], [r["qual"] for r in kat], 33, fill="C")
    compare_all(fixture_index, test_param, rb)


def test_fixture_random_reads(fixture_index, test_param):
    ro = compare_all(fixture_index, test_param, fixture_random_reads(fixture_index, 20000))
    assert (ro.nhits > 0).mean() > 0.3 and (ro.nhits > 1).sum() > 50


# read lengths of BASELINE.json config 5's sweep (50-255 bp) and config 4 (150 bp)
@pytest.mark.parametrize("seed,ng,R", [(1, 300, 101), (2, 300, 50), (3, 200, 255), (4, 300, 76), (5, 200, 125), (6, 200, 150), (7, 200, 200)])
def test_synthetic(seed, ng, R):
    idx = synth.make_index(n_genes=ng, K=9, seed=seed, paralog_frac=0.1)
    rb = synth.simulate_reads(idx, 30000, R, seed=seed + 10, sub_rate=0.01)
    ro = compare_all(idx, wq.AlignParam(), rb)
    assert (ro.nhits > 0).mean() > 0.9
    assert (ro.aln["npath"] >= 3).sum() > 20      # long paths exercised


def test_ragged_and_short_reads():
    idx = synth.make_index(n_genes=150, K=9, seed=9, paralog_frac=0.1)
    rb = ragged_reads(idx, 20000)
    compare_all(idx, wq.AlignParam(), rb)


def test_stranded_and_small_k():
    idx = synth.make_index(n_genes=150, K=5, seed=12, paralog_frac=0.2, exon_median=40.0)
    rb = synth.simulate_reads(idx, 20000, 60, seed=3, sub_rate=0.02)
    p = wq.AlignParam(mismatches=2, kmer_size=5, seed_try=4, seed_tolerate=6, seed_length=12, seed_buffer=2,
                      seed_inc=7, is_stranded=True)
    compare_all(idx, p, rb)
    p2 = wq.AlignParam(mismatches=4, kmer_size=5, seed_try=2, seed_tolerate=8, seed_length=10, seed_buffer=1, seed_inc=3)
    compare_all(idx, p2, rb)


@pytest.mark.parametrize("seed,ng,R,pr,rc", PE_CASES)
def test_paired_end(seed, ng, R, pr, rc):
    idx, p, f, r = pe_case(seed, ng, R, pr, rc, 20000)
    o, e = Oracle(idx), Emu(idx)
    ro, re = o.align_pe(p, f, r), e.align_pe(p, f, r)
    assert results_equal(ro, re, f.n, paired=True) == []
    assert (ro.nhits > 0).mean() > 0.85
    o.quant_reset()
    o.process_paired_reads(p, f, r)
    st, c = o.stats(), e.counters()
    for k in ("total", "mapped", "multi", "sum_readlen"):
        assert st[k] == c[k], k
    assert np.array_equal(o.node_counts(), e.node_counts().astype(float))
    ok, ov = o.edges()
    ek, ev = e.edges()
    assert np.array_equal(ok, ek) and np.array_equal(ov, ev.astype(float))
    el, em = e.keyed()
    assert o.keyed(0) == {k: float(v) for k, v in el.items()}
    assert o.keyed(1) == {k: float(v) for k, v in em.items()}


def test_real_structure_minus_strand_and_ambiguous_bases():
    """'-' strand genes, retained introns and zero-length nodes of the refFlat-derived structures; runs of N in the graph text
    (the ambiguity mask of wq_text_mism, which no round-1 input reached)."""
    idx = refseq_subset(600, seed=3, n_frac=0.002)
    assert (np.asarray(idx.gene_strand) == 0).sum() > 200 and (~np.isin(idx.seq4, (1, 2, 4, 8))).sum() > 1000
    rb = synth.simulate_reads(idx, 30000, 101, seed=4, sub_rate=0.01)
    ro = compare_all(idx, wq.AlignParam(), rb)
    assert (ro.nhits > 0).mean() > 0.9


def test_path_overflow_reported_and_second_pass():
    idx = tiny_node_index()
    p = wq.AlignParam()
    rb = synth.simulate_reads(idx, 5000, 101, seed=5, sub_rate=0.002)
    o, e = Oracle(idx), Emu(idx)
    d0 = e.ext_stats()["deferred"]
    ro, re = o.align_se(p, rb), e.align_se(p, rb)
    assert results_equal(ro, re, rb.n) == [] and ro.aln["npath"].max() > 12
    assert e.ext_stats()["deferred"] - d0 > 500 and e.counters()["path_overflow"] == 0
    rb = synth.simulate_reads(idx, 5000, 255, seed=6, sub_rate=0.002)
    o, e = Oracle(idx), Emu(idx)
    ro, re = o.align_se(p, rb), e.align_se(p, rb)
    assert ro.aln["npath"].max() > 24 and e.counters()["path_overflow"] > 1000
    for i in range(rb.n):
        if any(len(a[4]) > 24 for a in ro.read(i)):
            assert re.nhits[i] == 0
        if re.nhits[i]:
            assert re.read(i) == ro.read(i)


def test_small_k_parked_candidates():
    idx = synth.make_index(n_genes=100, K=2, seed=23, paralog_frac=0.2, exon_median=12.0, utr_median=36.0)
    rb = synth.simulate_reads(idx, 20000, 40, seed=24, sub_rate=0.02)
    p = wq.AlignParam(mismatches=3, kmer_size=2, seed_try=3, seed_tolerate=6, seed_length=12, seed_buffer=2, seed_inc=7)
    e = Emu(idx)
    s0 = e.ext_stats()
    compare_all(idx, p, rb)
    s1 = Emu(idx).ext_stats()
    assert s1["parks"] - s0["parks"] > 100 and s1["unparks"] - s0["unparks"] > 50
