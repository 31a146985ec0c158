"""GPU parity on the branches round 1 left untested (VERDICT weak item 3), all through the C ABI against the oracle:
'-' strand genes / retained introns / zero-length nodes of the real annotation structure, the ambiguity mask of the graph
text, path-capacity overflow and its reporting, the second-pass extension kernel (long paths, parked candidates), the
seed_tolerate bound, events beyond 1024 paths, the stale iset of the paired assign_ambig!, the running mean_readlen, and the
all-or-nothing re-run of a chunk whose path pool was too small."""
import ctypes as C
import os

import numpy as np
import pytest

import whippet_jl_b200 as wq
from whippet_jl_b200.lib import WhippetB200Error
from wq_inputs import synth
from data_gen import pe_case, ragged_reads, refseq_subset, tiny_node_index
from helpers_cmp import full_pipeline_parity, results_equal
from oracle import Oracle

pytestmark = pytest.mark.gpu


def test_real_structure_minus_strand_and_ambiguous_bases():
    idx = refseq_subset(2000, seed=3, n_frac=0.002)
    assert (np.asarray(idx.gene_strand) == 0).sum() > 500 and (~np.isin(idx.seq4, (1, 2, 4, 8))).sum() > 1000
    rb = synth.simulate_reads(idx, 300_000, 101, seed=4, sub_rate=0.01)
    q, o, p = wq.GraphLibQuant(idx), Oracle(idx), wq.AlignParam()
    assert results_equal(o.align_se(p, rb), q.ungapped_align(p, rb), rb.n) == []
    s = full_pipeline_parity(q, o, p, rb, readlen=101)
    assert s["mapped"] > 0.9 * rb.n and s["events"] > 5000
    q.close()


def test_real_structure_paired_with_paralogs():
    idx = refseq_subset(1500, seed=5, paralog_families=60)
    f, r = synth.simulate_reads(idx, 150_000, 101, seed=6, sub_rate=0.01, paired=True)
    p = wq.AlignParam(seed_pair_range=5000, is_paired=True, is_pair_rc=True)
    q, o = wq.GraphLibQuant(idx), Oracle(idx)
    assert results_equal(o.align_pe(p, f, r), q.ungapped_align(p, f, r), f.n, paired=True) == []
    s = full_pipeline_parity(q, o, p, f, r, readlen=101)
    assert s["multi"] > 100
    q.close()


def test_path_overflow_is_reported_and_second_pass_handles_long_paths():
    idx = tiny_node_index()
    p = wq.AlignParam()
    q, o = wq.GraphLibQuant(idx), Oracle(idx)
    # 101 bp: 13-17 nodes per path -> beyond the first-pass store (12), inside WQ_MAX_PATH: everything equal, nothing dropped
    rb = synth.simulate_reads(idx, 20000, 101, seed=5, sub_rate=0.002)
    ro, rg = o.align_se(p, rb), q.ungapped_align(p, rb)
    assert results_equal(ro, rg, rb.n) == [] and ro.aln["npath"].max() > wq.abi.WQ_FAST_PATH
    assert q.last_deferred() > 1000 and q.stats().path_overflow == 0
    # 255 bp: more than WQ_MAX_PATH nodes -> reported, the read is not counted; reads that did fit are equal to the oracle's
    rb = synth.simulate_reads(idx, 20000, 255, seed=6, sub_rate=0.002)
    ro = o.align_se(p, rb)
    q.reset()
    rg = q.ungapped_align(p, rb)
    st = q.stats()
    assert ro.aln["npath"].max() > wq.abi.WQ_MAX_PATH and st.path_overflow > 1000
    long_reads = {int(i) for i in range(rb.n) if any(len(a[4]) > wq.abi.WQ_MAX_PATH for a in ro.read(i))}
    assert long_reads and all(rg.nhits[i] == 0 for i in long_reads)
    for i in range(rb.n):
        if rg.nhits[i]:
            assert rg.read(i) == ro.read(i)
    assert st.mapped == int((rg.nhits > 0).sum())
    q.close()


def test_small_k_many_candidates_parked():
    idx = synth.make_index(n_genes=100, K=2, seed=23, paralog_frac=0.2, exon_median=12.0, utr_median=36.0)
    rb = synth.simulate_reads(idx, 50000, 40, seed=24, sub_rate=0.02)
    p = wq.AlignParam(mismatches=3, kmer_size=2, seed_try=3, seed_tolerate=6, seed_length=12, seed_buffer=2, seed_inc=7)
    q, o = wq.GraphLibQuant(idx), Oracle(idx)
    assert results_equal(o.align_se(p, rb), q.ungapped_align(p, rb), rb.n) == []
    assert q.last_deferred() > 100          # candidates had to be parked: second pass with side buffers
    full_pipeline_parity(q, o, p, rb, readlen=40)
    q.close()


def test_seed_tolerate_above_bound_is_rejected(fixture_index, test_param):
    q = wq.GraphLibQuant(fixture_index)
    rb = wq.ReadBatch.from_strings(["CGCGGATTACA"], ["#BBBBBBBBBB"], 33)
    bad = wq.AlignParam(1, 2, 4, 9, 4, 5, 1, 2, 1000, 0.05, 0.7, False, False, True, False, True)
    with pytest.raises(WhippetB200Error, match="seed_tolerate"):
        q.process_reads(bad, rb)
    assert q.stats().total == 0             # nothing was counted
    q.process_reads(test_param, rb)         # the handle is still usable
    assert q.stats().total == 1
    q.close()


def test_events_beyond_1024_paths_are_flagged():
    idx = synth.make_stress_index(n_genes=6, K=9, seed=9, cassettes=(11, 11))
    tx = synth.Transcriptome(idx, synth.stress_paths(idx, per_gene=400, seed=10))
    rb = synth.simulate_reads(idx, 150_000, 101, seed=11, tx=tx, random_frac=0.0)
    q, o = wq.GraphLibQuant(idx), Oracle(idx)
    s = full_pipeline_parity(q, o, wq.AlignParam(), rb, readlen=101)
    assert (s["hdr"][:, 4] & 2).any()        # flags bit 1: more than 1024 paths (the reference subsamples at random there)
    assert max(s["hdr"][:, 6].max(), s["hdr"][:, 7].max()) <= 1024     # each graph is cut at 1024 paths before every round (reduce_graph, src/events.jl:234)
    q.close()


def test_paired_assign_ambig_stale_iset():
    """Few genes, many paralogs: multi-map classes whose last alignment lies in a gene with small isoform ids, so mate-2 node
    ids of the first alignment collide with them (src/quant.jl:283-323, 493-501, 541-551)."""
    idx = synth.make_index(n_genes=12, K=9, seed=31, paralog_frac=1.0)
    f, r = synth.simulate_reads(idx, 60000, 76, seed=32, sub_rate=0.005, paired=True)
    p = wq.AlignParam(seed_pair_range=5000, is_paired=True, is_pair_rc=True)
    q, o = wq.GraphLibQuant(idx), Oracle(idx)
    s = full_pipeline_parity(q, o, p, f, r, readlen=76)
    assert s["multi"] > 500
    q.close()


def test_running_mean_readlen_ragged():
    idx = synth.make_index(n_genes=150, K=9, seed=9, paralog_frac=0.1)
    rb = ragged_reads(idx, 50000)
    os.environ["WQ_CHUNK_READS"] = "7000"
    try:
        q = wq.GraphLibQuant(idx)
    finally:
        os.environ.pop("WQ_CHUNK_READS", None)
    o = Oracle(idx)
    s = full_pipeline_parity(q, o, wq.AlignParam(), rb)          # readlen = floor(running mean), as the CLI does
    st, ost = q.stats(), o.stats()
    assert st.mean_readlen == ost["mean_readlen"] and st.mean_readlen != st.sum_readlen / st.mapped
    q.close()


def test_path_pool_regrows_without_double_counting():
    idx = synth.make_index(n_genes=300, K=9, seed=1, paralog_frac=0.1)
    rb = synth.simulate_reads(idx, 100_000, 101, seed=11, sub_rate=0.01)
    os.environ["WQ_POOL_NODES_PER_READ"] = "0"          # 1024 path nodes for the whole chunk: the first attempt must fail
    try:
        q = wq.GraphLibQuant(idx)
    finally:
        os.environ.pop("WQ_POOL_NODES_PER_READ", None)
    o = Oracle(idx)
    full_pipeline_parity(q, o, wq.AlignParam(), rb, readlen=101)
    q.close()


@pytest.mark.gpu
def test_fixed_len_batches_equal_offset_batches(monkeypatch):
    """wq_read_batch.fixed_len: a batch handed over without its offset arrays (the device derives them) gives the same alignments, counts
    and SAM text as the same reads with explicit offsets -- single end with ragged lengths below the capacity, paired end, several chunks
    per call (the chunked H2D path computes its slices from the strides) and the one-chunk path."""
    from data_gen import pe_case
    from helpers_cmp import results_equal
    from wq_inputs import synth
    idx, p, f, r = pe_case(8, 200, 101, 5000, True, 30000)
    rng = np.random.default_rng(4)
    rag = synth.simulate_reads(idx, 30000, 101, seed=12, lens=rng.integers(37, 102, 30000))
    idx.gene_strand = rng.integers(0, 2, idx.n_genes).astype("u1")
    idx.gene_seqname = [f"chr{int(x)}" for x in rng.integers(1, 4, idx.n_genes)]
    for chunk in ("7000", "4194304"):
        monkeypatch.setenv("WQ_CHUNK_READS", chunk)
        a, b = wq.GraphLibQuant(idx), wq.GraphLibQuant(idx)
        for q in (a, b):
            q.set_gene_info(idx.gene_seqname, idx.gene_strand)
        fx = rag.to_fixed()
        assert fx is not None and fx.c().fixed_len == 101 and not fx.c().seq_off and not fx.c().qual_off
        ra, rb_ = a.ungapped_align(wq.AlignParam(), rag), b.ungapped_align(wq.AlignParam(), fx)
        assert results_equal(ra, rb_, rag.n) == []
        names = [f"r{i}" for i in range(rag.n)]
        assert a.sam_batch(rag, names, ra) == b.sam_batch(fx, names, rb_)
        # qualities as 4-bit dictionary indices (qual_bits = 4): expanded on the device, decoded by the SAM formatter
        pk = rag.to_fixed(pack_qual=True)
        assert pk.c().qual_bits == 4 and pk.c().qual_bytes == rag.n * 52 and sorted(set(pk.qual_dict[:2])) == [2, 40]
        rc_ = b.ungapped_align(wq.AlignParam(), pk)
        assert results_equal(ra, rc_, rag.n) == []
        assert a.sam_batch(rag, names, ra) == b.sam_batch(pk, names, rc_)
        a.reset(); b.reset()
        a.process_reads(wq.AlignParam(), rag); b.process_reads(wq.AlignParam(), fx)
        for x, y in zip(a.counts(), b.counts()):
            assert np.array_equal(x, y)
        assert a.keyed(0) == b.keyed(0) and a.keyed(1) == b.keyed(1)
        a.reset(); b.reset()
        ff, rr = f.to_fixed(pack_qual=True), r.to_fixed(pack_qual=True)
        assert ff.c().qual_bits == 4
        sa, sb = a.process_paired_reads(p, f, r), b.process_paired_reads(p, ff, rr)
        assert (sa.total, sa.mapped, sa.multi, sa.sum_readlen) == (sb.total, sb.mapped, sb.multi, sb.sum_readlen) and sa.mapped > 1000
        for x, y in zip(a.counts(), b.counts()):
            assert np.array_equal(x, y)
        assert a.keyed(0) == b.keyed(0) and a.keyed(1) == b.keyed(1)
        a.close(); b.close()
    from whippet_jl_b200 import lib
    bad = fx.c()
    bad.seq2_words = 10                                            # smaller than n_reads strides: refused before anything is read
    q = wq.GraphLibQuant(idx)
    with pytest.raises(lib.WhippetB200Error, match="fixed_len"):
        lib.check(q.L.wq_align_se(q.h, C.byref(wq.AlignParam().c()), C.byref(bad), None))
    q.close()
