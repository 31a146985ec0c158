"""The library's HOST stages on the CPU (tests/emu/wq_host_emu.cu compiles the same headers the product library uses:
class building, class shares, event construction) against the oracle.  A development aid for a container without a
GPU, like test_kernel_logic_cpu.py; the parity tests proper are the -m gpu ones."""
import numpy as np
import pytest

import whippet_jl_b200 as wq
from wq_inputs import synth
from emu.host_emu import HostEmu
from oracle import Oracle, em_psi


@pytest.fixture(scope="module")
def case():
    idx = synth.make_index(n_genes=2300, K=9, seed=71, paralog_frac=0.1)       # >= 2048 genes: the threaded paths run
    rb = synth.simulate_reads(idx, 250000, 101, seed=72, sub_rate=0.01)
    o = Oracle(idx)
    o.process_reads(wq.AlignParam(), rb)
    ek, ev = o.edges()
    counts = dict(node=o.node_counts().astype("<u8"), ekey=ek, ecount=ev.astype("<u8"), longs=o.keyed(0), multis=o.keyed(1))
    o.em_tpm(101)
    return idx, o, counts


def ensure_ambig(o):
    """assign_ambig! adds to the oracle's stores: apply it once per fixture."""
    if not getattr(o, "_ambig_applied", False):
        o.assign_ambig()
        o._ambig_applied = True


def oracle_classes(o):
    # the oracle numbers isoforms from 1 (Julia), the library from 0
    conv = lambda c: ([i - 1 for i in c[0]], c[2], c[3], c[4])
    return [conv(c) for c in o.classes(multi=True)] + [conv(c) for c in o.classes(multi=False)]


def test_class_building_equals_oracle(case):
    idx, o, counts = case
    e = HostEmu(idx)
    e.set_counts(**counts)
    e.build_classes()
    got, nm = e.classes()
    want = oracle_classes(o)
    assert nm == len(o.classes(multi=True)) and len(got) == len(want)
    for (gi, gc, gd, gp), (wi, wc, wd, wp) in zip(got, want):
        assert gi == wi and gc == wc and gp == wp
    e.close()


def test_class_shares_assemble_to_the_same_classes(case):
    idx, o, counts = case
    full = HostEmu(idx)
    full.set_counts(**counts)
    full.build_classes()
    want, nm = full.classes()
    parts = []
    for k in range(3):
        e = HostEmu(idx)
        e.set_counts(**counts)
        e.build_classes(k, 3)
        parts.append(e)
    shares = [e.share() for e in parts]
    for e in parts:
        e.assemble(shares)
        got, gm = e.classes()
        assert gm == nm and got == want
        assert np.array_equal(e.iso_count(), full.iso_count())
        e.close()
    full.close()


def test_assign_ambig_equals_oracle(case):
    """assign_ambig! on the host (with the oracle's EM results injected) equals the oracle's, unpartitioned and as three
    gene partitions (each complete for its own genes).  Runs before the event test, which moves the oracle on."""
    idx, o, counts = case
    props = [x for c in o.classes(multi=True) for x in c[1]] + [x for c in o.classes(multi=False) for x in c[1]]
    import ctypes as C
    from oracle import lib
    tpm, cnt, genetpm = np.zeros(idx.n_isoforms), np.zeros(idx.n_isoforms), np.zeros(idx.n_genes)
    v = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().wqo_em_results(C.c_void_p(o.h), v(tpm), v(cnt), v(genetpm))
    full = HostEmu(idx)
    full.set_counts(**counts)
    full.build_classes()
    node, ek, ev = full.assign_ambig(props, genetpm)
    ensure_ambig(o)
    wk, wv = o.edges()
    assert np.array_equal(node, o.node_counts())
    assert np.array_equal(ek, wk) and np.array_equal(ev, wv)
    G = idx.n_genes
    for k in range(3):
        parts = []
        for j in range(3):
            e = HostEmu(idx)
            e.set_counts(**counts)
            e.build_classes(j, 3)
            parts.append(e)
        shares = [e.share() for e in parts]
        e = parts[k]
        e.assemble(shares)
        pn, pk, pv = e.assign_ambig(props, genetpm)
        g_lo, g_hi = G * k // 3, G * (k + 1) // 3
        n_lo, n_hi = int(idx.node_base[g_lo]), int(idx.node_base[g_hi])
        assert np.array_equal(pn[n_lo:n_hi], node[n_lo:n_hi]), k
        own = lambda keys: ((keys >> np.uint64(32)) > g_lo) & ((keys >> np.uint64(32)) <= g_hi)
        assert np.array_equal(pk[own(pk)], ek[own(ek)]) and np.array_equal(pv[own(pk)], ev[own(ek)]), k
        for x in parts:
            x.close()
    full.close()


def test_event_construction_equals_oracle(case):
    idx, o, counts = case
    ensure_ambig(o)
    compare_events(idx, o, 101, 5000, 100)


def test_event_construction_on_the_fixture(fixture_index, test_param):
    """The reference's own toy graph (zero-length node, both strands, tiny genes), reads of 11-21 bases."""
    from data_gen import fixture_random_reads
    o = Oracle(fixture_index)
    o.process_reads(test_param, fixture_random_reads(fixture_index, 30000))
    o.em_tpm(15)
    o.assign_ambig()
    compare_events(fixture_index, o, 15, 5, 1)


def test_event_construction_in_pieces_and_on_real_structures():
    """Genes of the RefSeq-derived structures ('-' strand, retained introns, many isoforms), built whole and in pieces of 1 and 5
    visited nodes (the way the library spreads a gene with many nodes over its host threads): same records either way."""
    from data_gen import refseq_subset
    idx = refseq_subset(400, seed=3)
    o = Oracle(idx)
    o.process_reads(wq.AlignParam(), synth.simulate_reads(idx, 150000, 101, seed=4, sub_rate=0.01))
    o.em_tpm(101)
    o.assign_ambig()
    for pieces in (0, 1, 5):
        compare_events(idx, o, 101, 2000, 50, pieces=pieces)


def compare_events(idx, o, readlen, min_records, min_em, pieces=0):
    ek, ev = o.edges()
    want = o.process_events(readlen)
    e = HostEmu(idx)
    e.set_event_pieces(pieces)
    got = e.events(o.node_counts(), ek, ev, readlen)
    e.set_event_pieces(0)
    assert len(got) == len(want) and len(want) > min_records
    nem = 0
    for g, w in zip(got, want):
        key = (w["gene"], w["node"])
        assert (g["gene"], g["node"], g["motif"]) == (w["gene"], w["node"], w["motif"]), key
        if w["kind"] == 2 or g["kind"] == 2:
            assert g["kind"] == 2 and w["kind"] == 2 and g["n_utr"] == len(w["utr_psi"]), key
            if g["problem"] >= 0:
                psi, it = em_psi(*e.problem(g["problem"]), readlen)
                if not np.isnan(psi).any():
                    assert [float(np.round(x * 10000.0) / 10000.0) for x in psi] == w["utr_psi"] and it == w["iterations"], key
                    nem += 1
            continue
        assert [p for p in g["paths"][:g["n_inc"]]] == [p for p, _ in w["inc"]], key
        assert [p for p in g["paths"][g["n_inc"]:]] == [p for p, _ in w["exc"]], key
        assert g["total"] == w["total"], key
        if g["problem"] >= 0:
            ninc, cnt, ln, sets, mult, tandem = e.problem(g["problem"])
            psi, it = em_psi(ninc, cnt, ln, sets, mult, tandem, readlen)
            s = 0.0
            for x in psi[:ninc]:
                s += x
            kind = 1 if (0.0 <= s <= 1.0 and g["total"] > 0) else 0
            assert kind == w["kind"], key
            if kind == 1:
                assert s == w["psi"] and it == w["iterations"], key
                assert [float(x) for x in psi[:ninc]] == [v for _, v in w["inc"]] and [float(x) for x in psi[ninc:]] == [v for _, v in w["exc"]], key
            nem += 1
        else:
            kind = g["kind"] if g["kind"] != 1 else (1 if g["total"] > 0 else 0)
            assert kind == w["kind"], key
            if kind == 1:
                assert g["psi"] == w["psi"], key
    assert nem >= min_em
    e.close()


def test_event_cuts_spread_one_wide_gene_over_the_ranks():
    """Partition of the event stage (wq_event_cuts): the cuts are monotone and cover the work list; a gene of 200 nodes whose middle
    100 nodes sit under one long exon-skipping junction beside a dozen shorter ones (the shape of the titin-like gene that holds nearly
    every 1000-path event of a deep human job) is shared by all ranks, each getting about the same number of those nodes, while
    its nodes outside the junction and the junction-free genes cost next to nothing."""
    idx = synth.make_index(n_genes=40, K=9, seed=5, paralog_frac=0.0, mean_exons=150.0)
    nn = np.diff(idx.node_base)
    g = int(np.argmax(nn)) + 1                              # 1-based gene id
    assert nn[g - 1] >= 200
    others = [int(x) + 1 for x in np.argsort(nn)[:2]]
    e = HostEmu(idx)
    keys, vals = [], []
    for n in range(1, 200):
        keys.append((g << 32) | (n << 16) | (n + 1)); vals.append(50.0)
    keys.append((g << 32) | (50 << 16) | 151); vals.append(20.0)
    for a in range(60, 140, 7):
        keys.append((g << 32) | (a << 16) | (a + 3)); vals.append(10.0)
    for gg in others:                                       # two ordinary genes: adjacent edges only
        for n in range(1, 30):
            keys.append((gg << 32) | (n << 16) | (n + 1)); vals.append(30.0)
    order = np.argsort(np.array(keys, "<u8"))
    ek, ev = np.array(keys, "<u8")[order], np.array(vals)[order]
    for nparts in (2, 4, 8):
        pieces, cum, cuts = e.event_cuts(ek, ev, nparts)
        assert cuts[0] == 0 and cuts[-1] == len(pieces) and np.all(np.diff(cuts.astype(np.int64)) >= 0)
        assert np.all(np.diff(cum) > 0)
        under = []                                          # nodes 51..150 of gene 7 per rank
        for r in range(nparts):
            k = 0
            for g_lo, g_hi, n_lo, n_hi in pieces[cuts[r]:cuts[r + 1]]:
                if g_lo == g and g_hi == g and n_hi != 0xFFFFFFFF:
                    k += len([n for n in range(int(n_lo), int(n_hi) + 1) if 51 <= n <= 150])
            under.append(k)
        assert sum(under) == 100 and min(under) >= 100 // nparts - 4 and max(under) <= 100 // nparts + 4, (nparts, under)
    e.close()


def test_node_set_windows():
    """WqBitWin (the node sets of event paths): word-mask adjacency, first / last / list against the per-node definitions,
    including windows of several 64-bit words (genes with hundreds of nodes)."""
    import ctypes as C
    from emu import host_emu
    L = C.CDLL(host_emu.build())
    L.hemu_bitwin_selftest.restype = C.c_uint64
    for span in (5, 64, 65, 128, 129, 400):
        assert L.hemu_bitwin_selftest(C.c_uint32(span), C.c_uint32(3000), C.c_uint64(12345 + span)) == 0, span


def test_paired_assign_ambig_stale_iset_equals_oracle():
    """Paired multi-map classes: the library's host stage and the oracle both hand the FIRST alignment of a class the isoform
    ids left in ambig.iset by the last one (src/quant.jl:493-501, 541-551); few genes and many paralogs make the collision
    of mate-2 node ids with small isoform ids common."""
    idx = synth.make_index(n_genes=12, K=9, seed=31, paralog_frac=1.0)
    f, r = synth.simulate_reads(idx, 40000, 76, seed=32, sub_rate=0.005, paired=True)
    p = wq.AlignParam(seed_pair_range=5000, is_paired=True, is_pair_rc=True)
    o = Oracle(idx)
    o.process_paired_reads(p, f, r)
    ek, ev = o.edges()
    counts = dict(node=o.node_counts().astype("<u8"), ekey=ek, ecount=ev.astype("<u8"), longs=o.keyed(0), multis=o.keyed(1))
    assert len(counts["multis"]) > 50
    o.em_tpm(76)
    props = [x for c in o.classes(multi=True) for x in c[1]] + [x for c in o.classes(multi=False) for x in c[1]]
    import ctypes as C
    from oracle import lib
    tpm, cnt, genetpm = np.zeros(idx.n_isoforms), np.zeros(idx.n_isoforms), np.zeros(idx.n_genes)
    v = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().wqo_em_results(C.c_void_p(o.h), v(tpm), v(cnt), v(genetpm))
    e = HostEmu(idx)
    e.set_counts(**counts)
    e.build_classes()
    node, gk, gv = e.assign_ambig(props, genetpm)
    o.assign_ambig()
    wk, wv = o.edges()
    assert np.array_equal(node, o.node_counts())
    assert np.array_equal(gk, wk) and np.array_equal(gv, wv)
    e.close()
