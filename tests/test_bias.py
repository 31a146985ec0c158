"""--biascorrect: JointBiasMod / JointBiasCounter (src/bias.jl:103-270) on the hot path.

Known answers from the reference's own "Bias Models" test set (test/runtests.jl:62-103) pin the oracle's restatement; the kernel
bodies (csrc/wq_bias.cuh, run on the CPU by tests/emu) and, on the GPU box, the library through the C ABI are then compared with
the oracle: model tables, every counter's primer-adjusted count and GC factor, TPM, assign_ambig! values and every event record."""
import ctypes as C

import numpy as np
import pytest

import whippet_jl_b200 as wq
from helpers_cmp import bits_equal
from oracle import Oracle, canonical_events, lib as olib


def small_case(paired=False, n=30000, seed=3):
    from wq_inputs import synth
    if paired:
        from data_gen import pe_case
        idx, p, f, r = pe_case(seed, 300, 101, 5000, True, n)
        return idx, p, f, r
    idx = synth.make_index(n_genes=200, K=9, seed=seed, paralog_frac=0.1)
    return idx, wq.AlignParam(), synth.simulate_reads(idx, n, 101, seed=seed + 1, sub_rate=0.01), None


def test_reference_known_answers():
    """test/runtests.jl:86-102: a counter that holds one hexamer twice adjusts to exactly value!(mod, hept) (two tallies over forenpos = 2);
    ExpectedGC has Int(div(1.0, 0.05) + 1) == 20 bins, and a window with GC content 0.5 falls into bin Int(div(0.5, 0.05) + 1) == 10:
    both only under Julia's exact floor division for floats, which the GC-bin table of the library must follow as well."""
    L = olib()
    L.wqo_bias_gc_bin.restype = C.c_int64
    L.wqo_bias_expected_gc_bins.restype = C.c_int64
    assert L.wqo_bias_expected_gc_bins(C.c_double(0.05)) == 20
    assert L.wqo_bias_gc_bin(C.c_int64(25), C.c_int64(50), C.c_double(0.05)) == 10          # dna"AAAAAGCGCG" ^ 5
    bins = [L.wqo_bias_gc_bin(C.c_int64(k), C.c_int64(50), C.c_double(0.05)) for k in range(51)]
    assert bins[0] == 1 and bins[50] == 20 and all(b2 - b1 in (0, 1) for b1, b2 in zip(bins, bins[1:]))
    # independent check of the table: exact rational floor of (k / 50) / 0.05 with 0.05 and k / 50 taken as the doubles they are
    from fractions import Fraction
    assert bins == [int(Fraction(k / 50.0) / Fraction(0.05)) + 1 for k in range(51)]


def check_bias_counters(o, node_pg, edges, longs, multi, fore, back):
    """`o`: an oracle that has seen the same reads.  Compares the model and every counter, bit for bit, before and after gc_adjust!."""
    o.bias_primer_adjust()
    ofore, oback, short = o.bias_model()
    assert short == 0 and bits_equal(fore, ofore) and bits_equal(back, oback)
    on = o.node_counts()
    assert bits_equal(node_pg[0], on)
    ek, ev = o.edges()
    assert sorted(edges) == [int(k) for k in ek]
    assert all(np.float64(edges[int(k)][0]).tobytes() == np.float64(v).tobytes() for k, v in zip(ek, ev))
    ol, om = o.keyed(0), o.keyed(1)
    assert set(ol) == set(longs) and set(om) == set(multi)
    assert all(np.float64(longs[k][0]).tobytes() == np.float64(ol[k]).tobytes() for k in ol)
    assert all(np.float64(multi[k][0]).tobytes() == np.float64(om[k]).tobytes() for k in om)
    # gc_adjust! on the untouched stores (no class building in between): count * factor
    o.bias_gc_adjust()
    gc = lambda p, g: p if g < 0 else p * g
    on2 = o.node_counts()
    assert bits_equal(np.array([gc(p, g) for p, g in zip(*node_pg)]), on2)
    ek2, ev2 = o.edges()
    assert all(np.float64(gc(*edges[int(k)])).tobytes() == np.float64(v).tobytes() for k, v in zip(ek2, ev2))
    om2 = o.keyed(1)
    assert all(np.float64(gc(*multi[k])).tobytes() == np.float64(om2[k]).tobytes() for k in om2)
    return int((np.asarray(node_pg[1]) >= 0).sum())


@pytest.mark.parametrize("paired", [False, True])
def test_kernel_bodies_on_cpu(paired):
    """The per-read tally code of wq_k_bias / wq_k_bias_pe and the arithmetic of wq_k_bias_model / wq_k_bias_adjust, run on the CPU."""
    from emu.emu import Emu
    idx, p, f, r = small_case(paired, n=12000)
    e = Emu(idx)
    e.set_bias(True)
    o = Oracle(idx)
    o.set_bias(True)
    o.quant_reset()
    if paired:
        e.align_pe(p, f, r)
        o.process_paired_reads(p, f, r)
    else:
        e.process_reads(p, f)
        o.process_reads(p, f)
    node_pg, edges, longs, multi, fore, back, short = e.bias_values()
    assert short == 0
    with_gc = check_bias_counters(o, node_pg, edges, longs, multi, fore, back)
    assert with_gc > 50 and len(edges) > 50 and (len(multi) > 10 or paired)


@pytest.mark.parametrize("paired", [False, True])
def test_host_stages_with_bias_on_cpu(paired):
    """Class building, the long-path re-count, gc_adjust! and assign_ambig! of the library's host code (tests/emu/wq_host_emu.cu) fed with
    the tallies of the kernel bodies, against the oracle in the reference's order (bin/whippet-quant.jl:156-176).  Paired: a "long path"
    may be two single-node mates, whose re-count lands on NODE counters before gc_adjust! scales them."""
    import ctypes as C
    from emu.emu import Emu
    from emu.host_emu import HostEmu
    from oracle import lib as olib_
    from wq_inputs import synth
    if paired:
        from data_gen import pe_case
        idx, p, rb, rb2 = pe_case(5, 2300, 101, 5000, True, 60000)
    else:
        idx = synth.make_index(n_genes=2300, K=9, seed=71, paralog_frac=0.1)       # >= 2048 genes: the threaded paths run
        rb, rb2 = synth.simulate_reads(idx, 120000, 101, seed=72, sub_rate=0.01), None
        p = wq.AlignParam()
    e = Emu(idx)
    e.set_bias(True)
    if paired:
        e.align_pe(p, rb, rb2)
    else:
        e.process_reads(p, rb)
    node_pg, edges, longs, multi, fore, back, short = e.bias_values()
    ek, ec = e.edges()
    kl, km = e.keyed()
    lkeys, mkeys = sorted(kl), sorted(km)
    he = HostEmu(idx)
    he.set_counts(e.node_counts(), ek, ec, {k: kl[k] for k in lkeys}, {k: km[k] for k in mkeys})
    col = lambda d, keys, j: np.array([d[k][j] for k in keys] or [0.0])
    he.set_bias(node_pg, (col(edges, [int(k) for k in ek], 0), col(edges, [int(k) for k in ek], 1)),
                (col(longs, lkeys, 0), col(longs, lkeys, 1)), (col(multi, mkeys, 0), col(multi, mkeys, 1)))
    he.build_classes()
    o = Oracle(idx)
    o.set_bias(True)
    o.quant_reset()
    (o.process_paired_reads(p, rb, rb2) if paired else o.process_reads(p, rb))
    o.bias_primer_adjust()
    o.em_tpm(101)
    conv = lambda c: ([i - 1 for i in c[0]], c[2], c[3], c[4])
    want = [conv(c) for c in o.classes(multi=True)] + [conv(c) for c in o.classes(multi=False)]
    got, nm = he.classes()
    assert nm == len(o.classes(multi=True)) and len(got) == len(want) and nm > (10 if paired else 100)
    for (gi, gc, gd, gp), (wi, wc, wd, wp) in zip(got, want):
        assert gi == wi and np.float64(gc).tobytes() == np.float64(wc).tobytes() and gp == wp
    assert any(gc != round(gc) for _, gc, _, _ in got)                          # the counts really are fractional
    props = [x for c in o.classes(multi=True) for x in c[1]] + [x for c in o.classes(multi=False) for x in c[1]]
    tpm, cnt, genetpm = np.zeros(idx.n_isoforms), np.zeros(idx.n_isoforms), np.zeros(idx.n_genes)
    v = lambda a: a.ctypes.data_as(C.c_void_p)
    olib_().wqo_em_results(C.c_void_p(o.h), v(tpm), v(cnt), v(genetpm))
    node, k2, v2 = he.assign_ambig(props, genetpm)
    o.bias_gc_adjust()
    o.assign_ambig()
    wk, wv = o.edges()
    assert bits_equal(node, o.node_counts())
    assert np.array_equal(k2, wk) and bits_equal(v2, wv)
    he.close()


def test_short_mapped_read_is_an_error():
    from wq_inputs import synth
    idx = synth.make_index(n_genes=50, K=9, seed=9)
    rb = synth.simulate_reads(idx, 500, 30, seed=10)
    o = Oracle(idx)
    o.set_bias(True)
    o.quant_reset()
    with pytest.raises(ValueError, match="shorter than 37"):
        o.process_reads(wq.AlignParam(), rb)


def full_bias_parity(q, o, p, f, r):
    q.set_biascorrect(True)
    o.set_bias(True)
    o.quant_reset()
    if r is None:
        sq = q.process_reads(p, f)
        o.process_reads(p, f)
    else:
        sq = q.process_paired_reads(p, f, r)
        o.process_paired_reads(p, f, r)
    so = o.stats()
    assert (sq.total, sq.mapped, sq.multi) == (so["total"], so["mapped"], so["multi"])
    readlen = int(sq.mean_readlen)
    # counters
    b = q.bias_download()
    node, ek, ec, ck, cc = q.counts()
    edges = {int(k): (b["edge"][0][i], b["edge"][1][i]) for i, k in enumerate(ek)}
    edges.update({int(k): (b["circ"][0][i], b["circ"][1][i]) for i, k in enumerate(ck)})
    kl, km = q.keyed(0), q.keyed(1)
    longs = {k: (b["long"][0][i], b["long"][1][i]) for i, k in enumerate(kl)}
    multi = {k: (b["multi"][0][i], b["multi"][1][i]) for i, k in enumerate(km)}
    o2 = Oracle(o.index)                      # a second oracle for the counter-level check (it applies gc_adjust! before the EM)
    o2.set_bias(True)
    o2.quant_reset()
    (o2.process_reads(p, f) if r is None else o2.process_paired_reads(p, f, r))
    with_gc = check_bias_counters(o2, b["node"], edges, longs, multi, b["fore"], b["back"])
    # the quant stage in the reference's order (bin/whippet-quant.jl:156-181)
    o.bias_primer_adjust()
    it_o, tpm_o, cnt_o, gt_o = o.em_tpm(readlen)
    it_q, tpm_q, cnt_q, gt_q = q.gene_em(readlen)
    assert it_q == it_o and bits_equal(tpm_q, tpm_o) and bits_equal(cnt_q, cnt_o) and bits_equal(gt_q, gt_o)
    o.bias_gc_adjust()
    o.assign_ambig()
    vq = q.assign_ambig()
    assert bits_equal(vq[0], o.node_counts())
    ek_o, ev_o = o.edges()
    keys_q = np.concatenate([vq[1], vq[3]])
    vals_q = np.concatenate([vq[2], vq[4]])
    order = np.argsort(keys_q, kind="stable")
    assert np.array_equal(keys_q[order], ek_o) and bits_equal(vals_q[order], ev_o)
    got = canonical_events(*q.process_events_raw(readlen))
    want = o.events_flat(readlen)
    for name, x, y in zip(("hdr", "psi/total/iterations", "values", "path lengths", "path nodes"), got, want):
        assert bits_equal(x, y), f"event records differ in {name}"
    return dict(with_gc=with_gc, iterations=it_q, events=len(want[0]), multi=len(multi), longs=len(longs))


@pytest.mark.gpu
@pytest.mark.parametrize("paired", [False, True])
def test_gpu_bias_parity(paired):
    idx, p, f, r = small_case(paired, n=60000 if not paired else 30000)
    q, o = wq.GraphLibQuant(idx), Oracle(idx)
    s = full_bias_parity(q, o, p, f, r)
    assert s["with_gc"] > 100 and s["events"] > 500
    # several batches (and several chunks inside a batch) accumulate into the same tallies
    one = q.bias_download()
    q.reset()
    cut = [0, f.n // 3, f.n // 3 + 1, f.n]
    sub = lambda rb, lo, hi: rb.slice(lo, hi)
    for lo, hi in zip(cut, cut[1:]):
        if r is None:
            q.process_reads(p, sub(f, lo, hi))
        else:
            q.process_paired_reads(p, sub(f, lo, hi), sub(r, lo, hi))
    two = q.bias_download()
    for k in ("node", "edge", "circ", "long", "multi"):
        assert bits_equal(one[k][0], two[k][0]) and bits_equal(one[k][1], two[k][1]), k
    assert bits_equal(one["fore"], two["fore"]) and bits_equal(one["back"], two["back"])
    q.close()


@pytest.mark.gpu
def test_gpu_bias_refused_cases():
    from wq_inputs import synth
    from whippet_jl_b200 import lib
    idx = synth.make_index(n_genes=50, K=9, seed=9)
    q = wq.GraphLibQuant(idx)
    with pytest.raises(lib.WhippetB200Error, match="bias correction is off"):
        q.bias_download()
    q.set_biascorrect(True)
    with pytest.raises(lib.WhippetB200Error, match="shorter than 37"):
        q.process_reads(wq.AlignParam(), synth.simulate_reads(idx, 500, 30, seed=10))
    q.set_biascorrect(False)
    q.process_reads(wq.AlignParam(), synth.simulate_reads(idx, 500, 30, seed=10))       # fine with DefaultCounter
    q.close()
