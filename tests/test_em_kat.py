"""EM numerics pinned against something other than the oracle's own past:

  * tests/golden/em_kat.json: known answers worked in exact rational arithmetic (tests/golden/make_em_kat.py, no floating
    point and no shared code): TPM after iterations 1, 2, 3 and at convergence, PSI of a 2-path and a 3-path event;
  * oracle/em_restatement.py: a second restatement of gene_em! / spliced_em! / rec_tandem_em!, written from the Julia;
  * the C++ oracle (and, under -m gpu, the CUDA PSI kernels through wq_em_psi_batch).
All three must agree bit for bit; the tests whose names start with `test_assumption_` state the readings of the reference
that cannot be checked without Julia."""
import json
import os

import numpy as np
import pytest

import whippet_jl_b200 as wq
from wq_inputs import synth
from conftest import GOLDEN
from oracle import Oracle, em_psi, gene_em_direct
import em_restatement as R

KAT = json.load(open(os.path.join(GOLDEN, "em_kat.json")))


def test_kat_file_is_what_the_generator_writes(tmp_path):
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(GOLDEN, "make_em_kat.py")], capture_output=True, text=True, check=True).stdout
    assert json.loads(out) == KAT


def test_transcript_em_known_answers():
    k = KAT["tpm"]
    classes = [(ids, c) for ids, c in k["classes"]]
    it, tpm, cnt, trace = gene_em_direct(k["lengths"], k["count"], [], classes, k["readlen"], trace_n=3)
    assert it == k["iterations"]
    assert trace[0].tolist() == k["tpm_after_iteration_1"] and trace[1].tolist() == k["tpm_after_iteration_2"]
    assert trace[2].tolist() == k["tpm_after_iteration_3"]
    assert tpm.tolist() == k["tpm_final"]
    assert np.allclose(cnt, k["count_final"], rtol=0, atol=1e-9)         # sums of products: exact decimals up to binary rounding
    tr = []
    it2, tpm2, cnt2 = R.gene_em(k["count"], k["lengths"], [], [R.EqClass(i, c) for i, c in classes], k["readlen"], trace=tr)
    assert it2 == it and tpm2.tolist() == tpm.tolist() and cnt2.tolist() == cnt.tolist()
    assert [t.tolist() for t in tr[:3]] == [trace[0].tolist(), trace[1].tolist(), trace[2].tolist()]


def _psi_case(k):
    inc, exc = k["inc"], k["exc"]
    cnt = [c for c, _ in inc + exc]
    ln = [l for _, l in inc + exc]
    sets = [[p - 1 for p in paths] for paths, _ in k["ambig"]]
    mult = [m for _, m in k["ambig"]]
    return len(inc), cnt, ln, sets, mult


@pytest.mark.parametrize("name", ["psi2", "psi3"])
def test_psi_em_known_answers(name):
    k = KAT[name]
    ni, cnt, ln, sets, mult = _psi_case(k)
    psi, it = em_psi(ni, cnt, ln, sets, mult, False, 101)
    assert it == k["iterations"] and psi.tolist() == k["psi_final"]
    tr = []
    it2, ip, ep = R.spliced_em(cnt[:ni], ln[:ni], cnt[ni:], ln[ni:], [R.Ambig(p, m) for p, m in k["ambig"]], trace=tr)
    assert it2 == it and ip + ep == psi.tolist()
    assert tr[0][0] + tr[0][1] == k["psi_after_sweep_1"] and tr[1][0] + tr[1][1] == k["psi_after_sweep_2"]


def _classes_of(o):
    return [(c[0], c[2]) for c in o.classes(multi=True)], [(c[0], c[2]) for c in o.classes(multi=False)]


@pytest.mark.parametrize("seed,ng,R_", [(1, 120, 101), (3, 200, 76)])
def test_restatement_equals_oracle_on_synthetic_counts(seed, ng, R_):
    """Whole-index run: the oracle's classes (ids, counts, in its canonical order) and unique counts go through the NumPy
    restatement; TPM, isoform counts and the iteration count must come out identical."""
    idx = synth.make_index(n_genes=ng, K=9, seed=seed, paralog_frac=0.15)
    rb = synth.simulate_reads(idx, 30000, R_, seed=seed + 1)
    o = Oracle(idx)
    o.process_reads(wq.AlignParam(), rb)
    it0, tpm0, cnt0, _ = o.em_tpm(R_, 1, 1)             # maxit = 1: classes built, unique counts, no sweep
    multi, classes = _classes_of(o)
    lengths = []
    for i in range(idx.n_isoforms):
        g = int(np.searchsorted(idx.iso_base, i, "right") - 1)
        nodes = idx.iso_nodes[int(idx.iso_node_ptr[i]):int(idx.iso_node_ptr[i + 1])]
        lengths.append(int(idx.nodelen[idx.node_base[g] + nodes.astype(np.int64) - 1].sum()))
    o2 = Oracle(idx)
    o2.process_reads(wq.AlignParam(), rb)
    it, tpm, cnt, _ = o2.em_tpm(R_, 1, 10000)
    post = [c[4] for c in o2.classes(multi=True)]
    m = [R.EqClass(ids, c, isdone=p) if ids else None for (ids, c), p in zip(multi, post)]
    rit, rtpm, rcnt = R.gene_em(cnt0, lengths, [x for x in m if x is not None], [R.EqClass(i, c) for i, c in classes], R_)
    assert rit == it and rtpm.tolist() == tpm.tolist() and rcnt.tolist() == cnt.tolist()
    assert len(classes) > 20


def test_restatement_equals_oracle_on_event_problems():
    rng = np.random.default_rng(5)
    n1500 = 0
    for t in range(300):
        npth = int(rng.integers(2, 30))
        ni = int(rng.integers(1, npth))
        cnt = np.where(rng.random(npth) < 0.3, 0.0, rng.integers(0, 200, npth)).astype(float).tolist()
        ln = rng.integers(1, 6, npth).astype(float).tolist()
        sets = [sorted(rng.choice(npth, int(rng.integers(2, min(npth, 8) + 1)), replace=False).tolist()) for _ in range(int(rng.integers(0, 6)))]
        mult = rng.integers(1, 100, len(sets)).astype(float).tolist()
        psi, it = em_psi(ni, cnt, ln, sets, mult, False, 101)
        it2, ip, ep = R.spliced_em(cnt[:ni], ln[:ni], cnt[ni:], ln[ni:], [R.Ambig([p + 1 for p in s], m) for s, m in zip(sets, mult)])
        got = np.array(ip + ep)
        assert it2 == it and np.array_equal(got.view(np.uint64), psi.view(np.uint64)), t
        n1500 += it == 1500
        if t < 60:
            tl = rng.integers(50, 900, npth).astype(float).tolist()
            psi, it = em_psi(npth, cnt, tl, sets, mult, True, 101)
            it2, tp = R.tandem_em(cnt, tl, [R.Ambig([p + 1 for p in s], m) for s, m in zip(sets, mult)], 101)
            assert it2 == it and np.array_equal(np.array(tp).view(np.uint64), psi.view(np.uint64)), ("tandem", t)
    assert n1500 > 0          # the never-converging kind (1500 sweeps) is among them


# ---- readings of the reference that Julia alone could confirm ---------------------------------------------------------
def test_assumption_fastmath_does_not_reassociate_tpm_scaling():
    """`@fastmath quant.tpm[i] * SCALING_FACTOR / rpk_sum` (src/quant.jl:662) is taken as (tpm * 1e6) / rpk_sum; the other
    association, tpm * (1e6 / rpk_sum), differs in the last bit for some inputs, so the choice is observable."""
    a, s = 0.07296554464299441, 1.878278103012795
    assert (a * 1e6) / s != a * (1e6 / s)
    raw = [a, s - a]                                 # tpm_i = count_i / (length_i - readlen) with unit effective lengths
    assert R.calculate_tpm(raw, [101, 101], 100, 0)[0] == (a * 1e6) / R.canonical_sum(raw)


def test_assumption_pairwise_sum_is_the_declared_tile_tree():
    """`sum(quant.tpm)` (src/quant.jl:659): Julia's blocked SIMD sum is machine-dependent; the declared order is an
    adjacent-pair tree over tiles of 1024.  A left-to-right sum gives different bits."""
    v = np.random.default_rng(1).random(5000) * 1e-3
    seq = 0.0
    for x in v:
        seq += float(x)
    assert R.canonical_sum(v) != seq
    assert abs(R.canonical_sum(v) - seq) < 1e-12


def test_assumption_class_order_is_observable():
    """Classes are visited in Dict order in the reference; the declared order is the canonical one.  The order can change the
    low bits of the temp counts (floating-point additions onto one isoform), hence the need to declare it."""
    k = KAT["tpm"]
    a = R.gene_em(k["count"], k["lengths"], [], [R.EqClass(i, c) for i, c in k["classes"]], k["readlen"])
    b = R.gene_em(k["count"], k["lengths"], [], [R.EqClass(i, c) for i, c in reversed(k["classes"])], k["readlen"])
    assert a[0] == b[0] and np.allclose(a[1], b[1], rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["psi2", "psi3"])
def test_psi_em_known_answers_on_gpu(name, fixture_index):
    k = KAT[name]
    ni, cnt, ln, sets, mult = _psi_case(k)
    q = wq.GraphLibQuant(fixture_index)
    (psi, it), = q.em_psi_batch([(ni, cnt, ln, sets, mult, False)], 101)
    assert it == k["iterations"] and psi.tolist() == k["psi_final"]
    q.close()
