"""Known answers of the reference's quantification tests, checked on the oracle
(test/runtests.jl:440-522): isoform/multi-map classes, assign_ambig! split, TPM EM invariants."""
import json
import os

import numpy as np

import whippet_jl_b200 as wq
from conftest import GOLDEN
from helpers import identity
from oracle import Oracle


def count_kat_reads(o, idx, p):
    """test/runtests.jl:332-358: count!(gquant, align.value[best_ind], 1.0) for the ten reads."""
    kat = json.load(open(os.path.join(GOLDEN, "kat_reads.json")))["reads"]
    rb = wq.ReadBatch.from_strings([r["seq"] for r in kat], [r["qual"] for r in kat], 33, fill="C")
    res = o.align_se(p, rb)
    for i, r in enumerate(kat):
        alns = res.read(i)
        best = alns[int(np.argmax([identity(a, len(r["seq"])) for a in alns]))]
        path = best[4]
        g = path[0][0]
        if len(path) == 1:
            o.add_node(g, path[0][1], 1.0)
        elif len(path) == 2:
            o.add_edge(g, path[0][1], path[1][1], 1.0)
        else:
            o.add_keyed(0, [len(path), g] + [n[1] for n in path], 1.0)


def test_multimapping_classes_and_assign_ambig(fixture_index, test_param):
    """test/runtests.jl:472-511."""
    idx = fixture_index
    o = Oracle(idx)
    o.quant_reset()
    # alignment to gene 2 nodes [1,7] + gene 4 node [1]: path [1,7] matches no isoform -> isdone, empty class
    o.add_keyed(1, [(1 << 28) | 2, 2, 2, 1, 7, 1, 4, 1], 1.0)
    o.em_tpm(20)
    (ids, props, cnt, isdone, post), = o.classes(multi=True)
    assert isdone and post and ids == []
    # gene 2 nodes [1,2] + gene 2 node [1] -> class [3,4,5]
    o.quant_reset()
    o.add_keyed(1, [(1 << 28) | 2, 2, 2, 1, 2, 1, 2, 1], 10.0)
    o.em_tpm(20, maxit=1)        # maxit=1: classes built, no EM iteration
    (ids, props, cnt, isdone, post), = o.classes(multi=True)
    assert ids == [3, 4, 5] and not isdone and cnt == 10.0
    nb = int(idx.node_base[1])
    prev_node = o.node_counts()[nb]
    o.assign_ambig()
    assert o.node_counts()[nb] == prev_node + 7.5
    k, v = o.edges()
    assert dict(zip(k.tolist(), v.tolist()))[(2 << 32) | (1 << 16) | 2] == 2.5


def test_paired_assign_ambig_stale_iset_known_answer(fixture_index):
    """Hand-worked: a paired multi-map class of two alignments in gene `one` (geneidx 2, isoform ids 3,4,5),
    A1 = (mate 1 [node 1], mate 2 [node 3]) and A2 = (mate 1 [node 1], mate 2 [node 5]), count 10.  Both are compatible with all
    three isoforms, so each gets 5.  assign_ambig! (src/quant.jl:520-553) leaves ambig.iset = {3,4,5} (the isoform ids of A2, set
    by the last set_equivalence_class!, :499-508) when it spreads A1: assign_path! (:283-323) adds node 1 to the set, then finds
    mate-2 node 3 `in temp_iset` -- it is isoform id 3 -- and SKIPS it; the set is emptied, A2 then adds node 1 and node 5.
    Expected: node 1 += 10, node 3 += 0 (5 without the stale ids), node 5 += 5."""
    idx = fixture_index
    o = Oracle(idx)
    o.quant_reset()
    pair = lambda f, r: [(2 << 28) | len(f) | (len(r) << 8), 2] + f + r
    o.add_keyed(1, [(3 << 28) | 2] + pair([1], [3]) + pair([1], [5]), 10.0)
    o.em_tpm(20, maxit=1)
    (ids, props, cnt, isdone, post), = o.classes(multi=True)
    assert ids == [3, 4, 5] and not post and cnt == 10.0
    nb = int(idx.node_base[1])
    before = o.node_counts().copy()
    o.assign_ambig()
    d = o.node_counts() - before
    assert d[nb + 0] == 10.0 and d[nb + 2] == 0.0 and d[nb + 4] == 5.0
    assert d.sum() == 15.0


def test_isoform_classes_and_tpm(fixture_index, test_param):
    """test/runtests.jl:440-470, 513-522: EM changes TPM, converges in < 5000 iterations, sum(tpm) == 1e6."""
    idx = fixture_index
    o = Oracle(idx)
    o.quant_reset()
    count_kat_reads(o, idx, test_param)
    it0, tpm0, cnt0, _ = o.em_tpm(20, sig=1, maxit=1)       # calculate_tpm! only
    cls = o.classes()
    assert len(cls) >= 1 and all(len(c[0]) > 1 and c[2] > 0 for c in cls)
    assert all(2 < i <= 5 for c in cls for i in c[0])          # only gene `one` (isoforms 3..5) has shared nodes
    o.quant_reset()
    count_kat_reads(o, idx, test_param)
    it, tpm, cnt, gtpm = o.em_tpm(20, sig=1, maxit=10000)
    assert not np.array_equal(tpm0, tpm)
    assert it < 5000
    assert tpm.sum() == 1000000
    assert np.allclose(gtpm, [tpm[0:2].sum(), tpm[2:5].sum(), tpm[5], tpm[6]])


def test_graph_enumeration_known_answer():
    """test/runtests.jl:535-580: 15 weighted edges of ENSMUSG00000038685 node 23 -> exactly 18 minimal paths."""
    from oracle import reduce_graph
    edgestr = "1-2:616.0,1-3:4.0,2-3:569.0,2-4:20.0,3-4:629.0,3-6:1.0,3-10:3.0,4-5:1.0,4-6:664.0,5-6:5.0,6-7:789.0,7-8:790.0,8-9:2.0,8-10:606.0,9-10:15.0"
    edges = []
    for s in edgestr.split(","):
        a, rest = s.split("-")
        b, v = rest.split(":")
        edges.append((int(a), int(b), float(v)))
    expected = [[1, 2, 3, 4, 5, 6, 7, 8, 9, 10], [1, 3, 4, 5, 6, 7, 8, 9, 10], [1, 2, 4, 5, 6, 7, 8, 9, 10], [1, 2, 3, 6, 7, 8, 9, 10],
                [1, 3, 6, 7, 8, 9, 10], [1, 2, 3, 10], [1, 3, 10], [1, 2, 3, 4, 6, 7, 8, 9, 10], [1, 3, 4, 6, 7, 8, 9, 10],
                [1, 2, 4, 6, 7, 8, 9, 10], [1, 2, 3, 4, 5, 6, 7, 8, 10], [1, 3, 4, 5, 6, 7, 8, 10], [1, 2, 4, 5, 6, 7, 8, 10],
                [1, 2, 3, 6, 7, 8, 10], [1, 3, 6, 7, 8, 10], [1, 2, 3, 4, 6, 7, 8, 10], [1, 3, 4, 6, 7, 8, 10], [1, 2, 4, 6, 7, 8, 10]]
    got = reduce_graph(edges)
    assert len(got) == len(expected) == 18
    for p in expected:
        assert frozenset(p) in got


def test_event_building_on_fixture(fixture_index, test_param):
    """test/runtests.jl:582-599 analogue: every node of every multi-node gene yields one event record, PSI in [0,1]."""
    o = Oracle(fixture_index)
    o.quant_reset()
    count_kat_reads(o, fixture_index, test_param)
    o.em_tpm(20)
    o.assign_ambig()
    ev = o.process_events(20)
    genes = {e["gene"] for e in ev}
    assert genes == {1, 2}                                  # `single` genes have <= 2 edges: no events
    for e in ev:
        assert e["kind"] in (0, 1, 2, 3)
        if e["kind"] == 1:
            assert 0.0 <= e["psi"] <= 1.0 and e["total"] > 0
    # gene `one`: exon2 (node 3) is skipped by the exon1-exon3def and exon1-exon4 reads and included by exon1-exon2
    e3 = [e for e in ev if e["gene"] == 2 and e["node"] == 3][0]
    assert e3["motif"] == 5 and e3["kind"] == 1 and 0.0 < e3["psi"] < 1.0


def test_path_overlap_known_answers():
    """test/runtests.jl:421-438 ("SGAlignPath Overlap"): the sub-path test paired count! relies on (src/quant.jl:47-62)."""
    from oracle import path_in_path
    single, single_two = [(1, 1)], [(1, 2)]
    two_three, larger = [(1, 2), (1, 3)], [(1, 1), (1, 2), (1, 3)]
    assert path_in_path(single, single)
    assert not path_in_path(single, single_two)
    assert path_in_path(single_two, two_three)
    assert not path_in_path(single, two_three)
    assert path_in_path(single, larger) and path_in_path(single_two, larger) and path_in_path(two_three, larger)


def test_oracle_regression_vectors():
    """The oracle still computes what it computed when it was pinned (tests/golden/make_oracle_regression.py): counts, EM
    iterations, TPM and every event record on the reference's fixture index, to the last bit."""
    import importlib.util
    import json
    import os
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_oracle_regression", os.path.join(GOLDEN, "make_oracle_regression.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    got = mod.compute()
    want = json.load(open(os.path.join(GOLDEN, "oracle_fixture_regression.json")))
    assert json.loads(json.dumps(got)) == want
