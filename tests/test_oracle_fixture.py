"""Pins the oracle (and the .jls reader) against the reference's own fixture and known answers
(SURVEY.md §8c): test/test_index.jls and the ten reads of test/runtests.jl:290-330."""
import json
import os

import numpy as np
import pytest

import whippet_jl_b200 as wq
from conftest import GOLDEN, REFERENCE
from helpers import cigar_string, identity
from oracle import Oracle

ENC = {c: i for i, c in enumerate("ACGT")}


def text_of(idx):
    return "".join({1: "A", 2: "C", 4: "G", 8: "T"}[int(x)] for x in idx.seq4)


def test_fixture_layout(fixture_index):
    idx = fixture_index
    # SURVEY.md §8c "fixture facts"
    assert idx.gene_names == ["kissing", "one", "single_rev", "single"]
    assert list(idx.gene_offset) == [0, 20, 85, 95]
    assert idx.kmer == 2 and idx.text_len == 105
    nb = idx.node_base
    assert list(idx.nodeoffset[nb[1]:nb[2]]) == [1, 16, 26, 36, 39, 48, 51, 61]
    assert list(idx.nodecoord[nb[1]:nb[2]]) == [6, 21, 31, 51, 54, 63, 76, 86]
    assert list(idx.nodelen[nb[1]:nb[2]]) == [15, 10, 10, 3, 9, 3, 10, 5]
    assert list(idx.edgetype[nb[1] + 1:nb[2] + 2]) == [0, 5, 6, 4, 6, 5, 4, 3, 3]
    assert list(idx.nodelen[nb[0]:nb[1]]) == [10, 0, 10]
    assert list(idx.iso_base) == [0, 2, 5, 6, 7]          # geneidx == [0,2,5,6], test/runtests.jl:465
    iso = [list(idx.iso_nodes[idx.iso_node_ptr[i]:idx.iso_node_ptr[i + 1]]) for i in range(idx.n_isoforms)]
    assert iso[2:5] == [[1, 3, 5, 7], [1, 2, 3, 4, 5, 7], [1, 3, 5, 6, 7, 8]]   # test/runtests.jl:205-207
    assert text_of(idx)[:20] == "AAAAAAAAAATGTAATCCGC"


def test_fm_index_is_textbook(fixture_index):
    """The stored FMIndex{2,UInt8} equals a textbook construction byte for byte: suffix array with the
    implicit `$` smallest, bwt[1] = text[end], `$` row dropped, count = [1, 1+#A, ...]."""
    idx = fixture_index
    t = text_of(idx)
    n = len(t)
    sa = sorted(range(n), key=lambda i: t[i:])
    assert list(idx.sa) == sa
    bwt = [t[-1]] + [t[i - 1] for i in sa if i > 0]
    assert "".join("ACGT"[x] for x in idx.bwt) == "".join(bwt)
    assert idx.sentinel == sa.index(0) + 2 == 3
    cnt = [t.count(c) for c in "ACGT"]
    assert list(idx.count) == [1, 1 + cnt[0], 1 + cnt[0] + cnt[1], 1 + sum(cnt[:3])] == [1, 37, 54, 71]


def test_edge_tables(fixture_index):
    """test/runtests.jl:241-268: which of the 16 K=2 slots are assigned."""
    idx = fixture_index
    kmer = lambda s: ENC[s[0]] * 4 + ENC[s[1]]
    left = {kmer(s) for s in ["CA", "AG", "AG", "TC", "AA"]}
    right = {kmer(s) for s in ["GC", "CC", "CT", "TT", "TG"]}
    for i in range(16):
        assert (idx.left_ptr[i + 1] > idx.left_ptr[i]) == (i in left)
        assert (idx.right_ptr[i + 1] > idx.right_ptr[i]) == (i in right)
    l = idx.left_nodes.reshape(-1, 2)
    assert l[idx.left_ptr[kmer("AG")]:idx.left_ptr[kmer("AG") + 1]].tolist() == [[2, 4], [2, 6]]


@pytest.mark.skipif(not os.path.exists(REFERENCE), reason="reference checkout only exists in the build container")
def test_jls_reader_matches_golden(fixture_index):
    idx = wq.FlatIndex.from_jls(os.path.join(REFERENCE, "test", "test_index.jls"))
    for k in wq.FlatIndex._ARRAYS:
        assert np.array_equal(getattr(idx, k), getattr(fixture_index, k)), k
    assert idx.sentinel == fixture_index.sentinel


def test_sa_range_vs_bruteforce(fixture_index):
    """src/fmindex_patch.jl:19-31 + sa_value: hit sets equal brute-force substring search, in row order."""
    idx = fixture_index
    o = Oracle(idx)
    t = text_of(idx)
    rng = np.random.default_rng(1)
    queries = ["GCGGA", "CTATG", "TTACA", "AAAAA", "ACA", "A", "GGGG", "T", "CGCG"]
    queries += ["".join("ACGT"[x] for x in rng.integers(0, 4, rng.integers(1, 7))) for _ in range(300)]
    for q in queries:
        sp, ep = o.sa_range([ENC[c] for c in q])
        hits = [o.sa_value(r) + 1 for r in range(sp, ep + 1)]
        brute = [i + 1 for i in range(len(t)) if t.startswith(q, i)]
        assert sorted(hits) == brute, q
        # rows are in suffix order
        assert [t[h - 1:] for h in hits] == sorted(t[h - 1:] for h in hits)
    sp, ep = o.sa_range([ENC[c] for c in "GCGGA"])
    assert (sp, ep) == (63, 64) and [o.sa_value(r) + 1 for r in (63, 64)] == [96, 26]


def test_ten_known_answer_reads(fixture_index, test_param):
    """test/runtests.jl:332-382: per read non-null, all valid, score spread, strand, path length,
    CIGAR string and leftmost coordinate encoded in the read name."""
    kat = json.load(open(os.path.join(GOLDEN, "kat_reads.json")))
    reads = kat["reads"]
    idx = fixture_index
    o = Oracle(idx)
    rb = wq.ReadBatch.from_strings([r["seq"] for r in reads], [r["qual"] for r in reads], 33, fill=kat["fill_ambiguous"])
    res = o.align_se(test_param, rb)
    for i, r in enumerate(reads):
        alns = res.read(i)
        name, readlen = r["name"], len(r["seq"])
        assert len(alns) >= 1
        assert all(a[3] for a in alns)
        scores = [identity(a, readlen) for a in alns]
        assert max(scores) - min(scores) <= 0.05
        assert alns[0][2] == (0 if ":rc" in name else 1)
        best = alns[int(np.argmax(scores))]
        assert len(best[4]) == len([x for x in name.split("-") if x])
        gene, first = best[4][0][0], best[4][0][1]
        cigar, positions = [x for x in name.split("%") if x][:2]
        firstpos = int(positions.split(",")[0])
        assert cigar_string(best, idx, gene, True, readlen)[0] == cigar, name
        nb = int(idx.node_base[gene - 1])
        assert firstpos == int(idx.nodecoord[nb + first - 1]) + (best[0] - int(idx.nodeoffset[nb + first - 1])), name


def test_hand_traces(fixture_index, test_param):
    """four full hand traces of the reference aligner."""
    kat = json.load(open(os.path.join(GOLDEN, "kat_reads.json")))["reads"]
    o = Oracle(fixture_index)
    rb = wq.ReadBatch.from_strings([r["seq"] for r in kat], [r["qual"] for r in kat], 33, fill="C")
    res = o.align_se(test_param, rb)
    tol = int(np.float32(1 - 10 ** -0.2).view(np.uint32))
    assert res.read(0) == [(1, 2, 1, 1, [(4, 1, 10, 0, 0)]), (5, 1, 1, 1, [(2, 1, 10, 1, tol)])]
    assert res.read(3) == [(5, 1, 1, 1, [(2, 1, 10, 1, tol), (2, 3, 10, 0, 0)])]
    assert res.read(8) == [(5, 1, 1, 1, [(2, 1, 10, 1, tol), (2, 7, 4, 0, 0)])]
    assert res.read(9) == [(14, 1, 0, 1, [(2, 1, 2, 0, 0), (2, 7, 10, 0, 0)])]
    assert res.flipped[9] == 1 and res.flipped[0] == 0
