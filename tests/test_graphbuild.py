"""Input fabrication check: the SpliceGraph constructor restatement (wq_inputs/graphbuild.py) reproduces the graphs the
reference stored in test/test_index.jls for the GTF of test/runtests.jl:105-122 (node coordinates, lengths, edge types,
annotated paths; '+' and '-' strand genes, the zero-length node of gene `kissing`)."""
import os

import numpy as np

import whippet_jl_b200 as wq
from wq_inputs import graphbuild as gb
from conftest import GOLDEN

# test/runtests.jl:106-121 as (gene, strand, transcript, exons)
GTF = [("one", "+", "def", [(6, 20), (31, 40), (54, 62), (76, 85)]),
       ("one", "+", "int1_alt3", [(6, 40), (51, 62), (76, 85)]),
       ("one", "+", "apa_alt5", [(6, 20), (31, 40), (54, 65), (76, 90)]),
       ("single", "+", "ex1_single", [(11, 20)]),
       ("single_rev", "-", "single_rev", [(11, 20)]),
       ("kissing", "-", "def_kiss", [(11, 20), (21, 30)]),
       ("kissing", "-", "ret_kiss", [(11, 30)])]


def genes():
    out = {}
    for g, s, t, ex in GTF:
        out.setdefault(g, gb.Gene(g, "chr0", s == "+")).tx.append((t, [a for a, _ in ex], [b for _, b in ex]))
    for x in out.values():
        x.finish()
    return out


def test_gene_annotations_known_answers():
    g = genes()["one"]                       # test/runtests.jl:130-131
    assert tuple(g.don) == (20, 40, 62, 65) and tuple(g.acc) == (31, 51, 54, 76)


def test_graphs_equal_the_fixture():
    idx = wq.FlatIndex.load(os.path.join(GOLDEN, "test_index_flat.npz"))
    G = genes()
    for gi, name in enumerate(idx.gene_names):
        nc, nl, et = gb.splice_graph(G[name])
        lo, hi = int(idx.node_base[gi]), int(idx.node_base[gi + 1])
        assert list(idx.nodecoord[lo:hi]) == nc, name
        assert list(idx.nodelen[lo:hi]) == nl, name
        assert list(idx.edgetype[lo + gi:hi + gi + 1]) == et, name
        paths = [gb.annotated_path(nc, st, en, G[name].strand) for _, st, en in G[name].tx]
        for k, p in enumerate(paths):
            a, b = int(idx.iso_node_ptr[idx.iso_base[gi] + k]), int(idx.iso_node_ptr[idx.iso_base[gi] + k + 1])
            assert list(idx.iso_nodes[a:b]) == p, (name, k)
    # test/runtests.jl:207-209
    one = [gb.annotated_path(gb.splice_graph(G["one"])[0], st, en, True) for _, st, en in G["one"].tx]
    assert one == [[1, 3, 5, 7], [1, 2, 3, 4, 5, 7], [1, 3, 5, 6, 7, 8]]


def test_refseq_structure_fixture_is_well_formed():
    S = gb.load_structure()
    G = len(S["node_base"]) - 1
    assert G > 20000 and (S["strand"] == 0).sum() > 5000 and len(S["edgetype"]) == len(S["nodelen"]) + G
    # every gene starts with a transcript start / ends with a transcript end on its own strand orientation
    first = S["edgetype"][S["node_base"][:-1].astype(np.int64) + np.arange(G)]
    last = S["edgetype"][S["node_base"][1:].astype(np.int64) + np.arange(G)]
    assert set(first.tolist()) <= {gb.SL} and set(last.tolist()) <= {gb.RS}
    idx = gb.index_from_structure(S, K=9, seed=3, n_genes=300)
    assert idx.n_genes == 300 and idx.n_isoforms == int(S["iso_base"][300])
