"""Splice graphs from real transcript structures (test / benchmark input fabrication, not product code).

BASELINE config 2 names a GENCODE GTF that is not available offline; SURVEY.md section 8d names the stand-in:
the reference's own `anno/refseq_hg19.flat.gz` read as exon lines (gene = name2 + chromosome + strand, so one locus
per gene; transcript = column 1), laid over a random-sequence genome.  This module

  * reads the refFlat table into per-gene transcript structures the way load_gtf does
    (src/refset.jl:68-160: donors = exon ends but the last, acceptors = exon starts but the first, txst/txen with the
    splice sites removed, the set of unique exons);
  * restates the SpliceGraph constructor (src/graph.jl:143-230: the four-way merge of txen/acc/don/txst coordinates,
    node-vs-gap decision by exon containment, edge type tables INDEX_TO_EDGETYPE(_NODE), '-' strand genes built by
    prepending and inverting edge types) and build_annotated_path (:259-277);
  * stores the resulting STRUCTURE (node lengths / coordinates / edge types / isoform node paths / strands) as a
    compact fixture, `anno/refseq_hg19_structure.npz` (made in the build container from /root/reference by
    `python -m wq_inputs.graphbuild`; the GPU box has no /root/reference);
  * turns a structure into a FlatIndex: random ACGT per node (seeded), k-mer edge tables, FM index.

Gene order: chromosomes in first-appearance order of the table, genes by leftmost coordinate (the reference uses
FASTA order and Dict order, which cannot be reproduced without Julia and does not matter to the path).
"""
from __future__ import annotations

import gzip
import os
import sys

import numpy as np

SL, SR, LS, RS, LR, LL, RR = 0, 1, 2, 3, 4, 5, 6
INF = (1 << 32) - 1          # typemax(CoordInt)

# src/graph.jl:15-25 (rows: category of the left coordinate 1 txen, 2 acc, 3 don, 4 txst; columns: of the right one)
INDEX_TO_EDGETYPE = [[RS, SR, RS, RS], [RR, RR, RR, RR], [LL, LR, LL, LS], [SL, SL, SL, SL]]
INDEX_TO_EDGETYPE_NODE = [[RS] * 4, [RR] * 4, [LL] * 4, [SL] * 4]
_HERE = os.path.dirname(os.path.abspath(__file__))
STRUCTURE = os.path.join(_HERE, "anno", "refseq_hg19_structure.npz")


def invert_edgetype(e):
    """src/graph.jl:80-86"""
    return e if e == LR else e ^ 0b011


class Gene:
    __slots__ = ("name", "chrom", "strand", "tx", "don", "acc", "txst", "txen", "exons")

    def __init__(self, name, chrom, strand):
        self.name, self.chrom, self.strand = name, chrom, strand
        self.tx = []            # [(name, [exon starts], [exon ends])] 1-based inclusive, sorted

    def finish(self):
        """private_add_transcript! + the clean-up of load_gtf (src/refset.jl:93-125, 240-242)."""
        don, acc, txst, txen, exons = set(), set(), set(), set(), set()
        for _, st, en in self.tx:
            don.update(en[:-1])
            acc.update(st[1:])
            txst.add(min(st))
            txen.add(max(en))
            exons.update(zip(st, en))
        self.don, self.acc = sorted(don), sorted(acc)
        self.txst, self.txen = sorted(txst - acc), sorted(txen - don)
        self.exons = sorted(exons)


def read_refflat(path):
    """refFlat rows -> {gene key: Gene}; coordinates converted to 1-based inclusive exon lines (start + 1)."""
    genes, order = {}, []
    seen_tx = set()
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as fh:
        for line in fh:
            f = line.rstrip("\n").split("\t")
            if len(f) < 12:
                continue
            name, chrom, strand = f[0], f[1], f[2]
            if "_" in chrom:                      # alternate haplotypes / unplaced contigs: one locus per gene
                continue
            st = [int(x) + 1 for x in f[8].split(",") if x]
            en = [int(x) for x in f[9].split(",") if x]
            if not st or len(st) != len(en):
                continue
            key = (f[11], chrom, strand)
            tid = name
            k = 1
            while (tid, key) in seen_tx:          # a transcript id placed twice in one locus
                k += 1
                tid = f"{name}.{k}"
            seen_tx.add((tid, key))
            if key not in genes:
                genes[key] = Gene(f"{f[11]}|{chrom}|{strand}", chrom, strand == "+")
                order.append(key)
            o = np.argsort(st)
            genes[key].tx.append((tid, [st[i] for i in o], [en[i] for i in o]))
    for g in genes.values():
        g.finish()
    return [genes[k] for k in order]


def splice_graph(g: Gene):
    """SpliceGraph(gene, genome, k) (src/graph.jl:143-230) without the sequence: returns
    (nodecoord, nodelen, edgetype) in graph order ('-' strand genes are built by pushfirst!)."""
    lists = [g.txen, g.acc, g.don, g.txst]          # category order of getmin_ind_val
    idx = [0, 0, 0, 0]
    alen, dlen, slen, plen = len(g.acc), len(g.don), len(g.txst), len(g.txen)
    strand = g.strand
    nodecoord, nodelen, edgetype = [], [], []

    def getmin():
        best, bv = 0, INF + 1
        vals = [lst[i] if i < len(lst) else INF for lst, i in zip(lists, idx)]
        for c in range(4):
            if vals[c] < bv:
                best, bv = c, vals[c]
        return best, bv

    def push(coll, v):
        if strand:
            coll.append(v)
        else:
            coll.insert(0, v)

    def issub(lo, hi):
        for a, b in g.exons:
            if a <= lo and b >= hi:
                return True
        return False

    # the reference's loop condition compares idx[1..4] (txen, acc, don, txst cursors) with the lengths of acc, don,
    # txst, txen in that order (src/graph.jl:176); it is kept as written
    while idx[0] < alen or idx[1] < dlen or idx[2] < slen or idx[3] < plen:
        minidx, minval = getmin()
        idx[minidx] += 1
        secidx, secval = getmin()
        if secval == INF:
            push(edgetype, RS if strand else SL)
            break
        if issub(minval, secval):
            leftadj = 1 if (minidx in (0, 2) and minval != secval) else 0
            rightadj = 1 if (secidx in (1, 3) and minval != secval) else 0
            nodesize = secval - minval - leftadj - rightadj + 1
            edge = INDEX_TO_EDGETYPE_NODE[minidx][secidx]
            pushval = minval + leftadj
        else:
            idx[secidx] += 1
            thridx, thrval = getmin()
            if thrval == INF:
                raise ValueError(f"gene {g.name}: a gap is not followed by a coordinate")
            rightadj = 1 if (thridx in (1, 3) and secval != thrval) else 0
            nodesize = thrval - secval - rightadj + 1
            edge = INDEX_TO_EDGETYPE[minidx][secidx]
            pushval = secval
        if not strand:
            edge = invert_edgetype(edge)
        if nodesize < 0:
            raise ValueError(f"gene {g.name}: negative node size")
        push(nodecoord, pushval)
        push(nodelen, nodesize)
        push(edgetype, edge)
    if len(edgetype) != len(nodelen) + 1:
        raise ValueError(f"gene {g.name}: {len(edgetype)} edges for {len(nodelen)} nodes")
    return nodecoord, nodelen, edgetype


def annotated_path(nodecoord, st, en, strand):
    """build_annotated_path (src/graph.jl:259-277): 1-based node ids of one transcript."""
    nc = np.asarray(nodecoord, np.int64)
    key = nc if strand else -nc
    path = set()
    for a, d in zip(st, en):
        v = a if strand else -a
        lo, hi = np.searchsorted(key, v, "left"), np.searchsorted(key, v, "right")
        if hi <= lo:
            raise ValueError("exon start is not a node boundary")
        ind = hi                                   # last of the equal range, 1-based
        path.add(int(ind))
        cur = ind + (1 if strand else -1)
        while 1 <= cur <= len(nc) and nc[cur - 1] <= d:
            path.add(int(cur))
            cur += 1 if strand else -1
    return sorted(path)


def build_structure(genes, max_genes=None, log=None):
    """All genes -> flat structure arrays (no sequence yet)."""
    by_chrom = {}
    for g in genes:
        by_chrom.setdefault(g.chrom, []).append(g)
    node_base, nodelen, nodecoord, edgetype = [0], [], [], []
    iso_base, iso_ptr, iso_nodes, strands, names = [0], [0], [], [], []
    skipped = 0
    for chrom, gs in by_chrom.items():
        for g in sorted(gs, key=lambda x: min(min(t[1]) for t in x.tx)):
            try:
                nc, nl, et = splice_graph(g)
                paths = [annotated_path(nc, st, en, g.strand) for _, st, en in g.tx]
            except ValueError as e:
                skipped += 1
                if log:
                    print(f"[graphbuild] skipped: {e}", file=log)
                continue
            if len(nl) > 65535 or len(nl) == 0:
                skipped += 1
                continue
            # annotated isoforms as node sets; identical node sets of different transcripts stay separate isoforms, as
            # in the reference (annopath has one BitSet per RefTx)
            for pth in paths:
                iso_nodes.extend(pth)
                iso_ptr.append(len(iso_nodes))
            iso_base.append(iso_base[-1] + len(paths))
            node_base.append(node_base[-1] + len(nl))
            nodelen += nl
            nodecoord += nc
            edgetype += et
            strands.append(1 if g.strand else 0)
            names.append(g.name)
            if max_genes and len(names) >= max_genes:
                break
        if max_genes and len(names) >= max_genes:
            break
    return dict(node_base=np.array(node_base, "<u4"), nodelen=np.array(nodelen, "<u4"), nodecoord=np.array(nodecoord, "<u4"),
                edgetype=np.array(edgetype, "u1"), iso_base=np.array(iso_base, "<u4"), iso_ptr=np.array(iso_ptr, "<u4"),
                iso_nodes=np.array(iso_nodes, "<u2"), strand=np.array(strands, "u1"),
                names=np.array(names), skipped=np.array([skipped]))


def load_structure(path=STRUCTURE):
    z = np.load(path, allow_pickle=False)
    return {k: z[k] for k in z.files}


def index_from_structure(S, K=9, seed=1, n_genes=None, paralog_families=0, n_frac=0.0, threads=None):
    """Structure -> FlatIndex: seeded uniform ACGT per node, optional paralog families (BASELINE config 3: copies of a
    gene's sequence with 1-3 % divergence so seeds hit <= seed_tolerate loci) and optional runs of N (ambiguity mask;
    N is encoded as A in the FM text, src/index.jl:34-42, and never equals a read base)."""
    from whippet_jl_b200.flat import FlatIndex
    from . import synth
    rng = np.random.default_rng(seed)
    G = len(S["node_base"]) - 1 if n_genes is None else min(n_genes, len(S["node_base"]) - 1)
    node_base = S["node_base"][:G + 1].astype(np.int64)
    N = int(node_base[-1])
    nodelen = S["nodelen"][:N].astype(np.int64)
    edgetype = S["edgetype"][:N + G]
    iso_base = S["iso_base"][:G + 1].astype(np.int64)
    iso_ptr = S["iso_ptr"][:int(iso_base[-1]) + 1].astype(np.int64)
    gene_of_node = np.repeat(np.arange(G), np.diff(node_base))
    glen = np.zeros(G, np.int64)
    np.add.at(glen, gene_of_node, nodelen)
    gene_offset = np.cumsum(glen) - glen
    text_len = int(glen.sum())
    cum = np.cumsum(nodelen) - nodelen
    nodeoffset = cum - gene_offset[gene_of_node] + 1
    codes = rng.integers(0, 4, text_len, dtype=np.uint8)
    for _ in range(paralog_families):
        a = int(rng.integers(0, G))
        for _c in range(int(rng.integers(1, 4))):          # 2-4 copies in the family
            b = int(rng.integers(0, G))
            if a == b:
                continue
            ln = int(min(glen[a], glen[b]))
            src = codes[gene_offset[a]:gene_offset[a] + ln].copy()
            mut = rng.random(ln) < rng.uniform(0.01, 0.03)
            src[mut] = (src[mut] + rng.integers(1, 4, int(mut.sum()))) % 4
            codes[gene_offset[b]:gene_offset[b] + ln] = src
    seq4 = (1 << codes).astype("u1")
    if n_frac > 0:
        nruns = max(1, int(text_len * n_frac / 8))
        for s0 in rng.integers(0, max(1, text_len - 16), nruns):
            ln = int(rng.integers(1, 16))
            seq4[s0:s0 + ln] = 15
            codes[s0:s0 + ln] = 0
    el, er, (lp, ln_), (rp, rn) = synth.build_edges(K, gene_offset, node_base, nodeoffset, edgetype, codes, seq4=seq4)
    sa, bwt, sentinel = synth.suffix_array(codes, threads)
    cnt = np.bincount(codes, minlength=4)
    count = np.array([1, 1 + cnt[0], 1 + cnt[0] + cnt[1], 1 + cnt[0] + cnt[1] + cnt[2]], np.int64)
    names = [str(x) for x in S["names"][:G]]
    return FlatIndex(kmer=K, gene_offset=gene_offset, node_base=node_base, nodeoffset=nodeoffset,
                     nodecoord=S["nodecoord"][:N].astype(np.int64), nodelen=nodelen, edgetype=edgetype, edgeleft=el, edgeright=er,
                     seq4=seq4, bwt=bwt, sentinel=sentinel, count=count, sa=sa,
                     left_ptr=lp, left_nodes=ln_, right_ptr=rp, right_nodes=rn,
                     iso_base=iso_base, iso_node_ptr=iso_ptr, iso_nodes=S["iso_nodes"][:int(iso_ptr[-1])].astype("<u4"),
                     gene_names=names, iso_names=[], gene_strand=S["strand"][:G].astype("u1"))


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/anno/refseq_hg19.flat.gz"
    genes = read_refflat(src)
    S = build_structure(genes, log=sys.stderr)
    os.makedirs(os.path.dirname(STRUCTURE), exist_ok=True)
    np.savez_compressed(STRUCTURE, **S)
    G = len(S["node_base"]) - 1
    print(f"{G} genes ({int(S['skipped'][0])} skipped), {len(S['nodelen'])} nodes, {len(S['iso_ptr']) - 1} isoforms, "
          f"text {int(S['nodelen'].sum())} nt, '-' strand genes {int((S['strand'] == 0).sum())}, zero-length nodes "
          f"{int((S['nodelen'] == 0).sum())}, {os.path.getsize(STRUCTURE) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
