"""Synthetic GraphLib bundles and read sets at the scales BASELINE.json names.

The GENCODE GTFs of the configs are not in the reference checkout and there is no network, so
inputs are fabricated: random-sequence genes whose node/edge structure follows the rules of
SpliceGraph (src/graph.jl:104-220: edge codes SL/SR/LS/RS/LR/LL/RR of src/graph.jl:39-46, nodes
contiguous inside the gene sequence), the k-mer edge tables of build_edges (src/edges.jl:96-116)
and the FM index of trans_index! (src/index.jl:64-110, sigma=4, r=1).  Index construction is
outside the accelerated path; this module exists to feed it (test / benchmark infrastructure, not product code).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from whippet_jl_b200 import abi
from whippet_jl_b200.flat import FlatIndex, ReadBatch

SL, SR, LS, RS, LR, LL, RR = 0, 1, 2, 3, 4, 5, 6
_HERE = os.path.dirname(os.path.abspath(__file__))
_HOST_SO = os.path.join(_HERE, "libwq_host.so")


def build_host_lib(force=False):
    src = os.path.join(_HERE, "sa_build.cpp")
    if force or not os.path.exists(_HOST_SO) or os.path.getmtime(_HOST_SO) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", _HOST_SO, src])
    return _HOST_SO


_host = None


def _hostlib():
    global _host
    if _host is None:
        build_host_lib()
        _host = C.CDLL(_HOST_SO)
    return _host


def suffix_array(codes: np.ndarray, threads: int | None = None):
    """(sa u32[n], bwt u8[n], sentinel) in the FMIndex{2,T} layout."""
    codes = np.ascontiguousarray(codes, "u1")
    n = len(codes)
    sa = np.zeros(n, "<u4")
    bwt = np.zeros(n, "u1")
    sent = C.c_uint64(0)
    threads = threads or max(1, min(32, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    L = _hostlib()
    rc = L.wq_build_sa(abi.ptr(codes, abi.u8p), C.c_uint64(n), abi.ptr(sa, abi.u32p), C.c_int(threads))
    assert rc == 0
    rc = L.wq_build_bwt(abi.ptr(codes, abi.u8p), C.c_uint64(n), abi.ptr(sa, abi.u32p), abi.ptr(bwt, abi.u8p), C.byref(sent))
    assert rc == 0
    return sa, bwt, int(sent.value)


def _is_edge(et, left):
    """is_edge, src/edges.jl:27-33."""
    if left:
        return (et == LR) | (et == LS) | (et == LL)
    return (et == LR) | (et == SR) | (et == RR)


def build_edges(K, gene_offset, node_base, nodeoffset, edgetype, codes, seq4=None):
    """build_edges / add_kmer_edge!, src/edges.jl:58-116.  A window that leaves the gene sequence or holds an ambiguous base
    (seq4 not one-hot; `N` -> no table entry, k-mer value 0, :65-67) adds nothing."""
    G = len(gene_offset)
    N = len(nodeoffset)
    n = len(codes)
    gene_of_node = np.repeat(np.arange(G), np.diff(node_base))
    node_in_gene = np.arange(N) - node_base[gene_of_node] + 1
    eidx = np.arange(N) + gene_of_node                  # edge j of node j
    et = edgetype[eidx]
    seqlen = np.diff(np.append(gene_offset, n)).astype(np.int64)[gene_of_node]
    off = nodeoffset.astype(np.int64)
    base = gene_offset.astype(np.int64)[gene_of_node]
    pw = (4 ** np.arange(K - 1, -1, -1)).astype(np.uint64)
    ambcum = None
    if seq4 is not None:
        amb = ~np.isin(np.asarray(seq4), (1, 2, 4, 8))
        if amb.any():
            ambcum = np.concatenate([[0], np.cumsum(amb)])

    def kmers(start1):   # k-mer beginning at 1-based in-gene position start1
        ok = (start1 >= 1) & (start1 + K - 1 <= seqlen)
        pos = np.where(ok, base + start1 - 1, 0)
        if ambcum is not None:
            ok &= (ambcum[np.minimum(pos + K, n)] - ambcum[pos]) == 0
        v = np.zeros(N, np.uint64)
        for t in range(K):
            v += codes[np.minimum(pos + t, n - 1)].astype(np.uint64) * pw[t]
        return ok, v

    edgeleft = np.zeros(N + G, "<u8")
    edgeright = np.zeros(N + G, "<u8")
    out = []
    for left in (True, False):
        m = _is_edge(et, left)
        ok, v = kmers(off - K if left else off)
        sel = m & ok
        (edgeleft if left else edgeright)[eidx[sel]] = v[sel]
        order = np.lexsort((node_in_gene[sel], gene_of_node[sel], v[sel]))
        kv = v[sel][order]
        pairs = np.stack([gene_of_node[sel][order] + 1, node_in_gene[sel][order]], axis=1).astype("<u4")
        ptr_ = np.zeros(4 ** K + 1, np.int64)
        np.add.at(ptr_, kv.astype(np.int64) + 1, 1)
        out.append((np.cumsum(ptr_).astype("<u4"), pairs.reshape(-1)))
    return edgeleft, edgeright, out[0], out[1]


def make_index(n_genes=2000, K=9, seed=1, mean_exons=8.0, paralog_frac=0.02, threads=None,
               exon_median=130.0, utr_median=400.0, max_isoforms=6) -> FlatIndex:
    """Random gene models -> FlatIndex.  All genes are '+' strand (strand only changes nodecoord)."""
    rng = np.random.default_rng(seed)
    node_base, gene_offset = [0], []
    nodelen, nodecoord, edgetype = [], [], []
    iso_base, iso_ptr, iso_nodes = [0], [0], []
    text_len = 0
    coord = 1000
    for g in range(n_genes):
        m = 1 + rng.poisson(mean_exons - 1) if rng.random() > 0.05 else 1
        nl, nc, et = [], [], [SL]
        # units: each exon contributes 1..2 nodes; 'choices' records alternative features
        feats = []          # per exon: dict(nodes=[ids], kind, ...)
        for e in range(m):
            last, first = e == m - 1, e == 0
            L = int(np.clip(rng.lognormal(np.log(utr_median if (first or last) else exon_median), 0.6), 12, 4000))
            kind = "plain"
            r = rng.random()
            if m > 1 and not first and not last:
                kind = "alt5" if r < 0.12 else "alt3" if r < 0.24 else "plain"
            elif last and m > 1 and r < 0.2:
                kind = "apa"
            elif first and m > 1 and r < 0.15:
                kind = "alttss"
            ids = []
            if kind in ("alt5", "apa"):
                ext = int(rng.integers(3, 120))
                nl += [L, ext]
                nc += [coord, coord + L]
                ids = [len(nl) - 1, len(nl)]
                coord += L + ext
                inner = LL if kind == "alt5" else RS
            elif kind in ("alt3", "alttss"):
                ext = int(rng.integers(3, 120))
                nl += [ext, L]
                nc += [coord, coord + ext]
                ids = [len(nl) - 1, len(nl)]
                coord += L + ext
                inner = RR if kind == "alt3" else SL
            else:
                nl.append(L)
                nc.append(coord)
                ids = [len(nl)]
                coord += L
                inner = None
            if inner is not None:
                et.append(inner)
            feats.append(dict(kind=kind, ids=ids))
            if not last:
                # boundary to the next exon: intron (LR), retained intron (LL node RR), tandem alt first/last exon
                r = rng.random()
                if r < 0.05:
                    il = int(rng.integers(60, 300))
                    et.append(LL)
                    nl.append(il)
                    nc.append(coord)
                    coord += il
                    feats.append(dict(kind="ri", ids=[len(nl)]))
                    et.append(RR)
                elif first and r < 0.12 and m > 2:
                    et.append(LS)
                    coord += int(rng.integers(80, 3000))
                    feats[-1]["tandem_first"] = True
                elif e == m - 2 and r < 0.12 and m > 2:
                    et.append(SR)
                    coord += int(rng.integers(80, 3000))
                    feats[-1]["tandem_last"] = True
                else:
                    et.append(LR)
                    coord += int(rng.integers(80, 3000))
            else:
                et.append(RS)
        nn = len(nl)
        assert len(et) == nn + 1
        # ---- isoforms ----
        exons = [f for f in feats if f["kind"] != "ri"]
        ris = {id(f) for f in feats if f["kind"] == "ri"}
        niso = 1 + int(rng.integers(0, max_isoforms)) if m > 1 else 1
        paths = set()
        tandem_first = any(f.get("tandem_first") for f in feats)
        tandem_last = any(f.get("tandem_last") for f in feats)
        for t in range(niso * 2):
            if len(paths) >= niso:
                break
            default = t == 0
            nodes = []
            ex = list(exons)
            # tandem alternative first exon: use exon 0 (skip exon 1) or start at exon 1
            if tandem_first:
                ex = ex[1:] if (not default and rng.random() < 0.5) else [ex[0]] + ex[2:]
            if tandem_last:
                ex = ex[:-1] if (default or rng.random() < 0.5) else ex[:-2] + [ex[-1]]
            for i, f in enumerate(ex):
                internal = 0 < i < len(ex) - 1
                if internal and not default and rng.random() < 0.15:
                    continue        # cassette exon skipped
                k = f["kind"]
                alt = (not default) and rng.random() < 0.5
                if k == "plain":
                    nodes += f["ids"]
                elif k in ("alt5", "apa"):
                    nodes += f["ids"] if alt else f["ids"][:1]
                else:
                    nodes += f["ids"] if alt else f["ids"][1:]
            nodes = set(nodes)
            # retained introns: include when both flanking nodes are present
            for f in feats:
                if id(f) in ris and not default and rng.random() < 0.5:
                    i0 = f["ids"][0]
                    if (i0 - 1) in nodes and (i0 + 1) in nodes:
                        nodes.add(i0)
            paths.add(tuple(sorted(nodes)))
        for pth in sorted(paths):
            iso_nodes.extend(pth)
            iso_ptr.append(len(iso_nodes))
        iso_base.append(iso_base[-1] + len(paths))
        gene_offset.append(text_len)
        text_len += sum(nl)
        node_base.append(node_base[-1] + nn)
        nodelen += nl
        nodecoord += nc
        edgetype += et
        coord += int(rng.integers(1000, 20000))

    nodelen = np.array(nodelen, np.int64)
    node_base = np.array(node_base, np.int64)
    gene_offset = np.array(gene_offset, np.int64)
    G = n_genes
    gene_of_node = np.repeat(np.arange(G), np.diff(node_base))
    cum = np.cumsum(nodelen) - nodelen
    nodeoffset = cum - gene_offset[gene_of_node] + 1
    codes = rng.integers(0, 4, text_len, dtype=np.uint8)
    # paralog families: overwrite a gene with a diverged copy of an earlier gene's sequence (multi-mapping)
    npar = int(paralog_frac * G)
    glen = np.diff(np.append(gene_offset, text_len))
    for _ in range(npar):
        a, b = rng.integers(0, G, 2)
        if a == b:
            continue
        ln = int(min(glen[a], glen[b]))
        src = codes[gene_offset[a]:gene_offset[a] + ln].copy()
        mut = rng.random(ln) < rng.uniform(0.0, 0.03)
        src[mut] = (src[mut] + rng.integers(1, 4, int(mut.sum()))) % 4
        codes[gene_offset[b]:gene_offset[b] + ln] = src
    edgetype = np.array(edgetype, "u1")
    el, er, (lp, ln_), (rp, rn) = build_edges(K, gene_offset, node_base, nodeoffset, edgetype, codes)
    sa, bwt, sentinel = suffix_array(codes, threads)
    cnt = np.bincount(codes, minlength=4)
    count = np.array([1, 1 + cnt[0], 1 + cnt[0] + cnt[1], 1 + cnt[0] + cnt[1] + cnt[2]], np.int64)
    return FlatIndex(kmer=K, gene_offset=gene_offset, node_base=node_base, nodeoffset=nodeoffset,
                     nodecoord=np.array(nodecoord), nodelen=nodelen, edgetype=edgetype, edgeleft=el, edgeright=er,
                     seq4=(1 << codes).astype("u1"), bwt=bwt, sentinel=sentinel, count=count, sa=sa,
                     left_ptr=lp, left_nodes=ln_, right_ptr=rp, right_nodes=rn,
                     iso_base=np.array(iso_base), iso_node_ptr=np.array(iso_ptr), iso_nodes=np.array(iso_nodes, "<u4"),
                     gene_names=[f"G{g}" for g in range(G)], iso_names=[],
                     gene_strand=np.ones(G, "u1"))


# ------------------------------------------------------------------------------------------
# read simulation
# ------------------------------------------------------------------------------------------
class Transcriptome:
    """Sequences of node paths (by default the annotated isoforms): concatenated node sequences of each path.
    `paths` = (gene_of_path 0-based [P], ptr [P+1], nodes 1-based within the gene) samples reads from arbitrary paths,
    e.g. every exon-skipping combination of a stress gene, annotated or not."""

    def __init__(self, idx: FlatIndex, paths=None):
        codes = np.zeros(len(idx.seq4), np.uint8)
        for v, c in ((1, 0), (2, 1), (4, 2), (8, 3)):
            codes[idx.seq4 == v] = c
        if paths is None:
            gene_of_iso = np.repeat(np.arange(idx.n_genes), np.diff(idx.iso_base))
            ptr, nodes = np.asarray(idx.iso_node_ptr, np.int64), np.asarray(idx.iso_nodes, np.int64)
        else:
            gene_of_iso, ptr, nodes = (np.asarray(x, np.int64) for x in paths)
        I = len(ptr) - 1
        iso_of_entry = np.repeat(np.arange(I), np.diff(ptr))
        g = gene_of_iso[iso_of_entry]
        slot = np.asarray(idx.node_base, np.int64)[g] + nodes - 1
        start = np.asarray(idx.gene_offset, np.int64)[g] + np.asarray(idx.nodeoffset, np.int64)[slot] - 1
        ln = np.asarray(idx.nodelen, np.int64)[slot]
        tot = int(ln.sum())
        # expand (start, len) runs into positions
        run_start = np.cumsum(ln) - ln
        pos = np.arange(tot, dtype=np.int64) - np.repeat(run_start, ln) + np.repeat(start, ln)
        self.seq = codes[pos]
        iso_len = np.zeros(I, np.int64)
        np.add.at(iso_len, iso_of_entry, ln)
        self.iso_len = iso_len
        self.iso_start = np.cumsum(iso_len) - iso_len
        self.genome = codes


def simulate_reads(idx: FlatIndex, n_reads: int, readlen: int = 101, seed: int = 2, sub_rate: float = 0.005,
                   n_rate: float = 1e-4, random_frac: float = 0.02, lens=None, tx: Transcriptome | None = None,
                   sigma: float = 2.0, chunk: int = 1 << 20, paired: bool = False, frag_mean=250.0, frag_sd=50.0):
    """Single-end (or paired) reads sampled from annotated isoforms: log-normal abundances, uniform
    position, strand 50/50, substitutions at `sub_rate` with phred 2 ('#') at the substituted base and
    phred 40 elsewhere, N -> A at `n_rate` (src/reads.jl:9), plus a fraction of random-sequence reads.
    Returns ReadBatch (or a (fwd, rev) pair)."""
    rng = np.random.default_rng(seed)
    tx = tx or Transcriptome(idx)
    need = readlen if not paired else readlen
    ok = np.nonzero(tx.iso_len >= need)[0]
    w = rng.lognormal(0.0, sigma, len(ok)) * np.maximum(tx.iso_len[ok] - need + 1, 1)
    w /= w.sum()
    R = readlen
    out_codes = [np.zeros((n_reads, R), np.uint8)]
    out_qual = [np.zeros((n_reads, R), np.uint8)]
    if paired:
        out_codes.append(np.zeros((n_reads, R), np.uint8))
        out_qual.append(np.zeros((n_reads, R), np.uint8))
    ar = np.arange(R, dtype=np.int64)
    for lo in range(0, n_reads, chunk):
        hi = min(n_reads, lo + chunk)
        m = hi - lo
        iso = ok[rng.choice(len(ok), m, p=w)]
        L = tx.iso_len[iso]
        if paired:
            frag = np.clip(rng.normal(frag_mean, frag_sd, m).astype(np.int64), R, L)
        else:
            frag = np.full(m, R, np.int64)
        st = (rng.random(m) * (L - frag + 1)).astype(np.int64)
        base = tx.iso_start[iso] + st
        flip = rng.random(m) < 0.5
        pos_a, pos_b = base, base + frag - R
        for k in range(2 if paired else 1):
            fw = tx.seq[pos_a[:, None] + ar[None, :]]
            if paired:
                # sense fragment: mate 1 = start, mate 2 = revcomp(end); antisense: mate 1 = revcomp(end), mate 2 = start
                bw = 3 - tx.seq[pos_b[:, None] + ar[None, :]][:, ::-1]
                c = np.where(flip[:, None], bw, fw) if k == 0 else np.where(flip[:, None], fw, bw)
            else:
                c = np.where(flip[:, None], 3 - fw[:, ::-1], fw)
            c = np.ascontiguousarray(c)
            rnd = rng.random(m) < random_frac if k == 0 else rnd
            c[rnd] = rng.integers(0, 4, (int(rnd.sum()), R), dtype=np.uint8)
            q = np.full((m, R), 40, np.uint8)
            sub = rng.random((m, R)) < sub_rate
            c[sub] = (c[sub] + rng.integers(1, 4, int(sub.sum()), dtype=np.uint8)) % 4
            q[sub] = 2
            nn = rng.random((m, R)) < n_rate
            c[nn] = 0
            q[nn] = 2
            out_codes[k][lo:hi] = c
            out_qual[k][lo:hi] = q
    if lens is not None:
        lens = np.asarray(lens, "u1")
    batches = [ReadBatch.from_codes(c, q, lens) for c, q in zip(out_codes, out_qual)]
    return batches[0] if not paired else tuple(batches)


# ------------------------------------------------------------------------------------------
# BASELINE config 5: "complex-event PSI EM stress (high-entropy AS events)"
# ------------------------------------------------------------------------------------------
def make_stress_index(n_genes=2000, K=9, seed=5, cassettes=(6, 10), exon_len=(60, 140), threads=None):
    """Genes of one first exon, c consecutive cassette exons (c in `cassettes`) and one last exon, every boundary an intron
    (LR).  Two isoforms are annotated (all cassettes in / all out); the reads come from random subsets (stress_paths), so
    every skip junction is expressed and the events of a cassette node have 2^6 .. 2^10 = 64 .. 1024 paths, the bound below
    which the reference does not subsample at random (src/events.jl:234-238)."""
    rng = np.random.default_rng(seed)
    node_base, nodelen, nodecoord, edgetype = [0], [], [], []
    iso_base, iso_ptr, iso_nodes = [0], [0], []
    coord = 1000
    for g in range(n_genes):
        c = int(rng.integers(cassettes[0], cassettes[1] + 1))
        nn = c + 2
        nl = rng.integers(exon_len[0], exon_len[1] + 1, nn)
        nl[0] += 150
        nl[-1] += 150
        for j in range(nn):
            nodelen.append(int(nl[j]))
            nodecoord.append(coord)
            coord += int(nl[j]) + int(rng.integers(100, 2000))
        edgetype += [SL] + [LR] * (nn - 1) + [RS]
        for pth in (list(range(1, nn + 1)), [1, nn]):
            iso_nodes.extend(pth)
            iso_ptr.append(len(iso_nodes))
        iso_base.append(iso_base[-1] + 2)
        node_base.append(node_base[-1] + nn)
        coord += 5000
    nodelen = np.array(nodelen, np.int64)
    node_base = np.array(node_base, np.int64)
    G = n_genes
    gene_of_node = np.repeat(np.arange(G), np.diff(node_base))
    glen = np.zeros(G, np.int64)
    np.add.at(glen, gene_of_node, nodelen)
    gene_offset = np.cumsum(glen) - glen
    text_len = int(glen.sum())
    nodeoffset = (np.cumsum(nodelen) - nodelen) - gene_offset[gene_of_node] + 1
    codes = rng.integers(0, 4, text_len, dtype=np.uint8)
    edgetype = np.array(edgetype, "u1")
    el, er, (lp, ln_), (rp, rn) = build_edges(K, gene_offset, node_base, nodeoffset, edgetype, codes)
    sa, bwt, sentinel = suffix_array(codes, threads)
    cnt = np.bincount(codes, minlength=4)
    count = np.array([1, 1 + cnt[0], 1 + cnt[0] + cnt[1], 1 + cnt[0] + cnt[1] + cnt[2]], np.int64)
    return FlatIndex(kmer=K, gene_offset=gene_offset, node_base=node_base, nodeoffset=nodeoffset,
                     nodecoord=np.array(nodecoord), nodelen=nodelen, edgetype=edgetype, edgeleft=el, edgeright=er,
                     seq4=(1 << codes).astype("u1"), bwt=bwt, sentinel=sentinel, count=count, sa=sa,
                     left_ptr=lp, left_nodes=ln_, right_ptr=rp, right_nodes=rn,
                     iso_base=np.array(iso_base), iso_node_ptr=np.array(iso_ptr), iso_nodes=np.array(iso_nodes, "<u4"),
                     gene_names=[f"S{g}" for g in range(G)], iso_names=[], gene_strand=np.ones(G, "u1"))


def stress_paths(idx: FlatIndex, per_gene=48, seed=6):
    """Random exon-skipping combinations of every stress gene (first and last exon always in) as Transcriptome paths."""
    rng = np.random.default_rng(seed)
    gene, ptr, nodes = [], [0], []
    nb = np.asarray(idx.node_base, np.int64)
    for g in range(idx.n_genes):
        nn = int(nb[g + 1] - nb[g])
        for _ in range(per_gene):
            keep = rng.random(nn - 2) < rng.uniform(0.25, 0.75)
            pth = [1] + [j + 2 for j in range(nn - 2) if keep[j]] + [nn]
            gene.append(g)
            nodes.extend(pth)
            ptr.append(len(nodes))
    return np.array(gene), np.array(ptr), np.array(nodes)


def fixture_reads(idx: FlatIndex, n=100000, seed=20261019, lo=11, hi=21):
    """BASELINE config 1 as SURVEY.md section 8d restates it: reads of length lo..hi (the fixture's own read-length range) sampled
    from the annotated isoform paths of test/test_index.jls, strand 50/50, 1 % substitutions (phred '#' = 2 there, 'B' = 33
    elsewhere), first base N -> A with p = 0.1 (src/reads.jl:9)."""
    rng = np.random.default_rng(seed)
    tx = Transcriptome(idx)
    lens = rng.integers(lo, hi + 1, n)
    cands = np.nonzero(tx.iso_len >= lo)[0]
    iso = cands[rng.integers(0, len(cands), n)]
    lens = np.minimum(lens, tx.iso_len[iso]).astype(np.int64)
    st = (rng.random(n) * (tx.iso_len[iso] - lens + 1)).astype(np.int64)
    ar = np.arange(hi, dtype=np.int64)
    pos = tx.iso_start[iso][:, None] + st[:, None] + np.minimum(ar[None, :], lens[:, None] - 1)
    c = tx.seq[pos]
    valid = ar[None, :] < lens[:, None]
    flip = rng.random(n) < 0.5
    # reverse complement inside each read's own length
    rc_idx = np.clip(lens[:, None] - 1 - ar[None, :], 0, hi - 1)
    rc = 3 - np.take_along_axis(c, rc_idx, axis=1)
    c = np.where(flip[:, None], rc, c)
    q = np.full((n, hi), 33, np.uint8)
    sub = (rng.random((n, hi)) < 0.01) & valid
    c = np.where(sub, (c + rng.integers(1, 4, (n, hi))) % 4, c).astype(np.uint8)
    q[sub] = 2
    firstn = rng.random(n) < 0.1
    c[firstn, 0] = 0
    q[firstn, 0] = 2
    c[~valid] = 0
    return ReadBatch.from_codes(np.ascontiguousarray(c), q, lens.astype("u1"))


# ------------------------------------------------------------------------------------------
# FASTQ text of a read batch (the file -> result measurements and tests of the ingest path)
# ------------------------------------------------------------------------------------------
def fastq_bytes(rb: ReadBatch, lo=0, hi=None, qualoffset=33, tag="r") -> bytes:
    """4-line FASTQ records of reads [lo, hi) of a FIXED-length batch, names `<tag><9 digits>` (vectorised)."""
    hi = rb.n if hi is None else hi
    n = hi - lo
    R = int(rb.len[lo]) if n else 0
    assert n == 0 or (np.asarray(rb.len[lo:hi]) == R).all(), "fastq_bytes needs fixed-length reads"
    W = (R + 31) // 32
    so = np.asarray(rb.seq_off[lo:hi]).astype(np.int64)
    qo = np.asarray(rb.qual_off[lo:hi]).astype(np.int64)
    name_w = len(tag) + 10
    reclen = name_w + 1 + R + 3 + R + 1
    out = np.empty((n, reclen), np.uint8)
    out[:, 0] = ord("@")
    for k, ch in enumerate(tag.encode()):
        out[:, 1 + k] = ch
    idx = np.arange(lo, hi, dtype=np.int64)
    for d in range(9):
        out[:, name_w - 1 - d] = 48 + (idx // (10 ** d)) % 10
    out[:, name_w] = 10
    lut = np.frombuffer(b"ACGT", np.uint8)
    words = rb.seq2[(so[:, None] + np.arange(W)[None, :]).ravel()].reshape(n, W)
    for j in range(R):
        out[:, name_w + 1 + j] = lut[((words[:, j >> 5] >> np.uint64(2 * (j & 31))) & np.uint64(3)).astype(np.int64)]
    p = name_w + 1 + R
    out[:, p] = 10
    out[:, p + 1] = ord("+")
    out[:, p + 2] = 10
    out[:, p + 3:p + 3 + R] = rb.qual[(qo[:, None] + np.arange(R)[None, :]).ravel()].reshape(n, R) + np.uint8(qualoffset)
    out[:, p + 3 + R] = 10
    return out.tobytes()


def bgzf_compress(data: bytes, level=1, block=65280, threads=8) -> bytes:
    """Blocked gzip (BGZF, as bgzip / bcl2fastq write it): independent deflate blocks with their size in a 'BC' extra field."""
    import struct
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    def one(i):
        chunk = data[i:i + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        hdr = struct.pack("<BBBBIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, ord("B"), ord("C"), 2, len(comp) + 25)
        return hdr + comp + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(one, range(0, len(data), block)))
    eof = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    return b"".join(parts) + eof
