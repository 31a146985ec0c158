"""Seeded inputs shared by the CPU and GPU parity tests."""
import numpy as np

import whippet_jl_b200 as wq
from wq_inputs import synth


def fixture_random_reads(idx, n=20000, seed=7):
    """Reads of length 11..21 on the reference's toy index: isoform-derived,
    genome-derived and random sequences, both strands, substitutions and low-quality bases."""
    rng = np.random.default_rng(seed)
    tx = synth.Transcriptome(idx)
    codes = np.zeros((n, 21), np.uint8)
    quals = np.full((n, 21), 33, np.uint8)
    lens = rng.integers(11, 22, n).astype(np.uint8)
    for i in range(n):
        L = int(lens[i])
        if rng.random() < 0.15:
            c = rng.integers(0, 4, L)
        else:
            cands = np.nonzero(tx.iso_len >= L)[0]
            if rng.random() < 0.5 and len(cands):
                iso = rng.choice(cands)
                st = rng.integers(0, tx.iso_len[iso] - L + 1)
                c = tx.seq[tx.iso_start[iso] + st: tx.iso_start[iso] + st + L].copy()
            else:
                st = rng.integers(0, len(tx.genome) - L + 1)
                c = tx.genome[st:st + L].copy()
            if rng.random() < 0.5:
                c = 3 - c[::-1]
        sub = rng.random(L) < 0.03
        c = np.where(sub, (c + rng.integers(1, 4, L)) % 4, c)
        codes[i, :L] = c
        quals[i, :L][sub] = rng.integers(0, 10, int(sub.sum()))
        lowq = rng.random(L) < 0.05
        quals[i, :L][lowq] = rng.integers(0, 6, int(lowq.sum()))
    return wq.ReadBatch.from_codes(codes, quals, lens)


def ragged_reads(idx, n=20000, seed=5, lo=30, hi=255):
    """Variable-length reads (down to shorter than the seed window) from a synthetic index."""
    rng = np.random.default_rng(seed)
    rb = synth.simulate_reads(idx, n, hi, seed=seed, sub_rate=0.01)
    lens = rng.integers(lo, hi + 1, n).astype(np.uint8)
    lens[:50] = rng.integers(1, 30, 50)
    codes = np.zeros((n, hi), np.uint8)
    W = (hi + 31) // 32
    words = rb.seq2.reshape(n, W)
    for j in range(hi):
        codes[:, j] = (words[:, j >> 5] >> np.uint64(2 * (j & 31))) & np.uint64(3)
    Q = len(rb.qual) // n
    quals = rb.qual.reshape(n, Q)[:, :hi]
    return wq.ReadBatch.from_codes(codes, quals, lens)


def same_strand_mate(r, R):
    """Turn a head-to-head mate 2 into a same-orientation one (for is_pair_rc = false)."""
    W = len(r.seq2) // r.n
    codes = np.zeros((r.n, R), np.uint8)
    words = r.seq2.reshape(r.n, W)
    for j in range(R):
        codes[:, j] = (words[:, j >> 5] >> np.uint64(2 * (j & 31))) & np.uint64(3)
    Q = len(r.qual) // r.n
    q = r.qual.reshape(r.n, Q)[:, :R]
    return wq.ReadBatch.from_codes(np.ascontiguousarray(3 - codes[:, ::-1]), np.ascontiguousarray(q[:, ::-1]))


PE_CASES = [(1, 300, 101, 5000, True), (2, 300, 50, 2500, True), (3, 200, 150, 1000, True), (4, 300, 76, 5000, False)]


def pe_case(seed, ng, R, pair_range, rc, n):
    idx = synth.make_index(n_genes=ng, K=9, seed=seed, paralog_frac=0.15)
    p = wq.AlignParam(seed_pair_range=pair_range, is_paired=True, is_pair_rc=rc)
    f, r = synth.simulate_reads(idx, n, R, seed=seed + 10, sub_rate=0.01, paired=True)
    if not rc:
        r = same_strand_mate(r, R)
    return idx, p, f, r


# ---- inputs for the branches round 1 left untested (VERDICT weak item 3) ----------------------------------------
def refseq_subset(n_genes=600, seed=3, n_frac=0.0, paralog_families=0, K=9):
    """The first `n_genes` loci of the refFlat-derived structure fixture ('+' and '-' strand genes, retained introns,
    zero-length nodes), optionally with runs of N in the graph sequence (ambiguity mask of wq_text_mism)."""
    from wq_inputs import graphbuild
    return graphbuild.index_from_structure(graphbuild.load_structure(), K=K, seed=seed, n_genes=n_genes, n_frac=n_frac,
                                           paralog_families=paralog_families)


def tiny_node_index(n_chain_genes=6, chain_nodes=70, node_len=6, seed=11, K=9):
    """Genes made of MANY tiny contiguous nodes (alternating alt-5' / alt-3' boundaries, LL / RR) next to ordinary genes:
    a 255-bp read crosses far more than WQ_MAX_PATH = 24 nodes there, which the library must report as path_overflow
    (the reference's path vectors are unbounded)."""
    from wq_inputs import graphbuild as gb
    rng = np.random.default_rng(seed)
    node_base, nodelen, nodecoord, edgetype = [0], [], [], []
    iso_base, iso_ptr, iso_nodes, strands, names = [0], [0], [], [], []
    coord = 1000
    for g in range(n_chain_genes):
        nn = chain_nodes
        for j in range(nn):
            nodelen.append(node_len + int(rng.integers(0, 3)))
            nodecoord.append(coord)
            coord += nodelen[-1]
        edgetype += [gb.SL] + [gb.LL if j % 2 == 0 else gb.RR for j in range(nn - 1)] + [gb.RS]
        iso_nodes += list(range(1, nn + 1))
        iso_ptr.append(len(iso_nodes))
        iso_base.append(iso_base[-1] + 1)
        node_base.append(node_base[-1] + nn)
        strands.append(1)
        names.append(f"chain{g}")
        coord += 5000
    S = dict(node_base=np.array(node_base, "<u4"), nodelen=np.array(nodelen, "<u4"), nodecoord=np.array(nodecoord, "<u4"),
             edgetype=np.array(edgetype, "u1"), iso_base=np.array(iso_base, "<u4"), iso_ptr=np.array(iso_ptr, "<u4"),
             iso_nodes=np.array(iso_nodes, "<u2"), strand=np.array(strands, "u1"), names=np.array(names))
    return gb.index_from_structure(S, K=K, seed=seed)
