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
