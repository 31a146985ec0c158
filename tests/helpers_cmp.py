import numpy as np


def results_equal(a, b, n, paired=False):
    """Bit-exact comparison of two AlignResult objects read by read.  Returns list of differing reads."""
    bad = []
    if not (np.array_equal(a.nhits, b.nhits)):
        bad = list(np.nonzero(a.nhits != b.nhits)[0][:20])
    # fast vectorised path when CSR layouts coincide
    same_layout = np.array_equal(a.hit_start, b.hit_start) and len(a.aln) == len(b.aln) and len(a.nodes) == len(b.nodes)
    if same_layout and not bad:
        ok = (a.aln.tobytes() == b.aln.tobytes()) and (a.nodes.tobytes() == b.nodes.tobytes()) and \
             np.array_equal(a.flipped * (a.nhits > 0), b.flipped * (b.nhits > 0))
        if ok and (not paired or a.aln_mate.tobytes() == b.aln_mate.tobytes()):
            return []
    for i in range(n):
        if a.read(i) != b.read(i) or (paired and a.read(i, True) != b.read(i, True)):
            bad.append(i)
        elif a.nhits[i] and a.flipped[i] != b.flipped[i]:
            bad.append(i)
        if len(bad) > 20:
            break
    return sorted(set(int(x) for x in bad))


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype.itemsize == b.dtype.itemsize and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def full_pipeline_parity(q, o, p, rb, mate=None, readlen=None, expect_overflow=False):
    """The whole path through the C ABI (`q`, a GraphLibQuant) against the oracle (`o`) on the same reads, bit for bit:
    statistics incl. the running mean_readlen, node / edge / circ counts, long-path and multi-map keyed stores, TPM /
    isoform counts / gene TPM / EM iterations, assign_ambig! values, and every event record.  Returns a small summary."""
    from oracle import Oracle, canonical_events
    q.reset()
    o.quant_reset()
    if mate is None:
        q.process_reads(p, rb)
        o.process_reads(p, rb)
    else:
        q.process_paired_reads(p, rb, mate)
        o.process_paired_reads(p, rb, mate)
    st, ost = q.stats(), o.stats()
    assert (st.total, st.mapped, st.multi, st.sum_readlen) == (ost["total"], ost["mapped"], ost["multi"], ost["sum_readlen"]), (st, ost)
    assert st.mean_readlen == ost["mean_readlen"], (st.mean_readlen, ost["mean_readlen"])
    if not expect_overflow:
        assert st.path_overflow == 0
    node, ek, ec, ck, cc = q.counts()
    assert np.array_equal(node.astype(float), o.node_counts())
    oek, oev = o.edges()
    gk, gv = np.concatenate([ek, ck]), np.concatenate([ec, cc]).astype(float)
    order = np.argsort(gk, kind="stable")
    assert np.array_equal(gk[order], oek) and np.array_equal(gv[order], oev)
    assert q.keyed(0) == {k: int(v) for k, v in o.keyed(0).items()}
    assert q.keyed(1) == {k: int(v) for k, v in o.keyed(1).items()}
    rl = readlen if readlen is not None else int(np.floor(ost["mean_readlen"]))
    it, tpm, cnt, gt = q.gene_em(rl, 1, 10000)
    oit, otpm, ocnt, ogt = o.em_tpm(rl, 1, 10000)
    assert it == oit, (it, oit)
    assert bits_equal(tpm, otpm) and bits_equal(cnt, ocnt) and bits_equal(gt, ogt)
    gnode, gek, gev, gck, gcv = q.assign_ambig()
    o.assign_ambig()
    oek2, oev2 = o.edges()
    gk2, gv2 = np.concatenate([gek, gck]), np.concatenate([gev, gcv])
    order = np.argsort(gk2, kind="stable")
    assert bits_equal(gnode, o.node_counts())
    assert np.array_equal(gk2[order], oek2) and bits_equal(gv2[order], oev2)
    got = canonical_events(*q.process_events_raw(rl))
    want = o.events_flat(rl)
    for name, x, y in zip(("hdr", "psi/total/iterations", "values", "path lengths", "path nodes"), got, want):
        assert bits_equal(x, y), f"event records differ in {name}"
    return dict(mapped=st.mapped, multi=st.multi, iterations=it, events=len(want[0]), hdr=want[0], dbl=want[1])
