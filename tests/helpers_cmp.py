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
