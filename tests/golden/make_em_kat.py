"""Known-answer vectors for the two EM procedures, worked in EXACT rational arithmetic (fractions.Fraction) with explicit
round-half-even at the decimal places / significant digits the reference rounds at -- no floating point, no code shared with
oracle/ or the library.  Writes tests/golden/em_kat.json; tests/test_em_kat.py checks the oracle (C++), the NumPy restatement
(oracle/em_restatement.py) and the CUDA kernels against it.

    python tests/golden/make_em_kat.py

The procedures (src/quant.jl:655-769, src/events.jl:802-932) are deterministic sequences of divisions and decimal roundings;
the cases are chosen so that no intermediate value sits on a rounding boundary, where exact and binary64 arithmetic could
round differently.
"""
import json
import os
from fractions import Fraction as F

HERE = os.path.dirname(os.path.abspath(__file__))


def rnd(x: F, digits: int) -> F:
    """round to `digits` decimals, ties to even, on an exact rational"""
    sc = F(10) ** digits
    y = x * sc
    fl = y.numerator // y.denominator
    rem = y - fl
    if rem > F(1, 2) or (rem == F(1, 2) and fl % 2 == 1):
        fl += 1
    return F(fl) / sc


def rnd_sig(x: F, sig: int) -> F:
    if x == 0:
        return x
    h, a = 0, abs(x)             # hidigit: 1 + floor(log10 |x|)
    while a >= 10:
        a /= 10
        h += 1
    while a < 1:
        a *= 10
        h -= 1
    return rnd(x, sig - (h + 1))


def tpm_of(counts, eff, sig):
    raw = [c / e for c, e in zip(counts, eff)]
    s = max(sum(raw), F(1))
    return [rnd(r * 1_000_000 / s, sig) for r in raw]


def transcript_em(count, eff, classes, sig=1, maxit=10000, steps=None):
    """gene_em! on exact rationals; classes = [(ids 1-based, count)], all not done at the start."""
    count = list(count)
    tpm = tpm_of(count, eff, 1)
    cls = [dict(ids=ids, prop=[F(1, len(ids))] * len(ids), count=F(c), done=False) for ids, c in classes]
    tpm_temp, it, trace = None, 1, []
    while tpm_temp != tpm and it < maxit:
        count_temp, tpm_temp = list(count), list(tpm)
        for c in cls:
            if c["done"]:
                continue
            if it > 1:
                s = sum(tpm[i - 1] for i in c["ids"])
                new = [rnd(tpm[i - 1] / s, 4) for i in c["ids"]]
                c["done"] = new == c["prop"]
                c["prop"] = new
            for k, i in enumerate(c["ids"]):
                if c["done"]:
                    count[i - 1] += c["prop"][k] * c["count"]
                count_temp[i - 1] += c["prop"][k] * c["count"]
        tpm = tpm_of(count_temp, eff, sig)
        trace.append(list(tpm))
        it += 1
    return it, tpm, count, trace


def psi_em(inc, exc, ambig, sig=4, maxit=1500):
    """spliced_em! as _process_spliced calls it; inc / exc = [(count, length)], ambig = [(paths 1-based, multiplier)]."""
    ni = len(inc)
    cnt = [F(c) for c, _ in inc + exc]
    ln = [F(l) for _, l in inc + exc]

    def psi_from(counts, s):
        raw = [c / l for c, l in zip(counts, ln)]
        tot = sum(raw[ni:]) + sum(raw[:ni])
        return [rnd_sig(r / tot, s) if s else r / tot for r in raw]
    psi = psi_from(cnt, 0)
    prev, it, trace = None, 1, []
    props = [[F(1, len(p))] * len(p) for p, _ in ambig]
    while prev != psi and it < maxit:
        ct, prev = list(cnt), list(psi)
        for a, (paths, mult) in enumerate(ambig):
            if it > 1:
                props[a] = [psi[p - 1] for p in paths]
            s = sum(props[a])
            props[a] = [x / s for x in props[a]]
            for k, p in enumerate(paths):
                ct[p - 1] += props[a][k] * F(mult)
        psi = psi_from(ct, sig)
        trace.append(list(psi))
        it += 1
    return it, psi, trace


def fl(v):
    return [float(x) for x in v]


def main():
    out = {"note": "exact rational arithmetic, see make_em_kat.py; floats are the binary64 values nearest to the exact decimals"}
    # ---- transcript EM: 3 isoforms, effective lengths (length - readlen) 1000 / 500 / 2000, one class over {1,2}, one over {2,3}
    lengths, readlen = [1100, 600, 2100], 100
    eff = [F(l - readlen) for l in lengths]
    count = [F(10000), F(5000), F(8000)]
    classes = [([1, 2], 30000), ([2, 3], 12000)]
    it, tpm, cnt, trace = transcript_em(count, eff, classes)
    out["tpm"] = dict(lengths=lengths, readlen=readlen, count=fl(count), classes=[[ids, c] for ids, c in classes],
                      tpm_after_iteration_1=fl(trace[0]), tpm_after_iteration_2=fl(trace[1]), tpm_after_iteration_3=fl(trace[2]),
                      iterations=it, tpm_final=fl(tpm), count_final=fl(cnt))
    # ---- PSI: 2 paths (1 inclusion, 1 exclusion), one ambiguous set over both
    it, psi, trace = psi_em([(10, 2)], [(5, 1)], [([1, 2], 6)])
    out["psi2"] = dict(inc=[[10, 2]], exc=[[5, 1]], ambig=[[[1, 2], 6]], psi_after_sweep_1=fl(trace[0]), psi_after_sweep_2=fl(trace[1]),
                       iterations=it, psi_final=fl(psi))
    # ---- PSI: 3 paths (2 inclusion, 1 exclusion), one ambiguous set over paths 1 and 3, one over 1 and 2
    it, psi, trace = psi_em([(12, 3), (4, 2)], [(9, 1)], [([1, 3], 7), ([1, 2], 3)])
    out["psi3"] = dict(inc=[[12, 3], [4, 2]], exc=[[9, 1]], ambig=[[[1, 3], 7], [[1, 2], 3]], psi_after_sweep_1=fl(trace[0]),
                       psi_after_sweep_2=fl(trace[1]), iterations=it, psi_final=fl(psi))
    json.dump(out, open(os.path.join(HERE, "em_kat.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
