"""TEST INFRASTRUCTURE: a second, independent restatement of the two EM procedures of the path, written in plain
Python / NumPy float64 FROM THE JULIA SOURCE (not from oracle/whippet_oracle.cpp), so that a transcription slip shared by
the C++ oracle and the CUDA kernels would show up as a disagreement:

  calculate_tpm!   src/quant.jl:655-667        gene_em!        src/quant.jl:719-769      copy_isdone!  src/quant.jl:702-712
  calculate_psi!   src/events.jl:831-850       spliced_em!     src/events.jl:853-893     rec_tandem_em! src/events.jl:895-932
  divsignif!       src/events.jl:802-812       Base.round(x, digits=d) / Base.round(x, sigdigits=n)

Readings of the reference that cannot be checked without Julia are explicit here and named in tests/test_em_kat.py:
  * `@fastmath a * SCALING_FACTOR / rpk_sum` is evaluated left to right, `(a * 1e6) / rpk_sum` (no reassociation);
  * `sum(quant.tpm)` (Julia's blocked / SIMD sum, whose association depends on the machine) is taken as the declared
    canonical order of DESIGN.md section 2: an adjacent-pair tree over tiles of 1024, tiles folded left to right;
  * classes are visited in the order given by the caller (Julia iterates a Dict);
  * sums over the few paths of an event (`sum(egraph.psi) + sum(igraph.psi)`) run left to right.
Only tests/ imports this module.
"""
import math

import numpy as np

SCALING_FACTOR = 1_000_000.0


def round_digits(x, digits):
    """Base.round(x, digits=d) -> _round_digits: round(x * 10^d) / 10^d with ties to even; a non-finite result gives x back
    (digits > 0)."""
    with np.errstate(all="ignore"):              # IEEE semantics as in Julia: 10.0^310 = Inf, Inf / Inf = NaN
        x64 = np.float64(x)
        if digits >= 0:
            sc = np.float64(10.0) ** np.float64(digits)
            r = float(np.rint(x64 * sc) / sc)
        else:
            sc = np.float64(10.0) ** np.float64(-digits)
            r = float(np.rint(x64 / sc) * sc)
    if not math.isfinite(r):
        return x if digits > 0 else r
    return r


def round_sigdigits(x, sig):
    """Base.round(x, sigdigits=n) -> hidigit(x, 10) = 1 + floor(log10(|x|)) (0 for x == 0), then _round_digits(x, n - h)."""
    if not math.isfinite(x):
        return x
    h = 0 if x == 0 else 1 + math.floor(math.log10(abs(x)))
    return round_digits(x, sig - h)


def canonical_sum(v):
    """Declared canonical order of sum(quant.tpm): per tile of 1024 an adjacent-pair tree (v[t] + v[t + s] for t a multiple
    of 2s, s = 1, 2, 4, ...), tiles added left to right."""
    v = np.asarray(v, np.float64)
    total = 0.0
    for t in range(0, len(v), 1024):
        buf = np.zeros(1024)
        buf[:len(v[t:t + 1024])] = v[t:t + 1024]
        s = 1
        while s < 1024:
            buf[0::2 * s] = buf[0::2 * s] + buf[s::2 * s]
            s *= 2
        total = buf[0] if t == 0 else total + buf[0]
    return float(total)


def calculate_tpm(counts, lengths, readlen, sig):
    """calculate_tpm!: tpm_i = counts_i / max(length_i - readlen, 1.0); rpk_sum = max(sum(tpm), 1.0);
    tpm_i = round(tpm_i * 1e6 / rpk_sum, digits=sig) (sig > 0)."""
    tpm = np.array([c / max(float(int(l) - int(readlen)), 1.0) for c, l in zip(counts, lengths)], np.float64)
    rpk_sum = max(canonical_sum(tpm), 1.0)
    out = np.empty_like(tpm)
    for i in range(len(tpm)):
        v = (tpm[i] * SCALING_FACTOR) / rpk_sum
        out[i] = round_digits(float(v), sig) if sig > 0 else v
    return out


class EqClass:
    """IsoCompat / MultiCompat: class (1-based isoform ids), prop = 1/len each, prop_sum = 1, count, isdone."""

    def __init__(self, ids, count, isdone=False):
        self.ids = list(ids)
        self.prop = [1.0 / len(ids)] * len(ids) if ids else []
        self.prop_sum = 1.0
        self.count = float(count)
        self.isdone = bool(isdone)


def copy_isdone(dest, src, denominator, sig=4):
    """copy_isdone!: dest_i = round(src_i / denominator, digits=sig) (NaN -> 0); returns `nothing changed || denominator == 0`."""
    isdifferent = False
    for i in range(len(dest)):
        q = src[i] / denominator if denominator != 0 else (math.nan if src[i] == 0 else math.copysign(math.inf, src[i]))
        val = round_digits(q, sig) if math.isfinite(q) else q
        if val != dest[i]:                      # NaN != anything
            isdifferent = True
        dest[i] = 0.0 if math.isnan(val) else val
    return (not isdifferent) or denominator == 0


def gene_em(count, lengths, multi, classes, readlen, sig=1, maxit=10000, init_sig=1, trace=None):
    """bin/whippet-quant.jl:163-164: calculate_tpm!(quant, readlen=readlen) [its own default sig = 1], then
    gene_em!(quant, multi, sig, readlen, maxit).  `multi` and `classes` are lists of EqClass, visited in that order, multi-map
    classes first (src/quant.jl:756-761).  Returns (it, tpm, count); `trace` collects the TPM vector after every iteration."""
    count = np.array(count, np.float64)
    tpm = calculate_tpm(count, lengths, readlen, init_sig)
    count_temp = np.ones(len(count))
    tpm_temp = np.ones(len(count))
    prop_temp = [0.0] * 500
    it = 1

    def maximize_and_assign(eq, maximize):
        if eq.isdone:
            return
        if maximize:
            eq.prop_sum = 0.0
            for c, i in enumerate(eq.ids):
                cur = tpm[i - 1]
                prop_temp[c] = cur
                eq.prop_sum += cur
            eq.isdone = copy_isdone(eq.prop, prop_temp, eq.prop_sum)
        for c, i in enumerate(eq.ids):
            if eq.isdone:
                count[i - 1] += eq.prop[c] * eq.count
            count_temp[i - 1] += eq.prop[c] * eq.count

    while not np.array_equal(tpm_temp, tpm) and it < maxit:
        count_temp[:] = count
        tpm_temp[:] = tpm
        for eq in multi:
            maximize_and_assign(eq, it > 1)
        for eq in classes:
            maximize_and_assign(eq, it > 1)
        tpm = calculate_tpm(count_temp, lengths, readlen, sig)
        if trace is not None:
            trace.append(tpm.copy())
        it += 1
    return it, tpm, count


# ---------------------------------------------------------------------------------------------------- PSI
class Ambig:
    """AmbigCounts(paths, ones(n) / n, 1.0, multiplier) (src/events.jl:300-305, 428-430); paths are 1-based indices into the
    inclusion paths followed by the exclusion paths."""

    def __init__(self, paths, multiplier):
        self.paths = list(paths)
        self.prop = [1.0 / len(paths)] * len(paths)
        self.prop_sum = 1.0
        self.multiplier = float(multiplier)


def _divsignif(arr, divisor, sig):
    for i in range(len(arr)):
        q = _div(arr[i], divisor)
        arr[i] = round_sigdigits(q, sig) if sig > 0 else q


def _div(a, b):
    if b == 0:
        return math.nan if (a == 0 or math.isnan(a)) else math.copysign(math.inf, a)
    return a / b


def _spread(ambig, psi_of, count_temp, it):
    for ac in ambig:
        if it > 1:
            ac.prop_sum = 0.0
            for k, p in enumerate(ac.paths):
                prev = psi_of(p)
                ac.prop[k] = prev
                ac.prop_sum += prev
        for k, p in enumerate(ac.paths):
            prop = _div(ac.prop[k], ac.prop_sum)
            ac.prop[k] = prop
            count_temp[p - 1] += prop * ac.multiplier


def calculate_psi_spliced(ipsi, epsi, inc_length, exc_length, counts, sig):
    """calculate_psi!(igraph, egraph, counts; sig): psi = count / length, divided by sum(exclusion) + sum(inclusion),
    rounded to `sig` significant digits when sig > 0 (src/events.jl:831-841)."""
    ni = len(ipsi)
    for i in range(ni):
        ipsi[i] = _div(counts[i], inc_length[i])
    for i in range(len(epsi)):
        epsi[i] = _div(counts[ni + i], exc_length[i])
    cnt_sum = sum_lr(epsi) + sum_lr(ipsi)
    _divsignif(ipsi, cnt_sum, sig)
    _divsignif(epsi, cnt_sum, sig)


def spliced_em(inc_count, inc_length, exc_count, exc_length, ambig, sig=4, maxit=1500, trace=None):
    """The call site of _process_spliced (src/events.jl:733-736): psi = zeros; calculate_psi!(inc, exc, [inc.count; exc.count])
    WITHOUT rounding (sig defaults to 0); then spliced_em!(inc, exc, ambig, sig=4).  Returns (it, inc_psi, exc_psi)."""
    ni, ne = len(inc_count), len(exc_count)
    ipsi, epsi = [0.0] * ni, [0.0] * ne
    calculate_psi_spliced(ipsi, epsi, inc_length, exc_length, list(inc_count) + list(exc_count), 0)
    inc_temp, exc_temp = [0.0] * ni, [0.0] * ne
    count_temp = [1.0] * (ni + ne)
    it = 1
    while (not _same(inc_temp, ipsi) or not _same(epsi, exc_temp)) and it < maxit:
        count_temp[:ni] = inc_count
        count_temp[ni:] = exc_count
        inc_temp[:] = ipsi
        exc_temp[:] = epsi
        _spread(ambig, lambda p: ipsi[p - 1] if p <= ni else epsi[p - ni - 1], count_temp, it)
        calculate_psi_spliced(ipsi, epsi, inc_length, exc_length, count_temp, sig)
        if trace is not None:
            trace.append((list(ipsi), list(epsi)))
        it += 1
    return it, ipsi, epsi


def _same(a, b):
    """Vector `==` of Julia: element-wise IEEE equality (NaN is not equal to NaN)."""
    return len(a) == len(b) and all(x == y for x, y in zip(a, b))


def sum_lr(v):
    s = 0.0
    for x in v:
        s += x
    return s


def tandem_em(count, length, ambig, readlen, sig=4, maxit=1500):
    """rec_tandem_em! written as the loop its tail recursion is; calculate_psi!(pgraph, counts; sig, readlen)."""
    n = len(count)
    psi = [0.0] * n
    it = 1
    while True:
        count_temp = list(count)
        utr_temp = list(psi)
        _spread(ambig, lambda p: psi[p - 1], count_temp, it)
        for i in range(n):
            psi[i] = _div(count_temp[i], max(length[i] - readlen, 1))
        _divsignif(psi, sum_lr(psi), sig)
        if not _same(utr_temp, psi) and it < maxit:
            it += 1
            continue
        return it, psi
