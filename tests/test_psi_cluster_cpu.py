"""The cluster variant of the PSI EM kernel (wq_k_em_psi_cl, csrc/wq_em.cuh: one large event shared by 2 / 4 / 8 thread blocks) on the
CPU: its phases are __host__ __device__ functions, and tests/emu runs them for every block and thread in turn against the oracle's
spliced_em! / rec_tandem_em! (src/events.jl:853-932).  This pins the ownership maps, the inverted index, the order of every
addition and the convergence flags; the barriers and the distributed-shared-memory writes are covered by the -m gpu tests."""
import os

import numpy as np
import pytest

from emu.host_emu import psi_cluster
from oracle import em_psi

HERE = os.path.dirname(os.path.abspath(__file__))


def large_events():
    d = np.load(os.path.join(HERE, "golden", "psi_large_events.npz"))
    pp, ap, app = d["path_ptr"], d["amb_ptr"], d["amb_path_ptr"]
    out = []
    for e in range(len(pp) - 1):
        sets = [d["amb_paths"][app[a]:app[a + 1]].astype(np.uint32) for a in range(ap[e], ap[e + 1])]
        out.append((int(d["n_inc"][e]), d["count"][pp[e]:pp[e + 1]], d["length"][pp[e]:pp[e + 1]], sets, d["amb_mult"][ap[e]:ap[e + 1]], bool(d["is_tandem"][e])))
    return out


def random_problems(seed, sizes):
    rng = np.random.default_rng(seed)
    problems = []
    for np_ in sizes:
        for tandem in (False, True):
            ni = np_ if tandem else int(rng.integers(1, np_))
            cnt = np.where(rng.random(np_) < 0.3, 0.0, np.round(rng.gamma(1.0, 30.0, np_)))
            ln = np.round(rng.uniform(2, 400 if tandem else 12, np_))
            sets, mult = [], []
            for _ in range(int(rng.integers(0, min(2 * np_, 40) + 1))):
                k = int(rng.integers(2, min(np_, 12) + 1))
                a = sorted(int(x) for x in rng.choice(np_, size=k, replace=False))
                if a in sets:
                    continue
                sets.append(a); mult.append(float(np.round(rng.gamma(1.0, 20.0), 1 if rng.random() < 0.3 else 0) + 1))
            problems.append((ni, cnt, ln, sets, mult, tandem))
    return problems


def check(problems, got, readlen, maxit):
    capped = 0
    for (ni, cnt, ln, sets, mult, tandem), (psi, it) in zip(problems, got):
        want, wit = em_psi(ni, cnt, ln, sets, mult, tandem, readlen, maxit=maxit)
        assert it == wit, (len(cnt), tandem, it, wit)
        assert psi.tobytes() == want.tobytes(), (len(cnt), tandem)
        capped += it == maxit
    return capped


@pytest.mark.parametrize("blocks", [2, 4, 8])
def test_cluster_phases_equal_oracle_small_and_medium_events(blocks):
    """2 to 400 paths, spliced and tandem, converging events and ones that decay to the 1500-sweep cap; 64 threads per block so that
    the strided loops wrap; path counts around the group size of the ownership map (32) and its multiples."""
    problems = random_problems(5, [2, 3, 5, 31, 32, 33, 64, 65, 100, 255, 256, 257, 400])
    got, need = psi_cluster(problems, blocks, 64, readlen=101)
    assert got is not None
    capped = check(problems, got, 101, 1500)
    assert 0 < capped < len(problems)


@pytest.mark.parametrize("blocks", [4, 8])
def test_cluster_phases_equal_oracle_large_events(blocks):
    """The events the kernel exists for (1000-1900 paths, up to 140 k set members), 40 sweeps; the first one also as a tandem event."""
    problems = large_events()
    ni, cnt, ln, sets, mult, _ = problems[1]
    problems.append((len(cnt), cnt, ln + 150.0, sets, mult, True))
    got, need = psi_cluster(problems, blocks, 96, readlen=101, maxit=40)
    assert got is not None and need <= 227 * 1024
    assert check(problems, got, 101, 40) == len(problems)


def test_cluster_layout_limits():
    """What the host refuses (the library then keeps the events on the block kernel): a layout beyond the 227 KB of shared memory a
    block can have -- a synthetic event of 20 000 paths in 40 sets on two blocks -- and member lists that are not in ascending path
    order (the bit-mask walk of phase 1 would add them in another order than the reference; a path twice would also race in the
    inverted-index build).  Two blocks do hold a 1900-path event."""
    problems = large_events()[:1]
    got, need = psi_cluster(problems, 2, 64, readlen=101, maxit=3)
    assert got is not None and need <= 227 * 1024
    assert check(problems, got, 101, 3) == 1
    n = 20000
    huge = (n // 2, np.ones(n), np.full(n, 100.0), [np.arange(k, n, 2, dtype=np.uint32) for k in range(40)], np.ones(40), False)
    got, need = psi_cluster([huge], 2, 64, readlen=101, maxit=2)
    assert got is None and need > 227 * 1024
    ni, cnt, ln, sets, mult, t = random_problems(9, [40])[0]
    for bad in ([list(sets[0]) + [sets[0][0]]], [list(sets[0])[::-1]]):
        got, _ = psi_cluster([(ni, cnt, ln, bad + sets[1:], mult, t)], 4, 64, readlen=101)
        assert got is None
