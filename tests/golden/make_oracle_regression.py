"""Regression vectors of the ORACLE itself on the reference's fixture index (not reference outputs: Julia cannot run
here).  They freeze what the restatement computed when it was pinned against the reference's
known answers, so that a later edit of oracle/whippet_oracle.cpp that changes any quantity is noticed.

    python tests/golden/make_oracle_regression.py        # rewrites tests/golden/oracle_fixture_regression.json
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import whippet_jl_b200 as wq  # noqa: E402
from data_gen import fixture_random_reads  # noqa: E402
from oracle import Oracle  # noqa: E402


def compute():
    idx = wq.FlatIndex.load(os.path.join(HERE, "test_index_flat.npz"))
    p = wq.AlignParam(1, 2, 4, 4, 4, 5, 1, 2, 1000, 0.05, 0.7, False, False, True, False, True)     # test/runtests.jl:283-284
    rb = fixture_random_reads(idx, 30000)
    o = Oracle(idx)
    o.process_reads(p, rb)
    st = o.stats()
    ek, ev = o.edges()
    it, tpm, cnt, gt = o.em_tpm(15)
    o.assign_ambig()
    ev2 = o.process_events(15)
    return dict(
        reads=int(rb.n), stats={k: int(st[k]) for k in ("total", "mapped", "multi", "sum_readlen")},
        edges=[[int(k), float(v)] for k, v in zip(ek, ev)],
        em_iterations=int(it), tpm=[float(x) for x in tpm], isoform_counts=[float(x) for x in cnt], gene_tpm=[float(x) for x in gt],
        events=[dict(gene=e["gene"], node=e["node"], motif=e["motif"], kind=e["kind"], psi=e["psi"], total=e["total"],
                     iterations=e["iterations"], inc=[[list(pp), v] for pp, v in e["inc"]], exc=[[list(pp), v] for pp, v in e["exc"]],
                     utr_psi=e["utr_psi"]) for e in ev2])


if __name__ == "__main__":
    out = compute()
    json.dump(out, open(os.path.join(HERE, "oracle_fixture_regression.json"), "w"), indent=0)
    print("wrote", len(out["events"]), "events,", len(out["tpm"]), "isoforms")
