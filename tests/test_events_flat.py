"""The two canonical event forms agree: Oracle.events_flat (C++ export) and canonical_events applied to arrays in the
library's wq_events_out layout (here re-encoded from the oracle's per-event dicts, following include/whippet_b200.h).
The GPU parity tests and bench.py's parity_check compare the library's real arrays the same way."""
import numpy as np

import whippet_jl_b200 as wq
from whippet_jl_b200 import abi
from wq_inputs import synth
from oracle import Oracle, canonical_events


def library_layout(events):
    ev = np.zeros(len(events), abi.EVENT_DT)
    vals, pptr, pnodes = [], [0], []
    for i, e in enumerate(events):
        r = ev[i]
        r["gene"], r["node"], r["motif"], r["kind"], r["flags"] = e["gene"], e["node"], e["motif"], e["kind"], e["flags"]
        r["psi"], r["total"], r["iterations"] = e["psi"], e["total"], e["iterations"]
        r["n_utr"], r["n_inc"], r["n_exc"] = len(e["utr_psi"]), len(e["inc"]), len(e["exc"])
        r["value_start"], r["path_start"] = len(vals), len(pptr) - 1
        if e["kind"] == 2:
            vals += e["utr_psi"]
        else:
            if e["kind"] == 1 and e["inc"] and e["exc"]:
                vals += [v for _, v in e["inc"]] + [v for _, v in e["exc"]]
            for pth, _ in e["inc"] + e["exc"]:
                pnodes += list(pth)
                pptr.append(len(pnodes))
    return ev, np.array(vals or [0.0]), np.array(pptr, "<u8"), np.array(pnodes or [0], "<u4")


def test_canonical_forms_agree():
    idx = synth.make_index(n_genes=300, K=9, seed=4, paralog_frac=0.1)
    rb = synth.simulate_reads(idx, 40000, 101, seed=5)
    o = Oracle(idx)
    o.process_reads(wq.AlignParam(), rb)
    o.em_tpm(101)
    o.assign_ambig()
    flat = o.events_flat(101)
    dicts = o.process_events(101)
    got = canonical_events(*library_layout(dicts))
    assert len(dicts) == len(flat[0]) > 500 and (flat[0][:, 3] == 2).any() and (flat[0][:, 3] == 1).sum() > 100
    for a, b in zip(got, flat):
        assert a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))
