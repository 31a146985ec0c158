"""Generator of tests/golden/psi_large_events.npz: EM problems of a few LARGE events (1000-1900 paths, 130-150 ambiguous sets,
60-140 k set members) as the library's own event construction (wq_events.h, checked against the oracle's process_events by
tests/test_host_stages_cpu.py) produces them for the titin-like gene of the hg19-scale annotation on the counts of an 80 M-read
job.  They are inputs only -- the tests solve them with the oracle and compare; nothing here is an expected output.

The source arrays were dumped once by running the host harness on simulated counts (about ten minutes of read simulation with the
oracle), `tmp_psi/big169.npz`, events in decreasing order of member count; this script only selects from that dump."""
import sys

import numpy as np


def subset(d, sel):
    pp, ap, app = d["path_ptr"], d["amb_ptr"], d["amb_path_ptr"]
    o = dict(path_ptr=[0], n_inc=[], count=[], length=[], amb_ptr=[0], amb_path_ptr=[0], amb_paths=[], amb_mult=[], is_tandem=[])
    for e in sel:
        p0, p1, a0, a1 = pp[e], pp[e + 1], ap[e], ap[e + 1]
        o["n_inc"].append(d["n_inc"][e]); o["count"].append(d["count"][p0:p1]); o["length"].append(d["length"][p0:p1])
        o["path_ptr"].append(o["path_ptr"][-1] + int(p1 - p0))
        for a in range(a0, a1):
            o["amb_path_ptr"].append(o["amb_path_ptr"][-1] + int(app[a + 1] - app[a]))
        o["amb_paths"].append(d["amb_paths"][app[a0]:app[a1]]); o["amb_mult"].append(d["amb_mult"][a0:a1])
        o["amb_ptr"].append(o["amb_ptr"][-1] + int(a1 - a0)); o["is_tandem"].append(d["is_tandem"][e])
    cat = lambda k, dt: np.concatenate(o[k]).astype(dt)
    return dict(path_ptr=np.array(o["path_ptr"], "<u4"), n_inc=np.array(o["n_inc"], "<u4"), count=cat("count", "<f8"), length=cat("length", "<f8"),
                amb_ptr=np.array(o["amb_ptr"], "<u4"), amb_path_ptr=np.array(o["amb_path_ptr"], "<u4"), amb_paths=cat("amb_paths", "<u2"),
                amb_mult=cat("amb_mult", "<f8"), is_tandem=np.array(o["is_tandem"], "u1"))


if __name__ == "__main__":
    d = np.load(sys.argv[1] if len(sys.argv) > 1 else "tmp_psi/big169.npz")
    out = subset(d, [0, 5, 60, 168])
    np.savez_compressed("tests/golden/psi_large_events.npz", **out)
    print({k: v.shape for k, v in out.items()})
