"""The quant stage of ONE rank of an 8-GPU weak-scaling job (80 M reads over the hg19-scale annotation), emulated on one GPU: the
node / edge counts of such a job (tmp_psi/counts_depth80M.npz: the oracle's counts of 80 M simulated reads, made by a CPU script;
not in the repository) are put into a handle through wq_counts_merge, and the stage runs as part k of 8 with 4 host threads, once
with the large PSI events on the block kernel (WQ_PSI_CLUSTER=0) and once as the library chooses (cluster kernel).
Prints wall milliseconds of process_events per part (event build on the host threads, not overlapped with anything here, + PSI EM);
the transcript EM is replicated on every rank and not part of the figure."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
os.environ.setdefault("WQ_HOST_THREADS", "4")
import whippet_jl_b200 as wq  # noqa: E402
from whippet_jl_b200 import abi, lib  # noqa: E402
from wq_inputs import graphbuild  # noqa: E402

src = os.path.join(ROOT, "tmp_psi", "counts_depth80M.npz")
if not os.path.exists(src):
    print("no", src)
    sys.exit(0)
z = np.load(src)
idx = graphbuild.index_from_structure(graphbuild.load_structure(), K=9, seed=1)
q = wq.GraphLibQuant(idx, device=0)
node, ek, ec = np.ascontiguousarray(z["node"], "<u8"), np.ascontiguousarray(z["ek"], "<u8"), np.ascontiguousarray(z["ec"], "<u8")
def inject():
    """a fresh count state holding the saved counts (gene_em! mutates the state, so every run starts from here)"""
    q.reset()
    c = abi.CountsC()
    c.n_nodes, c.node_count = len(node), abi.ptr(node, abi.u64p)
    c.n_edge, c.edge_key, c.edge_count = len(ek), abi.ptr(ek, abi.u64p), abi.ptr(ec, abi.u64p)
    c.n_circ, c.circ_key, c.circ_count = 0, None, None
    lib.check(q.L.wq_counts_merge(q.h, C.byref(c), None, None))


nparts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for mode in ("0", None):
    if mode is None:
        os.environ.pop("WQ_PSI_CLUSTER", None)
    else:
        os.environ["WQ_PSI_CLUSTER"] = mode
    for rep in range(2):
        out = []
        for k in range(nparts):
            inject()
            q.set_gene_partition(0, 1)
            q.gene_em(101, 1, 10000)                     # complete class set (on a real job: wq_classes_sync), EM replicated
            q.assign_ambig_inplace()
            q.set_gene_partition(k, nparts)              # drops the event batches built under the EM kernel: the part's build is timed below
            t = time.perf_counter()
            ev = q.process_events_view(101)
            out.append((time.perf_counter() - t, len(ev[0])))
        print(f"parts of {nparts}, WQ_PSI_CLUSTER={mode}, pass {rep}: ms per part {[round(a * 1e3, 1) for a, _ in out]} records {[n for _, n in out]}", flush=True)
inject()
q.set_gene_partition(0, 1)
t = time.perf_counter(); q.gene_em(101, 1, 10000); q.assign_ambig_inplace(); ev = q.process_events_view(101)
print(f"unpartitioned: {(time.perf_counter() - t) * 1e3:.1f} ms, {len(ev[0])} records")
q.close()
