"""Timing of wq_em_psi_batch on LARGE events (tests/golden/psi_large_events.npz replicated to the ~24 events one rank of an 8-GPU
weak-scaling job holds, and to 48 / 96): block kernel (WQ_PSI_CLUSTER=0) against the cluster kernel with 4 / 8 blocks per event and
the library's own choice.  Wall time of the call (uploads + kernels + download), best of 3; the results of every mode are compared
with the block kernel's bit for bit."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import whippet_jl_b200 as wq  # noqa: E402
from test_psi_cluster_cpu import large_events  # noqa: E402

idx = wq.FlatIndex.load(os.path.join(ROOT, "tests", "golden", "test_index_flat.npz"))
q = wq.GraphLibQuant(idx, device=0)
base = large_events()
for reps in (6, 12, 24):
    problems = base * reps
    ref = None
    for mode in ("0", "4", "8", None):
        if mode is None:
            os.environ.pop("WQ_PSI_CLUSTER", None)
        else:
            os.environ["WQ_PSI_CLUSTER"] = mode
        best = 1e9
        for _ in range(3):
            t = time.perf_counter()
            got = q.em_psi_batch(problems, readlen=101)
            best = min(best, time.perf_counter() - t)
        sig = b"".join(p.tobytes() for p, _ in got) + bytes(int(it) % 256 for _, it in got)
        if ref is None:
            ref = sig
        print(f"{len(problems):3d} large events, WQ_PSI_CLUSTER={mode}: {best * 1e3:8.1f} ms  equal_to_block_kernel={sig == ref}", flush=True)
q.close()
