"""Timing of wq_em_psi_batch on LARGE events (tests/golden/psi_large_events.npz replicated to the ~24 events one rank of an 8-GPU
weak-scaling job holds, and to 48 / 96): block kernel (WQ_PSI_CLUSTER=0) against the cluster kernel with 4 / 8 blocks per event and
the library's own choice.  The kernel time is the library's per-class diagnostic (WQ_PSI_CLASS_TIMING: host clock around a stream
synchronisation after the launch, printed on stderr as "psi class 5"); the wall time of the Python call is dominated by building the
argument arrays and only given for completeness.  One more call per cluster mode prints the per-phase clocks of one block
(WQ_PSI_CL_TIMING).  WQ_PSI_CL_SPLIT=1 switches the lane-split shares phase on.  The results of every mode are compared with the block kernel's bit for bit."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
os.environ["WQ_PSI_CLASS_TIMING"] = "1"
import whippet_jl_b200 as wq  # noqa: E402
from test_psi_cluster_cpu import large_events  # noqa: E402

idx = wq.FlatIndex.load(os.path.join(ROOT, "tests", "golden", "test_index_flat.npz"))
q = wq.GraphLibQuant(idx, device=0)
base = large_events()
for reps in [int(x) for x in sys.argv[1:]] or [6, 12, 24]:
    problems = base * reps
    ref = None
    for mode, split in (("0", None), ("2", None), ("4", None), ("4", "1"), ("8", None), (None, None)):
        for name, val in (("WQ_PSI_CLUSTER", mode), ("WQ_PSI_CL_SPLIT", split)):
            if val is None:
                os.environ.pop(name, None)
            else:
                os.environ[name] = val
        q.em_psi_batch(problems[:4], readlen=101)
        print(f"--- {len(problems)} large events, WQ_PSI_CLUSTER={mode} WQ_PSI_CL_SPLIT={split}", file=sys.stderr, flush=True)
        t = time.perf_counter()
        got = q.em_psi_batch(problems, readlen=101)
        wall = time.perf_counter() - t
        sig = b"".join(p.tobytes() for p, _ in got) + bytes(int(it) % 256 for _, it in got)
        if ref is None:
            ref = sig
        print(f"{len(problems):3d} large events, WQ_PSI_CLUSTER={mode} WQ_PSI_CL_SPLIT={split}: call {wall * 1e3:8.1f} ms  equal_to_block_kernel={sig == ref}", flush=True)
        if mode in ("2", "4", "8") and split is None and reps == 6:
            os.environ["WQ_PSI_CL_TIMING"] = "1"
            q.em_psi_batch(problems, readlen=101)
            os.environ.pop("WQ_PSI_CL_TIMING")
q.close()
