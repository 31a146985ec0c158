"""EM kernel timing for several cooperative grid sizes (experiment helper)."""
import os, sys, time
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import numpy as np
import whippet_jl_b200 as wq
from wq_inputs import synth
idx = synth.make_index(n_genes=int(os.environ.get("WQ_BENCH_GENES", 40000)), K=9, seed=1)
rb = synth.simulate_reads(idx, int(os.environ.get("WQ_BENCH_READS", 4000000)), 101, seed=2)
q = wq.GraphLibQuant(idx)
p = wq.AlignParam()
ref = None
for blocks in sys.argv[1:]:
    os.environ["WQ_EM_BLOCKS"] = blocks
    os.environ["WQ_PROFILE"] = "1"
    q.reset(); q.process_reads(p, rb)
    t = time.perf_counter(); r = q.gene_em(101); dt = time.perf_counter() - t
    if ref is None: ref = r
    print(f"blocks {blocks}: gene_em {dt*1000:.1f} ms, it {r[0]}, same {all(np.array_equal(a,b) for a,b in zip(ref[1:], r[1:]))}", flush=True)
