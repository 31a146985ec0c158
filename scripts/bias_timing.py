"""Cost of --biascorrect on the bench index: process_reads + the quant stage with DefaultCounter and with JointBiasCounter (experiment helper)."""
import os, sys, time
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import numpy as np
import whippet_jl_b200 as wq
from wq_inputs import synth, graphbuild
n = int(os.environ.get("WQ_BENCH_READS", 4000000))
idx = graphbuild.index_from_structure(graphbuild.load_structure(), K=9, seed=1)
rb = synth.simulate_reads(idx, n, 101, seed=2)
q = wq.GraphLibQuant(idx)
p = wq.AlignParam()
for bias in (False, True, False, True):
    q.set_biascorrect(bias)
    t0 = time.perf_counter(); st = q.process_reads(p, rb); t1 = time.perf_counter()
    it, tpm, cnt, gt = q.gene_em(101); t2 = time.perf_counter()
    q.assign_ambig_inplace(); ev = q.process_events_view(101); t3 = time.perf_counter()
    print(f"bias={bias}: process_reads {1e3*(t1-t0):.1f} ms ({n/(t1-t0)/1e6:.1f} M reads/s from pageable host memory), classes+EM {1e3*(t2-t1):.1f} ms ({it} it), assign+events {1e3*(t3-t2):.1f} ms, mapped {st.mapped}", flush=True)
