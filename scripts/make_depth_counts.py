"""Counts of deep jobs over the hg19-scale annotation (provenance of tmp_psi/counts_depth80M.npz and of tests/golden/psi_large_events.npz):
the oracle aligns 8 x 10 M simulated reads (seeds as the ranks of an 8-GPU weak-scaling bench use them) and the node / edge counts are saved after
every 10 M.  CPU only, about twelve minutes.  Output: /tmp/evprof_counts_depth{10..80}M.npz (copy the 80M one to tmp_psi/counts_depth80M.npz)."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import numpy as np
import whippet_jl_b200 as wq
from wq_inputs import synth, graphbuild
from oracle import Oracle
idx = graphbuild.index_from_structure(graphbuild.load_structure(), K=9, seed=1)
o = Oracle(idx); o.quant_reset()
tx = synth.Transcriptome(idx)
for r in range(8):
    t=time.time()
    rb = synth.simulate_reads(idx, 10_000_000, 101, seed=2 + 1000*r, tx=tx)
    o.process_reads(wq.AlignParam(), rb)
    print("rank", r, time.time()-t, flush=True)
    node=o.node_counts(); ek,ec=o.edges()
    np.savez(f'/tmp/evprof_counts_depth{(r+1)*10}M.npz', node=node, ek=ek, ec=ec)
