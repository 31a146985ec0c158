"""Per-kernel device times of the alignment path on a device-resident batch (experiment helper)."""
import ctypes as C
import os
import sys

sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import numpy as np
import torch
import whippet_jl_b200 as wq
from whippet_jl_b200 import abi, lib
from wq_inputs import synth

genes = int(os.environ.get("WQ_BENCH_GENES", 40000))
nreads = int(os.environ.get("WQ_BENCH_READS", 4000000))
cache = f"/tmp/wq_cache_{genes}_{nreads}.npz"
idx = synth.make_index(n_genes=genes, K=9, seed=1)
rb = synth.simulate_reads(idx, nreads, 101, seed=2)
dev = torch.device("cuda", 0)
t = [torch.from_numpy(a).to(dev) for a in (rb.len, rb.seq_off.view(np.int64), rb.seq2.view(np.int64), rb.qual_off.view(np.int64), rb.qual)]
db = abi.ReadBatchC()
db.n_reads = rb.n
db.len = C.cast(t[0].data_ptr(), abi.u8p); db.seq_off = C.cast(t[1].data_ptr(), abi.u64p)
db.seq2 = C.cast(t[2].data_ptr(), abi.u64p); db.seq2_words = len(rb.seq2)
db.qual_off = C.cast(t[3].data_ptr(), abi.u64p); db.qual = C.cast(t[4].data_ptr(), abi.u8p); db.qual_bytes = len(rb.qual)
pc = wq.AlignParam().c()
for so in sys.argv[1:]:
    os.environ["WQ_LIB"] = so
    lib._lib = None
    q = wq.GraphLibQuant(idx)
    ms = np.zeros(3)
    for it in range(5):
        lib.check(q.L.wq_quant_reset(q.h))
        lib.check(q.L.wq_align_se_device(q.h, C.byref(pc), C.byref(db)))
        if it >= 2:
            ms += np.array(q.last_kernel_ms())
    st = q.stats()
    print(f"{os.path.basename(so):40s} seed {ms[0]/3:7.3f} extend {ms[1]/3:7.3f} resolve {ms[2]/3:7.3f} ms   mapped {st.mapped}", flush=True)
    q.close()
