"""Round trip of the bench index (refFlat-derived hg19-scale GraphLib: ~24 k genes, text 72.7 M) through the .jls writer, the .jls reader
and the converter to the native index file; prints sizes and times.  Host only.  Output kept in profiles/r02_jls_roundtrip_scale.txt."""
import os, sys, time, tempfile
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import numpy as np
import whippet_jl_b200 as wq
from whippet_jl_b200 import jls, convert
from wq_inputs import graphbuild

t = time.time(); idx = graphbuild.index_from_structure(graphbuild.load_structure(), K=9, seed=1); print(f"index built in {time.time()-t:.1f} s: G={idx.n_genes} N={idx.n_nodes} I={idx.n_isoforms} n={idx.text_len}")
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
p, w = os.path.join(d, "g.jls"), os.path.join(d, "g.wqx")
t = time.time(); out = jls.dump_jls(jls.graphlib_tree(idx)); open(p, "wb").write(out); print(f".jls written in {time.time()-t:.1f} s: {len(out)} bytes")
t = time.time(); r = jls.JLSReader(out); r.read_header(); lib = r.value(); print(f".jls parsed in {time.time()-t:.1f} s: {r.counter} back-reference slots, index type {lib.index.type}, consumed all bytes: {r.p == len(out)}")
del out, lib, r
t = time.time(); b = wq.FlatIndex.from_jls(p); print(f"FlatIndex.from_jls in {time.time()-t:.1f} s")
for k in wq.FlatIndex._ARRAYS:
    assert np.array_equal(getattr(idx, k), getattr(b, k)), k
print("every flat array equals the source index")
t = time.time(); wq.write_index(b, w); print(f"native index written in {time.time()-t:.1f} s: {os.path.getsize(w)} bytes")
for f in (p, w):
    os.remove(f)
os.rmdir(d)
