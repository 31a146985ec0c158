"""TEST INFRASTRUCTURE ONLY: ctypes driver of tests/emu/wq_emu.cpp (kernel bodies run on the CPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

import whippet_jl_b200 as wq
from whippet_jl_b200 import abi
from oracle import AlignResult

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwq_emu.so")
_CSRC = os.path.join(_HERE, "..", "..", "whippet.jl_b200", "csrc")


def build():
    srcs = [os.path.join(_HERE, "wq_emu.cpp")] + [os.path.join(_CSRC, f) for f in
                                                   ("wq_kernels.cuh", "wq_align.cuh", "wq_types.h", "wq_host_index.h", "wq_bias.cuh")]
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < max(map(os.path.getmtime, srcs)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas",
                               "-ffp-contract=off", "-o", _SO, srcs[0]])
    return _SO


class Counters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("total", "mapped", "multi", "sum_readlen", "path_overflow", "edge_full",
                                          "keyed_full", "pool_overflow")]


class Emu:
    def __init__(self, index: wq.FlatIndex):
        build()
        self.L = C.CDLL(_SO)
        self.L.emu_create.restype = C.c_void_p
        self.L.emu_result_sizes.restype = C.c_uint64
        self.L.emu_node_counts.restype = C.POINTER(C.c_uint64)
        self.L.emu_counters.restype = C.POINTER(Counters)
        self.index = index
        self._d = index.desc()
        self.h = C.c_void_p(self.L.emu_create(C.byref(self._d)))
        assert self.h

    def align_se(self, param, reads):
        p, b = param.c(), reads.c()
        rc = self.L.emu_align_se(self.h, C.byref(p), C.byref(b))
        assert rc == 0, rc
        nn = C.c_uint64()
        na = self.L.emu_result_sizes(self.h, C.byref(nn))
        n = reads.n
        hs, nh, fl = np.zeros(n, "<u4"), np.zeros(n, "u1"), np.zeros(n, "u1")
        aln, nodes = np.zeros(na, abi.ALIGNMENT_DT), np.zeros(nn.value, abi.ALIGN_NODE_DT)
        v = lambda a: a.ctypes.data_as(C.c_void_p)
        self.L.emu_result_copy(self.h, v(hs), v(nh), v(fl), v(aln), v(nodes))
        return AlignResult(hs, nh, fl, aln, nodes)

    def align_pe(self, param, fwd, rev):
        p, f, r = param.c(), fwd.c(), rev.c()
        rc = self.L.emu_align_pe(self.h, C.byref(p), C.byref(f), C.byref(r))
        assert rc == 0, rc
        nn = C.c_uint64()
        na = self.L.emu_result_sizes(self.h, C.byref(nn))
        n = fwd.n
        hs, nh, fl = np.zeros(n, "<u4"), np.zeros(n, "u1"), np.zeros(n, "u1")
        aln, alm, nodes = np.zeros(na, abi.ALIGNMENT_DT), np.zeros(na, abi.ALIGNMENT_DT), np.zeros(nn.value, abi.ALIGN_NODE_DT)
        v = lambda a: a.ctypes.data_as(C.c_void_p)
        self.L.emu_result_copy_pe(self.h, v(hs), v(nh), v(fl), v(aln), v(alm), v(nodes))
        return AlignResult(hs, nh, fl, aln, nodes, alm)

    def ext_stats(self):
        """(deferred tasks, parks, unparks, rollbacks, nested extensions, rollbacks below the checkpoint) since the library was loaded."""
        self.L.emu_deferred.restype = C.c_uint64
        self.L.emu_ext_stats.restype = C.POINTER(C.c_uint64)
        st = self.L.emu_ext_stats()
        return dict(deferred=int(self.L.emu_deferred(self.h)), parks=int(st[0]), unparks=int(st[1]), rollbacks=int(st[2]),
                    nested=int(st[3]), rollbacks_below=int(st[4]))

    def seeds(self, n):
        out = np.zeros((n, 4), np.int64)
        self.L.emu_seeds(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def node_counts(self):
        return np.ctypeslib.as_array(self.L.emu_node_counts(self.h), (self.index.n_nodes,)).copy()

    def counters(self):
        c = self.L.emu_counters(self.h).contents
        return {k: getattr(c, k) for k, _ in Counters._fields_}

    def process_reads(self, param, reads):
        p, b = param.c(), reads.c()
        rc = self.L.emu_align_se(self.h, C.byref(p), C.byref(b))
        assert rc == 0, rc

    # ---- --biascorrect: per-counter values in the oracle's form ----
    def set_bias(self, on=True):
        self.L.emu_set_bias(self.h, int(bool(on)))

    def bias_values(self):
        """After all reads: ({node: (p, g)} as arrays, {edge key: (p, g)}, {long key: (p, g)}, {multi key: (p, g)}, fore, back, short) with
        p = the count after primer_adjust!, g = the gc_adjust! factor (< 0: none)."""
        self.L.emu_bias_finish.restype = C.c_uint64
        short = int(self.L.emu_bias_finish(self.h))
        v = lambda a: a.ctypes.data_as(C.c_void_p)
        N = self.index.n_nodes
        np_, ng = np.zeros(N), np.zeros(N)
        self.L.emu_bias_nodes(self.h, v(np_), v(ng))
        self.L.emu_edges.restype = C.c_uint64
        ne = self.L.emu_edges(self.h, None, None)
        k, c, ep, eg = np.zeros(ne, "<u8"), np.zeros(ne, "<u8"), np.zeros(ne), np.zeros(ne)
        self.L.emu_edges(self.h, v(k), v(c))
        self.L.emu_bias_edges(self.h, v(ep), v(eg))
        edges = {int(k[i]): (ep[i], eg[i]) for i in range(ne)}
        self.L.emu_keyed.restype = C.c_uint64
        n = self.L.emu_keyed(self.h, None)
        w = np.zeros(max(1, n), "<u4")
        self.L.emu_keyed(self.h, v(w))
        keys, i = [], 0
        while i < n:
            nw = int(w[i])
            keys.append(tuple(int(x) for x in w[i + 1:i + 1 + nw]))
            i += nw + 3
        kp, kg = np.zeros(max(1, len(keys))), np.zeros(max(1, len(keys)))
        self.L.emu_bias_keyed(self.h, v(kp), v(kg))
        longs = {key: (kp[j], kg[j]) for j, key in enumerate(keys) if not (key[0] >> 28) & 1}
        multi = {key: (kp[j], kg[j]) for j, key in enumerate(keys) if (key[0] >> 28) & 1}
        fore, back = np.zeros(4096), np.zeros(4096)
        self.L.emu_bias_model(self.h, v(fore), v(back))
        return (np_, ng), edges, longs, multi, fore, back, short

    def edges(self):
        self.L.emu_edges.restype = C.c_uint64
        n = self.L.emu_edges(self.h, None, None)
        k, v = np.zeros(n, "<u8"), np.zeros(n, "<u8")
        self.L.emu_edges(self.h, k.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
        o = np.argsort(k)
        return k[o], v[o]

    def keyed(self):
        self.L.emu_keyed.restype = C.c_uint64
        n = self.L.emu_keyed(self.h, None)
        w = np.zeros(max(1, n), "<u4")
        self.L.emu_keyed(self.h, w.ctypes.data_as(C.c_void_p))
        longs, multi, i = {}, {}, 0
        while i < n:
            nw = int(w[i])
            key = tuple(int(x) for x in w[i + 1:i + 1 + nw])
            cnt = int(w[i + 1 + nw]) | (int(w[i + 2 + nw]) << 32)
            (multi if (key[0] >> 28) & 1 else longs)[key] = cnt
            i += nw + 3
        return longs, multi
