"""ctypes wrapper of the CPU oracle (oracle/whippet_oracle.cpp) — TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import whippet_jl_b200 as wq
from whippet_jl_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwhippet_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "whippet_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "whippet_b200.h")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libwhippet_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.wqo_create.restype = C.c_void_p
        _lib.wqo_create.argtypes = [C.POINTER(abi.IndexDesc)]
        _lib.wqo_destroy.argtypes = [C.c_void_p]
        _lib.wqo_sa_value.restype = C.c_int64
        _lib.wqo_sa_value.argtypes = [C.c_void_p, C.c_int64]
        _lib.wqo_result_n_aln.restype = C.c_uint64
        _lib.wqo_result_n_nodes.restype = C.c_uint64
        _lib.wqo_result_n_aln.argtypes = [C.c_void_p]
        _lib.wqo_result_n_nodes.argtypes = [C.c_void_p]
    return _lib


class AlignResult:
    """Per-read alignments: CSR over `aln`, each alignment owning nodes[path_start:path_start+npath]."""

    def __init__(self, hit_start, nhits, flipped, aln, nodes, aln_mate=None):
        self.hit_start, self.nhits, self.flipped = hit_start, nhits, flipped
        self.aln, self.nodes, self.aln_mate = aln, nodes, aln_mate

    def read(self, i, mate=False):
        """List of (offset, offsetread, strand, isvalid, [(gene,node,matches,mismatches,mistol_bits)...])."""
        out = []
        src = self.aln_mate if mate else self.aln
        for k in range(int(self.hit_start[i]), int(self.hit_start[i]) + int(self.nhits[i])):
            a = src[k]
            ns = self.nodes[int(a["path_start"]): int(a["path_start"]) + int(a["npath"])]
            out.append((int(a["offset"]), int(a["offsetread"]), int(a["strand"]), int(a["isvalid"]),
                        [(int(n["gene"]), int(n["node"]), int(n["matches"]), int(n["mismatches"]),
                          int(np.float32(n["mistolerance"]).view(np.uint32))) for n in ns]))
        return out


class Oracle:
    def __init__(self, index: wq.FlatIndex):
        self.index = index
        self._desc = index.desc()
        self.h = lib().wqo_create(C.byref(self._desc))

    def __del__(self):
        try:
            lib().wqo_destroy(C.c_void_p(self.h))
        except Exception:
            pass

    def sa_range(self, codes):
        q = np.ascontiguousarray(codes, "u1")
        sp, ep = C.c_int64(), C.c_int64()
        lib().wqo_sa_range(C.c_void_p(self.h), abi.ptr(q, abi.u8p), C.c_int64(len(q)), C.byref(sp), C.byref(ep))
        return sp.value, ep.value

    def sa_value(self, row):
        return lib().wqo_sa_value(C.c_void_p(self.h), row)

    def seed_locate(self, param: wq.AlignParam, reads: wq.ReadBatch):
        out = np.zeros((reads.n, 4), np.int64)
        p, b = param.c(), reads.c()
        lib().wqo_seed_locate(C.c_void_p(self.h), C.byref(p), C.byref(b), out.ctypes.data_as(C.c_void_p))
        return out

    def _collect(self, n, paired):
        L = lib()
        na, nn = L.wqo_result_n_aln(C.c_void_p(self.h)), L.wqo_result_n_nodes(C.c_void_p(self.h))
        hs, nh, fl = np.zeros(n, "<u4"), np.zeros(n, "u1"), np.zeros(n, "u1")
        aln = np.zeros(na, abi.ALIGNMENT_DT)
        alm = np.zeros(na, abi.ALIGNMENT_DT) if paired else None
        nodes = np.zeros(nn, abi.ALIGN_NODE_DT)
        L.wqo_result_copy(C.c_void_p(self.h), hs.ctypes.data_as(C.c_void_p), nh.ctypes.data_as(C.c_void_p),
                          fl.ctypes.data_as(C.c_void_p), aln.ctypes.data_as(C.c_void_p),
                          alm.ctypes.data_as(C.c_void_p) if paired else None, nodes.ctypes.data_as(C.c_void_p))
        return AlignResult(hs, nh, fl, aln, nodes, alm)

    def align_se(self, param: wq.AlignParam, reads: wq.ReadBatch) -> AlignResult:
        p, b = param.c(), reads.c()
        lib().wqo_align_se(C.c_void_p(self.h), C.byref(p), C.byref(b))
        return self._collect(reads.n, False)

    def align_pe(self, param: wq.AlignParam, fwd: wq.ReadBatch, rev: wq.ReadBatch) -> AlignResult:
        p, f, r = param.c(), fwd.c(), rev.c()
        lib().wqo_align_pe(C.c_void_p(self.h), C.byref(p), C.byref(f), C.byref(r))
        return self._collect(fwd.n, True)

    # ---- counting (process_reads!, src/reads.jl:41-80) -------------------------------------------
    def quant_reset(self):
        lib().wqo_quant_reset(C.c_void_p(self.h))

    def process_reads(self, param: wq.AlignParam, reads: wq.ReadBatch):
        p, b = param.c(), reads.c()
        rc = lib().wqo_process_reads_se(C.c_void_p(self.h), C.byref(p), C.byref(b))
        if rc == -2:
            raise ValueError("bias correction: a mapped read is shorter than 37 bases (the reference throws BoundsError, src/bias.jl:163-165)")

    def process_paired_reads(self, param: wq.AlignParam, fwd: wq.ReadBatch, rev: wq.ReadBatch):
        p, f, r = param.c(), fwd.c(), rev.c()
        rc = lib().wqo_process_reads_pe(C.c_void_p(self.h), C.byref(p), C.byref(f), C.byref(r))
        if rc == -2:
            raise ValueError("bias correction: a mapped read is shorter than 37 bases (the reference throws BoundsError, src/bias.jl:163-165)")

    # ---- --biascorrect (JointBiasMod / JointBiasCounter, src/bias.jl:103-270; bin/whippet-quant.jl:116-122, 156-169) ----
    def set_bias(self, on=True):
        """CounterType = JointBiasCounter, BiasModelType = JointBiasMod; takes effect at the next quant_reset()."""
        lib().wqo_set_bias(C.c_void_p(self.h), int(bool(on)))

    def bias_primer_adjust(self):
        """primer_normalize!(mod); primer_adjust!(quant, mod); primer_adjust!(multi, mod) -- before build_equivalence_classes!."""
        lib().wqo_bias_primer_adjust(C.c_void_p(self.h))

    def bias_gc_adjust(self):
        """gc_adjust!(quant, mod); gc_adjust!(multi, mod) -- after gene_em! / set_gene_tpm!, before assign_ambig!."""
        lib().wqo_bias_gc_adjust(C.c_void_p(self.h))

    def bias_model(self):
        fore, back, short = np.zeros(4096), np.zeros(4096), C.c_uint64()
        lib().wqo_bias_model(C.c_void_p(self.h), fore.ctypes.data_as(C.c_void_p), back.ctypes.data_as(C.c_void_p), C.byref(short))
        return fore, back, short.value

    def merge(self, other: "Oracle"):
        lib().wqo_merge(C.c_void_p(self.h), C.c_void_p(other.h))

    def process_events_count(self, readlen):
        L = lib()
        L.wqo_process_events.restype = C.c_uint64
        return L.wqo_process_events(C.c_void_p(self.h), C.c_int64(readlen))

    def stats(self):
        out = (C.c_uint64 * 5)()
        m = C.c_double()
        lib().wqo_stats(C.c_void_p(self.h), out, C.byref(m))
        return dict(total=out[0], mapped=out[1], multi=out[2], sum_readlen=out[3], mean_readlen=m.value)

    def node_counts(self):
        a = np.zeros(self.index.n_nodes, "<f8")
        lib().wqo_node_counts(C.c_void_p(self.h), a.ctypes.data_as(C.c_void_p))
        return a

    def edges(self):
        L = lib()
        L.wqo_edge_count.restype = C.c_uint64
        n = L.wqo_edge_count(C.c_void_p(self.h))
        k, v = np.zeros(n, "<u8"), np.zeros(n, "<f8")
        L.wqo_edges(C.c_void_p(self.h), k.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
        return k, v

    def keyed(self, kind):
        L = lib()
        nr, nw = C.c_uint64(), C.c_uint64()
        L.wqo_keyed(C.c_void_p(self.h), kind, C.byref(nr), C.byref(nw), None, None, None)
        ptr_, words, vals = np.zeros(nr.value + 1, "<u8"), np.zeros(max(1, nw.value), "<u4"), np.zeros(max(1, nr.value), "<f8")
        v = lambda a: a.ctypes.data_as(C.c_void_p)
        L.wqo_keyed(C.c_void_p(self.h), kind, C.byref(nr), C.byref(nw), v(ptr_), v(words), v(vals))
        return {tuple(int(x) for x in words[int(ptr_[i]):int(ptr_[i + 1])]): float(vals[i]) for i in range(nr.value)}

    # ---- quantification (bin/whippet-quant.jl:161-177) --------------------------------------------
    def add_node(self, gene, node, v):
        lib().wqo_add_node(C.c_void_p(self.h), C.c_uint32(gene), C.c_uint32(node), C.c_double(v))

    def add_edge(self, gene, l, r, v):
        lib().wqo_add_edge(C.c_void_p(self.h), C.c_uint32(gene), C.c_uint32(l), C.c_uint32(r), C.c_double(v))

    def add_keyed(self, kind, words, v):
        w = np.ascontiguousarray(words, "<u4")
        lib().wqo_add_keyed(C.c_void_p(self.h), kind, w.ctypes.data_as(C.c_void_p), C.c_uint64(len(w)), C.c_double(v))

    def em_tpm(self, readlen, sig=1, maxit=10000):
        """build_equivalence_classes! + calculate_tpm! + gene_em! + set_gene_tpm!; returns
        (iterations, tpm[I], count[I], genetpm[G])."""
        L = lib()
        L.wqo_em_tpm.restype = C.c_int64
        it = L.wqo_em_tpm(C.c_void_p(self.h), C.c_int64(readlen), C.c_int64(sig), C.c_int64(maxit))
        tpm, cnt = np.zeros(self.index.n_isoforms), np.zeros(self.index.n_isoforms)
        gt = np.zeros(self.index.n_genes)
        v = lambda a: a.ctypes.data_as(C.c_void_p)
        L.wqo_em_results(C.c_void_p(self.h), v(tpm), v(cnt), v(gt))
        return int(it), tpm, cnt, gt

    def classes(self, multi=False):
        """[(ids, props, count, isdone, postassign)] in canonical order."""
        L = lib()
        L.wqo_n_classes.restype = C.c_uint64
        L.wqo_class.restype = C.c_uint64
        out = []
        ids, props = np.zeros(4096, "<i4"), np.zeros(4096, "<f8")
        for i in range(L.wqo_n_classes(C.c_void_p(self.h), int(multi))):
            cnt, fl = C.c_double(), C.c_int32()
            n = L.wqo_class(C.c_void_p(self.h), int(multi), C.c_uint64(i), ids.ctypes.data_as(C.c_void_p),
                            props.ctypes.data_as(C.c_void_p), C.byref(cnt), C.byref(fl))
            out.append((ids[:n].tolist(), props[:n].tolist(), cnt.value, bool(fl.value & 1), bool(fl.value & 2)))
        return out

    def assign_ambig(self):
        lib().wqo_assign_ambig(C.c_void_p(self.h))

    # ---- events (src/events.jl) -------------------------------------------------------------------
    def events_flat(self, readlen):
        """process_events over all genes, every record in the canonical flat form (see canonical_events):
        (hdr[E,8] u32, dbl[E,3] f64, values f64, path_len u32, path_nodes u32)."""
        L = lib()
        L.wqo_process_events.restype = C.c_uint64
        E = int(L.wqo_process_events(C.c_void_p(self.h), C.c_int64(readlen)))
        n = (C.c_uint64 * 3)()
        L.wqo_events_flat(C.c_void_p(self.h), n, None, None, None, None, None)
        hdr, dbl = np.zeros((E, 8), "<u4"), np.zeros((E, 3), "<f8")
        values, plen, nodes = np.zeros(int(n[0]), "<f8"), np.zeros(int(n[1]), "<u4"), np.zeros(int(n[2]), "<u4")
        v = lambda a: a.ctypes.data_as(C.c_void_p)
        L.wqo_events_flat(C.c_void_p(self.h), n, v(hdr), v(dbl), v(values), v(plen), v(nodes))
        return hdr, dbl, values, plen, nodes

    def process_events(self, readlen):
        """process_events over all genes.  Returns a list of dicts (one per event record, gene/node order)."""
        L = lib()
        L.wqo_process_events.restype = C.c_uint64
        L.wqo_event_path.restype = C.c_uint64
        n = L.wqo_process_events(C.c_void_p(self.h), C.c_int64(readlen))
        out = []
        hdr = (C.c_uint32 * 8)()
        vals = (C.c_double * 3)()
        buf = np.zeros(70000, "<u4")
        for i in range(n):
            L.wqo_event(C.c_void_p(self.h), C.c_uint64(i), hdr, vals)
            e = dict(gene=hdr[0], node=hdr[1], motif=hdr[2], kind=hdr[3], flags=hdr[4], psi=vals[0], total=vals[1],
                     iterations=int(vals[2]), utr_psi=[], inc=[], exc=[])
            for which, key, cnt in ((0, "utr_psi", hdr[5]), (1, "inc", hdr[6]), (2, "exc", hdr[7])):
                if not cnt:
                    continue
                v = np.zeros(cnt)
                if which == 0 or e["kind"] == 1 and hdr[6] and hdr[7]:
                    L.wqo_event_psi(C.c_void_p(self.h), C.c_uint64(i), which, v.ctypes.data_as(C.c_void_p))
                if which == 0:
                    e[key] = v.tolist()
                else:
                    paths = []
                    for k in range(cnt):
                        m = L.wqo_event_path(C.c_void_p(self.h), C.c_uint64(i), which, C.c_uint64(k), buf.ctypes.data_as(C.c_void_p))
                        paths.append((tuple(int(x) for x in buf[:m]), float(v[k])))
                    e[key] = paths
            out.append(e)
        return out


def canonical_events(ev, vals, pptr, pnodes):
    """The library's wq_events_out arrays (abi.EVENT_DT records, values, path_ptr, path_nodes) in the canonical flat form of
    Oracle.events_flat, so that two numpy comparisons check every record: (hdr[E,8], dbl[E,3], values, path_len, path_nodes)."""
    E = len(ev)
    hdr = np.zeros((E, 8), "<u4")
    for k, f in enumerate(("gene", "node", "motif", "kind", "flags", "n_utr", "n_inc", "n_exc")):
        hdr[:, k] = ev[f]
    dbl = np.stack([ev["psi"], ev["total"], ev["iterations"].astype("<f8")], axis=1) if E else np.zeros((0, 3))
    kind, n_utr, n_inc, n_exc = hdr[:, 3], hdr[:, 5].astype(np.int64), hdr[:, 6].astype(np.int64), hdr[:, 7].astype(np.int64)
    nv = np.where(kind == 2, n_utr, np.where((kind == 1) & (n_inc > 0) & (n_exc > 0), n_inc + n_exc, 0))

    def expand(start, cnt):          # concatenated ranges start[i] .. start[i] + cnt[i]
        tot = int(cnt.sum())
        if tot == 0:
            return np.zeros(0, np.int64)
        first = np.cumsum(cnt) - cnt
        return np.arange(tot, dtype=np.int64) - np.repeat(first, cnt) + np.repeat(start, cnt)
    values = np.asarray(vals, "<f8")[expand(ev["value_start"].astype(np.int64), nv)] if E else np.zeros(0)
    npaths = np.where(kind == 2, 0, n_inc + n_exc)
    pidx = expand(ev["path_start"].astype(np.int64), npaths) if E else np.zeros(0, np.int64)
    pptr = np.asarray(pptr, np.int64)
    plen = (pptr[pidx + 1] - pptr[pidx]).astype("<u4") if len(pidx) else np.zeros(0, "<u4")
    nodes = np.asarray(pnodes, "<u4")[expand(pptr[pidx], plen.astype(np.int64))] if len(pidx) else np.zeros(0, "<u4")
    return hdr, dbl, values, plen, nodes


def gene_em_direct(lengths, count, multi, classes, readlen, sig=1, maxit=10000, trace_n=0):
    """The oracle's calculate_tpm! + gene_em! on a problem given directly: multi / classes = [(ids 1-based, count)], visited in
    that order (multi-map classes first).  Returns (iterations, tpm, count, trace[trace_n][I])."""
    L = lib()
    L.wqo_gene_em_direct.restype = C.c_int64
    I = len(lengths)
    ln, cnt = np.ascontiguousarray(lengths, "<i8"), np.ascontiguousarray(count, "<f8")
    allc = list(multi) + list(classes)
    ptr_ = np.zeros(len(allc) + 1, "<u4")
    ptr_[1:] = np.cumsum([len(c[0]) for c in allc])
    ids = np.ascontiguousarray([i for c in allc for i in c[0]] or [0], "<i4")
    cc = np.ascontiguousarray([c[1] for c in allc] or [0.0], "<f8")
    tpm, cout, trace = np.zeros(I), np.zeros(I), np.zeros((max(1, trace_n), I))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    it = L.wqo_gene_em_direct(C.c_uint32(I), p(ln), p(cnt), C.c_uint32(len(multi)), C.c_uint32(len(allc)), p(ptr_), p(ids), p(cc),
                              C.c_int64(readlen), C.c_int64(sig), C.c_int64(maxit), p(tpm), p(cout), p(trace), C.c_uint32(trace_n))
    return int(it), tpm, cout, trace[:trace_n]


def em_psi(n_inc, count, length, amb_sets, amb_mult, tandem, readlen, sig=4, maxit=1500):
    """spliced_em! / rec_tandem_em! (src/events.jl:853-932) on one problem given directly: the first n_inc paths are
    inclusion paths; amb_sets = list of lists of 0-based path indices.  Returns (psi[np], iterations)."""
    L = lib()
    L.wqo_em_psi.restype = C.c_int64
    cnt, ln = np.ascontiguousarray(count, "<f8"), np.ascontiguousarray(length, "<f8")
    ptr_ = np.zeros(len(amb_sets) + 1, "<u4")
    ptr_[1:] = np.cumsum([len(a) for a in amb_sets])
    paths = np.ascontiguousarray([x for a in amb_sets for x in a] or [0], "<u4")
    mult = np.ascontiguousarray(list(amb_mult) or [0.0], "<f8")
    psi = np.zeros(len(cnt))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    it = L.wqo_em_psi(C.c_uint32(len(cnt)), C.c_uint32(n_inc), p(cnt), p(ln), C.c_uint32(len(amb_sets)), p(ptr_), p(paths), p(mult),
                      C.c_int(1 if tandem else 0), C.c_int64(readlen), C.c_int(sig), C.c_int64(maxit), p(psi))
    return psi, int(it)


def path_in_path(first, last):
    """Base.in(first::SGAlignPath, last::SGAlignPath) (src/quant.jl:47-62); paths = [(gene, node), ...]."""
    a = np.ascontiguousarray([x for p in first for x in p] or [0], "<u4")
    b = np.ascontiguousarray([x for p in last for x in p] or [0], "<u4")
    v = lambda x: x.ctypes.data_as(C.c_void_p)
    return bool(lib().wqo_path_in_path(v(a), C.c_uint64(len(first)), v(b), C.c_uint64(len(last))))


def reduce_graph(edges):
    """reduce_graph (src/events.jl:223-266) on [(first, last, value)]; returns the list of node sets."""
    L = lib()
    L.wqo_reduce_graph.restype = C.c_uint64
    f = np.array([e[0] for e in edges], "<u4")
    l = np.array([e[1] for e in edges], "<u4")
    v = np.array([e[2] for e in edges], "<f8")
    out = np.zeros(1 << 16, "<u4")
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = L.wqo_reduce_graph(p(f), p(l), p(v), C.c_uint64(len(edges)), p(out), C.c_uint64(len(out)))
    sets, w = [], 0
    for _ in range(n):
        k = int(out[w])
        sets.append(frozenset(int(x) for x in out[w + 1:w + 1 + k]))
        w += 1 + k
    return sets


def counters(reset=False):
    """Instrumentation totals since the last reset: fm_steps, hits, nodes_touched, bases, out_nodes,
    out_alns, count_updates (process-wide; only meaningful for single-threaded runs)."""
    out = (C.c_uint64 * 7)()
    hist = (C.c_uint64 * 34)()
    lib().wqo_steps_hist(hist)            # before the reset below
    lib().wqo_counters(out, 1 if reset else 0)
    d = dict(zip(("fm_steps", "hits", "nodes_touched", "bases", "out_nodes", "out_alns", "count_updates"), map(int, out)))
    d["steps_hist"] = [int(x) for x in hist]     # steps_hist[s] = backward searches that executed s steps
    return d
