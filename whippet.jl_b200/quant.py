"""Host-side mirror of the reference's quantification interface for the accelerated path.

Names follow the calls bin/whippet-quant.jl:124-181 makes (GraphLibQuant, process_reads!,
ungapped_align, build_equivalence_classes!/calculate_tpm!/gene_em!, assign_ambig!); each method
cites the reference function it stands for.  All compute happens behind the C ABI."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import abi, lib
from .flat import AlignParam, FlatIndex, ReadBatch


class AlignResult:
    """Batch form of `Nullable{Vector{SGAlignment}}` (src/align.jl:86): read i owns
    aln[hit_start[i] : hit_start[i]+nhits[i]], each alignment owns nodes[path_start : +npath]."""

    def __init__(self, hit_start, nhits, flipped, aln, nodes, aln_mate=None):
        self.hit_start, self.nhits, self.flipped = hit_start, nhits, flipped
        self.aln, self.nodes, self.aln_mate = aln, nodes, aln_mate

    def read(self, i, mate=False):
        out = []
        src = self.aln_mate if mate else self.aln
        for k in range(int(self.hit_start[i]), int(self.hit_start[i]) + int(self.nhits[i])):
            a = src[k]
            ns = self.nodes[int(a["path_start"]): int(a["path_start"]) + int(a["npath"])]
            out.append((int(a["offset"]), int(a["offsetread"]), int(a["strand"]), int(a["isvalid"]),
                        [(int(n["gene"]), int(n["node"]), int(n["matches"]), int(n["mismatches"]),
                          int(np.float32(n["mistolerance"]).view(np.uint32))) for n in ns]))
        return out


@dataclass
class MapStats:
    total: int
    mapped: int
    multi: int
    sum_readlen: int
    path_overflow: int
    mean_readlen: float = 0.0      # the reference's running mean in read order (src/reads.jl:69), floored by the caller


def write_index(index: FlatIndex, path: str, with_info: bool = True):
    """wq_index_write: the re-laid index of `index` as one file (host only, no GPU needed).  lib.info (sequence name and
    strand of every gene, SAM output) is stored when the FlatIndex carries it."""
    L = lib.load()
    d = index.desc()
    names = getattr(index, "gene_seqname", None)
    strands = getattr(index, "gene_strand", None)
    if with_info and names is not None and strands is not None and len(names) == index.n_genes and len(strands) == index.n_genes:
        arr = (C.c_char_p * index.n_genes)(*[str(x).encode() for x in names])
        st = np.ascontiguousarray([1 if x else 0 for x in strands], "u1")
        lib.check(L.wq_index_write(C.byref(d), arr, abi.ptr(st, abi.u8p), os.fsencode(path)))
    else:
        lib.check(L.wq_index_write(C.byref(d), None, None, os.fsencode(path)))


class GraphLibQuant:
    """`GraphLibQuant{C,DefaultCounter}(lib)` + `MultiMapping{C,DefaultCounter}()`
    (src/quant.jl:165-198, 439-446) living on one GPU."""

    def __init__(self, index: FlatIndex, device: int = 0, index_file: str | None = None):
        """From the flat arrays of a GraphLib (wq_index_create), or, with `index_file`, from the library's native on-disk
        index (wq_index_load; `index` then only serves the host-side helpers that look at the annotation)."""
        self.L = lib.load()
        self.index = index
        h = C.c_void_p()
        if index_file is not None:
            lib.check(self.L.wq_index_load(os.fsencode(index_file), int(device), C.byref(h)))
        else:
            self._desc = index.desc()
            lib.check(self.L.wq_index_create(C.byref(self._desc), int(device), C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.wq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        lib.check(self.L.wq_quant_reset(self.h))

    # ---- --biascorrect (JointBiasMod / JointBiasCounter, src/bias.jl:103-270; bin/whippet-quant.jl:116-122, 156-169) ----
    def set_biascorrect(self, on: bool = True):
        """CounterType = JointBiasCounter, BiasModelType = JointBiasMod: call before the first batch (it resets the count stores).  From then
        on gene_em / assign_ambig / process_events work with the primer- and GC-adjusted Float64 counts."""
        lib.check(self.L.wq_set_biascorrect(self.h, int(bool(on))))

    def bias_download(self):
        """Every counter's count after primer_adjust! and gc_adjust! factor (< 0: none), aligned with counts() and keyed(kind), plus the
        normalised model tables: dict(node=(p, g), edge=(p, g), circ=(p, g), long=(p, g), multi=(p, g), fore, back)."""
        o = abi.BiasCountsC()
        lib.check(self.L.wq_bias_download(self.h, C.byref(o)))
        dp = C.POINTER(C.c_double)
        arrs = {k: (np.zeros(max(1, n)), np.zeros(max(1, n))) for k, n in
                (("node", o.n_nodes), ("edge", o.n_edge), ("circ", o.n_circ), ("long", o.n_long), ("multi", o.n_multi))}
        sizes = dict(node=o.n_nodes, edge=o.n_edge, circ=o.n_circ, long=o.n_long, multi=o.n_multi)
        for k, (p, g) in arrs.items():
            setattr(o, k + "_primer", p.ctypes.data_as(dp)); setattr(o, k + "_gc", g.ctypes.data_as(dp))
        fore, back = np.zeros(4096), np.zeros(4096)
        o.fore, o.back = fore.ctypes.data_as(dp), back.ctypes.data_as(dp)
        lib.check(self.L.wq_bias_download(self.h, C.byref(o)))
        out = {k: (p[:sizes[k]], g[:sizes[k]]) for k, (p, g) in arrs.items()}
        out["fore"], out["back"] = fore, back
        return out

    # ---- alignment ------------------------------------------------------------------------------
    def ungapped_align(self, param: AlignParam, reads: ReadBatch, mate: ReadBatch | None = None) -> AlignResult:
        """ungapped_align for every read (src/align.jl:225-270; paired: src/paired.jl:42-125).
        Counts are accumulated as well, exactly as process_reads! does."""
        n = reads.n
        T = max(1, param.seed_tolerate)
        cap = n * T + 1
        hs, nh, fl = np.zeros(n, "<u4"), np.zeros(n, "u1"), np.zeros(n, "u1")
        aln = np.zeros(cap, abi.ALIGNMENT_DT)
        alm = np.zeros(cap, abi.ALIGNMENT_DT) if mate is not None else None
        nodes = np.zeros(cap * (2 if mate is not None else 1) * abi.WQ_MAX_PATH // 4 + 64, abi.ALIGN_NODE_DT)
        out = abi.AlignOutC()
        out.hit_start = abi.ptr(hs, abi.u32p)
        out.nhits = abi.ptr(nh, abi.u8p)
        out.flipped = abi.ptr(fl, abi.u8p)
        out.aln = aln.ctypes.data
        out.aln_mate = alm.ctypes.data if alm is not None else None
        out.nodes = nodes.ctypes.data
        out.aln_cap, out.nodes_cap = len(aln), len(nodes)
        p, b = param.c(), reads.c()
        if mate is None:
            lib.check(self.L.wq_align_se(self.h, C.byref(p), C.byref(b), C.byref(out)))
        else:
            m = mate.c()
            lib.check(self.L.wq_align_pe(self.h, C.byref(p), C.byref(b), C.byref(m), C.byref(out)))
        na, nn = int(out.aln_used), int(out.nodes_used)
        return AlignResult(hs, nh, fl, aln[:na], nodes[:nn], alm[:na] if alm is not None else None)

    def process_reads(self, param: AlignParam, reads: ReadBatch) -> MapStats:
        """process_reads! (src/reads.jl:41-80) over one in-memory batch: align, then count!/push!(multi)
        on the device.  Returns the running (mapped, total, ...) statistics."""
        p, b = param.c(), reads.c()
        lib.check(self.L.wq_align_se(self.h, C.byref(p), C.byref(b), None))
        return self.stats()

    def process_paired_reads(self, param: AlignParam, fwd: ReadBatch, rev: ReadBatch) -> MapStats:
        """process_paired_reads! (src/reads.jl:83-129)."""
        p, f, r = param.c(), fwd.c(), rev.c()
        lib.check(self.L.wq_align_pe(self.h, C.byref(p), C.byref(f), C.byref(r), None))
        return self.stats()

    def process_fastq(self, param: AlignParam, fwd_path: str, rev_path: str | None = None, qualoffset: int = 33,
                      batch_reads: int = 0, sam_fd: int = -1) -> MapStats:
        """process_reads! / process_paired_reads! over FASTQ files (gzip or plain): the library parses and packs the
        next batch on its host threads while the device aligns and counts the current one.  sam_fd >= 0 is `sam=true`
        (src/reads.jl:49-52, 63-67): header + records of every batch go to that file descriptor in read order.
        batch_reads = 0 takes the library's default (512 k records: parsing batch k+1 hides behind aligning batch k)."""
        p = param.c()
        s = abi.MapStatsC()
        lib.check(self.L.wq_process_fastq_sam(self.h, C.byref(p), fwd_path.encode(), rev_path.encode() if rev_path else None,
                                              int(qualoffset), int(batch_reads), int(sam_fd), C.byref(s)))
        return MapStats(s.total, s.mapped, s.multi, s.sum_readlen, s.path_overflow, s.mean_readlen)

    # ---- SAM output (src/io.jl:69-276) -------------------------------------------------------------
    def set_gene_info(self, seqnames, strands):
        """lib.info (src/types.jl:33-37): reference sequence name and strand ('+' = True) of every gene."""
        G = self.index.n_genes
        assert len(seqnames) == G and len(strands) == G
        arr = (C.c_char_p * G)(*[str(x).encode() for x in seqnames])
        st = np.ascontiguousarray([1 if x else 0 for x in strands], "u1")
        lib.check(self.L.wq_index_set_gene_info(self.h, arr, abi.ptr(st, abi.u8p)))

    def sam_header(self) -> str:
        used = C.c_uint64()
        lib.check(self.L.wq_sam_header(self.h, None, 0, C.byref(used)))
        buf = C.create_string_buffer(int(used.value) + 1)
        lib.check(self.L.wq_sam_header(self.h, buf, used.value, C.byref(used)))
        return buf.raw[:used.value].decode()

    def sam_batch(self, reads: ReadBatch, names, res: AlignResult, mate: ReadBatch | None = None, mate_names=None,
                  qualoffset: int = 33) -> str:
        """write_sam for every read of an aligned batch in read order (process_reads! with sam=true)."""
        def pack(nm):
            enc = [x.encode() for x in nm]
            ln = np.array([len(x) for x in enc], "<u4")
            off = np.zeros(len(enc), "<u8")
            if len(enc) > 1:
                off[1:] = np.cumsum(ln[:-1])
            return off, ln, b"".join(enc)
        out = abi.AlignOutC()
        out.hit_start = abi.ptr(res.hit_start, abi.u32p)
        out.nhits = abi.ptr(res.nhits, abi.u8p)
        out.flipped = abi.ptr(res.flipped, abi.u8p)
        out.aln = res.aln.ctypes.data
        out.aln_mate = res.aln_mate.ctypes.data if res.aln_mate is not None else None
        out.nodes = res.nodes.ctypes.data
        out.aln_cap, out.nodes_cap = len(res.aln), len(res.nodes)
        out.aln_used, out.nodes_used = len(res.aln), len(res.nodes)
        b = reads.c()
        o1, l1, t1 = pack(names)
        if mate is not None:
            m = mate.c()
            o2, l2, t2 = pack(mate_names)
            args = (C.byref(b), C.byref(m), abi.ptr(o1, abi.u64p), abi.ptr(l1, abi.u32p), t1,
                    abi.ptr(o2, abi.u64p), abi.ptr(l2, abi.u32p), t2)
        else:
            args = (C.byref(b), None, abi.ptr(o1, abi.u64p), abi.ptr(l1, abi.u32p), t1, None, None, None)
        used = C.c_uint64()
        lib.check(self.L.wq_sam_batch(self.h, *args, C.byref(out), int(qualoffset), None, 0, C.byref(used)))
        buf = C.create_string_buffer(int(used.value) + 1)
        lib.check(self.L.wq_sam_batch(self.h, *args, C.byref(out), int(qualoffset), buf, used.value, C.byref(used)))
        return buf.raw[:used.value].decode("latin-1")

    def stats(self) -> MapStats:
        s = abi.MapStatsC()
        lib.check(self.L.wq_map_stats_get(self.h, C.byref(s)))
        return MapStats(s.total, s.mapped, s.multi, s.sum_readlen, s.path_overflow, s.mean_readlen)

    def last_deferred(self) -> int:
        """Extension tasks of the last alignment call that went to the second-pass kernel."""
        return int(self.L.wq_last_deferred(self.h))

    def last_kernel_ms(self):
        ms = (C.c_float * 3)()
        lib.check(self.L.wq_last_kernel_ms(self.h, ms))
        return [float(x) for x in ms]

    def quant_kernel_stats(self):
        """Device times (CUDA events) of the transcript EM and PSI EM kernels of the last gene_em / process_events and
        the sizes of their algorithmic-byte models (wq_quant_kernel_stats)."""
        out = (C.c_double * 8)()
        lib.check(self.L.wq_quant_kernel_stats(self.h, out))
        keys = ("em_tpm_ms", "em_psi_ms", "em_iterations", "classes", "class_members", "isoforms", "psi_problems", "psi_bytes")
        return dict(zip(keys, (float(x) for x in out)))

    # ---- count stores ---------------------------------------------------------------------------
    def counts(self):
        """SpliceGraphQuant.node / .edge / .circ (src/quant.jl:129-135) as
        (node_count[N], edge_key, edge_count, circ_key, circ_count); keys = gene<<32 | l<<16 | r, sorted."""
        c = abi.CountsC()
        lib.check(self.L.wq_counts_download(self.h, C.byref(c)))
        node = np.zeros(c.n_nodes, "<u8")
        ek, ec = np.zeros(c.n_edge, "<u8"), np.zeros(c.n_edge, "<u8")
        ck, cc = np.zeros(c.n_circ, "<u8"), np.zeros(c.n_circ, "<u8")
        c.node_count = abi.ptr(node, abi.u64p)
        c.edge_key, c.edge_count = abi.ptr(ek, abi.u64p), abi.ptr(ec, abi.u64p)
        c.circ_key, c.circ_count = abi.ptr(ck, abi.u64p), abi.ptr(cc, abi.u64p)
        lib.check(self.L.wq_counts_download(self.h, C.byref(c)))
        return node, ek, ec, ck, cc

    def keyed(self, kind):
        """kind 0: SpliceGraphQuant.long; kind 1: MultiMapping.map.  Returns {tuple(words): count}."""
        k = abi.KeyedC()
        lib.check(self.L.wq_keyed_download(self.h, kind, C.byref(k)))
        ptr_ = np.zeros(k.n_rec + 1, "<u8")
        words = np.zeros(max(1, k.n_words), "<u4")
        cnt = np.zeros(max(1, k.n_rec), "<u8")
        k.rec_ptr, k.words, k.count = abi.ptr(ptr_, abi.u64p), abi.ptr(words, abi.u32p), abi.ptr(cnt, abi.u64p)
        lib.check(self.L.wq_keyed_download(self.h, kind, C.byref(k)))
        return {tuple(int(x) for x in words[int(ptr_[i]):int(ptr_[i + 1])]): int(cnt[i]) for i in range(k.n_rec)}

    # ---- quantification -------------------------------------------------------------------------
    def gene_em(self, readlen: int, sig: int = 1, maxit: int = 10000):
        """build_equivalence_classes!(quant, lib) + calculate_tpm!(quant, readlen) + gene_em!(quant, multi;
        sig, readlen, maxit) + set_gene_tpm! (bin/whippet-quant.jl:161-166).  Returns
        (iterations, tpm[I], count[I], genetpm[G])."""
        I, G = self.index.n_isoforms, self.index.n_genes
        tpm, cnt, gt = np.zeros(I), np.zeros(I), np.zeros(G)
        it = C.c_int64()
        p = abi.EmParamC(int(readlen), int(sig), int(maxit))
        lib.check(self.L.wq_em_tpm(self.h, C.byref(p), abi.ptr(tpm, abi.f64p), abi.ptr(cnt, abi.f64p),
                                   abi.ptr(gt, abi.f64p), C.byref(it)))
        return int(it.value), tpm, cnt, gt

    def assign_ambig(self):
        """assign_ambig!(quant, lib, multi) (src/quant.jl:520-553).  Returns the node / edge / circ stores
        as Float64: (node_value[N], edge_key, edge_value, circ_key, circ_value)."""
        v = abi.ValuesC()
        lib.check(self.L.wq_assign_ambig(self.h, C.byref(v)))
        node = np.zeros(v.n_nodes)
        ek, ev = np.zeros(v.n_edge, "<u8"), np.zeros(v.n_edge)
        ck, cv = np.zeros(v.n_circ, "<u8"), np.zeros(v.n_circ)
        v.node_value = abi.ptr(node, abi.f64p)
        v.edge_key, v.edge_value = abi.ptr(ek, abi.u64p), abi.ptr(ev, abi.f64p)
        v.circ_key, v.circ_value = abi.ptr(ck, abi.u64p), abi.ptr(cv, abi.f64p)
        lib.check(self.L.wq_assign_ambig(self.h, C.byref(v)))
        return node, ek, ev, ck, cv

    def assign_ambig_inplace(self):
        """assign_ambig! without copying the stores out (the event stage reads them inside the library)."""
        v = abi.ValuesC()
        lib.check(self.L.wq_assign_ambig(self.h, C.byref(v)))

    def process_events_raw(self, readlen: int):
        """Same as process_events but returns the flat arrays of wq_events_out:
        (events[EVENT_DT], values f64, path_ptr u64, path_nodes u32)."""
        o = abi.EventsOutC()
        lib.check(self.L.wq_process_events(self.h, int(readlen), C.byref(o)))
        ev = np.zeros(o.n_events, abi.EVENT_DT)
        vals = np.zeros(max(1, o.n_values))
        pptr = np.zeros(o.n_paths + 1, "<u8")
        pnodes = np.zeros(max(1, o.n_path_nodes), "<u4")
        o.events, o.values = ev.ctypes.data, abi.ptr(vals, abi.f64p)
        o.path_ptr, o.path_nodes = abi.ptr(pptr, abi.u64p), abi.ptr(pnodes, abi.u32p)
        lib.check(self.L.wq_process_events(self.h, int(readlen), C.byref(o)))
        return ev, vals, pptr, pnodes

    def em_psi_batch(self, problems, readlen: int, sig: int = 4, maxit: int = 1500):
        """wq_em_psi_batch: spliced_em! / rec_tandem_em! (src/events.jl:853-932) for a list of problems
        (n_inc, count[np], length[np], amb_sets (lists of 0-based path indices), amb_mult, tandem).
        Returns [(psi[np], iterations)]."""
        E = len(problems)
        path_ptr, amb_ptr, amb_path_ptr = [0], [0], [0]
        n_inc, cnt, ln, apaths, amult, tan = [], [], [], [], [], []
        for (ni, c, l, sets, mult, t) in problems:
            n_inc.append(ni); cnt += list(c); ln += list(l); tan.append(1 if t else 0)
            path_ptr.append(len(cnt))
            for a, m in zip(sets, mult):
                apaths += list(a); amult.append(m); amb_path_ptr.append(len(apaths))
            amb_ptr.append(len(amult))
        arr = lambda x, dt: np.ascontiguousarray(x if len(x) else [0], dt)
        a_pp, a_ni, a_ap, a_app = arr(path_ptr, "<u4"), arr(n_inc, "<u4"), arr(amb_ptr, "<u4"), arr(amb_path_ptr, "<u4")
        a_paths, a_cnt, a_ln, a_mult, a_tan = arr(apaths, "<u4"), arr(cnt, "<f8"), arr(ln, "<f8"), arr(amult, "<f8"), arr(tan, "u1")
        psi, its = np.zeros(max(1, len(cnt))), np.zeros(max(1, E), "<i4")
        b = abi.PsiBatchC()
        b.n_events = E
        b.path_ptr, b.n_inc, b.path_count, b.path_length = abi.ptr(a_pp, abi.u32p), abi.ptr(a_ni, abi.u32p), abi.ptr(a_cnt, abi.f64p), abi.ptr(a_ln, abi.f64p)
        b.amb_ptr, b.amb_path_ptr, b.amb_paths, b.amb_mult = abi.ptr(a_ap, abi.u32p), abi.ptr(a_app, abi.u32p), abi.ptr(a_paths, abi.u32p), abi.ptr(a_mult, abi.f64p)
        b.is_tandem = abi.ptr(a_tan, abi.u8p)
        b.readlen, b.sig, b.maxit = int(readlen), int(sig), int(maxit)
        b.psi, b.path_total, b.iterations = abi.ptr(psi, abi.f64p), None, abi.ptr(its, abi.i32p)
        lib.check(self.L.wq_em_psi_batch(self.h, C.byref(b)))
        return [(psi[path_ptr[e]:path_ptr[e + 1]].copy(), int(its[e])) for e in range(E)]

    def process_events_view(self, readlen: int):
        """process_events without copies: numpy views of the arrays the library owns (wq_process_events_view), valid
        until the count state changes or the handle is closed."""
        o = abi.EventsOutC()
        lib.check(self.L.wq_process_events_view(self.h, int(readlen), C.byref(o)))
        n = int(o.n_events)
        if n == 0:
            return np.zeros(0, abi.EVENT_DT), np.zeros(0), np.zeros(1, "<u8"), np.zeros(0, "<u4")
        ev = np.frombuffer((C.c_char * (n * abi.EVENT_DT.itemsize)).from_address(o.events), dtype=abi.EVENT_DT)
        vals = np.ctypeslib.as_array(o.values, shape=(int(o.n_values),)) if o.n_values else np.zeros(0)
        pptr = np.ctypeslib.as_array(o.path_ptr, shape=(int(o.n_paths) + 1,))
        pnodes = np.ctypeslib.as_array(o.path_nodes, shape=(int(o.n_path_nodes),)) if o.n_path_nodes else np.zeros(0, "<u4")
        return ev, vals, pptr, pnodes

    def process_events(self, readlen: int):
        """process_events(out, lib, quant; isnodeok=false, readlen) (src/events.jl:744-800) without the writers:
        returns a list of dicts, one per event record in gene/node order (same shape as the oracle's)."""
        o = abi.EventsOutC()
        lib.check(self.L.wq_process_events(self.h, int(readlen), C.byref(o)))
        ev = np.zeros(o.n_events, abi.EVENT_DT)
        vals = np.zeros(max(1, o.n_values))
        pptr = np.zeros(o.n_paths + 1, "<u8")
        pnodes = np.zeros(max(1, o.n_path_nodes), "<u4")
        o.events, o.values = ev.ctypes.data, abi.ptr(vals, abi.f64p)
        o.path_ptr, o.path_nodes = abi.ptr(pptr, abi.u64p), abi.ptr(pnodes, abi.u32p)
        lib.check(self.L.wq_process_events(self.h, int(readlen), C.byref(o)))
        out = []
        for e in ev:
            d = dict(gene=int(e["gene"]), node=int(e["node"]), motif=int(e["motif"]), kind=int(e["kind"]), flags=int(e["flags"]),
                     psi=float(e["psi"]), total=float(e["total"]), iterations=int(e["iterations"]), utr_psi=[], inc=[], exc=[],
                     ci=(float(e["ci_lo"]), float(e["ci_hi"])))
            vs, ps = int(e["value_start"]), int(e["path_start"])
            path = lambda k: tuple(int(x) for x in pnodes[int(pptr[ps + k]):int(pptr[ps + k + 1])])
            if d["kind"] == 2:
                d["utr_psi"] = vals[vs:vs + int(e["n_utr"])].tolist()
            else:
                ni, nx = int(e["n_inc"]), int(e["n_exc"])
                have = d["kind"] == 1 and ni and nx
                d["inc"] = [(path(k), float(vals[vs + k]) if have else 0.0) for k in range(ni)]
                d["exc"] = [(path(ni + k), float(vals[vs + ni + k]) if have else 0.0) for k in range(nx)]
            out.append(d)
        return out

    # ---- multi-GPU exchange ---------------------------------------------------------------------
    def comm_init(self, dist=None, nranks=None, rank=None, unique_id: bytes | None = None):
        """NCCL communicator of this handle inside the library (wq_comm_unique_id / wq_comm_init_rank).  With a
        torch.distributed module the 128-byte id is made on rank 0 and broadcast over that process group; a host without
        torch passes (nranks, rank, unique_id) obtained by any other means."""
        if dist is not None:
            nranks, rank = dist.get_world_size(), dist.get_rank()
            box = [None]
            if rank == 0:
                buf = C.create_string_buffer(128)
                lib.check(self.L.wq_comm_unique_id(buf))
                box[0] = buf.raw
            dist.broadcast_object_list(box, src=0)
            unique_id = box[0]
        lib.check(self.L.wq_comm_init_rank(self.h, int(nranks), int(rank), C.c_char_p(unique_id)))
        self._lib_comm = True

    def counts_sync(self):
        """wq_counts_sync: ONE ncclAllReduce of the dense counts + ONE ncclAllGather of the compacted sparse stores, merged on
        the device, one stream synchronisation."""
        lib.check(self.L.wq_counts_sync(self.h))

    def classes_sync(self):
        """wq_classes_sync: this rank's share of the classes, all-gathered over the library's communicator, assembled."""
        lib.check(self.L.wq_classes_sync(self.h))

    def dense_tensor(self):
        """torch view of the device-resident dense vector (node counts + total/mapped/multi/sum_readlen)."""
        import torch
        dptr, nd = C.c_void_p(), C.c_uint64()
        lib.check(self.L.wq_counts_device_ptr(self.h, C.byref(dptr), C.byref(nd)))

        class _Wrap:
            __cuda_array_interface__ = {"shape": (int(nd.value),), "typestr": "<i8", "data": (int(dptr.value), False), "version": 2}
        return torch.as_tensor(_Wrap(), device=torch.device("cuda", self.device))

    def allreduce_counts(self):
        """Combine the count stores of all ranks (torch.distributed must be initialised): one NCCL all-reduce of
        the dense vector on the device, one all-gather of the packed sparse stores merged on every rank."""
        import torch
        import torch.distributed as dist
        from . import parallel
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        world, rank = dist.get_world_size(), dist.get_rank()
        dev = torch.device("cuda", self.device)
        dist.all_reduce(self.dense_tensor())
        # sparse stores: compacted on the device, all-gathered over NCCL, merged into the device tables by a kernel
        dptr, nbytes = C.c_void_p(), C.c_uint64()
        lib.check(self.L.wq_sparse_device_pack(self.h, C.byref(dptr), C.byref(nbytes)))
        nw = int(nbytes.value) // 8

        class _Wrap:
            __cuda_array_interface__ = {"shape": (nw,), "typestr": "<i8", "data": (int(dptr.value), False), "version": 2}
        mine = torch.as_tensor(_Wrap(), device=dev)
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([nw], dtype=torch.int64, device=dev))
        sizes = [int(x) for x in sizes.tolist()]
        m = max(sizes)
        padded = torch.zeros(m, dtype=torch.int64, device=dev)
        padded[:nw] = mine
        gathered = torch.empty(world * m, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, padded)
        torch.cuda.current_stream().synchronize()
        for r in range(world):
            if r != rank:
                lib.check(self.L.wq_sparse_device_merge(self.h, C.c_void_p(gathered[r * m:].data_ptr()), C.c_uint64(sizes[r] * 8)))

    def classes_share(self) -> np.ndarray:
        """This part's share of the classes (wq_classes_build_local + wq_classes_pack) as a byte array (a copy)."""
        lib.check(self.L.wq_classes_build_local(self.h))
        ptr_, n = C.c_void_p(), C.c_uint64()
        lib.check(self.L.wq_classes_pack(self.h, C.byref(ptr_), C.byref(n)))
        return np.frombuffer((C.c_char * int(n.value)).from_address(ptr_.value), dtype="u1").copy()

    def classes_assemble(self, shares):
        """wq_classes_assemble: the shares of all parts, in part order."""
        shares = [np.ascontiguousarray(s, "u1") for s in shares]
        bufs = (C.c_void_p * len(shares))(*[s.ctypes.data for s in shares])
        lens = (C.c_uint64 * len(shares))(*[s.nbytes for s in shares])
        lib.check(self.L.wq_classes_assemble(self.h, bufs, lens, len(shares)))

    def build_classes_distributed(self):
        """Class building partitioned by gene across ranks: inside the library when it has a communicator (comm_init),
        else over torch.distributed (after allreduce_counts and set_gene_partition(rank, world)): local share -> NCCL
        all-gather -> complete class set on every rank."""
        if getattr(self, "_lib_comm", False):
            return self.classes_sync()
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        dev = torch.device("cuda", self.device)
        mine = self.classes_share()
        nb = (len(mine) + 7) // 8 * 8
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([len(mine)], dtype=torch.int64, device=dev))
        sizes = [int(x) for x in sizes.tolist()]
        m = (max(sizes) + 7) // 8 * 8
        padded = torch.zeros(m, dtype=torch.uint8, device=dev)
        padded[:len(mine)] = torch.from_numpy(mine).to(dev)
        gathered = torch.empty(world * m, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered, padded)
        host = gathered.cpu().numpy()
        self.classes_assemble([host[r * m:r * m + sizes[r]] for r in range(world)])

    def set_gene_partition(self, part: int, nparts: int):
        """Event stage partitioned by gene: this handle builds and solves the events of gene range `part` of `nparts`;
        process_events then returns that range only (the parts concatenate to the full output)."""
        lib.check(self.L.wq_set_gene_partition(self.h, int(part), int(nparts)))

    def sparse_pack(self) -> np.ndarray:
        used = C.c_uint64()
        lib.check(self.L.wq_sparse_pack(self.h, None, 0, C.byref(used)))
        buf = np.zeros(used.value // 8, "<u8")
        lib.check(self.L.wq_sparse_pack(self.h, buf.ctypes.data_as(C.c_void_p), buf.nbytes, C.byref(used)))
        return buf

    def sparse_merge(self, buf: np.ndarray):
        buf = np.ascontiguousarray(buf, "<u8")
        lib.check(self.L.wq_sparse_merge(self.h, buf.ctypes.data_as(C.c_void_p), buf.nbytes))

    def merge_sparse(self, ek, ec, ck, cc, longs, multis):
        c = abi.CountsC()
        ek, ec, ck, cc = (np.ascontiguousarray(a, "<u8") for a in (ek, ec, ck, cc))
        c.n_nodes, c.node_count = 0, None
        c.n_edge, c.edge_key, c.edge_count = len(ek), abi.ptr(ek, abi.u64p), abi.ptr(ec, abi.u64p)
        c.n_circ, c.circ_key, c.circ_count = len(ck), abi.ptr(ck, abi.u64p), abi.ptr(cc, abi.u64p)
        ks, keep = [], []
        for d in (longs, multis):
            k = abi.KeyedC()
            keys = sorted(d)
            ptr_ = np.zeros(len(keys) + 1, "<u8")
            words = np.zeros(max(1, sum(len(x) for x in keys)), "<u4")
            cnt = np.zeros(max(1, len(keys)), "<u8")
            w = 0
            for i, key in enumerate(keys):
                ptr_[i] = w
                words[w:w + len(key)] = key
                w += len(key)
                cnt[i] = d[key]
            ptr_[len(keys)] = w
            k.n_rec, k.n_words = len(keys), w
            k.rec_ptr, k.words, k.count = abi.ptr(ptr_, abi.u64p), abi.ptr(words, abi.u32p), abi.ptr(cnt, abi.u64p)
            ks.append(k)
            keep.append((ptr_, words, cnt))
        lib.check(self.L.wq_counts_merge(self.h, C.byref(c), C.byref(ks[0]), C.byref(ks[1])))
